// tflite_model.cc -- see tflite_model.h.
#include "tflite_model.h"

#include <cstdio>
#include <cstring>

namespace fdl {

const char* op_name(int code) {
  switch (code) {
    case OP_ADD: return "ADD";
    case OP_CONCATENATION: return "CONCATENATION";
    case OP_CONV_2D: return "CONV_2D";
    case OP_DEPTHWISE_CONV_2D: return "DEPTHWISE_CONV_2D";
    case OP_DEPTH_TO_SPACE: return "DEPTH_TO_SPACE";
    case OP_DEQUANTIZE: return "DEQUANTIZE";
    case OP_MAX_POOL_2D: return "MAX_POOL_2D";
    case OP_RELU: return "RELU";
    case OP_RESHAPE: return "RESHAPE";
    case OP_RESIZE_BILINEAR: return "RESIZE_BILINEAR";
    case OP_PAD: return "PAD";
    case OP_PRELU: return "PRELU";
    case OP_DENSIFY: return "DENSIFY";
    default: return "UNKNOWN";
  }
}

float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1f;
  uint32_t man = h & 0x3ffu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else {  // subnormal: renormalise
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400u));
      man &= 0x3ffu;
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | (man << 13);
  } else {
    bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}

uint16_t float_to_half(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
  const uint32_t ax = x & 0x7fffffffu;
  if (ax > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);                 // NaN
  if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7bffu);                // >= 65520 rounds past the largest half: saturate
  if (ax < 0x33000001u) return sign;                                       // <= 2^-25: rounds to zero
  int e = (int)(ax >> 23) - 127;
  uint32_t man = (ax & 0x7fffffu) | 0x800000u;                             // 24-bit significand
  int shift = e >= -14 ? 13 : 13 + (-14 - e);                              // bits dropped (subnormal halves drop more)
  uint32_t q = man >> shift, rem = man & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (q & 1u))) ++q;
  uint32_t h = e >= -14 ? (((uint32_t)(e + 15) << 10) + (q - 0x400u)) : q;  // a carry out of the significand bumps the exponent
  return (uint16_t)(sign | h);
}

namespace {

// Cursor over the file image; every access is bounds-checked and failure is sticky.
struct FB {
  const uint8_t* b;
  size_t n;
  bool ok = true;

  template <typename T> T rd(size_t p) {
    if (p + sizeof(T) > n || p + sizeof(T) < p) { ok = false; return T(0); }
    T v; std::memcpy(&v, b + p, sizeof(T)); return v;
  }
  // absolute position of field `fid` in `table`, 0 if absent
  size_t field(size_t table, int fid) {
    int32_t soff = rd<int32_t>(table);
    size_t vt = (size_t)((int64_t)table - soff);
    uint16_t vsize = rd<uint16_t>(vt);
    size_t slot = 4 + 2 * (size_t)fid;
    if (!ok || slot + 2 > vsize) return 0;
    uint16_t off = rd<uint16_t>(vt + slot);
    return off ? table + off : 0;
  }
  template <typename T> T scalar(size_t table, int fid, T dflt) {
    size_t p = field(table, fid);
    return p ? rd<T>(p) : dflt;
  }
  size_t indirect(size_t p) { return p + rd<uint32_t>(p); }
  size_t table(size_t t, int fid) { size_t p = field(t, fid); return p ? indirect(p) : 0; }
  // vector field -> (start of elements, count)
  bool vec(size_t t, int fid, size_t* start, uint32_t* count) {
    size_t p = field(t, fid);
    if (!p) { *start = 0; *count = 0; return false; }
    size_t v = indirect(p);
    *count = rd<uint32_t>(v);
    *start = v + 4;
    return ok;
  }
  std::vector<int> vec_i32(size_t t, int fid) {
    size_t s; uint32_t c; std::vector<int> out;
    if (!vec(t, fid, &s, &c)) return out;
    if (s + 4ull * c > n) { ok = false; return out; }
    out.resize(c);
    for (uint32_t i = 0; i < c; ++i) out[i] = rd<int32_t>(s + 4ull * i);
    return out;
  }
  std::vector<size_t> vec_tables(size_t t, int fid) {
    size_t s; uint32_t c; std::vector<size_t> out;
    if (!vec(t, fid, &s, &c)) return out;
    if (s + 4ull * c > n) { ok = false; return out; }
    out.resize(c);
    for (uint32_t i = 0; i < c; ++i) out[i] = indirect(s + 4ull * i);
    return out;
  }
  std::string str(size_t t, int fid) {
    size_t s; uint32_t c;
    if (!vec(t, fid, &s, &c) || s + c > n) return std::string();
    return std::string((const char*)b + s, c);
  }
};

}  // namespace

bool TfModel::load(const std::string& path, std::string* err) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { *err = "cannot open model file: " + path; return false; }
  std::fseek(f, 0, SEEK_END);
  long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (sz < 16) { std::fclose(f); *err = "model file too small: " + path; return false; }
  file.resize((size_t)sz);
  size_t got = std::fread(file.data(), 1, (size_t)sz, f);
  std::fclose(f);
  if (got != (size_t)sz) { *err = "short read: " + path; return false; }
  if (std::memcmp(file.data() + 4, "TFL3", 4) != 0) { *err = "not a TFLite (TFL3) flatbuffer: " + path; return false; }

  FB fb{file.data(), file.size()};
  size_t root = fb.indirect(0);
  version = fb.scalar<uint32_t>(root, 0, 0);

  std::vector<int> codes;
  for (size_t oc : fb.vec_tables(root, 1)) {
    int dep = fb.scalar<int8_t>(oc, 0, 0);
    int neu = fb.scalar<int32_t>(oc, 3, 0);
    codes.push_back(dep > neu ? dep : neu);
  }
  struct Buf { size_t start; uint32_t len; };
  std::vector<Buf> bufs;
  for (size_t bt : fb.vec_tables(root, 4)) {
    size_t s; uint32_t c;
    fb.vec(bt, 0, &s, &c);
    if (s && s + c > file.size()) fb.ok = false;
    bufs.push_back({s, c});
  }
  auto subs = fb.vec_tables(root, 2);
  if (!fb.ok || subs.size() != 1) { *err = "malformed model or != 1 subgraph: " + path; return false; }
  size_t sub = subs[0];

  for (size_t t : fb.vec_tables(sub, 0)) {
    TfTensor tt;
    tt.shape = fb.vec_i32(t, 0);
    tt.type = fb.scalar<int8_t>(t, 1, 0);
    tt.buffer = fb.scalar<uint32_t>(t, 2, 0);
    tt.name = fb.str(t, 3);
    tt.has_sparsity = fb.field(t, 6) != 0;
    // sizes that come from the file reach std::vector and the arena planner: every dimension must be positive and the element
    // count bounded (the largest tensor of the shipped graphs has 3.1 M elements)
    {
      int64_t total = 1;
      for (int d : tt.shape) {
        if (d <= 0 || d > (1 << 24)) { *err = "malformed model: tensor '" + tt.name + "' has a non-positive or oversized dimension: " + path; return false; }
        total *= d;
        if (total > (int64_t)1 << 31) { *err = "malformed model: tensor '" + tt.name + "' is too large: " + path; return false; }
      }
      if (tt.shape.size() > 8) { *err = "malformed model: tensor rank > 8: " + path; return false; }
    }
    if (tt.has_sparsity) {
      // SparsityParameters{0: traversal_order, 1: block_map, 2: dim_metadata[]};
      // DimensionMetadata{0: format (0 DENSE, 1 SPARSE_CSR), 1: dense_size, 2/3: array_segments (union), 4/5: array_indices (union)};
      // SparseIndexVector union: 1 Int32Vector, 2 Uint16Vector, 3 Uint8Vector, each {0: values}.
      const size_t sp = fb.table(t, 6);
      const std::vector<int> order = sp ? fb.vec_i32(sp, 0) : std::vector<int>();
      const std::vector<int> bmap = sp ? fb.vec_i32(sp, 1) : std::vector<int>();
      const std::vector<size_t> dims = sp ? fb.vec_tables(sp, 2) : std::vector<size_t>();
      const size_t rank = tt.shape.size();
      bool good = sp && rank >= 1 && order.size() == rank && bmap.empty() && dims.size() == rank;
      for (size_t d = 0; good && d < rank; ++d) good = order[d] == (int)d;
      for (size_t d = 0; good && d + 1 < rank; ++d)
        good = fb.scalar<int8_t>(dims[d], 0, 0) == 0 && fb.scalar<int32_t>(dims[d], 1, 0) == tt.shape[d];
      auto index_vector = [&](size_t dm, int type_fid, int value_fid, std::vector<int32_t>* out) {
        const int kind = fb.scalar<uint8_t>(dm, type_fid, 0);
        const size_t tb = fb.table(dm, value_fid);
        size_t s; uint32_t c;
        if (kind < 1 || kind > 3 || !tb || !fb.vec(tb, 0, &s, &c)) return false;
        const size_t width = kind == 1 ? 4 : (kind == 2 ? 2 : 1);
        if (s + (size_t)c * width > file.size()) return false;
        out->resize(c);
        for (uint32_t i = 0; i < c; ++i) {
          if (kind == 1) { int32_t v; std::memcpy(&v, file.data() + s + 4ull * i, 4); (*out)[i] = v; }
          else if (kind == 2) { uint16_t v; std::memcpy(&v, file.data() + s + 2ull * i, 2); (*out)[i] = v; }
          else (*out)[i] = file[s + i];
        }
        return true;
      };
      if (good) {
        const size_t last = dims[rank - 1];
        good = fb.scalar<int8_t>(last, 0, 0) == 1 && index_vector(last, 2, 3, &tt.sp_segments) && index_vector(last, 4, 5, &tt.sp_indices);
      }
      if (good) {
        int64_t rows = 1;
        for (size_t d = 0; d + 1 < rank; ++d) rows *= tt.shape[d];
        good = rows >= 1 && !tt.sp_segments.empty() && (int64_t)tt.sp_segments.size() == rows + 1 && tt.sp_segments.front() == 0 && tt.sp_segments.back() == (int32_t)tt.sp_indices.size();
        for (size_t i = 0; good && i + 1 < tt.sp_segments.size(); ++i) good = tt.sp_segments[i] <= tt.sp_segments[i + 1];
        for (size_t i = 0; good && i < tt.sp_indices.size(); ++i) good = tt.sp_indices[i] >= 0 && tt.sp_indices[i] < tt.shape[rank - 1];
      }
      tt.sparse_ok = good;
    }
    if (tt.buffer < bufs.size() && bufs[tt.buffer].start && bufs[tt.buffer].len) {
      tt.data = file.data() + bufs[tt.buffer].start;
      tt.nbytes = bufs[tt.buffer].len;
    }
    tensors.push_back(std::move(tt));
  }
  for (size_t o : fb.vec_tables(sub, 3)) {
    TfOp op;
    uint32_t ci = fb.scalar<uint32_t>(o, 0, 0);
    if (ci >= codes.size()) { fb.ok = false; break; }
    op.code = codes[ci];
    op.inputs = fb.vec_i32(o, 1);
    op.outputs = fb.vec_i32(o, 2);
    size_t t = fb.table(o, 4);
    if (t) {
      switch (op.code) {
        case OP_CONV_2D:
          op.padding = fb.scalar<int8_t>(t, 0, 0);
          op.stride_w = fb.scalar<int32_t>(t, 1, 0);
          op.stride_h = fb.scalar<int32_t>(t, 2, 0);
          op.fused_act = fb.scalar<int8_t>(t, 3, 0);
          op.dil_w = fb.scalar<int32_t>(t, 4, 1);
          op.dil_h = fb.scalar<int32_t>(t, 5, 1);
          break;
        case OP_DEPTHWISE_CONV_2D:
          op.padding = fb.scalar<int8_t>(t, 0, 0);
          op.stride_w = fb.scalar<int32_t>(t, 1, 0);
          op.stride_h = fb.scalar<int32_t>(t, 2, 0);
          op.depth_multiplier = fb.scalar<int32_t>(t, 3, 0);
          op.fused_act = fb.scalar<int8_t>(t, 4, 0);
          op.dil_w = fb.scalar<int32_t>(t, 5, 1);
          op.dil_h = fb.scalar<int32_t>(t, 6, 1);
          break;
        case OP_MAX_POOL_2D:
          op.padding = fb.scalar<int8_t>(t, 0, 0);
          op.stride_w = fb.scalar<int32_t>(t, 1, 0);
          op.stride_h = fb.scalar<int32_t>(t, 2, 0);
          op.filter_w = fb.scalar<int32_t>(t, 3, 0);
          op.filter_h = fb.scalar<int32_t>(t, 4, 0);
          op.fused_act = fb.scalar<int8_t>(t, 5, 0);
          break;
        case OP_ADD:
          op.fused_act = fb.scalar<int8_t>(t, 0, 0);
          break;
        case OP_CONCATENATION:
          op.axis = fb.scalar<int32_t>(t, 0, 0);
          op.fused_act = fb.scalar<int8_t>(t, 1, 0);
          break;
        case OP_RESHAPE:
          op.new_shape = fb.vec_i32(t, 0);
          break;
        case OP_DEPTH_TO_SPACE:
          op.block_size = fb.scalar<int32_t>(t, 0, 0);
          break;
        case OP_RESIZE_BILINEAR:
          op.align_corners = fb.scalar<uint8_t>(t, 2, 0) != 0;
          op.half_pixel_centers = fb.scalar<uint8_t>(t, 3, 0) != 0;
          break;
        default: break;
      }
    }
    for (int i : op.inputs) if (i < -1 || i >= (int)tensors.size()) fb.ok = false;
    for (int i : op.outputs) if (i < 0 || i >= (int)tensors.size()) fb.ok = false;
    ops.push_back(std::move(op));
  }
  inputs = fb.vec_i32(sub, 1);
  outputs = fb.vec_i32(sub, 2);
  for (int i : inputs) if (i < 0 || i >= (int)tensors.size()) fb.ok = false;
  for (int i : outputs) if (i < 0 || i >= (int)tensors.size()) fb.ok = false;
  if (!fb.ok) { *err = "malformed flatbuffer (offset out of range): " + path; return false; }
  return true;
}

bool TfModel::const_f32(int t, std::vector<float>* out) const {
  if (t < 0 || t >= (int)tensors.size()) return false;
  const TfTensor& tt = tensors[t];
  if (!tt.data) return false;
  int64_t n = tt.elems();
  if (tt.has_sparsity) {
    // DENSIFY folded at load: scatter the stored values of every CSR row into a zero tensor
    if (!tt.sparse_ok || (tt.type != TT_F32 && tt.type != TT_F16)) return false;
    const size_t nnz = tt.sp_indices.size(), width = tt.type == TT_F32 ? 4 : 2;
    if (tt.nbytes < nnz * width) return false;
    out->assign((size_t)n, 0.f);
    const int64_t cols = tt.shape.back();
    for (size_t r = 0; r + 1 < tt.sp_segments.size(); ++r)
      for (int32_t k = tt.sp_segments[r]; k < tt.sp_segments[r + 1]; ++k) {
        float v;
        if (tt.type == TT_F32) std::memcpy(&v, tt.data + 4ull * (size_t)k, 4);
        else { uint16_t h; std::memcpy(&h, tt.data + 2ull * (size_t)k, 2); v = half_to_float(h); }
        (*out)[(size_t)((int64_t)r * cols + tt.sp_indices[(size_t)k])] = v;
      }
    return true;
  }
  out->resize((size_t)n);
  if (tt.type == TT_F32) {
    if ((int64_t)tt.nbytes < n * 4) return false;
    std::memcpy(out->data(), tt.data, (size_t)n * 4);
  } else if (tt.type == TT_F16) {
    if ((int64_t)tt.nbytes < n * 2) return false;
    for (int64_t i = 0; i < n; ++i) {
      uint16_t h; std::memcpy(&h, tt.data + 2 * i, 2);
      (*out)[(size_t)i] = half_to_float(h);
    }
  } else {
    return false;
  }
  return true;
}

bool TfModel::const_i32(int t, std::vector<int>* out) const {
  if (t < 0 || t >= (int)tensors.size()) return false;
  const TfTensor& tt = tensors[t];
  if (!tt.data || tt.type != TT_I32) return false;
  int64_t n = tt.elems();
  if ((int64_t)tt.nbytes < n * 4) return false;
  out->resize((size_t)n);
  std::memcpy(out->data(), tt.data, (size_t)n * 4);
  return true;
}

}  // namespace fdl
