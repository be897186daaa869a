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
  if (!tt.data || tt.has_sparsity) return false;
  int64_t n = tt.elems();
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
