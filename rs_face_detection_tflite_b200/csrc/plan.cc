// plan.cc -- see plan.h.
#include "plan.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>

namespace fdl {

namespace {

constexpr int64_t kAlign = 64;  // floats (256 B)
inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct Group {
  int kind = -1;
  std::vector<int> ops;
  int last_op = -1;
  int main_op = -1;   // CONV / DW / POOL / PAD / ADD / ACT / RESIZE op index
  int dw_op = -1;     // BLOCK: the depthwise op
  int act_op = -1;
  int add_op = -1;
  int in_tensor = -1, out_tensor = -1;
  int skip_tensor = -1, skip_pool = 0, skip_c = 0;
  int pad_op = -1;            // spatial PAD folded into the (VALID) depthwise / convolution: explicit padding
  int pad_t = 0, pad_b = 0, pad_l = 0, pad_r = 0;
  int fused_relu = 0;         // fused_activation_function = RELU on the group's last CONV_2D / ADD
};

struct Builder {
  const TfModel& m;
  std::string* err;
  std::vector<int> producer;                 // tensor -> op index (-1: none)
  std::vector<std::vector<int>> consumers;   // tensor -> op indices
  std::vector<char> is_output;
  std::vector<char> is_const;
  std::vector<int> group_of;                 // op -> group id, -2 = folded away (dequantize / alias)
  std::vector<Group> groups;
  // alias: tensor -> (root tensor, offset within root item)
  std::vector<int> root;
  std::vector<int64_t> root_off;

  Builder(const TfModel& mm, std::string* e) : m(mm), err(e) {}

  bool fail(const std::string& s) { *err = s; return false; }

  int n_consumers(int t) const { return (int)consumers[t].size() + (is_output[t] ? 1 : 0); }
  const std::vector<int>& shape(int t) const { return m.tensors[t].shape; }

  bool nhwc(int t, int* H, int* W, int* C) const {
    const auto& s = shape(t);
    if (s.size() == 4 && s[0] == 1) { *H = s[1]; *W = s[2]; *C = s[3]; return true; }
    if (s.size() == 3 && s[0] == 1) { *H = 1; *W = s[1]; *C = s[2]; return true; }
    if (s.size() == 2 && s[0] == 1) { *H = 1; *W = 1; *C = s[1]; return true; }
    return false;
  }
  int64_t elems(int t) const { return m.tensors[t].elems(); }

  bool is_act(int op) const { return m.ops[op].code == OP_RELU || m.ops[op].code == OP_PRELU; }

  // sole consumer op of tensor t (and t is not a graph output), else -1
  int sole_consumer(int t) const {
    if (is_output[t] || consumers[t].size() != 1) return -1;
    return consumers[t][0];
  }

  bool index() {
    size_t nt = m.tensors.size();
    producer.assign(nt, -1);
    consumers.assign(nt, {});
    is_output.assign(nt, 0);
    is_const.assign(nt, 0);
    for (size_t t = 0; t < nt; ++t) is_const[t] = m.tensors[t].data != nullptr;
    for (int o : m.outputs) is_output[o] = 1;
    for (size_t i = 0; i < m.ops.size(); ++i) {
      const TfOp& op = m.ops[i];
      if (op.outputs.size() != 1) return fail("op with != 1 output is unsupported");
      // TFLite marks an omitted optional input with -1; none of the supported graphs has one, and the planner indexes by input
      for (int t : op.inputs) if (t < 0) return fail("operator with an omitted optional input (-1) is unsupported");
      if (producer[op.outputs[0]] != -1) return fail("tensor produced twice");
      producer[op.outputs[0]] = (int)i;
      for (int t : op.inputs) if (t >= 0 && !is_const[t]) consumers[t].push_back((int)i);
    }
    // DEQUANTIZE outputs of constants are constants themselves (folded at load)
    for (size_t i = 0; i < m.ops.size(); ++i) {
      const TfOp& op = m.ops[i];
      if (op.code == OP_DEQUANTIZE) {
        if (op.inputs.empty() || !is_const[op.inputs[0]]) return fail("DEQUANTIZE of a non-constant tensor is unsupported");
        is_const[op.outputs[0]] = 2;  // 2 = via dequantize
      } else if (op.code == OP_DENSIFY) {
        if (op.inputs.empty() || !is_const[op.inputs[0]]) return fail("DENSIFY of a non-constant tensor is unsupported");
        if (!m.tensors[op.inputs[0]].has_sparsity || !m.tensors[op.inputs[0]].sparse_ok)
          return fail("DENSIFY: only row-major CSR on the last dimension is supported");
        is_const[op.outputs[0]] = 2;  // via densify (ops are in topological order: a DEQUANTIZE of it comes later)
      }
    }
    // rebuild consumers without the now-constant tensors
    for (auto& c : consumers) c.clear();
    for (size_t i = 0; i < m.ops.size(); ++i)
      for (int t : m.ops[i].inputs) if (t >= 0 && !is_const[t]) consumers[t].push_back((int)i);
    return true;
  }

  // constant tensor -> f32 values, looking through a DEQUANTIZE
  bool const_values(int t, std::vector<float>* out) const {
    while (is_const[t] == 2) t = m.ops[producer[t]].inputs[0];   // through DEQUANTIZE and DENSIFY
    return m.const_f32(t, out);
  }
  // PAD with paddings [[0,0],[t,b],[l,r],[0,0]]
  bool check_spatial_pad(const TfOp& op, int* pt, int* pb, int* pl, int* pr) const {
    std::vector<int> p;
    if (op.inputs.size() < 2 || !m.const_i32(op.inputs[1], &p) || p.size() != 8) return false;
    if (p[0] || p[1] || p[6] || p[7]) return false;
    *pt = p[2]; *pb = p[3]; *pl = p[4]; *pr = p[5];
    return p[2] >= 0 && p[3] >= 0 && p[4] >= 0 && p[5] >= 0 && (p[2] + p[3] + p[4] + p[5]) > 0;
  }
  // Is op `i` a spatial PAD whose only consumer is a VALID depthwise / convolution?  (Then that consumer's group takes it.)
  bool pad_feeds_valid_conv(int i) const {
    const TfOp& op = m.ops[i];
    int pt, pb, pl, pr;
    if (op.code != OP_PAD || !check_spatial_pad(op, &pt, &pb, &pl, &pr)) return false;
    const int c = sole_consumer(op.outputs[0]);
    if (c < 0) return false;
    const TfOp& cv = m.ops[c];
    return (cv.code == OP_DEPTHWISE_CONV_2D || cv.code == OP_CONV_2D) && cv.padding == 1 && cv.inputs[0] == op.outputs[0];
  }
  // If the (VALID) conv op `c` reads a foldable spatial PAD, fold it: the group reads the PAD's input with explicit padding.
  void fold_input_pad(int c, Group* g, const std::function<void(int)>& take) {
    const int in = m.ops[c].inputs[0];
    const int p = producer[in];
    if (p < 0 || group_of[p] != -1 || !pad_feeds_valid_conv(p)) return;
    check_spatial_pad(m.ops[p], &g->pad_t, &g->pad_b, &g->pad_l, &g->pad_r);
    g->pad_op = p;
    g->in_tensor = m.ops[p].inputs[0];
    take(p);
  }

  bool check_channel_pad(const TfOp& op, int* extra) const {
    std::vector<int> p;
    if (op.inputs.size() < 2 || !m.const_i32(op.inputs[1], &p) || p.size() != 8) return false;
    for (int i = 0; i < 7; ++i) if (p[i] != 0) return false;
    *extra = p[7];
    return p[7] >= 0;
  }
  bool check_pool2(const TfOp& op) const {
    int H, W, C, OH, OW, OC;
    if (!nhwc(op.inputs[0], &H, &W, &C) || !nhwc(op.outputs[0], &OH, &OW, &OC)) return false;
    return op.filter_w == 2 && op.filter_h == 2 && op.stride_w == 2 && op.stride_h == 2 && op.fused_act == 0 &&
           H % 2 == 0 && W % 2 == 0 && OH == H / 2 && OW == W / 2 && OC == C;
  }

  bool aliases() {
    size_t nt = m.tensors.size();
    root.resize(nt);
    root_off.assign(nt, 0);
    for (size_t t = 0; t < nt; ++t) root[t] = (int)t;
    group_of.assign(m.ops.size(), -1);
    for (int i = (int)m.ops.size() - 1; i >= 0; --i) {
      const TfOp& op = m.ops[i];
      if (op.code == OP_DEQUANTIZE || op.code == OP_DENSIFY) { group_of[i] = -2; continue; }
      if (op.code == OP_RESHAPE) {
        int in = op.inputs[0], out = op.outputs[0];
        if (is_const[in]) return fail("RESHAPE of a constant is unsupported");
        if (n_consumers(in) != 1) return fail("RESHAPE input with several consumers is unsupported");
        if (elems(in) != elems(out)) return fail("RESHAPE changes the element count");
        root[in] = root[out];
        root_off[in] = root_off[out];
        group_of[i] = -2;
      } else if (op.code == OP_CONCATENATION) {
        int out = op.outputs[0];
        const auto& so = shape(out);
        if (op.fused_act != 0 || so.size() != 3 || so[0] != 1 || op.axis != 1)
          return fail("only CONCATENATION(axis=1) of [1,n,c] tensors is supported");
        int64_t off = 0;
        for (int in : op.inputs) {
          const auto& si = shape(in);
          if (is_const[in] || si.size() != 3 || si[0] != 1 || si[2] != so[2] || n_consumers(in) != 1)
            return fail("unsupported CONCATENATION operand");
          root[in] = root[out];
          root_off[in] = root_off[out] + off;
          off += (int64_t)si[1] * si[2];
        }
        if (off != elems(out)) return fail("CONCATENATION size mismatch");
        group_of[i] = -2;
      }
    }
    return true;
  }

  bool make_groups() {
    for (int i = 0; i < (int)m.ops.size(); ++i) {
      if (group_of[i] != -1) continue;
      const TfOp& op = m.ops[i];
      if (pad_feeds_valid_conv(i)) continue;       // taken by the group of the conv that consumes it
      Group g;
      std::function<void(int)> take = [&](int o) { g.ops.push_back(o); group_of[o] = (int)groups.size(); g.last_op = std::max(g.last_op, o); };
      auto take_act = [&](int t) {  // fuse a trailing RELU/PRELU consuming tensor t; returns the new tail tensor
        int c = sole_consumer(t);
        if (c >= 0 && group_of[c] == -1 && is_act(c)) { g.act_op = c; take(c); return m.ops[c].outputs[0]; }
        return t;
      };
      switch (op.code) {
        case OP_DEPTHWISE_CONV_2D: {
          if (op.depth_multiplier != 1 || op.dil_w != 1 || op.dil_h != 1 || op.fused_act != 0 || op.stride_w != op.stride_h)
            return fail("unsupported DEPTHWISE_CONV_2D options");
          const auto& ws = shape(op.inputs[1]);
          if (ws.size() != 4 || ws[1] != 3 || ws[2] != 3) return fail("only 3x3 depthwise kernels are supported");
          g.in_tensor = op.inputs[0];
          g.dw_op = i;
          take(i);
          fold_input_pad(i, &g, take);
          int d = op.outputs[0];
          int c = sole_consumer(d);
          bool pw = false;
          if (c >= 0 && group_of[c] == -1 && m.ops[c].code == OP_CONV_2D) {
            const TfOp& cv = m.ops[c];
            const auto& cs = shape(cv.inputs[1]);
            pw = cs.size() == 4 && cs[1] == 1 && cs[2] == 1 && cv.stride_w == 1 && cv.stride_h == 1 && (cv.fused_act == 0 || cv.fused_act == 1) &&
                 cv.inputs[0] == d;
          }
          if (!pw) {
            g.kind = STEP_DW; g.main_op = i;
            g.out_tensor = take_act(d);
            break;
          }
          g.kind = STEP_BLOCK; g.main_op = c;
          take(c);
          int t = m.ops[c].outputs[0];
          if (m.ops[c].fused_act == 1) { g.fused_relu = 1; g.out_tensor = t; break; }   // RELU before anything else: no residual to fuse
          int a = sole_consumer(t);
          if (a >= 0 && group_of[a] == -1 && m.ops[a].code == OP_ADD && (m.ops[a].fused_act == 0 || m.ops[a].fused_act == 1)) {
            const TfOp& add = m.ops[a];
            int other = add.inputs[0] == t ? add.inputs[1] : add.inputs[0];
            if (other != t && !is_const[other]) {
              // walk the skip branch back through PAD / MAX_POOL
              int s = other, pool = 0, pad_op = -1, pool_op = -1;
              int p = producer[s];
              if (p >= 0 && group_of[p] == -1 && m.ops[p].code == OP_PAD && n_consumers(s) == 1) {
                int extra;
                if (check_channel_pad(m.ops[p], &extra)) { pad_op = p; s = m.ops[p].inputs[0]; p = producer[s]; }
              }
              if (p >= 0 && group_of[p] == -1 && m.ops[p].code == OP_MAX_POOL_2D && n_consumers(s) == 1 && check_pool2(m.ops[p])) {
                pool_op = p; pool = 1; s = m.ops[p].inputs[0];
              }
              int H, W, C, OH, OW, OC;
              if (nhwc(s, &H, &W, &C) && nhwc(t, &OH, &OW, &OC) && C <= OC &&
                  ((pool && H == 2 * OH && W == 2 * OW) || (!pool && H == OH && W == OW)) && !is_const[s] &&
                  (pad_op >= 0 || C == OC)) {
                g.skip_tensor = s; g.skip_pool = pool; g.skip_c = C; g.add_op = a;
                if (pad_op >= 0) take(pad_op);
                if (pool_op >= 0) take(pool_op);
                take(a);
                t = add.outputs[0];
                if (add.fused_act == 1) { g.fused_relu = 1; g.out_tensor = t; break; }
              }
            }
          }
          g.out_tensor = take_act(t);
          break;
        }
        case OP_CONV_2D: {
          if (op.dil_w != 1 || op.dil_h != 1 || (op.fused_act != 0 && op.fused_act != 1) || op.stride_w != op.stride_h)
            return fail("unsupported CONV_2D options");
          g.kind = STEP_CONV; g.main_op = i; g.in_tensor = op.inputs[0];
          take(i);
          fold_input_pad(i, &g, take);
          if (op.fused_act == 1) { g.fused_relu = 1; g.out_tensor = op.outputs[0]; break; }
          g.out_tensor = take_act(op.outputs[0]);
          break;
        }
        case OP_MAX_POOL_2D: {
          if (!check_pool2(op)) return fail("only 2x2 stride-2 MAX_POOL_2D on even sizes is supported");
          g.kind = STEP_POOL; g.main_op = i; g.in_tensor = op.inputs[0];
          take(i);
          g.out_tensor = op.outputs[0];
          break;
        }
        case OP_PAD: {
          int extra;
          if (!check_channel_pad(op, &extra)) return fail("only trailing channel PAD is supported");
          g.kind = STEP_PADC; g.main_op = i; g.in_tensor = op.inputs[0];
          take(i);
          g.out_tensor = op.outputs[0];
          break;
        }
        case OP_ADD: {
          if ((op.fused_act != 0 && op.fused_act != 1) || is_const[op.inputs[0]] || is_const[op.inputs[1]] || elems(op.inputs[0]) != elems(op.inputs[1]))
            return fail("unsupported ADD");
          g.kind = STEP_ADD; g.main_op = i; g.in_tensor = op.inputs[0]; g.skip_tensor = op.inputs[1];
          take(i);
          if (op.fused_act == 1) { g.fused_relu = 1; g.out_tensor = op.outputs[0]; break; }
          g.out_tensor = take_act(op.outputs[0]);
          break;
        }
        case OP_RELU: case OP_PRELU: {
          g.kind = STEP_ACT; g.main_op = i; g.act_op = i; g.in_tensor = op.inputs[0];
          take(i);
          g.out_tensor = op.outputs[0];
          break;
        }
        case OP_RESIZE_BILINEAR: {
          if (op.align_corners || !op.half_pixel_centers) return fail("RESIZE_BILINEAR must use half_pixel_centers");
          g.kind = STEP_RESIZE; g.main_op = i; g.in_tensor = op.inputs[0];
          take(i);
          int t = op.outputs[0];
          int a = sole_consumer(t);
          if (a >= 0 && group_of[a] == -1 && m.ops[a].code == OP_ADD && (m.ops[a].fused_act == 0 || m.ops[a].fused_act == 1)) {
            const TfOp& add = m.ops[a];
            int other = add.inputs[0] == t ? add.inputs[1] : add.inputs[0];
            if (other != t && !is_const[other] && elems(other) == elems(t)) {
              g.skip_tensor = other; g.add_op = a;
              take(a);
              t = add.outputs[0];
              if (add.fused_act == 1) { g.fused_relu = 1; g.out_tensor = t; break; }
            }
          }
          g.out_tensor = take_act(t);
          break;
        }
        case OP_DEPTH_TO_SPACE: {
          int H, W, C, OH, OW, OC;
          const int bs = op.block_size;
          if (bs < 1 || !nhwc(op.inputs[0], &H, &W, &C) || !nhwc(op.outputs[0], &OH, &OW, &OC) || OH != H * bs || OW != W * bs || OC * bs * bs != C)
            return fail("unsupported DEPTH_TO_SPACE");
          g.kind = STEP_D2S; g.main_op = i; g.in_tensor = op.inputs[0];
          take(i);
          g.out_tensor = op.outputs[0];
          break;
        }
        default:
          return fail(std::string("unsupported builtin operator ") + op_name(op.code) + " (code " + std::to_string(op.code) + ")");
      }
      groups.push_back(std::move(g));
    }
    for (size_t i = 0; i < m.ops.size(); ++i)
      if (group_of[i] == -1) return fail(std::string("internal: ") + op_name(m.ops[i].code) + " was left out of every fused step");
    return true;
  }
};

struct Interval { int64_t off, size; };

}  // namespace

bool Plan::build(const TfModel& m, std::string* err) {
  Builder b(m, err);
  if (m.inputs.size() != 1) { *err = "graph must have exactly one input"; return false; }
  if (!b.index() || !b.aliases() || !b.make_groups()) return false;
  num_tflite_ops = (int)m.ops.size();

  // emission order: by the position of each group's last op (op order is topological)
  std::vector<int> order(b.groups.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return b.groups[x].last_op < b.groups[y].last_op; });

  // ---- weights -------------------------------------------------------------------------------
  auto push = [&](const std::vector<float>& v) {
    int64_t off = align_up((int64_t)weights.size(), kAlign);
    weights.resize((size_t)off, 0.f);
    weights.insert(weights.end(), v.begin(), v.end());
    return off;
  };

  steps.clear();
  for (int gi : order) {
    const Group& g = b.groups[gi];
    Step s;
    s.kind = g.kind;
    s.ops = g.ops;
    std::sort(s.ops.begin(), s.ops.end());
    int H = 0, W = 0, C = 0;
    if (!b.nhwc(g.in_tensor, &H, &W, &C)) { *err = "unsupported tensor rank"; return false; }
    s.in.tensor = g.in_tensor; s.in.H = H; s.in.W = W; s.in.C = C;
    if (!b.nhwc(g.out_tensor, &H, &W, &C)) { *err = "unsupported tensor rank"; return false; }
    s.out.tensor = g.out_tensor; s.out.H = H; s.out.W = W; s.out.C = C;
    if (g.skip_tensor >= 0) {
      if (!b.nhwc(g.skip_tensor, &H, &W, &C)) { *err = "unsupported tensor rank"; return false; }
      s.skip.tensor = g.skip_tensor; s.skip.H = H; s.skip.W = W; s.skip.C = C;
      s.skip_pool = g.skip_pool; s.skip_c = g.skip_c;
    }
    auto same_pad = [](int in, int k, int stride, int out, int* before) {
      int total = std::max(0, (out - 1) * stride + k - in);
      *before = total / 2;
    };
    if (g.fused_relu) s.act = ACT_RELU;
    if (g.act_op >= 0) {
      const TfOp& a = m.ops[g.act_op];
      if (a.code == OP_RELU) s.act = ACT_RELU;
      else {
        s.act = ACT_PRELU;
        std::vector<float> al;
        if (!b.const_values(a.inputs[1], &al) || (int)al.size() != s.out.C) { *err = "PRELU alpha must be a [1,1,C] constant"; return false; }
        s.alpha = push(al);
      }
    }
    if (g.kind == STEP_BLOCK || g.kind == STEP_DW) {
      const TfOp& dw = m.ops[g.dw_op];
      std::vector<float> w, bias;
      if (!b.const_values(dw.inputs[1], &w) || !b.const_values(dw.inputs[2], &bias) || (int)w.size() != 9 * s.in.C ||
          (int)bias.size() != s.in.C) { *err = "bad depthwise weights"; return false; }
      s.kh = s.kw = 3; s.stride = dw.stride_w;
      int dh = 0, dwid = 0, dc = 0;
      b.nhwc(dw.outputs[0], &dh, &dwid, &dc);
      if (dw.padding == 0) { same_pad(s.in.H, 3, s.stride, dh, &s.pad_t); same_pad(s.in.W, 3, s.stride, dwid, &s.pad_l); }
      else { s.pad_t = g.pad_t; s.pad_l = g.pad_l; }   // VALID: no padding, or the folded spatial PAD's
      // output-size sanity (SAME: ceil(in/stride); VALID: (in + pads - k)/stride + 1)
      int eh = dw.padding == 0 ? (s.in.H + s.stride - 1) / s.stride : (s.in.H + g.pad_t + g.pad_b - 3) / s.stride + 1;
      int ew = dw.padding == 0 ? (s.in.W + s.stride - 1) / s.stride : (s.in.W + g.pad_l + g.pad_r - 3) / s.stride + 1;
      if (eh != dh || ew != dwid || dc != s.in.C) { *err = "depthwise output shape mismatch"; return false; }
      s.w_dw = push(w);   // tflite layout [1,3,3,C] == [9][C]
      s.b_dw = push(bias);
      flops_per_item += 2LL * 9 * dc * dh * dwid;
    }
    if (g.kind == STEP_BLOCK || g.kind == STEP_CONV) {
      const TfOp& cv = m.ops[g.main_op];
      const auto& ws = b.shape(cv.inputs[1]);
      std::vector<float> w, bias;
      if (ws.size() != 4 || !b.const_values(cv.inputs[1], &w) || !b.const_values(cv.inputs[2], &bias)) { *err = "bad conv weights"; return false; }
      int co = ws[0], kh = ws[1], kw = ws[2], ci = ws[3];
      int ih = 0, iw = 0, ic = 0, oh = 0, ow = 0, oc = 0;
      b.nhwc(g.kind == STEP_CONV ? g.in_tensor : cv.inputs[0], &ih, &iw, &ic);   // STEP_CONV: the folded PAD's input, if any
      b.nhwc(cv.outputs[0], &oh, &ow, &oc);
      if (ic != ci || oc != co || (int)bias.size() != co) { *err = "conv shape mismatch"; return false; }
      if (g.kind == STEP_CONV) {
        s.kh = kh; s.kw = kw; s.stride = cv.stride_w;
        if (cv.padding == 0) { same_pad(ih, kh, s.stride, oh, &s.pad_t); same_pad(iw, kw, s.stride, ow, &s.pad_l); }
        else { s.pad_t = g.pad_t; s.pad_l = g.pad_l; }
        int eh = cv.padding == 0 ? (ih + s.stride - 1) / s.stride : (ih + g.pad_t + g.pad_b - kh) / s.stride + 1;
        int ew = cv.padding == 0 ? (iw + s.stride - 1) / s.stride : (iw + g.pad_l + g.pad_r - kw) / s.stride + 1;
        if (eh != oh || ew != ow) { *err = "conv output shape mismatch"; return false; }
      }
      s.K = kh * kw * ci; s.K4 = (int)align_up(s.K, 4); s.N = co; s.Npad = (int)align_up(co, 4);
      std::vector<float> wt((size_t)s.K4 * s.Npad, 0.f), bp((size_t)s.Npad, 0.f);  // rows K..K4 stay zero
      for (int o = 0; o < co; ++o) {
        bp[o] = bias[o];
        for (int k = 0; k < s.K; ++k) wt[(size_t)k * s.Npad + o] = w[(size_t)o * s.K + k];  // OHWI -> [(ky,kx,ci)][co]
      }
      s.w = push(wt);
      s.b = push(bp);
      flops_per_item += 2LL * s.K * co * oh * ow;
      if (ci % 4 == 0 || g.kind == STEP_CONV) {
        auto hi_of = [](float v) { uint32_t u; std::memcpy(&u, &v, 4); u &= 0xffffe000u; float r; std::memcpy(&r, &u, 4); return r; };
        bool exact = true;
        for (float v : w) if (hi_of(v) != v) { exact = false; break; }
        const int ws = exact ? 1 : 2;
        s.Kp = (int)align_up(s.K, 32);
        s.Nt = std::min((int)align_up(co, 16), 128);
        s.n_tiles = (co + s.Nt - 1) / s.Nt;
        const size_t plane = (size_t)(s.Kp / 4) * s.Nt * 4;       // floats of one (hi or lo) copy of one N tile
        std::vector<float> pk((size_t)s.n_tiles * ws * plane, 0.f);
        for (int o = 0; o < co; ++o) {
          const int t = o / s.Nt, nn = o - t * s.Nt;
          for (int k = 0; k < s.K; ++k) {
            float v = w[(size_t)o * s.K + k], h = hi_of(v);
            size_t at = (size_t)t * ws * plane + ((size_t)(k / 4) * s.Nt + nn) * 4 + (k % 4);
            pk[at] = h;
            if (ws == 2) pk[at + plane] = v - h;
          }
        }
        s.w_tc = push(pk);
        s.wsplit = ws;
      }
      if (g.kind == STEP_BLOCK && ci % 8 == 0) {
        // tensor-core packing: plane q (4 input channels) x output channel n x 4 floats
        s.Np = (int)align_up(co, 16);
        auto hi_of = [](float v) { uint32_t u; std::memcpy(&u, &v, 4); u &= 0xffffe000u; float r; std::memcpy(&r, &u, 4); return r; };
        bool exact = true;
        for (float v : w) if (hi_of(v) != v) { exact = false; break; }
        s.wsplit = exact ? 1 : 2;
        const int Q = ci / 4;
        std::vector<float> pk((size_t)s.wsplit * Q * s.Np * 4, 0.f);
        for (int o = 0; o < co; ++o)
          for (int k = 0; k < ci; ++k) {
            float v = w[(size_t)o * ci + k], h = hi_of(v);
            size_t at = ((size_t)(k / 4) * s.Np + o) * 4 + (k % 4);
            pk[at] = h;
            if (s.wsplit == 2) pk[(size_t)Q * s.Np * 4 + at] = v - h;
          }
        s.w_umma = push(pk);
        // f16-split packing (block_ws_kernel, kind::f16)
        bool exact16 = true;
        for (float v : w) if (half_to_float(float_to_half(v)) != v) { exact16 = false; break; }
        s.wsplit16 = exact16 ? 1 : 2;
        std::vector<uint16_t> hk((size_t)s.wsplit16 * Q * s.Np * 8, 0);
        for (int o = 0; o < co; ++o)
          for (int k = 0; k < ci; ++k) {
            const float v = w[(size_t)o * ci + k];
            const uint16_t h = float_to_half(v);
            const size_t at = ((size_t)(k / 4) * s.Np + o) * 8 + (k % 4);
            hk[at] = h; hk[at + 4] = h;                       // against the hi half and against the lo half of the activation
            if (s.wsplit16 == 2) hk[(size_t)Q * s.Np * 8 + at] = float_to_half(v - half_to_float(h));   // (w_lo, 0): only the hi half
          }
        std::vector<float> hf(hk.size() / 2);
        std::memcpy(hf.data(), hk.data(), hk.size() * 2);
        s.w_f16 = push(hf);
      }
    }
    if (g.kind == STEP_RESIZE) {
      if (s.out.C != s.in.C) { *err = "resize channel mismatch"; return false; }
    }
    if (g.kind == STEP_D2S) s.stride = m.ops[g.main_op].block_size;
    steps.push_back(std::move(s));
  }

  // ---- buffers ---------------------------------------------------------------------------------
  // root tensor -> [first writer step, last reader step]
  size_t nt = m.tensors.size();
  std::vector<int> first_def(nt, INT32_MAX), last_use(nt, -1);
  int in_t = m.inputs[0];
  first_def[b.root[in_t]] = -1;
  for (size_t si = 0; si < steps.size(); ++si) {
    const Step& s = steps[si];
    int ro = b.root[s.out.tensor];
    first_def[ro] = std::min(first_def[ro], (int)si);
    last_use[ro] = std::max(last_use[ro], (int)si);
    int ri = b.root[s.in.tensor];
    last_use[ri] = std::max(last_use[ri], (int)si);
    if (s.skip.tensor >= 0) { int rs = b.root[s.skip.tensor]; last_use[rs] = std::max(last_use[rs], (int)si); }
  }
  for (int o : m.outputs) last_use[b.root[o]] = INT32_MAX;
  last_use[b.root[in_t]] = std::max(last_use[b.root[in_t]], 0);

  // ---- branches: which graph outputs does each step feed? ----------------------------------------
  // (FaceMesh: landmarks / face flag from the 6x6 map on; iris: eye contour / iris from the 8x8 map on; detectors: the
  // regressor / classificator heads.)  Steps are in topological order, so one reverse sweep propagates the output sets.
  num_streams = 1;
  if (m.outputs.size() >= 2 && m.outputs.size() <= 32) {
    std::vector<uint32_t> reach(nt, 0u);
    for (size_t k = 0; k < m.outputs.size(); ++k) reach[b.root[m.outputs[k]]] |= 1u << k;
    for (size_t si = steps.size(); si-- > 0;) {
      Step& s = steps[si];
      const uint32_t mask = reach[b.root[s.out.tensor]];
      reach[b.root[s.in.tensor]] |= mask;
      if (s.skip.tensor >= 0) reach[b.root[s.skip.tensor]] |= mask;
      s.stream = 0;
      if (mask && (mask & (mask - 1)) == 0) {          // exactly one output
        int k = 0;
        while (!((mask >> k) & 1u)) ++k;
        s.stream = k;
      }
      num_streams = std::max(num_streams, s.stream + 1);
    }
    // buffers touched by an auxiliary-stream step are never recycled: no write-after-read hazards between streams
    for (const Step& s : steps) {
      if (s.stream == 0) continue;
      last_use[b.root[s.out.tensor]] = INT32_MAX;
      last_use[b.root[s.in.tensor]] = INT32_MAX;
      if (s.skip.tensor >= 0) last_use[b.root[s.skip.tensor]] = INT32_MAX;
    }
  }

  std::vector<int64_t> buf_off(nt, -1);
  std::vector<std::pair<int, Interval>> live;  // (root tensor, interval)
  int64_t high = 0;
  auto alloc = [&](int rt) {
    if (buf_off[rt] >= 0) return;
    int64_t size = align_up(m.tensors[rt].elems(), kAlign);   // (skewing consecutive buffers against each other was measured: no effect)
    std::vector<Interval> iv;
    for (auto& l : live) iv.push_back(l.second);
    std::sort(iv.begin(), iv.end(), [](const Interval& a, const Interval& c) { return a.off < c.off; });
    int64_t pos = 0;
    for (auto& i : iv) {
      if (pos + size <= i.off) break;
      pos = std::max(pos, i.off + i.size);
    }
    buf_off[rt] = pos;
    live.push_back({rt, {pos, size}});
    high = std::max(high, pos + size);
  };
  alloc(b.root[in_t]);
  for (size_t si = 0; si < steps.size(); ++si) {
    const Step& s = steps[si];
    for (int t : {s.in.tensor, s.skip.tensor}) {
      if (t >= 0 && buf_off[b.root[t]] < 0) { *err = "internal: step reads an unallocated tensor"; return false; }
    }
    alloc(b.root[s.out.tensor]);
    for (size_t k = 0; k < live.size();) {
      if (last_use[live[k].first] <= (int)si) live.erase(live.begin() + k); else ++k;
    }
  }
  arena_per_item = high;

  auto fill = [&](TensorRef& r) {
    if (r.tensor < 0) return;
    int rt = b.root[r.tensor];
    r.buf_offset = buf_off[rt];
    r.batch_stride = m.tensors[rt].elems();
    r.offset = b.root_off[r.tensor];
  };
  algo_bytes_per_item = 0;
  for (auto& s : steps) {
    fill(s.in); fill(s.out); fill(s.skip);
    int64_t bytes = (int64_t)s.in.H * s.in.W * s.in.C + (int64_t)s.out.H * s.out.W * s.out.C;
    if (s.skip.tensor >= 0 && s.skip.tensor != s.in.tensor) bytes += (int64_t)s.skip.H * s.skip.W * s.skip.C;
    algo_bytes_per_item += bytes * 4;
  }
  {
    int H = 0, W = 0, C = 0;
    b.nhwc(in_t, &H, &W, &C);
    input.tensor = in_t; input.H = H; input.W = W; input.C = C;
    fill(input);
    outputs.clear();
    for (int o : m.outputs) {
      TensorRef r;
      if (!b.nhwc(o, &H, &W, &C)) { *err = "unsupported output rank"; return false; }
      r.tensor = o; r.H = H; r.W = W; r.C = C;
      fill(r);
      outputs.push_back(r);
    }
  }

  // ---- text ------------------------------------------------------------------------------------
  static const char* kActName[] = {"none", "relu", "prelu"};
  for (size_t si = 0; si < steps.size(); ++si) {
    Step& s = steps[si];
    char buf[512];
    std::string ops;
    for (int o : s.ops) { ops += (ops.empty() ? "" : ","); ops += std::to_string(o); }
    switch (s.kind) {
      case STEP_BLOCK:
        std::snprintf(buf, sizeof buf, "#%zu BLOCK dw3x3/s%d+pw %d->%d in %dx%d out %dx%d skip=%s%s(t%d,c%d) act=%s ops=[%s]", si, s.stride,
                      s.in.C, s.out.C, s.in.H, s.in.W, s.out.H, s.out.W, s.skip.tensor < 0 ? "none" : (s.skip_pool ? "maxpool" : "direct"),
                      (s.skip.tensor >= 0 && s.skip_c < s.out.C) ? "+chanpad" : "", s.skip.tensor, s.skip_c, kActName[s.act], ops.c_str());
        break;
      case STEP_CONV:
        std::snprintf(buf, sizeof buf, "#%zu CONV %dx%d/s%d %d->%d in %dx%d out %dx%d pad(t%d,l%d) act=%s ops=[%s]", si, s.kh, s.kw, s.stride,
                      s.in.C, s.out.C, s.in.H, s.in.W, s.out.H, s.out.W, s.pad_t, s.pad_l, kActName[s.act], ops.c_str());
        break;
      case STEP_DW:
        std::snprintf(buf, sizeof buf, "#%zu DW 3x3/s%d c%d in %dx%d out %dx%d act=%s ops=[%s]", si, s.stride, s.in.C, s.in.H, s.in.W,
                      s.out.H, s.out.W, kActName[s.act], ops.c_str());
        break;
      case STEP_POOL:
        std::snprintf(buf, sizeof buf, "#%zu MAXPOOL 2x2/s2 c%d in %dx%d ops=[%s]", si, s.in.C, s.in.H, s.in.W, ops.c_str());
        break;
      case STEP_PADC:
        std::snprintf(buf, sizeof buf, "#%zu CHANPAD %d->%d @%dx%d ops=[%s]", si, s.in.C, s.out.C, s.in.H, s.in.W, ops.c_str());
        break;
      case STEP_ADD:
        std::snprintf(buf, sizeof buf, "#%zu ADD c%d @%dx%d act=%s ops=[%s]", si, s.out.C, s.out.H, s.out.W, kActName[s.act], ops.c_str());
        break;
      case STEP_ACT:
        std::snprintf(buf, sizeof buf, "#%zu ACT %s c%d @%dx%d ops=[%s]", si, kActName[s.act], s.out.C, s.out.H, s.out.W, ops.c_str());
        break;
      case STEP_D2S:
        std::snprintf(buf, sizeof buf, "#%zu DEPTH_TO_SPACE x%d %dx%dx%d->%dx%dx%d ops=[%s]", si, s.stride, s.in.H, s.in.W, s.in.C, s.out.H, s.out.W,
                      s.out.C, ops.c_str());
        break;
      case STEP_RESIZE:
        std::snprintf(buf, sizeof buf, "#%zu RESIZE_BILINEAR %dx%d->%dx%d c%d add=%s act=%s ops=[%s]", si, s.in.H, s.in.W, s.out.H, s.out.W,
                      s.out.C, s.skip.tensor >= 0 ? "yes" : "no", kActName[s.act], ops.c_str());
        break;
      default: buf[0] = 0;
    }
    s.text = buf;
    if (s.stream > 0) s.text += " stream=" + std::to_string(s.stream);
  }
  build_chain(*this);
  return true;
}

std::string Plan::describe() const {
  std::string out;
  char buf[256];
  std::snprintf(buf, sizeof buf, "plan: %d tflite ops -> %zu launches; arena %lld floats/item; weights %zu floats; "
                "block-fused floor %lld bytes/item; %lld flop/item\n", num_tflite_ops, steps.size(), (long long)arena_per_item,
                weights.size(), (long long)algo_bytes_per_item, (long long)flops_per_item);
  out += buf;
  for (const auto& s : steps) { out += s.text; out += "\n"; }
  for (const ChainPlan& c : chains) out += c.text;
  return out;
}

}  // namespace fdl
