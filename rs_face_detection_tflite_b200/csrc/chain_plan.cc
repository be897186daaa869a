// chain_plan.cc -- see chain.h: finds the chainable tail of a planned graph and compiles it into the chain kernel's program.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>

#include "chain.h"
#include "plan.h"

namespace fdl {

namespace {

int align_up_i(int v, int a) { return (v + a - 1) / a * a; }

bool dense(const TensorRef& r) { return r.tensor >= 0 && r.offset == 0 && r.batch_stride == (int64_t)r.H * r.W * r.C; }

// What kind of chain step is this?  0: not chainable, 1: CONV 1x1, 2: CONV k x k / stride k (non-overlapping patches), 3: BLOCK
int classify(const Step& s) {
  if (!dense(s.in) || !dense(s.out)) return 0;
  const int N = s.out.C, C = s.in.C;
  if (N % 16 != 0 || N > 128 || C % 16 != 0 || C > 128) return 0;
  if (s.in.H * s.in.W > 64 || s.out.H * s.out.W > 64) return 0;
  if (s.kind == STEP_CONV) {
    if (s.skip.tensor >= 0 || s.pad_t != 0 || s.pad_l != 0 || s.kh != s.kw) return 0;
    if (s.kh == 1 && s.stride == 1) return 1;
    if (s.kh == s.stride && s.kh == 2 && s.in.H == 2 * s.out.H && s.in.W == 2 * s.out.W && s.K == 4 * C) return 2;
    return 0;
  }
  if (s.kind == STEP_BLOCK) {
    if (s.w_dw < 0 || s.K != C) return 0;
    if (s.stride == 1 && !(s.pad_t == 1 && s.pad_l == 1 && s.in.H == s.out.H && s.in.W == s.out.W)) return 0;
    if (s.stride == 2 && !(s.pad_t == 0 && s.pad_l == 0 && s.in.H == 2 * s.out.H && s.in.W == 2 * s.out.W)) return 0;
    if (s.stride != 1 && s.stride != 2) return 0;
    if (s.skip.tensor >= 0) {
      if (!dense(s.skip) || s.skip_c % 8 != 0 || s.skip_c > N || s.skip.C != s.skip_c) return 0;
      if (s.skip_pool && !(s.skip.H == 2 * s.out.H && s.skip.W == 2 * s.out.W)) return 0;
      if (!s.skip_pool && !(s.skip.H == s.out.H && s.skip.W == s.out.W)) return 0;
      if (s.skip.H * s.skip.W > 64) return 0;
    }
    return 3;
  }
  return 0;
}

uint16_t f2h(float f) { return float_to_half(f); }

struct Inst {            // one shared-memory tensor instance
  ChainTensor t;
  int rows = 0;          // rows allocated (items * H * W rounded up to 8)
  int size = 0;          // bytes
  int first = -1, last = -1;   // op indices
  int tensor = -1;       // tflite tensor (-1: temporary)
};

}  // namespace

namespace {

int step_rows(const Step& s) {
  int r = std::max(s.in.H * s.in.W, s.out.H * s.out.W);
  if (s.skip.tensor >= 0) r = std::max(r, s.skip.H * s.skip.W);
  return r;
}

// One chain over the plan steps [s0, s1] (all chainable).
bool build_one(Plan& plan, int s0, int s1, ChainPlan& ch) {
  ch = ChainPlan();
  const std::vector<Step>& steps = plan.steps;
  const int ns = (int)steps.size();
  int maxrows = 1;
  for (int i = s0; i <= s1; ++i) maxrows = std::max(maxrows, step_rows(steps[i]));
  const int G = 128 / maxrows;
  if (G < 1) return false;

  // ---- order: stream by stream (a stream-k step, k >= 1, feeds graph output k only: nothing in a lower stream waits for it) ----
  std::vector<int> order;
  for (int i = s0; i <= s1; ++i) order.push_back(i);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return steps[a].stream < steps[b].stream; });

  auto used_outside = [&](int tensor, int stream) {    // read by a step outside the chain, by another stream's run, or a graph output
    for (const TensorRef& o : plan.outputs) if (o.tensor == tensor) return true;
    for (int i = 0; i < ns; ++i) {
      const Step& s = steps[i];
      const bool reads = s.in.tensor == tensor || s.skip.tensor == tensor;
      if (!reads) continue;
      if (i < s0 || i > s1 || s.stream != stream) return true;
    }
    return false;
  };
  // a tensor may also be aliased into a graph output's buffer (heads): such steps are not dense() and never get here

  std::vector<ChainOp> ops;
  std::vector<Inst> inst;
  auto new_inst = [&](int fmt, int C, int H, int W, int tensor) {
    Inst in;
    in.rows = align_up_i(G * H * W, 8);
    in.t.fmt = (short)fmt; in.t.C = (short)C; in.t.H = (short)H; in.t.W = (short)W;
    if (fmt == CH_F32) { in.t.pl = C + 4; in.size = in.rows * (C + 4) * 4; }
    else { in.t.pl = in.rows * 16 + 16; in.size = 2 * (C / 8) * in.t.pl; }
    in.size = align_up_i(in.size, 128);
    in.tensor = tensor;
    inst.push_back(in);
    return (int)inst.size() - 1;
  };
  // op operands are instance indices until the allocation below: stored in ChainTensor::off
  auto ref = [&](int ii) { ChainTensor t = inst[ii].t; t.off = ii; return t; };
  auto none = [] { return ChainTensor(); };
  auto touch = [&](int op, int ii) {
    if (inst[ii].first < 0) inst[ii].first = op;
    inst[ii].last = std::max(inst[ii].last, op);
  };

  std::vector<float>& W = plan.weights;
  auto pad_weights = [&] { while (W.size() % 4) W.push_back(0.f); };

  int cur_stream = -1;
  std::map<std::pair<int, int>, int> live;             // (tensor, fmt) -> instance, within the current run
  auto emit_load = [&](const TensorRef& r, int fmt) {
    const int ii = new_inst(fmt, r.C, r.H, r.W, r.tensor);
    ChainOp o;
    o.kind = CH_LOAD; o.out = ref(ii); o.in = none(); o.out2 = none(); o.skip = none();
    o.g_buf_offset = r.buf_offset; o.g_bstride = r.batch_stride; o.g_offset = r.offset;
    ops.push_back(o);
    touch((int)ops.size() - 1, ii);
    live[{r.tensor, fmt}] = ii;
    return ii;
  };
  auto need = [&](const TensorRef& r, int fmt) {       // fmt < 0: any format
    if (fmt >= 0) {
      auto it = live.find({r.tensor, fmt});
      return it != live.end() ? it->second : emit_load(r, fmt);
    }
    for (int f : {CH_F32, CH_P16}) {
      auto it = live.find({r.tensor, f});
      if (it != live.end()) return it->second;
    }
    return emit_load(r, CH_F32);
  };

  for (int si : order) {
    const Step& s = steps[si];
    const int kind = classify(s);
    if (s.stream != cur_stream) { cur_stream = s.stream; live.clear(); }
    const int N = s.out.C, Np = N;
    // ---- parameter block ----
    ChainPar par;
    pad_weights();
    par.w_off = (long long)W.size();
    if (kind == 3) {
      par.dw_c = s.in.C;
      for (int i = 0; i < 9 * s.in.C; ++i) { const float v = W[s.w_dw + i]; W.push_back(v); }
      for (int i = 0; i < s.in.C; ++i) { const float v = W[s.b_dw + i]; W.push_back(v); }
    }
    par.np = Np;
    for (int n = 0; n < Np; ++n) { const float v = W[s.b + n]; W.push_back(v); }
    par.has_alpha = s.act == ACT_PRELU ? 1 : 0;
    if (par.has_alpha) for (int n = 0; n < Np; ++n) { const float v = W[s.alpha + n]; W.push_back(v); }
    par.bytes = (int)(((long long)W.size() - par.w_off) * 4);
    if (par.bytes > kChainParSlot || par.bytes % 16) return false;
    const int par_index = (int)ch.pars.size();
    ch.pars.push_back(par);
    ch.loads.push_back(ChainLoad{par.w_off, par.bytes, 1});

    // ---- the residual ----
    int skip_inst = -1;
    if (s.skip.tensor >= 0) {
      if (s.skip_pool) {
        const int src = need(s.skip, -1);
        skip_inst = new_inst(CH_F32, s.skip_c, s.out.H, s.out.W, -1);
        ChainOp o;
        o.kind = CH_POOL; o.in = ref(src); o.out = ref(skip_inst); o.out2 = none(); o.skip = none(); o.step = (short)si;
        ops.push_back(o);
        touch((int)ops.size() - 1, src); touch((int)ops.size() - 1, skip_inst);
      } else {
        skip_inst = need(s.skip, -1);
      }
    }
    // ---- the A operand ----
    int a_inst = -1;
    if (kind == 1) {
      a_inst = need(s.in, CH_P16);
    } else if (kind == 2) {
      const int src = need(s.in, CH_P16);
      a_inst = new_inst(CH_P16, s.K, s.out.H, s.out.W, -1);
      ChainOp o;
      o.kind = CH_GATHER; o.in = ref(src); o.out = ref(a_inst); o.out2 = none(); o.skip = none();
      o.k = (short)s.kh; o.stride = (short)s.stride; o.step = (short)si;
      ops.push_back(o);
      touch((int)ops.size() - 1, src); touch((int)ops.size() - 1, a_inst);
    } else {
      const int src = need(s.in, CH_F32);
      a_inst = new_inst(CH_P16, s.in.C, s.out.H, s.out.W, -1);
      ChainOp o;
      o.kind = CH_DW; o.in = ref(src); o.out = ref(a_inst); o.out2 = none(); o.skip = none();
      o.stride = (short)s.stride; o.pad_t = (short)s.pad_t; o.pad_l = (short)s.pad_l; o.par = (short)par_index; o.step = (short)si;
      o.par_dw_c = (short)s.in.C;
      if (!ops.empty() && ops.back().kind == CH_POOL && ops.back().step == (short)si) ops.back().no_barrier = 1;   // disjoint data: one phase
      ops.push_back(o);
      touch((int)ops.size() - 1, src); touch((int)ops.size() - 1, a_inst);
    }
    // ---- output formats: what the consumers inside this run need ----
    bool want_f32 = false, want_p16 = false;
    for (int sj : order) {
      const Step& c = steps[sj];
      if (c.stream != s.stream || sj <= si) continue;
      const int ck = classify(c);
      if (c.in.tensor == s.out.tensor) { if (ck == 3) want_f32 = true; else want_p16 = true; }
    }
    const bool store = used_outside(s.out.tensor, s.stream);
    if (!want_f32 && !want_p16) want_f32 = true;       // residual-only / store-only tensors (a residual is read from either format)
    const int o1 = new_inst(want_p16 ? CH_P16 : CH_F32, N, s.out.H, s.out.W, s.out.tensor);
    const int o2 = (want_p16 && want_f32) ? new_inst(CH_F32, N, s.out.H, s.out.W, s.out.tensor) : -1;
    // ---- the GEMM and its weight chunks ----
    ChainOp g;
    g.kind = CH_GEMM; g.in = ref(a_inst); g.out = ref(o1); g.out2 = o2 >= 0 ? ref(o2) : none();
    g.skip = skip_inst >= 0 ? ref(skip_inst) : none();
    g.has_skip = skip_inst >= 0 ? 1 : 0; g.skip_c = (short)s.skip_c;
    g.K = (short)s.K; g.N = (short)N; g.Np = (short)Np; g.act = (short)s.act; g.par = (short)par_index; g.par_release = 1; g.step = (short)si;
    g.par_dw_c = (short)(kind == 3 ? s.in.C : 0);
    g.chunk0 = (short)ch.chunks.size();
    if (s.K % 16) return false;
    const int kc_max = chain_kc_max(Np);
    for (int k0 = 0; k0 < s.K; k0 += kc_max) {
      const int kc = std::min(kc_max, s.K - k0);
      ChainChunk c;
      pad_weights();
      c.w_off = (long long)W.size(); c.k0 = k0; c.kc = kc; c.bytes = kc * Np * 4;
      std::vector<uint16_t> hk((size_t)2 * kc * Np, 0);
      for (int k = 0; k < kc; ++k)
        for (int n = 0; n < N; ++n) {
          const float v = plan.weights[s.w + (size_t)(k0 + k) * s.Npad + n];
          const uint16_t h = f2h(v);
          const size_t at = ((size_t)(k / 8) * Np + n) * 8 + (k % 8);
          hk[at] = h;
          hk[(size_t)kc * Np + at] = f2h(v - half_to_float(h));
        }
      const size_t base = W.size();
      W.resize(base + hk.size() / 2);
      std::memcpy(W.data() + base, hk.data(), hk.size() * 2);
      ch.loads.push_back(ChainLoad{c.w_off, c.bytes, 0});
      ch.chunks.push_back(c);
    }
    g.nchunks = (short)((int)ch.chunks.size() - g.chunk0);
    ops.push_back(g);
    {
      const int oi = (int)ops.size() - 1;
      touch(oi, a_inst); touch(oi, o1);
      if (o2 >= 0) touch(oi, o2);
      if (skip_inst >= 0) touch(oi, skip_inst);
    }
    live[{s.out.tensor, inst[o1].t.fmt}] = o1;
    if (o2 >= 0) live[{s.out.tensor, CH_F32}] = o2;
    if (store) {
      ChainOp o;
      o.kind = CH_STORE; o.in = ref(o2 >= 0 ? o2 : o1); o.out = none(); o.out2 = none(); o.skip = none(); o.step = (short)si;
      o.g_buf_offset = s.out.buf_offset; o.g_bstride = s.out.batch_stride; o.g_offset = s.out.offset;
      ops.push_back(o);
      touch((int)ops.size() - 1, o2 >= 0 ? o2 : o1);
    }
  }
  // liveness within a run ends with the run: an instance is never referenced by a later run (live is cleared), so `last` is exact.
  if ((int)ops.size() > kChainMaxOps || (int)ch.loads.size() > kChainMaxLoads) return false;

  // ---- shared-memory allocation by liveness ----
  std::vector<int> off(inst.size(), -1);
  struct Seg { int off, size, ii; };
  std::vector<Seg> segs;
  auto place = [&](int ii, int at_hint) {
    if (at_hint >= 0) { off[ii] = at_hint; segs.push_back({at_hint, inst[ii].size, ii}); return true; }
    std::sort(segs.begin(), segs.end(), [](const Seg& a, const Seg& b) { return a.off < b.off; });
    int pos = 0;
    for (const Seg& sg : segs) {
      if (pos + inst[ii].size <= sg.off) break;
      pos = std::max(pos, sg.off + sg.size);
    }
    if (pos + inst[ii].size > kChainArena) return false;
    off[ii] = pos;
    segs.push_back({pos, inst[ii].size, ii});
    return true;
  };
  auto release = [&](int ii) {
    for (size_t k = 0; k < segs.size(); ++k) if (segs[k].ii == ii) { segs.erase(segs.begin() + k); return; }
  };
  std::vector<int> deferred;
  for (int oi = 0; oi < (int)ops.size(); ++oi) {
    ChainOp& o = ops[oi];
    const int iin = o.in.off, iout = o.out.off, iout2 = o.out2.off, iskip = o.skip.off;
    int hint = -1;
    if (o.kind == CH_GEMM) {
      // the A operand is dead once the accumulator is complete: its space may hold the output
      if (iin >= 0 && inst[iin].last == oi) release(iin);
      // a dying residual of the same layout is overwritten in place (row-local read, then write, by the same thread)
      if (iskip >= 0 && inst[iskip].last == oi && off[iskip] >= 0 && inst[iskip].t.fmt == inst[iout].t.fmt && inst[iskip].t.C == inst[iout].t.C &&
          inst[iskip].rows == inst[iout].rows && iout2 < 0) {
        hint = off[iskip];
        release(iskip);
      }
    }
    if (iout >= 0 && off[iout] < 0 && !place(iout, hint)) return false;
    if (iout2 >= 0 && off[iout2] < 0 && !place(iout2, -1)) return false;
    // (an op that shares its phase with the next one keeps its dying operands until that op's outputs are placed)
    for (int ii : deferred) release(ii);
    deferred.clear();
    for (int ii : {iin, iskip, iout, iout2})
      if (ii >= 0 && inst[ii].last <= oi) { if (o.no_barrier) deferred.push_back(ii); else release(ii); }
  }
  auto fix = [&](ChainTensor& t) { if (t.off >= 0) t.off = off[t.off]; };
  for (ChainOp& o : ops) { fix(o.in); fix(o.out); fix(o.out2); fix(o.skip); }

  ch.ops = ops;
  ch.first_step = s0; ch.last_step = s1; ch.items = G;
  ch.valid = true;
  char buf[256];
  std::snprintf(buf, sizeof buf, "chain: steps #%d..#%d -> one launch, %d items per group, %zu ops, %zu weight chunks, %zu parameter blocks\n", s0, s1, G,
                ops.size(), ch.chunks.size(), ch.pars.size());
  ch.text = buf;
  static const char* kName[] = {"LOAD", "STORE", "POOL", "GATHER", "DW", "GEMM"};
  for (const ChainOp& o : ops) {
    std::snprintf(buf, sizeof buf, "  %-6s step #%d in@%d(%s c%d %dx%d) out@%d(%s c%d %dx%d) out2@%d skip@%d K=%d N=%d chunks=%d\n", kName[o.kind], o.step, o.in.off,
                  o.in.fmt ? "P16" : "F32", o.in.C, o.in.H, o.in.W, o.out.off, o.out.fmt ? "P16" : "F32", o.out.C, o.out.H, o.out.W, o.out2.off, o.skip.off,
                  o.K, o.N, o.nchunks);
    ch.text += buf;
  }
  return true;
}

}  // namespace

bool build_chain(Plan& plan) {
  plan.chains.clear();
  const std::vector<Step>& steps = plan.steps;
  const int ns = (int)steps.size();
  // ---- the longest run of chainable steps ----
  int best0 = 0, best1 = -1;
  for (int i = 0; i < ns;) {
    if (!classify(steps[i])) { ++i; continue; }
    int j = i;
    while (j + 1 < ns && classify(steps[j + 1])) ++j;
    if (j - i > best1 - best0) { best0 = i; best1 = j; }
    i = j + 1;
  }
  if (best1 - best0 + 1 < 6) return false;
  for (int i = 0; i < best0; ++i) if (steps[i].stream != 0) return false;
  // ---- two chains when the maps shrink along the run: an op costs the same ~3 us of latency whether its 128 rows are full or
  // not, and after two halvings a group of the first chain's size fills 8 .. 32 of them.  The run is cut after the last step that
  // touches a map of more than a quarter of the largest one; the second chain packs four times as many items into a group. ----
  int maxrows = 1;
  for (int i = best0; i <= best1; ++i) maxrows = std::max(maxrows, step_rows(steps[i]));
  int cut = -1;
  for (int i = best0; i <= best1; ++i) if (step_rows(steps[i]) * 4 > maxrows) cut = i;
  static const bool split_on = [] { const char* e = getenv("FDL_CHAIN_SPLIT"); return e ? atoi(e) != 0 : true; }();
  // (worth it for a long second part only -- iris: 20 steps, 384 -> 327 us; the landmark graph's 5 + 5 steps measured slower apart)
  if (split_on && cut >= best0 + 1 && best1 - cut >= 12) {
    std::vector<float> saved = plan.weights;
    ChainPlan a, b;
    if (build_one(plan, best0, cut, a) && build_one(plan, cut + 1, best1, b) && b.items >= 2 * a.items) {
      plan.chains.push_back(a);
      plan.chains.push_back(b);
      return true;
    }
    plan.weights = saved;
  }
  ChainPlan one;
  if (!build_one(plan, best0, best1, one)) return false;
  plan.chains.push_back(one);
  return true;
}

}  // namespace fdl