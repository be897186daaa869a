// chain.h -- the "tail chain": every step of a planned graph that works on maps of at most 8 x 8 pixels, run by ONE
// persistent kernel with the activations resident in shared memory (chain_kernel.cu).
//
// The FaceMesh graph from its 6 x 6 map on and the iris graph from its 8 x 8 map on are 10 and 32 launches of 17 .. 45 us
// each for a few hundred kilobytes of data per item: launch latency, not work.  The planner (chain_plan.cc) turns such a
// run of steps into a small program over shared-memory tensors; a CTA takes a group of items (as many as fit in the 128
// rows of one UMMA tile), walks the program, and only the chain's inputs and outputs touch global memory.
//
// Tensor formats in shared memory (rows = item-major pixels of the group, at most 128):
//   F32  row-major fp32, pixel stride C + 4 floats          -- depthwise input, residual source, store source
//   P16  "planes": for every 8 channels one hi plane and one lo plane of [row][8 halves] (x ~= hi + lo, 2^-22), plane stride
//        pl bytes -- exactly the K-major core-matrix layout tcgen05.mma kind::f16 reads, so a P16 tensor IS an A operand
// Program ops (each is one phase of the CTA, closed by a barrier):
//   LOAD / STORE   global fp32 NHWC <-> a tensor
//   POOL           MAX_POOL 2x2 of a tensor -> F32 (the residual of the stride-2 / bottleneck-downsampling blocks)
//   GATHER         im2col of a k x k / stride k convolution: P16 [HxW][C] -> P16 [H/k x W/k][k*k*C]
//   DW             depthwise 3x3 (stride 1 or 2) + bias: F32 -> P16
//   GEMM           [rows x K] (P16) x [K x N] on the tensor cores (3 passes: hi*hi, lo*hi, hi*lo; accumulator in TMEM),
//                  + bias, + residual, RELU / PRELU, written as F32 and / or P16
// Weights stream through a two-slot ring of 32 KB chunks (cp.async.bulk by a producer warp, one layer ahead); the small
// per-layer parameters (depthwise weights, biases, slopes) through a second ring.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

namespace fdl {

struct Plan;

enum ChainOpKind : int { CH_LOAD = 0, CH_STORE = 1, CH_POOL = 2, CH_GATHER = 3, CH_DW = 4, CH_GEMM = 5 };
enum ChainFmt : int { CH_F32 = 0, CH_P16 = 1 };

// A tensor in the shared-memory arena (off < 0: none).
struct ChainTensor {
  int off = -1;       // byte offset in the arena
  int pl = 0;         // P16: plane stride in bytes; F32: pixel stride in floats
  short C = 0;        // channels
  short fmt = CH_F32;
  short H = 0, W = 0; // per item
};

struct ChainOp {      // 128 bytes: the program travels as a kernel parameter (constant bank: uniform loads)
  ChainTensor in, out, out2, skip;   // out2: second copy of a GEMM's output in the other format
  // LOAD / STORE: the global tensor (floats): arena + buf_offset * B + item * bstride + offset
  long long g_buf_offset = 0, g_bstride = 0, g_offset = 0;
  short kind = 0;
  short stride = 1, pad_t = 0, pad_l = 0, k = 1;   // DW / GATHER geometry
  short act = 0, has_skip = 0;
  short par_release = 0;             // the parameter block is dead after this op
  short K = 0, N = 0, Np = 0, skip_c = 0;          // GEMM
  short chunk0 = 0, nchunks = 0;     // this GEMM's weight chunks: chunk j covers K range [j * kc_max, ...), kc_max = chain_kc_max(Np)
  short par = -1;                    // parameter block this op reads (the DW and the GEMM of one step share it)
  short step = -1;                   // plan step (diagnostics)
  short par_dw_c = 0;                // channels of the depthwise part at the head of the parameter block (0: none)
  short no_barrier = 0;              // the next op works on disjoint data: no barrier between the two
  short _pad[2] = {0, 0};
};
static_assert(sizeof(ChainOp) == 128, "ChainOp: 64 of them are passed by value as a kernel parameter");

struct ChainChunk {   // one weight chunk: K range [k0, k0 + kc) of a GEMM, hi planes then lo planes, [kc/8][Np][8 halves] each
  long long w_off = 0;   // floats into the weight arena
  int bytes = 0;
  int k0 = 0, kc = 0, _pad = 0;
};
struct ChainPar {     // one parameter block: [dw_w 9*C][dw_b C][bias Np][alpha Np] floats (absent parts have size 0)
  long long w_off = 0;
  int bytes = 0;
  int dw_c = 0;          // C of the depthwise part (0: none)
  int np = 0;
  int has_alpha = 0;
};

struct ChainLoad {    // one bulk copy of the producer warp, in issue order: a weight chunk or a parameter block
  long long w_off = 0;   // floats into the weight arena
  int bytes = 0;
  int is_par = 0;
};

// Fixed shared-memory map of the kernel.
constexpr int kChainArena = 141312;          // tensors
constexpr int kChainWSlot = 32768;           // x2
constexpr int kChainParSlot = 6144;          // x2
constexpr int kChainMaxOps = 64, kChainMaxLoads = 96;
constexpr int kChainThreads = 512;           // workers; + one producer warp
// K values per weight chunk: as many as fit in a slot (hi + lo halves: 4 bytes per weight), a multiple of 16
inline __host__ __device__ int chain_kc_max(int Np) { int v = (kChainWSlot / (4 * Np)) / 16 * 16; return v < 16 ? 16 : v; }

struct ChainPlan {
  bool valid = false;
  int first_step = 0, last_step = -1;         // plan steps [first, last] replaced by the chain
  int items = 0;                              // items per group (rows of the largest map * items <= 128)
  std::vector<ChainOp> ops;
  std::vector<ChainChunk> chunks;
  std::vector<ChainPar> pars;
  std::vector<ChainLoad> loads;               // the chunk / parameter loads in the order the producer issues them
  std::string text;
};

// Appends the chains' packed weights to plan.weights and fills plan.chains (two chains when the maps shrink along the run: the
// second packs four times as many items into a group).  Returns false (no chains) when the graph has no chainable tail; never
// fails the plan.
bool build_chain(Plan& plan);

}  // namespace fdl
