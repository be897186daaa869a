// chain_kernel.cuh -- launch interface of the tail-chain kernel (chain_kernel.cu; program format in chain.h).
#pragma once
#include <cuda_runtime.h>

#include "chain.h"

namespace fdl {

struct ChainArgs {
  ChainOp ops[kChainMaxOps];           // the program, by value: every thread reads the same op -> constant-bank (uniform) loads,
                                       // which keeps the tcgen05.mma operands of the issuing thread in uniform registers
  ChainLoad loads[kChainMaxLoads];     // the producer warp's copies, in order (also uniform loads: no dependent global reads)
  int n_ops = 0, n_loads = 0;
  const float* weights = nullptr;      // the net's weight arena (chunks and parameter blocks are offsets into it)
  float* arena = nullptr;              // the net's activation arena
  int B = 0;                           // batch the arena views are laid out for
  int items = 1;                       // items per group
  const int* n_active = nullptr;
};

cudaError_t chain_init();              // once per device
bool chain_enabled();                  // FDL_CHAIN (default 1)
cudaError_t launch_chain(const ChainArgs& a, cudaStream_t stream);

}  // namespace fdl
