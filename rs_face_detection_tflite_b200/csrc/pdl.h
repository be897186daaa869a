// pdl.h -- programmatic dependent launch (PDL) for the chains of network kernels.
//
// A planned graph is 21..55 launches on one stream, many of them 20..60 us long.  With PDL the prologue of launch
// i+1 (mbarrier init, TMEM allocation, weights -> shared memory: nothing that depends on launch i) runs while launch
// i drains; `pdl_wait()` then blocks until launch i has completed and its writes are visible, so the data flow is
// exactly the stream order.  Every kernel launched through launch_pdl() executes, in every CTA,
//     prologue (incl. tcgen05.alloc)  ->  __syncthreads  ->  pdl_launch_dependents()  ->  pdl_wait()  ->  body
// The trigger comes AFTER the CTA's own TMEM allocation: a dependent grid only starts once every CTA of the primary
// has triggered, i.e. is resident and owns its TMEM columns, so a waiting dependent CTA can never starve a primary CTA.
// Both instructions are no-ops for a kernel launched without the attribute (FDL_PDL=0, or a plain <<<>>> launch).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace fdl {

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("FDL_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}

// SMs the persistent kernels spread over.  148 by default; a smaller number (FDL_PERSIST_SMS) leaves SMs free for the
// small launches of other pipeline lanes' streams to run beside a persistent kernel instead of queueing behind it.
inline int persist_sms() {
  static const int n = [] { const char* e = getenv("FDL_PERSIST_SMS"); int v = e ? atoi(e) : 148; return v < 1 ? 1 : (v > 148 ? 148 : v); }();
  return n;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace fdl
