// fdl_status.h -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <string>

#include "../../include/fdl.h"

namespace fdl {
// Records `msg` as the calling thread's last error and returns `code`.
int set_error(int code, const std::string& msg);
void clear_error();
}  // namespace fdl

#define FDL_CUDA_TRY(expr)                                                                             \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) return ::fdl::set_error(FDL_ERR_CUDA, std::string("CUDA: ") + cudaGetErrorString(_e) + " at " #expr); \
  } while (0)
