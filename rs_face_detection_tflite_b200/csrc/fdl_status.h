// fdl_status.h -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <new>
#include <exception>
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

// Nothing throws across the C ABI (include/fdl.h): every extern "C" entry point is a function-try-block ending in this handler.
#define FDL_ABI_CATCH                                                                                   \
  catch (const std::bad_alloc&) { return ::fdl::set_error(FDL_ERR_INTERNAL, "out of host memory"); }    \
  catch (const std::exception& e) { return ::fdl::set_error(FDL_ERR_INTERNAL, std::string("internal error: ") + e.what()); } \
  catch (...) { return ::fdl::set_error(FDL_ERR_INTERNAL, "internal error"); }
