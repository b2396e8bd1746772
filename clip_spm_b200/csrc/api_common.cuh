// Error plumbing shared by the C-ABI translation units: no C++ exception crosses the boundary,
// every entry point returns an int status and records a message readable via spm_last_error().
#pragma once
#include <cuda_runtime.h>
#include <string>

namespace spm {

void set_error(const std::string& msg);
const char* get_error();

#define SPM_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::spm::set_error(std::string(#call) + ": " + cudaGetErrorString(_e));                   \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)

#define SPM_CHECK(cond, msg)                                                                  \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      ::spm::set_error(std::string(msg));                                                     \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)

#define SPM_TRY(expr)                                                                         \
  do {                                                                                        \
    if ((expr) != 0) return 1;                                                                \
  } while (0)

}  // namespace spm
