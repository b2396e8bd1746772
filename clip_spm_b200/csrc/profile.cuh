// Process-wide instrumentation: a counter of this library's kernel launches and optional CUDA-event timing of the
// GEMM launches (the dominant kernel) so bench.py can report a live roofline number for the timed region.
#pragma once
#include <cuda_runtime.h>

namespace spm {
void count_launch();
// returns false when profiling is off or the record pool is exhausted
bool profile_gemm_begin(cudaStream_t st, int tag, double flops, int* slot);
void profile_gemm_end(cudaStream_t st, int slot);
}  // namespace spm
