// Process-wide instrumentation: a counter of this library's kernel launches and optional CUDA-event timing of the
// GEMM launches (the dominant kernel) so bench.py can report a live roofline number for the timed region.
#pragma once
#include <cuda_runtime.h>

namespace spm {
void count_launch();
// returns false when profiling is off or the record pool is exhausted
bool profile_gemm_begin(cudaStream_t st, int tag, double flops, int* slot, int M = 0, int N = 0, int K = 0);
void profile_gemm_end(cudaStream_t st, int slot);
// true while GEMM launches are being event-timed: the encoder then keeps to one stream so that an event pair
// brackets exactly one kernel (with two streams the other chunk's kernels would run inside the bracket)
bool profile_armed();
}  // namespace spm
