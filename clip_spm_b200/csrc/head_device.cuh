// Device helpers shared by the metric-head kernels (head_kernels.cu, sibling_heads.cu): float4 arithmetic, warp /
// block reductions, and the label -> class-index rule of the reference (rank among the sorted distinct labels).
#pragma once
#include <cuda_runtime.h>

namespace spm {

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result valid in every thread; `red` is >= 32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : 0.f;
  t = warp_sum(t);
  return t;
}
// class index of every support video: rank of its label among the episode's sorted distinct labels
// (== position in torch.unique(labels), model_clipspm.py:133).  Returns the number of classes.
__device__ __forceinline__ int class_indices(const float* __restrict__ labels, int S, int* cls /* smem [S] */) {
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float me = labels[s];
    int rank = 0;
    for (int j = 0; j < S; ++j) {
      const float o = labels[j];
      if (o < me) {
        bool first = true;  // count each distinct smaller label once
        for (int k = 0; k < j; ++k) first = first && (labels[k] != o);
        rank += first ? 1 : 0;
      }
    }
    cls[s] = rank;
  }
  __syncthreads();
  int W = 0;
  for (int s = 0; s < S; ++s) W = max(W, cls[s] + 1);
  return W;
}

}  // namespace spm
