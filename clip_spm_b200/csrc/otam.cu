// Metric tail: frame-to-frame cosine distances + (bi)directional OTAM soft-DTW, fused in one kernel.
//   cos_sim            models/myRes.py:756-765      x.y / (|x||y| + 0.01)
//   otam_distance      models/model_clipspm.py:348-362   d = 1 - cos_sim ; OTAM(d) + OTAM(d^T)
//   OTAM_cum_dist_v2   models/myRes.py:821-855      soft-min DP (lambda 0.5) over the [T, T+2] zero-padded grid
//
// One CTA per (pair problem, query video).  Phase 1: the query's T frames are staged in shared memory; each warp
// streams support frames (coalesced float4, the frame stays in registers) and produces T dot products + the norm
// per frame with warp shuffles.  Phase 2: every (class, direction) DP is an anti-diagonal wavefront inside ONE warp:
// lane m owns column m of the padded grid, diagonal k = l + m is one step, the three neighbours arrive by
// __shfl_up from lane m-1 (its last and second-to-last values) and from the lane's own last value:
// 2T+1 dependent steps instead of the reference's T*(T+2) sequential host-launched cells.
#include "head_kernels.cuh"
#include "profile.cuh"

namespace spm {

namespace {
constexpr float LBDA = 0.5f;

__device__ __forceinline__ float softmin2(float a, float b) {
  return -LBDA * logf(expf(-a / LBDA) + expf(-b / LBDA));
}
__device__ __forceinline__ float softmin3(float a, float b, float c) {
  return -LBDA * logf(expf(-a / LBDA) + expf(-b / LBDA) + expf(-c / LBDA));
}

// One warp, one DP.  dist(l, j): l = row (0..T-1), j = unpadded column (0..T-1).  Returns C[T-1, T+1] in every lane.
template <class DistFn>
__device__ __forceinline__ float otam_wavefront(int T, DistFn dist) {
  const int m = threadIdx.x & 31;  // padded column owned by this lane
  float v1 = 0.f, v2 = 0.f;        // this lane's last / second-to-last computed cells
  for (int k = 0; k <= 2 * T; ++k) {
    const float left = __shfl_up_sync(0xffffffffu, v1, 1);   // C[l,   m-1]
    const float diag = __shfl_up_sync(0xffffffffu, v2, 1);   // C[l-1, m-1]
    const int l = k - m;
    if (l >= 0 && l < T && m <= T + 1) {
      const float d = (m >= 1 && m <= T) ? dist(l, m - 1) : 0.f;
      float c;
      if (m == 0) c = 0.f;                                    // column 0 is never written (stays 0)
      else if (l == 0) c = d + left;                          // top row: plain prefix sum
      else if (m == 1 || m == T + 1) c = d + softmin3(diag, v1, left);   // (l-1,m-1), (l-1,m), (l,m-1)
      else c = d + softmin2(diag, left);                      // interior: no vertical neighbour
      v2 = v1;
      v1 = c;
    }
  }
  return __shfl_sync(0xffffffffu, v1, T + 1);
}
}  // namespace

template <int NV>  // D = NV * 128
__global__ void __launch_bounds__(256)
otam_kernel(const float* __restrict__ sup, long long s_p, long long s_w, long long s_t, const float* __restrict__ tgt,
            long long t_p, long long t_q, long long t_t, int W, int Q, int T, int single_direct, float alpha,
            float beta, float* __restrict__ out) {
  constexpr int D = NV * 128;
  extern __shared__ __align__(16) float sm_ot[];
  float* sq = sm_ot;               // [T][D] query frames
  float* qn = sq + T * D;          // [T] query norms
  float* dist = qn + 32;           // [W][T][T]  dist[w][tq][ts] = 1 - cos_sim
  float* res = dist + W * T * T;   // [W][2]
  const int q = blockIdx.x, p = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* tq = tgt + p * t_p + q * t_q;
  for (int i = threadIdx.x; i < T * (D / 4); i += blockDim.x) {
    const int t = i / (D / 4), c = i % (D / 4);
    reinterpret_cast<float4*>(sq + t * D)[c] = reinterpret_cast<const float4*>(tq + t * t_t)[c];
  }
  __syncthreads();
  for (int t = warp; t < T; t += 8) {
    float a = 0.f;
    for (int c = lane; c < D / 4; c += 32) {
      const float4 v = reinterpret_cast<const float4*>(sq + t * D)[c];
      a += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) qn[t] = sqrtf(a);
  }
  __syncthreads();
  // ---- phase 1: cosine distances
  for (int j = warp; j < W * T; j += 8) {
    const int w = j / T, ts = j % T;
    const float4* sp = reinterpret_cast<const float4*>(sup + p * s_p + w * s_w + ts * s_t);
    float4 sv[NV];
    float nn = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      sv[i] = sp[i * 32 + lane];
      nn += (sv[i].x * sv[i].x + sv[i].y * sv[i].y) + (sv[i].z * sv[i].z + sv[i].w * sv[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
    const float sn = sqrtf(nn);
    for (int t = 0; t < T; ++t) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 qv = reinterpret_cast<const float4*>(sq + t * D)[i * 32 + lane];
        a += (sv[i].x * qv.x + sv[i].y * qv.y) + (sv[i].z * qv.z + sv[i].w * qv.w);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) dist[(w * T + t) * T + ts] = 1.f - a / (qn[t] * sn + 0.01f);
    }
  }
  __syncthreads();
  // ---- phase 2: one warp per (class, direction) DP
  const int ndir = single_direct ? 1 : 2;
  for (int j = warp; j < W * ndir; j += 8) {
    const int w = j / ndir, dir = j % ndir;
    const float* dw = dist + w * T * T;
    float r;
    if (dir == 0) r = otam_wavefront(T, [&](int l, int c) { return dw[l * T + c]; });   // rows: query frames
    else r = otam_wavefront(T, [&](int l, int c) { return dw[c * T + l]; });            // transposed
    if (lane == 0) res[w * 2 + dir] = r;
  }
  __syncthreads();
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    const float r = res[w * 2] + (single_direct ? 0.f : res[w * 2 + 1]);
    float* o = out + ((long long)p * Q + q) * W + w;
    *o = (beta != 0.f ? beta * (*o) : 0.f) + alpha * r;
  }
}

constexpr int OTAM_SMEM_MAX = 200 * 1024;
static size_t otam_smem(int W, int T, int D) { return (size_t)(T * D + 32 + W * T * T + 2 * W) * sizeof(float); }

int k_otam_init() {
  const int bytes = OTAM_SMEM_MAX;
  cudaError_t e = cudaFuncSetAttribute(otam_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(otam_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  return (int)e;
}

int k_otam(cudaStream_t st, const float* sup, long long s_p, long long s_w, long long s_t, const float* tgt,
           long long t_p, long long t_q, long long t_t, int P, int W, int Q, int T, int D, int single_direct,
           float alpha, float beta, float* out) {
  if (T < 2 || T > 30 || W > 32 || (D != 512 && D != 1024)) return -2;
  if (P <= 0 || Q <= 0) return 0;
  dim3 grid(Q, P);
  const size_t smem = otam_smem(W, T, D);
  if (smem > (size_t)OTAM_SMEM_MAX) return -2;
  if (D == 512)
    otam_kernel<4><<<grid, 256, smem, st>>>(sup, s_p, s_w, s_t, tgt, t_p, t_q, t_t, W, Q, T, single_direct, alpha, beta, out);
  else
    otam_kernel<8><<<grid, 256, smem, st>>>(sup, s_p, s_w, s_t, tgt, t_p, t_q, t_t, W, Q, T, single_direct, alpha, beta, out);
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}

}  // namespace spm
