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
#include "otam_dp.cuh"
#include "profile.cuh"
#include <cstdlib>
#include <cstring>

namespace spm {

using namespace otam_dp;

// NV: D = NV*128.  TP: number of dot-product accumulators kept per lane (power of two >= T).
template <int NV, int TP>
__global__ void __launch_bounds__(256, (TP <= 8 ? 5 : (TP <= 16 ? 4 : 2)))
otam_kernel(const float* __restrict__ sup, long long s_p, long long s_w, long long s_t, const float* __restrict__ tgt,
            long long t_p, long long t_q, long long t_t, int W, int Q, int T, int single_direct, float alpha,
            float beta, float* __restrict__ out, int dp_log) {
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
      const float4 v = reinterpret_cast<float4*>(sq + t * D)[c];
      a += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) qn[t] = sqrtf(a);
  }
  __syncthreads();
  // ---- phase 1: cosine distances.  A warp streams one support frame (coalesced float4, kept in registers) and
  // accumulates its dot product with every query frame; the TP partial sums per lane are then combined with a
  // reduce-scatter butterfly (TP-1 + log2(32/TP) shuffles instead of 5 per dot product).
  for (int j = warp; j < W * T; j += 8) {
    const int w = j / T, ts = j % T;
    const float4* sp = reinterpret_cast<const float4*>(sup + p * s_p + w * s_w + ts * s_t);
    float4 sv[NV];
    float nn = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      sv[i] = __ldg(sp + i * 32 + lane);
      nn += (sv[i].x * sv[i].x + sv[i].y * sv[i].y) + (sv[i].z * sv[i].z + sv[i].w * sv[i].w);
    }
    float acc[TP];
#pragma unroll
    for (int t = 0; t < TP; ++t) {
      float a = 0.f;
      if (t < T) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float4 qv = reinterpret_cast<const float4*>(sq + t * D)[i * 32 + lane];
          a = fmaf(sv[i].x, qv.x, a); a = fmaf(sv[i].y, qv.y, a); a = fmaf(sv[i].z, qv.z, a); a = fmaf(sv[i].w, qv.w, a);
        }
      }
      acc[t] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
    // reduce-scatter: after the round with offset `off`, a lane keeps the half of the accumulators selected by
    // its bit `off`; the index it finally owns is t_own (built from the same bits)
    int t_own = 0;
#pragma unroll
    for (int off = 16, n = TP; n > 1; off >>= 1, n >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int jj = 0; jj < n / 2; ++jj) {
        const float mine = upper ? acc[jj + n / 2] : acc[jj];
        const float other = upper ? acc[jj] : acc[jj + n / 2];
        acc[jj] = mine + __shfl_xor_sync(0xffffffffu, other, off);
      }
      if (upper) t_own += n / 2;
    }
    float tot = acc[0];
#pragma unroll
    for (int off = 16 / TP; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
    if ((lane & (32 / TP - 1)) == 0 && t_own < T)
      dist[(w * T + t_own) * T + ts] = 1.f - tot / (qn[t_own] * sqrtf(nn) + 0.01f);
  }
  __syncthreads();
  // ---- phase 2: the (class, direction) DPs as anti-diagonal wavefronts, 32 / (T + 2) of them per warp
  const int ndir = single_direct ? 1 : 2;
  const int n_dp = W * ndir;
  const int per_warp = otam_dps_per_warp(T), seg = lane / (T + 2), m = lane % (T + 2);
  for (int j0 = warp * per_warp; j0 < n_dp; j0 += 8 * per_warp) {
    const int j = j0 + seg;
    const bool valid = seg < per_warp && j < n_dp;
    const int w = valid ? j / ndir : 0, dir = valid ? j % ndir : 0;
    const float r = otam_wavefront_auto(T, m, valid, dist + w * T * T, dir, dp_log);
    if (valid && m == T + 1) res[w * 2 + dir] = r;
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

#define SPM_OTAM_FOR_ALL(X) X(4, 8) X(4, 16) X(4, 32) X(8, 8) X(8, 16) X(8, 32)

int k_otam_init() {
  const int bytes = OTAM_SMEM_MAX;
#define SPM_OTAM_ATTR(NV, TP)                                                                                      \
  {                                                                                                                \
    cudaError_t e = cudaFuncSetAttribute(otam_kernel<NV, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); \
    if (e != cudaSuccess) return (int)e;                                                                           \
  }
  SPM_OTAM_FOR_ALL(SPM_OTAM_ATTR)
#undef SPM_OTAM_ATTR
  return 0;
}

int k_otam(cudaStream_t st, const float* sup, long long s_p, long long s_w, long long s_t, const float* tgt,
           long long t_p, long long t_q, long long t_t, int P, int W, int Q, int T, int D, int single_direct,
           float alpha, float beta, float* out) {
  if (T < 2 || T > 30 || W > 32 || (D != 512 && D != 1024)) return -2;
  if (P <= 0 || Q <= 0) return 0;
  // tensor-core kernel (otam_mma.cu) for every shape it is instantiated for; SPM_OTAM=stream keeps this kernel
  static const bool allow_mma = [] { const char* e = getenv("SPM_OTAM"); return e == nullptr || strcmp(e, "stream") != 0; }();
  if (allow_mma) {
    // batch scale (P >= 2 x #SM, headline shapes): the tcgen05 kernel of otam_tc.cu, then the persistent mma.sync kernel
    const int rt = k_otam_tc(st, sup, s_p, s_w, s_t, tgt, t_p, t_q, t_t, P, W, Q, T, D, single_direct, alpha, beta, out);
    if (rt != -3) return rt;
    const int rf = k_otam_fused(st, sup, s_p, s_w, s_t, tgt, t_p, t_q, t_t, P, W, Q, T, D, single_direct, alpha, beta, out);
    if (rf != -3) return rf;
    const int r = k_otam_mma(st, sup, s_p, s_w, s_t, tgt, t_p, t_q, t_t, P, W, Q, T, D, single_direct, alpha, beta, out);
    if (r != -3) return r;
  }
  dim3 grid(Q, P);
  const size_t smem = otam_smem(W, T, D);
  if (smem > (size_t)OTAM_SMEM_MAX) return -2;
  const int nv = D / 128, tp = T <= 8 ? 8 : (T <= 16 ? 16 : 32);
#define SPM_OTAM_LAUNCH(NV, TP)                                                                                    \
  if (nv == NV && tp == TP)                                                                                        \
    otam_kernel<NV, TP><<<grid, 256, smem, st>>>(sup, s_p, s_w, s_t, tgt, t_p, t_q, t_t, W, Q, T, single_direct,   \
                                                 alpha, beta, out, otam_dp_force_log());
  SPM_OTAM_FOR_ALL(SPM_OTAM_LAUNCH)
#undef SPM_OTAM_LAUNCH
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}

}  // namespace spm
