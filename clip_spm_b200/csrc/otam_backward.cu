// Backward of the metric tail (first piece of the training step): gradients of
//   out[p,q,w] = alpha * ( OTAM(1 - cos_sim(target[p,q], support[p,w])) + OTAM(transposed) )
// with respect to the frame features, i.e. what autograd derives for models/model_clipspm.py:348-362 +
// myRes.py:756-765 (cos_sim) + myRes.py:821-855 (OTAM_cum_dist_v2) when the reference trains (run/main_run.py:245-254).
//
// Kernel 1 (one CTA per (pair problem, query video)) recomputes the T x T dot products / norms, then one warp per
// class runs the soft-min DP forward as an anti-diagonal wavefront keeping the whole cumulative table in shared
// memory, and sweeps it BACKWARD along the same anti-diagonals: g(l,m) = d out / d C(l,m) pulls from its (up to
// three) successors with their soft-min weights exp(-(C_pred - softmin)/lambda).  d out / d dist(l,j) = g(l, j+1).
// Both directions add into one T x T table, which is turned into the coefficients of the feature gradients:
//   sim = x.y / (|x||y| + eps):  d sim/dx = y/den - (x.y)|y| / (den^2 |x|) x
// Kernels 2 / 3 then form grad_target[p,q,t,:] and grad_support[p,w,t,:] as small deterministic reductions
// (no atomics: every output row is owned by one CTA).
#include <algorithm>

#include "../../include/clipspm_b200.h"
#include "api_common.cuh"
#include "profile.cuh"

namespace spm {
namespace {
constexpr float LBDA = 0.5f, INV_LBDA = 2.0f, EPS = 0.01f;
constexpr int TMAX = 30, CW = 32;  // cumulative table row pitch (T + 2 <= 32)

// forward DP of one direction into c[T][CW] (lane m owns padded column m), exact expf/logf
template <class DistFn>
__device__ void otam_forward_table(int T, float* c, DistFn dist) {
  const int m = threadIdx.x & 31;
  for (int k = 0; k <= 2 * T; ++k) {
    const int l = k - m;
    if (l >= 0 && l < T && m <= T + 1) {
      const float d = (m >= 1 && m <= T) ? dist(l, m - 1) : 0.f;
      float v;
      if (m == 0) v = 0.f;
      else if (l == 0) v = d + c[m - 1];
      else {
        const float diag = c[(l - 1) * CW + m - 1], left = c[l * CW + m - 1];
        if (m == 1 || m == T + 1) {
          const float up = c[(l - 1) * CW + m];
          const float mn = fminf(diag, fminf(up, left));
          v = d + mn - LBDA * logf(expf((mn - diag) * INV_LBDA) + expf((mn - up) * INV_LBDA) + expf((mn - left) * INV_LBDA));
        } else {
          const float mn = fminf(diag, left);
          v = d + mn - LBDA * logf(expf((mn - diag) * INV_LBDA) + expf((mn - left) * INV_LBDA));
        }
      }
      c[l * CW + m] = v;
    }
    __syncwarp();
  }
}

// backward sweep: adds scale * d out / d dist(l, j) into G through `add(l, j, value)`
template <class DistFn, class AddFn>
__device__ void otam_backward_table(int T, const float* c, float* g, DistFn dist, AddFn add) {
  const int m = threadIdx.x & 31;
  auto dpad = [&](int l, int mm) { return (mm >= 1 && mm <= T) ? dist(l, mm - 1) : 0.f; };
  // weight with which predecessor value `cp` enters the soft-min of successor (ls, ms)
  auto wgt = [&](float cp, int ls, int ms) { return expf(-(cp - (c[ls * CW + ms] - dpad(ls, ms))) * INV_LBDA); };
  for (int k = 2 * T; k >= 1; --k) {
    const int l = k - m;
    if (l >= 0 && l < T && m >= 1 && m <= T + 1) {
      const float cp = c[l * CW + m];
      float v = (l == T - 1 && m == T + 1) ? 1.f : 0.f;
      if (m + 1 <= T + 1) {                                       // successor (l, m+1) takes us as its "left" input
        const float gs = g[l * CW + m + 1];
        v += (l == 0) ? gs : gs * wgt(cp, l, m + 1);
        if (l + 1 < T) v += g[(l + 1) * CW + m + 1] * wgt(cp, l + 1, m + 1);   // (l+1, m+1): "diagonal" input
      }
      if (l + 1 < T && (m == 1 || m == T + 1)) v += g[(l + 1) * CW + m] * wgt(cp, l + 1, m);   // (l+1, m): "up"
      g[l * CW + m] = v;
      if (m <= T) add(l, m - 1, v);
    }
    __syncwarp();
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
otam_bwd_coeff_kernel(const float* __restrict__ sup, const float* __restrict__ tgt, const float* __restrict__ grad_out,
                      int W, int Q, int T, int single_direct, float alpha, float* __restrict__ coefA,
                      float* __restrict__ coefB, float* __restrict__ coefC) {
  constexpr int D = NV * 128;
  extern __shared__ __align__(16) float sm[];
  float* sq = sm;                       // [T][D]
  float* qn = sq + T * D;               // [32]
  float* sn = qn + 32;                  // [W][T]
  float* dot = sn + W * T;              // [W][T][T]  dot[w][tq][ts]
  float* G = dot + W * T * T;           // [W][T][T]  d out / d dist
  float* tab = G + W * T * T;           // [8 warps][2][T][CW]  cumulative table + its gradient
  const int q = blockIdx.x, p = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* tq = tgt + ((long long)p * Q + q) * T * D;
  for (int i = threadIdx.x; i < T * (D / 4); i += blockDim.x)
    reinterpret_cast<float4*>(sq)[i] = reinterpret_cast<const float4*>(tq)[i];
  for (int i = threadIdx.x; i < W * T * T; i += blockDim.x) G[i] = 0.f;
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
  for (int j = warp; j < W * T; j += 8) {
    const int w = j / T, ts = j % T;
    const float4* sp = reinterpret_cast<const float4*>(sup + (((long long)p * W + w) * T + ts) * D);
    float4 sv[NV];
    float nn = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      sv[i] = __ldg(sp + i * 32 + lane);
      nn += (sv[i].x * sv[i].x + sv[i].y * sv[i].y) + (sv[i].z * sv[i].z + sv[i].w * sv[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
    if (lane == 0) sn[w * T + ts] = sqrtf(nn);
    for (int t = 0; t < T; ++t) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 qv = reinterpret_cast<const float4*>(sq + t * D)[i * 32 + lane];
        a = fmaf(sv[i].x, qv.x, a); a = fmaf(sv[i].y, qv.y, a); a = fmaf(sv[i].z, qv.z, a); a = fmaf(sv[i].w, qv.w, a);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) dot[(w * T + t) * T + ts] = a;
    }
  }
  __syncthreads();
  // ---- DPs: one warp per class, direction after direction
  float* c = tab + warp * 2 * T * CW;
  float* g = c + T * CW;
  for (int w = warp; w < W; w += 8) {
    const float* dw = dot + w * T * T;
    const float* snw = sn + w * T;
    float* Gw = G + w * T * T;
    auto d0 = [&](int l, int j) { return 1.f - dw[l * T + j] / (qn[l] * snw[j] + EPS); };   // rows = query frames
    auto d1 = [&](int l, int j) { return 1.f - dw[j * T + l] / (qn[j] * snw[l] + EPS); };   // transposed problem
    for (int i = lane; i < 2 * T * CW; i += 32) c[i] = 0.f;
    __syncwarp();
    otam_forward_table(T, c, d0);
    otam_backward_table(T, c, g, d0, [&](int l, int j, float v) { Gw[l * T + j] += v; });
    if (!single_direct) {
      for (int i = lane; i < 2 * T * CW; i += 32) c[i] = 0.f;
      __syncwarp();
      otam_forward_table(T, c, d1);
      otam_backward_table(T, c, g, d1, [&](int l, int j, float v) { Gw[j * T + l] += v; });
    }
    __syncwarp();
    // ---- coefficients of the feature gradients for this (p, q, w)
    const float go = alpha * grad_out[((long long)p * Q + q) * W + w];
    const long long base = ((long long)p * Q + q) * W + w;
    for (int i = lane; i < T * T; i += 32) {
      const int l = i / T, j = i % T;
      const float den = qn[l] * snw[j] + EPS;
      coefA[base * T * T + i] = -go * Gw[i] / den;
    }
    for (int l = lane; l < T; l += 32) {   // B[l] = sum_j Gs * dot * |y_j| / (den^2 |x_l|)
      float b = 0.f;
      for (int j = 0; j < T; ++j) {
        const float den = qn[l] * snw[j] + EPS;
        b += -go * Gw[l * T + j] * dw[l * T + j] * snw[j] / (den * den * fmaxf(qn[l], 1e-30f));
      }
      coefB[base * T + l] = b;
    }
    for (int j = lane; j < T; j += 32) {   // C[j] = sum_l Gs * dot * |x_l| / (den^2 |y_j|)
      float cc = 0.f;
      for (int l = 0; l < T; ++l) {
        const float den = qn[l] * snw[j] + EPS;
        cc += -go * Gw[l * T + j] * dw[l * T + j] * qn[l] / (den * den * fmaxf(snw[j], 1e-30f));
      }
      coefC[base * T + j] = cc;
    }
    __syncwarp();
  }
}

// grad_target[p,q,l,:] = sum_w sum_j A[p,q,w][l][j] * support[p,w,j,:] - (sum_w B[p,q,w][l]) * target[p,q,l,:]
__global__ void otam_bwd_target_kernel(const float* __restrict__ sup, const float* __restrict__ tgt,
                                       const float* __restrict__ coefA, const float* __restrict__ coefB, int W, int Q,
                                       int T, int D, float* __restrict__ grad_tgt) {
  const int l = blockIdx.x, q = blockIdx.y, p = blockIdx.z;
  for (int d4 = threadIdx.x; d4 < D / 4; d4 += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float bsum = 0.f;
    for (int w = 0; w < W; ++w) {
      const long long base = ((long long)p * Q + q) * W + w;
      bsum += coefB[base * T + l];
      for (int j = 0; j < T; ++j) {
        const float a = coefA[base * T * T + l * T + j];
        const float4 y = __ldg(reinterpret_cast<const float4*>(sup + (((long long)p * W + w) * T + j) * D) + d4);
        acc.x = fmaf(a, y.x, acc.x); acc.y = fmaf(a, y.y, acc.y); acc.z = fmaf(a, y.z, acc.z); acc.w = fmaf(a, y.w, acc.w);
      }
    }
    const long long row = (((long long)p * Q + q) * T + l) * D;
    const float4 x = __ldg(reinterpret_cast<const float4*>(tgt + row) + d4);
    reinterpret_cast<float4*>(grad_tgt + row)[d4] =
        make_float4(acc.x - bsum * x.x, acc.y - bsum * x.y, acc.z - bsum * x.z, acc.w - bsum * x.w);
  }
}

// grad_support[p,w,j,:] = sum_q sum_l A[p,q,w][l][j] * target[p,q,l,:] - (sum_q C[p,q,w][j]) * support[p,w,j,:]
__global__ void otam_bwd_support_kernel(const float* __restrict__ sup, const float* __restrict__ tgt,
                                        const float* __restrict__ coefA, const float* __restrict__ coefC, int W, int Q,
                                        int T, int D, float* __restrict__ grad_sup) {
  const int j = blockIdx.x, w = blockIdx.y, p = blockIdx.z;
  for (int d4 = threadIdx.x; d4 < D / 4; d4 += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float csum = 0.f;
    for (int q = 0; q < Q; ++q) {
      const long long base = ((long long)p * Q + q) * W + w;
      csum += coefC[base * T + j];
      for (int l = 0; l < T; ++l) {
        const float a = coefA[base * T * T + l * T + j];
        const float4 x = __ldg(reinterpret_cast<const float4*>(tgt + (((long long)p * Q + q) * T + l) * D) + d4);
        acc.x = fmaf(a, x.x, acc.x); acc.y = fmaf(a, x.y, acc.y); acc.z = fmaf(a, x.z, acc.z); acc.w = fmaf(a, x.w, acc.w);
      }
    }
    const long long row = (((long long)p * W + w) * T + j) * D;
    const float4 y = __ldg(reinterpret_cast<const float4*>(sup + row) + d4);
    reinterpret_cast<float4*>(grad_sup + row)[d4] =
        make_float4(acc.x - csum * y.x, acc.y - csum * y.y, acc.z - csum * y.z, acc.w - csum * y.w);
  }
}
}  // namespace
}  // namespace spm

using namespace spm;

extern "C" int spm_otam_distance_backward(void* stream, int n_pairs, int W, int Q, int T, int D, const float* support,
                                          const float* target, int single_direct, float alpha, const float* grad_out,
                                          float* grad_support, float* grad_target) {
  SPM_CHECK(support && target && grad_out && grad_support && grad_target, "spm_otam_distance_backward: null argument");
  SPM_CHECK(T >= 2 && T <= TMAX && W >= 1 && W <= 32 && Q >= 1 && (D == 512 || D == 1024),
            "spm_otam_distance_backward: unsupported shape (2 <= T <= 30, W <= 32, D in {512, 1024})");
  if (n_pairs <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(T * D + 32 + W * T + 2 * W * T * T + 8 * 2 * T * CW) * sizeof(float);
  SPM_CHECK(smem <= 200 * 1024, "spm_otam_distance_backward: problem too large for shared memory");
  static bool attr = false;
  if (!attr) {
    SPM_CUDA(cudaFuncSetAttribute(otam_bwd_coeff_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SPM_CUDA(cudaFuncSetAttribute(otam_bwd_coeff_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  const long long npq = (long long)n_pairs * Q * W;
  float* scratch = nullptr;  // stream-ordered scratch: coefficient tables A [npq,T,T], B [npq,T], C [npq,T]
  SPM_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), (size_t)npq * (T * T + 2 * T) * sizeof(float), st));
  float* coefA = scratch;
  float* coefB = coefA + npq * T * T;
  float* coefC = coefB + npq * T;
  const dim3 grid(Q, n_pairs);
  if (D == 512)
    otam_bwd_coeff_kernel<4><<<grid, 256, smem, st>>>(support, target, grad_out, W, Q, T, single_direct, alpha, coefA, coefB, coefC);
  else
    otam_bwd_coeff_kernel<8><<<grid, 256, smem, st>>>(support, target, grad_out, W, Q, T, single_direct, alpha, coefA, coefB, coefC);
  count_launch();
  otam_bwd_target_kernel<<<dim3(T, Q, n_pairs), 128, 0, st>>>(support, target, coefA, coefB, W, Q, T, D, grad_target);
  count_launch();
  otam_bwd_support_kernel<<<dim3(T, W, n_pairs), 128, 0, st>>>(support, target, coefA, coefC, W, Q, T, D, grad_support);
  count_launch();
  const cudaError_t e = cudaGetLastError();
  cudaFreeAsync(scratch, st);
  if (e != cudaSuccess) { set_error(std::string("spm_otam_distance_backward launch: ") + cudaGetErrorString(e)); return 1; }
  return 0;
}
