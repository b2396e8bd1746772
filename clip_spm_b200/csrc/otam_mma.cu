// Metric tail at batch scale: the frame-to-frame products of cos_sim (models/myRes.py:756-765) as ONE small GEMM per
// CTA on the tensor cores, followed by the OTAM wavefronts (otam_dp.cuh) of every (query, class, direction).
//
// Why: the streaming kernel (otam.cu) is instruction-bound -- 36 k warp instructions per (problem, query) CTA, the
// support set re-read by every query's CTA -- and sits at ~0.12 of the HBM roofline for 1000 problems.  Here a CTA
// owns a problem (or one query of it when there are too few problems to fill the GPU): its [QT x WT x D] product is
// split over the 8 warps as (M groups) x (N groups) x (K slices); every operand byte is read from HBM exactly once,
// straight into mma fragments, no shared-memory staging of the operands.
//
// Exactness: the products must be fp32-accurate (distances are 1 - cos of nearly parallel frames), so each operand
// is split x = hi + lo with hi = tf32(x) and three m16n8k8 tf32 MMAs accumulate lo*hi + hi*lo + hi*hi in fp32
// ("3xTF32"; the dropped lo*lo term is 2^-22 relative).
//
// Fragment trick: a dot product does not care in which order k is visited, so lane (g, t) of a warp loads ONE float2
// -- columns kc + 2t, kc + 2t + 1 of row g -- and uses .x/.y as its (k = t, k = t + 4) elements of the k-step; A and B
// use the same permutation.  All operand traffic is 8-byte loads covering one full 32-byte sector per row.
#include "head_kernels.cuh"
#include "otam_dp.cuh"
#include "profile.cuh"

namespace spm {

using namespace otam_dp;

namespace {

// x rounded to tf32 (nearest, ties away from zero) on the bit pattern: 2 integer instructions instead of the 4 that
// cvt.rna.tf32.f32 expands to on sm_100a (its inf/nan guard is not needed: such inputs give nan distances anyway)
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// MT x NT: m16 / n8 tiles per warp.  MG x NG x KG = 8 warps.  The CTA covers QG queries (rows = QG*T <= MG*MT*16)
// against all W classes (columns = W*T <= NG*NT*8).
template <int MT, int NT, int MG, int NG, int KG>
__global__ void __launch_bounds__(256, 2)
otam_mma_kernel(const float* __restrict__ sup, long long s_p, long long s_w, long long s_t,
                const float* __restrict__ tgt, long long t_p, long long t_q, long long t_t, int W, int Q, int QG, int T,
                int D, int single_direct, float alpha, float beta, float* __restrict__ out) {
  static_assert(MG * NG * KG == 8, "8 warps");
  constexpr int MP = MG * MT * 16, NP = NG * NT * 8;
  extern __shared__ __align__(16) float sm_om[];
  float* part = sm_om;                  // [KG][MP][NP] partial products of the K slices
  float* an = part + KG * MP * NP;      // [KG][MP] partial squared norms of the query frames
  float* bn = an + KG * MP;             // [KG][NP] ... of the support frames
  float* dist = bn + KG * NP;           // [QG][W][T][T]
  float* res = dist + QG * W * T * T;   // [QG][W][2]
  const int p = blockIdx.y, q0 = blockIdx.x * QG;
  const int nq = min(QG, Q - q0);
  const int QT = nq * T, WT = W * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int kg = warp % KG, ng = (warp / KG) % NG, mg = warp / (KG * NG);

  // ---- phase 1: products and norms.  Rows past the end are clamped to the last valid row (results unused).
  const float* tgt_p = tgt + p * t_p + q0 * t_q + 2 * t;   // 32-bit row offsets from the problem's base: fewer registers
  const float* sup_p = sup + p * s_p + 2 * t;
  int arow[MT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = min((mg * MT + i) * 16 + g + 8 * h, QT - 1);
      arow[i][h] = (r / T) * (int)t_q + (r % T) * (int)t_t;
    }
  int brow[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int c = min((ng * NT + j) * 8 + g, WT - 1);
    brow[j] = (c / T) * (int)s_w + (c % T) * (int)s_t;
  }
  float acc[MT][NT][4];
  float na[MT][2], nb[NT];
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    na[i][0] = na[i][1] = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) nb[j] = 0.f;

  const int kslice = D / KG;
#pragma unroll 1
  for (int kc = kg * kslice; kc < (kg + 1) * kslice; kc += 8) {
    float2 va[MT][2], vb[NT];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      va[i][0] = __ldg(reinterpret_cast<const float2*>(tgt_p + arow[i][0] + kc));
      va[i][1] = __ldg(reinterpret_cast<const float2*>(tgt_p + arow[i][1] + kc));
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) vb[j] = __ldg(reinterpret_cast<const float2*>(sup_p + brow[j] + kc));
    uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      nb[j] = fmaf(vb[j].x, vb[j].x, fmaf(vb[j].y, vb[j].y, nb[j]));
      bh[j][0] = tf32_hi(vb[j].x); bl[j][0] = __float_as_uint(vb[j].x - __uint_as_float(bh[j][0]));
      bh[j][1] = tf32_hi(vb[j].y); bl[j][1] = __float_as_uint(vb[j].y - __uint_as_float(bh[j][1]));
    }
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      na[i][0] = fmaf(va[i][0].x, va[i][0].x, fmaf(va[i][0].y, va[i][0].y, na[i][0]));
      na[i][1] = fmaf(va[i][1].x, va[i][1].x, fmaf(va[i][1].y, va[i][1].y, na[i][1]));
      // a0 (row g, k t)  a1 (row g+8, k t)  a2 (row g, k t+4)  a3 (row g+8, k t+4)
      const float x[4] = {va[i][0].x, va[i][1].x, va[i][0].y, va[i][1].y};
      uint32_t ah[4], al[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        ah[e] = tf32_hi(x[e]);
        al[e] = __float_as_uint(x[e] - __uint_as_float(ah[e]));
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        mma_tf32(acc[i][j], al, bh[j][0], bh[j][1]);
        mma_tf32(acc[i][j], ah, bl[j][0], bl[j][1]);
        mma_tf32(acc[i][j], ah, bh[j][0], bh[j][1]);
      }
    }
  }
  // partial results of this warp's K slice -> shared memory (c0,c1: row g, cols 2t,2t+1; c2,c3: row g+8)
  float* pk = part + kg * MP * NP;
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int r = (mg * MT + i) * 16 + g, c = (ng * NT + j) * 8 + 2 * t;
      *reinterpret_cast<float2*>(pk + r * NP + c) = make_float2(acc[i][j][0], acc[i][j][1]);
      *reinterpret_cast<float2*>(pk + (r + 8) * NP + c) = make_float2(acc[i][j][2], acc[i][j][3]);
    }
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v = na[i][h];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (ng == 0 && t == 0) an[kg * MP + (mg * MT + i) * 16 + g + 8 * h] = v;
    }
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    float v = nb[j];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (mg == 0 && t == 0) bn[kg * NP + (ng * NT + j) * 8 + g] = v;
  }
  __syncthreads();
  // K slices summed in a fixed order (deterministic), norms finished in place
  for (int i = threadIdx.x; i < MP + NP; i += blockDim.x) {
    float* v = i < MP ? an + i : bn + (i - MP);
    const int stride = i < MP ? MP : NP;
    float s = v[0];
#pragma unroll
    for (int k = 1; k < KG; ++k) s += v[k * stride];
    v[0] = sqrtf(s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < QT * WT; i += blockDim.x) {
    const int m = i / WT, n = i % WT;
    float s = part[m * NP + n];
#pragma unroll
    for (int k = 1; k < KG; ++k) s += part[k * MP * NP + m * NP + n];
    // dist[q][w][tq][ts] = 1 - cos_sim  (myRes.py:756-765: x.y / (|x||y| + 0.01))
    dist[(((m / T) * W + n / T) * T + m % T) * T + n % T] = 1.f - s / (an[m] * bn[n] + 0.01f);
  }
  __syncthreads();
  // ---- phase 2: the (query, class, direction) DPs as anti-diagonal wavefronts, two per warp when a DP fits 16 lanes
  const int ndir = single_direct ? 1 : 2;
  const int n_dp = nq * W * ndir;
  if (T + 2 <= 16) {
    for (int j0 = warp * 2; j0 < n_dp; j0 += 16) {
      const int j = j0 + (lane >> 4);
      const bool valid = j < n_dp;
      const int qw = valid ? j / ndir : 0, dir = valid ? j % ndir : 0;
      const float* dw = dist + qw * T * T;
      const float r = otam_wavefront<16>(T, valid, dw, dir);
      if (valid && (lane & 15) == 0) res[qw * 2 + dir] = r;
    }
  } else {
    for (int j = warp; j < n_dp; j += 8) {
      const int qw = j / ndir, dir = j % ndir;
      const float* dw = dist + qw * T * T;
      const float r = otam_wavefront<32>(T, true, dw, dir);
      if (lane == 0) res[qw * 2 + dir] = r;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nq * W; i += blockDim.x) {
    const float r = res[i * 2] + (single_direct ? 0.f : res[i * 2 + 1]);
    float* o = out + ((long long)p * Q + q0) * W + i;
    *o = (beta != 0.f ? beta * (*o) : 0.f) + alpha * r;
  }
}

template <int MT, int NT, int MG, int NG, int KG>
size_t mma_smem(int QG, int W, int T) {
  constexpr int MP = MG * MT * 16, NP = NG * NT * 8;
  return (size_t)(KG * MP * NP + KG * MP + KG * NP + QG * W * T * T + QG * W * 2) * sizeof(float);
}

template <int MT, int NT, int MG, int NG, int KG>
int launch(cudaStream_t st, const float* sup, long long s_p, long long s_w, long long s_t, const float* tgt,
           long long t_p, long long t_q, long long t_t, int P, int W, int Q, int QG, int T, int D, int single_direct,
           float alpha, float beta, float* out) {
  const size_t smem = mma_smem<MT, NT, MG, NG, KG>(QG, W, T);
  if (smem > 100 * 1024) return -3;
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(otam_mma_kernel<MT, NT, MG, NG, KG>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((Q + QG - 1) / QG, P);
  otam_mma_kernel<MT, NT, MG, NG, KG><<<grid, 256, smem, st>>>(sup, s_p, s_w, s_t, tgt, t_p, t_q, t_t, W, Q, QG, T, D,
                                                               single_direct, alpha, beta, out);
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}

}  // namespace

// Returns -3 when the shape has no tensor-core instantiation (the caller then uses the streaming kernel).
int k_otam_mma(cudaStream_t st, const float* sup, long long s_p, long long s_w, long long s_t, const float* tgt,
               long long t_p, long long t_q, long long t_t, int P, int W, int Q, int T, int D, int single_direct,
               float alpha, float beta, float* out) {
  static const int sms = [] {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  if (T < 2 || T > 30 || D % 128 != 0) return -3;
  if (((s_p | s_w | s_t | t_p | t_q | t_t) & 1) != 0) return -3;                      // 8-byte fragment loads
  if ((reinterpret_cast<uintptr_t>(sup) | reinterpret_cast<uintptr_t>(tgt)) & 7) return -3;
  const long long lim = 1LL << 30;   // within-problem row offsets are 32-bit
  if (s_w < 0 || s_t < 0 || t_q < 0 || t_t < 0 || W * s_w + T * s_t >= lim || Q * t_q + T * t_t >= lim) return -3;
  const int WT = W * T;
  // a CTA per problem when the problems alone fill the GPU twice over, else a CTA per (problem, query)
  const int QG = (P >= 2 * sms && Q * T <= 96) ? Q : 1;
  const int QT = QG * T;
#define SPM_OTAM_MMA(MT, NT, MG, NG, KG) \
  return launch<MT, NT, MG, NG, KG>(st, sup, s_p, s_w, s_t, tgt, t_p, t_q, t_t, P, W, Q, QG, T, D, single_direct, alpha, beta, out)
  if (WT <= 40) {
    if (QT <= 16) SPM_OTAM_MMA(1, 5, 1, 1, 8);
    if (QT <= 48) SPM_OTAM_MMA(3, 5, 1, 1, 8);
    if (QT <= 96) SPM_OTAM_MMA(3, 5, 2, 1, 4);
  } else if (WT <= 80) {
    if (QT <= 16) SPM_OTAM_MMA(1, 5, 1, 2, 4);
    if (QT <= 48) SPM_OTAM_MMA(3, 5, 1, 2, 4);
    if (QT <= 96) SPM_OTAM_MMA(3, 5, 2, 2, 2);
  }
#undef SPM_OTAM_MMA
  return -3;
}

}  // namespace spm
