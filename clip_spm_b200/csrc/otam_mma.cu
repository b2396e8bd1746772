// Metric tail at batch scale, in two kernels:
//   cos_dist_mma_kernel   the frame-to-frame products of cos_sim (models/myRes.py:756-765) as ONE small GEMM per CTA
//                         on the tensor cores -> dist[p][q][w][tq][ts] = 1 - cos_sim (a [P,Q,W,T,T] scratch tensor)
//   otam_dp_kernel        the OTAM wavefronts (otam_dp.cuh) of every (problem, query, class, direction) at full
//                         occupancy, result accumulated into out[p][q][w]
//
// Why: the streaming kernel (otam.cu) is instruction-bound -- 36 k warp instructions per (problem, query) CTA, the
// support set re-read by every query's CTA -- and sits at ~0.12 of the HBM roofline for 1000 problems.  Here a CTA
// owns a problem (or one query of it when there are too few problems to fill the GPU): its [QT x WT x D] product is
// split over the 8 warps as (M groups) x (N groups) x (K slices); every operand byte is read from HBM exactly once.
// Why two kernels: the product needs ~128 registers per thread (60 accumulators), the DP ~40 and nothing but
// latency hiding; fused, the DP phase ran at 16 warps per SM and took a quarter of the kernel (ncu).  The scratch
// tensor is 4 % of the operand bytes and is read back from L2.
//
// Exactness: the products must be fp32-accurate (distances are 1 - cos of nearly parallel frames), so each operand
// is split x = hi + lo with hi = tf32(x) and three m16n8k8 tf32 MMAs accumulate lo*hi + hi*lo + hi*hi in fp32
// ("3xTF32"; the dropped lo*lo term is 2^-22 relative).
//
// Operand path: every warp owns a K slice and feeds itself -- a private ring of cp.async stages (16 floats of every
// row per stage), so a warp needs no block-wide barrier until its slice is done and keeps one stage in flight while
// it computes on the other.  16 warps per SM (128 registers each), as CTAs of 4 warps where the tile shape allows:
// four independent CTAs per SM de-phase their load / MMA / reduce phases better than two of 8 warps (82 -> 76 us).  Fragment trick: a dot product does not care in which order k is visited, so lane (g, t)
// reads ONE float2 -- columns 2t, 2t + 1 of an 8-float half stage of row g -- as its (k = t, k = t + 4) elements of
// the k-step; A and B use the same permutation, and the 32 lanes read 256 contiguous bytes (no bank conflicts).
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
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int NS = 2;        // ring stages per warp
constexpr int STAGE_K = 16;  // floats of every row per stage (two k-steps)

// MT x NT: m16 / n8 tiles per warp.  MG x NG x KG = NW warps (4 or 8; 16 warps per SM either way).  The CTA covers QG queries (rows = QG*T <= MG*MT*16)
// against all W classes (columns = W*T <= NG*NT*8).
template <int MT, int NT, int MG, int NG, int KG>
__global__ void __launch_bounds__(32 * MG * NG * KG, 16 / (MG * NG * KG))
cos_dist_mma_kernel(const float* __restrict__ sup, long long s_p, long long s_w, long long s_t,
                    const float* __restrict__ tgt, long long t_p, long long t_q, long long t_t, int W, int Q, int QG,
                    int T, int D, float* __restrict__ dist) {
  constexpr int NW = MG * NG * KG;
  static_assert(NW == 4 || NW == 8, "4 or 8 warps");
  constexpr int MP = MG * MT * 16, NP = NG * NT * 8;
  constexpr int R = MT * 16 + NT * 8;           // rows a warp stages: its A rows, then its B rows
  constexpr int STAGE = 2 * R * 8;              // floats per stage: [2 halves][R rows][8]
  constexpr int RING = NW * NS * STAGE, PART = KG * MP * NP;
  extern __shared__ __align__(16) float sm_om[];
  float* ring = sm_om;                          // [NW warps][NS][STAGE]; reused as `part` once every slice is done
  float* part = sm_om;                          // [KG][MP][NP] partial products of the K slices
  float* an = sm_om + (RING > PART ? RING : PART);   // [KG][MP] partial squared norms of the query frames
  float* bn = an + KG * MP;                     // [KG][NP] ... of the support frames
  const int p = blockIdx.y, q0 = blockIdx.x * QG;
  const int nq = min(QG, Q - q0);
  const int QT = nq * T, WT = W * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int kg = warp % KG, ng = (warp / KG) % NG, mg = warp / (KG * NG);
  const float* tgt_p = tgt + p * t_p + q0 * t_q;
  const float* sup_p = sup + p * s_p;

  // This lane's share of a stage: piece `it` = 16 bytes (column quarter c = lane & 3) of staged row it*8 + (lane >> 2).
  // Staged rows 0 .. MT*16-1 are the warp's A rows, the rest its B rows; MT*16 is a multiple of 8, so whether a piece
  // is A or B depends on `it` alone.  Offsets (floats, from the problem's base) are resolved once; rows past the end
  // are not loaded: whatever the ring holds there only reaches result rows/columns nobody reads.
  constexpr int NPIECE = R / 8;
  int goff[NPIECE];
#pragma unroll
  for (int it = 0; it < NPIECE; ++it) {
    if (it < MT * 2) {
      const int r = mg * MT * 16 + it * 8 + g;
      goff[it] = r < QT ? (r / T) * (int)t_q + (r % T) * (int)t_t + t * 4 : -1;
    } else {
      const int c = ng * NT * 8 + (it - MT * 2) * 8 + g;
      goff[it] = c < WT ? (c / T) * (int)s_w + (c % T) * (int)s_t + t * 4 : -1;
    }
  }
  float* my_ring = ring + warp * NS * STAGE;
  // destination of piece 0 inside a stage: half (t >> 1), row g, 16-byte part (t & 1); piece `it` is 8 rows further
  const int dst0 = (t >> 1) * (R * 8) + g * 8 + (t & 1) * 4;
  const int k_begin = kg * (D / KG), n_stage = (D / KG) / STAGE_K;
  auto issue = [&](int stage_idx) {
    float* dst = my_ring + (stage_idx % NS) * STAGE + dst0;
    const int kc = k_begin + stage_idx * STAGE_K;
#pragma unroll
    for (int it = 0; it < NPIECE; ++it)
      if (goff[it] >= 0) cp_async16(dst + it * 64, (it < MT * 2 ? tgt_p : sup_p) + goff[it] + kc);
    cp_async_commit();
  };

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

#pragma unroll
  for (int s = 0; s < NS - 1; ++s) {
    if (s < n_stage) issue(s); else cp_async_commit();
  }
#pragma unroll 1
  for (int s = 0; s < n_stage; ++s) {
    if (s + NS - 1 < n_stage) issue(s + NS - 1); else cp_async_commit();   // slot (s-1)%NS: consumed last iteration
    cp_async_wait<NS - 1>();
    __syncwarp();
    const float* st = my_ring + (s % NS) * STAGE;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float* sh = st + h * (R * 8) + 2 * t;
      uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float2 v = *reinterpret_cast<const float2*>(sh + (MT * 16 + j * 8 + g) * 8);
        nb[j] = fmaf(v.x, v.x, fmaf(v.y, v.y, nb[j]));
        bh[j][0] = tf32_hi(v.x); bl[j][0] = __float_as_uint(v.x - __uint_as_float(bh[j][0]));
        bh[j][1] = tf32_hi(v.y); bl[j][1] = __float_as_uint(v.y - __uint_as_float(bh[j][1]));
      }
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const float2 v0 = *reinterpret_cast<const float2*>(sh + (i * 16 + g) * 8);
        const float2 v1 = *reinterpret_cast<const float2*>(sh + (i * 16 + g + 8) * 8);
        na[i][0] = fmaf(v0.x, v0.x, fmaf(v0.y, v0.y, na[i][0]));
        na[i][1] = fmaf(v1.x, v1.x, fmaf(v1.y, v1.y, na[i][1]));
        // a0 (row g, k t)  a1 (row g+8, k t)  a2 (row g, k t+4)  a3 (row g+8, k t+4)
        const float x[4] = {v0.x, v1.x, v0.y, v1.y};
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
    __syncwarp();   // every lane is done with this slot before the next iteration refills it
  }
  cp_async_wait<0>();
  __syncthreads();  // all rings idle: `part` may overwrite them
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
  // dist[p][q][w][tq][ts] = 1 - cos_sim  (myRes.py:756-765: x.y / (|x||y| + 0.01)).  A warp per (query, class) pair,
  // lanes walk the pair's T*T table in storage order; (tq, ts) advance incrementally (no division per element).
  float* dp = dist + ((long long)p * Q + q0) * W * T * T;
  const int TT = T * T, step_q = 32 / T, step_s = 32 % T;
  for (int pair = warp; pair < nq * W; pair += NW) {
    const int q = pair / W, w = pair - q * W;
    int tq = lane / T, ts = lane - tq * T;
    for (int c = lane; c < TT; c += 32) {
      const int m = q * T + tq, n = w * T + ts;
      float s = part[m * NP + n];
#pragma unroll
      for (int k = 1; k < KG; ++k) s += part[k * MP * NP + m * NP + n];
      dp[pair * TT + c] = 1.f - s / (an[m] * bn[n] + 0.01f);
      tq += step_q; ts += step_s;
      if (ts >= T) { ts -= T; ++tq; }
    }
  }
}

// One warp runs 32/(T+2) DPs; a CTA of 8 warps owns SL = 8*per_warp slots.  With two directions, slot s and slot
// s + SL/2 are the two directions of the same (problem, query, class), so their sum is formed inside the CTA.
__global__ void __launch_bounds__(256)
otam_dp_kernel(const float* __restrict__ dist, long long n_pairs, int T, int single_direct, float alpha, float beta,
               float* __restrict__ out, int dp_log) {
  extern __shared__ __align__(16) float sm_dp[];   // [pairs per CTA][T*T] distance tables, then [SL] results
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = otam_dps_per_warp(T), SL = 8 * per_warp;
  const int ndir = single_direct ? 1 : 2, pairs_per_cta = SL / ndir;
  const long long pair0 = (long long)blockIdx.x * pairs_per_cta;
  const int n_here = (int)min((long long)pairs_per_cta, n_pairs - pair0);
  float* res = sm_dp + pairs_per_cta * T * T;
  // T is even (launcher), so every table is a whole number of float4 and 16-byte aligned
  const float4* src = reinterpret_cast<const float4*>(dist + pair0 * T * T);
  const bool exp_mode = otam_exp_mode(T, dp_log);   // the exponentials are taken here, in parallel, off the wavefront's chain
  for (int i = threadIdx.x; i < n_here * T * T / 4; i += blockDim.x) {
    float4 v = __ldg(src + i);
    v.x = otam_table_value(v.x, exp_mode); v.y = otam_table_value(v.y, exp_mode);
    v.z = otam_table_value(v.z, exp_mode); v.w = otam_table_value(v.w, exp_mode);
    reinterpret_cast<float4*>(sm_dp)[i] = v;
  }
  __syncthreads();
  const int seg = lane / (T + 2), m = lane % (T + 2);
  const int slot = warp * per_warp + seg;
  const int pair = slot % pairs_per_cta, dir = slot / pairs_per_cta;
  const bool valid = seg < per_warp && pair < n_here;
  const float r = otam_wavefront_table(T, m, valid, sm_dp + (valid ? pair : 0) * T * T, dir, exp_mode);
  if (valid && m == T + 1) res[slot] = r;
  __syncthreads();
  for (int i = threadIdx.x; i < n_here; i += blockDim.x) {
    const float r2 = res[i] + (single_direct ? 0.f : res[i + pairs_per_cta]);
    float* o = out + pair0 + i;
    *o = (beta != 0.f ? beta * (*o) : 0.f) + alpha * r2;
  }
}

// Scratch for the distance tensor: a stream-ordered pool owned by the library (no device-wide sync, reuses memory).
cudaError_t scratch_pool(cudaMemPool_t* pool) {
  static cudaMemPool_t the_pool = nullptr;
  if (the_pool == nullptr) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    e = cudaMemPoolCreate(&the_pool, &props);
    if (e != cudaSuccess) return e;
    unsigned long long keep = ~0ull;   // never hand memory back to the driver between calls
    e = cudaMemPoolSetAttribute(the_pool, cudaMemPoolAttrReleaseThreshold, &keep);
    if (e != cudaSuccess) return e;
  }
  *pool = the_pool;
  return cudaSuccess;
}

template <int MT, int NT, int MG, int NG, int KG>
int launch(cudaStream_t st, const float* sup, long long s_p, long long s_w, long long s_t, const float* tgt,
           long long t_p, long long t_q, long long t_t, int P, int W, int Q, int QG, int T, int D, float* dist) {
  constexpr int MP = MG * MT * 16, NP = NG * NT * 8, R = MT * 16 + NT * 8;
  constexpr int NW = MG * NG * KG;
  constexpr int RING = NW * NS * 2 * R * 8, PART = KG * MP * NP;
  constexpr size_t smem = (size_t)((RING > PART ? RING : PART) + KG * MP + KG * NP) * sizeof(float);
  static_assert(smem * (16 / NW) <= 226 * 1024, "16 warps per SM");
  if ((D / KG) % STAGE_K != 0) return -3;
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(cos_dist_mma_kernel<MT, NT, MG, NG, KG>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((Q + QG - 1) / QG, P);
  cos_dist_mma_kernel<MT, NT, MG, NG, KG><<<grid, 32 * NW, smem, st>>>(sup, s_p, s_w, s_t, tgt, t_p, t_q, t_t, W, Q, QG, T,
                                                                  D, dist);
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
  if (T < 2 || T > 30 || (T & 1) || D % 128 != 0) return -3;
  if (((s_p | s_w | s_t | t_p | t_q | t_t) & 3) != 0) return -3;                      // 16-byte cp.async pieces
  if ((reinterpret_cast<uintptr_t>(sup) | reinterpret_cast<uintptr_t>(tgt)) & 15) return -3;
  const long long lim = 1LL << 30;   // within-problem row offsets are 32-bit
  if (s_w < 0 || s_t < 0 || t_q < 0 || t_t < 0 || W * s_w + T * s_t >= lim || Q * t_q + T * t_t >= lim) return -3;
  const int WT = W * T;
  // a CTA per problem when the problems alone fill the GPU twice over, else a CTA per (problem, query)
  const int QG = (P >= 2 * sms && Q * T <= 96) ? Q : 1;
  const int QT = QG * T;
  if (WT > 80 || QT > 96) return -3;

  cudaMemPool_t pool;
  cudaError_t e = scratch_pool(&pool);
  if (e != cudaSuccess) return (int)e;
  const long long n_pairs = (long long)P * Q * W;
  float* dist = nullptr;
  e = cudaMallocFromPoolAsync((void**)&dist, (size_t)n_pairs * T * T * sizeof(float), pool, st);
  if (e != cudaSuccess) return (int)e;
  int r = -3;
#define SPM_OTAM_MMA(MT, NT, MG, NG, KG) \
  r = launch<MT, NT, MG, NG, KG>(st, sup, s_p, s_w, s_t, tgt, t_p, t_q, t_t, P, W, Q, QG, T, D, dist)
  if (WT <= 40) {
    if (QT <= 16) SPM_OTAM_MMA(1, 5, 1, 1, 8);
    else if (QT <= 48) SPM_OTAM_MMA(3, 5, 1, 1, 4);   // K slices of 1, 2, 4, 8 warps measured within 5 % of each other
    else SPM_OTAM_MMA(3, 5, 2, 1, 4);
  } else {
    if (QT <= 16) SPM_OTAM_MMA(1, 5, 1, 2, 4);
    else if (QT <= 48) SPM_OTAM_MMA(3, 5, 1, 2, 4);
    else SPM_OTAM_MMA(3, 5, 2, 2, 2);
  }
#undef SPM_OTAM_MMA
  if (r == 0) {
    const int per_warp = 32 / (T + 2), ndir = single_direct ? 1 : 2, pairs_per_cta = 8 * per_warp / ndir;
    const size_t smem = (size_t)(pairs_per_cta * T * T + 8 * per_warp) * sizeof(float);
    const long long ctas = (n_pairs + pairs_per_cta - 1) / pairs_per_cta;
    otam_dp_kernel<<<(unsigned)ctas, 256, smem, st>>>(dist, n_pairs, T, single_direct, alpha, beta, out,
                                                      otam_dp_force_log());
    r = (int)cudaGetLastError();
    count_launch();
  }
  cudaFreeAsync(dist, st);
  return r;
}

}  // namespace spm
