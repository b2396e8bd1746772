// Metric tail at batch scale as ONE persistent, warp-specialised kernel (r02): cos_sim (models/myRes.py:756-765) +
// bidirectional OTAM (models/myRes.py:821-855, models/model_clipspm.py:348-362) for P >= 2 x #SM independent problems.
//
// Why a third formulation.  The two-kernel batch path (otam_mma.cu) reads every operand byte once but reaches only
// 0.3-0.4 of the HBM roofline: a CTA per problem runs load -> product -> reduce -> store back to back, so a K slice is
// 4-8 cp.async stages deep and the pipeline ramps up and drains once per problem (ncu: 23 % warps active, tensor pipe
// 40 %).  Here one CTA per SM stays resident and streams ALL its problems through one ring:
//   warp 0      producer: per 32-column chunk two TMA boxes ([Q,T,32] query rows, [W,T,32] class rows; 4-D tensor
//               maps, so arbitrary 16-byte-aligned strides work) into a 10-stage SWIZZLE_128B ring -- ~110 KB in
//               flight per SM, running across problem boundaries
//   warps 1-8   products: chunk n belongs to warp n % 8 (a K split); 3xTF32 mma.sync m16n8k8 (lo*hi + hi*lo + hi*hi,
//               fp32-accurate: the distances are 1 - cos of nearly parallel frames), fragments read conflict-free from the
//               swizzled rows, squared norms accumulated on the way; per problem the eight partial tiles are summed
//               in a fixed order (deterministic) into the [Q*W][T][T] distance tables of a double-buffered smem slot
//   warps 9-12  OTAM wavefronts (otam_dp.cuh) of problem i while the product warps already work on problem i + 1;
//               out[p,q,w] = beta*out + alpha*(dir0 + dir1)
// Nothing but the operands (read once, by TMA) and the Q*W results touches HBM; the distance tensor never leaves
// shared memory.  Roofline: algorithmic bytes (Q+W)*T*D*4 per problem against the measured HBM copy bandwidth; the
// legacy tensor pipe needs ~5.8 k cycles per problem per SM (2880 HMMA.1688 at one per 8 clk per sub-partition), the
// HBM share of an SM ~7 k cycles, so the two overlap with little slack -- tools/time_head_kernels.py.
#include <cstdlib>

#include "gemm.cuh"
#include "head_kernels.cuh"
#include "otam_dp.cuh"
#include "profile.cuh"
#include "ptx.cuh"

namespace spm {

using namespace otam_dp;

namespace {

constexpr int F_MMA_WARPS = 8, F_DP_WARPS = 4;
constexpr int F_THREADS = 32 * (1 + F_MMA_WARPS + F_DP_WARPS);
constexpr int F_STAGES = 10;
constexpr int F_MT = 3, F_NT = 5;                      // m16 tiles over query rows, n8 tiles over class rows
constexpr int F_MP = F_MT * 16, F_NP = F_NT * 8;       // 48 x 40
constexpr int F_A_BYTES = F_MP * 128, F_B_BYTES = F_NP * 128, F_STAGE_BYTES = F_A_BYTES + F_B_BYTES;
static_assert(F_STAGE_BYTES % 1024 == 0 && F_A_BYTES % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");

// shared-memory carve-up (bytes from the 1024-aligned base)
constexpr int OFF_PART = F_STAGES * F_STAGE_BYTES;                       // float [8][48][40]
constexpr int OFF_PN = OFF_PART + F_MMA_WARPS * F_MP * F_NP * 4;         // float [8][88]   partial squared norms
constexpr int OFF_NRM = OFF_PN + F_MMA_WARPS * (F_MP + F_NP) * 4;        // float [88]      norms of the current problem
constexpr int OFF_DIST = OFF_NRM + (F_MP + F_NP) * 4;                    // float [2][48*40] distance tables
constexpr int OFF_RES = OFF_DIST + 2 * F_MP * F_NP * 4;                  // float [2][64]   DP results
constexpr int OFF_BAR = OFF_RES + 2 * 64 * 4;                            // mbarriers
constexpr int F_SMEM_BYTES = OFF_BAR + (2 * F_STAGES + 4) * 8 + 1024;    // + alignment slack
static_assert(F_SMEM_BYTES <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(F_THREADS, 1)
otam_fused_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmT, int P, int W, int Q,
                  int T, int D, int single_direct, float alpha, float beta, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* part = reinterpret_cast<float*>(smem + OFF_PART);
  float* pn = reinterpret_cast<float*>(smem + OFF_PN);
  float* nrm = reinterpret_cast<float*>(smem + OFF_NRM);
  float* dist = reinterpret_cast<float*>(smem + OFF_DIST);
  float* res = reinterpret_cast<float*>(smem + OFF_RES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + F_STAGES;
  uint64_t* dfull = empty + F_STAGES;
  uint64_t* dempty = dfull + 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int QT = Q * T, WT = W * T, NC = D / 32;   // NC % F_MMA_WARPS == 0 (launcher)
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmT);
    for (int s = 0; s < F_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&dfull[b], 32 * F_MMA_WARPS); mbar_init(&dempty[b], 32 * F_DP_WARPS); }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(QT + WT) * 128u;
      long long n = 0;
      for (int p = blockIdx.x; p < P; p += gridDim.x) {
        for (int c = 0; c < NC; ++c, ++n) {
          const int s = (int)(n % F_STAGES);
          mbar_wait(&empty[s], (uint32_t)(((n / F_STAGES) & 1) ^ 1));
          mbar_expect_tx(&full[s], bytes);
          uint8_t* st = smem + s * F_STAGE_BYTES;
          tma_load_4d(st, &tmT, &full[s], c * 32, 0, 0, p);
          tma_load_4d(st + F_A_BYTES, &tmS, &full[s], c * 32, 0, 0, p);
        }
      }
    }
  } else if (warp <= F_MMA_WARPS) {
    // =============================== products (3xTF32) ===============================
    const int mw = warp - 1, g = lane >> 2, t = lane & 3;
    const int tid_m = threadIdx.x - 32;   // 0 .. 255 inside the product group
    const int TT = T * T, n_dist = Q * W * TT;
    long long n = 0;
    int it = 0;
    for (int p = blockIdx.x; p < P; p += gridDim.x, ++it) {
      float acc[F_MT][F_NT][4];
      float na[F_MT][2], nb[F_NT];
#pragma unroll
      for (int i = 0; i < F_MT; ++i) {
        na[i][0] = na[i][1] = 0.f;
#pragma unroll
        for (int j = 0; j < F_NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < F_NT; ++j) nb[j] = 0.f;
#pragma unroll 1
      for (int c = mw; c < NC; c += F_MMA_WARPS) {
        const long long nn = n + c;
        const int s = (int)(nn % F_STAGES);
        mbar_wait(&full[s], (uint32_t)((nn / F_STAGES) & 1));
        const float* sa = reinterpret_cast<const float*>(smem + s * F_STAGE_BYTES);   // [48 rows][32 floats], swizzled
        const float* sb = sa + F_A_BYTES / 4;                                          // [40 rows][32 floats]
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          // element (row r, column kk*8 + x) of a SWIZZLE_128B row sits in 16-byte unit ((kk*2 + x/4) ^ (r & 7)); every
          // fragment row of this lane has r & 7 == g, so the 32 lanes of a load hit 32 different banks
          const int o0 = (((2 * kk) ^ g) << 2) + t, o1 = (((2 * kk + 1) ^ g) << 2) + t;
          // x = hi + lo with hi = x truncated to tf32 (one LOP3; masked explicitly so the split does not depend on how the
          // tensor core treats the low mantissa bits of a raw fp32 operand) and lo = x - hi (|lo| < 2^-10 |x|, exact)
          uint32_t bx[F_NT][2], bl[F_NT][2];
#pragma unroll
          for (int j = 0; j < F_NT; ++j) {
            const float* row = sb + (j * 8 + g) * 32;
            const float x0 = row[o0], x1 = row[o1];
            nb[j] = fmaf(x0, x0, fmaf(x1, x1, nb[j]));
            bx[j][0] = __float_as_uint(x0) & 0xffffe000u; bl[j][0] = __float_as_uint(x0 - __uint_as_float(bx[j][0]));
            bx[j][1] = __float_as_uint(x1) & 0xffffe000u; bl[j][1] = __float_as_uint(x1 - __uint_as_float(bx[j][1]));
          }
          uint32_t ax[F_MT][4], al[F_MT][4];
#pragma unroll
          for (int i = 0; i < F_MT; ++i) {
            const float* r0 = sa + (i * 16 + g) * 32;
            const float* r1 = r0 + 8 * 32;
            // a0 (row g, k t)  a1 (row g+8, k t)  a2 (row g, k t+4)  a3 (row g+8, k t+4)
            const float x[4] = {r0[o0], r1[o0], r0[o1], r1[o1]};
            na[i][0] = fmaf(x[0], x[0], fmaf(x[2], x[2], na[i][0]));
            na[i][1] = fmaf(x[1], x[1], fmaf(x[3], x[3], na[i][1]));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              ax[i][e] = __float_as_uint(x[e]) & 0xffffe000u;
              al[i][e] = __float_as_uint(x[e] - __uint_as_float(ax[i][e]));
            }
          }
          // three passes over the 15 accumulator tiles: consecutive MMAs never touch the same accumulator, so the tensor
          // pipe is not serialised on the accumulate dependency (two product warps per scheduler cannot hide it otherwise)
#pragma unroll
          for (int i = 0; i < F_MT; ++i)
#pragma unroll
            for (int j = 0; j < F_NT; ++j) mma_tf32(acc[i][j], al[i], bx[j][0], bx[j][1]);
#pragma unroll
          for (int i = 0; i < F_MT; ++i)
#pragma unroll
            for (int j = 0; j < F_NT; ++j) mma_tf32(acc[i][j], ax[i], bl[j][0], bl[j][1]);
#pragma unroll
          for (int i = 0; i < F_MT; ++i)
#pragma unroll
            for (int j = 0; j < F_NT; ++j) mma_tf32(acc[i][j], ax[i], bx[j][0], bx[j][1]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);   // the stage may be refilled
      }
      n += NC;
      // ---- this warp's partial tile and partial norms -> shared memory
      float* pk = part + mw * F_MP * F_NP;
#pragma unroll
      for (int i = 0; i < F_MT; ++i)
#pragma unroll
        for (int j = 0; j < F_NT; ++j) {
          const int r = i * 16 + g, cc = j * 8 + 2 * t;
          *reinterpret_cast<float2*>(pk + r * F_NP + cc) = make_float2(acc[i][j][0], acc[i][j][1]);
          *reinterpret_cast<float2*>(pk + (r + 8) * F_NP + cc) = make_float2(acc[i][j][2], acc[i][j][3]);
        }
      float* pnw = pn + mw * (F_MP + F_NP);
#pragma unroll
      for (int i = 0; i < F_MT; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v = na[i][h];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          if (t == 0) pnw[i * 16 + g + 8 * h] = v;
        }
#pragma unroll
      for (int j = 0; j < F_NT; ++j) {
        float v = nb[j];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (t == 0) pnw[F_MP + j * 8 + g] = v;
      }
      named_bar(1, 32 * F_MMA_WARPS);
      if (tid_m < F_MP + F_NP) {
        float s = pn[tid_m];
#pragma unroll
        for (int k = 1; k < F_MMA_WARPS; ++k) s += pn[k * (F_MP + F_NP) + tid_m];
        nrm[tid_m] = sqrtf(s);
      }
      named_bar(1, 32 * F_MMA_WARPS);
      // ---- distance tables of this problem -> slot it & 1 (freed by the DP warps two problems ago)
      const int b = it & 1;
      mbar_wait(&dempty[b], (uint32_t)(((it >> 1) & 1) ^ 1));
      float* db = dist + b * F_MP * F_NP;
      for (int e = tid_m; e < n_dist; e += 32 * F_MMA_WARPS) {
        const int pair = e / TT, cidx = e - pair * TT;
        const int q = pair / W, w = pair - q * W, tq = cidx / T, ts = cidx - tq * T;
        const int m = q * T + tq, nn2 = w * T + ts;
        float s = part[m * F_NP + nn2];
#pragma unroll
        for (int k = 1; k < F_MMA_WARPS; ++k) s += part[k * F_MP * F_NP + m * F_NP + nn2];
        db[e] = 1.f - s / (nrm[m] * nrm[F_MP + nn2] + 0.01f);   // myRes.py:756-765: x.y / (|x||y| + 0.01)
      }
      mbar_arrive(&dfull[b]);
      named_bar(1, 32 * F_MMA_WARPS);   // `part` / `pn` / `nrm` may be overwritten by the next problem
    }
  } else {
    // =============================== OTAM wavefronts ===============================
    const int dwarp = warp - 1 - F_MMA_WARPS, tid_d = threadIdx.x - 32 * (1 + F_MMA_WARPS);
    const int npairs = Q * W, ndir = single_direct ? 1 : 2, ndp = npairs * ndir;
    const int per_warp = otam_dps_per_warp(T), seg = lane / (T + 2), m = lane % (T + 2);
    int it = 0;
    for (int p = blockIdx.x; p < P; p += gridDim.x, ++it) {
      const int b = it & 1;
      mbar_wait(&dfull[b], (uint32_t)((it >> 1) & 1));
      const float* db = dist + b * F_MP * F_NP;
      float* rb = res + b * 64;
      for (int base = dwarp * per_warp; base < ndp; base += F_DP_WARPS * per_warp) {
        const int slot = base + seg;
        const bool valid = seg < per_warp && slot < ndp;
        const int dir = valid ? slot / npairs : 0, pair = valid ? slot - dir * npairs : 0;
        const float r = otam_wavefront(T, m, valid, db + pair * T * T, dir);
        if (valid && m == T + 1) rb[slot] = r;
      }
      named_bar(2, 32 * F_DP_WARPS);
      for (int i = tid_d; i < npairs; i += 32 * F_DP_WARPS) {
        const float r2 = rb[i] + (single_direct ? 0.f : rb[i + npairs]);
        float* o = out + (long long)p * npairs + i;
        *o = (beta != 0.f ? beta * (*o) : 0.f) + alpha * r2;
      }
      mbar_arrive(&dempty[b]);   // every DP thread is done with the slot's tables and results
    }
  }
}

}  // namespace

// Returns -3 when the shape is outside this kernel's envelope (the caller falls back to the two-kernel path).
int k_otam_fused(cudaStream_t st, const float* sup, long long s_p, long long s_w, long long s_t, const float* tgt,
                 long long t_p, long long t_q, long long t_t, int P, int W, int Q, int T, int D, int single_direct,
                 float alpha, float beta, float* out) {
  static const int sms = [] {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  // Opt-in (SPM_OTAM_FUSED=1).  Measured r02 (tools/time_head_kernels.py, P = 1000): 88 us against 76 us for the
  // two-kernel path -- correct, every operand byte read once (ncu: 164 MB), but the four wavefront warps run their five
  // rounds of 17 dependent diagonal steps at ~270 cycles per step, ~23 k cycles per problem, where the product warps need
  // ~6 k; hiding that chain takes >= 16 wavefront warps, which the register budget of one CTA per SM (8 product warps at
  // 128 registers) does not leave without setmaxnreg re-balancing.  Kept for that next step; the default stays two kernels.
  static const bool enabled = [] { const char* e = getenv("SPM_OTAM_FUSED"); return e != nullptr && atoi(e) != 0; }();
  if (!enabled || P < 2 * sms) return -3;
  if (T < 2 || T > 30 || Q * T > F_MP || W * T > F_NP || Q * W * (single_direct ? 1 : 2) > 64) return -3;
  if (D % (32 * F_MMA_WARPS) != 0) return -3;
  if (((s_p | s_w | s_t | t_p | t_q | t_t) & 3) != 0 || s_w <= 0 || s_t <= 0 || t_q <= 0 || t_t <= 0) return -3;
  if ((reinterpret_cast<uintptr_t>(sup) | reinterpret_cast<uintptr_t>(tgt)) & 15) return -3;
  CUtensorMap tmS, tmT;
  {
    const unsigned long long dims[4] = {(unsigned long long)D, (unsigned long long)T, (unsigned long long)W,
                                        (unsigned long long)P};
    const unsigned long long strides[3] = {(unsigned long long)s_t * 4, (unsigned long long)s_w * 4,
                                           (unsigned long long)(P > 1 ? s_p : (long long)W * s_w) * 4};
    const unsigned box[4] = {32, (unsigned)T, (unsigned)W, 1};
    if (make_tensor_map_f32_nd(&tmS, sup, 4, dims, strides, box) != 0) return -3;
  }
  {
    const unsigned long long dims[4] = {(unsigned long long)D, (unsigned long long)T, (unsigned long long)Q,
                                        (unsigned long long)P};
    const unsigned long long strides[3] = {(unsigned long long)t_t * 4, (unsigned long long)t_q * 4,
                                           (unsigned long long)(P > 1 ? t_p : (long long)Q * t_q) * 4};
    const unsigned box[4] = {32, (unsigned)T, (unsigned)Q, 1};
    if (make_tensor_map_f32_nd(&tmT, tgt, 4, dims, strides, box) != 0) return -3;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(otam_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int grid = P < sms ? P : sms;
  otam_fused_kernel<<<grid, F_THREADS, F_SMEM_BYTES, st>>>(tmS, tmT, P, W, Q, T, D, single_direct, alpha, beta, out);
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}

}  // namespace spm
