// Metric tail at batch scale as ONE persistent, warp-specialised kernel (r02): cos_sim (models/myRes.py:756-765) +
// bidirectional OTAM (models/myRes.py:821-855, models/model_clipspm.py:348-362) for P >= 2 x #SM independent problems.
//
// Why a third formulation.  The two-kernel batch path (otam_mma.cu) reads every operand byte once but reaches only
// 0.3-0.4 of the HBM roofline: a CTA per problem runs load -> product -> reduce -> store back to back, so a K slice is
// 4-8 cp.async stages deep and the pipeline ramps up and drains once per problem (ncu: 23 % warps active, tensor pipe
// 40 %).  Here one CTA per SM stays resident and streams ALL its problems through one ring:
//   warp 0      producer: per 16-column chunk two TMA boxes ([Q,T,16] query rows, [W,T,16] class rows; 4-D tensor
//               maps, so arbitrary 16-byte-aligned strides work) into a 30-stage SWIZZLE_64B ring running across problem
//               boundaries.  What counts is the part of the ring that is IN FLIGHT: a stage stays occupied while its
//               consumer warp works on it, so with W product warps only (stages - W) of them hide memory latency -- the
//               first version (8 warps, 10 stages of 32 columns) kept 2 stages = 22 KB per SM in flight and sat at
//               2.5 TB/s (ncu: DRAM 24 %, product warps 24 % of their time in the full-barrier wait); now 26 x 5.5 KB
//   warps 1-4   products: chunk n belongs to warp n % 4 (a K split, one warp per scheduler: a single warp saturates the
//               legacy tensor pipe, tools/micro/mma_rate.cu); 3xTF32 mma.sync m16n8k8 (lo*hi + hi*lo + hi*hi,
//               fp32-accurate: the distances are 1 - cos of nearly parallel frames), fragments read conflict-free from the
//               swizzled rows, squared norms accumulated on the way; per problem the four partial tiles are summed
//               in a fixed order (deterministic) into the [Q*W][T][T] tables of a double-buffered smem slot, stored as
//               exp(-d / lambda) when the exponent-domain wavefront reads them
//   warps 5-7   OTAM wavefronts (otam_dp.cuh) of problem i while the product warps already work on problem i + 1;
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

// 16 warps: warp 0 producer; warps 4-11 products (two per scheduler, so one warp's fragment loads and hi/lo splits
// fill the issue slots between the other's HMMAs: a single warp per scheduler kept the tensor pipe only 61 % busy inside
// its own MMA region); warps 1-3 and 12-15 wavefronts.
constexpr int F_KSPLIT = 4;                            // chunk n belongs to the product-warp PAIR n % 4
constexpr int F_MMA_WARPS = 2 * F_KSPLIT, F_DP_WARPS = 7;
constexpr int F_THREADS = 32 * (1 + F_MMA_WARPS + F_DP_WARPS);
static_assert(F_THREADS == 512, "role layout below assumes sixteen warps");
constexpr int F_MT = 3, F_NT = 5;                      // m16 tiles over query rows, n8 tiles over class rows
constexpr int F_MP = F_MT * 16, F_NP = F_NT * 8;       // 48 x 40
constexpr int F_EPI = (F_MP * F_NP + 32 * F_MMA_WARPS - 1) / (32 * F_MMA_WARPS);   // tile elements per product thread

// Ring geometry: KC columns per stage (32: 128-byte rows, SWIZZLE_128B; 16: 64-byte rows, SWIZZLE_64B), STAGES stages.
// Shared-memory carve-up in bytes from the 1024-aligned base.
template <int KC, int STAGES>
struct FCfg {
  static constexpr int A_BYTES = F_MP * KC * 4, B_BYTES = F_NP * KC * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  static_assert(STAGE_BYTES % 1024 == 0 || (KC == 16 && STAGE_BYTES % 512 == 0), "swizzle atoms: 1024 B (128B) / 512 B (64B)");
  static_assert(A_BYTES % 1024 == 0 || (KC == 16 && A_BYTES % 512 == 0), "swizzle atoms");
  static constexpr int OFF_PART = STAGES * STAGE_BYTES;                       // float [4][48][40]
  static constexpr int OFF_PN = OFF_PART + F_KSPLIT * F_MP * F_NP * 4;         // float [8][88]   partial squared norms
  static constexpr int OFF_NRM = OFF_PN + F_MMA_WARPS * (F_MP + F_NP) * 4;     // float [88]      norms of the current problem
  static constexpr int OFF_DIST = OFF_NRM + (F_MP + F_NP) * 4;                 // float [2][48*40] distance tables
  static constexpr int OFF_RES = OFF_DIST + 2 * F_MP * F_NP * 4;               // float [2][64]   DP results
  static constexpr int OFF_BAR = OFF_RES + 2 * 64 * 4;                         // mbarriers
  static constexpr int SMEM_BYTES = OFF_BAR + (2 * STAGES + 4) * 8 + 1024;     // + alignment slack
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

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

template <int KC, int STAGES>
__global__ void __launch_bounds__(F_THREADS, 1)
otam_fused_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmT, int P, int W, int Q,
                  int T, int D, int single_direct, float alpha, float beta, float* __restrict__ out, int dp_log) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by OFFSET, not through an integer round trip: the pointer keeps its shared-memory provenance, so every access
  // below compiles to LDS / STS (the uintptr_t form made all 250 of them generic LD / ST -- ncu source page, r02)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  using C = FCfg<KC, STAGES>;
  constexpr int F_KC = KC, F_STAGES = STAGES, F_A_BYTES = C::A_BYTES, F_STAGE_BYTES = C::STAGE_BYTES;
  float* part = reinterpret_cast<float*>(smem + C::OFF_PART);
  float* pn = reinterpret_cast<float*>(smem + C::OFF_PN);
  float* nrm = reinterpret_cast<float*>(smem + C::OFF_NRM);
  float* dist = reinterpret_cast<float*>(smem + C::OFF_DIST);
  float* res = reinterpret_cast<float*>(smem + C::OFF_RES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* empty = full + F_STAGES;
  uint64_t* dfull = empty + F_STAGES;
  uint64_t* dempty = dfull + 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int QT = Q * T, WT = W * T, NC = D / F_KC;   // NC % F_MMA_WARPS == 0 (launcher)
  const bool exp_mode = otam_exp_mode(T, dp_log);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmT);
    for (int s = 0; s < F_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); }   // a pair consumes a stage
    for (int b = 0; b < 2; ++b) { mbar_init(&dfull[b], 32 * F_MMA_WARPS); mbar_init(&dempty[b], 32 * F_DP_WARPS); }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(QT + WT) * (uint32_t)(F_KC * 4);
      long long n = 0;
      for (int p = blockIdx.x; p < P; p += gridDim.x) {
        for (int c = 0; c < NC; ++c, ++n) {
          const int s = (int)(n % F_STAGES);
          mbar_wait(&empty[s], (uint32_t)(((n / F_STAGES) & 1) ^ 1));
          mbar_expect_tx(&full[s], bytes);
          uint8_t* st = smem + s * F_STAGE_BYTES;
          tma_load_4d(st, &tmT, &full[s], c * F_KC, 0, 0, p);
          tma_load_4d(st + F_A_BYTES, &tmS, &full[s], c * F_KC, 0, 0, p);
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + F_MMA_WARPS) {
    // =============================== products (3xTF32) ===============================
    // pair mw = warps 4 + mw and 8 + mw (same scheduler): both read the pair's chunks, `half` picks the k8 steps
    const int pw = warp - 4, mw = pw & (F_KSPLIT - 1), half = pw / F_KSPLIT, g = lane >> 2, t = lane & 3;
    const int tid_m = threadIdx.x - 128;   // 0 .. 255 inside the product group
    constexpr int KKH = KC / 16;           // k8 steps of a chunk per warp of the pair
    // where element (m = q*T + tq, n = w*T + ts) of the 48 x 40 product tile goes in the [Q*W][T][T] tables: the same for
    // every problem, so the runtime divisions by T are done once (they were a third of the product warps' time)
    int dst[F_EPI];
#pragma unroll
    for (int k = 0; k < F_EPI; ++k) {
      const int idx = tid_m + k * 32 * F_MMA_WARPS, m = idx / F_NP, nn2 = idx - m * F_NP;
      const int q = m / T, tq = m - q * T, w = nn2 / T, ts = nn2 - w * T;
      dst[k] = (idx < F_MP * F_NP && m < QT && nn2 < WT) ? ((q * W + w) * T + tq) * T + ts : -1;
    }
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
      for (int c = mw; c < NC; c += F_KSPLIT) {
        const long long nn = n + c;
        const int s = (int)(nn % F_STAGES);
        mbar_wait(&full[s], (uint32_t)((nn / F_STAGES) & 1));
        const float* sa = reinterpret_cast<const float*>(smem + s * F_STAGE_BYTES);   // [48 rows][16 floats], swizzled
        const float* sb = sa + F_A_BYTES / 4;                                          // [40 rows][16 floats]
#pragma unroll
        for (int kk = half * KKH; kk < (half + 1) * KKH; ++kk) {
          // element (row r, column kk*8 + x) of a swizzled row sits in 16-byte unit ((kk*2 + x/4) ^ sw(r)); every fragment
          // row of this lane has r & 7 == g, so (128-byte rows) the eight rows start 8 units apart modulo the XOR, or
          // (64-byte rows) odd rows sit 16 banks further: either way the 32 lanes of a load hit 32 different banks
          const int sw = KC == 32 ? g : (g >> 1) & 3;   // SWIZZLE_128B: unit ^= r & 7;  SWIZZLE_64B: unit ^= (r >> 1) & 3
          const int o0 = (((2 * kk) ^ sw) << 2) + t, o1 = (((2 * kk + 1) ^ sw) << 2) + t;
          // x = hi + lo with hi = x truncated to tf32 (one LOP3; masked explicitly so the split does not depend on how the
          // tensor core treats the low mantissa bits of a raw fp32 operand) and lo = x - hi (|lo| < 2^-10 |x|, exact)
          uint32_t bx[F_NT][2], bl[F_NT][2];
#pragma unroll
          for (int j = 0; j < F_NT; ++j) {
            const float* row = sb + (j * 8 + g) * F_KC;
            const float x0 = row[o0], x1 = row[o1];
            nb[j] = fmaf(x0, x0, fmaf(x1, x1, nb[j]));
            bx[j][0] = __float_as_uint(x0) & 0xffffe000u; bl[j][0] = __float_as_uint(x0 - __uint_as_float(bx[j][0]));
            bx[j][1] = __float_as_uint(x1) & 0xffffe000u; bl[j][1] = __float_as_uint(x1 - __uint_as_float(bx[j][1]));
          }
          uint32_t ax[F_MT][4], al[F_MT][4];
#pragma unroll
          for (int i = 0; i < F_MT; ++i) {
            const float* r0 = sa + (i * 16 + g) * F_KC;
            const float* r1 = r0 + 8 * F_KC;
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
      // ---- the pair's partial tile -> shared memory: the second warp stores, the first adds its own and stores the sum
      float* pk = part + mw * F_MP * F_NP;
      if (half == 1) {
#pragma unroll
        for (int i = 0; i < F_MT; ++i)
#pragma unroll
          for (int j = 0; j < F_NT; ++j) {
            const int r = i * 16 + g, cc = j * 8 + 2 * t;
            *reinterpret_cast<float2*>(pk + r * F_NP + cc) = make_float2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<float2*>(pk + (r + 8) * F_NP + cc) = make_float2(acc[i][j][2], acc[i][j][3]);
          }
      }
      named_bar(3 + mw, 64);
      if (half == 0) {
#pragma unroll
        for (int i = 0; i < F_MT; ++i)
#pragma unroll
          for (int j = 0; j < F_NT; ++j) {
            const int r = i * 16 + g, cc = j * 8 + 2 * t;
            float2* p0 = reinterpret_cast<float2*>(pk + r * F_NP + cc);
            float2* p1 = reinterpret_cast<float2*>(pk + (r + 8) * F_NP + cc);
            const float2 o0v = *p0, o1v = *p1;
            *p0 = make_float2(acc[i][j][0] + o0v.x, acc[i][j][1] + o0v.y);
            *p1 = make_float2(acc[i][j][2] + o1v.x, acc[i][j][3] + o1v.y);
          }
      }
      float* pnw = pn + pw * (F_MP + F_NP);
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
#pragma unroll
      for (int k = 0; k < F_EPI; ++k) {
        if (dst[k] < 0) continue;
        const int idx = tid_m + k * 32 * F_MMA_WARPS, m = idx / F_NP, nn2 = idx - m * F_NP;
        float s = part[idx];
#pragma unroll
        for (int kw = 1; kw < F_KSPLIT; ++kw) s += part[kw * F_MP * F_NP + idx];
        // myRes.py:756-765: x.y / (|x||y| + 0.01); stored as exp(-d / lambda) for the exponent-domain wavefront
        db[dst[k]] = otam_table_value(1.f - s / (nrm[m] * nrm[F_MP + nn2] + 0.01f), exp_mode);
      }
      mbar_arrive(&dfull[b]);
      named_bar(1, 32 * F_MMA_WARPS);   // `part` / `pn` / `nrm` may be overwritten by the next problem
    }
  } else {
    // =============================== OTAM wavefronts ===============================
    const int dwarp = warp < 4 ? warp - 1 : warp - (4 + F_MMA_WARPS) + 3, tid_d = dwarp * 32 + lane;
    const int npairs = Q * W, ndir = single_direct ? 1 : 2, ndp = npairs * ndir;
    const int per_warp = otam_dps_per_warp(T), seg = lane / (T + 2), m = lane % (T + 2);
    const int pass = F_DP_WARPS * per_warp;   // wavefronts the wavefront warps run side by side
    int it = 0;
    for (int p = blockIdx.x; p < P; p += gridDim.x, ++it) {
      const int b = it & 1;
      mbar_wait(&dfull[b], (uint32_t)((it >> 1) & 1));
      const float* db = dist + b * F_MP * F_NP;
      float* rb = res + b * 64;
      constexpr int NW = 3;   // wavefronts interleaved per lane: 7 warps x 3 segments x 3 = 63 >= the 50 DPs of a 5 x 5 problem
      for (int base = dwarp * per_warp; base < ndp; base += NW * pass) {
        int slot[NW], dir[NW];
        bool valid[NW];
        const float* dw[NW];
#pragma unroll
        for (int u = 0; u < NW; ++u) {
          slot[u] = base + u * pass + seg;
          valid[u] = seg < per_warp && slot[u] < ndp;
          dir[u] = valid[u] ? slot[u] / npairs : 0;
          dw[u] = db + (valid[u] ? slot[u] - dir[u] * npairs : 0) * T * T;
        }
        float r[NW];
        if (exp_mode) {
          otam_wavefront_exp_pre_n<NW>(T, m, valid, dw, dir, r);
        } else {
#pragma unroll
          for (int u = 0; u < NW; ++u) r[u] = otam_wavefront(T, m, valid[u], dw[u], dir[u]);
        }
#pragma unroll
        for (int u = 0; u < NW; ++u)
          if (valid[u] && m == T + 1) rb[slot[u]] = r[u];
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
  // Default for batch scale since its r02 rework (SPM_OTAM_FUSED=0 keeps the two-kernel path).  tools/time_head_kernels.py,
  // P = 1000 / 4000 / 1000 at D = 1024: 64 / 211 / 92 us against 74 / 222 / 109 us for the two kernels.  What the ncu source
  // pages showed on the way (profiles/r02_ncu_otam_fused_notes.txt): every shared-memory access was a GENERIC LD/ST because
  // the buffer was aligned through uintptr_t; the table index arithmetic (three runtime divisions per element) was a third of
  // the product warps' time; the log-domain wavefront ran at ~270 cycles per diagonal step on its own scheduler.  What is
  // left: 2880 HMMA.1688 per problem at one per 8 cycles per scheduler (tools/micro/mma_rate.cu) = 5.8 k cycles, and the
  // hi/lo splits and norm FMAs that feed them do not overlap with it (~10 k cycles per problem in the MMA region with one
  // or with two product warps per scheduler alike) against an HBM share of 7 k cycles.
  static const bool enabled = [] { const char* e = getenv("SPM_OTAM_FUSED"); return e == nullptr || atoi(e) != 0; }();
  if (!enabled || P < 2 * sms) return -3;
  if (T < 2 || T > 30 || Q * T > F_MP || W * T > F_NP || Q * W * (single_direct ? 1 : 2) > 64) return -3;
  // ring geometry: SPM_OTAM_KC=16 -> 30 stages of 16 columns (64-byte rows), default 15 stages of 32 columns
  static const int kc = [] { const char* e = getenv("SPM_OTAM_KC"); return (e != nullptr && atoi(e) == 16) ? 16 : 32; }();
  if (D % (kc * F_KSPLIT) != 0) return -3;
  if (((s_p | s_w | s_t | t_p | t_q | t_t) & 3) != 0 || s_w <= 0 || s_t <= 0 || t_q <= 0 || t_t <= 0) return -3;
  if ((reinterpret_cast<uintptr_t>(sup) | reinterpret_cast<uintptr_t>(tgt)) & 15) return -3;
  CUtensorMap tmS, tmT;
  {
    const unsigned long long dims[4] = {(unsigned long long)D, (unsigned long long)T, (unsigned long long)W,
                                        (unsigned long long)P};
    const unsigned long long strides[3] = {(unsigned long long)s_t * 4, (unsigned long long)s_w * 4,
                                           (unsigned long long)(P > 1 ? s_p : (long long)W * s_w) * 4};
    const unsigned box[4] = {(unsigned)kc, (unsigned)T, (unsigned)W, 1};
    if (make_tensor_map_f32_nd(&tmS, sup, 4, dims, strides, box) != 0) return -3;
  }
  {
    const unsigned long long dims[4] = {(unsigned long long)D, (unsigned long long)T, (unsigned long long)Q,
                                        (unsigned long long)P};
    const unsigned long long strides[3] = {(unsigned long long)t_t * 4, (unsigned long long)t_q * 4,
                                           (unsigned long long)(P > 1 ? t_p : (long long)Q * t_q) * 4};
    const unsigned box[4] = {(unsigned)kc, (unsigned)T, (unsigned)Q, 1};
    if (make_tensor_map_f32_nd(&tmT, tgt, 4, dims, strides, box) != 0) return -3;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(otam_fused_kernel<32, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         FCfg<32, 15>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(otam_fused_kernel<16, 30>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               FCfg<16, 30>::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int grid = P < sms ? P : sms;
  if (kc == 32)
    otam_fused_kernel<32, 15><<<grid, F_THREADS, FCfg<32, 15>::SMEM_BYTES, st>>>(tmS, tmT, P, W, Q, T, D, single_direct, alpha,
                                                                              beta, out, otam_dp_force_log());
  else
    otam_fused_kernel<16, 30><<<grid, F_THREADS, FCfg<16, 30>::SMEM_BYTES, st>>>(tmS, tmT, P, W, Q, T, D, single_direct, alpha,
                                                                              beta, out, otam_dp_force_log());
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}

}  // namespace spm
