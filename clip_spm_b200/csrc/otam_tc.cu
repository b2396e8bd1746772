// Metric tail at batch scale on the 5th-generation tensor cores (r02): cos_sim (models/myRes.py:756-765) + bidirectional
// OTAM (models/myRes.py:821-855, models/model_clipspm.py:348-362) for P >= 2 x #SM independent problems whose rows stack
// three to a 128-row tile (3*Q*T <= 128, the headline 5-way shapes with T = 8).
//
// Why another formulation.  The mma.sync kernels (otam_mma.cu, otam_fused.cu) are bound by the legacy tensor path: 2880
// HMMA.1688 per problem at one per 8 cycles per scheduler plus the per-fragment hi/lo splits that do not overlap with them
// (~10 k cycles per problem against an HBM share of 7 k, profiles/r02_ncu_otam_fused_notes.txt).  tcgen05 takes the
// products off the warps altogether:
//   warp 0       TMA producer: per 32-column chunk the query rows of up to three consecutive problems (A tile, 3*Q*T rows)
//                and their class rows (B tile, 3*W*T rows) land in a SWIZZLE_128B stage -- the K-major tf32 operand layout
//   warps 2-9    converters, one thread per row: lo = x - (x & 0xffffe000) into the stage's second tile (hi is the landed
//                operand itself: kind::tf32 truncates the low mantissa bits, measured), squared norms accumulated in a
//                register on the way (fp32, exact operands)
//   warp 1       one thread issues, per 8-column step, three tcgen05.mma kind::tf32 (128 x 128 x 8) on the stacked tiles:
//                hi*hi -> D_main, lo*hi + hi*lo -> D_corr in tensor memory (3xTF32: fp32-accurate, the distances are
//                1 - cos of nearly parallel frames).  All 3 x 3 problem blocks are computed and only the diagonal ones
//                used: one 128-wide MMA reads 8 KB of shared memory where three 48-wide ones read 16.5 KB (measured 241 ->
//                195 us at P = 4000); the tensor pipe has the headroom.
//                Two accumulators because the tensor core's fp32 adder truncates: 192 accumulations into one large sum
//                left a biased 1e-5 relative error on near-parallel videos; the 64 of hi*hi alone stay below 4e-6, and
//                the correction terms are 2^-11 of that magnitude
//   warps 12-15  epilogue: tcgen05.ld of each row's own problem, 1 - dot / (|q||s| + 0.01), stored as exp(-d / lambda)
//                into the [Q*W][T][T] tables of a double-buffered smem slot
//   the rest     OTAM wavefronts (otam_dp.cuh, exponent domain, three interleaved per lane) of unit i while unit i + 1 is
//                being multiplied; out[p,q,w] = beta*out + alpha*(dir0 + dir1)
// Variants measured and dropped (r02, P = 4000, this version 182-185 us): the A operands through tensor memory (tcgen05.st by
// the converters, no A reads from shared memory, a fourth ring stage instead: 189 us -- shared-memory bandwidth was not the
// limit after all); two 2-D boxes of three problems per stage instead of six 4-D ones (189 us: nor is the TMA instruction
// count); bursts of L2 prefetches a unit ahead (217 us).  With conversion AND MMAs switched off the ring alone streams the
// same bytes in 150-155 us (0.65-0.67 of the HBM roofline), L2-resident or not: 120 KB in flight per SM over a
// TMA -> converters -> MMA -> commit -> producer cycle of ~4 us is what bounds this design; the products add 25 % on top.
// Nothing but the operands (read once, by TMA) and the Q*W results touches HBM.  A CTA owns a contiguous range of
// problems (P / grid or one more), walked in units of three, the last unit shorter: every SM gets the same number of
// problems to within one.  Roofline: algorithmic bytes (Q+W)*T*D*4 per problem against the measured HBM copy bandwidth.
#include <cstdlib>

#include "gemm.cuh"
#include "head_kernels.cuh"
#include "otam_dp.cuh"
#include "profile.cuh"
#include "ptx.cuh"

namespace spm {

using namespace otam_dp;

namespace {

constexpr int C_STAGES = 3;
constexpr int C_KC = 32;                   // fp32 columns per stage: 128-byte rows
constexpr int C_ROWS = 128;                // rows of the stacked tiles the MMA reads
constexpr int C_TILE = 120 * 128;          // bytes reserved per operand tile: 3 x 40 rows (the MMA's rows 120..127 read on
                                           // into the next tile: rows that only feed unused outputs)
constexpr int C_HALF = 2 * C_TILE;         // A tile + B tile
constexpr int C_STAGE = 2 * C_HALF;        // hi (TMA destination, masked in place) + lo
constexpr int C_UNIT = 3;                  // problems per unit
constexpr int C_TAB = 40 * 40;             // floats per problem table (>= Q*W*T*T = Q*T * W*T)
constexpr int C_CONV_WARPS = 8, C_EPI_WARPS = 4, C_DP_WARPS = 10;
constexpr int C_THREADS = 32 * 24;         // warps: 0 producer, 1 MMA, 2-9 converters, 10-11 + 16-23 wavefronts, 12-15 epilogue
constexpr int C_RES = 64;                  // DP results per problem (>= 2*Q*W)

constexpr int OFF_TAB = C_STAGES * C_STAGE;                       // float [2][3][C_TAB]
constexpr int OFF_NRM = OFF_TAB + 2 * C_UNIT * C_TAB * 4;         // float [2][256]
constexpr int OFF_RES = OFF_NRM + 2 * 256 * 4;                    // float [2][3][C_RES]
constexpr int OFF_BAR = OFF_RES + 2 * C_UNIT * C_RES * 4;         // mbarriers
constexpr int N_BARS = 3 * C_STAGES + 8;
constexpr int OFF_TMEM = OFF_BAR + N_BARS * 8;                    // uint32 TMEM base address
constexpr int C_SMEM_BYTES = OFF_TMEM + 16 + 1024;                // + alignment slack
static_assert(C_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(C_TILE % 1024 == 0, "SWIZZLE_128B atoms are 1024 bytes");

__device__ __forceinline__ void named_bar_tc(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d_tc(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d_tc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                               int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(C_THREADS, 1)
otam_tc_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmT, int P, int W, int Q, int T,
               int D, int single_direct, float alpha, float beta, float* __restrict__ out, int dp_log, int mask_hi, int pf_on) {
  extern __shared__ uint8_t smem_raw_tc[];
  uint8_t* smem = smem_raw_tc + ((1024u - (smem_u32(smem_raw_tc) & 1023u)) & 1023u);   // offset form: stays a shared pointer
  float* tab = reinterpret_cast<float*>(smem + OFF_TAB);
  float* nrm = reinterpret_cast<float*>(smem + OFF_NRM);
  float* res = reinterpret_cast<float*>(smem + OFF_RES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // TMA data landed                      (tx bytes)
  uint64_t* conv = full + C_STAGES;                               // hi masked, lo written               (8 converter warps)
  uint64_t* mdone = conv + C_STAGES;                              // the stage's MMAs have read it       (tcgen05.commit)
  uint64_t* dfull = mdone + C_STAGES;                             // [2] unit's accumulators complete    (tcgen05.commit)
  uint64_t* dempty = dfull + 2;                                   // [2] accumulators + norms read       (128 epilogue threads)
  uint64_t* nfull = dempty + 2;                                   // [2] unit's norms written            (256 converter threads)
  uint64_t* tfull = nfull + 2;                                    // [2] unit's tables written           (128 epilogue threads)
  uint64_t* tempty = tfull + 2;                                   // [2] tables + results consumed       (wavefront threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int QT = Q * T, WT = W * T, NC = D / C_KC;
  const int nA = C_UNIT * QT;                       // rows of the stacked query tile; B tile starts right after (nA % 8 == 0)
  const bool exp_mode = otam_exp_mode(T, dp_log);
  // this CTA's problems: a contiguous range, sizes differ by at most one across the grid
  const int base_n = P / (int)gridDim.x, rem_n = P % (int)gridDim.x;
  const int my_n = base_n + ((int)blockIdx.x < rem_n ? 1 : 0);
  const int my_p0 = (int)blockIdx.x * base_n + min((int)blockIdx.x, rem_n);
  const int n_units = (my_n + C_UNIT - 1) / C_UNIT;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmT);
    for (int s = 0; s < C_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], C_CONV_WARPS); mbar_init(&mdone[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&dfull[b], 1);
      mbar_init(&dempty[b], 32 * C_EPI_WARPS);
      mbar_init(&nfull[b], 32 * C_CONV_WARPS);
      mbar_init(&tfull[b], 32 * C_EPI_WARPS);
      mbar_init(&tempty[b], 32 * C_DP_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      long long n = 0;
      // The ring's boxes are 128-byte row segments 2-4 KB apart, spread over the life of a unit: fetched straight from DRAM
      // each opens a page for 128 bytes.  While unit u streams, every problem of unit u + 1 is requested as ONE burst of L2
      // prefetches (all 16 segments of its rows together); the ring's loads then hit L2.
      const int pf_every = (NC + C_UNIT - 1) / C_UNIT;
      if (pf_on)
        for (int j = 0; j < min(C_UNIT, my_n); ++j)
          for (int cc = 0; cc < NC; ++cc) {
            tma_prefetch_4d_tc(&tmT, cc * C_KC, 0, 0, my_p0 + j);
            tma_prefetch_4d_tc(&tmS, cc * C_KC, 0, 0, my_p0 + j);
          }
      for (int u = 0; u < n_units; ++u) {
        const int p0 = my_p0 + u * C_UNIT, cnt = min(C_UNIT, my_n - u * C_UNIT);
        const int cnt_next = min(C_UNIT, my_n - (u + 1) * C_UNIT);   // <= 0 after the last unit
        for (int c = 0; c < NC; ++c, ++n) {
          if (pf_on && c % pf_every == 0 && c / pf_every < cnt_next) {
            const int pp = p0 + C_UNIT + c / pf_every;
            for (int cc = 0; cc < NC; ++cc) {
              tma_prefetch_4d_tc(&tmT, cc * C_KC, 0, 0, pp);
              tma_prefetch_4d_tc(&tmS, cc * C_KC, 0, 0, pp);
            }
          }
          const int s = (int)(n % C_STAGES);
          mbar_wait(&mdone[s], (uint32_t)(((n / C_STAGES) & 1) ^ 1));
          mbar_expect_tx(&full[s], (uint32_t)(cnt * (QT + WT)) * 128u);
          uint8_t* st = smem + s * C_STAGE;
          for (int j = 0; j < cnt; ++j) {
            tma_load_4d_tc(st + j * QT * 128, &tmT, &full[s], c * C_KC, 0, 0, p0 + j);
            tma_load_4d_tc(st + (nA + j * WT) * 128, &tmS, &full[s], c * C_KC, 0, 0, p0 + j);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(2, C_ROWS, C_ROWS);
      long long n = 0;
      for (int u = 0; u < n_units; ++u) {
        const int b = u & 1;
        mbar_wait(&dempty[b], (uint32_t)(((u >> 1) & 1) ^ 1));
        tc_fence_after_sync();
        for (int c = 0; c < NC; ++c, ++n) {
          const int s = (int)(n % C_STAGES);
          mbar_wait(&conv[s], (uint32_t)((n / C_STAGES) & 1));
          tc_fence_after_sync();
          const uint32_t hi = smem_u32(smem + s * C_STAGE), lo = hi + C_HALF;
          const uint64_t a_hi = umma_desc_k_sw128(hi), a_lo = umma_desc_k_sw128(lo);
          const uint32_t boff = (uint32_t)nA * 128u;
          const uint64_t b_hi = umma_desc_k_sw128(hi + boff), b_lo = umma_desc_k_sw128(lo + boff);
          const uint32_t d_main = tmem_base + (uint32_t)(b * 256), d_corr = d_main + 128u;
#pragma unroll
          for (int k = 0; k < C_KC / 8; ++k) {   // +32 bytes along K inside the swizzle atom == +2 in the address field
            mma_tf32_ss(d_corr, a_lo + 2u * k, b_hi + 2u * k, idesc, (c | k) != 0);
            mma_tf32_ss(d_corr, a_hi + 2u * k, b_lo + 2u * k, idesc, 1u);
            mma_tf32_ss(d_main, a_hi + 2u * k, b_hi + 2u * k, idesc, (c | k) != 0);
          }
          tc_commit(&mdone[s]);
        }
        tc_commit(&dfull[b]);
      }
    }
  } else if (warp >= 2 && warp < 2 + C_CONV_WARPS) {
    // =============================== converters: one thread per stacked row ===============================
    const int r = threadIdx.x - 64;                 // 0 .. 255; rows [0, nA) are query rows, [nA, nA + 3*WT) class rows
    const int nrows = nA + C_UNIT * WT;
    long long n = 0;
    for (int u = 0; u < n_units; ++u) {
      const int cnt = min(C_UNIT, my_n - u * C_UNIT), b = u & 1;
      // rows of problems the unit does not have hold stale data: skipped (their outputs are never read)
      const bool live = r < nrows && (r < nA ? r < cnt * QT : r - nA < cnt * WT);
      float nn = 0.f;
      for (int c = 0; c < NC; ++c, ++n) {
        const int s = (int)(n % C_STAGES);
        mbar_wait(&full[s], (uint32_t)((n / C_STAGES) & 1));
        if (live) {
          uint8_t* row_hi = smem + s * C_STAGE + r * 128;
          uint8_t* row_lo = row_hi + C_HALF;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const int off = ((v ^ (r & 7)) << 4);   // physical 16-byte unit of the row: conflict-free across 8 rows
            float4 x = *reinterpret_cast<float4*>(row_hi + off);
            nn = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, nn))));
            float4 h;
            h.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
            h.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
            h.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
            h.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
            if (mask_hi) *reinterpret_cast<float4*>(row_hi + off) = h;
            *reinterpret_cast<float4*>(row_lo + off) = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
          }
        }
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[s]);
      }
      // the unit's norms -> slot b (freed together with the accumulators two units ago)
      mbar_wait(&dempty[b], (uint32_t)(((u >> 1) & 1) ^ 1));
      nrm[b * 256 + r] = sqrtf(nn);
      mbar_arrive(&nfull[b]);
    }
  } else if (warp >= 12 && warp < 16) {
    // =============================== epilogue: accumulators -> exp(-d / lambda) tables ===============================
    const int q4 = warp & 3, r = q4 * 32 + lane;   // this thread's TMEM lane = stacked query row
    const int jr = r / QT, mloc = r - jr * QT;      // the row's problem and its (query, frame)
    const int qi = mloc / T, tq = mloc - qi * T;
    const int jlo = (q4 * 32) / QT, jhi = min((q4 * 32 + 31) / QT, C_UNIT - 1);
    for (int u = 0; u < n_units; ++u) {
      const int cnt = min(C_UNIT, my_n - u * C_UNIT), b = u & 1;
      const uint32_t ph = (uint32_t)((u >> 1) & 1);
      mbar_wait(&dfull[b], ph);
      mbar_wait(&nfull[b], ph);
      mbar_wait(&tempty[b], ph ^ 1u);
      tc_fence_after_sync();
      const float* nb = nrm + b * 256;
      const float nq = nb[r];
      // the warp's 32 rows belong to problems jlo..jhi: their column blocks are loaded by the whole warp (tcgen05.ld is
      // warp-collective), each thread keeps the block of its own problem
      for (int j = jlo; j <= jhi; ++j) {
        if (j >= cnt) break;                        // warp-uniform
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(b * 256 + j * WT);
        float* tb = tab + (b * C_UNIT + j) * C_TAB;
        for (int g = 0; g * 8 < WT; ++g) {
          uint32_t vm[8], vc[8];
          tmem_ld_32x32b_x8(taddr + (uint32_t)(g * 8), vm);
          tmem_ld_32x32b_x8(taddr + 128u + (uint32_t)(g * 8), vc);
          tmem_ld_wait();
          if (j == jr) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int col = g * 8 + e;             // < WT: WT % 8 == 0
              const int w = col / T, ts = col - w * T;
              const float dot = __uint_as_float(vm[e]) + __uint_as_float(vc[e]);
              const float dval = 1.f - dot / (nq * nb[nA + j * WT + col] + 0.01f);
              tb[((qi * W + w) * T + tq) * T + ts] = otam_table_value(dval, exp_mode);
            }
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&dempty[b]);
      mbar_arrive(&tfull[b]);
    }
  } else {
    // =============================== OTAM wavefronts ===============================
    const int dwarp = warp < 12 ? warp - 10 : warp - 16 + 2, tid_d = dwarp * 32 + lane;
    const int npairs = Q * W, ndir = single_direct ? 1 : 2, ndp = npairs * ndir;
    const int per_warp = otam_dps_per_warp(T), seg = lane / (T + 2), m = lane % (T + 2);
    const int pass = C_DP_WARPS * per_warp;
    constexpr int NW = 3;
    for (int u = 0; u < n_units; ++u) {
      const int p0 = my_p0 + u * C_UNIT, cnt = min(C_UNIT, my_n - u * C_UNIT), b = u & 1;
      mbar_wait(&tfull[b], (uint32_t)((u >> 1) & 1));
      const float* tb = tab + b * C_UNIT * C_TAB;
      float* rb = res + b * C_UNIT * C_RES;
      const int total = cnt * ndp;
      for (int base = dwarp * per_warp; base < total; base += NW * pass) {
        int slot[NW], dir[NW], jj[NW];
        bool valid[NW];
        const float* dw[NW];
#pragma unroll
        for (int x = 0; x < NW; ++x) {
          slot[x] = base + x * pass + seg;
          valid[x] = seg < per_warp && slot[x] < total;
          jj[x] = valid[x] ? slot[x] / ndp : 0;
          const int sl = valid[x] ? slot[x] - jj[x] * ndp : 0;
          dir[x] = sl / npairs;
          dw[x] = tb + jj[x] * C_TAB + (sl - dir[x] * npairs) * T * T;
          slot[x] = jj[x] * C_RES + sl;
        }
        float rr[NW];
        if (exp_mode) {
          otam_wavefront_exp_pre_n<NW>(T, m, valid, dw, dir, rr);
        } else {
#pragma unroll
          for (int x = 0; x < NW; ++x) rr[x] = otam_wavefront(T, m, valid[x], dw[x], dir[x]);
        }
#pragma unroll
        for (int x = 0; x < NW; ++x)
          if (valid[x] && m == T + 1) rb[slot[x]] = rr[x];
      }
      named_bar_tc(2, 32 * C_DP_WARPS);
      for (int i = tid_d; i < cnt * npairs; i += 32 * C_DP_WARPS) {
        const int j = i / npairs, pr = i - j * npairs;
        const float r2 = rb[j * C_RES + pr] + (single_direct ? 0.f : rb[j * C_RES + pr + npairs]);
        float* o = out + (long long)(p0 + j) * npairs + pr;
        *o = (beta != 0.f ? beta * (*o) : 0.f) + alpha * r2;
      }
      mbar_arrive(&tempty[b]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// Returns -3 when the shape is outside this kernel's envelope (the caller falls back to the mma.sync kernels).
int k_otam_tc(cudaStream_t st, const float* sup, long long s_p, long long s_w, long long s_t, const float* tgt, long long t_p,
              long long t_q, long long t_t, int P, int W, int Q, int T, int D, int single_direct, float alpha, float beta,
              float* out) {
  static const int sms = [] {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  static const bool enabled = [] { const char* e = getenv("SPM_OTAM_TC"); return e == nullptr || atoi(e) != 0; }();
  // Crossover measured r02 (tools/time_head_kernels.py, 5 x 5 x 8, D = 512): P = 1000 -> 70 us here against 64 us for the
  // persistent mma.sync kernel (units of three problems drain slowly: the last unit's epilogue and wavefronts are exposed);
  // P = 4000 -> 182 us against 211 us.  SPM_OTAM_TC_MINP overrides the threshold (tests run this kernel at P = 300).
  static const int min_p = [] { const char* e = getenv("SPM_OTAM_TC_MINP"); return e != nullptr ? atoi(e) : 0; }();
  if (!enabled || P < (min_p > 0 ? min_p : 12 * sms)) return -3;
  // SPM_OTAM_TC_MASK=0: leave hi unmasked in shared memory (test of how kind::tf32 reads the low mantissa bits)
  // kind::tf32 TRUNCATES the low 13 mantissa bits of its fp32 operands (measured r02: identical results with hi masked in
  // shared memory and with the raw operand), so hi needs no write-back; SPM_OTAM_TC_MASK=1 masks it anyway
  static const int mask_hi = [] { const char* e = getenv("SPM_OTAM_TC_MASK"); return (e != nullptr && atoi(e) != 0) ? 1 : 0; }();
  // opt-in: measured slower (P = 1000 / 4000: 83 / 217 us against 70 / 182 us without) -- DRAM page locality is not the limit
  static const int pf_on = [] { const char* e = getenv("SPM_OTAM_TC_PF"); return (e != nullptr && atoi(e) != 0) ? 1 : 0; }();
  const int QT = Q * T, WT = W * T;
  // D <= 512: the tensor core's truncating fp32 adder leaves an error that grows with the number of accumulations (64 per
  // 512 columns into D_main: 5e-6 on a video against itself; 1.0e-5 at D = 1024, over this library's 1e-5 bar)
  if (T < 2 || T > 30 || QT % 8 != 0 || WT % 8 != 0 || QT > 40 || WT > 40 || D % C_KC != 0 || D > 512) return -3;
  if (Q * W * T * T > C_TAB || Q * W * (single_direct ? 1 : 2) > C_RES) return -3;
  if (((s_p | s_w | s_t | t_p | t_q | t_t) & 3) != 0 || s_w <= 0 || s_t <= 0 || t_q <= 0 || t_t <= 0) return -3;
  if ((reinterpret_cast<uintptr_t>(sup) | reinterpret_cast<uintptr_t>(tgt)) & 15) return -3;
  CUtensorMap tmS, tmT;
  {
    const unsigned long long dims[4] = {(unsigned long long)D, (unsigned long long)T, (unsigned long long)W,
                                        (unsigned long long)P};
    const unsigned long long strides[3] = {(unsigned long long)s_t * 4, (unsigned long long)s_w * 4,
                                           (unsigned long long)(P > 1 ? s_p : (long long)W * s_w) * 4};
    const unsigned box[4] = {C_KC, (unsigned)T, (unsigned)W, 1};
    if (make_tensor_map_f32_nd(&tmS, sup, 4, dims, strides, box) != 0) return -3;
  }
  {
    const unsigned long long dims[4] = {(unsigned long long)D, (unsigned long long)T, (unsigned long long)Q,
                                        (unsigned long long)P};
    const unsigned long long strides[3] = {(unsigned long long)t_t * 4, (unsigned long long)t_q * 4,
                                           (unsigned long long)(P > 1 ? t_p : (long long)Q * t_q) * 4};
    const unsigned box[4] = {C_KC, (unsigned)T, (unsigned)Q, 1};
    if (make_tensor_map_f32_nd(&tmT, tgt, 4, dims, strides, box) != 0) return -3;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(otam_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int grid = P < sms ? P : sms;
  otam_tc_kernel<<<grid, C_THREADS, C_SMEM_BYTES, st>>>(tmS, tmT, P, W, Q, T, D, single_direct, alpha, beta, out,
                                                       otam_dp_force_log(), mask_hi, pf_on);
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}

}  // namespace spm
