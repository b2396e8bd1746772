// 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM: a cluster of two CTAs on one TPC computes a
// 256 x 256 output tile.  Each CTA stages ITS 128 rows of A and ITS half (128 rows) of B per k-block, so the shared
// memory written by TMA and read by the tensor core per CTA drops from 48 KB to 32 KB per 128x256x64 of work
// (the 1-CTA kernel is bound by that traffic: 96 B/clk in + 96 B/clk out against 128 B/clk of shared memory) and the
// pipeline deepens from 4 to 6 stages.  The leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 (M = 256);
// every CTA's TMEM receives its own 128 accumulator rows and runs the same epilogue as the 1-CTA kernel.
//   producer (both CTAs)  TMA .cta_group::2 loads into the local smem, completion bytes counted on the LEADER's
//                         full barrier (peer bit of the barrier address cleared)
//   MMA (leader)          tcgen05.commit ... multicast::cluster frees the stage in both CTAs / publishes the
//                         accumulator to both epilogues
//   epilogue (both CTAs)  arrive remotely on the leader's tmem-empty barrier (16 warp arrivals)
// bf16 operands only, N % 256 == 0; used for the large frame-encoder GEMMs.
#include "gemm.cuh"
#include <cstdlib>

#include "gemm_epilogue.cuh"
#include "profile.cuh"
#include "ptx.cuh"

namespace spm {

namespace {
// RES: residual boxes arrive by TMA into a ring of two 4 KB tiles per epilogue warp (which double as the staging
// tiles); the operand pipeline gives up one stage for them.
template <bool RES>
struct G2T {
  static constexpr int BM = 128, BN = 256, BK = 64;
  static constexpr int A_BYTES = 128 * 128, B_BYTES = 128 * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = RES ? 5 : 6;
  static constexpr int TMEM_COLS = 512;
  static constexpr int BAR_BYTES = 1024;  // keeps the staging / residual tiles 1024-byte aligned (TMA SWIZZLE_128B)
  static constexpr int EPI_WARPS = 8;
  static constexpr int STG_BUFS = RES ? 2 : 1;
  static constexpr int STG_BYTES_PER_WARP = STG_BUFS * 32 * 128;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + EPI_WARPS * STG_BYTES_PER_WARP + 1024;
};
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_ss_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2cta_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive on a barrier that lives in the leader CTA's shared memory (works from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* local_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(local_bar) & PEER_MASK)
               : "memory");
}
}  // namespace

struct Gemm2Args {
  GemmEpilogue ep;
  int M, N, K;
  int reverse;
  int conv_w2, conv_cblocks;   // EPI == 3: implicit 3x3 convolution taps (see GemmOp), 0 = 1x1 convolution
};

// EPI: 0 = general epilogue; 1 / 2 = bf16 output + bias (+ QuickGELU) with identity rows (gemm_epilogue_tile_bf16_bias);
// 3 = convolution (rn50.cu): the producer addresses the 3x3 taps as row offsets of the zero-bordered image matrix and
// the general epilogue is compiled with its zero-border handling
// LNF: LayerNorm folded into the epilogues (GemmEpilogue::ln_*): RES kernels also emit the bf16 copy and the row statistics,
// EPI 1 / 2 kernels normalise their accumulator rows with them
template <bool RES, int EPI = 0, int LNF = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmR, const Gemm2Args args) {
  using T = G2T<RES>;
  extern __shared__ uint8_t smem_raw2[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + T::STAGES * T::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + T::STAGES;
  uint64_t* tfull_bar = empty_bar + T::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;  // [EPI_WARPS][2], RES only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * T::EPI_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int M = args.M, N = args.N, K = args.K;
  const int num_n = N / T::BN;
  const int num_tiles = ((M + 2 * T::BM - 1) / (2 * T::BM)) * num_n;
  const int num_kb = (K + T::BK - 1) / T::BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < T::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * T::EPI_WARPS);  // epilogue warps of BOTH CTAs (only the leader's copy is used)
    }
    if (RES)
      for (int s = 0; s < 2 * T::EPI_WARPS; ++s) mbar_init(&res_bar[s], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2cta(tmem_slot, T::TMEM_COLS);
  tc_fence_before_sync();
  cluster_sync_all();  // both CTAs' barriers and TMEM exist before any cross-CTA signal
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the prologue above overlapped the predecessor's tail; from here on its results are needed
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (both CTAs) =====================
      int stage = 0;
      uint32_t phase = 0;
      int pf_tile = cluster_id, pf_kb = 0;
      auto prefetch_next = [&]() {
        if (pf_tile < num_tiles) {
          tma_prefetch_l2_2d(&tmA, pf_kb * T::BK,
                             ((args.reverse ? num_tiles - 1 - pf_tile : pf_tile) / num_n) * (2 * T::BM) + (int)rank * T::BM);
          if (++pf_kb == num_kb) { pf_kb = 0; pf_tile += n_clusters; }
        }
      };
      constexpr bool conv = EPI == 3;   // taps re-read an L2-resident activation: no prefetch needed
      if (!conv)
        for (int i = 0; i < GEMM_L2_PREFETCH_KB; ++i) prefetch_next();
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int rt = args.reverse ? num_tiles - 1 - tile : tile;
        const int m0 = (rt / num_n) * (2 * T::BM) + (int)rank * T::BM;
        const int n0 = (rt % num_n) * T::BN + (int)rank * (T::BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!conv) prefetch_next();
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * T::STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * T::STAGE_BYTES);  // bytes of both CTAs land on the leader
          const uint32_t lbar = smem_u32(&full_bar[stage]) & PEER_MASK;
          if (conv && args.conv_cblocks > 0) {
            const int tap = kb / args.conv_cblocks, cb = kb - tap * args.conv_cblocks;
            const int roff = (tap / 3 - 1) * args.conv_w2 + (tap % 3 - 1);  // rows outside the matrix read as zeros
            tma_load_2d_2sm(sa, &tmA, lbar, cb * T::BK, m0 + roff);
          } else
          tma_load_2d_2sm(sa, &tmA, lbar, kb * T::BK, m0);
          tma_load_2d_2sm(sa + T::A_BYTES, &tmB, lbar, kb * T::BK, n0);
          if (++stage == T::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===================== MMA issuer (leader CTA) =====================
      constexpr uint32_t idesc = umma_idesc(1, 2 * T::BM, T::BN);
      int stage = 0;
      uint32_t phase = 0;
      int t = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++t) {
        const int acc = t & 1;
        const uint32_t acc_phase = (t >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * T::BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * T::STAGE_BYTES);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + T::A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_bf16_ss_2cta(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
          tc_commit_2cta_mc(&empty_bar[stage], 3);  // stage free in both CTAs once these MMAs have read it
          if (++stage == T::STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit_2cta_mc(&tfull_bar[acc], 3);  // accumulator complete -> both epilogues
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const GemmEpilogue& ep = args.ep;
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t stg_u =
        smem_u32(smem + T::STAGES * T::STAGE_BYTES + T::BAR_BYTES + (warp - 4) * T::STG_BYTES_PER_WARP);
    int t = 0;
    if (RES) {
      // This warp's column groups, in processing order: tile by tile, groups half, half+2, half+4, half+6 of the
      // tile's eight 32-column groups.  The residual box of group i lands in ring buffer i & 1; it is requested when
      // buffer i & 1 was last released, i.e. one whole group of work before it is needed.
      constexpr int GPT = T::BN / 64;  // groups per tile and warp
      uint64_t* my_bar = res_bar + (warp - 4) * 2;
      int is_tile = cluster_id, is_g = 0;  // issue cursor
      uint32_t n_issued = 0, n_done = 0;
      auto issue = [&]() {
        if (is_tile < num_tiles) {
          if (lane == 0) {
            const int b = n_issued & 1;
            const int rt = args.reverse ? num_tiles - 1 - is_tile : is_tile;
            const int row0 = (rt / num_n) * (2 * T::BM) + (int)rank * T::BM + q * 32;
            const int col0 = (rt % num_n) * T::BN + (half + 2 * is_g) * 32;
            fence_proxy_async_smem();  // the buffer was last touched by generic-proxy loads/stores of this warp
            mbar_expect_tx(&my_bar[b], 32 * 128);
            tma_load_2d(reinterpret_cast<void*>(smem + T::STAGES * T::STAGE_BYTES + T::BAR_BYTES +
                                                (warp - 4) * T::STG_BYTES_PER_WARP + b * 4096),
                        &tmR, &my_bar[b], col0, row0);
          }
          ++n_issued;
          if (++is_g == GPT) { is_g = 0; is_tile += n_clusters; }
        }
      };
      issue();
      issue();
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++t) {
        const int acc = t & 1;
        const uint32_t acc_phase = (t >> 1) & 1;
        const int rt = args.reverse ? num_tiles - 1 - tile : tile;
        const int m_base = (rt / num_n) * (2 * T::BM) + (int)rank * T::BM + q * 32;
        const int n0 = (rt % num_n) * T::BN;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * T::BN);
        float s1 = 0.f, s2 = 0.f;   // LNF: row sums over this warp's 128 columns of the tile
#pragma unroll 1
        for (int gi = 0; gi < GPT; ++gi) {
          const int g = half + 2 * gi;
          const int b = n_done & 1;
          mbar_wait(&my_bar[b], (n_done >> 1) & 1);
          gemm_epilogue_group_restma<(LNF != 0)>(ep, stg_u + (uint32_t)(b * 4096), taddr + (uint32_t)(g * 32), m_base,
                                          n0 + g * 32, M, lane, s1, s2);
          ++n_done;
          if (LNF && gi == GPT - 1 && m_base + lane < M)
            *reinterpret_cast<float2*>(ep.ln_stats_out + ((long long)(m_base + lane) * LN_FOLD_SLOTS +
                                                          (n0 / T::BN) * 2 + half) * 2) = make_float2(s1, s2);
          if (gi == GPT - 1) {  // last TMEM read of this tile is done: hand the accumulator back before refilling
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
          }
          issue();
        }
      }
    } else {
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++t) {
        const int acc = t & 1;
        const uint32_t acc_phase = (t >> 1) & 1;
        const int rt = args.reverse ? num_tiles - 1 - tile : tile;
        const int m_base = (rt / num_n) * (2 * T::BM) + (int)rank * T::BM + q * 32;
        const int n0 = (rt % num_n) * T::BN;
        if (EPI == 3) prefetch_residual_bf16<T::BN>(ep, m_base, n0, M, N, lane, half);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * T::BN);
        if (EPI == 1) gemm_epilogue_tile_bf16_bias<T::BN, ACT_NONE, LNF>(ep, stg_u, taddr, m_base, n0, M, lane, half);
        else if (EPI == 2) gemm_epilogue_tile_bf16_bias<T::BN, ACT_QUICKGELU, LNF>(ep, stg_u, taddr, m_base, n0, M, lane, half);
        else if (EPI == 3) {
          if (conv_epilogue_applies(ep)) {   // (warp-uniform) compile-time specialised convolution epilogue
            if (ep.residual_bf16 != nullptr) gemm_epilogue_tile_conv<T::BN, true>(ep, stg_u, taddr, m_base, n0, M, N, lane, half);
            else gemm_epilogue_tile_conv<T::BN, false>(ep, stg_u, taddr, m_base, n0, M, N, lane, half);
          } else {
            gemm_epilogue_tile<T::BN, true>(ep, stg_u, taddr, m_base, n0, M, N, lane, half);
          }
        }
        else gemm_epilogue_tile<T::BN>(ep, stg_u, taddr, m_base, n0, M, N, lane, half);
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
      }
    }
  }

  tc_fence_before_sync();
  cluster_sync_all();  // nobody frees TMEM / exits while the peer may still signal or read
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc_2cta(tmem_base, T::TMEM_COLS);
  }
}

int gemm2_init(const char** err) {
  if (cudaFuncSetAttribute(gemm2_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           G2T<false>::SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm2_tcgen05_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           G2T<false>::SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm2_tcgen05_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           G2T<false>::SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm2_tcgen05_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           G2T<false>::SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm2_tcgen05_kernel<false, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           G2T<false>::SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm2_tcgen05_kernel<false, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           G2T<false>::SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm2_tcgen05_kernel<false, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           G2T<false>::SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm2_tcgen05_kernel<false, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           G2T<false>::SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm2_tcgen05_kernel<true, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           G2T<true>::SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm2_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           G2T<true>::SMEM_BYTES) != cudaSuccess) {
    *err = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed for the 2-CTA GEMM kernel";
    return 1;
  }
  return 0;
}

template <class... KArgs, class... Args>
static void launch2_pdl(void (*kernel)(KArgs...), int grid, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = gemm_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, args...);
}

int gemm2_launch(const GemmOp* op, cudaStream_t stream) {
  Gemm2Args a;
  a.ep = op->ep; a.M = op->M; a.N = op->N; a.K = op->K; a.reverse = op->reverse;
  a.conv_w2 = op->conv_w2; a.conv_cblocks = op->conv_cblocks;
  if (op->conv_cblocks > 0 || op->ep.border_w2 > 0)
    launch2_pdl(gemm2_tcgen05_kernel<false, 3>, op->grid, G2T<false>::SMEM_BYTES, stream, op->ta, op->tb, op->ta, a);
  else if (op->res_tma && op->ep.ln_stats_out != nullptr)
    launch2_pdl(gemm2_tcgen05_kernel<true, 0, 1>, op->grid, G2T<true>::SMEM_BYTES, stream, op->ta, op->tb, op->tr, a);
  else if (op->res_tma)
    launch2_pdl(gemm2_tcgen05_kernel<true>, op->grid, G2T<true>::SMEM_BYTES, stream, op->ta, op->tb, op->tr, a);
  else {
    // the two hot consumer shapes of the ViT encoder get the compile-time-specialised epilogue (SPM_GEMM_EPI=0: off)
    static const bool allow_spec = [] { const char* e = getenv("SPM_GEMM_EPI"); return e == nullptr || atoi(e) != 0; }();
    const GemmEpilogue& e = op->ep;
    const bool simple = allow_spec && e.out_bf16 && e.bias != nullptr && e.residual == nullptr &&
                        e.residual_bf16 == nullptr && e.out_row_group == 0 && e.border_w2 == 0 &&
                        op->N % G2T<false>::BN == 0 && (e.ldo % 8) == 0 &&
                        (reinterpret_cast<uintptr_t>(e.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0;
    if (simple && e.ln_stats_in != nullptr && e.ln_colsum == nullptr && e.act == ACT_NONE)
      launch2_pdl(gemm2_tcgen05_kernel<false, 1, 2>, op->grid, G2T<false>::SMEM_BYTES, stream, op->ta, op->tb, op->ta, a);
    else if (simple && e.ln_stats_in != nullptr && e.ln_colsum == nullptr && e.act == ACT_QUICKGELU)
      launch2_pdl(gemm2_tcgen05_kernel<false, 2, 2>, op->grid, G2T<false>::SMEM_BYTES, stream, op->ta, op->tb, op->ta, a);
    else if (simple && e.ln_stats_in != nullptr && e.act == ACT_NONE)
      launch2_pdl(gemm2_tcgen05_kernel<false, 1, 1>, op->grid, G2T<false>::SMEM_BYTES, stream, op->ta, op->tb, op->ta, a);
    else if (simple && e.ln_stats_in != nullptr && e.act == ACT_QUICKGELU)
      launch2_pdl(gemm2_tcgen05_kernel<false, 2, 1>, op->grid, G2T<false>::SMEM_BYTES, stream, op->ta, op->tb, op->ta, a);
    else if (simple && e.act == ACT_NONE)
      launch2_pdl(gemm2_tcgen05_kernel<false, 1>, op->grid, G2T<false>::SMEM_BYTES, stream, op->ta, op->tb, op->ta, a);
    else if (simple && e.act == ACT_QUICKGELU)
      launch2_pdl(gemm2_tcgen05_kernel<false, 2>, op->grid, G2T<false>::SMEM_BYTES, stream, op->ta, op->tb, op->ta, a);
    else
      launch2_pdl(gemm2_tcgen05_kernel<false>, op->grid, G2T<false>::SMEM_BYTES, stream, op->ta, op->tb, op->ta, a);
  }
  return (int)cudaGetLastError();
}

}  // namespace spm
