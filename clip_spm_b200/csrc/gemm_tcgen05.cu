// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   out[orow(m), n] = act( sum_k A[m,k] * B[n,k] + bias[n] ) (+ residual[rrow(m), n])
//
// A [M,K] and B [N,K] are K-contiguous (a torch.nn.Linear weight [out,in] is already "B").
// Roles inside one 384-thread CTA (one CTA per SM, grid = min(#tiles, #SMs), static round-robin tiles):
//   warp 0 (1 lane)  TMA producer: 128B-swizzled [128 x 128B] A box + [BN x 128B] B box per stage
//   warp 1 (1 lane)  MMA issuer  : 4 x tcgen05.mma (128 x BN x 32 bytes of K) per stage, fp32 accum in TMEM
//   warp 2           TMEM allocator (2 accumulator buffers of BN columns -> epilogue overlaps the next tile)
//   warps 4..11      epilogue    : two warps per TMEM lane quarter, alternating 128-byte column groups:
//                                  tcgen05.ld (one accumulator row per thread) -> bias/act -> swizzled smem staging
//                                  -> row-contiguous read-back (+ prefetched residual) -> full-line global stores
//
// This replaces the cuBLASLt / cuDNN calls the reference makes through ATen for
// models/clip_fsar.py:626-632,673,687 (ViT linears, patch-embed conv, projection) and
// models/myRes.py:944-996 + models/model_clipspm.py:76-99,171-174 (head linears, gates, temporal convs).
#include <cstdlib>

#include "gemm.cuh"
#include "gemm_epilogue.cuh"
#include "profile.cuh"
#include "ptx.cuh"

namespace spm {

struct GemmArgs {
  GemmEpilogue ep;
  int M, N, K;
  int conv_w2, conv_cblocks;  // implicit 3x3 convolution (see GemmOp), 0 = plain GEMM
  int conv_pair;
  int reverse;
};

template <int BN, int KIND>
struct GemmTile {
  static constexpr int BM = 128;
  static constexpr int ELEM = (KIND == GEMM_BF16) ? 2 : 4;
  static constexpr int BK = 128 / ELEM;  // elements per 128-byte swizzled row
  static constexpr int A_BYTES = BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : (BN == 64 ? 7 : 8));   // 32 / 64-wide tiles: small-N convolutions
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STG_BYTES_PER_WARP = 32 * 128;  // epilogue staging tile of one warp
  static constexpr int EPI_WARPS = 8;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + EPI_WARPS * STG_BYTES_PER_WARP + 1024;  // +1024: alignment slack
};

// CONV: implicit-3x3-convolution addressing in the producer + zero-border epilogue (rn50.cu); the plain GEMM
// instantiations (CONV = false) contain none of that code
template <int BN, int KIND, bool CONV = false>
__global__ void __launch_bounds__(384, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmArgs args) {
  using T = GemmTile<BN, KIND>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + T::STAGES * T::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + T::STAGES;
  uint64_t* tfull_bar = empty_bar + T::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int M = args.M, N = args.N, K = args.K;
  const int num_m = (M + T::BM - 1) / T::BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
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
      mbar_init(&tempty_bar[s], T::EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, T::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: the prologue above overlapped the predecessor's tail; from here on its results are needed
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      // L2 prefetch iterator over this CTA's (tile, k-block) sequence, GEMM_L2_PREFETCH_KB ahead of the loads
      int pf_tile = blockIdx.x, pf_kb = 0;
      auto prefetch_next = [&]() {
        if (pf_tile < num_tiles) {
          tma_prefetch_l2_2d(&tmA, pf_kb * T::BK, ((args.reverse ? num_tiles - 1 - pf_tile : pf_tile) / num_n) * T::BM);
          if (++pf_kb == num_kb) { pf_kb = 0; pf_tile += gridDim.x; }
        }
      };
      constexpr bool conv = CONV;  // taps re-read an L2-resident activation: no prefetch needed
      if (!conv)
        for (int i = 0; i < GEMM_L2_PREFETCH_KB; ++i) prefetch_next();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int rt = args.reverse ? num_tiles - 1 - tile : tile;
        const int m0 = (rt / num_n) * T::BM;
        const int n0 = (rt % num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!conv) prefetch_next();
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * T::STAGE_BYTES;
          uint8_t* sb = sa + T::A_BYTES;
          mbar_expect_tx(&full_bar[stage], T::STAGE_BYTES);
          if (conv && args.conv_pair) {
            const int roff = ((kb >> 1) - 1) * args.conv_w2 + ((kb & 1) ? 1 : -1);
            tma_load_2d(sa, &tmA, &full_bar[stage], 0, m0 + roff);
          } else if (conv && args.conv_cblocks > 0) {
            const int tap = kb / args.conv_cblocks, cb = kb - tap * args.conv_cblocks;
            const int roff = (tap / 3 - 1) * args.conv_w2 + (tap % 3 - 1);  // rows outside the matrix read as zeros
            tma_load_2d(sa, &tmA, &full_bar[stage], cb * T::BK, m0 + roff);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * T::BK, m0);
          }
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * T::BK, n0);
          if (++stage == T::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma_idesc(KIND == GEMM_BF16 ? 1 : 2, T::BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const int acc = t & 1;
        const uint32_t acc_phase = (t >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * T::STAGE_BYTES);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + T::A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // +32 bytes along K inside the 128B swizzle atom == +2 in the (addr >> 4) field
            if (KIND == GEMM_BF16)
              mma_bf16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
            else
              mma_tf32_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == T::STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const GemmEpilogue& ep = args.ep;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    // warp-private staging tile: 32 rows x 128 B, 16-byte units XOR-swizzled by (row & 7) -> conflict-free both ways
    const int half = (warp - 4) >> 2;  // which of the two warps of this lane quarter: takes column groups g % 2 == half
    const uint32_t stg_u = smem_u32(smem + T::STAGES * T::STAGE_BYTES + T::BAR_BYTES + (warp - 4) * T::STG_BYTES_PER_WARP);
    int t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int acc = t & 1;
      const uint32_t acc_phase = (t >> 1) & 1;
      const int rt = args.reverse ? num_tiles - 1 - tile : tile;
      const int m_base = (rt / num_n) * T::BM + q * 32;
      const int n0 = (rt % num_n) * BN;
      if (CONV) prefetch_residual_bf16<BN>(ep, m_base, n0, M, N, lane, half);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      if (CONV && conv_epilogue_applies(ep)) {   // (warp-uniform) compile-time specialised convolution epilogue
        if (ep.residual_bf16 != nullptr) gemm_epilogue_tile_conv<BN, true>(ep, stg_u, taddr, m_base, n0, M, N, lane, half);
        else gemm_epilogue_tile_conv<BN, false>(ep, stg_u, taddr, m_base, n0, M, N, lane, half);
      } else {
        gemm_epilogue_tile<BN, CONV>(ep, stg_u, taddr, m_base, n0, M, N, lane, half);
      }
      // accumulator buffer drained -> hand it back to the MMA warp
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, T::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  return fn;
}

// 2-D K-contiguous operand [rows, K] with row stride ld (elements): box = [box_rows x 128 bytes], SWIZZLE_128B.
static int make_operand_map(CUtensorMap* map, int kind, const void* ptr, long long ld, int rows, int K, int box_rows,
                            const char** err) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) { *err = "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)"; return 1; }
  const int elem = kind == GEMM_BF16 ? 2 : 4;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * elem) & 15)) {
    *err = "GEMM operand must be 16-byte aligned with a 16-byte-multiple row stride";
    return 1;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)(ld * elem)};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elem), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, kind == GEMM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { *err = "cuTensorMapEncodeTiled failed"; return 1; }
  return 0;
}

int make_tensor_map_f32_nd(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                           const unsigned long long* strides_bytes, const unsigned* box) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc || rank < 2 || rank > 5) return 1;
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    if (strides_bytes[i] % 16 != 0) return 1;
    gstride[i] = strides_bytes[i];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (box[0] != 32 && box[0] != 16)) return 1;
  // the innermost box extent picks the swizzle: 32 floats = 128-byte rows, 16 floats = 64-byte rows
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, bx, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, box[0] == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 1;
}

int make_operand_map_rows(CUtensorMap* map, int kind, const void* ptr, long long ld, int rows, int K, int box_rows,
                          const char** err) {
  return make_operand_map(map, kind, ptr, ld, rows, K, box_rows, err);
}

bool gemm_conv_pair_supported() {
  static const int ok = [] {
    if (const char* e = getenv("SPM_CONV_PAIR")) if (atoi(e) == 0) return 0;
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc) return 0;
    alignas(64) static unsigned char dummy[64];
    CUtensorMap m;
    cuuint64_t gdim[2] = {64, 1000};
    cuuint64_t gstride[1] = {64};  // 32 bf16: consecutive rows overlap by half
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    return enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dummy, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 1 : 0;
  }();
  return ok != 0;
}

int gemm_init(const char** err) {
#define SPM_SET_SMEM(BN, KIND)                                                                               \
  if (cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                           GemmTile<BN, KIND>::SMEM_BYTES) != cudaSuccess) {                                 \
    *err = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed for the GEMM kernel";                    \
    return 1;                                                                                                \
  }
  SPM_SET_SMEM(256, GEMM_BF16)
  SPM_SET_SMEM(128, GEMM_BF16)
  SPM_SET_SMEM(256, GEMM_TF32)
  SPM_SET_SMEM(128, GEMM_TF32)
#undef SPM_SET_SMEM
  if (cudaFuncSetAttribute(gemm_tcgen05_kernel<256, GEMM_BF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           GemmTile<256, GEMM_BF16>::SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_tcgen05_kernel<128, GEMM_BF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           GemmTile<128, GEMM_BF16>::SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_tcgen05_kernel<64, GEMM_BF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           GemmTile<64, GEMM_BF16>::SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_tcgen05_kernel<32, GEMM_BF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           GemmTile<32, GEMM_BF16>::SMEM_BYTES) != cudaSuccess) {
    *err = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed for the conv GEMM kernel";
    return 1;
  }
  if (conv_win_init()) { *err = "cudaFuncSetAttribute failed for the window convolution kernel"; return 1; }
  return gemm2_init(err);
}

int gemm_plan(GemmOp* op, int kind, const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
              const GemmEpilogue& ep, int num_sms, const char** err) {
  if (M <= 0 || N <= 0 || K <= 0) { *err = "GEMM: empty problem"; return 1; }
  if (N % 32 != 0) { *err = "GEMM: N must be a multiple of 32"; return 1; }
  if (ep.out == nullptr) { *err = "GEMM: null output"; return 1; }
  if (ep.residual != nullptr && ep.out_bf16) { *err = "GEMM: an fp32 residual needs an fp32 output"; return 1; }
  if (ep.residual_bf16 != nullptr && (!ep.out_bf16 || ep.ldr % 8 != 0)) {
    *err = "GEMM: a bf16 residual needs a bf16 output and 16-byte aligned rows";
    return 1;
  }
  if (ep.out_bf16 ? (ep.ldo % 8 != 0) : (ep.ldo % 4 != 0)) { *err = "GEMM: output rows must be 16-byte aligned"; return 1; }
  if (ep.residual != nullptr && ep.ldr % 4 != 0) { *err = "GEMM: residual rows must be 16-byte aligned"; return 1; }
  op->M = M; op->N = N; op->K = K; op->kind = kind; op->ep = ep;
  if (kind == GEMM_F32_SIMT) {
    if (ep.out_bf16 || ep.residual_bf16 != nullptr) { *err = "GEMM: the fp32 SIMT kernel has fp32 outputs only"; return 1; }
    if (K % 4 != 0 || lda % 4 != 0 || ldb % 4 != 0) { *err = "GEMM: fp32 SIMT operands need 16-byte aligned rows"; return 1; }
    op->simt_a = A; op->simt_b = B; op->simt_lda = lda; op->simt_ldb = ldb; op->two_cta = 0;
    return 0;
  }
  // Tile width: 256 for the big encoder GEMMs; 128 when that yields more CTAs than SMs can use otherwise
  const long long tiles256 = (long long)((M + 127) / 128) * ((N + 255) / 256);
  op->bn = (N % 256 == 0 && tiles256 >= num_sms) ? 256 : 128;
  // convolutions with few output channels (RN50 stem: 32 / 64, layer1: 64): a tile as wide as the layer instead of a
  // 128-wide one whose B rows are mostly TMA zero fill (4x / 2x the tensor, shared-memory and L2 work for nothing)
  static const bool allow_narrow = [] { const char* e = getenv("SPM_CONV_NARROW"); return e == nullptr || atoi(e) != 0; }();
  if (allow_narrow && kind == GEMM_BF16 && ep.border_w2 > 0 && N <= 64) op->bn = N <= 32 ? 32 : 64;
  const long long tiles = (long long)((M + 127) / 128) * ((N + op->bn - 1) / op->bn);
  op->grid = (int)(tiles < num_sms ? tiles : num_sms);
  // Large bf16 problems run on CTA pairs (cta_group::2): 256x256 tile per pair, each CTA stages half of B
  static const bool allow_2cta = [] { const char* e = getenv("SPM_GEMM_2CTA"); return e == nullptr || atoi(e) != 0; }();
  const long long pair_tiles = (long long)((M + 255) / 256) * (N / 256);
  // Measured twice.  Kernel alone under ncu at max clock (profiles/r01_ncu_gemm_2cta_vs_1cta.txt): the pair kernel wins
  // on long reductions (K = 3072: 178 vs 188 us) and loses on K = 768 (cluster-wide barriers per tile).  Inside the
  // real step, which runs at the 1000 W power cap (~1.5 GHz), it wins everywhere: a pair stages 32 KB of operands per
  // k-block and CTA instead of 48 KB, and the step gets 4 % faster (median 85.9 vs 89.5 ms, interleaved A/B runs on
  // one box) -- energy per flop is what bounds a power-capped step.  So it is the default for every eligible shape;
  // SPM_GEMM_2CTA=1 restricts it to K >= 2048, 0 disables it.
  static const int mode_2cta = [] { const char* e = getenv("SPM_GEMM_2CTA"); return e == nullptr ? 2 : atoi(e); }();
  // Convolutions (zero-bordered outputs, rn50.cu) with N % 256 == 0 take the pair kernel whatever their tile count: the
  // alternative, 128-wide 1-CTA tiles, is bound by shared-memory traffic at half the tensor rate (SPM_CONV_2CTA=0: off)
  static const bool conv_2cta = [] { const char* e = getenv("SPM_CONV_2CTA"); return e == nullptr || atoi(e) != 0; }();
  if (ep.border_w2 > 0)
    op->two_cta = (allow_2cta && conv_2cta && kind == GEMM_BF16 && N % 256 == 0) ? 1 : 0;
  else
    op->two_cta = (allow_2cta && kind == GEMM_BF16 && N % 256 == 0 && pair_tiles >= num_sms / 2 &&
                   (K >= 2048 || mode_2cta == 2)) ? 1 : 0;
  // LayerNorm folding (GemmEpilogue::ln_*) lives in the pair kernel's epilogues only: taken whatever the tile count, so that
  // a frame's features do not depend on the size of the chunk it was encoded in
  const bool ln_fold = ep.ln_stats_in != nullptr || ep.ln_stats_out != nullptr;
  if (ln_fold) {
    if (kind != GEMM_BF16 || N % 256 != 0 || (ep.ln_stats_out != nullptr && N != LN_FOLD_C)) {
      *err = "LayerNorm folding needs the bf16 pair kernel (N % 256 == 0; producer N == 768)";
      return 1;
    }
    op->two_cta = 1;
  }
  if (op->two_cta) {
    op->bn = 256;
    const long long pairs = pair_tiles < num_sms / 2 ? pair_tiles : num_sms / 2;
    op->grid = (int)(2 * pairs);
  }
  if (make_operand_map(&op->ta, kind, A, lda, M, K, 128, err)) return 1;
  if (make_operand_map(&op->tb, kind, B, ldb, N, K, op->two_cta ? 128 : op->bn, err)) return 1;
  static const bool allow_res_tma = [] { const char* e = getenv("SPM_GEMM_RES_TMA"); return e == nullptr || atoi(e) != 0; }();
  op->res_tma = 0;
  if (allow_res_tma && op->two_cta && ep.residual != nullptr && !ep.out_bf16 && ep.act == ACT_NONE &&
      !ep.relu_after_residual && ep.res_row_mod == 0 && ep.out_row_group == 0 &&
      (reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0 && (ep.ldr & 3) == 0) {
    if (make_operand_map(&op->tr, GEMM_TF32, ep.residual, ep.ldr, M, N, 32, err)) return 1;
    op->res_tma = 1;
  }
  op->conv_win = 0;
  if (ep.border_w2 > 0 && !op->two_cta) conv_win_plan_1x1(op, A, lda, num_sms);   // few-channel 1x1 convolutions (conv_win.cu)
  return 0;
}

int gemm_plan_conv3x3(GemmOp* op, const void* A, int C, int rows, int W2, const void* B, int Cout, const GemmEpilogue& ep,
                      int num_sms, const char** err) {
  if (C % 8 != 0 || Cout % 32 != 0) { *err = "conv3x3: C must be a multiple of 8 and Cout of 32"; return 1; }
  const bool pair = C == 32 && gemm_conv_pair_supported();
  const int cpad = (C + 63) / 64 * 64;
  const int Kv = pair ? 6 * 64 : 9 * cpad;  // k extent of the folded weight matrix
  // plan as a [rows, Kv] x [Cout, Kv] GEMM (fictitious A stride), then describe the real A
  if (gemm_plan(op, GEMM_BF16, A, Kv, B, Kv, rows, Cout, Kv, ep, num_sms, err)) return 1;
  if (pair) {
    // virtual row p = channels of pixels p and p+1 (row stride 32 elements, extent 64); the last pixel has no successor
    if (make_operand_map(&op->ta, GEMM_BF16, A, C, rows - 1, 64, 128, err)) return 1;
  } else {
    // (columns >= C of a box read as zeros, rows outside [0, rows) too)
    if (make_operand_map(&op->ta, GEMM_BF16, A, C, rows, C, 128, err)) return 1;
  }
  op->res_tma = 0;   // (both kernels address the taps in their producers; the pair kernel's A box is the same 128 rows)
  op->conv_w2 = W2;
  op->conv_cblocks = pair ? 1 : cpad / 64;
  op->conv_pair = pair ? 1 : 0;
  if (!op->two_cta) conv_win_plan(op, A, C, rows, W2, Cout, num_sms);   // few-channel layers: one shared-memory window per tile
  return 0;
}

// Launch with the programmatic-stream-serialization attribute (PDL, ptx.cuh) unless SPM_PDL=0.
bool gemm_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("SPM_PDL"); return e == nullptr || atoi(e) != 0; }();
  return on;
}
template <class... KArgs, class... Args>
static void launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = gemm_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, args...);
}

int gemm_run(const GemmOp* op, cudaStream_t stream, const char** err) {
  GemmArgs a;
  a.ep = op->ep; a.M = op->M; a.N = op->N; a.K = op->K;
  a.conv_w2 = op->conv_w2; a.conv_cblocks = op->conv_cblocks; a.conv_pair = op->conv_pair; a.reverse = op->reverse;
  int slot = -1;
  const bool prof = profile_gemm_begin(stream, (op->kind & 1) * 2 + (op->bn == 256 ? 1 : 0),
                                       2.0 * (double)op->M * (double)op->N * (double)op->K, &slot, op->M, op->N, op->K);
#define SPM_LAUNCH(BN, KIND) \
  launch_pdl(gemm_tcgen05_kernel<BN, KIND>, op->grid, 384, GemmTile<BN, KIND>::SMEM_BYTES, stream, op->ta, op->tb, a)
  if (op->kind == GEMM_F32_SIMT) {
    sgemm_f32_run(op, stream);
  } else if (op->conv_win) {
    conv_win_launch(op, stream);
  } else if (op->two_cta) {
    gemm2_launch(op, stream);
  } else if (op->conv_cblocks > 0 || op->ep.border_w2 > 0) {  // convolution path (bf16): taps / zero borders
#define SPM_LAUNCH_CONV(BN) \
  launch_pdl(gemm_tcgen05_kernel<BN, GEMM_BF16, true>, op->grid, 384, GemmTile<BN, GEMM_BF16>::SMEM_BYTES, stream, op->ta, op->tb, a)
    if (op->bn == 256) SPM_LAUNCH_CONV(256);
    else if (op->bn == 128) SPM_LAUNCH_CONV(128);
    else if (op->bn == 64) SPM_LAUNCH_CONV(64);
    else SPM_LAUNCH_CONV(32);
#undef SPM_LAUNCH_CONV
  } else if (op->kind == GEMM_BF16) {
    if (op->bn == 256) SPM_LAUNCH(256, GEMM_BF16); else SPM_LAUNCH(128, GEMM_BF16);
  } else {
    if (op->bn == 256) SPM_LAUNCH(256, GEMM_TF32); else SPM_LAUNCH(128, GEMM_TF32);
  }
#undef SPM_LAUNCH
  cudaError_t e = cudaGetLastError();
  if (prof) profile_gemm_end(stream, slot);
  if (e != cudaSuccess) { *err = cudaGetErrorString(e); return 1; }
  count_launch();
  return 0;
}

}  // namespace spm
