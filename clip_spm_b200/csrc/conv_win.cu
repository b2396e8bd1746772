// 3x3 convolution (stride 1, pad 1) over zero-bordered NHWC bf16 images with few channels (RN50 stem conv2 / conv3:
// 32 -> 32 / 64; layer1 conv2: 64 -> 64), as an implicit tcgen05 GEMM whose nine taps are served from ONE shared-memory
// window per tile (r02).
//
// Why: the generic implicit-GEMM path (gemm_tcgen05.cu, CONV) fetches a 128-row TMA box per tap, i.e. reads every input
// row nine times from L2.  For the early, wide-image layers that traffic IS the bound -- ncu / launch lists of a 216-frame
// chunk: layer1 conv2 817 MB of operand reads in 138 us = 5.9 TB/s, stem conv2 2.1 GB in 317 us = 6.6 TB/s, against an
// L2 -> SM limit of ~6.3 kB/clk -- while tensor pipe and HBM idle.  The taps of a tile of 128 consecutive pixel rows all
// lie in the rows [m0 - (W2+1), m0 + 127 + (W2+1)] of the same matrix (tap (dy,dx) = row offset dy*W2 + dx), so:
//   * the producer loads that window once per tile (two TMA boxes, SWIZZLE_128B, rows = 64 channels = 128 bytes; for
//     32 channels a row is the "pixel pair" p | p+1 of gemm_plan_conv3x3): 246-360 rows instead of 9 x 128,
//   * the MMA warp addresses tap t by a shared-memory descriptor that starts (offset_t + W2 + 1) rows into the window.
//     TMA and the tensor core both derive the 128-byte swizzle phase from the absolute shared-memory address, so a start
//     that is not a multiple of 8 rows needs nothing else: the descriptor's base-offset field stays 0 (measured: with
//     (address >> 7) & 7 in it the results are wrong, with 0 they match the goldens),
//   * the folded weights (<= 74 KB) stay resident in shared memory for the whole kernel: no B traffic per tile.
//   * a tile is 256 pixel rows = TWO M = 128 accumulators side by side in tensor memory: the eight epilogue warps split
//     by accumulator (all of them busy even when there are only 32 output channels) and the per-tile barrier round trips
//     (TMA -> MMA -> epilogue -> MMA), which bound these short-K layers at ~1.4 us per 128-row tile whatever their K, are
//     paid once per 256 rows; the halo also amortises better (stem: 486 window rows per 256 instead of 358 per 128).
// The same kernel runs the few-channel 1x1 convolutions (stem conv1 on its im2col matrix, layer1.0 conv1): one tap, no halo.
// Roles / epilogue as in gemm_tcgen05.cu (warp 0 TMA, warp 1 MMA, warp 2 TMEM, warps 4-11 the convolution epilogue).
#include <cstdlib>

#include "gemm.cuh"
#include "gemm_epilogue.cuh"
#include "profile.cuh"
#include "ptx.cuh"

namespace spm {

constexpr int CW_TILE_M = 256;
struct ConvWinArgs {
  GemmEpilogue ep;
  int M, N;          // pixel rows (all padded pixels), output channels (<= BN)
  int num_kb;        // k-blocks of 64 bf16 = taps: 9 (64 channels), 6 (32 channels as pixel pairs) or 1 (1x1)
  int roff[9];       // row offset of each tap relative to the output row
  int halo;          // rows loaded before / after the tile: max |roff|
  int win_rows;      // rows of a window (multiple of 16): 256 + 2*halo, rounded up
  int n_win;         // window stages
};

namespace {
constexpr int CW_EPI_WARPS = 8;
constexpr int CW_BAR_BYTES = 256;

}  // namespace

template <int BN>
__global__ void __launch_bounds__(384, 1)
conv3x3_win_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmB, const ConvWinArgs args) {
  extern __shared__ uint8_t smem_raw_cw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_cw) + 1023) & ~uintptr_t(1023));
  constexpr int B_KB_BYTES = BN * 128;                       // one k-block of the folded weights
  const int b_bytes = args.num_kb * B_KB_BYTES;              // multiple of 1024 (BN >= 32: 4096 per k-block)
  const int win_bytes = args.win_rows * 128;                 // multiple of 1024 (win_rows % 16 == 0)
  uint8_t* sB = smem;
  uint8_t* sW = smem + b_bytes;
  uint8_t* sBar = sW + args.n_win * win_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sBar);    // [n_win <= 8]
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* b_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_bar + 1);
  const uint32_t stg_base = smem_u32(sBar + CW_BAR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = args.M, N = args.N;
  const int num_tiles = (M + CW_TILE_M - 1) / CW_TILE_M;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < args.n_win; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], CW_EPI_WARPS); }
    mbar_init(b_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 4 * BN);   // two buffers of two accumulators
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer: weights once, then one window per tile =====================
      mbar_expect_tx(b_bar, (uint32_t)b_bytes);
      for (int kb = 0; kb < args.num_kb; ++kb) tma_load_2d(sB + kb * B_KB_BYTES, &tmB, b_bar, kb * 64, 0);
      int stage = 0;
      uint32_t phase = 0;
      const int half_rows = args.win_rows / 2;   // a TMA box is at most 256 rows
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int r0 = tile * CW_TILE_M - args.halo;   // rows outside the matrix read as zeros
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* w = sW + stage * win_bytes;
        mbar_expect_tx(&full_bar[stage], (uint32_t)win_bytes);
        tma_load_2d(w, &tmW, &full_bar[stage], 0, r0);
        tma_load_2d(w + half_rows * 128, &tmW, &full_bar[stage], 0, r0 + half_rows);
        if (++stage == args.n_win) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma_idesc(1, 128, BN);
      mbar_wait(b_bar, 0);
      tc_fence_after_sync();
      int stage = 0;
      uint32_t phase = 0;
      int t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const int acc = t & 1;
        mbar_wait(&tempty_bar[acc], (uint32_t)(((t >> 1) & 1) ^ 1));
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        const uint32_t w_u = smem_u32(sW + stage * win_bytes), b_u = smem_u32(sB);
#pragma unroll 1
        for (int mb = 0; mb < 2; ++mb) {   // the two 128-row accumulators of the tile
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * BN + mb * BN);
          for (int kb = 0; kb < args.num_kb; ++kb) {
            // tap kb = the window from row (roff + halo) on: any 128-byte row is a legal descriptor start
            const uint64_t adesc = umma_desc_k_sw128(w_u + (uint32_t)((args.roff[kb] + args.halo + mb * 128) * 128));
            const uint64_t bdesc = umma_desc_k_sw128(b_u + (uint32_t)(kb * B_KB_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_bf16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
          }
        }
        tc_commit(&empty_bar[stage]);   // window reusable once these MMAs have read it
        tc_commit(&tfull_bar[acc]);     // accumulator complete -> epilogue
        if (++stage == args.n_win) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const GemmEpilogue& ep = args.ep;
    const int q = warp & 3, mb = (warp - 4) >> 2;   // lane quarter, accumulator (rows mb*128 ..) of this warp
    const uint32_t stg_u = stg_base + (uint32_t)((warp - 4) * 4096);
    int t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int acc = t & 1;
      const int m_base = tile * CW_TILE_M + mb * 128 + q * 32;
      mbar_wait(&tfull_bar[acc], (uint32_t)((t >> 1) & 1));
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * BN + mb * BN);
      gemm_epilogue_tile_conv<BN, false>(ep, stg_u, taddr, m_base, 0, M, N, lane, 0);   // all column groups of its rows
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 4 * BN);
  }
}

namespace {
constexpr int CW_SMEM_MAX = 227 * 1024;
int conv_win_smem(int b_bytes, int win_bytes, int n_win) {
  return b_bytes + n_win * win_bytes + CW_BAR_BYTES + CW_EPI_WARPS * 4096 + 1024;
}
}  // namespace

static bool conv_win_enabled() {
  static const bool on = [] { const char* e = getenv("SPM_CONV_WIN"); return e == nullptr || atoi(e) != 0; }();
  return on;
}
static bool conv_win_epilogue_ok(const GemmEpilogue& e) {
  return e.out_bf16 && e.bias != nullptr && e.residual == nullptr && e.residual_bf16 == nullptr && e.border_w2 > 0 &&
         e.out_row_group == 0 && e.res_row_mod == 0 && (e.act == ACT_NONE || e.act == ACT_RELU);
}
// shared tail of the two planners: window geometry, stage count, window tensor map (`ld` / `map_rows`: the matrix the
// generic path reads; 64 bf16 per box row)
static bool conv_win_finish(GemmOp* op, const void* A, long long ld, int map_rows, int map_k, int num_kb, int halo, int Cout,
                            int num_sms) {
  const int win_rows = (CW_TILE_M + 2 * halo + 15) / 16 * 16;
  if (win_rows / 2 > 256) return false;
  const int b_bytes = num_kb * Cout * 128, win_bytes = win_rows * 128;
  int n_win = (CW_SMEM_MAX - conv_win_smem(b_bytes, 0, 0)) / win_bytes;
  if (n_win < 2) return false;
  if (n_win > 4) n_win = 4;
  const char* err = "";
  if (make_operand_map_rows(&op->tr, GEMM_BF16, A, ld, map_rows, map_k, win_rows / 2, &err)) return false;
  op->conv_win = 1;
  op->conv_win_rows = win_rows;
  op->conv_win_stages = n_win;
  op->conv_win_taps = num_kb;
  op->conv_win_halo = halo;
  const long long tiles = ((long long)op->M + CW_TILE_M - 1) / CW_TILE_M;
  op->grid = (int)(tiles < num_sms ? tiles : num_sms);
  return true;
}

// 3x3 convolution: fills op->conv_win_* when it qualifies (C in {32 as pixel pairs, 64}, Cout in {32, 64}, the window
// ring fits); returns false otherwise and leaves the generic implicit-GEMM plan untouched.
bool conv_win_plan(GemmOp* op, const void* A, int C, int rows, int W2, int Cout, int num_sms) {
  if (!conv_win_enabled() || (Cout != 32 && Cout != 64) || !(C == 64 || (C == 32 && op->conv_pair))) return false;
  if (!conv_win_epilogue_ok(op->ep)) return false;
  const int num_kb = op->conv_pair ? 6 : 9;
  for (int kb = 0; kb < num_kb; ++kb)
    op->conv_win_roff[kb] = op->conv_pair ? ((kb >> 1) - 1) * W2 + ((kb & 1) ? 1 : -1) : (kb / 3 - 1) * W2 + (kb % 3 - 1);
  return conv_win_finish(op, A, op->conv_pair ? 32 : C, op->conv_pair ? rows - 1 : rows, 64, num_kb, W2 + 1, Cout, num_sms);
}
// few-channel 1x1 convolution / plain GEMM with the convolution epilogue (K <= 64, N in {32, 64}): one tap, no halo
bool conv_win_plan_1x1(GemmOp* op, const void* A, long long lda, int num_sms) {
  if (!conv_win_enabled() || (op->N != 32 && op->N != 64) || op->K > 64 || op->kind != GEMM_BF16) return false;
  if (!conv_win_epilogue_ok(op->ep)) return false;
  op->conv_win_roff[0] = 0;
  return conv_win_finish(op, A, lda, op->M, op->K, 1, 0, op->N, num_sms);
}

int conv_win_init() {
  if (cudaFuncSetAttribute(conv3x3_win_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM_MAX) != cudaSuccess ||
      cudaFuncSetAttribute(conv3x3_win_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM_MAX) != cudaSuccess)
    return 1;
  return 0;
}

void conv_win_launch(const GemmOp* op, cudaStream_t stream) {
  ConvWinArgs a;
  a.ep = op->ep; a.M = op->M; a.N = op->N;
  a.num_kb = op->conv_win_taps; a.halo = op->conv_win_halo;
  for (int i = 0; i < 9; ++i) a.roff[i] = i < a.num_kb ? op->conv_win_roff[i] : 0;
  a.win_rows = op->conv_win_rows; a.n_win = op->conv_win_stages;
  const int smem = conv_win_smem(a.num_kb * op->N * 128, a.win_rows * 128, a.n_win);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)op->grid); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = gemm_pdl_enabled() ? 1 : 0;
  if (op->N == 32) cudaLaunchKernelEx(&cfg, conv3x3_win_kernel<32>, op->tr, op->tb, a);
  else cudaLaunchKernelEx(&cfg, conv3x3_win_kernel<64>, op->tr, op->tb, a);
}

}  // namespace spm
