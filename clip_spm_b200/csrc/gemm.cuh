// Host-side interface of the tcgen05 GEMM used by every dense contraction on the path
// (frame-encoder linears, patch embedding, head linears, temporal convolutions as GEMMs).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace spm {

// GEMM_F32_SIMT: exact fp32 FFMA kernel (sgemm_f32.cu) used by the SPM_PRECISION_FP32 parity mode
enum GemmKind { GEMM_BF16 = 0, GEMM_TF32 = 1, GEMM_F32_SIMT = 2 };
enum GemmAct { ACT_NONE = 0, ACT_QUICKGELU = 1, ACT_GELU_ERF = 2, ACT_LEAKY = 3, ACT_SIGMOID = 4, ACT_RELU = 5 };

// out[orow(m), n] = act(sum_k A[m,k] * B[n,k] + bias[n]) (+ residual[rrow(m), n])
struct GemmEpilogue {
  const float* bias = nullptr;      // [N] fp32 or null
  const float* residual = nullptr;  // fp32, leading dim ldr, added AFTER the activation
  int ldr = 0;
  const void* residual_bf16 = nullptr;  // bf16 residual (needs a bf16 output), same row mapping / ldr
  int relu_after_residual = 0;          // out = relu(act(acc + bias) + residual)  (ResNet bottleneck tail)
  // zero-bordered NHWC images (rn50.cu): rows are pixels of [H2 x W2] padded images; rows on the 1-pixel border are
  // written as zeros so the tensor stays a valid zero-padded input of the next 3x3 convolution (0 = off)
  int border_w2 = 0, border_h2w2 = 0;
  int res_row_mod = 0;    // 0: rrow = orow ; >0: rrow = (m % res_row_mod) + res_row_off  (positional table)
  int res_row_off = 0;
  // 0: orow = m ; >0: orow = (m / out_row_group) * out_group_stride + (m % out_row_group) + out_row_off
  // (patch rows skip one class-token row per frame: 196 / 197 / 1; query tokens land after the supports of their episode)
  int out_row_group = 0;
  int out_group_stride = 0;
  int out_row_off = 0;
  void* out = nullptr;    // fp32 or bf16, leading dim ldo (elements)
  int ldo = 0;
  int out_bf16 = 0;
  int act = ACT_NONE;
  float slope = 0.f;      // LeakyReLU negative slope
  // LayerNorm folded into the neighbouring GEMMs (2-CTA kernel only, N == LN_FOLD_C on the producer side):
  //   producer (fp32 residual GEMM): also writes a bf16 copy of its output rows (out2_bf16, leading dim ldo) and, per row
  //   and (256-column tile, epilogue half), the partial sums {sum x, sum x^2} of the columns it produced (ln_stats_out);
  //   consumer (bf16 + bias GEMM whose A operand is that bf16 copy and whose weights carry gamma): out = act(rstd * (acc -
  //   mean * ln_colsum[n]) + bias[n]) with mean / rstd rebuilt from the row's partial sums (ln_stats_in).
  void* out2_bf16 = nullptr;
  float* ln_stats_out = nullptr;
  const float* ln_stats_in = nullptr;
  const float* ln_colsum = nullptr;
};
constexpr int LN_FOLD_C = 768;                        // row width the statistics cover (ViT-B/16 residual stream)
constexpr int LN_FOLD_SLOTS = 2 * (LN_FOLD_C / 256);  // partial sums per row: (n-tile, epilogue half)

struct GemmOp {
  CUtensorMap ta, tb;
  GemmEpilogue ep;
  int M = 0, N = 0, K = 0;
  int bn = 256;
  int kind = GEMM_BF16;
  int grid = 0;
  const void* simt_a = nullptr;  // GEMM_F32_SIMT: raw operand pointers / row strides (no tensor maps)
  const void* simt_b = nullptr;
  long long simt_lda = 0, simt_ldb = 0;
  // implicit 3x3 convolution (stride 1, pad 1) over a zero-bordered NHWC image matrix [rows, C]: k-block kb reads
  // channel block kb % conv_cblocks of tap kb / conv_cblocks, i.e. the same matrix at the constant row offset
  // (tap/3 - 1) * conv_w2 + (tap%3 - 1) -- only the TMA coordinate changes, no im2col is materialised (0 = plain GEMM)
  int conv_w2 = 0, conv_cblocks = 0;
  // C == 32 variant ("pixel pairs"): A is described as [rows-1, 64] with a row stride of 32 elements, so that virtual row
  // p holds the channels of pixels p and p+1 and every TMA box is a full 128-byte line; the 9 taps become 6 k-blocks
  // (dy, {x-1, x}) and (dy, {x+1, -}) instead of 9 half-empty ones
  int conv_pair = 0;
  // 3x3 convolutions with few channels (conv_win.cu): all nine taps served from one shared-memory window per tile, the
  // folded weights resident in shared memory; `tr` then holds the window tensor map
  int conv_win = 0, conv_win_rows = 0, conv_win_stages = 0, conv_win_taps = 0, conv_win_halo = 0;
  int conv_win_roff[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // row offset of each tap
  int reverse = 0;  // visit the output tiles in descending order (same results; L2 reuse between consecutive kernels)
  int two_cta = 0;  // 1: launched as 2-CTA clusters (gemm2_tcgen05.cu), 256x256 tile per CTA pair
  // 2-CTA kernel, fp32 out + fp32 residual with identity row mapping: the epilogue warps prefetch the residual tile
  // by TMA (`tr`: fp32 [M, N], 32 x 32 boxes, SWIZZLE_128B) one column group ahead instead of loading it into registers
  int res_tma = 0;
  CUtensorMap tr;
};

// A: [M, K] (row stride lda elements), B: [N, K] (row stride ldb) -- both K-contiguous ("K-major"),
// bf16 (GEMM_BF16) or fp32 (GEMM_TF32).  Returns 0 on success, else sets an error string.
int gemm_plan(GemmOp* op, int kind, const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
              const GemmEpilogue& ep, int num_sms, const char** err);
int gemm_run(const GemmOp* op, cudaStream_t stream, const char** err);
// 3x3/stride-1/pad-1 convolution as an implicit GEMM: A = zero-bordered NHWC activations [rows, C] bf16 (rows = all
// padded pixels), B = folded weights [Cout, 9 * cpad] with column tap*cpad + c (cpad = C rounded up to 64)
int gemm_plan_conv3x3(GemmOp* op, const void* A, int C, int rows, int W2, const void* B, int Cout, const GemmEpilogue& ep,
                      int num_sms, const char** err);
// true if the driver accepts a tensor map whose row stride is smaller than its row extent (needed by conv_pair)
bool gemm_conv_pair_supported();
// conv_win.cu
bool conv_win_plan(GemmOp* op, const void* A, int C, int rows, int W2, int Cout, int num_sms);
bool conv_win_plan_1x1(GemmOp* op, const void* A, long long lda, int num_sms);
int conv_win_init();
void conv_win_launch(const GemmOp* op, cudaStream_t stream);
// K-contiguous 2-D operand map with a caller-chosen box height (gemm_tcgen05.cu)
int make_operand_map_rows(CUtensorMap* map, int kind, const void* ptr, long long ld, int rows, int K, int box_rows,
                          const char** err);
// fp32 tensor map of rank 2..5 (dims / box innermost first, strides in BYTES for dims 1..rank-1, each a multiple of 16),
// SWIZZLE_128B (the box's innermost extent must be 32 floats = 128 bytes); 0 on success
int make_tensor_map_f32_nd(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                           const unsigned long long* strides_bytes, const unsigned* box);
bool gemm_pdl_enabled();   // programmatic dependent launch of the GEMM kernels (SPM_PDL=0 turns it off)
// one-time: opt into large dynamic shared memory for every instantiation
int gemm_init(const char** err);
// 2-CTA kernel (gemm2_tcgen05.cu)
int gemm2_init(const char** err);
int gemm2_launch(const GemmOp* op, cudaStream_t stream);
// fp32 SIMT kernel (sgemm_f32.cu)
int sgemm_f32_run(const GemmOp* op, cudaStream_t stream);

}  // namespace spm
