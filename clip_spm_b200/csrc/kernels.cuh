// Launchers of the non-GEMM kernels (all fp32 unless stated).  Each returns 0 / sets an error string.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace spm {

// ---- packing (weights are repacked once at load time) ---------------------------------------------------
int k_cast_bf16(cudaStream_t st, const float* in, __nv_bfloat16* out, long long n);
// in [R, C] fp32 -> out [C, R] bf16   (ViT `proj` [768,512] -> B operand [512,768])
int k_transpose_cast_bf16(cudaStream_t st, const float* in, __nv_bfloat16* out, int R, int C);
// Conv1d weight [O, I, 3] -> GEMM B operand [O, 3*I] with column index kk*I + i
int k_repack_conv1d(cudaStream_t st, const float* in, float* out, int O, int I);
int k_add_vec(cudaStream_t st, const float* a, const float* b, float* out, int n);

// ---- ViT frame encoder -------------------------------------------------------------------------------------
// images [F,3,224,224] fp32 -> patches [F*196, 768] bf16, column = c*256 + ky*16 + kx  (clip_fsar.py:673)
int k_patch_im2col(cudaStream_t st, const float* images, __nv_bfloat16* patches, int n_frames);
// LayerNorm over the last dim C (eps 1e-5, fp32 statistics; clip_fsar.py:610-616 / myRes.py:1036).
//   row r of the input is `in + r*in_stride`; if cls_period > 0 and r % cls_period == 0 the input row is
//   `cls_row` instead (class-token row = class_embedding + pos[0], clip_fsar.py:676-677).
//   Writes fp32 (out_f32) and/or bf16 (out_bf16), contiguous rows of C (row stride out_stride).
int k_layernorm(cudaStream_t st, const float* in, long long in_stride, int rows, int C, const float* gamma,
                const float* beta, const float* cls_row, int cls_period, float* out_f32, __nv_bfloat16* out_bf16,
                long long out_stride, int reverse = 0, float* stats_out = nullptr);
// LayerNorm folded into the next GEMM (GemmEpilogue::ln_*): gamma into the weights, beta into the bias, column sums
int k_fold_ln(cudaStream_t st, const float* W, const float* gamma, const float* beta, const float* bias, int N, int K,
              __nv_bfloat16* Wf, float* colsum, float* bias2, int center = 0);
int k_layernorm_bf16in(cudaStream_t st, const __nv_bfloat16* in, long long in_stride, int rows, int C, const float* gamma,
                       const float* beta, __nv_bfloat16* out_bf16, long long out_stride, int reverse = 0);
// softmax(Q K^T / 8) V for 12 heads x 64 dims over 197 tokens per frame (clip_fsar.py:626,638), bf16 tensor cores.
//   qkv [F*197, 2304] bf16 (q | k | v, head h at columns h*64) -> out [F*197, 768] bf16
int k_vit_attention(cudaStream_t st, const __nv_bfloat16* qkv, __nv_bfloat16* out, int n_frames);
int k_vit_attention_init();
// same contract on tcgen05/TMEM (vit_attention_tc.cu): S and O accumulate in tensor memory, P feeds the second
// MMA straight from TMEM.  `sms` = number of SMs (persistent grid).
// reverse != 0: items / rows / tiles are visited in descending order (same results; used to alternate the sweep
// direction of consecutive kernels so that each one starts on the rows its producer wrote last, still in L2)
int k_vit_attention_tc(cudaStream_t st, const __nv_bfloat16* qkv, __nv_bfloat16* out, int n_frames, int sms,
                       int reverse = 0);
int k_vit_attention_tc_init();

// fp32 parity-mode kernels (sgemm_f32.cu)
int k_vit_attention_f32(cudaStream_t st, const float* qkv, float* out, int n_frames);
int k_vit_attention_f32_init();
int k_patch_im2col_f32(cudaStream_t st, const float* images, float* patches, int n_frames);

// decoded RGB frames [F,H,W,3] uint8 -> Resize(256) / CenterCrop(224) / ToTensor (frame_transform.cu): exactly one of
// out_f32 [F,3,224,224] and out_patch (the bf16 patch matrix of k_patch_im2col) is non-null
int k_frame_transform(cudaStream_t st, const uint8_t* frames, int n_frames, int H, int W, float* out_f32,
                      __nv_bfloat16* out_patch);

}  // namespace spm
