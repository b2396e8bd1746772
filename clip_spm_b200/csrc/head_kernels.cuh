// Launchers of the metric-head kernels (head_kernels.cu, otam.cu).  Return 0 on success, a cudaError_t value on a
// launch failure, or a negative number for an unsupported shape.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spm {

int k_temporal_im2col(cudaStream_t st, const float* x, long long vid_stride, int V, int T, int D, float* out);
int k_motion_reduce(cudaStream_t st, const float* conv, const float* x, long long vid_stride, int V, int T, int D,
                    float* out);
int k_mo_dist(cudaStream_t st, const float* new_m, const float* z_tok, long long tok_stride, int E, int S, int Q, int D,
              const float* mo_alpha1, float* dists);
int k_token_prepare(cudaStream_t st, const float* text, int n_cls, const float* real_support, const float* real_target,
                    const float* X, int E, int S, int Q, int T, int D, float* tok_b, float* tt_in, int* err_flag);
int k_gather4(cudaStream_t st, const float* src, int n0, int n1, int n2, int n3, long long s0, long long s1,
              long long s2, float* out);
int k_seq_build(cudaStream_t st, const float* tok, const float* gt, const float* X, const float* gv, int n_calls,
                int V, int T, int D, float alpha, float* seq);
int k_seq_attention_init();
int k_seq_attention(cudaStream_t st, const float* qkv, float* out, int n_batch, int rows_per_batch, int n_groups,
                    int off0, int len0, int off1, int len1, int heads, int dh);
int k_padm_build(cudaStream_t st, const float* z, const float* labels, int E, int S, int Q, int W, int T, int D,
                 float* su_pro, float* seq1, int* err_flag);
int k_class_mean_padm(cudaStream_t st, const float* z1, const float* labels, int E, int S, int Q, int W, int T, int D,
                      float* su_pro2);
int k_finalize(cudaStream_t st, const float* accd, const float* d3, int E, int Q, int W, const long long* target,
               float tasks_per_batch, const float* dists, float* logits, float* loss, float* accuracy, int* pred,
               const int* err_flag);
// ---- sibling head CLIP-FSAR (sibling_heads.cu; models/model_clipfsar.py:325-383)
int k_fsar_seq_build(cudaStream_t st, const float* X, const float* text, int n_cls, const float* real_s, int E, int S,
                     int Q, int T, int D, float* seq);
int k_fsar_merge_seq_build(cudaStream_t st, const float* X, const float* text, int n_cls, const float* labels,
                           const float* real_s, int E, int S, int Q, int W, int T, int D, float* seq, int* err_flag);
int k_fsar_class_mean(cudaStream_t st, const float* z, const float* labels, int E, int S, int W, int T, int D,
                      float* su_pro, int* err_flag);
int k_fsar_class_logits(cudaStream_t st, const float* X, const float* text_train, int n_cls, const float* scale, int V,
                        int T, int D, float* out);
int k_fsar_class_ce_add(cudaStream_t st, const float* cls, const float* real_s, const float* real_t, int E, int S, int Q,
                        int n_cls, float coef, float* loss);
// sibling head STEN as shipped (models/model_sten.py:62-113): frame means, class-mean prototypes of features and
// prompts, product of the two cosine softmaxes; negsim = -logits (k_finalize negates)
int k_sten_head(cudaStream_t st, const float* X, const float* text, int n_cls, const float* labels,
                const float* real_s, int E, int S, int Q, int W, int T, int D, float* frame_mean, float* proto,
                float* negsim, int* err_flag);
// ---- sibling head CPM2C (cpm2c_kernels.cu; models/model_cpm2c.py:207-312); Z = context2 outputs [2][V][L][D]
int k_temporal_im2col_dil(cudaStream_t st, const float* x, int V, int T, int D, int dil, float* out);
int k_cpm2c_motion_diff(cudaStream_t st, const float* conv, const float* x, int V, int T, int D, float* out);
int k_cpm2c_tokens(cudaStream_t st, const float* text, int n_cls, const float* real_s, const float* real_t,
                   const float* cls_token, int E, int S, int Q, int D, float* tok, int* err_flag);
int k_cpm2c_class_mean(cudaStream_t st, const float* z, const float* labels, int E, int S, int Q, int W, int L, int D,
                       float* su_pro, int* err_flag);
int k_cpm2c_consist(cudaStream_t st, const float* z, int E, int S, int Q, int L, int D, float coeff, float beta, float* consist);
int k_cpm2c_global(cudaStream_t st, const float* z, const float* labels, int E, int S, int Q, int W, int L, int D,
                   float coeff, float beta, float* g);
int k_cpm2c_finalize(cudaStream_t st, const float* loc, const float* glob, const float* cls, int n_cls, const float* real_s,
                     const float* real_t, int E, int S, int Q, int W, const long long* target, float l0, float l1, float l2,
                     float tasks_per_batch, float* out_local, float* out_global, float* out_total, float* loss,
                     float* accuracy, int* pred, const int* err_flag);
// soft-DTW of TA2N (softdtw.cu; models/OTAM.py:34-203): D [B,N,M] -> R [B,N+2,M+2], out [B]; backward -> E [B,N,M]
int k_softdtw_forward(cudaStream_t st, const float* D, int B, int N, int M, float gamma, float bandwidth, float* R,
                      float* out);
int k_softdtw_backward(cudaStream_t st, const float* D, const float* R, int B, int N, int M, float gamma,
                       float bandwidth, float* E);
int k_otam_init();
// out[p,q,w] = beta*out + alpha * otam(support[p,w,:,:], target[p,q,:,:]); element (p,w,t,d) of the support set is at
// sup + p*s_p + w*s_w + t*s_t + d (strides in floats), likewise for the target set.
int k_otam(cudaStream_t st, const float* sup, long long s_p, long long s_w, long long s_t, const float* tgt,
           long long t_p, long long t_q, long long t_t, int P, int W, int Q, int T, int D, int single_direct,
           float alpha, float beta, float* out);
// same contract on the tensor cores (3xTF32 products, one CTA per problem at batch scale); -3 = shape not instantiated
int k_otam_mma(cudaStream_t st, const float* sup, long long s_p, long long s_w, long long s_t, const float* tgt,
               long long t_p, long long t_q, long long t_t, int P, int W, int Q, int T, int D, int single_direct,
               float alpha, float beta, float* out);

// same contract as ONE persistent warp-specialised kernel (TMA ring across problems, 3xTF32 products, wavefronts on their
// own warps; otam_fused.cu) for P >= 2 x #SM problems of the headline shapes; -3 = outside its envelope
int k_otam_fused(cudaStream_t st, const float* sup, long long s_p, long long s_w, long long s_t, const float* tgt,
                 long long t_p, long long t_q, long long t_t, int P, int W, int Q, int T, int D, int single_direct,
                 float alpha, float beta, float* out);


// same contract on the tcgen05 tensor cores (otam_tc.cu): three problems stacked in a 128-row tile, 3xTF32 products into
// tensor memory, for P >= 2 x #SM problems with 3*Q*T <= 128 (the headline 5-way, T = 8 shapes); -3 = outside its envelope
int k_otam_tc(cudaStream_t st, const float* sup, long long s_p, long long s_w, long long s_t, const float* tgt, long long t_p,
              long long t_q, long long t_t, int P, int W, int Q, int T, int D, int single_direct, float alpha, float beta,
              float* out);

}  // namespace spm
