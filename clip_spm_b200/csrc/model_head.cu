// Metric-head side of the handle: weight packing, workspace, cached plans and the kernel sequences of the CLIP-SPM head
// (models/model_clipspm.py:116-143) and of the sibling heads CLIP-FSAR (models/model_clipfsar.py:325-383) and STEN
// (models/model_sten.py:62-113), all behind head_run.
#include "model_internal.cuh"

namespace spm {
namespace detail {

// one Transformer_v1 layer (models/myRes.py:1053-1064): LayerNorm, to_q/k/v fused into one [3*inner, D] B operand
// (myRes.py:957-959), to_out, FeedForward
int load_ctx(spm_handle* h, cudaStream_t st, const WeightTable& wt, const std::string& p, long long inner, CtxW* out) {
  const long long D = h->D;
  CtxW& x = *out;
  const float* src;
  SPM_TRY(copy_f32(h, st, wt, p + "0.norm.weight", D, &x.ln_g));
  SPM_TRY(copy_f32(h, st, wt, p + "0.norm.bias", D, &x.ln_b));
  SPM_TRY(dalloc_t(h, &x.qkv_w, 3LL * inner * D));
  const char* names[3] = {"0.fn.to_q.weight", "0.fn.to_k.weight", "0.fn.to_v.weight"};
  for (int i = 0; i < 3; ++i) {
    SPM_TRY(wt.get(p + names[i], inner * D, &src));
    SPM_CUDA(cudaMemcpyAsync(x.qkv_w + (long long)i * inner * D, src, (size_t)inner * D * 4, cudaMemcpyDeviceToDevice, st));
  }
  SPM_TRY(copy_f32(h, st, wt, p + "0.fn.to_out.0.weight", D * inner, &x.out_w));
  SPM_TRY(copy_f32(h, st, wt, p + "0.fn.to_out.0.bias", D, &x.out_b));
  SPM_TRY(copy_f32(h, st, wt, p + "1.net.0.weight", HEAD_MLP * D, &x.ff0_w));
  SPM_TRY(copy_f32(h, st, wt, p + "1.net.0.bias", HEAD_MLP, &x.ff0_b));
  SPM_TRY(copy_f32(h, st, wt, p + "1.net.3.weight", D * HEAD_MLP, &x.ff3_w));
  SPM_TRY(copy_f32(h, st, wt, p + "1.net.3.bias", D, &x.ff3_b));
  return 0;
}

// CNN_OTAM_CLIPFSAR's own parameters (models/model_clipfsar.py:137-145): scale, context2 with inner width D
int load_head_fsar(spm_handle* h, cudaStream_t st, const WeightTable& wt) {
  SPM_TRY(copy_f32(h, st, wt, "scale", 1, &h->fsar_scale));
  SPM_TRY(load_ctx(h, st, wt, "context2.layers.0.", h->D, &h->fsar_ctx));
  h->fsar_ctx_more.assign(std::max(h->cfg.fsar_depth, 1) - 1, CtxW{});
  for (size_t i = 0; i < h->fsar_ctx_more.size(); ++i)
    SPM_TRY(load_ctx(h, st, wt, "context2.layers." + std::to_string(i + 1) + ".", h->D, &h->fsar_ctx_more[i]));
  return 0;
}

// CLIP_CPMMC_FSAR's parameters the forward reads (models/model_cpm2c.py:73-141): scale, context2 (inner width D), the
// gates, the two class tokens, the multi-scale motion convolutions (k = 1, 3, 3 dilated) and their 1x1 fusion, whose
// `* motion_residual_ratio` (:175) is folded into its weights and bias here.
__global__ void scale_copy_kernel(const float* __restrict__ in, float s, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * s;
}
int load_head_cpm2c(spm_handle* h, cudaStream_t st, const WeightTable& wt) {
  HeadW& w = h->head;
  spm_handle::Cpm2cW& c = h->cpm;
  const long long D = h->D, HT = h->HT, HV = h->HV;
  SPM_TRY(copy_f32(h, st, wt, "scale", 1, &h->fsar_scale));
  SPM_TRY(load_ctx(h, st, wt, "context2.layers.0.", D, &h->fsar_ctx));
  SPM_TRY(copy_f32(h, st, wt, "gate_text.0.weight", HT * D, &w.gt0_w));
  SPM_TRY(copy_f32(h, st, wt, "gate_text.0.bias", HT, &w.gt0_b));
  SPM_TRY(copy_f32(h, st, wt, "gate_text.2.weight", D * HT, &w.gt2_w));
  SPM_TRY(copy_f32(h, st, wt, "gate_text.2.bias", D, &w.gt2_b));
  SPM_TRY(copy_f32(h, st, wt, "gate_vision.0.weight", HV * D, &w.gv0_w));
  SPM_TRY(copy_f32(h, st, wt, "gate_vision.0.bias", HV, &w.gv0_b));
  SPM_TRY(copy_f32(h, st, wt, "gate_vision.2.weight", D * HV, &w.gv2_w));
  SPM_TRY(copy_f32(h, st, wt, "gate_vision.2.bias", D, &w.gv2_b));
  SPM_TRY(copy_f32(h, st, wt, "class_token", D, &c.cls_tok));
  SPM_TRY(copy_f32(h, st, wt, "class_token_motion", D, &c.cls_tok_motion));
  SPM_TRY(copy_f32(h, st, wt, "motion_conv1_1.weight", D * D, &c.m1_w));   // [D, D, 1] is already the GEMM's B operand
  SPM_TRY(copy_f32(h, st, wt, "motion_conv1_1.bias", D, &c.m1_b));
  const float* src;
  SPM_TRY(wt.get("motion_conv1_3.weight", D * D * 3, &src));
  SPM_TRY(dalloc_t(h, &c.m3_w, D * D * 3));
  SPM_KERNEL(k_repack_conv1d(st, src, c.m3_w, (int)D, (int)D));
  SPM_TRY(copy_f32(h, st, wt, "motion_conv1_3.bias", D, &c.m3_b));
  SPM_TRY(wt.get("motion_conv1_5.weight", D * D * 3, &src));
  SPM_TRY(dalloc_t(h, &c.m5_w, D * D * 3));
  SPM_KERNEL(k_repack_conv1d(st, src, c.m5_w, (int)D, (int)D));
  SPM_TRY(copy_f32(h, st, wt, "motion_conv1_5.bias", D, &c.m5_b));
  const float ratio = h->cfg.motion_residual_ratio;
  SPM_TRY(wt.get("scale_conv.weight", D * 3 * D, &src));
  SPM_TRY(dalloc_t(h, &c.sc_w, D * 3 * D));
  scale_copy_kernel<<<(unsigned)((D * 3 * D + 255) / 256), 256, 0, st>>>(src, ratio, c.sc_w, D * 3 * D);
  count_launch();
  SPM_TRY(wt.get("scale_conv.bias", D, &src));
  SPM_TRY(dalloc_t(h, &c.sc_b, D));
  scale_copy_kernel<<<(unsigned)((D + 255) / 256), 256, 0, st>>>(src, ratio, c.sc_b, D);
  count_launch();
  SPM_CUDA(cudaGetLastError());
  return 0;
}

int load_head(spm_handle* h, cudaStream_t st, const WeightTable& wt) {
  HeadW& w = h->head;
  const long long D = h->D, HT = h->HT, HV = h->HV;
  const float* src;
  SPM_TRY(wt.get("motion_conv1.weight", D * D * 3, &src));
  SPM_TRY(dalloc_t(h, &w.mc1_w, D * D * 3));
  SPM_KERNEL(k_repack_conv1d(st, src, w.mc1_w, (int)D, (int)D));
  SPM_TRY(wt.get("motion_conv2.weight", D * D * 3, &src));
  SPM_TRY(dalloc_t(h, &w.mc2_w, D * D * 3));
  SPM_KERNEL(k_repack_conv1d(st, src, w.mc2_w, (int)D, (int)D));
  SPM_TRY(copy_f32(h, st, wt, "motion_conv1.bias", D, &w.mc1_b));
  SPM_TRY(copy_f32(h, st, wt, "motion_conv2.bias", D, &w.mc2_b));
  SPM_TRY(copy_f32(h, st, wt, "token_tr.mlp.net.0.weight", HEAD_MLP * D, &w.tt0_w));
  SPM_TRY(copy_f32(h, st, wt, "token_tr.mlp.net.0.bias", HEAD_MLP, &w.tt0_b));
  SPM_TRY(copy_f32(h, st, wt, "token_tr.mlp.net.3.weight", D * HEAD_MLP, &w.tt3_w));
  SPM_TRY(copy_f32(h, st, wt, "token_tr.mlp.net.3.bias", D, &w.tt3_b));
  SPM_TRY(copy_f32(h, st, wt, "gate_text.0.weight", HT * D, &w.gt0_w));
  SPM_TRY(copy_f32(h, st, wt, "gate_text.0.bias", HT, &w.gt0_b));
  SPM_TRY(copy_f32(h, st, wt, "gate_text.2.weight", D * HT, &w.gt2_w));
  SPM_TRY(copy_f32(h, st, wt, "gate_text.2.bias", D, &w.gt2_b));
  SPM_TRY(copy_f32(h, st, wt, "gate_vision.0.weight", HV * D, &w.gv0_w));
  SPM_TRY(copy_f32(h, st, wt, "gate_vision.0.bias", HV, &w.gv0_b));
  SPM_TRY(copy_f32(h, st, wt, "gate_vision.2.weight", D * HV, &w.gv2_w));
  SPM_TRY(copy_f32(h, st, wt, "gate_vision.2.bias", D, &w.gv2_b));
  SPM_TRY(copy_f32(h, st, wt, "mo_alpha1", 1, &w.mo_alpha1));
  SPM_TRY(load_ctx(h, st, wt, "context1.layers.0.", HEAD_INNER, &w.ctx[0]));
  SPM_TRY(load_ctx(h, st, wt, "context2.layers.0.", HEAD_INNER, &w.ctx[1]));
  return 0;
}


// ---------------------------------------------------------------------------------------------------------
// metric head
// ---------------------------------------------------------------------------------------------------------
int ensure_head_workspace(spm_handle* h, int E, int S, int Q, int W) {
  if (E <= h->head_cap_E && S <= h->head_cap_S && Q <= h->head_cap_Q && W <= h->head_cap_W) return 0;
  // grow-only: plans that point into the old buffers are dropped
  h->head_plans.clear();
  h->fsar_plans.clear();
  h->cpm2c_plans.clear();
  const long long cE = std::max<long long>(E, h->head_cap_E), cS = std::max<long long>(S, h->head_cap_S),
                  cQ = std::max<long long>(Q, h->head_cap_Q), cW = std::max<long long>(W, h->head_cap_W);
  const long long T = h->cfg.seq_len, D = h->D, N = cS + cQ, V = cE * N;
  const long long R2 = 2 * V * (T + 1), R1 = cE * T * (cW + cS + 1 + cQ), R = std::max(R1, R2);
  SPM_TRY(drealloc_t(h, &h->Xhead, V * T * D));
  h->X = h->Xhead;
  SPM_TRY(drealloc_t(h, &h->XC, V * T * 3 * D));
  SPM_TRY(drealloc_t(h, &h->C1, V * T * D));
  SPM_TRY(drealloc_t(h, &h->C2, V * T * D));
  SPM_TRY(drealloc_t(h, &h->TOK, 2 * V * D));
  SPM_TRY(drealloc_t(h, &h->TTIN, cE * cQ * D));
  SPM_TRY(drealloc_t(h, &h->TTH, cE * cQ * HEAD_MLP));
  SPM_TRY(drealloc_t(h, &h->GTH, 2 * V * h->HT));
  SPM_TRY(drealloc_t(h, &h->GT, 2 * V * D));
  SPM_TRY(drealloc_t(h, &h->GVH, V * T * h->HV));
  SPM_TRY(drealloc_t(h, &h->GV, V * T * D));
  SPM_TRY(drealloc_t(h, &h->SEQ, R * D));
  SPM_TRY(drealloc_t(h, &h->HN, R * D));
  SPM_TRY(drealloc_t(h, &h->QKVH, R * 3 * HEAD_INNER));
  SPM_TRY(drealloc_t(h, &h->AO, R * HEAD_INNER));
  SPM_TRY(drealloc_t(h, &h->Y, R * D));
  SPM_TRY(drealloc_t(h, &h->FFH, R * HEAD_MLP));
  SPM_TRY(drealloc_t(h, &h->Z, R2 * D));
  SPM_TRY(drealloc_t(h, &h->Z1, R1 * D));
  SPM_TRY(drealloc_t(h, &h->NEWM, V * D));
  SPM_TRY(drealloc_t(h, &h->SUPRO, cE * cW * T * D));
  SPM_TRY(drealloc_t(h, &h->SUPRO2, cE * cW * T * D));
  SPM_TRY(drealloc_t(h, &h->ACC, cE * cQ * cW));
  SPM_TRY(drealloc_t(h, &h->D3, cE * cW));
  if (h->err_flag == nullptr) {
    SPM_TRY(dalloc_t(h, &h->err_flag, 1));
    SPM_CUDA(cudaMemset(h->err_flag, 0, sizeof(int)));
  }
  h->head_cap_E = cE; h->head_cap_S = cS; h->head_cap_Q = cQ; h->head_cap_W = cW;
  return 0;
}

// inner = heads * dim_head of the attention (2048 for CLIP-SPM's context1/2, D for CLIP-FSAR's context2)
int plan_ctx(spm_handle* h, CtxPlan* p, const CtxW& w, int R, float* seq, float* out, int inner = HEAD_INNER) {
  const int D = h->D;
  GemmEpilogue e1;
  e1.out = h->QKVH; e1.ldo = 3 * inner;
  SPM_TRY(plan_gemm(&p->qkv, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, h->HN, D, w.qkv_w, D, R, 3 * inner, D, e1, h->sms));
  GemmEpilogue e2;  // to_out + bias + the un-normalised sequence (myRes.py:1040)
  e2.bias = w.out_b; e2.residual = seq; e2.ldr = D; e2.out = h->Y; e2.ldo = D;
  SPM_TRY(plan_gemm(&p->outp, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, h->AO, inner, w.out_w, inner, R, D, inner, e2, h->sms));
  GemmEpilogue e3;
  e3.bias = w.ff0_b; e3.act = ACT_GELU_ERF; e3.out = h->FFH; e3.ldo = HEAD_MLP;
  SPM_TRY(plan_gemm(&p->ff0, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, h->Y, D, w.ff0_w, D, R, HEAD_MLP, D, e3, h->sms));
  GemmEpilogue e4;  // x = ff(x) + x (myRes.py:1069)
  e4.bias = w.ff3_b; e4.residual = h->Y; e4.ldr = D; e4.out = out; e4.ldo = D;
  SPM_TRY(plan_gemm(&p->ff3, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, h->FFH, HEAD_MLP, w.ff3_w, HEAD_MLP, R, D, HEAD_MLP, e4, h->sms));
  return 0;
}

int get_head_plan(spm_handle* h, int E, int S, int Q, int W, HeadPlan** out) {
  SPM_TRY(ensure_head_workspace(h, E, S, Q, W));
  for (auto& p : h->head_plans)
    if (p->E == E && p->S == S && p->Q == Q && p->W == W && p->X == h->X) { *out = p.get(); return 0; }
  std::unique_ptr<HeadPlan> pl(new HeadPlan());
  pl->E = E; pl->S = S; pl->Q = Q; pl->W = W; pl->X = h->X;
  const HeadW& w = h->head;
  const int T = h->cfg.seq_len, D = h->D, N = S + Q, V = E * N;
  {
    GemmEpilogue e;
    e.bias = w.mc1_b; e.out = h->C1; e.ldo = D;
    SPM_TRY(plan_gemm(&pl->mc1, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, h->XC, 3 * D, w.mc1_w, 3 * D, V * T, D, 3 * D, e, h->sms));
    e.bias = w.mc2_b; e.out = h->C2;
    SPM_TRY(plan_gemm(&pl->mc2, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, h->XC, 3 * D, w.mc2_w, 3 * D, V * T, D, 3 * D, e, h->sms));
  }
  {
    GemmEpilogue e;
    e.bias = w.tt0_b; e.act = ACT_GELU_ERF; e.out = h->TTH; e.ldo = HEAD_MLP;
    SPM_TRY(plan_gemm(&pl->tt0, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, h->TTIN, D, w.tt0_w, D, E * Q, HEAD_MLP, D, e, h->sms));
    GemmEpilogue e2;  // query tokens of the `sem` call land after the S support tokens of their episode
    e2.bias = w.tt3_b; e2.out = h->TOK + (long long)V * D; e2.ldo = D;
    e2.out_row_group = Q; e2.out_group_stride = N; e2.out_row_off = S;
    SPM_TRY(plan_gemm(&pl->tt3, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, h->TTH, HEAD_MLP, w.tt3_w, HEAD_MLP, E * Q, D, HEAD_MLP, e2, h->sms));
  }
  {
    GemmEpilogue e;
    e.bias = w.gt0_b; e.act = ACT_LEAKY; e.slope = h->cfg.negative_slope; e.out = h->GTH; e.ldo = h->HT;
    SPM_TRY(plan_gemm(&pl->gt0, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, h->TOK, D, w.gt0_w, D, 2 * V, h->HT, D, e, h->sms));
    GemmEpilogue e2;
    e2.bias = w.gt2_b; e2.act = ACT_SIGMOID; e2.out = h->GT; e2.ldo = D;
    SPM_TRY(plan_gemm(&pl->gt2, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, h->GTH, h->HT, w.gt2_w, h->HT, 2 * V, D, h->HT, e2, h->sms));
    GemmEpilogue e3;
    e3.bias = w.gv0_b; e3.act = ACT_LEAKY; e3.slope = h->cfg.negative_slope; e3.out = h->GVH; e3.ldo = h->HV;
    SPM_TRY(plan_gemm(&pl->gv0, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, h->X, D, w.gv0_w, D, V * T, h->HV, D, e3, h->sms));
    GemmEpilogue e4;
    e4.bias = w.gv2_b; e4.act = ACT_SIGMOID; e4.out = h->GV; e4.ldo = D;
    SPM_TRY(plan_gemm(&pl->gv2, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, h->GVH, h->HV, w.gv2_w, h->HV, V * T, D, h->HV, e4, h->sms));
  }
  SPM_TRY(plan_ctx(h, &pl->c2, w.ctx[1], 2 * V * (T + 1), h->SEQ, h->Z));
  SPM_TRY(plan_ctx(h, &pl->c1, w.ctx[0], E * T * (W + S + 1 + Q), h->SEQ, h->Z1));
  *out = pl.get();
  h->head_plans.push_back(std::move(pl));
  return 0;
}

int run_ctx(spm_handle* h, cudaStream_t st, const CtxPlan& p, const CtxW& w, int R, int n_batch, int rows_per_batch,
            int n_groups, int off0, int len0, int off1, int len1) {
  const int D = h->D;
  SPM_KERNEL(k_layernorm(st, h->SEQ, D, R, D, w.ln_g, w.ln_b, nullptr, 0, h->HN, nullptr, D));
  SPM_GEMM_RUN(p.qkv);
  SPM_KERNEL(k_seq_attention(st, h->QKVH, h->AO, n_batch, rows_per_batch, n_groups, off0, len0, off1, len1,
                             HEAD_HEADS, HEAD_DH));
  SPM_GEMM_RUN(p.outp);
  SPM_GEMM_RUN(p.ff0);
  SPM_GEMM_RUN(p.ff3);
  return 0;
}

// STEN head as shipped (models/model_sten.py:62-113) on frame features in h->X [E, N, T, D]; no learned parameters
int sten_head_run(spm_handle* h, cudaStream_t st, int E, int S, int Q, int W, const float* labels, const float* real_s,
                  const long long* target_labels, float tasks_per_batch, float* logits, float* dists, float* loss,
                  float* acc, int* pred) {
  SPM_CHECK(h->text_set, "head: text features not set (spm_set_text_features)");
  SPM_TRY(ensure_head_workspace(h, E, S, Q, W));
  const int T = h->cfg.seq_len, D = h->D;
  // scratch: NEWM [V, D] frame means, SUPRO [E, W, T, D] >= [E, W, 2, D] prototypes, ACC [E, Q, W]
  SPM_KERNEL(k_sten_head(st, h->X, h->text, h->n_cls, labels, real_s, E, S, Q, W, T, D, h->NEWM, h->SUPRO, h->ACC,
                         h->err_flag));
  SPM_CUDA(cudaMemsetAsync(h->D3, 0, (size_t)E * W * sizeof(float), st));
  SPM_CUDA(cudaMemsetAsync(dists, 0, (size_t)E * sizeof(float), st));
  SPM_KERNEL(k_finalize(st, h->ACC, h->D3, E, Q, W, target_labels, tasks_per_batch, dists, logits, loss, acc, pred,
                        h->err_flag));
  return 0;
}

// CLIP-FSAR head (models/model_clipfsar.py:325-383) on frame features in h->X [E, N, T, D]:
//   target  = context2(target)                       self-attention over the T frames of each query video
//   support = context2(cat[support, prompt])[:, :T]  T frames + the class prompt of the video's real label
//   prototypes = per-class mean;  logits = -(OTAM(d) + OTAM(d^T));  class_logits = cos_sim(mean_t feats, text_train)*scale
// cfg.fsar_merge_before (:341-349): the per-class means move in front of context2 (E*W support sequences, their outputs are
// the prototypes); cfg.fsar_depth (:143-144): context2 has that many layers, run back to back between SEQ and Z.
int fsar_head_run(spm_handle* h, cudaStream_t st, int E, int S, int Q, int W, const float* labels, const float* real_s,
                  const float* real_t, const long long* target_labels, float tasks_per_batch, float* logits,
                  float* dists, float* loss, float* acc, int* pred) {
  SPM_CHECK(h->text_set, "head: text features not set (spm_set_text_features)");
  SPM_TRY(ensure_head_workspace(h, E, S, Q, W));
  const int T = h->cfg.seq_len, D = h->D, N = S + Q, V = E * N, dh = D / HEAD_HEADS;
  const bool merge = h->cfg.fsar_merge_before != 0;
  const int depth = 1 + (int)h->fsar_ctx_more.size();
  const int SB = merge ? W : S;   // support sequences per episode that go through context2
  const long long RS = (long long)E * SB * (T + 1), R = RS + (long long)E * Q * T, TD = (long long)T * D;
  FsarPlan* pl = nullptr;
  for (auto& p : h->fsar_plans)
    if (p->E == E && p->S == S && p->Q == Q && p->W == W) pl = p.get();
  if (pl == nullptr) {
    std::unique_ptr<FsarPlan> np(new FsarPlan());
    np->E = E; np->S = S; np->Q = Q; np->W = W;
    np->c2.resize(depth);
    for (int l = 0; l < depth; ++l)   // layer l reads buf[l % 2], writes buf[(l + 1) % 2]  (buf = {SEQ, Z})
      SPM_TRY(plan_ctx(h, &np->c2[l], l == 0 ? h->fsar_ctx : h->fsar_ctx_more[l - 1], (int)R, l % 2 ? h->Z : h->SEQ,
                       l % 2 ? h->SEQ : h->Z, D));
    pl = np.get();
    h->fsar_plans.push_back(std::move(np));
  }
  if (merge)
    SPM_KERNEL(k_fsar_merge_seq_build(st, h->X, h->text, h->n_cls, labels, real_s, E, S, Q, W, T, D, h->SEQ, h->err_flag));
  else
    SPM_KERNEL(k_fsar_seq_build(st, h->X, h->text, h->n_cls, real_s, E, S, Q, T, D, h->SEQ));
  for (int l = 0; l < depth; ++l) {
    const CtxW& w = l == 0 ? h->fsar_ctx : h->fsar_ctx_more[l - 1];
    SPM_KERNEL(k_layernorm(st, l % 2 ? h->Z : h->SEQ, D, (int)R, D, w.ln_g, w.ln_b, nullptr, 0, h->HN, nullptr, D));
    SPM_GEMM_RUN(pl->c2[l].qkv);
    SPM_KERNEL(k_seq_attention(st, h->QKVH, h->AO, E * SB, T + 1, 1, 0, T + 1, 0, 0, HEAD_HEADS, dh));
    SPM_KERNEL(k_seq_attention(st, h->QKVH + RS * 3 * D, h->AO + RS * D, E * Q, T, 1, 0, T, 0, 0, HEAD_HEADS, dh));
    SPM_GEMM_RUN(pl->c2[l].outp);
    SPM_GEMM_RUN(pl->c2[l].ff0);
    SPM_GEMM_RUN(pl->c2[l].ff3);
  }
  const float* Zout = depth % 2 ? h->Z : h->SEQ;
  if (merge) {   // the class sequences' frame rows are the prototypes: rows 1..T of each (T+1)-row sequence
    SPM_KERNEL(k_otam(st, Zout + D, (long long)W * (T + 1) * D, (long long)(T + 1) * D, D, Zout + RS * D,
                      (long long)Q * TD, TD, D, E, W, Q, T, D, h->cfg.single_direct, 1.f, 0.f, h->ACC));
  } else {
    SPM_KERNEL(k_fsar_class_mean(st, Zout, labels, E, S, W, T, D, h->SUPRO, h->err_flag));
    SPM_KERNEL(k_otam(st, h->SUPRO, (long long)W * TD, TD, D, Zout + RS * D, (long long)Q * TD, TD, D, E, W, Q, T, D,
                      h->cfg.single_direct, 1.f, 0.f, h->ACC));
  }
  SPM_CUDA(cudaMemsetAsync(h->D3, 0, (size_t)E * W * sizeof(float), st));
  SPM_CUDA(cudaMemsetAsync(dists, 0, (size_t)E * sizeof(float), st));   // this head has no auxiliary distance
  h->cls_rows = 0;
  if (h->text_train != nullptr) {
    const long long need = (long long)V * h->n_cls_train;
    if (need > h->cls_cap) {
      SPM_TRY(drealloc_t(h, &h->CLS, need));
      h->cls_cap = need;
    }
    SPM_KERNEL(k_fsar_class_logits(st, h->X, h->text_train, h->n_cls_train, h->fsar_scale, V, T, D, h->CLS));
    h->cls_rows = V;
  }
  SPM_KERNEL(k_finalize(st, h->ACC, h->D3, E, Q, W, target_labels, tasks_per_batch, dists, logits, loss, acc, pred,
                        h->err_flag));
  if (loss != nullptr && target_labels != nullptr) {
    // run/main_run.py:355-356: (CE(logits) + USE_CLASSIFICATION_VALUE * CE(class_logits, real labels)) / TASKS_PER_BATCH
    SPM_CHECK(h->text_train != nullptr, "CLIP-FSAR loss needs text_features_train (spm_set_text_features_train)");
    SPM_KERNEL(k_fsar_class_ce_add(st, h->CLS, real_s, real_t, E, S, Q, h->n_cls_train,
                                   h->cfg.cls_value / tasks_per_batch, loss));
  }
  return 0;
}

// CPM2C head (models/model_cpm2c.py:207-312, evaluation) on frame features in h->X [E, N, T, D]:
//   motion = multi-scale temporal convolutions + frame differences [V, T-1, D]
//   per branch (motion with class_token_motion, normal with class_token): gates, two context2 batches (sequences on the
//   real prompt / on the class token), class prototypes, consistency distance, global (token) and local (OTAM) distances
//   logits_out = lambdas1 * (-local) + lambdas2 * (-global); dists_out = target_consist_distance
int cpm2c_head_run(spm_handle* h, cudaStream_t st, int E, int S, int Q, int W, const float* labels, const float* real_s,
                   const float* real_t, const long long* target_labels, float tasks_per_batch, float* logits, float* dists,
                   float* loss, float* acc, int* pred) {
  SPM_CHECK(h->text_set, "head: text features not set (spm_set_text_features)");
  SPM_TRY(ensure_head_workspace(h, E, S, Q, W));
  const int T = h->cfg.seq_len, D = h->D, N = S + Q, V = E * N, dh = D / HEAD_HEADS;
  SPM_CHECK(T >= 3, "CPM2C head: seq_len must be at least 3 (its motion branch aligns T-1 frame differences)");
  const int kind = h->fp32 ? GEMM_F32_SIMT : GEMM_TF32;
  // ---- workspace of this head
  if (V > h->cp_cap_V) {
    SPM_TRY(drealloc_t(h, &h->CP_FCAT, (long long)V * T * 3 * D));
    SPM_TRY(drealloc_t(h, &h->CP_CONV, (long long)V * T * D));
    SPM_TRY(drealloc_t(h, &h->CP_MOT, (long long)V * T * D));
    SPM_TRY(drealloc_t(h, &h->CP_TOK, 2LL * V * D));
    SPM_TRY(drealloc_t(h, &h->CP_GT, 2LL * V * D));
    h->cp_cap_V = V;
    h->cpm2c_plans.clear();
  }
  if ((long long)E * Q * W > h->cp_cap_EQW) {
    for (float** p : {&h->CP_LOC, &h->CP_GLOB, &h->CP_OUT_L, &h->CP_OUT_G}) SPM_TRY(drealloc_t(h, p, (long long)E * Q * W));
    h->cp_cap_EQW = (long long)E * Q * W;
  }
  if ((long long)E * W > h->cp_cap_EW) {
    SPM_TRY(drealloc_t(h, &h->CP_PRO, (long long)E * W * (T + 1) * D));
    h->cp_cap_EW = (long long)E * W;
  }
  const long long need_cls = (long long)V * h->n_cls;
  if (need_cls > h->cls_cap) { SPM_TRY(drealloc_t(h, &h->CLS, need_cls)); h->cls_cap = need_cls; }
  // ---- plans
  Cpm2cPlan* pl = nullptr;
  for (auto& p : h->cpm2c_plans)
    if (p->E == E && p->S == S && p->Q == Q && p->X == h->X) pl = p.get();
  if (pl == nullptr) {
    std::unique_ptr<Cpm2cPlan> np(new Cpm2cPlan());
    np->E = E; np->S = S; np->Q = Q; np->X = h->X;
    const spm_handle::Cpm2cW& c = h->cpm;
    const HeadW& w = h->head;
    GemmEpilogue e;
    e.ldo = 3 * D;
    e.bias = c.m1_b; e.out = h->CP_FCAT;
    SPM_TRY(plan_gemm(&np->f1, kind, h->X, D, c.m1_w, D, V * T, D, D, e, h->sms));
    e.bias = c.m3_b; e.out = h->CP_FCAT + D;
    SPM_TRY(plan_gemm(&np->f3, kind, h->XC, 3 * D, c.m3_w, 3 * D, V * T, D, 3 * D, e, h->sms));
    e.bias = c.m5_b; e.out = h->CP_FCAT + 2 * D;
    SPM_TRY(plan_gemm(&np->f5, kind, h->XC, 3 * D, c.m5_w, 3 * D, V * T, D, 3 * D, e, h->sms));
    GemmEpilogue es;   // fused * ratio + residual (ratio folded into the weights)
    es.bias = c.sc_b; es.residual = h->X; es.ldr = D; es.out = h->CP_CONV; es.ldo = D;
    SPM_TRY(plan_gemm(&np->sc, kind, h->CP_FCAT, 3 * D, c.sc_w, 3 * D, V * T, D, 3 * D, es, h->sms));
    GemmEpilogue g0;
    g0.bias = w.gt0_b; g0.act = ACT_LEAKY; g0.slope = h->cfg.negative_slope; g0.out = h->GTH; g0.ldo = h->HT;
    SPM_TRY(plan_gemm(&np->gt0, kind, h->CP_TOK, D, w.gt0_w, D, 2 * V, h->HT, D, g0, h->sms));
    GemmEpilogue g2;
    g2.bias = w.gt2_b; g2.act = ACT_SIGMOID; g2.out = h->CP_GT; g2.ldo = D;
    SPM_TRY(plan_gemm(&np->gt2, kind, h->GTH, h->HT, w.gt2_w, h->HT, 2 * V, D, h->HT, g2, h->sms));
    for (int b = 0; b < 2; ++b) {   // branch 0: motion (T-1 frames of CP_MOT), branch 1: normal (T frames of X)
      const int Tp = b == 0 ? T - 1 : T;
      const float* F = b == 0 ? h->CP_MOT : h->X;
      GemmEpilogue v0;
      v0.bias = w.gv0_b; v0.act = ACT_LEAKY; v0.slope = h->cfg.negative_slope; v0.out = h->GVH; v0.ldo = h->HV;
      SPM_TRY(plan_gemm(&np->gv0[b], kind, F, D, w.gv0_w, D, V * Tp, h->HV, D, v0, h->sms));
      GemmEpilogue v2;
      v2.bias = w.gv2_b; v2.act = ACT_SIGMOID; v2.out = h->GV; v2.ldo = D;
      SPM_TRY(plan_gemm(&np->gv2[b], kind, h->GVH, h->HV, w.gv2_w, h->HV, V * Tp, D, h->HV, v2, h->sms));
      SPM_TRY(plan_ctx(h, &np->c2[b], h->fsar_ctx, 2 * V * (Tp + 1), h->SEQ, h->Z, D));
    }
    pl = np.get();
    h->cpm2c_plans.push_back(std::move(np));
  }
  const spm_handle::Cpm2cW& c = h->cpm;
  const CtxW& cw = h->fsar_ctx;
  const float nc = h->cfg.normal_coeff, mc = h->cfg.motion_coeff;
  // ---- multi-scale motion features (:165-199)
  SPM_GEMM_RUN(pl->f1);
  SPM_KERNEL(k_temporal_im2col_dil(st, h->X, V, T, D, 1, h->XC));
  SPM_GEMM_RUN(pl->f3);
  SPM_KERNEL(k_temporal_im2col_dil(st, h->X, V, T, D, 2, h->XC));
  SPM_GEMM_RUN(pl->f5);
  SPM_GEMM_RUN(pl->sc);
  SPM_KERNEL(k_cpm2c_motion_diff(st, h->CP_CONV, h->X, V, T, D, h->CP_MOT));
  // ---- class_text_logits on the raw frames (:222, :420-432)
  if (h->cfg.use_classification) {
    SPM_KERNEL(k_fsar_class_logits(st, h->X, h->text, h->n_cls, h->fsar_scale, V, T, D, h->CLS));
    h->cls_rows = V;
  } else {
    h->cls_rows = 0;
  }
  // ---- the two branches (:238-241): motion first, then normal
  for (int b = 0; b < 2; ++b) {
    const int Tp = b == 0 ? T - 1 : T, L = Tp + 1;
    const float* F = b == 0 ? h->CP_MOT : h->X;
    const float coeff = b == 0 ? mc : nc, beta = b == 0 ? 0.f : 1.f;
    const long long LD = (long long)L * D;
    SPM_KERNEL(k_cpm2c_tokens(st, h->text, h->n_cls, real_s, real_t, b == 0 ? c.cls_tok_motion : c.cls_tok, E, S, Q, D,
                              h->CP_TOK, h->err_flag));
    SPM_GEMM_RUN(pl->gt0);
    SPM_GEMM_RUN(pl->gt2);
    SPM_GEMM_RUN(pl->gv0[b]);
    SPM_GEMM_RUN(pl->gv2[b]);
    SPM_KERNEL(k_seq_build(st, h->CP_TOK, h->CP_GT, F, h->GV, 2, V, Tp, D, h->cfg.alpha, h->SEQ));
    const int R = 2 * V * L;
    SPM_KERNEL(k_layernorm(st, h->SEQ, D, R, D, cw.ln_g, cw.ln_b, nullptr, 0, h->HN, nullptr, D));
    SPM_GEMM_RUN(pl->c2[b].qkv);
    SPM_KERNEL(k_seq_attention(st, h->QKVH, h->AO, 2 * V, L, 1, 0, L, 0, 0, HEAD_HEADS, dh));
    SPM_GEMM_RUN(pl->c2[b].outp);
    SPM_GEMM_RUN(pl->c2[b].ff0);
    SPM_GEMM_RUN(pl->c2[b].ff3);
    SPM_KERNEL(k_cpm2c_class_mean(st, h->Z, labels, E, S, Q, W, L, D, h->CP_PRO, h->err_flag));
    SPM_KERNEL(k_cpm2c_consist(st, h->Z, E, S, Q, L, D, coeff, beta, dists));
    SPM_KERNEL(k_cpm2c_global(st, h->Z, labels, E, S, Q, W, L, D, coeff, beta, h->CP_GLOB));
    // local: OTAM over rows 1.. of the prototypes and of the fake-token query sequences (:292-296)
    SPM_KERNEL(k_otam(st, h->CP_PRO + D, (long long)W * LD, LD, D, h->Z + ((long long)V + S) * LD + D, (long long)N * LD, LD, D,
                      E, W, Q, Tp, D, h->cfg.single_direct, coeff, beta, h->CP_LOC));
  }
  SPM_KERNEL(k_cpm2c_finalize(st, h->CP_LOC, h->CP_GLOB, h->cfg.use_classification ? h->CLS : nullptr, h->n_cls, real_s, real_t,
                              E, S, Q, W, target_labels, h->cfg.lambdas[0], h->cfg.lambdas[1], h->cfg.lambdas[2],
                              tasks_per_batch, h->CP_OUT_L, h->CP_OUT_G, logits, loss, acc, pred, h->err_flag));
  h->cp_last_E = E; h->cp_last_Q = Q; h->cp_last_W = W;
  h->last_E = E; h->last_S = S; h->last_Q = Q; h->last_W = W;
  return 0;
}

// Frame features already in h->X as [E, N, T, D] (supports first).  Produces logits [E,Q,W], dists [E] and, when
// target_labels is given, loss / accuracy / predictions.
int head_run(spm_handle* h, cudaStream_t st, int E, int S, int Q, int W, const float* labels, const float* real_s,
             const float* real_t, const long long* target_labels, float tasks_per_batch, float* logits, float* dists,
             float* loss, float* acc, int* pred) {
  if (h->cfg.head == SPM_HEAD_STEN)
    return sten_head_run(h, st, E, S, Q, W, labels, real_s, target_labels, tasks_per_batch, logits, dists, loss, acc, pred);
  if (h->cfg.head == SPM_HEAD_CPM2C)
    return cpm2c_head_run(h, st, E, S, Q, W, labels, real_s, real_t, target_labels, tasks_per_batch, logits, dists, loss,
                          acc, pred);
  if (h->cfg.head == SPM_HEAD_CLIPFSAR)
    return fsar_head_run(h, st, E, S, Q, W, labels, real_s, real_t, target_labels, tasks_per_batch, logits, dists, loss,
                         acc, pred);
  SPM_CHECK(h->text_set, "head: text features not set (spm_set_text_features)");
  HeadPlan* pl;
  SPM_TRY(get_head_plan(h, E, S, Q, W, &pl));
  const HeadW& w = h->head;
  const int T = h->cfg.seq_len, D = h->D, N = S + Q, V = E * N, L1 = W + S + 1 + Q;
  const long long TD = (long long)T * D, T1D = (long long)(T + 1) * D;
  // ---- HSMR: motion features of the raw frames (model_clipspm.py:195)
  SPM_KERNEL(k_temporal_im2col(st, h->X, TD, V, T, D, h->XC));
  SPM_GEMM_RUN(pl->mc1);
  SPM_KERNEL(k_temporal_im2col(st, h->C1, TD, V, T, D, h->XC));
  SPM_GEMM_RUN(pl->mc2);
  SPM_KERNEL(k_motion_reduce(st, h->C2, h->X, TD, V, T, D, h->TOK));  // tokens of the `mo` se_te call
  // ---- SPM tokens (model_clipspm.py:120-121,213-216)
  SPM_KERNEL(k_token_prepare(st, h->text, h->n_cls, real_s, real_t, h->X, E, S, Q, T, D, h->TOK + (long long)V * D,
                             h->TTIN, h->err_flag));
  SPM_GEMM_RUN(pl->tt0);
  SPM_GEMM_RUN(pl->tt3);
  // ---- gates + the two live se_te batches (mo: tokens = motion; sem: tokens = prompts), one context2 pass
  SPM_GEMM_RUN(pl->gt0);
  SPM_GEMM_RUN(pl->gt2);
  SPM_GEMM_RUN(pl->gv0);
  SPM_GEMM_RUN(pl->gv2);
  SPM_KERNEL(k_seq_build(st, h->TOK, h->GT, h->X, h->GV, 2, V, T, D, h->cfg.alpha, h->SEQ));
  SPM_TRY(run_ctx(h, st, pl->c2, w.ctx[1], 2 * V * (T + 1), 2 * V, T + 1, 1, 0, T + 1, 0, 0));
  // ---- HSMR: motion of the refined frames vs the refined motion token (model_clipspm.py:200-205)
  SPM_KERNEL(k_temporal_im2col(st, h->Z + D, T1D, V, T, D, h->XC));
  SPM_GEMM_RUN(pl->mc1);
  SPM_KERNEL(k_temporal_im2col(st, h->C1, TD, V, T, D, h->XC));
  SPM_GEMM_RUN(pl->mc2);
  SPM_KERNEL(k_motion_reduce(st, h->C2, h->Z + D, T1D, V, T, D, h->NEWM));
  SPM_KERNEL(k_mo_dist(st, h->NEWM, h->Z, T1D, E, S, Q, D, w.mo_alpha1, dists));
  // ---- prototypes, class_dists_l, PADM sequences (model_clipspm.py:231-239,269,275-287)
  const float* Zb = h->Z + (long long)V * T1D;  // outputs of the `sem` call
  SPM_KERNEL(k_padm_build(st, Zb, labels, E, S, Q, W, T, D, h->SUPRO, h->SEQ, h->err_flag));
  SPM_KERNEL(k_otam(st, h->SUPRO, (long long)W * TD, TD, D, Zb + ((long long)S * (T + 1) + 1) * D, (long long)N * T1D,
                    T1D, D, E, W, Q, T, D, h->cfg.single_direct, 0.5f, 0.f, h->ACC));
  SPM_TRY(run_ctx(h, st, pl->c1, w.ctx[0], E * T * L1, E * T, L1, 2, 0, W + S, W + S, 1 + Q));
  // ---- task distances on the PADM outputs (model_clipspm.py:133-138)
  SPM_KERNEL(k_class_mean_padm(st, h->Z1, labels, E, S, Q, W, T, D, h->SUPRO2));
  const long long L1D = (long long)L1 * D;
  SPM_KERNEL(k_otam(st, h->SUPRO2, (long long)W * TD, TD, D, h->Z1 + (long long)(W + S + 1) * D, (long long)T * L1D, D,
                    L1D, E, W, Q, T, D, h->cfg.single_direct, 1.f, 1.f, h->ACC));
  SPM_KERNEL(k_otam(st, h->Z1, (long long)T * L1D, D, L1D, h->Z1 + (long long)(W + S) * D, (long long)T * L1D, 0, L1D,
                    E, W, 1, T, D, h->cfg.single_direct, 1.f, 0.f, h->D3));
  SPM_KERNEL(k_finalize(st, h->ACC, h->D3, E, Q, W, target_labels, tasks_per_batch, dists, logits, loss, acc, pred,
                        h->err_flag));
  h->last_E = E; h->last_S = S; h->last_Q = Q; h->last_W = W;
  return 0;
}

// Stage tensors of the most recent CLIP-SPM head pass, gathered out of the workspace in the reference's layouts
// (the tensors SURVEY.md 8(c) lists; tests compare each with the golden written from the executed reference).
int head_stage(spm_handle* h, cudaStream_t st, const char* name, float* out, long long capacity, long long* numel) {
  SPM_CHECK(h->cfg.head == SPM_HEAD_CLIPSPM || h->cfg.head == SPM_HEAD_CPM2C,
            "spm_head_stage: only the CLIP-SPM and CPM2C heads keep named stage tensors");
  SPM_CHECK(h->last_E > 0, "spm_head_stage: no head pass has run on this handle yet");
  const int E = h->last_E, S = h->last_S, Q = h->last_Q, W = h->last_W;
  if (h->cfg.head == SPM_HEAD_CPM2C) {
    // model_cpm2c.py:165-199 (motion features) and :315-360 on the normal branch, whose buffers are the ones left over
    const int T = h->cfg.seq_len, D = h->D, N = S + Q, V = E * N, L = T + 1;
    const long long MD = (long long)(T - 1) * D, LD = (long long)L * D;
    struct View { const float* p; int n0, n1, n2; long long s0, s1, s2; };
    const std::string n(name ? name : "");
    View v{};
    if (n == "su_motion") v = {h->CP_MOT, E, S, T - 1, N * MD, MD, D};
    else if (n == "qu_motion") v = {h->CP_MOT + S * MD, E, Q, T - 1, N * MD, MD, D};
    else if (n == "su_real") v = {h->Z, E, S, L, N * LD, LD, D};
    else if (n == "qu_fake") v = {h->Z + ((long long)V + S) * LD, E, Q, L, N * LD, LD, D};
    else if (n == "su_pro") v = {h->CP_PRO, E, W, L, W * LD, LD, D};
    else { set_error("spm_head_stage: unknown CPM2C stage '" + n + "'"); return 1; }
    const long long total = (long long)v.n0 * v.n1 * v.n2 * D;
    if (numel != nullptr) *numel = total;
    if (out == nullptr) return 0;
    SPM_CHECK(capacity >= total, "spm_head_stage: output buffer too small");
    SPM_KERNEL(k_gather4(st, v.p, v.n0, v.n1, v.n2, D, v.s0, v.s1, v.s2, out));
    return 0;
  }
  const int T = h->cfg.seq_len, D = h->D, N = S + Q, V = E * N, L1 = W + S + 1 + Q;
  const long long TD = (long long)T * D, T1D = (long long)(T + 1) * D, L1D = (long long)L1 * D;
  const float* Zb = h->Z + (long long)V * T1D;            // outputs of the `sem` se_te batch
  const float* Zm = h->Z;                                 // outputs of the `mo` se_te batch
  const float* tokb = h->TOK + (long long)V * D;          // tokens of the `sem` batch
  struct View { const float* p; int n0, n1, n2; long long s0, s1, s2; };
  const std::string n(name ? name : "");
  View v{};
  if (n == "su_mo") v = {h->TOK, E, S, 1, (long long)N * D, D, 0};                       // model_clipspm.py:195 (b1)
  else if (n == "qu_mo") v = {h->TOK + (long long)S * D, E, Q, 1, (long long)N * D, D, 0};
  else if (n == "support_token") v = {tokb, E, S, 1, (long long)N * D, D, 0};           // :120 (a4)
  else if (n == "target_token") v = {tokb + (long long)S * D, E, Q, 1, (long long)N * D, D, 0};   // :216 token_tr (c3)
  else if (n == "su_mo_refined") v = {Zm + D, E, S, T, N * T1D, T1D, D};                // :197 x' of the mo se_te (c1)
  else if (n == "qu_mo_refined") v = {Zm + S * T1D + D, E, Q, T, N * T1D, T1D, D};
  else if (n == "su_mo_token") v = {Zm, E, S, 1, N * T1D, T1D, 0};                      // m' (c1)
  else if (n == "qu_mo_token") v = {Zm + S * T1D, E, Q, 1, N * T1D, T1D, 0};
  else if (n == "su_mo2") v = {h->NEWM, E, S, 1, (long long)N * D, D, 0};               // :200 m'' (b1 on refined frames)
  else if (n == "qu_mo2") v = {h->NEWM + (long long)S * D, E, Q, 1, (long long)N * D, D, 0};
  else if (n == "su_real") v = {Zb + D, E, S, T, N * T1D, T1D, D};                      // :218,221 (c1/c4)
  else if (n == "qu_fake") v = {Zb + S * T1D + D, E, Q, T, N * T1D, T1D, D};
  else if (n == "token_s_real") v = {Zb, E, S, 1, N * T1D, T1D, 0};
  else if (n == "token_q_fake") v = {Zb + S * T1D, E, Q, 1, N * T1D, T1D, 0};
  else if (n == "su_pro") v = {h->SUPRO, E, W, T, W * TD, TD, D};                       // :231-239 (c6)
  else if (n == "su_t2") v = {h->Z1, E, W, T, T * L1D, D, L1D};                         // :275-294 (c5)
  else if (n == "su_2") v = {h->Z1 + (long long)W * D, E, S, T, T * L1D, D, L1D};
  else if (n == "qu_t2") v = {h->Z1 + (long long)(W + S) * D, E, 1, T, T * L1D, D, L1D};
  else if (n == "qu_2") v = {h->Z1 + (long long)(W + S + 1) * D, E, Q, T, T * L1D, D, L1D};
  else if (n == "su_pro2") v = {h->SUPRO2, E, W, T, W * TD, TD, D};                     // :133-137 on the PADM output
  else { set_error("spm_head_stage: unknown stage '" + n + "'"); return 1; }
  const long long total = (long long)v.n0 * v.n1 * v.n2 * D;
  if (numel != nullptr) *numel = total;
  if (out == nullptr) return 0;   // size query
  SPM_CHECK(capacity >= total, "spm_head_stage: output buffer too small");
  SPM_KERNEL(k_gather4(st, v.p, v.n0, v.n1, v.n2, D, v.s0, v.s1, v.s2, out));
  return 0;
}


// the label/W mismatch flag is per call: cleared (stream-ordered) when a public entry point starts
int reset_err_flag(spm_handle* h, cudaStream_t st) {
  if (h->err_flag != nullptr) SPM_CUDA(cudaMemsetAsync(h->err_flag, 0, sizeof(int), st));
  return 0;
}

int check_shapes(spm_handle* h, int E, int S, int Q, int W) {
  SPM_CHECK(h != nullptr, "null handle");
  SPM_CHECK(E >= 1 && S >= 1 && Q >= 1 && W >= 1, "episode shape must be positive");
  SPM_CHECK(W <= S, "way cannot exceed the number of support videos");
  SPM_CHECK(W + S <= 64 && Q + 1 <= 64, "PADM sequences longer than 64 tokens are not supported");
  SPM_CHECK(W <= 32 && Q <= 64, "at most 32 classes / 64 queries per episode");
  SPM_CHECK(h->cfg.seq_len >= 2 && h->cfg.seq_len <= 30, "seq_len must be in [2, 30]");
  return 0;
}

}  // namespace detail
}  // namespace spm
