// Frame encoder side of the handle: ViT-B/16 weight packing (models/clip_fsar.py:672-689 parameter names), chunk
// workspace, cached GEMM plans, the per-chunk kernel sequence, and encode_segments (fp32 images or decoded uint8 frames
// -> features), shared with the RN50 tower (rn50.cu).
#include "model_internal.cuh"

namespace spm {
namespace detail {

int load_vit(spm_handle* h, cudaStream_t st, const WeightTable& wt) {
  VitW& v = h->vit;
  const std::string p = "backbone.";
  const int C = VIT_C;
  SPM_TRY(copy_bf16(h, st, wt, p + "conv1.weight", (long long)C * C, &v.conv1_w));
  SPM_TRY(copy_f32(h, st, wt, p + "positional_embedding", (long long)VIT_L * C, &v.pos));
  const float* ce;
  SPM_TRY(wt.get(p + "class_embedding", C, &ce));
  SPM_TRY(dalloc_t(h, &v.cls_pos, C));
  SPM_KERNEL(k_add_vec(st, ce, v.pos, v.cls_pos, C));  // class token row = class_embedding + pos[0]
  SPM_TRY(copy_f32(h, st, wt, p + "ln_pre.weight", C, &v.ln_pre_g));
  SPM_TRY(copy_f32(h, st, wt, p + "ln_pre.bias", C, &v.ln_pre_b));
  SPM_TRY(copy_f32(h, st, wt, p + "ln_post.weight", C, &v.ln_post_g));
  SPM_TRY(copy_f32(h, st, wt, p + "ln_post.bias", C, &v.ln_post_b));
  const float* proj;
  SPM_TRY(wt.get(p + "proj", (long long)C * VIT_OUT, &proj));
  SPM_TRY(dalloc_t(h, &v.projT, (long long)C * VIT_OUT));
  SPM_KERNEL(k_transpose_cast_bf16(st, proj, v.projT, C, VIT_OUT));
  for (int i = 0; i < VIT_LAYERS; ++i) {
    const std::string b = p + "transformer.resblocks." + std::to_string(i) + ".";
    VitLayerW& l = v.layer[i];
    SPM_TRY(copy_bf16(h, st, wt, b + "attn.in_proj_weight", 3LL * C * C, &l.qkv_w));
    SPM_TRY(copy_f32(h, st, wt, b + "attn.in_proj_bias", 3 * C, &l.qkv_b));
    SPM_TRY(copy_bf16(h, st, wt, b + "attn.out_proj.weight", (long long)C * C, &l.out_w));
    SPM_TRY(copy_f32(h, st, wt, b + "attn.out_proj.bias", C, &l.out_b));
    SPM_TRY(copy_bf16(h, st, wt, b + "mlp.c_fc.weight", 4LL * C * C, &l.fc_w));
    SPM_TRY(copy_f32(h, st, wt, b + "mlp.c_fc.bias", 4 * C, &l.fc_b));
    SPM_TRY(copy_bf16(h, st, wt, b + "mlp.c_proj.weight", 4LL * C * C, &l.proj_w));
    SPM_TRY(copy_f32(h, st, wt, b + "mlp.c_proj.bias", C, &l.proj_b));
    SPM_TRY(copy_f32(h, st, wt, b + "ln_1.weight", C, &l.ln1_g));
    SPM_TRY(copy_f32(h, st, wt, b + "ln_1.bias", C, &l.ln1_b));
    SPM_TRY(copy_f32(h, st, wt, b + "ln_2.weight", C, &l.ln2_g));
    SPM_TRY(copy_f32(h, st, wt, b + "ln_2.bias", C, &l.ln2_b));
    if (h->ln_fold) {
      const float* w;
      SPM_TRY(wt.get(b + "attn.in_proj_weight", 3LL * C * C, &w));
      SPM_TRY(dalloc_t(h, &l.qkv_wf, 3LL * C * C));
      SPM_TRY(dalloc_t(h, &l.qkv_c, 3 * C));
      SPM_TRY(dalloc_t(h, &l.qkv_bf, 3 * C));
      SPM_KERNEL(k_fold_ln(st, w, l.ln1_g, l.ln1_b, l.qkv_b, 3 * C, C, l.qkv_wf, l.qkv_c, l.qkv_bf, h->ln_fold == 2));
      SPM_TRY(wt.get(b + "mlp.c_fc.weight", 4LL * C * C, &w));
      SPM_TRY(dalloc_t(h, &l.fc_wf, 4LL * C * C));
      SPM_TRY(dalloc_t(h, &l.fc_c, 4 * C));
      SPM_TRY(dalloc_t(h, &l.fc_bf, 4 * C));
      SPM_KERNEL(k_fold_ln(st, w, l.ln2_g, l.ln2_b, l.fc_b, 4 * C, C, l.fc_wf, l.fc_c, l.fc_bf, h->ln_fold == 2));
    }
  }
  return 0;
}

// [R, C] fp32 -> [C, R] fp32 (proj for the fp32 mode), tiny: done with a strided 2-D copy per column block
__global__ void transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)R * C) return;
  const int r = (int)(i / C), c = (int)(i % C);
  out[(long long)c * R + r] = in[i];
}

int load_vit32(spm_handle* h, cudaStream_t st, const WeightTable& wt) {
  VitW32& v = h->vit32;
  const std::string p = "backbone.";
  const int C = VIT_C;
  SPM_TRY(copy_f32(h, st, wt, p + "conv1.weight", (long long)C * C, &v.conv1_w));
  const float* proj;
  SPM_TRY(wt.get(p + "proj", (long long)C * VIT_OUT, &proj));
  SPM_TRY(dalloc_t(h, &v.projT, (long long)C * VIT_OUT));
  transpose_f32_kernel<<<(C * VIT_OUT + 255) / 256, 256, 0, st>>>(proj, v.projT, C, VIT_OUT);
  SPM_CUDA(cudaGetLastError());
  for (int i = 0; i < VIT_LAYERS; ++i) {
    const std::string b = p + "transformer.resblocks." + std::to_string(i) + ".";
    VitLayerW32& l = v.layer[i];
    SPM_TRY(copy_f32(h, st, wt, b + "attn.in_proj_weight", 3LL * C * C, &l.qkv_w));
    SPM_TRY(copy_f32(h, st, wt, b + "attn.out_proj.weight", (long long)C * C, &l.out_w));
    SPM_TRY(copy_f32(h, st, wt, b + "mlp.c_fc.weight", 4LL * C * C, &l.fc_w));
    SPM_TRY(copy_f32(h, st, wt, b + "mlp.c_proj.weight", 4LL * C * C, &l.proj_w));
  }
  return 0;
}

namespace {
void select_vit_ws(spm_handle* h, int i) {
  const spm_handle::VitWs& w = h->vit_ws[i];
  h->patches = w.patches; h->xn = w.xn; h->qkv = w.qkv; h->attn = w.attn; h->hid = w.hid; h->cls = w.cls;
  h->xnc = w.xnc; h->x = w.x; h->xc = w.xc; h->xb = w.xb; h->xcb = w.xcb; h->ln_stats = w.ln_stats;
  h->cur_ws = i;
}

int ensure_vit_workspace(spm_handle* h) {
  if (h->x != nullptr) return 0;
  const long long M = (long long)h->frame_chunk * VIT_L;
  if (h->fp32) h->enc_streams = 1;
  for (int i = 0; i < h->enc_streams; ++i) {
    spm_handle::VitWs& w = h->vit_ws[i];
    SPM_TRY(dalloc_t(h, &w.patches, (long long)h->frame_chunk * VIT_P * VIT_C));
    SPM_TRY(dalloc_t(h, &w.x, M * VIT_C));
    SPM_TRY(dalloc_t(h, &w.xn, M * VIT_C));
    SPM_TRY(dalloc_t(h, &w.qkv, M * 3 * VIT_C));
    SPM_TRY(dalloc_t(h, &w.attn, M * VIT_C));
    SPM_TRY(dalloc_t(h, &w.hid, M * 4 * VIT_C));
    SPM_TRY(dalloc_t(h, &w.cls, (long long)h->frame_chunk * VIT_C));
    SPM_TRY(dalloc_t(h, &w.xc, (long long)h->frame_chunk * VIT_C));
    SPM_TRY(dalloc_t(h, &w.xnc, (long long)h->frame_chunk * VIT_C));
    w.xb = w.xcb = nullptr;
    w.ln_stats = nullptr;
    if (h->ln_fold) {
      SPM_TRY(dalloc_t(h, &w.xb, M * VIT_C));
      SPM_TRY(dalloc_t(h, &w.ln_stats, M * 2 * LN_FOLD_SLOTS));
    }
    if (h->resid_bf16) {
      SPM_TRY(dalloc_t(h, &w.xb, M * VIT_C));
      SPM_TRY(dalloc_t(h, &w.xcb, (long long)h->frame_chunk * VIT_C));
    }
  }
  if (h->enc_streams > 1) {
    for (int i = 0; i < 2; ++i) {
      SPM_CUDA(cudaStreamCreateWithFlags(&h->enc_stream[i], cudaStreamNonBlocking));
      SPM_CUDA(cudaEventCreateWithFlags(&h->enc_join[i], cudaEventDisableTiming));
    }
    SPM_CUDA(cudaEventCreateWithFlags(&h->enc_fork, cudaEventDisableTiming));
  }
  select_vit_ws(h, 0);
  if (h->fp32) {
    SPM_TRY(dalloc_t(h, &h->patches32, (long long)h->frame_chunk * VIT_P * VIT_C));
    SPM_TRY(dalloc_t(h, &h->xn32, M * VIT_C));
    SPM_TRY(dalloc_t(h, &h->qkv32, M * 3 * VIT_C));
    SPM_TRY(dalloc_t(h, &h->attn32, M * VIT_C));
    SPM_TRY(dalloc_t(h, &h->hid32, M * 4 * VIT_C));
    SPM_TRY(dalloc_t(h, &h->cls32, (long long)h->frame_chunk * VIT_C));
  }
  return 0;
}

int get_vit_plan(spm_handle* h, int F, VitPlan** out) {
  const int key = F * 2 + h->cur_ws;  // plans bake the workspace pointers into their tensor maps
  auto it = h->vit_plans.find(key);
  if (it != h->vit_plans.end()) { *out = it->second.get(); return 0; }
  std::unique_ptr<VitPlan> pl(new VitPlan());
  const int C = VIT_C, M = F * VIT_L;
  const VitW& v = h->vit;
  // operands of the two precisions: bf16 tensor-core path, or fp32 activations/weights on the exact SIMT kernel
  const bool f32 = h->fp32;
  const int kind = f32 ? GEMM_F32_SIMT : GEMM_BF16;
  const int obf = f32 ? 0 : 1;
  const void* a_patches = f32 ? (const void*)h->patches32 : (const void*)h->patches;
  const void* a_xn = f32 ? (const void*)h->xn32 : (const void*)h->xn;
  const void* a_attn = f32 ? (const void*)h->attn32 : (const void*)h->attn;
  const void* a_hid = f32 ? (const void*)h->hid32 : (const void*)h->hid;
  const void* a_cls = f32 ? (const void*)h->cls32 : (const void*)h->cls;
  void* o_qkv = f32 ? (void*)h->qkv32 : (void*)h->qkv;
  void* o_hid = f32 ? (void*)h->hid32 : (void*)h->hid;
  {
    GemmEpilogue ep;  // x[f*197 + 1 + p] = patch . W + pos[1 + p]
    ep.residual = v.pos; ep.ldr = C; ep.res_row_mod = VIT_P; ep.res_row_off = 1;
    ep.out_row_group = VIT_P; ep.out_group_stride = VIT_L; ep.out_row_off = 1;
    ep.out = h->x; ep.ldo = C;
    SPM_TRY(plan_gemm(&pl->patch, kind, a_patches, C, f32 ? (const void*)h->vit32.conv1_w : (const void*)v.conv1_w, C,
                      F * VIT_P, C, C, ep, h->sms));
  }
  for (int i = 0; i < VIT_LAYERS; ++i) {
    const VitLayerW& l = v.layer[i];
    const VitLayerW32& l32 = h->vit32.layer[i];
    GemmEpilogue e1;
    e1.bias = l.qkv_b; e1.out = o_qkv; e1.ldo = 3 * C; e1.out_bf16 = obf;
    SPM_TRY(plan_gemm(&pl->qkv[i], kind, a_xn, C, f32 ? (const void*)l32.qkv_w : (const void*)l.qkv_w, C, M, 3 * C, C, e1, h->sms));
    GemmEpilogue e2;
    e2.bias = l.out_b; e2.residual = h->x; e2.ldr = C; e2.out = h->x; e2.ldo = C;
    if (h->resid_bf16) {   // x (bf16) += out_proj(attn): bf16 residual in, bf16 out, in place
      e2.residual = nullptr; e2.residual_bf16 = h->xb; e2.out = h->xb; e2.out_bf16 = 1;
    }
    SPM_TRY(plan_gemm(&pl->outp[i], kind, a_attn, C, f32 ? (const void*)l32.out_w : (const void*)l.out_w, C, M, C, C, e2, h->sms));
    GemmEpilogue e3;
    e3.bias = l.fc_b; e3.act = ACT_QUICKGELU; e3.out = o_hid; e3.ldo = 4 * C; e3.out_bf16 = obf;
    SPM_TRY(plan_gemm(&pl->fc[i], kind, a_xn, C, f32 ? (const void*)l32.fc_w : (const void*)l.fc_w, C, M, 4 * C, C, e3, h->sms));
    GemmEpilogue e4;
    e4.bias = l.proj_b; e4.residual = h->x; e4.ldr = C; e4.out = h->x; e4.ldo = C;
    if (h->resid_bf16) {
      e4.residual = nullptr; e4.residual_bf16 = h->xb; e4.out = h->xb; e4.out_bf16 = 1;
    }
    SPM_TRY(plan_gemm(&pl->proj[i], kind, a_hid, 4 * C, f32 ? (const void*)l32.proj_w : (const void*)l.proj_w, 4 * C, M, C, 4 * C, e4, h->sms));
  }
  // LayerNorm folding: every block GEMM runs on the pair kernel with the epilogues that implement it (TMA residual epilogue
  // for out_proj / c_proj, bf16 + bias epilogue for QKV / c_fc), whatever the chunk size
  if (h->ln_fold && !f32 && !h->resid_bf16) {
    for (int i = 0; i < VIT_LAYERS; ++i) {
      const VitLayerW& l = v.layer[i];
      GemmEpilogue e1 = pl->qkv[i].ep;
      e1.bias = l.qkv_bf; e1.ln_stats_in = h->ln_stats; e1.ln_colsum = h->ln_fold == 2 ? nullptr : l.qkv_c;
      SPM_TRY(plan_gemm(&pl->qkv[i], kind, h->xb, C, l.qkv_wf, C, M, 3 * C, C, e1, h->sms));
      GemmEpilogue e3 = pl->fc[i].ep;
      e3.bias = l.fc_bf; e3.ln_stats_in = h->ln_stats; e3.ln_colsum = h->ln_fold == 2 ? nullptr : l.fc_c;
      SPM_TRY(plan_gemm(&pl->fc[i], kind, h->xb, C, l.fc_wf, C, M, 4 * C, C, e3, h->sms));
      GemmEpilogue e2 = pl->outp[i].ep;
      e2.out2_bf16 = h->xb; e2.ln_stats_out = h->ln_stats;
      SPM_TRY(plan_gemm(&pl->outp[i], kind, a_attn, C, l.out_w, C, M, C, C, e2, h->sms));
      GemmEpilogue e4 = pl->proj[i].ep;
      e4.out2_bf16 = h->xb; e4.ln_stats_out = h->ln_stats;
      SPM_TRY(plan_gemm(&pl->proj[i], kind, a_hid, 4 * C, l.proj_w, 4 * C, M, C, 4 * C, e4, h->sms));
      SPM_CHECK(pl->qkv[i].two_cta && pl->fc[i].two_cta && pl->outp[i].res_tma && pl->proj[i].res_tma,
                "vit plan: LayerNorm folding lost its kernels");
    }
    pl->ln_fold = true;
  }
  if (!f32) {
    // Last block, class-token rows only: attention output / residual rows are taken with a row stride of 197 tokens
    const VitLayerW& l = v.layer[VIT_LAYERS - 1];
    const long long LC = (long long)VIT_L * C;
    GemmEpilogue e2;
    e2.bias = l.out_b; e2.residual = h->x; e2.ldr = (int)LC; e2.out = h->xc; e2.ldo = C;
    if (h->resid_bf16) {
      e2.residual = nullptr; e2.residual_bf16 = h->xb; e2.out = h->xcb; e2.out_bf16 = 1;
    }
    SPM_TRY(plan_gemm(&pl->outp_cls, GEMM_BF16, h->attn, LC, l.out_w, C, F, C, C, e2, h->sms));
    GemmEpilogue e3;
    e3.bias = l.fc_b; e3.act = ACT_QUICKGELU; e3.out = h->hid; e3.ldo = 4 * C; e3.out_bf16 = 1;
    SPM_TRY(plan_gemm(&pl->fc_cls, GEMM_BF16, h->xnc, C, l.fc_w, C, F, 4 * C, C, e3, h->sms));
    GemmEpilogue e4;
    e4.bias = l.proj_b; e4.residual = h->xc; e4.ldr = C; e4.out = h->xc; e4.ldo = C;
    if (h->resid_bf16) {
      e4.residual = nullptr; e4.residual_bf16 = h->xcb; e4.out = h->xcb; e4.out_bf16 = 1;
    }
    SPM_TRY(plan_gemm(&pl->proj_cls, GEMM_BF16, h->hid, 4 * C, l.proj_w, 4 * C, F, C, 4 * C, e4, h->sms));
  }
  {
    GemmEpilogue ep;
    ep.out = h->x;  // patched per call
    ep.ldo = VIT_OUT;
    SPM_TRY(plan_gemm(&pl->fin, kind, a_cls, C, f32 ? (const void*)h->vit32.projT : (const void*)v.projT, C, F, VIT_OUT, C, ep, h->sms));
  }
  *out = pl.get();
  h->vit_plans[key] = std::move(pl);
  return 0;
}

// `F` frames already im2col'ed into h->patches -> feats_out [F, 512]
int vit_run(spm_handle* h, cudaStream_t st, int F, float* feats_out) {
  VitPlan* pl;
  SPM_TRY(get_vit_plan(h, F, &pl));
  const VitW& v = h->vit;
  const int C = VIT_C, M = F * VIT_L;
  if (h->fp32) {
    // parity mode: same graph, fp32 activations, exact FFMA GEMMs / attention, no pruning shortcuts
    SPM_GEMM_RUN(pl->patch);
    SPM_KERNEL(k_layernorm(st, h->x, C, M, C, v.ln_pre_g, v.ln_pre_b, v.cls_pos, VIT_L, h->x, nullptr, C));
    for (int i = 0; i < VIT_LAYERS; ++i) {
      const VitLayerW& l = v.layer[i];
      SPM_KERNEL(k_layernorm(st, h->x, C, M, C, l.ln1_g, l.ln1_b, nullptr, 0, h->xn32, nullptr, C));
      SPM_GEMM_RUN(pl->qkv[i]);
      SPM_KERNEL(k_vit_attention_f32(st, h->qkv32, h->attn32, F));
      SPM_GEMM_RUN(pl->outp[i]);
      SPM_KERNEL(k_layernorm(st, h->x, C, M, C, l.ln2_g, l.ln2_b, nullptr, 0, h->xn32, nullptr, C));
      SPM_GEMM_RUN(pl->fc[i]);
      SPM_GEMM_RUN(pl->proj[i]);
    }
    SPM_KERNEL(k_layernorm(st, h->x, (long long)VIT_L * C, F, C, v.ln_post_g, v.ln_post_b, nullptr, 0, h->cls32, nullptr, C));
    GemmOp fin32 = pl->fin;
    fin32.ep.out = feats_out;
    SPM_GEMM_RUN(fin32);
    return 0;
  }
  // Consecutive kernels sweep their rows in OPPOSITE directions (h->alt_dir): a 512-frame chunk's tensors (155-620 MB)
  // do not fit the 126 MB L2, but the rows a kernel wrote last are still there when the next kernel starts on them.
  int dir = 0;
  auto next_dir = [&]() { const int d = dir; dir ^= h->alt_dir; return d; };
#define SPM_GEMM_RUN_DIR(op)            \
  do {                                  \
    GemmOp _op = (op);                  \
    _op.reverse = next_dir();           \
    SPM_GEMM_RUN(_op);                  \
  } while (0)
  SPM_GEMM_RUN_DIR(pl->patch);
  const bool rb = h->resid_bf16;   // bf16 residual stream: ln_pre writes it, every later LayerNorm reads it
  const bool lf = pl->ln_fold;     // ln_1 / ln_2 live in the GEMM epilogues: ln_pre also emits the bf16 rows + row statistics
  if (lf)
    SPM_KERNEL(k_layernorm(st, h->x, C, M, C, v.ln_pre_g, v.ln_pre_b, v.cls_pos, VIT_L, h->x, h->xb, C, next_dir(), h->ln_stats));
  else
    SPM_KERNEL(k_layernorm(st, h->x, C, M, C, v.ln_pre_g, v.ln_pre_b, v.cls_pos, VIT_L, rb ? nullptr : h->x, rb ? h->xb : nullptr,
                           C, next_dir()));
  for (int i = 0; i < VIT_LAYERS; ++i) {
    const VitLayerW& l = v.layer[i];
    if (lf) {}
    else if (rb) SPM_KERNEL(k_layernorm_bf16in(st, h->xb, C, M, C, l.ln1_g, l.ln1_b, h->xn, C, next_dir()));
    else SPM_KERNEL(k_layernorm(st, h->x, C, M, C, l.ln1_g, l.ln1_b, nullptr, 0, nullptr, h->xn, C, next_dir()));
    SPM_GEMM_RUN_DIR(pl->qkv[i]);
    if (h->attn_mma)
      SPM_KERNEL(k_vit_attention(st, h->qkv, h->attn, F));
    else
      SPM_KERNEL(k_vit_attention_tc(st, h->qkv, h->attn, F, h->sms, next_dir()));
    if (i == VIT_LAYERS - 1 && h->prune_last) {
      // only x[:, 0, :] is read after the last block: run its out-proj / MLP on the F class-token rows
      SPM_GEMM_RUN(pl->outp_cls);
      if (rb) SPM_KERNEL(k_layernorm_bf16in(st, h->xcb, C, F, C, l.ln2_g, l.ln2_b, h->xnc, C));
      else SPM_KERNEL(k_layernorm(st, h->xc, C, F, C, l.ln2_g, l.ln2_b, nullptr, 0, nullptr, h->xnc, C));
      SPM_GEMM_RUN(pl->fc_cls);
      SPM_GEMM_RUN(pl->proj_cls);
      if (rb) SPM_KERNEL(k_layernorm_bf16in(st, h->xcb, C, F, C, v.ln_post_g, v.ln_post_b, h->cls, C));
      else SPM_KERNEL(k_layernorm(st, h->xc, C, F, C, v.ln_post_g, v.ln_post_b, nullptr, 0, nullptr, h->cls, C));
      GemmOp fin = pl->fin;
      fin.ep.out = feats_out;
      SPM_GEMM_RUN(fin);
      return 0;
    }
    SPM_GEMM_RUN_DIR(pl->outp[i]);
    if (lf) {}
    else if (rb) SPM_KERNEL(k_layernorm_bf16in(st, h->xb, C, M, C, l.ln2_g, l.ln2_b, h->xn, C, next_dir()));
    else SPM_KERNEL(k_layernorm(st, h->x, C, M, C, l.ln2_g, l.ln2_b, nullptr, 0, nullptr, h->xn, C, next_dir()));
    SPM_GEMM_RUN_DIR(pl->fc[i]);
    SPM_GEMM_RUN_DIR(pl->proj[i]);
  }
#undef SPM_GEMM_RUN_DIR
  if (rb) SPM_KERNEL(k_layernorm_bf16in(st, h->xb, (long long)VIT_L * C, F, C, v.ln_post_g, v.ln_post_b, h->cls, C));
  else SPM_KERNEL(k_layernorm(st, h->x, (long long)VIT_L * C, F, C, v.ln_post_g, v.ln_post_b, nullptr, 0, nullptr, h->cls, C));
  GemmOp fin = pl->fin;
  fin.ep.out = feats_out;
  SPM_GEMM_RUN(fin);
  return 0;
}


// fp32 images of frames [a, b) of a segment: the caller's own, or transformed into the handle's scratch
int segment_images(spm_handle* h, cudaStream_t st, const Segment& seg, long long a, long long b, const float** out) {
  if (seg.frames_u8 == nullptr) { *out = seg.images + a * FRAME_ELEMS; return 0; }
  if (b - a > h->img_scratch_cap) {
    SPM_TRY(drealloc_t(h, &h->img_scratch, (b - a) * FRAME_ELEMS));
    h->img_scratch_cap = b - a;
  }
  SPM_KERNEL(k_frame_transform(st, seg.frames_u8 + a * (long long)seg.H * seg.W * 3, (int)(b - a), seg.H, seg.W,
                               h->img_scratch, nullptr));
  *out = h->img_scratch;
  return 0;
}

}  // namespace

int encode_segments(spm_handle* h, cudaStream_t st, const Segment* segs, int nseg, float* feats_out,
                    const ChunkHook* after_chunk) {
  SPM_CHECK(h->weights_loaded, "encode: weights not loaded (spm_load_weights)");
  if (h->cfg.backbone == SPM_BACKBONE_RN50) {
    long long done = 0;
    for (int s = 0; s < nseg; ++s) {
      const long long step = segs[s].frames_u8 ? 256 : segs[s].n_frames;  // uint8 input: bounded fp32 scratch
      for (long long a = 0; a < segs[s].n_frames; a += step) {
        const long long b = std::min(segs[s].n_frames, a + step);
        const float* img;
        SPM_TRY(segment_images(h, st, segs[s], a, b, &img));
        SPM_TRY(rn50_encode(h->rn50, st, img, (int)(b - a), feats_out + (done + a) * h->D));
      }
      done += segs[s].n_frames;
    }
    return 0;
  }
  SPM_TRY(ensure_vit_workspace(h));
  long long total = 0;
  for (int s = 0; s < nseg; ++s) total += segs[s].n_frames;
  // more than one chunk: alternate chunks between the two encoder streams (forked from / joined back into `st`)
  const bool dual = h->enc_streams > 1 && total > h->frame_chunk && !profile_armed();
  cudaStream_t caller = st;
  if (dual) {
    SPM_CUDA(cudaEventRecord(h->enc_fork, caller));
    for (int i = 0; i < 2; ++i) SPM_CUDA(cudaStreamWaitEvent(h->enc_stream[i], h->enc_fork, 0));
  }
  int chunk_no = 0;
  for (long long f0 = 0; f0 < total; f0 += h->frame_chunk, ++chunk_no) {
    const long long f1 = std::min(total, f0 + h->frame_chunk);
    if (dual) {
      select_vit_ws(h, chunk_no & 1);
      st = h->enc_stream[chunk_no & 1];
    } else if (h->cur_ws != 0) {
      select_vit_ws(h, 0);
    }
    long long seg0 = 0;
    for (int s = 0; s < nseg; ++s) {
      const long long a = std::max(f0, seg0), b = std::min(f1, seg0 + segs[s].n_frames);
      if (a < b) {
        if (h->fp32) {
          const float* img;
          SPM_TRY(segment_images(h, st, segs[s], a - seg0, b - seg0, &img));
          SPM_KERNEL(k_patch_im2col_f32(st, img, h->patches32 + (a - f0) * VIT_P * VIT_C, (int)(b - a)));
        } else if (segs[s].frames_u8 != nullptr) {  // uint8 frames -> bf16 patch matrix in one kernel
          SPM_KERNEL(k_frame_transform(st, segs[s].frames_u8 + (a - seg0) * (long long)segs[s].H * segs[s].W * 3,
                                       (int)(b - a), segs[s].H, segs[s].W, nullptr,
                                       h->patches + (a - f0) * VIT_P * VIT_C));
        } else {
          SPM_KERNEL(k_patch_im2col(st, segs[s].images + (a - seg0) * FRAME_ELEMS,
                                    h->patches + (a - f0) * VIT_P * VIT_C, (int)(b - a)));
        }
      }
      seg0 += segs[s].n_frames;
    }
    SPM_TRY(vit_run(h, st, (int)(f1 - f0), feats_out + f0 * h->D));
    if (after_chunk != nullptr) SPM_TRY((*after_chunk)(f1, chunk_no, st));
  }
  if (dual) {
    for (int i = 0; i < 2; ++i) {
      SPM_CUDA(cudaEventRecord(h->enc_join[i], h->enc_stream[i]));
      SPM_CUDA(cudaStreamWaitEvent(caller, h->enc_join[i], 0));
    }
  }
  return 0;
}

}  // namespace detail
}  // namespace spm
