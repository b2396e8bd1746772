// Handle, weight packing, workspace and the stream-ordered pipelines behind the C ABI:
//   frame encoder (ViT-B/16)  models/clip_fsar.py:672-689
//   metric head               models/model_clipspm.py:116-143 (mo / sem / taskM / otam_distance / logits)
//   loss + accuracy           utils/utils.py:174-186,259-264 ; run/main_run.py:390-392
// No allocation happens inside a forward call once the shapes have been seen (plans and workspace are cached),
// and nothing synchronises with the host except spm_load_weights and spm_eval_host (whose contract is blocking).
#include "model_internal.cuh"

namespace spm {
namespace {

// su_img / qu_img: fp32 [.,3,224,224] images, or -- when img_h > 0 -- uint8 [., img_h, img_w, 3] decoded frames
int forward_impl(spm_handle* h, cudaStream_t st, int E, int S, int Q, int W, const void* su_img, const void* qu_img,
                 const float* labels, const float* real_s, const float* real_t, const long long* target_labels,
                 float tasks_per_batch, float* logits, float* dists, float* loss, float* acc, int* pred, int img_h = 0,
                 int img_w = 0) {
  SPM_TRY(check_shapes(h, E, S, Q, W));
  const int T = h->cfg.seq_len, D = h->D, N = S + Q;
  const long long nf = (long long)E * N * T;
  if (nf > h->xall_cap) {
    SPM_TRY(drealloc_t(h, &h->Xall, nf * D));
    h->xall_cap = nf;
    h->head_plans.clear();
  }
  // scratch outputs for callers that only want loss / accuracy: each buffer has its own capacity (a call with fewer
  // episodes but more logits per episode, or the reverse, must not reuse a buffer sized for the other)
  if (logits == nullptr && (long long)E * Q * W > h->tmp_logits_cap) {
    SPM_TRY(drealloc_t(h, &h->tmp_logits, (long long)E * Q * W));
    h->tmp_logits_cap = (long long)E * Q * W;
  }
  if (dists == nullptr && E > h->tmp_dists_cap) {
    SPM_TRY(drealloc_t(h, &h->tmp_dists, E));
    h->tmp_dists_cap = E;
  }
  if (logits == nullptr) logits = h->tmp_logits;
  if (dists == nullptr) dists = h->tmp_dists;
  // frames in episode-major order (supports, then queries of each episode): the encoder writes X [E, N, T, D] directly
  std::vector<Segment> segs(2 * (size_t)E);
  const long long fb = img_h > 0 ? (long long)img_h * img_w * 3 : (long long)FRAME_ELEMS * 4;  // bytes per frame
  for (int e = 0; e < E; ++e) {
    Segment& a = segs[2 * e];
    Segment& b = segs[2 * e + 1];
    a.n_frames = (long long)S * T;
    b.n_frames = (long long)Q * T;
    const uint8_t* pa = static_cast<const uint8_t*>(su_img) + (long long)e * S * T * fb;
    const uint8_t* pb = static_cast<const uint8_t*>(qu_img) + (long long)e * Q * T * fb;
    if (img_h > 0) {
      a.frames_u8 = pa; b.frames_u8 = pb;
      a.H = b.H = img_h; a.W = b.W = img_w;
    } else {
      a.images = reinterpret_cast<const float*>(pa);
      b.images = reinterpret_cast<const float*>(pb);
    }
  }
  // Pipelined form (bf16 ViT, several chunks, not while GEMM launches are being event-timed): episodes are split
  // into groups; as soon as the chunks holding a group's frames are enqueued its head is enqueued on head_stream,
  // where it overlaps the encoder chunks of the following groups.  Only the last group's head is exposed.
  const bool vit = h->cfg.backbone == SPM_BACKBONE_VIT_B16;
  const bool pipelined = vit && !h->fp32 && h->enc_streams > 1 && !profile_armed() && E >= 2 && nf > h->frame_chunk;
  if (!vit) {
    // RN50 consumes contiguous runs of frames in its own 64-frame chunks: encode [all supports | all queries] and
    // scatter the feature rows into the episode-major X
    if (nf > h->feats_cap) {
      SPM_TRY(drealloc_t(h, &h->feats, nf * D));
      h->feats_cap = nf;
    }
    Segment two[2] = {segs[0], segs[1]};
    two[0].n_frames = (long long)E * S * T;
    two[1].n_frames = (long long)E * Q * T;
    SPM_TRY(encode_segments(h, st, two, 2, h->feats));
    const size_t row = (size_t)T * D * 4;
    SPM_CUDA(cudaMemcpy2DAsync(h->Xall, (size_t)N * row, h->feats, (size_t)S * row, (size_t)S * row, E,
                               cudaMemcpyDeviceToDevice, st));
    SPM_CUDA(cudaMemcpy2DAsync(h->Xall + (long long)S * T * D, (size_t)N * row, h->feats + (long long)E * S * T * D,
                               (size_t)Q * row, (size_t)Q * row, E, cudaMemcpyDeviceToDevice, st));
  }
  if (!pipelined) {
    if (vit) SPM_TRY(encode_segments(h, st, segs.data(), (int)segs.size(), h->Xall));
    SPM_TRY(ensure_head_workspace(h, E, S, Q, W));
    h->X = h->Xall;
    const int rc = head_run(h, st, E, S, Q, W, labels, real_s, real_t, target_labels, tasks_per_batch, logits, dists,
                            loss, acc, pred);
    h->X = h->Xhead;
    return rc;
  }
  const int Eg = std::max(1, (E + 3) / 4), G = (E + Eg - 1) / Eg;
  SPM_TRY(ensure_head_workspace(h, Eg, S, Q, W));
  if (h->head_stream == nullptr) {
    SPM_CUDA(cudaStreamCreateWithFlags(&h->head_stream, cudaStreamNonBlocking));
    SPM_CUDA(cudaEventCreateWithFlags(&h->head_done, cudaEventDisableTiming));
  }
  const int n_chunks = (int)((nf + h->frame_chunk - 1) / h->frame_chunk);
  while ((int)h->chunk_ev.size() < n_chunks) {
    cudaEvent_t ev;
    SPM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    h->chunk_ev.push_back(ev);
  }
  int next_group = 0;
  ChunkHook hook = [&](long long frames_done, int chunk_no, cudaStream_t cs) -> int {
    SPM_CUDA(cudaEventRecord(h->chunk_ev[chunk_no], cs));
    while (next_group < G && (long long)std::min(E, (next_group + 1) * Eg) * N * T <= frames_done) {
      const int e0 = next_group * Eg, ne = std::min(Eg, E - e0);
      SPM_CUDA(cudaStreamWaitEvent(h->head_stream, h->chunk_ev[chunk_no], 0));
      if (chunk_no > 0) SPM_CUDA(cudaStreamWaitEvent(h->head_stream, h->chunk_ev[chunk_no - 1], 0));
      h->X = h->Xall + (long long)e0 * N * T * D;
      const int rc = head_run(h, h->head_stream, ne, S, Q, W, labels + (long long)e0 * S, real_s + (long long)e0 * S,
                              real_t + (long long)e0 * Q, target_labels ? target_labels + (long long)e0 * Q : nullptr,
                              tasks_per_batch, logits + (long long)e0 * Q * W, dists + e0, loss ? loss + e0 : nullptr,
                              acc ? acc + e0 : nullptr, pred ? pred + (long long)e0 * Q : nullptr);
      h->X = h->Xhead;
      if (rc) return rc;
      ++next_group;
    }
    return 0;
  };
  SPM_TRY(encode_segments(h, st, segs.data(), (int)segs.size(), h->Xall, &hook));
  SPM_CHECK(next_group == G, "forward: internal error (episode groups left without a head pass)");
  SPM_CUDA(cudaEventRecord(h->head_done, h->head_stream));
  SPM_CUDA(cudaStreamWaitEvent(st, h->head_done, 0));
  return 0;
}


}  // namespace
}  // namespace spm

// =============================================================================================================
// C ABI
// =============================================================================================================
using namespace spm;

extern "C" {

int spm_create(const spm_config* cfg, spm_handle** out) {
  SPM_CHECK(cfg != nullptr && out != nullptr, "spm_create: null argument");
  SPM_CHECK(cfg->backbone == SPM_BACKBONE_VIT_B16 || cfg->backbone == SPM_BACKBONE_RN50, "spm_create: unknown backbone");
  SPM_CHECK(cfg->seq_len >= 2 && cfg->seq_len <= 30, "spm_create: seq_len must be in [2, 30]");
  SPM_CHECK(cfg->precision == SPM_PRECISION_BF16 || cfg->precision == SPM_PRECISION_FP32 ||
            cfg->precision == SPM_PRECISION_BF16_RESID, "spm_create: unknown precision");
  SPM_CHECK(cfg->head == SPM_HEAD_CLIPSPM || cfg->head == SPM_HEAD_CLIPFSAR || cfg->head == SPM_HEAD_STEN ||
            cfg->head == SPM_HEAD_CPM2C, "spm_create: unknown head");
  SPM_CHECK(cfg->head != SPM_HEAD_STEN || cfg->seq_len == 8,
            "spm_create: the STEN head reshapes to 8 frames per video (models/model_sten.py:65-66)");
  SPM_CHECK(cfg->precision == SPM_PRECISION_BF16 || cfg->backbone == SPM_BACKBONE_VIT_B16,
            "spm_create: SPM_PRECISION_FP32 / SPM_PRECISION_BF16_RESID are implemented for the ViT-B/16 backbone only");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("spm_create: no CUDA device -- this library has no CPU path");
    return 1;
  }
  int dev = 0, major = 0;
  SPM_CUDA(cudaGetDevice(&dev));
  SPM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SPM_CHECK(major == 10, "spm_create: the kernels are built for sm_100a (B200) only");
  std::unique_ptr<spm_handle> h(new spm_handle());
  h->cfg = *cfg;
  h->fp32 = cfg->precision == SPM_PRECISION_FP32;
  h->resid_bf16 = cfg->precision == SPM_PRECISION_BF16_RESID;
  h->D = cfg->backbone == SPM_BACKBONE_VIT_B16 ? 512 : 1024;
  h->HT = (int)(h->D * cfg->mid_dim_text);
  h->HV = (int)(h->D * cfg->mid_dim_vision);
  SPM_CHECK(h->HT % 32 == 0 && h->HV % 32 == 0 && h->HT > 0 && h->HV > 0,
            "spm_create: gate hidden sizes must be positive multiples of 32");
  SPM_TRY(device_sm_count(&h->sms));
  if (const char* e = getenv("SPM_FRAME_CHUNK")) h->frame_chunk = std::max(1, atoi(e));
  if (const char* e = getenv("SPM_ENC_STREAMS")) h->enc_streams = atoi(e) >= 2 ? 2 : 1;
  if (const char* e = getenv("SPM_ALT_DIR")) h->alt_dir = atoi(e) != 0 ? 1 : 0;
  const char* err = "";
  if (gemm_init(&err)) { set_error(err); return 1; }
  SPM_KERNEL(k_vit_attention_init());
  SPM_KERNEL(k_vit_attention_tc_init());
  SPM_KERNEL(k_vit_attention_f32_init());
  if (const char* e = getenv("SPM_ATTN")) h->attn_mma = std::string(e) == "mma";
  if (const char* e = getenv("SPM_PRUNE_LAST")) h->prune_last = atoi(e) != 0;
  if (const char* e = getenv("SPM_LN_FOLD")) h->ln_fold = atoi(e);
  if (h->fp32 || h->resid_bf16 || cfg->backbone != SPM_BACKBONE_VIT_B16) h->ln_fold = 0;
  SPM_KERNEL(k_seq_attention_init());
  SPM_KERNEL(k_otam_init());
  *out = h.release();
  return 0;
}

int spm_destroy(spm_handle* h) {
  if (h == nullptr) return 0;
  cudaDeviceSynchronize();
  if (h->rn50) rn50_destroy(h->rn50);
  for (void* p : h->allocs) cudaFree(p);
  if (h->pin_res) cudaFreeHost(h->pin_res);
  for (cudaEvent_t e : h->ev_copied) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_done) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (h->enc_stream[i]) cudaStreamDestroy(h->enc_stream[i]);
    if (h->enc_join[i]) cudaEventDestroy(h->enc_join[i]);
  }
  if (h->enc_fork) cudaEventDestroy(h->enc_fork);
  if (h->head_stream) cudaStreamDestroy(h->head_stream);
  if (h->head_done) cudaEventDestroy(h->head_done);
  for (cudaEvent_t e : h->chunk_ev) cudaEventDestroy(e);
  if (h->pf_event) cudaEventDestroy(h->pf_event);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->compute_stream) cudaStreamDestroy(h->compute_stream);
  delete h;
  return 0;
}

int spm_load_weights(spm_handle* h, void* stream, int n, const char* const* names, const void* const* dev_ptrs,
                     const int64_t* numel) {
  SPM_CHECK(h != nullptr && names != nullptr && dev_ptrs != nullptr && numel != nullptr, "spm_load_weights: null argument");
  SPM_CHECK(!h->weights_loaded, "spm_load_weights: weights already loaded for this handle");
  cudaStream_t st = (cudaStream_t)stream;
  WeightTable wt;
  for (int i = 0; i < n; ++i) wt.m[names[i]] = {static_cast<const float*>(dev_ptrs[i]), (long long)numel[i]};
  if (h->cfg.backbone == SPM_BACKBONE_VIT_B16) {
    SPM_TRY(load_vit(h, st, wt));
    if (h->fp32) SPM_TRY(load_vit32(h, st, wt));
  } else {
    auto getter = [&](const std::string& name, long long ne, const float** out) { return wt.get(name, ne, out); };
    SPM_TRY(rn50_create(&h->rn50, st, h->sms, getter));
  }
  if (h->cfg.head == SPM_HEAD_CLIPFSAR) SPM_TRY(load_head_fsar(h, st, wt));
  else if (h->cfg.head == SPM_HEAD_CPM2C) SPM_TRY(load_head_cpm2c(h, st, wt));
  else if (h->cfg.head == SPM_HEAD_CLIPSPM) SPM_TRY(load_head(h, st, wt));
  // SPM_HEAD_STEN: the shipped model has no parameters besides the backbone
  SPM_CUDA(cudaStreamSynchronize(st));
  h->weights_loaded = true;
  return 0;
}

int spm_set_text_features(spm_handle* h, void* stream, const float* table, int n_cls, int dim) {
  SPM_CHECK(h != nullptr && table != nullptr, "spm_set_text_features: null argument");
  SPM_CHECK(dim == h->D, "spm_set_text_features: feature dim does not match the backbone's mid_dim");
  SPM_CHECK(n_cls >= 1, "spm_set_text_features: empty table");
  if (n_cls > h->n_cls) SPM_TRY(drealloc_t(h, &h->text, (long long)n_cls * dim));
  SPM_CUDA(cudaMemcpyAsync(h->text, table, (size_t)n_cls * dim * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  h->n_cls = n_cls;
  h->text_set = true;
  return 0;
}

int spm_set_text_features_train(spm_handle* h, void* stream, const float* table, int n_cls, int dim) {
  SPM_CHECK(h != nullptr && table != nullptr, "spm_set_text_features_train: null argument");
  SPM_CHECK(dim == h->D, "spm_set_text_features_train: feature dim does not match the backbone's mid_dim");
  SPM_CHECK(n_cls >= 1, "spm_set_text_features_train: empty table");
  if (n_cls > h->n_cls_train) SPM_TRY(drealloc_t(h, &h->text_train, (long long)n_cls * dim));
  SPM_CUDA(cudaMemcpyAsync(h->text_train, table, (size_t)n_cls * dim * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  h->n_cls_train = n_cls;
  return 0;
}

int spm_class_logits(spm_handle* h, void* stream, int n_rows, int n_cls, float* out) {
  SPM_CHECK(h != nullptr && out != nullptr, "spm_class_logits: null argument");
  SPM_CHECK(h->cfg.head == SPM_HEAD_CLIPFSAR || h->cfg.head == SPM_HEAD_CPM2C,
            "spm_class_logits: only the CLIP-FSAR and CPM2C heads produce class logits");
  SPM_CHECK(h->cls_rows > 0, "spm_class_logits: no class logits available (no head call yet, or the text table / USE_CLASSIFICATION is not set)");
  const int want_cls = h->cfg.head == SPM_HEAD_CPM2C ? h->n_cls : h->n_cls_train;   // CPM2C evaluates on the test prompts
  SPM_CHECK(n_rows == h->cls_rows && n_cls == want_cls, "spm_class_logits: shape does not match the last head call");
  SPM_CUDA(cudaMemcpyAsync(out, h->CLS, (size_t)n_rows * n_cls * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int spm_cpm2c_outputs(spm_handle* h, void* stream, int n_episodes, int Q, int W, float* logits_local, float* logits_global) {
  SPM_CHECK(h != nullptr && logits_local != nullptr && logits_global != nullptr, "spm_cpm2c_outputs: null argument");
  SPM_CHECK(h->cfg.head == SPM_HEAD_CPM2C, "spm_cpm2c_outputs: the handle does not run the CPM2C head");
  SPM_CHECK(h->cp_last_E == n_episodes && h->cp_last_Q == Q && h->cp_last_W == W && n_episodes > 0,
            "spm_cpm2c_outputs: shape does not match the last head call");
  const size_t bytes = (size_t)n_episodes * Q * W * 4;
  SPM_CUDA(cudaMemcpyAsync(logits_local, h->CP_OUT_L, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  SPM_CUDA(cudaMemcpyAsync(logits_global, h->CP_OUT_G, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int spm_encode_frames(spm_handle* h, void* stream, const float* images, int n_frames, float* feats_out) {
  SPM_CHECK(h != nullptr, "spm_encode_frames: null handle");
  if (n_frames <= 0) return 0;  // empty input: nothing to do (pointers may be null)
  SPM_CHECK(images != nullptr && feats_out != nullptr, "spm_encode_frames: null argument");
  Segment seg;
  seg.images = images;
  seg.n_frames = n_frames;
  return encode_segments(h, (cudaStream_t)stream, &seg, 1, feats_out);
}

int spm_encode_frames_u8(spm_handle* h, void* stream, const uint8_t* frames, int n_frames, int H, int W,
                         float* feats_out) {
  SPM_CHECK(h != nullptr, "spm_encode_frames_u8: null handle");
  if (n_frames <= 0) return 0;
  SPM_CHECK(frames != nullptr && feats_out != nullptr, "spm_encode_frames_u8: null argument");
  SPM_CHECK(H > 0 && W > 0, "spm_encode_frames_u8: bad frame size");
  Segment seg;
  seg.images = nullptr;
  seg.n_frames = n_frames;
  seg.frames_u8 = frames;
  seg.H = H;
  seg.W = W;
  return encode_segments(h, (cudaStream_t)stream, &seg, 1, feats_out);
}

int spm_head(spm_handle* h, void* stream, int n_episodes, int S, int Q, int W, const float* su, const float* qu,
             const float* support_labels, const float* real_support, const float* real_target, float* logits_out,
             float* dists_out) {
  SPM_TRY(check_shapes(h, n_episodes, S, Q, W));
  SPM_CHECK(su && qu && support_labels && real_support && real_target && logits_out && dists_out, "spm_head: null argument");
  SPM_CHECK(h->weights_loaded, "spm_head: weights not loaded");
  cudaStream_t st = (cudaStream_t)stream;
  SPM_TRY(ensure_head_workspace(h, n_episodes, S, Q, W));
  SPM_TRY(reset_err_flag(h, st));
  const int T = h->cfg.seq_len, D = h->D, N = S + Q;
  const size_t row = (size_t)T * D * 4;
  SPM_CUDA(cudaMemcpy2DAsync(h->X, (size_t)N * row, su, (size_t)S * row, (size_t)S * row, n_episodes,
                             cudaMemcpyDeviceToDevice, st));
  SPM_CUDA(cudaMemcpy2DAsync(h->X + (long long)S * T * D, (size_t)N * row, qu, (size_t)Q * row, (size_t)Q * row,
                             n_episodes, cudaMemcpyDeviceToDevice, st));
  return head_run(h, st, n_episodes, S, Q, W, support_labels, real_support, real_target, nullptr, 1.f, logits_out,
                  dists_out, nullptr, nullptr, nullptr);
}

int spm_head_stage(spm_handle* h, void* stream, const char* name, float* out, long long capacity, long long* numel) {
  SPM_CHECK(h != nullptr && name != nullptr, "spm_head_stage: null argument");
  return head_stage(h, (cudaStream_t)stream, name, out, capacity, numel);
}

int spm_forward(spm_handle* h, void* stream, int n_episodes, int S, int Q, int W, const float* support_images,
                const float* target_images, const float* support_labels, const float* real_support,
                const float* real_target, float* logits_out, float* dists_out) {
  SPM_CHECK(support_images && target_images && support_labels && real_support && real_target && logits_out && dists_out,
            "spm_forward: null argument");
  SPM_CHECK(h != nullptr, "spm_forward: null handle");
  SPM_TRY(reset_err_flag(h, (cudaStream_t)stream));
  return forward_impl(h, (cudaStream_t)stream, n_episodes, S, Q, W, support_images, target_images, support_labels,
                      real_support, real_target, nullptr, 1.f, logits_out, dists_out, nullptr, nullptr, nullptr);
}

int spm_eval(spm_handle* h, void* stream, int n_episodes, int S, int Q, int W, const float* support_images,
             const float* target_images, const float* support_labels, const float* real_support,
             const float* real_target, const int64_t* target_labels, float tasks_per_batch, float* logits_out,
             float* dists_out, float* loss_out, float* acc_out, int32_t* pred_out) {
  SPM_CHECK(support_images && target_images && support_labels && real_support && real_target && target_labels,
            "spm_eval: null argument");
  SPM_CHECK(h != nullptr, "spm_eval: null handle");
  SPM_TRY(reset_err_flag(h, (cudaStream_t)stream));
  return forward_impl(h, (cudaStream_t)stream, n_episodes, S, Q, W, support_images, target_images, support_labels,
                      real_support, real_target, reinterpret_cast<const long long*>(target_labels), tasks_per_batch,
                      logits_out, dists_out, loss_out, acc_out, pred_out);
}

int spm_eval_u8(spm_handle* h, void* stream, int n_episodes, int S, int Q, int W, int img_h, int img_w,
                const uint8_t* support_frames, const uint8_t* target_frames, const float* support_labels,
                const float* real_support, const float* real_target, const int64_t* target_labels, float tasks_per_batch,
                float* logits_out, float* dists_out, float* loss_out, float* acc_out, int32_t* pred_out) {
  SPM_CHECK(support_frames && target_frames && support_labels && real_support && real_target, "spm_eval_u8: null argument");
  SPM_CHECK(h != nullptr, "spm_eval_u8: null handle");
  SPM_CHECK(img_h > 0 && img_w > 0, "spm_eval_u8: bad frame size");
  SPM_TRY(reset_err_flag(h, (cudaStream_t)stream));
  return forward_impl(h, (cudaStream_t)stream, n_episodes, S, Q, W, support_frames, target_frames, support_labels,
                      real_support, real_target, reinterpret_cast<const long long*>(target_labels), tasks_per_batch,
                      logits_out, dists_out, loss_out, acc_out, pred_out, img_h, img_w);
}

// frame_bytes: bytes of one input frame in host memory (fp32 image, or img_h x img_w x 3 uint8 when img_h > 0)
static int eval_host_impl(spm_handle* h, int n_episodes, int S, int Q, int W, const uint8_t* su_h, const uint8_t* qu_h,
                          long long frame_bytes, int img_h, int img_w, const float* lab_h, const float* rs_h,
                          const float* rt_h, const int64_t* tl_h, float tasks_per_batch, float* logits_h,
                          float* dists_h, float* loss_h, float* acc_h, int32_t* pred_h) {
  SPM_CHECK(h && su_h && qu_h && lab_h && rs_h && rt_h && tl_h, "spm_eval_host: null argument");
  SPM_TRY(check_shapes(h, 1, S, Q, W));
  const int T = h->cfg.seq_len;
  const int EC = std::max(1, h->cfg.max_episodes);  // episodes per compute chunk
  const int R = 2 * EC;                             // staging ring: R episode slots
  const long long fs = (long long)S * T, fq = (long long)Q * T;  // frames per episode
  if (h->copy_stream == nullptr) {
    SPM_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    SPM_CUDA(cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
  }
  spm_handle::Stage& s = h->stage[0];
  // every ring has its own capacity: the image rings are sized in bytes, the label / result rings by S, Q, Q*W and R
  // (a later call with the same frame counts but a larger Q*W must regrow the logits ring, not write past it)
  if (R * fs * frame_bytes > h->stage_cap_bytes_s || R * fq * frame_bytes > h->stage_cap_bytes_q ||
      (long long)R * S > h->stage_cap_S || (long long)R * Q > h->stage_cap_Q || (long long)R * Q * W > h->stage_cap_QW ||
      R > h->stage_cap_R) {
    SPM_CUDA(cudaDeviceSynchronize());   // nothing may still be using the rings that are being replaced
    if (R * fs * frame_bytes > h->stage_cap_bytes_s) {
      SPM_TRY(drealloc_t(h, &s.su, R * fs * frame_bytes));
      h->stage_cap_bytes_s = R * fs * frame_bytes;
    }
    if (R * fq * frame_bytes > h->stage_cap_bytes_q) {
      SPM_TRY(drealloc_t(h, &s.qu, R * fq * frame_bytes));
      h->stage_cap_bytes_q = R * fq * frame_bytes;
    }
    if ((long long)R * S > h->stage_cap_S) {
      SPM_TRY(drealloc_t(h, &s.lab, (long long)R * S));
      SPM_TRY(drealloc_t(h, &s.rs, (long long)R * S));
      h->stage_cap_S = (long long)R * S;
    }
    if ((long long)R * Q > h->stage_cap_Q) {
      SPM_TRY(drealloc_t(h, &s.rt, (long long)R * Q));
      SPM_TRY(drealloc_t(h, &s.tl, (long long)R * Q));
      SPM_TRY(drealloc_t(h, &s.pred, (long long)R * Q));
      h->stage_cap_Q = (long long)R * Q;
    }
    if ((long long)R * Q * W > h->stage_cap_QW) {
      SPM_TRY(drealloc_t(h, &s.logits, (long long)R * Q * W));
      h->stage_cap_QW = (long long)R * Q * W;
    }
    if (R > h->stage_cap_R) {
      SPM_TRY(drealloc_t(h, &s.dists, R));
      SPM_TRY(drealloc_t(h, &s.loss, R));
      SPM_TRY(drealloc_t(h, &s.acc, R));
      h->stage_cap_R = R;
    }
  }
  // prefetch buffers for the first chunk of the next call (allocated before anything is enqueued: cudaMalloc syncs)
  // (a whole compute chunk: with its copy out of the way the call can run full-size chunks from the start)
  const int first_n = std::min(EC, n_episodes);
  const void *hint_su = h->next_su, *hint_qu = h->next_qu;
  const int hint_n = std::min(EC, h->next_n);   // first chunk of the hinted call (it may hold fewer episodes)
  h->next_su = h->next_qu = nullptr;   // a hint is consumed by exactly one call
  h->next_n = 0;
  const long long pf_s = (long long)hint_n * fs * frame_bytes, pf_q = (long long)hint_n * fq * frame_bytes;
  const long long my_s = (long long)first_n * fs * frame_bytes, my_q = (long long)first_n * fq * frame_bytes;
  if (hint_su != nullptr && (pf_s > h->pf_cap_s || pf_q > h->pf_cap_q)) {
    SPM_CUDA(cudaStreamSynchronize(h->copy_stream));   // nobody may still be writing the old buffers
    SPM_CUDA(cudaStreamSynchronize(h->compute_stream));
    SPM_TRY(drealloc_t(h, &h->pf_su, pf_s));
    SPM_TRY(drealloc_t(h, &h->pf_qu, pf_q));
    SPM_TRY(drealloc_t(h, &h->pf_su_alt, pf_s));
    SPM_TRY(drealloc_t(h, &h->pf_qu_alt, pf_q));
    h->pf_cap_s = pf_s; h->pf_cap_q = pf_q;
    h->pf_src_su = h->pf_src_qu = nullptr;
  }
  if (h->pf_event == nullptr) SPM_CUDA(cudaEventCreateWithFlags(&h->pf_event, cudaEventDisableTiming));
  // does the prefetch made by the previous call hold this call's first chunk?
  const bool use_pf = h->pf_src_su == (const void*)su_h && h->pf_src_qu == (const void*)qu_h && h->pf_bytes_s == my_s &&
                      h->pf_bytes_q == my_q && su_h != nullptr;
  // Chunk schedule.  The H2D copy of a chunk can only overlap the compute of EARLIER chunks, so the first chunks
  // are small (1, 1, 2, 4, ... up to EC when EC is a power of two: offsets stay aligned, a chunk never wraps the
  // ring) -- only one episode's copy is exposed per call instead of EC episodes'.
  std::vector<int> starts;
  {
    const bool pow2 = (EC & (EC - 1)) == 0 && !use_pf;   // first chunk prefetched: full-size chunks throughout
    int e = 0, sz = pow2 ? 1 : EC;
    bool first = true;
    while (e < n_episodes) {
      starts.push_back(e);
      e += std::min(sz, n_episodes - e);
      if (pow2 && sz < EC) { if (first) first = false; else sz *= 2; }
    }
    starts.push_back(n_episodes);
  }
  const int n_chunks = (int)starts.size() - 1;
  while ((int)h->ev_copied.size() < n_chunks) {
    cudaEvent_t a, b;
    SPM_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    SPM_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    h->ev_copied.push_back(a); h->ev_done.push_back(b);
  }
  const long long per_ep = (long long)Q * W + 3 + Q;  // logits, dists, loss, acc, pred (int32 in a float slot)
  if ((long long)n_episodes * per_ep > h->pin_cap) {
    if (h->pin_res) cudaFreeHost(h->pin_res);
    SPM_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->pin_res), (size_t)n_episodes * per_ep * 4));
    h->pin_cap = (long long)n_episodes * per_ep;
  }
  float* p_logits = h->pin_res;
  float* p_dists = p_logits + (long long)n_episodes * Q * W;
  float* p_loss = p_dists + n_episodes;
  float* p_acc = p_loss + n_episodes;
  int32_t* p_pred = reinterpret_cast<int32_t*>(p_acc + n_episodes);
  cudaStream_t cs = h->copy_stream, ks = h->compute_stream;
  SPM_TRY(reset_err_flag(h, ks));
  std::vector<int> chunk_of(n_episodes);
  for (int c = 0; c < n_chunks; ++c)
    for (int e = starts[c]; e < starts[c + 1]; ++e) chunk_of[e] = c;
  for (int c = 0; c < n_chunks; ++c) {
    const int e0 = starts[c], E = starts[c + 1] - e0, slot = e0 % R;
    // ring slot reuse: the chunk that last used these slots must have been consumed
    if (e0 + E - 1 >= R) SPM_CUDA(cudaStreamWaitEvent(cs, h->ev_done[chunk_of[e0 + E - 1 - R]], 0));
    const bool from_pf = use_pf && c == 0 && E == first_n;   // this chunk's frames were copied by the previous call
    const uint8_t* su_d = from_pf ? h->pf_su : s.su + slot * fs * frame_bytes;
    const uint8_t* qu_d = from_pf ? h->pf_qu : s.qu + slot * fq * frame_bytes;
    if (from_pf) {
      SPM_CUDA(cudaStreamWaitEvent(ks, h->pf_event, 0));
    } else {
      SPM_CUDA(cudaMemcpyAsync(s.su + slot * fs * frame_bytes, su_h + e0 * fs * frame_bytes,
                               (size_t)(E * fs * frame_bytes), cudaMemcpyHostToDevice, cs));
      SPM_CUDA(cudaMemcpyAsync(s.qu + slot * fq * frame_bytes, qu_h + e0 * fq * frame_bytes,
                               (size_t)(E * fq * frame_bytes), cudaMemcpyHostToDevice, cs));
    }
    SPM_CUDA(cudaMemcpyAsync(s.lab + (long long)slot * S, lab_h + (long long)e0 * S, (size_t)E * S * 4, cudaMemcpyHostToDevice, cs));
    SPM_CUDA(cudaMemcpyAsync(s.rs + (long long)slot * S, rs_h + (long long)e0 * S, (size_t)E * S * 4, cudaMemcpyHostToDevice, cs));
    SPM_CUDA(cudaMemcpyAsync(s.rt + (long long)slot * Q, rt_h + (long long)e0 * Q, (size_t)E * Q * 4, cudaMemcpyHostToDevice, cs));
    SPM_CUDA(cudaMemcpyAsync(s.tl + (long long)slot * Q, tl_h + (long long)e0 * Q, (size_t)E * Q * 8, cudaMemcpyHostToDevice, cs));
    SPM_CUDA(cudaEventRecord(h->ev_copied[c], cs));
    SPM_CUDA(cudaStreamWaitEvent(ks, h->ev_copied[c], 0));
    float* lg = s.logits + (long long)slot * Q * W;
    SPM_TRY(forward_impl(h, ks, E, S, Q, W, su_d, qu_d,
                         s.lab + (long long)slot * S, s.rs + (long long)slot * S, s.rt + (long long)slot * Q,
                         s.tl + (long long)slot * Q, tasks_per_batch, lg, s.dists + slot, s.loss + slot, s.acc + slot,
                         s.pred + (long long)slot * Q, img_h, img_w));
    SPM_CUDA(cudaMemcpyAsync(p_logits + (long long)e0 * Q * W, lg, (size_t)E * Q * W * 4, cudaMemcpyDeviceToHost, ks));
    SPM_CUDA(cudaMemcpyAsync(p_dists + e0, s.dists + slot, (size_t)E * 4, cudaMemcpyDeviceToHost, ks));
    SPM_CUDA(cudaMemcpyAsync(p_loss + e0, s.loss + slot, (size_t)E * 4, cudaMemcpyDeviceToHost, ks));
    SPM_CUDA(cudaMemcpyAsync(p_acc + e0, s.acc + slot, (size_t)E * 4, cudaMemcpyDeviceToHost, ks));
    SPM_CUDA(cudaMemcpyAsync(p_pred + (long long)e0 * Q, s.pred + (long long)slot * Q, (size_t)E * Q * 4, cudaMemcpyDeviceToHost, ks));
    SPM_CUDA(cudaEventRecord(h->ev_done[c], ks));
  }
  // Behind this call's own copies (same FIFO copy stream): the first chunk of the next call, while this call computes.
  // This call's chunk 0 may be reading the current prefetch buffers, so the copy goes into the other set (last read by the
  // PREVIOUS call, which synchronised its compute stream before returning) and the two sets swap roles.
  h->pf_src_su = h->pf_src_qu = nullptr;
  if (hint_su != nullptr && hint_qu != nullptr && hint_n > 0) {
    std::swap(h->pf_su, h->pf_su_alt);
    std::swap(h->pf_qu, h->pf_qu_alt);
    SPM_CUDA(cudaMemcpyAsync(h->pf_su, hint_su, (size_t)pf_s, cudaMemcpyHostToDevice, cs));
    SPM_CUDA(cudaMemcpyAsync(h->pf_qu, hint_qu, (size_t)pf_q, cudaMemcpyHostToDevice, cs));
    SPM_CUDA(cudaEventRecord(h->pf_event, cs));
    h->pf_src_su = hint_su; h->pf_src_qu = hint_qu;
    h->pf_bytes_s = pf_s; h->pf_bytes_q = pf_q;
  }
  SPM_CUDA(cudaStreamSynchronize(ks));
  if (logits_h) memcpy(logits_h, p_logits, (size_t)n_episodes * Q * W * 4);
  if (dists_h) memcpy(dists_h, p_dists, (size_t)n_episodes * 4);
  if (loss_h) memcpy(loss_h, p_loss, (size_t)n_episodes * 4);
  if (acc_h) memcpy(acc_h, p_acc, (size_t)n_episodes * 4);
  if (pred_h) memcpy(pred_h, p_pred, (size_t)n_episodes * Q * 4);
  int flag = 0;
  SPM_CUDA(cudaMemcpy(&flag, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  SPM_CHECK(flag != 2, "spm_eval_host: a real_support / real_target class id lies outside the text-feature table");
  SPM_CHECK(flag == 0, "spm_eval_host: an episode's number of distinct support labels differs from `W`");
  return 0;
}

int spm_eval_host_set_next(spm_handle* h, const void* next_support_host, const void* next_target_host,
                           int next_n_episodes) {
  SPM_CHECK(h != nullptr, "spm_eval_host_set_next: null handle");
  SPM_CHECK(next_n_episodes >= 0, "spm_eval_host_set_next: negative episode count");
  const bool on = next_support_host != nullptr && next_target_host != nullptr && next_n_episodes > 0;
  h->next_su = on ? next_support_host : nullptr;
  h->next_qu = on ? next_target_host : nullptr;
  h->next_n = on ? next_n_episodes : 0;
  return 0;
}

int spm_eval_host(spm_handle* h, int n_episodes, int S, int Q, int W, const float* su_h, const float* qu_h,
                  const float* lab_h, const float* rs_h, const float* rt_h, const int64_t* tl_h, float tasks_per_batch,
                  float* logits_h, float* dists_h, float* loss_h, float* acc_h, int32_t* pred_h) {
  return eval_host_impl(h, n_episodes, S, Q, W, reinterpret_cast<const uint8_t*>(su_h),
                        reinterpret_cast<const uint8_t*>(qu_h), (long long)FRAME_ELEMS * 4, 0, 0, lab_h, rs_h, rt_h, tl_h,
                        tasks_per_batch, logits_h, dists_h, loss_h, acc_h, pred_h);
}

int spm_eval_host_u8(spm_handle* h, int n_episodes, int S, int Q, int W, int img_h, int img_w, const uint8_t* su_h,
                     const uint8_t* qu_h, const float* lab_h, const float* rs_h, const float* rt_h,
                     const int64_t* tl_h, float tasks_per_batch, float* logits_h, float* dists_h, float* loss_h,
                     float* acc_h, int32_t* pred_h) {
  SPM_CHECK(img_h > 0 && img_w > 0, "spm_eval_host_u8: bad frame size");
  return eval_host_impl(h, n_episodes, S, Q, W, su_h, qu_h, (long long)img_h * img_w * 3, img_h, img_w, lab_h, rs_h, rt_h,
                        tl_h, tasks_per_batch, logits_h, dists_h, loss_h, acc_h, pred_h);
}

int spm_softdtw_forward(void* stream, int n_pairs, int N, int M, const float* D, float gamma, float bandwidth, float* R,
                        float* out) {
  SPM_CHECK(D && R, "spm_softdtw_forward: null argument");
  SPM_CHECK(gamma > 0.f, "spm_softdtw_forward: gamma must be positive");
  SPM_CHECK(N >= 1 && M >= 1 && N <= 1024 && M <= 1024, "spm_softdtw_forward: sequence lengths must be in [1, 1024]");
  SPM_KERNEL(k_softdtw_forward((cudaStream_t)stream, D, n_pairs, N, M, gamma, bandwidth, R, out));
  return 0;
}

int spm_softdtw_backward(void* stream, int n_pairs, int N, int M, const float* D, const float* R, float gamma,
                         float bandwidth, float* E) {
  SPM_CHECK(D && R && E, "spm_softdtw_backward: null argument");
  SPM_CHECK(gamma > 0.f, "spm_softdtw_backward: gamma must be positive");
  SPM_CHECK(N >= 1 && M >= 1 && N <= 1024 && M <= 1024, "spm_softdtw_backward: sequence lengths must be in [1, 1024]");
  SPM_KERNEL(k_softdtw_backward((cudaStream_t)stream, D, R, n_pairs, N, M, gamma, bandwidth, E));
  return 0;
}

int spm_otam_distance(void* stream, int n_pairs, int W, int Q, int T, int D, const float* support, const float* target,
                      int single_direct, float alpha, float beta, float* out) {
  SPM_CHECK(support && target && out, "spm_otam_distance: null argument");
  static bool inited = false;
  if (!inited) { SPM_KERNEL(k_otam_init()); inited = true; }
  SPM_KERNEL(k_otam((cudaStream_t)stream, support, (long long)W * T * D, (long long)T * D, D, target,
                    (long long)Q * T * D, (long long)T * D, D, n_pairs, W, Q, T, D, single_direct, alpha, beta, out));
  return 0;
}

}  // extern "C"

