/*
 * clipspm_b200 -- C ABI of the B200-native CLIP-SPM episode-evaluation hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes (no torch types), returns an int
 * status (0 = ok) and records a message readable with spm_last_error().  All tensor arguments are
 * DEVICE pointers unless the name says `host`; the caller owns every buffer; the library owns only
 * the opaque handle (packed weights + workspace).  `stream` is a cudaStream_t passed as void*.
 * A handle is not thread-safe: one handle per (process, device).  The *_host entry points run on the handle's own streams and
 * share its workspaces with the stream-taking ones: synchronise the caller's stream before switching from a device-pointer
 * call to a host call on the same handle (the Python mirror does).
 *
 * Reference interfaces replaced (paths relative to the reference repo root):
 *   spm_create / spm_destroy      models/model_clipspm.py:15-101   CNN.__init__ (module construction)
 *   spm_load_weights              torch.nn.Module.load_state_dict of that CNN (keys listed in SURVEY.md 8b)
 *   spm_set_text_features         models/model_clipspm.py:60,70    text_features_{train,test} attributes
 *   spm_forward                   models/model_clipspm.py:111-144  CNN.forward(inputs) -> logits, dists
 *   spm_eval                      run/main_run.py:390-392 + utils/utils.py:174-186,259-264 (loss, accuracy)
 *   spm_eval_host                 run/main_run.py:266-279 (prepare_task H2D + forward + loss/acc + .item())
 *   spm_eval_host_set_next        run/main_run.py:71 (DataLoader prefetch: the next batch is known while this one runs)
 *   spm_encode_frames             models/clip_fsar.py:672-689 VisionTransformer.forward / :593-608 ModifiedResNet
 *   spm_head                      models/model_clipspm.py:125-143  (everything after get_feats)
 *   spm_head_stage                the intermediate tensors of :125-143 (test hook; the reference exposes them as locals)
 *   spm_otam_distance             models/model_clipspm.py:348-362 + models/myRes.py:756-765,821-855
 *   spm_set_text_features_train / spm_class_logits   models/model_clipfsar.py:127,329-331 (sibling head CLIP-FSAR)
 *   spm_softdtw_forward/backward  models/OTAM.py:34-203 (TA2N's numba.cuda soft-DTW kernels, _SoftDTWCUDA)
 *   spm_jpeg_info / spm_jpeg_decode   video_reader.py:227-230 read_single_image (PIL JPEG decode of every frame)
 *   spm_adam_* / spm_sgd_step / spm_scaler_update    run/main_run.py:84-96,76,207-209 (torch.optim.Adam / SGD + GradScaler)
 *   spm_tv1_* / spm_linear_backward   autograd through models/myRes.py:1053-1075 Transformer_v1 and the head's nn.Linear layers
 *   spm_vitblock_* / spm_layernorm_*  autograd through models/clip_fsar.py:622-643,664,668 (the ViT-B/16 tower's blocks, ln_pre / ln_post)
 *   spm_dropout / spm_tv1_set_dropout / spm_dropout_seed_source   nn.Dropout of models/myRes.py:961-996 in train mode
 *   spm_transform_frames / spm_transform_frames_train   video_reader.py:83-111 (the loader's test / training transform)
 *   spm_gemm                      ATen linear / conv-as-GEMM calls (cuBLASLt) under all of the above
 */
#ifndef CLIPSPM_B200_H
#define CLIPSPM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPM_ABI_VERSION 8 /* 2: spm_config gained `head`, `cls_value`; sibling heads, soft-DTW, spm_eval_host_set_next
                           * 3: spm_head_stage (per-stage taps for the parity tests)
                           * 4: spm_jpeg_info / spm_jpeg_decode, spm_eval_u8
                           * 5: SPM_HEAD_CPM2C: spm_config gained its parameters; spm_cpm2c_outputs
                           * 6: spm_config gained fsar_depth, fsar_merge_before (CLIP-FSAR's optional branches)
                           * 7: spm_adam_*, spm_scaler_update
                           * 8: spm_tv1_* (head transformer block forward + backward), spm_linear_backward, spm_dropout,
                           *    spm_vitblock_* (encoder block forward + backward), spm_layernorm_forward / _backward */

typedef struct spm_handle spm_handle;

enum { SPM_BACKBONE_VIT_B16 = 0, SPM_BACKBONE_RN50 = 1 };
/* arithmetic of the dense contractions: bf16 tensor cores with fp32 accumulation (the reference's autocast(bfloat16)
 * mode), or the exact-arithmetic parity mode: fp32 operands, every product an fp32 FFMA on the CUDA cores
 * (csrc/sgemm_f32.cu; Blackwell has no fp32 tensor-core MMA), ViT-B/16 backbone only */
enum { SPM_PRECISION_BF16 = 0, SPM_PRECISION_FP32 = 1,
       /* bf16 tensor cores AND a bf16 residual stream (ViT-B/16 only): the arithmetic the reference's own
        * autocast(bfloat16) forward has (run/main_run.py:274 around models/clip_fsar.py:672-689: conv1, every Linear and
        * every `x + ...` produce bf16).  SPM_PRECISION_BF16 keeps the residual stream and LayerNorm inputs in fp32 instead
        * (tighter parity, ~8 % slower); both stay inside the 2e-2 tolerance of the goldens. */
       SPM_PRECISION_BF16_RESID = 2 };
/* metric head behind the same entry points: CLIP-SPM (models/model_clipspm.py, the hot path) or its sibling
 * CLIP-FSAR (models/model_clipfsar.py CNN_OTAM_CLIPFSAR, evaluation branch :325-383), which reuses the same
 * transformer-block, class-mean and OTAM kernels */
enum { SPM_HEAD_CLIPSPM = 0, SPM_HEAD_CLIPFSAR = 1, SPM_HEAD_STEN = 2 /* models/model_sten.py:62-113 as shipped */,
       SPM_HEAD_CPM2C = 3 /* models/model_cpm2c.py CLIP_CPMMC_FSAR, evaluation forward :207-312 */ };

typedef struct spm_config {
  int backbone;         /* cfg.MODEL.BACKBONE: SPM_BACKBONE_*                         model_clipspm.py:18,24 */
  int seq_len;          /* cfg.DATA.SEQ_LEN (T frames per video)                      model_clipspm.py:154   */
  int n_text_classes;   /* rows of the text-feature table                             model_clipspm.py:60,70 */
  float mid_dim_text;   /* cfg.params['mid_dim_text']   (gate_text hidden = D * this) model_clipspm.py:89    */
  float mid_dim_vision; /* cfg.params['mid_dim_vision']                               model_clipspm.py:95    */
  float negative_slope; /* cfg.params['negative_slope'] (LeakyReLU)                   model_clipspm.py:90,96 */
  float alpha;          /* cfg.params['alpha']                                        model_clipspm.py:304   */
  int single_direct;    /* cfg.MODEL.SINGLE_DIRECT (0 = bidirectional OTAM)           model_clipspm.py:358   */
  int precision;        /* SPM_PRECISION_*                                                                   */
  int max_episodes;     /* episodes per spm_forward call the workspace is sized for                          */
  int max_support;      /* S = way*shot support videos per episode (upper bound)                             */
  int max_query;        /* Q query videos per episode (upper bound)                                          */
  int max_way;          /* W distinct support labels per episode (upper bound)                               */
  int head;             /* SPM_HEAD_* (cfg.MODEL.NAME 'clipspm' / 'clipfsar')          run/main_run.py:123-130 */
  float cls_value;      /* cfg.MODEL.USE_CLASSIFICATION_VALUE (CLIP-FSAR loss only)    run/main_run.py:356    */
  /* SPM_HEAD_CPM2C only (ABI 5) */
  float motion_residual_ratio; /* cfg.params['motion_residual_ratio']                  model_cpm2c.py:175     */
  float lambdas[4];     /* cfg.params['lambdas0'..'lambdas3'] (loss / total-logit weights) run/main_run.py:372-376 */
  float motion_coeff;   /* cfg.MODEL.MOTION_COFF                                        model_cpm2c.py:87      */
  float normal_coeff;   /* cfg.MODEL.NORMAL_COFF                                        model_cpm2c.py:88      */
  int use_classification; /* cfg.MODEL.USE_CLASSIFICATION (class_logits branch)          model_cpm2c.py:426     */
  /* SPM_HEAD_CLIPFSAR only (ABI 6); 0 reads as the default (one layer, no merge) */
  int fsar_depth;       /* layers of context2: cfg.TRAIN.TRANSFORMER_DEPTH when cfg.MODEL.TRANSFORMER_DEPTH is set
                         *                                                            model_clipfsar.py:143-144 */
  int fsar_merge_before; /* cfg.MODEL.MERGE_BEFORE: class means of frames and prompts BEFORE context2
                         *                                                            model_clipfsar.py:341-349 */
} spm_config;

const char* spm_last_error(void);
int spm_abi_version(void);

/* Instrumentation for bench.py: number of this library's kernel launches so far (process-wide), and CUDA-event
 * timing of every GEMM launch between begin/end.  spm_profile_end synchronises the device and returns, per GEMM
 * flavour t = kind*2 + (tile N == 256) [kind 0 = bf16, 1 = tf32], the summed algorithmic FLOPs (2MNK), the summed
 * launch durations in ms and the launch count. */
long long spm_launch_count(void);
int spm_profile_begin(int max_records);
int spm_profile_disarm(void); /* stop timing further launches (no sync); spm_profile_end still reads what was recorded */
int spm_profile_end(double* flops4, double* ms4, int* count4);

int spm_create(const spm_config* cfg, spm_handle** out);
int spm_destroy(spm_handle* h);

/* Weights: `n` fp32 contiguous device tensors keyed by the reference state_dict names
 * (e.g. "backbone.transformer.resblocks.0.attn.in_proj_weight", "context2.layers.0.0.fn.to_q.weight").
 * The library repacks them (bf16 copies, fused QKV, conv->GEMM layouts, folded BatchNorm) into its own storage;
 * the caller's tensors are not referenced after the call returns (the call synchronises `stream`). */
int spm_load_weights(spm_handle* h, void* stream, int n, const char* const* names, const void* const* dev_ptrs,
                     const int64_t* numel);
/* text feature table [n_cls, D] fp32 (the reference's text_features_test / _train) */
int spm_set_text_features(spm_handle* h, void* stream, const float* table, int n_cls, int dim);

/* CLIP-FSAR only: text_features_train [n_cls, D] (models/model_clipfsar.py:127) for class_text_logits (:331), and the
 * class_text_logits [n_rows = E*(S+Q), n_cls] of the most recent spm_head / spm_forward / spm_eval call (rows of an
 * episode: its S supports, then its Q queries -- torch.cat([support_features, target_features]), :329).
 * With SPM_HEAD_CLIPFSAR, spm_eval's loss is run/main_run.py:355-356:
 *   (sum_q CE(logits) + cls_value * sum_v CE(class_logits[v], real label of v)) / tasks_per_batch, dists_out = 0. */
int spm_set_text_features_train(spm_handle* h, void* stream, const float* table, int n_cls, int dim);
int spm_class_logits(spm_handle* h, void* stream, int n_rows, int n_cls, float* out);

/* CPM2C only: the per-branch outputs of the most recent head / forward / eval call (models/model_cpm2c.py:229-236):
 * logits_local [E,Q,W] (frame alignment), logits_global [E,Q,W] (token matching); the generic logits_out of that call
 * is lambdas1 * local + lambdas2 * global (the logits run/main_run.py:373 takes the accuracy on), dists_out is
 * target_consist_distance, and spm_class_logits returns class_logits [E*(S+Q), n_cls] (test prompts). */
int spm_cpm2c_outputs(spm_handle* h, void* stream, int n_episodes, int Q, int W, float* logits_local, float* logits_global);

/* Frame encoder: images [F,3,224,224] fp32 NCHW in [0,1] -> features [F, D] fp32 */
int spm_encode_frames(spm_handle* h, void* stream, const float* images, int n_frames, float* feats_out);

/* Metric head on precomputed frame features, n_episodes episodes at once.
 *   su [E,S,T,D], qu [E,Q,T,D] fp32; support_labels [E,S], real_support [E,S], real_target [E,Q] fp32
 *   (labels arrive as float tensors: video_reader.py:322-326).
 *   W = number of distinct support labels of every episode (the reference derives it with torch.unique, a host
 *   sync; here the caller states it and the kernels verify it: a mismatch yields NaN logits).
 *   logits_out [E,Q,W] (column w <-> w-th smallest distinct support label), dists_out [E]. */
int spm_head(spm_handle* h, void* stream, int n_episodes, int S, int Q, int W, const float* su, const float* qu,
             const float* support_labels, const float* real_support, const float* real_target, float* logits_out,
             float* dists_out);

/* Test hook (SURVEY.md 8b "per-stage entry points for tests"): one named stage tensor of the most recent CLIP-SPM head
 * pass (spm_head / spm_forward / spm_eval, any episode count), gathered into the
 * reference's layout, episodes stacked along dim 0.  Names and the reference lines that produce them
 * (models/model_clipspm.py): su_mo, qu_mo [E,S|Q,D] :195 get_motion_feats; support_token [E,S,D] :120;
 * target_token [E,Q,D] :216 token_tr; su_real [E,S,T,D], qu_fake [E,Q,T,D], token_s_real [E,S,D],
 * token_q_fake [E,Q,D] :218-223 se_te; su_pro [E,W,T,D] :231-239; su_2 [E,S,T,D], qu_2 [E,Q,T,D], su_t2 [E,W,T,D],
 * qu_t2 [E,1,T,D] :275-294 taskM; su_pro2 [E,W,T,D] :133-137; su_mo_refined / qu_mo_refined, su_mo_token /
 * qu_mo_token :197 (se_te of the mo call), su_mo2 / qu_mo2 :200.
 * *numel receives the element count; out == NULL only queries it; capacity is out's size in floats. */
int spm_head_stage(spm_handle* h, void* stream, const char* name, float* out, long long capacity, long long* numel);

/* CNN.forward for n_episodes episodes: support_images [E,S*T,3,224,224], target_images [E,Q*T,3,224,224] */
int spm_forward(spm_handle* h, void* stream, int n_episodes, int S, int Q, int W, const float* support_images,
                const float* target_images, const float* support_labels, const float* real_support,
                const float* real_target, float* logits_out, float* dists_out);

/* forward + loss/accuracy of run/main_run.py:390-392: per episode
 *   loss_out[e] = sum_q CE(logits[e,q,:], target_labels[e,q]) / tasks_per_batch + 0.001 * dists[e]
 *   acc_out[e]  = mean_q [argmax_w logits[e,q,w] == target_labels[e,q]];  pred_out [E,Q] int32 (may be null) */
int spm_eval(spm_handle* h, void* stream, int n_episodes, int S, int Q, int W, const float* support_images,
             const float* target_images, const float* support_labels, const float* real_support,
             const float* real_target, const int64_t* target_labels, float tasks_per_batch, float* logits_out,
             float* dists_out, float* loss_out, float* acc_out, int32_t* pred_out);

/* Same with HOST buffers (pinned or pageable): stages the host->device copies of every episode on a copy
 * stream overlapped with compute, runs spm_eval, and copies logits/dists/loss/acc/pred back to host.
 * Blocks until the results are in the host buffers. */
int spm_eval_host(spm_handle* h, int n_episodes, int S, int Q, int W, const float* support_images_host,
                  const float* target_images_host, const float* support_labels_host, const float* real_support_host,
                  const float* real_target_host, const int64_t* target_labels_host, float tasks_per_batch,
                  float* logits_host, float* dists_host, float* loss_host, float* acc_host, int32_t* pred_host);
/* Double buffering across calls (what a DataLoader with prefetching gives the reference's Learner.test loop,
 * run/main_run.py:71,266): tell the library which host image buffers the NEXT spm_eval_host / spm_eval_host_u8 call
 * will read and how many episodes they hold (same S, Q, frame format as the current call).  The current call then
 * copies that call's first chunk host->device behind its own copies, while its last chunks compute, and the next call
 * starts computing at once.  The hint is consumed by one call; the buffers must not change until the call that reads
 * them returns; a call whose buffers or episode count do not match the prefetch ignores it. */
int spm_eval_host_set_next(spm_handle* h, const void* next_support_host, const void* next_target_host,
                           int next_n_episodes);

/* cos_sim + (bi)directional OTAM soft-DTW of n_pairs independent problems:
 *   support [P,W,T,D], target [P,Q,T,D] fp32 -> out [P,Q,W] (accumulated: out = beta*out + alpha*otam) */
int spm_otam_distance(void* stream, int n_pairs, int W, int Q, int T, int D, const float* support,
                      const float* target, int single_direct, float alpha, float beta, float* out);

/* Soft-DTW of TA2N: the sm_100a counterparts of the reference's two numba.cuda kernels (models/OTAM.py:34-130) behind
 * _SoftDTWCUDA.forward / .backward (:134-203).  D [n_pairs, N, M] fp32 pairwise distances; bandwidth = Sakoe-Chiba
 * pruning (0 = none).  Forward writes the whole cumulative table R [n_pairs, N+2, M+2] (+inf borders, R[0,0] = 0;
 * the tensor _SoftDTWCUDA saves for backward) and out [n_pairs] = R[:, N, M] (out may be null).  Backward takes that
 * R unmodified (the reference's in-place edits :160-162 are applied on the fly) and writes E [n_pairs, N, M] =
 * d out / d D. */
int spm_softdtw_forward(void* stream, int n_pairs, int N, int M, const float* D, float gamma, float bandwidth, float* R,
                        float* out);
int spm_softdtw_backward(void* stream, int n_pairs, int N, int M, const float* D, const float* R, float gamma,
                         float bandwidth, float* E);

/* Backward of spm_otam_distance (first piece of the training step, run/main_run.py:245-254): given
 * grad_out [P,Q,W] = d loss / d out, writes d loss / d support [P,W,T,D] and d loss / d target [P,Q,T,D]
 * (overwritten, not accumulated) -- what autograd derives through otam_distance / cos_sim / OTAM_cum_dist_v2. */
int spm_otam_distance_backward(void* stream, int n_pairs, int W, int Q, int T, int D, const float* support,
                               const float* target, int single_direct, float alpha, const float* grad_out,
                               float* grad_support, float* grad_target);

/* Training step, head blocks (run/main_run.py:245-254 `scaler.scale(loss).backward()` through the metric head):
 * forward AND backward of models/myRes.py:1053-1075 `Transformer_v1` (depth 1, q = k = v = x; LayerNorm -> to_q / to_k / to_v ->
 * softmax attention -> to_out + x -> FeedForward + residual) -- `context1` / `context2` of CLIP-SPM, `context2` of CLIP-FSAR and
 * CPM2C -- with dropout p = 0.  x, out, grads: fp32 [n_seq * seq_len, D] (seq_len <= 48).  A handle keeps the activations
 * of its LAST forward (and the caller's x pointer, which must stay alive) for the one backward that follows; weight
 * gradients are OVERWRITTEN, g_wq / g_wk / g_wv must be three consecutive [heads*dim_head, D] blocks.
 * precision: 0 = tf32 tensor-core products, 1 = exact fp32. */
typedef struct spm_tv1 spm_tv1;
int spm_tv1_create(int D, int heads, int dim_head, int mlp_dim, int precision, spm_tv1** out);
int spm_tv1_destroy(spm_tv1* h);
int spm_tv1_load_weights(spm_tv1* h, void* stream, const float* ln_g, const float* ln_b, const float* wq, const float* wk,
                         const float* wv, const float* wout, const float* bout, const float* w0, const float* b0,
                         const float* w3, const float* b3);   /* device pointers, reference layouts ([out, in]) */
int spm_tv1_forward(spm_tv1* h, void* stream, const float* x, int n_seq, int seq_len, float* out);
/* nn.Dropout of the block in the reference's train mode (models/myRes.py:961-962 after to_out: p_atte; :990,992 after GELU
 * and after net.3: p_ffn; model_clipspm.py:80-81 builds the blocks with 0.2 / 0.05) for the forwards that follow.  The masks
 * are a pure function of (seed, site, element index): Philox4x32-10 with counter (index / 4, site), key = seed, word
 * index % 4; keep iff (word >> 8) * 2^-24 >= p; kept values times 1 / (1 - p) -- replayable by the oracle, never stored.
 * spm_dropout applies the same mask to a plain tensor (y may alias x); calling it on the upstream gradient is its backward. */
int spm_tv1_set_dropout(spm_tv1* h, float p_atte, float p_ffn, unsigned long long seed);
int spm_dropout(void* stream, const float* x, long long n, float p, unsigned long long seed, unsigned site, float* y);
/* CUDA graphs: a captured training step replays with the kernel arguments it was captured with, the seed included.  With a
 * device counter registered here every dropout kernel of the process adds `*device_counter` to its seed when it RUNS, so a
 * step that increments the counter on the device (inside the graph) draws fresh masks at every replay; null: off. */
int spm_dropout_seed_source(const unsigned long long* device_counter);
int spm_tv1_backward(spm_tv1* h, void* stream, const float* grad_out, float* grad_x, float* g_ln_g, float* g_ln_b,
                     float* g_wq, float* g_wk, float* g_wv, float* g_wout, float* g_bout, float* g_w0, float* g_b0,
                     float* g_w3, float* g_b3);

/* The same block with the frame encoder's switches: models/clip_fsar.py:622-643 `ResidualAttentionBlock` of the CLIP ViT-B/16
 * tower (x + out_proj(attn(ln_1 x)); + c_proj(QuickGELU(c_fc(ln_2 .)))), 197 tokens per frame, 12 heads x 64 -- forward and
 * backward for the training step (the reference's optimiser steps the tower too, run/main_run.py:84-88).  Parameters in the
 * reference's layouts: attn.in_proj_weight [2304,768] / in_proj_bias, attn.out_proj, ln_1, ln_2, mlp.c_fc, mlp.c_proj.
 * x, out, grads: fp32 [n_frames * 197, 768]; same handle rules as the spm_tv1_* block: one backward per forward, gradients overwritten,
 * spm_tv1_set_dropout / spm_tv1_destroy apply; the backward scratch is shared process-wide, so training calls belong on
  * one stream. */
int spm_vitblock_create(int precision, spm_tv1** out);
int spm_vitblock_load_weights(spm_tv1* h, void* stream, const float* ln1_g, const float* ln1_b, const float* in_proj_w,
                              const float* in_proj_b, const float* out_w, const float* out_b, const float* ln2_g,
                              const float* ln2_b, const float* fc_w, const float* fc_b, const float* proj_w, const float* proj_b);
int spm_vitblock_forward(spm_tv1* h, void* stream, const float* x, int n_frames, float* out);
int spm_vitblock_backward(spm_tv1* h, void* stream, const float* grad_out, float* grad_x, float* g_ln1_g, float* g_ln1_b,
                          float* g_in_proj_w, float* g_in_proj_b, float* g_out_w, float* g_out_b, float* g_ln2_g, float* g_ln2_b,
                          float* g_fc_w, float* g_fc_b, float* g_proj_w, float* g_proj_b);

/* nn.LayerNorm (eps 1e-5) over rows of C, forward and backward (ln_pre / ln_post of the tower, clip_fsar.py:664,668):
 * dx, dgamma, dbeta overwritten; workspace: (rows + 64) * C floats. */
int spm_layernorm_forward(void* stream, const float* x, int rows, int C, const float* gamma, const float* beta, float* y);
int spm_layernorm_backward(void* stream, const float* x, const float* dy, const float* gamma, int rows, int C, float* dx,
                           float* dgamma, float* dbeta, float* workspace);

/* Backward of y = act(x W^T + bias) (nn.Linear [+ LeakyReLU | Sigmoid | GELU]: the gates models/model_clipspm.py:88-99, the
 * FeedForward of token_trans :371-378, the temporal convolutions :169-172 as im2col GEMMs; forward = spm_gemm):
 * x [M,K], W [N,K], y / dy [M,N] fp32 contiguous -> dx [M,K] (null: skipped), dW [N,K], db [N] (null: skipped), all
 * overwritten.  y = the layer's OUTPUT (needed for LeakyReLU / sigmoid; GELU recomputes its pre-activation, bias may be
 * null).  act codes as spm_gemm.  workspace: spm_linear_backward_workspace(M, N, K) floats, 16-byte aligned. */
long long spm_linear_backward_workspace(int M, int N, int K);
int spm_linear_backward(void* stream, int precision, const float* x, const float* W, const float* bias, const float* y,
                        const float* dy, int M, int N, int K, int act, float slope, float* dx, float* dW, float* db,
                        float* workspace, long long workspace_floats);

/* Optimiser half of the training step (run/main_run.py:84-88 torch.optim.Adam(betas=(0.5, 0.999), weight_decay); :76,207-209
 * GradScaler.step / .update): multi-tensor kernels over all parameters, no host synchronisation.
 *   spm_adam_create   n fp32 device tensors (the model's parameters); allocates exp_avg / exp_avg_sq (zero) and the step count
 *   spm_adam_step     one optimiser step on grads[i] (device pointers, null = parameter without gradient).  scaler_state =
 *                     device float[3] {scale, growth_tracker, found_inf} of a GradScaler, or null: with it the gradients are
 *                     first unscaled IN PLACE (g *= 1/scale) and checked, and the whole step is skipped when any is inf / nan
 *   spm_scaler_update torch's _amp_update_scale_ on that state (backoff on overflow, growth after `growth_interval` clean
 *                     steps) and found_inf = 0
 *   spm_adam_state    device pointers of tensor i's exp_avg / exp_avg_sq and of the fp32 step count (state_dict) */
typedef struct spm_adam spm_adam;
int spm_adam_create(int n_tensors, float* const* params, const long long* numel, spm_adam** out);
int spm_adam_destroy(spm_adam* a);
int spm_adam_step(spm_adam* a, void* stream, float* const* grads, double lr, double beta1, double beta2, double eps,
                  double weight_decay, float* scaler_state);   /* the hyper-parameters as the Python floats they are */
int spm_adam_state(spm_adam* a, int i, float** exp_avg, float** exp_avg_sq, float** step);
/* SOLVER.OPTIM_METHOD == "sgd" (run/main_run.py:92-96 torch.optim.SGD(lr, momentum, weight_decay)) on the same handle: its first
 * state tensor is the momentum buffer (spm_adam_state's exp_avg); same scaler_state contract as spm_adam_step */
int spm_sgd_step(spm_adam* a, void* stream, float* const* grads, double lr, double momentum, double weight_decay,
                 float* scaler_state);
int spm_scaler_update(void* stream, float* scaler_state, float growth_factor, float backoff_factor, int growth_interval);

/* Frame-encoder self-attention stage (models/clip_fsar.py:626,638): qkv [F*197, 2304] bf16 (q | k | v, head h at
 * columns h*64 of each third) -> out [F*197, 768] bf16.  use_mma_sync = 0: tcgen05/TMEM kernel (product path),
 * 1: the mma.sync kernel kept as a cross-check. */
int spm_vit_attention(void* stream, const void* qkv, void* out, int n_frames, int use_mma_sync);

/* out[orow(m), n] = act(sum_k A[m,k] B[n,k] + bias[n]) (+ residual[rrow(m), n]); see csrc/gemm.cuh.
 * kind: 0 = bf16 operands, 1 = tf32 (fp32 operands).  act: 0 none, 1 QuickGELU, 2 GELU(erf), 3 LeakyReLU, 4 sigmoid, 5 ReLU */
int spm_gemm(void* stream, int kind, const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
             const float* bias, int act, float slope, const float* residual, int ldr, int res_row_mod,
             int res_row_off, int out_row_group, int out_group_stride, int out_row_off, void* out, int ldo,
             int out_bf16);

/* ---- text-prompt tower (class names -> text_features): models/model_clipspm.py:45-70, clip_fsar.py:793-805 ----
 * What the reference's CNN.__init__ computes once per class list.  Tokenisation (byte-level BPE) stays on the host
 * (clip_spm_b200/tokenizer.py mirrors clip_fsar.py:144-180,322-392); the device side starts from token ids. */
typedef struct spm_text spm_text;

/* embed_dim: 512 (ViT-B/16 checkpoint) or 1024 (RN50); precision as in spm_config */
int spm_text_create(int embed_dim, int precision, spm_text** out);
int spm_text_destroy(spm_text* h);
/* fp32 device tensors named by the reference's CLIP state_dict keys: token_embedding.weight, positional_embedding,
 * transformer.resblocks.{i}.{attn.in_proj_weight, attn.in_proj_bias, attn.out_proj.*, ln_1.*, ln_2.*, mlp.c_fc.*,
 * mlp.c_proj.*}, ln_final.*, text_projection; other names are ignored, a missing or mis-sized one is an error */
int spm_text_load_weights(spm_text* h, void* stream, int n, const char* const* names, const void* const* dev_ptrs,
                          const int64_t* numel);
/* CLIP.encode_text: tokens [n_texts, 77] int32 (device; <sot> ids <eot> 0...) -> out [n_texts, embed_dim] fp32 */
int spm_text_encode(spm_text* h, void* stream, const int32_t* tokens, int n_texts, float* out);
/* the constructor's loop: tokens [n_templates, n_classes, 77] -> out[c] = mean_t encode_text(tokens[t, c]) */
int spm_text_class_features(spm_text* h, void* stream, const int32_t* tokens, int n_templates, int n_classes,
                            float* out);

/* ---- evaluation-time frame transform (video_reader.py:83-111,265-272) ----
 * Resize(256) [PIL BILINEAR with antialiasing, videotransforms/functional.py:24-73] -> CenterCrop(224)
 * [video_transforms.py:204-247] -> ToTensor [uint8 HWC -> fp32 CHW / 255], bit-exact.
 * frames: device uint8 [n_frames, H, W, 3] (decoded RGB, all of one size) -> images_out fp32 [n_frames, 3, 224, 224] */
int spm_transform_frames(void* stream, const uint8_t* frames, int n_frames, int H, int W, float* images_out);
/* The TRAINING transform of the loader (video_reader.py:83-103: Resize(256) -> RandomHorizontalFlip [not for ssv2] ->
 * RandomCrop(224) -> ToTensor), bit-exact given the clip's random draws: aug = device int32 [n_frames, 3] {crop y1, crop x1,
 * flip} per frame in the coordinates of the RESIZED (and, when flip != 0, already mirrored) frame -- the numbers
 * videotransforms/video_transforms.py:152-153 draws (clip_spm_b200.frames.train_augmentation mirrors the draws and their
 * order); origins are clamped to the resized frame. */
int spm_transform_frames_train(void* stream, const uint8_t* frames, int n_frames, int H, int W, const int32_t* aug,
                               float* images_out);
/* spm_encode_frames on decoded frames: transform + encoder (on the bf16 ViT path the transform kernel writes the
 * patch-embedding GEMM's bf16 operand directly; the fp32 image is never materialised) */
int spm_encode_frames_u8(spm_handle* h, void* stream, const uint8_t* frames, int n_frames, int H, int W,
                         float* feats_out);
/* spm_eval_host on decoded frames in HOST memory: support frames uint8 [E, S*T, img_h, img_w, 3], target frames
 * [E, Q*T, img_h, img_w, 3]; 4x fewer (or less) host->device bytes than fp32 images, same results bit for bit */
int spm_eval_host_u8(spm_handle* h, int n_episodes, int S, int Q, int W, int img_h, int img_w,
                     const uint8_t* support_frames_host, const uint8_t* target_frames_host,
                     const float* support_labels_host, const float* real_support_host, const float* real_target_host,
                     const int64_t* target_labels_host, float tasks_per_batch, float* logits_host, float* dists_host,
                     float* loss_host, float* acc_host, int32_t* pred_host);
/* spm_eval on decoded frames already in DEVICE memory (e.g. the output of spm_jpeg_decode): support frames uint8
 * [E, S*T, img_h, img_w, 3], target frames [E, Q*T, img_h, img_w, 3]; everything else as spm_eval */
int spm_eval_u8(spm_handle* h, void* stream, int n_episodes, int S, int Q, int W, int img_h, int img_w,
                const uint8_t* support_frames, const uint8_t* target_frames, const float* support_labels,
                const float* real_support, const float* real_target, const int64_t* target_labels, float tasks_per_batch,
                float* logits_out, float* dists_out, float* loss_out, float* acc_out, int32_t* pred_out);

/* ---- JPEG decode (video_reader.py:227-230 read_single_image: PIL Image.open(path).load(), one file per frame) ----
 * Baseline / extended-sequential Huffman JPEG, 8-bit YCbCr, 4:4:4 / 4:2:2 / 4:2:0, one interleaved scan, restart
 * intervals allowed (what ffmpeg-extracted frame dumps contain); progressive / arithmetic / CMYK files are rejected
 * with an error.  The decoded RGB bytes are identical to PIL's (libjpeg-turbo: ISLOW IDCT, fancy upsampling).
 * spm_jpeg_info: header of one file in HOST memory -> size and luma sampling factors (any output pointer may be null).
 * spm_jpeg_decode: n_images files in HOST memory (all H x W with the same chroma subsampling) -> frames_out, DEVICE
 * uint8 [n_images, H, W, 3].  Container parsing and byte un-stuffing run on host threads, entropy decoding, IDCT,
 * upsampling and colour conversion on the GPU; the call returns after the stream has consumed the host buffers. */
int spm_jpeg_info(const uint8_t* jpeg_host, long long n_bytes, int* height, int* width, int* h_samp, int* v_samp);
int spm_jpeg_decode(void* stream, int n_images, const uint8_t* const* jpeg_host, const int64_t* jpeg_bytes, int H, int W,
                    uint8_t* frames_out);
/* the geometry that transform uses for an H x W frame (host arithmetic only; any output pointer may be null) */
int spm_frame_geometry(int H, int W, int* resized_h, int* resized_w, int* crop_y, int* crop_x);

#ifdef __cplusplus
}
#endif
#endif /* CLIPSPM_B200_H */
