// Internal interface of the model translation units (model.cu: C ABI + forward / host pipelines; model_vit.cu: frame
// encoder; model_head.cu: metric heads): the handle, the packed-weight and plan structures, error / allocation helpers
// and the few functions that cross the files.  Not part of the public ABI (include/clipspm_b200.h).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/clipspm_b200.h"
#include "api_common.cuh"
#include "gemm.cuh"
#include "head_kernels.cuh"
#include "kernels.cuh"
#include "profile.cuh"
#include "rn50.cuh"

namespace spm {

int device_sm_count(int* out);

namespace detail {
constexpr int VIT_C = 768, VIT_L = 197, VIT_P = 196, VIT_LAYERS = 12, VIT_OUT = 512;
constexpr int HEAD_INNER = 2048, HEAD_HEADS = 8, HEAD_DH = 256, HEAD_MLP = 2048;
constexpr long long FRAME_ELEMS = 3LL * 224 * 224;

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
};

#define SPM_KERNEL(call)                                                                       \
  do {                                                                                         \
    int _r = (call);                                                                           \
    if (_r != 0) {                                                                             \
      set_error(std::string(#call) + (_r < 0 ? ": unsupported shape"                          \
                                              : std::string(": ") + cudaGetErrorString((cudaError_t)_r))); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

#define SPM_GEMM_RUN(op)                                 \
  do {                                                   \
    const char* _e = "";                                 \
    if (gemm_run(&(op), st, &_e)) {                      \
      set_error(std::string("gemm_run " #op ": ") + _e); \
      return 1;                                          \
    }                                                    \
  } while (0)

struct VitLayerW {
  __nv_bfloat16 *qkv_w, *out_w, *fc_w, *proj_w;
  float *qkv_b, *out_b, *fc_b, *proj_b, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  // ln_1 folded into the QKV projection and ln_2 into c_fc (k_fold_ln): weights * gamma, column sums, bias + W beta
  __nv_bfloat16 *qkv_wf = nullptr, *fc_wf = nullptr;
  float *qkv_c = nullptr, *qkv_bf = nullptr, *fc_c = nullptr, *fc_bf = nullptr;
};
struct VitW {
  __nv_bfloat16* conv1_w = nullptr;  // [768, 768]
  float *cls_pos = nullptr, *pos = nullptr, *ln_pre_g = nullptr, *ln_pre_b = nullptr, *ln_post_g = nullptr,
        *ln_post_b = nullptr;
  __nv_bfloat16* projT = nullptr;  // [512, 768]
  VitLayerW layer[VIT_LAYERS];
};
// fp32 copies for the SPM_PRECISION_FP32 parity mode (same layouts, no bf16 rounding anywhere)
struct VitLayerW32 {
  float *qkv_w, *out_w, *fc_w, *proj_w;
};
struct VitW32 {
  float* conv1_w = nullptr;  // [768, 768]
  float* projT = nullptr;    // [512, 768]
  VitLayerW32 layer[VIT_LAYERS];
};
struct CtxW {
  float *ln_g, *ln_b, *qkv_w, *out_w, *out_b, *ff0_w, *ff0_b, *ff3_w, *ff3_b;
};
struct HeadW {
  float *mc1_w, *mc1_b, *mc2_w, *mc2_b;
  float *tt0_w, *tt0_b, *tt3_w, *tt3_b;
  float *gt0_w, *gt0_b, *gt2_w, *gt2_b, *gv0_w, *gv0_b, *gv2_w, *gv2_b;
  CtxW ctx[2];  // [0] = context1 (PADM), [1] = context2 (SPM se_te)
  float* mo_alpha1;
};

struct VitPlan {
  GemmOp patch, qkv[VIT_LAYERS], outp[VIT_LAYERS], fc[VIT_LAYERS], proj[VIT_LAYERS], fin;
  // last block restricted to the class-token rows (the only rows ln_post reads, clip_fsar.py:684)
  GemmOp outp_cls, fc_cls, proj_cls;
  bool ln_fold = false;   // qkv / fc normalise in their epilogues, outp / proj emit the bf16 rows + statistics: no LayerNorm kernels
};
struct CtxPlan {
  GemmOp qkv, outp, ff0, ff3;
};
struct Cpm2cPlan {   // sibling head CPM2C: motion fusion GEMMs, gates and one context2 pass per branch (motion, normal)
  int E, S, Q;
  const float* X;
  GemmOp f1, f3, f5, sc, gt0, gt2;
  GemmOp gv0[2], gv2[2];
  CtxPlan c2[2];
};
struct FsarPlan {   // sibling head CLIP-FSAR: context2 (one plan per layer) over E*S*(T+1) + E*Q*T rows (E*W with MERGE_BEFORE)
  int E, S, Q, W;
  std::vector<CtxPlan> c2;
};
struct HeadPlan {
  int E, S, Q, W;
  const float* X;  // frame-feature base the plan's tensor maps point at
  GemmOp mc1, mc2, tt0, tt3, gt0, gt2, gv0, gv2;
  CtxPlan c2, c1;
};
}  // namespace detail
using namespace detail;
}  // namespace spm

struct spm_handle {
  spm_config cfg;
  int D = 512, HT = 768, HV = 256;
  int sms = 148;
  int frame_chunk = 512;
  int alt_dir = 1;  // SPM_ALT_DIR=0: every kernel sweeps its rows in ascending order
  bool prune_last = true;  // SPM_PRUNE_LAST=0 runs the last block on all tokens (same result, more work)
  bool attn_mma = false;  // SPM_ATTN=mma selects the mma.sync attention kernel instead of the tcgen05 one
  bool weights_loaded = false, text_set = false;
  std::vector<void*> allocs;
  spm::VitW vit;
  spm::VitW32 vit32;
  bool fp32 = false;  // SPM_PRECISION_FP32: CUDA-core fp32 GEMMs / attention, fp32 activations
  bool resid_bf16 = false;  // SPM_PRECISION_BF16_RESID: bf16 residual stream (xb / xcb below)
  __nv_bfloat16 *xb = nullptr, *xcb = nullptr;
  float *patches32 = nullptr, *xn32 = nullptr, *qkv32 = nullptr, *attn32 = nullptr, *hid32 = nullptr, *cls32 = nullptr;
  spm::Rn50* rn50 = nullptr;
  spm::HeadW head;
  float* text = nullptr;
  int n_cls = 0;
  // encoder workspace (sized for frame_chunk frames)
  __nv_bfloat16 *patches = nullptr, *xn = nullptr, *qkv = nullptr, *attn = nullptr, *hid = nullptr, *cls = nullptr;
  float* x = nullptr;
  float* xc = nullptr;          // [frame_chunk, 768] class-token rows of the residual stream in the last block
  __nv_bfloat16* xnc = nullptr; // their LayerNorm output
  float* feats = nullptr;  // [max frames per call, D]
  long long feats_cap = 0;
  // Opt-in schedule (SPM_ENC_STREAMS=2): two encoder workspaces, consecutive frame chunks alternate between two streams
  // and the heads of episode groups run on a third, so that ramp-up / tail / memory-bound kernels of one chunk overlap
  // the other's GEMMs.  Bit-identical results (tests), but measured NOT faster: the step sits at the 1000 W power cap,
  // where overlap buys nothing, and multi-stream runs showed sporadic 100-300 ms submission stalls.  Default: 1 stream.
  struct VitWs {
    __nv_bfloat16 *patches, *xn, *qkv, *attn, *hid, *cls, *xnc, *xb, *xcb;
    float *x, *xc, *ln_stats;
  } vit_ws[2] = {};
  int ln_fold = 1;          // 2 = centred weights (no column-sum term); SPM_LN_FOLD=0: separate LayerNorm kernels everywhere (bf16 precision, fp32 residual stream only)
  float* ln_stats = nullptr;
  int cur_ws = 0, enc_streams = 1;  // 2 = opt-in (SPM_ENC_STREAMS): measured no faster under the power cap
  cudaStream_t enc_stream[2] = {nullptr, nullptr};
  cudaEvent_t enc_fork = nullptr, enc_join[2] = {nullptr, nullptr};
  float* img_scratch = nullptr;  // fp32 images of uint8 input frames (fp32-mode ViT and RN50 paths)
  long long img_scratch_cap = 0;
  std::map<int, std::unique_ptr<spm::VitPlan>> vit_plans;
  // head workspace
  long long head_cap_E = 0, head_cap_S = 0, head_cap_Q = 0, head_cap_W = 0;
  float *X = nullptr, *XC = nullptr, *C1 = nullptr, *C2 = nullptr, *TOK = nullptr, *TTIN = nullptr, *TTH = nullptr,
        *GTH = nullptr, *GT = nullptr, *GVH = nullptr, *GV = nullptr, *SEQ = nullptr, *HN = nullptr, *QKVH = nullptr,
        *AO = nullptr, *Y = nullptr, *FFH = nullptr, *Z = nullptr, *Z1 = nullptr, *NEWM = nullptr, *SUPRO = nullptr,
        *SUPRO2 = nullptr, *ACC = nullptr, *D3 = nullptr;
  int* err_flag = nullptr;
  int last_E = 0, last_S = 0, last_Q = 0, last_W = 0;   // shape of the most recent CLIP-SPM head pass (spm_head_stage)
  std::vector<std::unique_ptr<spm::HeadPlan>> head_plans;
  // sibling head CLIP-FSAR (cfg.head == SPM_HEAD_CLIPFSAR; models/model_clipfsar.py)
  spm::CtxW fsar_ctx = {};                 // context2.layers.0 (shared with the CPM2C head)
  std::vector<spm::CtxW> fsar_ctx_more;    // context2.layers.1.. (cfg.fsar_depth > 1, model_clipfsar.py:143-144)
  float* fsar_scale = nullptr;
  float* text_train = nullptr;   // [n_cls_train, D] text_features_train (class_text_logits)
  int n_cls_train = 0;
  float* CLS = nullptr;          // [E, S+Q, n_cls_train] class_text_logits of the last head call
  long long cls_cap = 0, cls_rows = 0;
  std::vector<std::unique_ptr<spm::FsarPlan>> fsar_plans;
  // sibling head CPM2C (cfg.head == SPM_HEAD_CPM2C; models/model_cpm2c.py): context2 / scale share fsar_ctx / fsar_scale,
  // the gates share head.g*; motion fusion weights, class tokens and its own workspace below
  struct Cpm2cW {
    float *m1_w = nullptr, *m1_b = nullptr, *m3_w = nullptr, *m3_b = nullptr, *m5_w = nullptr, *m5_b = nullptr,
          *sc_w = nullptr, *sc_b = nullptr, *cls_tok = nullptr, *cls_tok_motion = nullptr;
  } cpm;
  float *CP_FCAT = nullptr, *CP_CONV = nullptr, *CP_MOT = nullptr, *CP_TOK = nullptr, *CP_GT = nullptr, *CP_PRO = nullptr,
        *CP_LOC = nullptr, *CP_GLOB = nullptr, *CP_OUT_L = nullptr, *CP_OUT_G = nullptr;
  long long cp_cap_V = 0, cp_cap_EQW = 0, cp_cap_EW = 0;
  int cp_last_E = 0, cp_last_Q = 0, cp_last_W = 0;
  std::vector<std::unique_ptr<spm::Cpm2cPlan>> cpm2c_plans;
  // `X` is the feature block the head currently reads: its own buffer (Xhead), or a group of episodes inside Xall
  // when the forward pipelines episode groups (encoder of group g+1 overlaps the head of group g on head_stream)
  float *Xhead = nullptr, *Xall = nullptr;
  long long xall_cap = 0, tmp_logits_cap = 0, tmp_dists_cap = 0;
  cudaStream_t head_stream = nullptr;
  cudaEvent_t head_done = nullptr;
  std::vector<cudaEvent_t> chunk_ev;
  // forward workspace: logits/dists when the caller only wants loss/acc, host staging for spm_eval_host
  float *tmp_logits = nullptr, *tmp_dists = nullptr;
  struct Stage {
    uint8_t *su = nullptr, *qu = nullptr;  // staged input frames (fp32 images or uint8 frames), byte-addressed
    float *lab = nullptr, *rs = nullptr, *rt = nullptr;
    long long* tl = nullptr;
    float *logits = nullptr, *dists = nullptr, *loss = nullptr, *acc = nullptr;
    int* pred = nullptr;
    cudaEvent_t copied = nullptr, done = nullptr;
  } stage[2];
  long long stage_cap_bytes_s = 0, stage_cap_bytes_q = 0;            // image staging rings
  long long stage_cap_S = 0, stage_cap_Q = 0, stage_cap_QW = 0, stage_cap_R = 0;   // label / result rings (R slots each)
  cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
  std::vector<cudaEvent_t> ev_copied, ev_done;  // per chunk of one spm_eval_host call
  // pinned host landing zone for the results: an async D2H into the caller's (possibly pageable) buffers would
  // block the enqueueing thread until the chunk has finished and starve the GPU of the next chunk's launches
  float* pin_res = nullptr;
  long long pin_cap = 0;
  // spm_eval_host_set_next: the first chunk of the NEXT spm_eval_host call is copied to these buffers behind the
  // current call's own copies, so that call starts computing at once (its one exposed H2D copy disappears)
  const void *next_su = nullptr, *next_qu = nullptr;   // hint given by the caller, consumed by the next call
  int next_n = 0;                                      // episodes the hinted call will evaluate
  uint8_t *pf_su = nullptr, *pf_qu = nullptr;
  // second set: the next call's first chunk is copied into the buffers the CURRENT call does not read, so that the copy
  // need not wait for this call's compute (a one-chunk call -- one episode per call -- would otherwise overlap nothing)
  uint8_t *pf_su_alt = nullptr, *pf_qu_alt = nullptr;
  long long pf_cap_s = 0, pf_cap_q = 0;
  const void *pf_src_su = nullptr, *pf_src_qu = nullptr;   // what the buffers hold (null = nothing)
  long long pf_bytes_s = 0, pf_bytes_q = 0;
  cudaEvent_t pf_event = nullptr;
};

namespace spm {
namespace detail {

inline int dalloc(spm_handle* h, void** p, size_t bytes) {
  SPM_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  h->allocs.push_back(*p);
  return 0;
}
template <class T>
int dalloc_t(spm_handle* h, T** p, long long n) {
  return dalloc(h, reinterpret_cast<void**>(p), (size_t)n * sizeof(T));
}
// Regrow: the superseded buffer is freed (after a device sync -- work enqueued earlier may still read it) instead of
// staying in `allocs` until spm_destroy, so a sweep over varying episode shapes does not accumulate dead workspace.
template <class T>
int drealloc_t(spm_handle* h, T** p, long long n) {
  if (*p != nullptr) {
    SPM_CUDA(cudaDeviceSynchronize());
    auto it = std::find(h->allocs.begin(), h->allocs.end(), static_cast<void*>(*p));
    if (it != h->allocs.end()) h->allocs.erase(it);
    SPM_CUDA(cudaFree(*p));
    *p = nullptr;
  }
  return dalloc_t(h, p, n);
}

// ---------------------------------------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------------------------------------
struct WeightTable {
  std::unordered_map<std::string, std::pair<const float*, long long>> m;
  int get(const std::string& name, long long numel, const float** out) const {
    auto it = m.find(name);
    if (it == m.end()) { set_error("spm_load_weights: missing tensor '" + name + "'"); return 1; }
    if (it->second.second != numel) {
      set_error("spm_load_weights: tensor '" + name + "' has " + std::to_string(it->second.second) +
                " elements, expected " + std::to_string(numel));
      return 1;
    }
    *out = it->second.first;
    return 0;
  }
};

inline int copy_f32(spm_handle* h, cudaStream_t st, const WeightTable& wt, const std::string& name, long long n, float** dst) {
  const float* src;
  SPM_TRY(wt.get(name, n, &src));
  SPM_TRY(dalloc_t(h, dst, n));
  SPM_CUDA(cudaMemcpyAsync(*dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}
inline int copy_bf16(spm_handle* h, cudaStream_t st, const WeightTable& wt, const std::string& name, long long n,
              __nv_bfloat16** dst) {
  const float* src;
  SPM_TRY(wt.get(name, n, &src));
  SPM_TRY(dalloc_t(h, dst, n));
  SPM_KERNEL(k_cast_bf16(st, src, *dst, n));
  return 0;
}

inline int plan_gemm(GemmOp* op, int kind, const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
              const GemmEpilogue& ep, int sms) {
  const char* err = "";
  if (gemm_plan(op, kind, A, lda, B, ldb, M, N, K, ep, sms, &err)) {
    set_error(std::string("gemm_plan: ") + err);
    return 1;
  }
  return 0;
}

// A run of frames: fp32 images [n,3,224,224], or (frames_u8 != null) decoded RGB uint8 frames [n,H,W,3] that go
// through the Resize/CenterCrop/ToTensor kernel first (frame_transform.cu)
struct Segment {
  const float* images;
  long long n_frames;
  const uint8_t* frames_u8 = nullptr;
  int H = 0, W = 0;
};

// after_chunk(frames_done, chunk_no, chunk_stream) is called once the kernels of a chunk have been enqueued
using ChunkHook = std::function<int(long long, int, cudaStream_t)>;

// ---- model_vit.cu
int load_vit(spm_handle* h, cudaStream_t st, const WeightTable& wt);
int load_vit32(spm_handle* h, cudaStream_t st, const WeightTable& wt);
// Encode the concatenation of the segments; feature rows come out in segment order.
int encode_segments(spm_handle* h, cudaStream_t st, const Segment* segs, int nseg, float* feats_out,
                    const ChunkHook* after_chunk = nullptr);
// ---- model_head.cu
int load_head(spm_handle* h, cudaStream_t st, const WeightTable& wt);
int load_head_fsar(spm_handle* h, cudaStream_t st, const WeightTable& wt);
int load_head_cpm2c(spm_handle* h, cudaStream_t st, const WeightTable& wt);
int ensure_head_workspace(spm_handle* h, int E, int S, int Q, int W);
// Frame features already in h->X as [E, N, T, D] (supports first).  Produces logits [E,Q,W], dists [E] and, when
// target_labels is given, loss / accuracy / predictions (the head is chosen by cfg.head).
int head_run(spm_handle* h, cudaStream_t st, int E, int S, int Q, int W, const float* labels, const float* real_s,
             const float* real_t, const long long* target_labels, float tasks_per_batch, float* logits, float* dists,
             float* loss, float* acc, int* pred);
int reset_err_flag(spm_handle* h, cudaStream_t st);
int head_stage(spm_handle* h, cudaStream_t st, const char* name, float* out, long long capacity, long long* numel);
int check_shapes(spm_handle* h, int E, int S, int Q, int W);

}  // namespace detail
}  // namespace spm
