// Handle, weight packing, workspace and the stream-ordered pipelines behind the C ABI:
//   frame encoder (ViT-B/16)  models/clip_fsar.py:672-689
//   metric head               models/model_clipspm.py:116-143 (mo / sem / taskM / otam_distance / logits)
//   loss + accuracy           utils/utils.py:174-186,259-264 ; run/main_run.py:390-392
// No allocation happens inside a forward call once the shapes have been seen (plans and workspace are cached),
// and nothing synchronises with the host except spm_load_weights and spm_eval_host (whose contract is blocking).
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

namespace {
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
};
struct CtxPlan {
  GemmOp qkv, outp, ff0, ff3;
};
struct FsarPlan {   // sibling head CLIP-FSAR: one context2 pass over E*S*(T+1) + E*Q*T rows
  int E, S, Q;
  CtxPlan c2;
};
struct HeadPlan {
  int E, S, Q, W;
  const float* X;  // frame-feature base the plan's tensor maps point at
  GemmOp mc1, mc2, tt0, tt3, gt0, gt2, gv0, gv2;
  CtxPlan c2, c1;
};
}  // namespace
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
    __nv_bfloat16 *patches, *xn, *qkv, *attn, *hid, *cls, *xnc;
    float *x, *xc;
  } vit_ws[2] = {};
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
  std::vector<std::unique_ptr<spm::HeadPlan>> head_plans;
  // sibling head CLIP-FSAR (cfg.head == SPM_HEAD_CLIPFSAR; models/model_clipfsar.py)
  spm::CtxW fsar_ctx = {};
  float* fsar_scale = nullptr;
  float* text_train = nullptr;   // [n_cls_train, D] text_features_train (class_text_logits)
  int n_cls_train = 0;
  float* CLS = nullptr;          // [E, S+Q, n_cls_train] class_text_logits of the last head call
  long long cls_cap = 0, cls_rows = 0;
  std::vector<std::unique_ptr<spm::FsarPlan>> fsar_plans;
  // `X` is the feature block the head currently reads: its own buffer (Xhead), or a group of episodes inside Xall
  // when the forward pipelines episode groups (encoder of group g+1 overlaps the head of group g on head_stream)
  float *Xhead = nullptr, *Xall = nullptr;
  long long xall_cap = 0, tmp_out_cap = 0;
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
  long long stage_cap_frames_s = 0, stage_cap_frames_q = 0, stage_cap_bytes_s = 0, stage_cap_bytes_q = 0;
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
  long long pf_cap_s = 0, pf_cap_q = 0;
  const void *pf_src_su = nullptr, *pf_src_qu = nullptr;   // what the buffers hold (null = nothing)
  long long pf_bytes_s = 0, pf_bytes_q = 0;
  cudaEvent_t pf_event = nullptr;
};

namespace spm {
namespace {

int dalloc(spm_handle* h, void** p, size_t bytes) {
  SPM_CUDA(cudaMalloc(p, bytes ? bytes : 16));
  h->allocs.push_back(*p);
  return 0;
}
template <class T>
int dalloc_t(spm_handle* h, T** p, long long n) {
  return dalloc(h, reinterpret_cast<void**>(p), (size_t)n * sizeof(T));
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

int copy_f32(spm_handle* h, cudaStream_t st, const WeightTable& wt, const std::string& name, long long n, float** dst) {
  const float* src;
  SPM_TRY(wt.get(name, n, &src));
  SPM_TRY(dalloc_t(h, dst, n));
  SPM_CUDA(cudaMemcpyAsync(*dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}
int copy_bf16(spm_handle* h, cudaStream_t st, const WeightTable& wt, const std::string& name, long long n,
              __nv_bfloat16** dst) {
  const float* src;
  SPM_TRY(wt.get(name, n, &src));
  SPM_TRY(dalloc_t(h, dst, n));
  SPM_KERNEL(k_cast_bf16(st, src, *dst, n));
  return 0;
}

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
  return load_ctx(h, st, wt, "context2.layers.0.", h->D, &h->fsar_ctx);
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
// ViT-B/16 frame encoder
// ---------------------------------------------------------------------------------------------------------
int plan_gemm(GemmOp* op, int kind, const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
              const GemmEpilogue& ep, int sms) {
  const char* err = "";
  if (gemm_plan(op, kind, A, lda, B, ldb, M, N, K, ep, sms, &err)) {
    set_error(std::string("gemm_plan: ") + err);
    return 1;
  }
  return 0;
}

void select_vit_ws(spm_handle* h, int i) {
  const spm_handle::VitWs& w = h->vit_ws[i];
  h->patches = w.patches; h->xn = w.xn; h->qkv = w.qkv; h->attn = w.attn; h->hid = w.hid; h->cls = w.cls;
  h->xnc = w.xnc; h->x = w.x; h->xc = w.xc;
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
    SPM_TRY(plan_gemm(&pl->outp[i], kind, a_attn, C, f32 ? (const void*)l32.out_w : (const void*)l.out_w, C, M, C, C, e2, h->sms));
    GemmEpilogue e3;
    e3.bias = l.fc_b; e3.act = ACT_QUICKGELU; e3.out = o_hid; e3.ldo = 4 * C; e3.out_bf16 = obf;
    SPM_TRY(plan_gemm(&pl->fc[i], kind, a_xn, C, f32 ? (const void*)l32.fc_w : (const void*)l.fc_w, C, M, 4 * C, C, e3, h->sms));
    GemmEpilogue e4;
    e4.bias = l.proj_b; e4.residual = h->x; e4.ldr = C; e4.out = h->x; e4.ldo = C;
    SPM_TRY(plan_gemm(&pl->proj[i], kind, a_hid, 4 * C, f32 ? (const void*)l32.proj_w : (const void*)l.proj_w, 4 * C, M, C, 4 * C, e4, h->sms));
  }
  if (!f32) {
    // Last block, class-token rows only: attention output / residual rows are taken with a row stride of 197 tokens
    const VitLayerW& l = v.layer[VIT_LAYERS - 1];
    const long long LC = (long long)VIT_L * C;
    GemmEpilogue e2;
    e2.bias = l.out_b; e2.residual = h->x; e2.ldr = (int)LC; e2.out = h->xc; e2.ldo = C;
    SPM_TRY(plan_gemm(&pl->outp_cls, GEMM_BF16, h->attn, LC, l.out_w, C, F, C, C, e2, h->sms));
    GemmEpilogue e3;
    e3.bias = l.fc_b; e3.act = ACT_QUICKGELU; e3.out = h->hid; e3.ldo = 4 * C; e3.out_bf16 = 1;
    SPM_TRY(plan_gemm(&pl->fc_cls, GEMM_BF16, h->xnc, C, l.fc_w, C, F, 4 * C, C, e3, h->sms));
    GemmEpilogue e4;
    e4.bias = l.proj_b; e4.residual = h->xc; e4.ldr = C; e4.out = h->xc; e4.ldo = C;
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
  SPM_KERNEL(k_layernorm(st, h->x, C, M, C, v.ln_pre_g, v.ln_pre_b, v.cls_pos, VIT_L, h->x, nullptr, C, next_dir()));
  for (int i = 0; i < VIT_LAYERS; ++i) {
    const VitLayerW& l = v.layer[i];
    SPM_KERNEL(k_layernorm(st, h->x, C, M, C, l.ln1_g, l.ln1_b, nullptr, 0, nullptr, h->xn, C, next_dir()));
    SPM_GEMM_RUN_DIR(pl->qkv[i]);
    if (h->attn_mma)
      SPM_KERNEL(k_vit_attention(st, h->qkv, h->attn, F));
    else
      SPM_KERNEL(k_vit_attention_tc(st, h->qkv, h->attn, F, h->sms, next_dir()));
    if (i == VIT_LAYERS - 1 && h->prune_last) {
      // only x[:, 0, :] is read after the last block: run its out-proj / MLP on the F class-token rows
      SPM_GEMM_RUN(pl->outp_cls);
      SPM_KERNEL(k_layernorm(st, h->xc, C, F, C, l.ln2_g, l.ln2_b, nullptr, 0, nullptr, h->xnc, C));
      SPM_GEMM_RUN(pl->fc_cls);
      SPM_GEMM_RUN(pl->proj_cls);
      SPM_KERNEL(k_layernorm(st, h->xc, C, F, C, v.ln_post_g, v.ln_post_b, nullptr, 0, nullptr, h->cls, C));
      GemmOp fin = pl->fin;
      fin.ep.out = feats_out;
      SPM_GEMM_RUN(fin);
      return 0;
    }
    SPM_GEMM_RUN_DIR(pl->outp[i]);
    SPM_KERNEL(k_layernorm(st, h->x, C, M, C, l.ln2_g, l.ln2_b, nullptr, 0, nullptr, h->xn, C, next_dir()));
    SPM_GEMM_RUN_DIR(pl->fc[i]);
    SPM_GEMM_RUN_DIR(pl->proj[i]);
  }
#undef SPM_GEMM_RUN_DIR
  SPM_KERNEL(k_layernorm(st, h->x, (long long)VIT_L * C, F, C, v.ln_post_g, v.ln_post_b, nullptr, 0, nullptr, h->cls, C));
  GemmOp fin = pl->fin;
  fin.ep.out = feats_out;
  SPM_GEMM_RUN(fin);
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

// fp32 images of frames [a, b) of a segment: the caller's own, or transformed into the handle's scratch
int segment_images(spm_handle* h, cudaStream_t st, const Segment& seg, long long a, long long b, const float** out) {
  if (seg.frames_u8 == nullptr) { *out = seg.images + a * FRAME_ELEMS; return 0; }
  if (b - a > h->img_scratch_cap) {
    SPM_TRY(dalloc_t(h, &h->img_scratch, (b - a) * FRAME_ELEMS));
    h->img_scratch_cap = b - a;
  }
  SPM_KERNEL(k_frame_transform(st, seg.frames_u8 + a * (long long)seg.H * seg.W * 3, (int)(b - a), seg.H, seg.W,
                               h->img_scratch, nullptr));
  *out = h->img_scratch;
  return 0;
}

// Encode the concatenation of the segments; feature rows come out in segment order.
// after_chunk(frames_done, chunk_no, chunk_stream) is called once the kernels of a chunk have been enqueued
using ChunkHook = std::function<int(long long, int, cudaStream_t)>;
int encode_segments(spm_handle* h, cudaStream_t st, const Segment* segs, int nseg, float* feats_out,
                    const ChunkHook* after_chunk = nullptr) {
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

// ---------------------------------------------------------------------------------------------------------
// metric head
// ---------------------------------------------------------------------------------------------------------
int ensure_head_workspace(spm_handle* h, int E, int S, int Q, int W) {
  if (E <= h->head_cap_E && S <= h->head_cap_S && Q <= h->head_cap_Q && W <= h->head_cap_W) return 0;
  // grow-only: plans that point into the old buffers are dropped
  h->head_plans.clear();
  h->fsar_plans.clear();
  const long long cE = std::max<long long>(E, h->head_cap_E), cS = std::max<long long>(S, h->head_cap_S),
                  cQ = std::max<long long>(Q, h->head_cap_Q), cW = std::max<long long>(W, h->head_cap_W);
  const long long T = h->cfg.seq_len, D = h->D, N = cS + cQ, V = cE * N;
  const long long R2 = 2 * V * (T + 1), R1 = cE * T * (cW + cS + 1 + cQ), R = std::max(R1, R2);
  SPM_TRY(dalloc_t(h, &h->Xhead, V * T * D));
  h->X = h->Xhead;
  SPM_TRY(dalloc_t(h, &h->XC, V * T * 3 * D));
  SPM_TRY(dalloc_t(h, &h->C1, V * T * D));
  SPM_TRY(dalloc_t(h, &h->C2, V * T * D));
  SPM_TRY(dalloc_t(h, &h->TOK, 2 * V * D));
  SPM_TRY(dalloc_t(h, &h->TTIN, cE * cQ * D));
  SPM_TRY(dalloc_t(h, &h->TTH, cE * cQ * HEAD_MLP));
  SPM_TRY(dalloc_t(h, &h->GTH, 2 * V * h->HT));
  SPM_TRY(dalloc_t(h, &h->GT, 2 * V * D));
  SPM_TRY(dalloc_t(h, &h->GVH, V * T * h->HV));
  SPM_TRY(dalloc_t(h, &h->GV, V * T * D));
  SPM_TRY(dalloc_t(h, &h->SEQ, R * D));
  SPM_TRY(dalloc_t(h, &h->HN, R * D));
  SPM_TRY(dalloc_t(h, &h->QKVH, R * 3 * HEAD_INNER));
  SPM_TRY(dalloc_t(h, &h->AO, R * HEAD_INNER));
  SPM_TRY(dalloc_t(h, &h->Y, R * D));
  SPM_TRY(dalloc_t(h, &h->FFH, R * HEAD_MLP));
  SPM_TRY(dalloc_t(h, &h->Z, R2 * D));
  SPM_TRY(dalloc_t(h, &h->Z1, R1 * D));
  SPM_TRY(dalloc_t(h, &h->NEWM, V * D));
  SPM_TRY(dalloc_t(h, &h->SUPRO, cE * cW * T * D));
  SPM_TRY(dalloc_t(h, &h->SUPRO2, cE * cW * T * D));
  SPM_TRY(dalloc_t(h, &h->ACC, cE * cQ * cW));
  SPM_TRY(dalloc_t(h, &h->D3, cE * cW));
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
int fsar_head_run(spm_handle* h, cudaStream_t st, int E, int S, int Q, int W, const float* labels, const float* real_s,
                  const float* real_t, const long long* target_labels, float tasks_per_batch, float* logits,
                  float* dists, float* loss, float* acc, int* pred) {
  SPM_CHECK(h->text_set, "head: text features not set (spm_set_text_features)");
  SPM_TRY(ensure_head_workspace(h, E, S, Q, W));
  const int T = h->cfg.seq_len, D = h->D, N = S + Q, V = E * N, dh = D / HEAD_HEADS;
  const long long RS = (long long)E * S * (T + 1), R = RS + (long long)E * Q * T, TD = (long long)T * D;
  FsarPlan* pl = nullptr;
  for (auto& p : h->fsar_plans)
    if (p->E == E && p->S == S && p->Q == Q) pl = p.get();
  if (pl == nullptr) {
    std::unique_ptr<FsarPlan> np(new FsarPlan());
    np->E = E; np->S = S; np->Q = Q;
    SPM_TRY(plan_ctx(h, &np->c2, h->fsar_ctx, (int)R, h->SEQ, h->Z, D));
    pl = np.get();
    h->fsar_plans.push_back(std::move(np));
  }
  const CtxW& w = h->fsar_ctx;
  SPM_KERNEL(k_fsar_seq_build(st, h->X, h->text, h->n_cls, real_s, E, S, Q, T, D, h->SEQ));
  SPM_KERNEL(k_layernorm(st, h->SEQ, D, (int)R, D, w.ln_g, w.ln_b, nullptr, 0, h->HN, nullptr, D));
  SPM_GEMM_RUN(pl->c2.qkv);
  SPM_KERNEL(k_seq_attention(st, h->QKVH, h->AO, E * S, T + 1, 1, 0, T + 1, 0, 0, HEAD_HEADS, dh));
  SPM_KERNEL(k_seq_attention(st, h->QKVH + RS * 3 * D, h->AO + RS * D, E * Q, T, 1, 0, T, 0, 0, HEAD_HEADS, dh));
  SPM_GEMM_RUN(pl->c2.outp);
  SPM_GEMM_RUN(pl->c2.ff0);
  SPM_GEMM_RUN(pl->c2.ff3);
  SPM_KERNEL(k_fsar_class_mean(st, h->Z, labels, E, S, W, T, D, h->SUPRO, h->err_flag));
  SPM_KERNEL(k_otam(st, h->SUPRO, (long long)W * TD, TD, D, h->Z + RS * D, (long long)Q * TD, TD, D, E, W, Q, T, D,
                    h->cfg.single_direct, 1.f, 0.f, h->ACC));
  SPM_CUDA(cudaMemsetAsync(h->D3, 0, (size_t)E * W * sizeof(float), st));
  SPM_CUDA(cudaMemsetAsync(dists, 0, (size_t)E * sizeof(float), st));   // this head has no auxiliary distance
  h->cls_rows = 0;
  if (h->text_train != nullptr) {
    const long long need = (long long)V * h->n_cls_train;
    if (need > h->cls_cap) {
      SPM_TRY(dalloc_t(h, &h->CLS, need));
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

// Frame features already in h->X as [E, N, T, D] (supports first).  Produces logits [E,Q,W], dists [E] and, when
// target_labels is given, loss / accuracy / predictions.
int head_run(spm_handle* h, cudaStream_t st, int E, int S, int Q, int W, const float* labels, const float* real_s,
             const float* real_t, const long long* target_labels, float tasks_per_batch, float* logits, float* dists,
             float* loss, float* acc, int* pred) {
  if (h->cfg.head == SPM_HEAD_STEN)
    return sten_head_run(h, st, E, S, Q, W, labels, real_s, target_labels, tasks_per_batch, logits, dists, loss, acc, pred);
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
  SPM_KERNEL(k_token_prepare(st, h->text, real_s, real_t, h->X, E, S, Q, T, D, h->TOK + (long long)V * D, h->TTIN));
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

// su_img / qu_img: fp32 [.,3,224,224] images, or -- when img_h > 0 -- uint8 [., img_h, img_w, 3] decoded frames
int forward_impl(spm_handle* h, cudaStream_t st, int E, int S, int Q, int W, const void* su_img, const void* qu_img,
                 const float* labels, const float* real_s, const float* real_t, const long long* target_labels,
                 float tasks_per_batch, float* logits, float* dists, float* loss, float* acc, int* pred, int img_h = 0,
                 int img_w = 0) {
  SPM_TRY(check_shapes(h, E, S, Q, W));
  const int T = h->cfg.seq_len, D = h->D, N = S + Q;
  const long long nf = (long long)E * N * T;
  if (nf > h->xall_cap) {
    SPM_TRY(dalloc_t(h, &h->Xall, nf * D));
    h->xall_cap = nf;
    h->head_plans.clear();
  }
  if ((long long)E * Q * W > h->tmp_out_cap) {
    SPM_TRY(dalloc_t(h, &h->tmp_logits, (long long)E * Q * W));
    SPM_TRY(dalloc_t(h, &h->tmp_dists, E));
    h->tmp_out_cap = (long long)E * Q * W;
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
      SPM_TRY(dalloc_t(h, &h->feats, nf * D));
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
  SPM_CHECK(cfg->precision == SPM_PRECISION_BF16 || cfg->precision == SPM_PRECISION_FP32, "spm_create: unknown precision");
  SPM_CHECK(cfg->head == SPM_HEAD_CLIPSPM || cfg->head == SPM_HEAD_CLIPFSAR || cfg->head == SPM_HEAD_STEN,
            "spm_create: unknown head");
  SPM_CHECK(cfg->head != SPM_HEAD_STEN || cfg->seq_len == 8,
            "spm_create: the STEN head reshapes to 8 frames per video (models/model_sten.py:65-66)");
  SPM_CHECK(cfg->precision == SPM_PRECISION_BF16 || cfg->backbone == SPM_BACKBONE_VIT_B16,
            "spm_create: SPM_PRECISION_FP32 is implemented for the ViT-B/16 backbone only");
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
  if (n_cls > h->n_cls) SPM_TRY(dalloc_t(h, &h->text, (long long)n_cls * dim));
  SPM_CUDA(cudaMemcpyAsync(h->text, table, (size_t)n_cls * dim * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  h->n_cls = n_cls;
  h->text_set = true;
  return 0;
}

int spm_set_text_features_train(spm_handle* h, void* stream, const float* table, int n_cls, int dim) {
  SPM_CHECK(h != nullptr && table != nullptr, "spm_set_text_features_train: null argument");
  SPM_CHECK(dim == h->D, "spm_set_text_features_train: feature dim does not match the backbone's mid_dim");
  SPM_CHECK(n_cls >= 1, "spm_set_text_features_train: empty table");
  if (n_cls > h->n_cls_train) SPM_TRY(dalloc_t(h, &h->text_train, (long long)n_cls * dim));
  SPM_CUDA(cudaMemcpyAsync(h->text_train, table, (size_t)n_cls * dim * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  h->n_cls_train = n_cls;
  return 0;
}

int spm_class_logits(spm_handle* h, void* stream, int n_rows, int n_cls, float* out) {
  SPM_CHECK(h != nullptr && out != nullptr, "spm_class_logits: null argument");
  SPM_CHECK(h->cfg.head == SPM_HEAD_CLIPFSAR, "spm_class_logits: only the CLIP-FSAR head produces class logits");
  SPM_CHECK(h->cls_rows > 0, "spm_class_logits: no class logits available (no head call yet, or text_features_train not set)");
  SPM_CHECK(n_rows == h->cls_rows && n_cls == h->n_cls_train, "spm_class_logits: shape does not match the last head call");
  SPM_CUDA(cudaMemcpyAsync(out, h->CLS, (size_t)n_rows * n_cls * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
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
  if (R * fs * frame_bytes > h->stage_cap_bytes_s || R * fq * frame_bytes > h->stage_cap_bytes_q ||
      R * fs > h->stage_cap_frames_s || R * fq > h->stage_cap_frames_q) {
    SPM_TRY(dalloc_t(h, &s.su, R * fs * frame_bytes));
    SPM_TRY(dalloc_t(h, &s.qu, R * fq * frame_bytes));
    h->stage_cap_bytes_s = R * fs * frame_bytes; h->stage_cap_bytes_q = R * fq * frame_bytes;
    SPM_TRY(dalloc_t(h, &s.lab, (long long)R * S));
    SPM_TRY(dalloc_t(h, &s.rs, (long long)R * S));
    SPM_TRY(dalloc_t(h, &s.rt, (long long)R * Q));
    SPM_TRY(dalloc_t(h, &s.tl, (long long)R * Q));
    SPM_TRY(dalloc_t(h, &s.logits, (long long)R * Q * W));
    SPM_TRY(dalloc_t(h, &s.dists, R));
    SPM_TRY(dalloc_t(h, &s.loss, R));
    SPM_TRY(dalloc_t(h, &s.acc, R));
    SPM_TRY(dalloc_t(h, &s.pred, (long long)R * Q));
    h->stage_cap_frames_s = R * fs; h->stage_cap_frames_q = R * fq;
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
    SPM_TRY(dalloc_t(h, &h->pf_su, pf_s));
    SPM_TRY(dalloc_t(h, &h->pf_qu, pf_q));
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
  // Behind this call's own copies (same FIFO copy stream): the first chunk of the next call, while the last chunks of
  // this one compute.  The buffers may still be read by this call's chunk 0.
  h->pf_src_su = h->pf_src_qu = nullptr;
  if (hint_su != nullptr && hint_qu != nullptr && hint_n > 0) {
    SPM_CUDA(cudaStreamWaitEvent(cs, h->ev_done[0], 0));
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
