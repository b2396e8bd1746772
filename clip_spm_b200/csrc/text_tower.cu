// CLIP text-prompt tower: token ids -> prompt features, and their per-class mean over the
// prompt templates -- what the reference computes once in CNN.__init__ (models/model_clipspm.py:45-70) with
// CLIP.encode_text (models/clip_fsar.py:793-805): token + positional embedding, 12 residual attention blocks of
// width 512 with a causal mask (:778-784), ln_final, the <|endoftext|> row, text_projection.
// Runs once per class list, so it simply reuses the library's pieces on fp32 activations: LayerNorm kernel, the
// tcgen05 GEMM in tf32 mode (or the exact fp32 SIMT GEMM in SPM_PRECISION_FP32) with fused bias / QuickGELU /
// residual epilogues, plus two small kernels of its own (embedding + EOT search, causal attention).
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/clipspm_b200.h"
#include "api_common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "profile.cuh"

namespace spm {
int device_sm_count(int* out);

namespace {
constexpr int TC = 512, TL = 77, TH = 8, THD = 64, TLAYERS = 12, TVOCAB = 49408, TB_MAX = 256;

#define TXT_LAUNCH_CHECK()                                                                                  \
  do {                                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                                    \
    if (_e != cudaSuccess) { set_error(std::string("text kernel launch: ") + cudaGetErrorString(_e)); return 1; } \
    count_launch();                                                                                         \
  } while (0)

// x[b,t,:] = token_embedding[tokens[b,t]] + positional_embedding[t]; eot[b] = argmax_t tokens[b,t] (first maximum)
__global__ void text_embed_kernel(const int* __restrict__ tokens, const float* __restrict__ emb,
                                  const float* __restrict__ pos, float* __restrict__ x, int* __restrict__ eot) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < TL * (TC / 4); i += blockDim.x) {
    const int t = i / (TC / 4), c = i % (TC / 4);
    int tok = tokens[b * TL + t];
    tok = min(max(tok, 0), TVOCAB - 1);
    const float4 e = __ldg(reinterpret_cast<const float4*>(emb + (long long)tok * TC) + c);
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos + t * TC) + c);
    reinterpret_cast<float4*>(x + ((long long)b * TL + t) * TC)[c] = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
  }
  if (threadIdx.x == 0) {
    int best = 0, bv = tokens[b * TL];
    for (int t = 1; t < TL; ++t) {
      const int v = tokens[b * TL + t];
      if (v > bv) { bv = v; best = t; }
    }
    eot[b] = best;
  }
}

// causal softmax(q k^T / 8) v, 77 tokens x 64 dims per (sequence, head), fp32; qkv [B*77, 1536] -> out [B*77, 512]
__global__ void __launch_bounds__(128)
text_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out) {
  __shared__ float sK[TL][THD + 1];
  __shared__ float sV[TL][THD];
  __shared__ float sQ[4][THD];
  const int b = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + (long long)b * TL * (3 * TC) + h * THD;
  for (int i = threadIdx.x; i < TL * THD; i += blockDim.x) {
    const int r = i / THD, d = i % THD;
    sK[r][d] = base[(long long)r * (3 * TC) + TC + d];
    sV[r][d] = base[(long long)r * (3 * TC) + 2 * TC + d];
  }
  __syncthreads();
  for (int r = warp; r < TL; r += 4) {
    sQ[warp][lane] = base[(long long)r * (3 * TC) + lane];
    sQ[warp][lane + 32] = base[(long long)r * (3 * TC) + lane + 32];
    __syncwarp();
    float s[3], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int key = lane + 32 * j;
      s[j] = -INFINITY;
      if (key <= r) {  // causal: a token attends to itself and the past only
        float a = 0.f;
        for (int d = 0; d < THD; ++d) a = fmaf(sQ[warp][d], sK[key][d], a);
        s[j] = a * 0.125f;
      }
      mx = fmaxf(mx, s[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      s[j] = (lane + 32 * j <= r) ? expf(s[j] - mx) : 0.f;
      l += s[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    float a0 = 0.f, a1 = 0.f;
    for (int key = 0; key <= r; ++key) {
      const float p = __shfl_sync(0xffffffffu, s[key >> 5], key & 31);
      a0 = fmaf(p, sV[key][lane], a0);
      a1 = fmaf(p, sV[key][lane + 32], a1);
    }
    const float inv = 1.f / l;
    float* o = out + ((long long)b * TL + r) * TC + h * THD;
    o[lane] = a0 * inv;
    o[lane + 32] = a1 * inv;
    __syncwarp();
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ x, const int* __restrict__ eot, float* __restrict__ out) {
  const int b = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(x + ((long long)b * TL + eot[b]) * TC);
  for (int c = threadIdx.x; c < TC / 4; c += blockDim.x) reinterpret_cast<float4*>(out + (long long)b * TC)[c] = src[c];
}

// out[c, :] = mean over templates t of feats[t * n_cls + c, :]
__global__ void template_mean_kernel(const float* __restrict__ feats, int n_templates, int n_cls, int D,
                                     float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cls * D) return;
  const int c = i / D, d = i % D;
  float a = 0.f;
  for (int t = 0; t < n_templates; ++t) a += feats[((long long)t * n_cls + c) * D + d];
  out[i] = a / (float)n_templates;
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)R * Cc) return;
  out[(long long)(i % Cc) * R + i / Cc] = in[i];
}

struct LayerW {
  float *qkv_w, *qkv_b, *out_w, *out_b, *fc_w, *fc_b, *proj_w, *proj_b, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};
struct Plan {
  GemmOp qkv[TLAYERS], outp[TLAYERS], fc[TLAYERS], proj[TLAYERS], fin;
};
}  // namespace
}  // namespace spm

struct spm_text {
  int D = 512, kind = spm::GEMM_TF32, sms = 148;
  bool loaded = false;
  std::vector<void*> allocs;
  float *emb = nullptr, *pos = nullptr, *lnf_g = nullptr, *lnf_b = nullptr, *projT = nullptr;
  spm::LayerW layer[spm::TLAYERS];
  float *x = nullptr, *xn = nullptr, *qkv = nullptr, *att = nullptr, *hid = nullptr, *xe = nullptr, *xen = nullptr,
        *feats = nullptr;
  long long feats_cap = 0;
  int* eot = nullptr;
  std::map<int, std::unique_ptr<spm::Plan>> plans;
};

namespace spm {
namespace {
template <class T>
int talloc(spm_text* h, T** p, long long n) {
  SPM_CUDA(cudaMalloc(reinterpret_cast<void**>(p), (size_t)std::max<long long>(n, 4) * sizeof(T)));
  h->allocs.push_back(*p);
  return 0;
}

int plan(spm_text* h, GemmOp* op, const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
         const GemmEpilogue& ep) {
  const char* err = "";
  if (gemm_plan(op, h->kind, A, lda, B, ldb, M, N, K, ep, h->sms, &err)) {
    set_error(std::string("text gemm_plan: ") + err);
    return 1;
  }
  return 0;
}

int get_plan(spm_text* h, int B, Plan** out) {
  auto it = h->plans.find(B);
  if (it != h->plans.end()) { *out = it->second.get(); return 0; }
  std::unique_ptr<Plan> pl(new Plan());
  const int M = B * TL;
  for (int i = 0; i < TLAYERS; ++i) {
    const LayerW& l = h->layer[i];
    GemmEpilogue e1; e1.bias = l.qkv_b; e1.out = h->qkv; e1.ldo = 3 * TC;
    SPM_TRY(plan(h, &pl->qkv[i], h->xn, TC, l.qkv_w, TC, M, 3 * TC, TC, e1));
    GemmEpilogue e2; e2.bias = l.out_b; e2.residual = h->x; e2.ldr = TC; e2.out = h->x; e2.ldo = TC;
    SPM_TRY(plan(h, &pl->outp[i], h->att, TC, l.out_w, TC, M, TC, TC, e2));
    GemmEpilogue e3; e3.bias = l.fc_b; e3.act = ACT_QUICKGELU; e3.out = h->hid; e3.ldo = 4 * TC;
    SPM_TRY(plan(h, &pl->fc[i], h->xn, TC, l.fc_w, TC, M, 4 * TC, TC, e3));
    GemmEpilogue e4; e4.bias = l.proj_b; e4.residual = h->x; e4.ldr = TC; e4.out = h->x; e4.ldo = TC;
    SPM_TRY(plan(h, &pl->proj[i], h->hid, 4 * TC, l.proj_w, 4 * TC, M, TC, 4 * TC, e4));
  }
  GemmEpilogue ef; ef.out = h->x /* patched per call */; ef.ldo = h->D;
  SPM_TRY(plan(h, &pl->fin, h->xen, TC, h->projT, TC, B, h->D, TC, ef));
  *out = pl.get();
  h->plans[B] = std::move(pl);
  return 0;
}

int run_gemm(const GemmOp& op, cudaStream_t st) {
  const char* err = "";
  if (gemm_run(&op, st, &err)) { set_error(std::string("text gemm_run: ") + err); return 1; }
  return 0;
}

// tokens [B,77] (device) -> out [B, D]
int encode_chunk(spm_text* h, cudaStream_t st, const int* tokens, int B, float* out) {
  Plan* pl;
  SPM_TRY(get_plan(h, B, &pl));
  const int M = B * TL;
  text_embed_kernel<<<B, 256, 0, st>>>(tokens, h->emb, h->pos, h->x, h->eot);
  TXT_LAUNCH_CHECK();
  for (int i = 0; i < TLAYERS; ++i) {
    const LayerW& l = h->layer[i];
    if (k_layernorm(st, h->x, TC, M, TC, l.ln1_g, l.ln1_b, nullptr, 0, h->xn, nullptr, TC)) { set_error("text: layernorm"); return 1; }
    SPM_TRY(run_gemm(pl->qkv[i], st));
    text_attention_kernel<<<dim3(B, TH), 128, 0, st>>>(h->qkv, h->att);
    TXT_LAUNCH_CHECK();
    SPM_TRY(run_gemm(pl->outp[i], st));
    if (k_layernorm(st, h->x, TC, M, TC, l.ln2_g, l.ln2_b, nullptr, 0, h->xn, nullptr, TC)) { set_error("text: layernorm"); return 1; }
    SPM_TRY(run_gemm(pl->fc[i], st));
    SPM_TRY(run_gemm(pl->proj[i], st));
  }
  gather_rows_kernel<<<B, 128, 0, st>>>(h->x, h->eot, h->xe);   // ln_final is per token: only the EOT rows are needed
  TXT_LAUNCH_CHECK();
  if (k_layernorm(st, h->xe, TC, B, TC, h->lnf_g, h->lnf_b, nullptr, 0, h->xen, nullptr, TC)) { set_error("text: layernorm"); return 1; }
  GemmOp fin = pl->fin;
  fin.ep.out = out;
  SPM_TRY(run_gemm(fin, st));
  return 0;
}
}  // namespace
}  // namespace spm

using namespace spm;

extern "C" {

int spm_text_create(int embed_dim, int precision, spm_text** out) {
  SPM_CHECK(out != nullptr, "spm_text_create: null argument");
  SPM_CHECK(embed_dim == 512 || embed_dim == 1024, "spm_text_create: embed_dim must be 512 (ViT-B/16) or 1024 (RN50)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("spm_text_create: no CUDA device -- this library has no CPU path");
    return 1;
  }
  std::unique_ptr<spm_text> h(new spm_text());
  h->D = embed_dim;
  h->kind = precision == SPM_PRECISION_FP32 ? GEMM_F32_SIMT : GEMM_TF32;
  SPM_TRY(device_sm_count(&h->sms));
  const char* err = "";
  if (gemm_init(&err)) { set_error(err); return 1; }
  *out = h.release();
  return 0;
}

int spm_text_destroy(spm_text* h) {
  if (h == nullptr) return 0;
  cudaDeviceSynchronize();
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return 0;
}

int spm_text_load_weights(spm_text* h, void* stream, int n, const char* const* names, const void* const* dev_ptrs,
                          const int64_t* numel) {
  SPM_CHECK(h && names && dev_ptrs && numel, "spm_text_load_weights: null argument");
  SPM_CHECK(!h->loaded, "spm_text_load_weights: weights already loaded");
  cudaStream_t st = (cudaStream_t)stream;
  std::unordered_map<std::string, std::pair<const float*, long long>> m;
  for (int i = 0; i < n; ++i) m[names[i]] = {static_cast<const float*>(dev_ptrs[i]), (long long)numel[i]};
  auto take = [&](const std::string& name, long long ne, float** dst) -> int {
    auto it = m.find(name);
    if (it == m.end()) { set_error("spm_text_load_weights: missing tensor '" + name + "'"); return 1; }
    if (it->second.second != ne) { set_error("spm_text_load_weights: wrong size for '" + name + "'"); return 1; }
    SPM_TRY(talloc(h, dst, ne));
    SPM_CUDA(cudaMemcpyAsync(*dst, it->second.first, (size_t)ne * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
  };
  SPM_TRY(take("token_embedding.weight", (long long)TVOCAB * TC, &h->emb));
  SPM_TRY(take("positional_embedding", (long long)TL * TC, &h->pos));
  SPM_TRY(take("ln_final.weight", TC, &h->lnf_g));
  SPM_TRY(take("ln_final.bias", TC, &h->lnf_b));
  float* proj = nullptr;
  SPM_TRY(take("text_projection", (long long)TC * h->D, &proj));
  SPM_TRY(talloc(h, &h->projT, (long long)TC * h->D));
  transpose_kernel<<<(TC * h->D + 255) / 256, 256, 0, st>>>(proj, h->projT, TC, h->D);
  TXT_LAUNCH_CHECK();
  for (int i = 0; i < TLAYERS; ++i) {
    const std::string p = "transformer.resblocks." + std::to_string(i) + ".";
    LayerW& l = h->layer[i];
    SPM_TRY(take(p + "attn.in_proj_weight", 3LL * TC * TC, &l.qkv_w));
    SPM_TRY(take(p + "attn.in_proj_bias", 3 * TC, &l.qkv_b));
    SPM_TRY(take(p + "attn.out_proj.weight", (long long)TC * TC, &l.out_w));
    SPM_TRY(take(p + "attn.out_proj.bias", TC, &l.out_b));
    SPM_TRY(take(p + "mlp.c_fc.weight", 4LL * TC * TC, &l.fc_w));
    SPM_TRY(take(p + "mlp.c_fc.bias", 4 * TC, &l.fc_b));
    SPM_TRY(take(p + "mlp.c_proj.weight", 4LL * TC * TC, &l.proj_w));
    SPM_TRY(take(p + "mlp.c_proj.bias", TC, &l.proj_b));
    SPM_TRY(take(p + "ln_1.weight", TC, &l.ln1_g));
    SPM_TRY(take(p + "ln_1.bias", TC, &l.ln1_b));
    SPM_TRY(take(p + "ln_2.weight", TC, &l.ln2_g));
    SPM_TRY(take(p + "ln_2.bias", TC, &l.ln2_b));
  }
  const long long M = (long long)TB_MAX * TL;
  SPM_TRY(talloc(h, &h->x, M * TC));
  SPM_TRY(talloc(h, &h->xn, M * TC));
  SPM_TRY(talloc(h, &h->qkv, M * 3 * TC));
  SPM_TRY(talloc(h, &h->att, M * TC));
  SPM_TRY(talloc(h, &h->hid, M * 4 * TC));
  SPM_TRY(talloc(h, &h->xe, (long long)TB_MAX * TC));
  SPM_TRY(talloc(h, &h->xen, (long long)TB_MAX * TC));
  SPM_TRY(talloc(h, &h->eot, TB_MAX));
  SPM_CUDA(cudaStreamSynchronize(st));
  h->loaded = true;
  return 0;
}

int spm_text_encode(spm_text* h, void* stream, const int32_t* tokens, int n_texts, float* out) {
  SPM_CHECK(h != nullptr, "spm_text_encode: null handle");
  if (n_texts <= 0) return 0;
  SPM_CHECK(tokens && out, "spm_text_encode: null argument");
  SPM_CHECK(h->loaded, "spm_text_encode: weights not loaded");
  for (int b0 = 0; b0 < n_texts; b0 += TB_MAX) {
    const int B = std::min(TB_MAX, n_texts - b0);
    SPM_TRY(encode_chunk(h, (cudaStream_t)stream, tokens + (long long)b0 * TL, B, out + (long long)b0 * h->D));
  }
  return 0;
}

int spm_text_class_features(spm_text* h, void* stream, const int32_t* tokens, int n_templates, int n_classes,
                            float* out) {
  SPM_CHECK(h && tokens && out, "spm_text_class_features: null argument");
  SPM_CHECK(n_templates >= 1 && n_classes >= 1, "spm_text_class_features: empty input");
  const long long n = (long long)n_templates * n_classes;
  if (n > h->feats_cap) {
    SPM_TRY(talloc(h, &h->feats, n * h->D));
    h->feats_cap = n;
  }
  SPM_TRY(spm_text_encode(h, stream, tokens, (int)n, h->feats));
  template_mean_kernel<<<(n_classes * h->D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->feats, n_templates,
                                                                                       n_classes, h->D, out);
  TXT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
