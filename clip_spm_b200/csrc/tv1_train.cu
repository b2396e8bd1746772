// Training step, next piece: forward AND backward of the head's transformer block
// models/myRes.py:1053-1075 `Transformer_v1` (depth 1, q = k = v = x) --
//     y   = to_out(softmax(q k^T / sqrt(dh)) v) + b_out + x        q, k, v = to_{q,k,v}(LayerNorm(x))      (:1039-1040, :964-982)
//     out = W3 gelu(W0 y + b0) + b3 + y                                                                    (:1069, FeedForward :944-955)
// -- the block behind `context1` / `context2` of CLIP-SPM and `context2` of CLIP-FSAR / CPM2C, i.e. what
// `scaler.scale(loss).backward()` (run/main_run.py:252) differentiates most of the head's parameters through.
//
// Forward = the library's own kernels (LayerNorm, tcgen05 GEMMs with fused bias / GELU / residual epilogues, the short-
// sequence attention kernel); it leaves LN(x), qkv, the attention output, y and gelu(.) in the handle.  Backward:
//   * every dX = dY W and dW = dY^T X is the SAME tcgen05 GEMM (out = A B^T, both operands K-contiguous) on transposed
//     copies -- weights transposed once at load, activations by a tiled transpose kernel (rows padded to 16 bytes with
//     zeros, which add nothing to the reduction over rows); gradient accumulation across the residual branches rides in the
//     GEMM's residual epilogue;
//   * the pre-activation of the MLP is recomputed by one more GEMM (the forward fuses GELU into its epilogue and keeps only
//     the activated tensor), gelu' applied elementwise;
//   * attention backward: one CTA per (sequence, head) recomputes P from the saved q, k and forms dV = P^T dO,
//     dS = P o (dO V^T - rowsum(dO V^T o P)), dQ = dS K / sqrt(dh), dK = dS^T Q / sqrt(dh) in shared memory;
//   * LayerNorm backward: a warp per row (statistics recomputed), d gamma / d beta by a column-sum kernel; biases likewise.
// Dropout (Attention_qkv / FeedForward, active in the reference's train mode): replayable Philox masks, p = 0 by default.
// fp32 data; products in tf32 (precision 0) or exact fp32 SIMT (precision 1).
//
// The same code with three switches is the frame encoder's residual block models/clip_fsar.py:622-643
// `ResidualAttentionBlock` (spm_vitblock_*): a second LayerNorm in front of the MLP, biases on the fused q/k/v projection,
// QuickGELU, and 197-token x 12-head x 64-dim attention (forward: the fp32 kernel of sgemm_f32.cu; backward: the two-phase
// kernel below) -- i.e. the backward of the CLIP ViT-B/16 tower the reference's optimiser also steps.
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/clipspm_b200.h"
#include "api_common.cuh"
#include "gemm.cuh"
#include "head_kernels.cuh"
#include "kernels.cuh"
#include "profile.cuh"

struct spm_tv1 {
  int D = 0, heads = 0, dh = 0, inner = 0, mlp = 0, fp32 = 0, sms = 148;
  bool loaded = false;
  // ViT variant (clip_fsar.py:622-643): ln_2 before the MLP, q/k/v bias, QuickGELU, 197-token attention kernels
  bool vit = false;
  float *ln2_g = nullptr, *ln2_b = nullptr, *bqkv = nullptr, *H2 = nullptr;
  float* Lsm = nullptr;      // [frames * 12][208] log2 softmax denominators of the tensor-core attention forward
  bool have_L = false;
  // weights (reference layouts) and their transposes
  float *ln_g = nullptr, *ln_b = nullptr, *wqkv = nullptr, *wout = nullptr, *bout = nullptr, *w0 = nullptr, *b0 = nullptr,
        *w3 = nullptr, *b3 = nullptr;
  float *wqkvT = nullptr, *woutT = nullptr, *w0T = nullptr, *w3T = nullptr;
  // saved activations of the last forward, sized for cap_rows rows (the backward scratch is process-wide, Tv1Scratch)
  long long cap_rows = 0;
  const float* x = nullptr;   // the caller's input of the last forward (must stay alive until backward)
  int B = 0, n = 0;
  float *HN = nullptr, *QKV = nullptr, *AO = nullptr, *Y = nullptr, *FFH = nullptr;
  float p_atte = 0.f, p_ffn = 0.f;          // dropout of the NEXT forward (spm_tv1_set_dropout); fwd_* = of the last one
  unsigned long long seed = 0;
  float fwd_p_atte = 0.f, fwd_p_ffn = 0.f;
  unsigned long long fwd_seed = 0;

  std::vector<void*> allocs;
};

namespace spm {
int device_sm_count(int* out);   // api_gemm.cu
namespace {

#define TV1_KERNEL(call)                                                                                   \
  do {                                                                                                      \
    int _r = (call);                                                                                        \
    if (_r != 0) {                                                                                          \
      set_error(std::string(#call) + (_r < 0 ? ": unsupported shape" : std::string(": ") + cudaGetErrorString((cudaError_t)_r))); \
      return 1;                                                                                             \
    }                                                                                                       \
  } while (0)

#define TV1_LAUNCH_CHECK()                                                                                  \
  do {                                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                                    \
    if (_e != cudaSuccess) { set_error(std::string("tv1 kernel launch: ") + cudaGetErrorString(_e)); return 1; } \
    count_launch();                                                                                         \
  } while (0)

// out[c, r] = in[r, c] for r < R, 0 for R <= r < ldo  (in [R, C] row-major, out [C, ldo])
__global__ void transpose_pad_kernel(const float* __restrict__ in, int R, int C, float* __restrict__ out, int ldo) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[(long long)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < ldo) out[(long long)c * ldo + r] = tile[threadIdx.x][i];
  }
}

// two-stage column sum for tall matrices: partial sums of CS_CHUNKS row ranges, then their sum in a fixed order (deterministic,
// no atomics); the one-stage kernel below walks all R rows with C / 32 blocks, far too few for the encoder's 15 760+ rows
constexpr int CS_CHUNKS = 64;
__global__ void colsum_part_kernel(const float* __restrict__ a, int R, int C, int rows_per_chunk, float* __restrict__ part) {
  __shared__ float red[8][32];
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * rows_per_chunk, r1 = min(R, r0 + rows_per_chunk);
  float s = 0.f;
  if (c < C)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) s += a[(long long)r * C + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    part[(long long)blockIdx.y * C + c] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int n_chunks, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int k = 0; k < n_chunks; ++k) t += part[(long long)k * C + c];
  out[c] = t;
}

// out[c] = sum_r a[r, c] (* b[r, c] when b != null); fixed summation order (deterministic)
__global__ void colsum_kernel(const float* __restrict__ a, const float* __restrict__ b, int R, int C, float* __restrict__ out) {
  __shared__ float part[8][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < C)
    for (int r = threadIdx.y; r < R; r += 8) {
      const float v = a[(long long)r * C + c];
      s += b != nullptr ? v * b[(long long)r * C + c] : v;
    }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
    out[c] = t;
  }
}

// g[i] *= gelu'(pre[i]), gelu(x) = x Phi(x) (erf form, nn.GELU default): gelu'(x) = Phi(x) + x phi(x)
__global__ void gelu_bwd_kernel(float* __restrict__ g, const float* __restrict__ pre, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pre[i];
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
  g[i] *= cdf + x * pdf;
}

// g[i] = dy[i] * act'(.) from the layer's OUTPUT y: LeakyReLU (the slope is positive, so y and the pre-activation share
// their sign), sigmoid (y (1 - y)); `g` holds the recomputed pre-activation on entry for GELU
__global__ void act_bwd_kernel(float* __restrict__ g, const float* __restrict__ dy, const float* __restrict__ y, int act,
                               float slope, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d = 1.f;
  if (act == ACT_LEAKY) d = y[i] > 0.f ? 1.f : slope;
  else if (act == ACT_SIGMOID) d = y[i] * (1.f - y[i]);
  else if (act == ACT_GELU_ERF) {
    const float x = g[i];
    d = 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
  }
  g[i] = dy[i] * d;
}

// Philox4x32-10 (Salmon et al., SC'11; the generator behind torch's CUDA dropout): counter (c.x..c.w), key (k.x, k.y)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// nn.Dropout(p) with a replayable mask: element i of `site` is kept iff (word i % 4 of philox(counter = (i / 4, site),
// key = seed) >> 8) * 2^-24 >= p, kept values are scaled by 1 / (1 - p);  y = dropout(x) (+ add).  The same call with the
// upstream gradient as x is the backward.  One thread per four consecutive elements (one Philox block).
__global__ void dropout_kernel(const float* x, const float* __restrict__ add, float* y, long long n,
                               float p, unsigned long long seed, unsigned site, const unsigned long long* __restrict__ seed_src) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q * 4 >= n) return;
  if (seed_src != nullptr) seed += *seed_src;     // device-resident seed offset (CUDA-graph replays: spm_dropout_seed_source)
  const uint4 r = philox4x32_10(make_uint4((unsigned)q, (unsigned)(q >> 32), site, 0u),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const unsigned w[4] = {r.x, r.y, r.z, r.w};
  const float scale = 1.f / (1.f - p);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long i = q * 4 + j;
    if (i < n) {
      const float u = (float)(w[j] >> 8) * 5.9604644775390625e-08f;
      float v = u >= p ? x[i] * scale : 0.f;
      if (add != nullptr) v += add[i];
      y[i] = v;
    }
  }
}

// LayerNorm backward, a warp per row: xhat = (x - mean) rstd;  dx = rstd (g gamma - mean(g gamma) - xhat mean(g gamma xhat)) + add
// also writes xhat o g (for d gamma = column sum) into `gx`
__global__ void ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ gamma,
                              const float* __restrict__ add, int R, int C, float* __restrict__ dx, float* __restrict__ gx) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* xr = x + (long long)row * C;
  const float* gr = g + (long long)row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; v += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / (float)C + 1e-5f);
  float a = 0.f, b = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float gg = gr[c] * gamma[c], xh = (xr[c] - mean) * rstd;
    a += gg;
    b += gg * xh;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  a /= (float)C;
  b /= (float)C;
  for (int c = lane; c < C; c += 32) {
    const float xh = (xr[c] - mean) * rstd;
    dx[(long long)row * C + c] = rstd * (gr[c] * gamma[c] - a - xh * b) + (add != nullptr ? add[(long long)row * C + c] : 0.f);
    gx[(long long)row * C + c] = gr[c] * xh;
  }
}

// Attention backward for one (sequence, head): qkv rows [n, 3*inner] (q | k | v, head h at columns h*dh of each third),
// dO rows [n, inner] -> dqkv rows in the same layout.  Shared memory: q, k, v, dO [n][dh+1], P, dS [n][n].
__global__ void __launch_bounds__(256)
seq_attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dO, float* __restrict__ dqkv, int n,
                         int heads, int dh) {
  extern __shared__ float sm_ab[];
  const int seq = blockIdx.x, head = blockIdx.y, inner = heads * dh, ld = 3 * inner, st = dh + 1;
  float* sq = sm_ab;
  float* sk = sq + n * st;
  float* sv = sk + n * st;
  float* so = sv + n * st;
  float* sp = so + n * st;      // [n][n]
  float* sd = sp + n * n;       // [n][n]
  const long long row0 = (long long)seq * n;
  for (int i = threadIdx.x; i < n * dh; i += blockDim.x) {
    const int j = i / dh, d = i % dh;
    const float* r = qkv + (row0 + j) * ld + head * dh + d;
    sq[j * st + d] = r[0];
    sk[j * st + d] = r[inner];
    sv[j * st + d] = r[2 * inner];
    so[j * st + d] = dO[(row0 + j) * inner + head * dh + d];
  }
  __syncthreads();
  const float scale = rsqrtf((float)dh);
  // S = q k^T scale and dP = dO v^T, one (i, j) per thread
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e % n;
    float s = 0.f, dp = 0.f;
    for (int d = 0; d < dh; ++d) {
      s += sq[i * st + d] * sk[j * st + d];
      dp += so[i * st + d] * sv[j * st + d];
    }
    sp[e] = s * scale;
    sd[e] = dp;
  }
  __syncthreads();
  // row softmax and dS = P o (dP - sum_j dP P), one row per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < n; i += 8) {
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) m = fmaxf(m, sp[i * n + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float z = 0.f;
    for (int j = lane; j < n; j += 32) { const float p = expf(sp[i * n + j] - m); sp[i * n + j] = p; z += p; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    const float inv = 1.f / z;
    float dot = 0.f;
    for (int j = lane; j < n; j += 32) { const float p = sp[i * n + j] * inv; sp[i * n + j] = p; dot += p * sd[i * n + j]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int j = lane; j < n; j += 32) sd[i * n + j] = sp[i * n + j] * (sd[i * n + j] - dot);
  }
  __syncthreads();
  // dQ[i] = scale sum_j dS[i,j] K[j];  dK[j] = scale sum_i dS[i,j] Q[i];  dV[j] = sum_i P[i,j] dO[i]
  for (int e = threadIdx.x; e < n * dh; e += blockDim.x) {
    const int i = e / dh, d = e % dh;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j < n; ++j) {
      dq += sd[i * n + j] * sk[j * st + d];
      dk += sd[j * n + i] * sq[j * st + d];
      dv += sp[j * n + i] * so[j * st + d];
    }
    float* r = dqkv + (row0 + i) * ld + head * dh + d;
    r[0] = dq * scale;
    r[inner] = dk * scale;
    r[2 * inner] = dv;
  }
}
// g[i] *= quickgelu'(pre[i]), quickgelu(x) = x sigmoid(1.702 x) (clip_fsar.py:618-620)
__global__ void quickgelu_bwd_kernel(float* __restrict__ g, const float* __restrict__ pre, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pre[i];
  const float sg = 1.f / (1.f + expf(-1.702f * x));
  g[i] *= sg * (1.f + 1.702f * x * (1.f - sg));
}

// Attention backward of the frame encoder: 197 tokens x 64 dims per (frame, head), 12 heads (clip_fsar.py:626,636-638).
// One CTA per (frame, head) keeps q, k, v, dO [197][65] in shared memory (205 KB) and works in two phases, each the
// forward kernel's shape (a warp per row, a lane per 7 columns of the score row, outputs as 2 dims per lane):
//   phase 1, per QUERY row i:  p_i = softmax(q_i K^T / 8), dp_ij = dO_i . v_j, delta_i = sum_j p_ij dp_ij,
//                              dq_i = 1/8 sum_j p_ij (dp_ij - delta_i) k_j;   keeps m_i, 1 / l_i, delta_i
//   phase 2, per KEY row j:    p_ij and dp_ij rebuilt from the kept row statistics,
//                              dk_j = 1/8 sum_i p_ij (dp_ij - delta_i) q_i,  dv_j = sum_i p_ij dO_i
// -- no atomics, no [197 x 197] table; the price is computing the two score products twice.
constexpr int VL = 197, VHD = 64, VHEADS = 12, VC = 768, VST = VHD + 1;
constexpr int VIT_ATT_BWD_SMEM = (4 * VL * VST + 3 * VL) * 4;

// s[r][t] = bc1[row_r] . st1[col_t], dp[r][t] = bc2[row_r] . st2[col_t] for 4 rows (warp-uniform: broadcast reads) and the
// lane's 7 columns col_t = lane + 32 t: a 4 x 7 register tile per lane, 8 broadcast + 14 strided shared-memory reads per
// 56 FMAs (the one-row version paid 2 reads per FMA and was bound by shared-memory bandwidth)
__device__ __forceinline__ void vit_scores4(const float* __restrict__ bc1, const float* __restrict__ bc2,
                                            const float* __restrict__ st1, const float* __restrict__ st2, const int (&rows)[4],
                                            int lane, float (&sc)[4][7], float (&dp)[4][7]) {
  int col[7];
#pragma unroll
  for (int t = 0; t < 7; ++t) col[t] = min(lane + 32 * t, VL - 1) * VST;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int t = 0; t < 7; ++t) { sc[r][t] = 0.f; dp[r][t] = 0.f; }
#pragma unroll 4
  for (int d = 0; d < VHD; ++d) {
    float a[4], b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { a[r] = bc1[rows[r] * VST + d]; b[r] = bc2[rows[r] * VST + d]; }
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      const float x = st1[col[t] + d], y = st2[col[t] + d];
#pragma unroll
      for (int r = 0; r < 4; ++r) { sc[r][t] = fmaf(a[r], x, sc[r][t]); dp[r][t] = fmaf(b[r], y, dp[r][t]); }
    }
  }
}

// out[r][0..1] = sum_j w[r][j] mat[j][lane, lane + 32] with w spread over the lanes (column j = lane + 32 t lives in lane j % 32)
__device__ __forceinline__ void vit_accum4(const float (&w)[4][7], const float* __restrict__ mat, int lane, float (&out)[4][2]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) { out[r][0] = 0.f; out[r][1] = 0.f; }
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    const int nsrc = t < 6 ? 32 : VL - 192;
    for (int src = 0; src < nsrc; ++src) {
      const int j = src + 32 * t;
      const float m0 = mat[j * VST + lane], m1 = mat[j * VST + lane + 32];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float wr = __shfl_sync(0xffffffffu, w[r][t], src);
        out[r][0] = fmaf(wr, m0, out[r][0]);
        out[r][1] = fmaf(wr, m1, out[r][1]);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
vit_attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dO, float* __restrict__ dqkv) {
  extern __shared__ float sm_vb[];
  float* sQ = sm_vb;
  float* sK = sQ + VL * VST;
  float* sV = sK + VL * VST;
  float* sO = sV + VL * VST;
  float* sM = sO + VL * VST;      // row max
  float* sL = sM + VL;            // 1 / row sum
  float* sD = sL + VL;            // delta
  const int frame = blockIdx.x / VHEADS, head = blockIdx.x % VHEADS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)frame * VL;
  const float* base = qkv + row0 * (3 * VC) + head * VHD;
  for (int i = threadIdx.x; i < VL * VHD; i += blockDim.x) {
    const int r = i / VHD, d = i % VHD;
    sQ[r * VST + d] = base[(long long)r * (3 * VC) + d];
    sK[r * VST + d] = base[(long long)r * (3 * VC) + VC + d];
    sV[r * VST + d] = base[(long long)r * (3 * VC) + 2 * VC + d];
    sO[r * VST + d] = dO[(row0 + r) * VC + head * VHD + d];
  }
  __syncthreads();
  float* out = dqkv + row0 * (3 * VC) + head * VHD;
  constexpr int GROUPS = (VL + 3) / 4;
  float sc[4][7], dp[4][7], acc[4][2];
  // ---- phase 1: query rows, four per warp pass
  for (int g = warp; g < GROUPS; g += 8) {
    int rows[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) rows[r] = min(4 * g + r, VL - 1);
    vit_scores4(sQ, sO, sK, sV, rows, lane, sc, dp);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int t = 0; t < 7; ++t) {
        sc[r][t] = (lane + 32 * t < VL) ? sc[r][t] * 0.125f : -INFINITY;
        mx = fmaxf(mx, sc[r][t]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float l = 0.f;
#pragma unroll
      for (int t = 0; t < 7; ++t) {
        sc[r][t] = (lane + 32 * t < VL) ? expf(sc[r][t] - mx) : 0.f;
        l += sc[r][t];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
      const float inv = 1.f / l;
      float delta = 0.f;
#pragma unroll
      for (int t = 0; t < 7; ++t) { sc[r][t] *= inv; delta = fmaf(sc[r][t], dp[r][t], delta); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, o);
      if (lane == 0 && 4 * g + r < VL) { sM[rows[r]] = mx; sL[rows[r]] = inv; sD[rows[r]] = delta; }
#pragma unroll
      for (int t = 0; t < 7; ++t) sc[r][t] *= dp[r][t] - delta;       // dS (up to the 1/8)
    }
    vit_accum4(sc, sK, lane, acc);
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (4 * g + r < VL) {
        out[(long long)rows[r] * (3 * VC) + lane] = acc[r][0] * 0.125f;
        out[(long long)rows[r] * (3 * VC) + lane + 32] = acc[r][1] * 0.125f;
      }
  }
  __syncthreads();
  // ---- phase 2: key rows, four per warp pass; the lane's columns are now QUERY rows
  for (int g = warp; g < GROUPS; g += 8) {
    int rows[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) rows[r] = min(4 * g + r, VL - 1);
    vit_scores4(sK, sV, sQ, sO, rows, lane, sc, dp);
    float m[7], li[7], de[7];
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      const int i = min(lane + 32 * t, VL - 1);
      m[t] = sM[i]; li[t] = sL[i]; de[t] = sD[i];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int t = 0; t < 7; ++t) {
        const float pr = (lane + 32 * t < VL) ? expf(sc[r][t] * 0.125f - m[t]) * li[t] : 0.f;
        sc[r][t] = pr;                              // P
        dp[r][t] = pr * (dp[r][t] - de[t]);         // dS (up to the 1/8)
      }
    vit_accum4(dp, sQ, lane, acc);
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (4 * g + r < VL) {
        out[(long long)rows[r] * (3 * VC) + VC + lane] = acc[r][0] * 0.125f;
        out[(long long)rows[r] * (3 * VC) + VC + lane + 32] = acc[r][1] * 0.125f;
      }
    vit_accum4(sc, sO, lane, acc);
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (4 * g + r < VL) {
        out[(long long)rows[r] * (3 * VC) + 2 * VC + lane] = acc[r][0];
        out[(long long)rows[r] * (3 * VC) + 2 * VC + lane + 32] = acc[r][1];
      }
  }
}

// ---- the same backward on the tensor cores (mma.sync m16n8k8 tf32, fp32 accumulation): the tf32 training path ----
// One CTA per (frame, head); q, k, v, dO rounded to tf32 in shared memory as [208][68] (rows >= 197 zero; a row stride of 68
// floats makes the scalar fragment reads conflict-free).  Row statistics come first, so that no pass needs a row reduction:
//   delta_i = dO_i . O_i            (= sum_j p_ij dp_ij; O is the forward's attention output, kept by the block)
//   L_i     = log2 sum_j 2^(s_ij)   (s = q k^T scaled to base 2; online over the key tiles, then across the 4 lanes of a row)
// phase 1, a warp per 16 QUERY rows, key tiles of 8: S and dP tiles by 2 x 8 MMAs, p = 2^(s - L_i), dS = p (dp - delta_i) in the
//   accumulator registers, which ARE the A fragment of the next product (slot t <-> column 2t, slot t+4 <-> column 2t+1, the
//   B rows read in the same order): dQ += dS K.
// phase 2, a warp per 16 KEY rows, query tiles of 8: the transposed tiles K Q^T and V dO^T with the column statistics,
//   dK += dS^T Q, dV += P^T dO.
// 25 key tiles x (16 + 8) MMAs per row tile instead of 35 MFLOP of FFMA per (frame, head).
constexpr int ML = 68, MROWS = 208, MTILES = MROWS / 16, NTILES = (VL + 7) / 8;
constexpr int VIT_ATT_BWD_MMA_SMEM = (4 * MROWS * ML + 2 * MROWS) * 4;
constexpr float V_SCALE_LOG2 = 0.125f * 1.4426950408889634f;

__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A fragments of rows r0 .. r0+15 of X over the 64 dims: k-slot t <-> dim 8 ks + 2t, slot t+4 <-> dim 8 ks + 2t + 1
__device__ __forceinline__ void vit_afrag(const float* __restrict__ X, int r0, int g, int t, uint32_t (&a)[8][4]) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const float2 v0 = *reinterpret_cast<const float2*>(X + (r0 + g) * ML + 8 * ks + 2 * t);
    const float2 v1 = *reinterpret_cast<const float2*>(X + (r0 + g + 8) * ML + 8 * ks + 2 * t);
    a[ks][0] = __float_as_uint(v0.x); a[ks][1] = __float_as_uint(v1.x);
    a[ks][2] = __float_as_uint(v0.y); a[ks][3] = __float_as_uint(v1.y);
  }
}
// c[16 x 8] = X rows (fragments a) . Y[j0 .. j0+7, :]^T
__device__ __forceinline__ void vit_score_tile(const uint32_t (&a)[8][4], const float* __restrict__ Y, int j0, int g, int t,
                                               float (&c)[4]) {
  c[0] = c[1] = c[2] = c[3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const float2 v = *reinterpret_cast<const float2*>(Y + (j0 + g) * ML + 8 * ks + 2 * t);
    mma_tf32(c, a[ks], __float_as_uint(v.x), __float_as_uint(v.y));
  }
}
// acc[16 x 64] += W[16 x 8] . Z[j0 .. j0+7, :64], W given in the accumulator layout (c0,c1: row g, cols 2t,2t+1; c2,c3: row g+8)
__device__ __forceinline__ void vit_accum_tile(const float (&w)[4], const float* __restrict__ Z, int j0, int g, int t,
                                               float (&acc)[8][4]) {
  const uint32_t a[4] = {__float_as_uint(tf32_round(w[0])), __float_as_uint(tf32_round(w[2])),
                         __float_as_uint(tf32_round(w[1])), __float_as_uint(tf32_round(w[3]))};
  const float* z0 = Z + (j0 + 2 * t) * ML + g;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) mma_tf32(acc[nd], a, __float_as_uint(z0[8 * nd]), __float_as_uint(z0[ML + 8 * nd]));
}

// Forward of the same attention on the tensor cores (tf32 training path): L_i first (as in the backward's phase 0), then
// p = 2^(s - L_i) is final at once and O += P V needs no rescaling; L is kept for the backward.
constexpr int VIT_ATT_FWD_MMA_SMEM = (3 * MROWS * ML + MROWS) * 4;

__global__ void __launch_bounds__(256, 1)
vit_attention_fwd_mma_kernel(const float* __restrict__ qkv, float* __restrict__ out, float* __restrict__ Lout) {
  extern __shared__ __align__(16) float sm_vf[];
  float* sQ = sm_vf;
  float* sK = sQ + MROWS * ML;
  float* sV = sK + MROWS * ML;
  const int frame = blockIdx.x / VHEADS, head = blockIdx.x % VHEADS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long row0 = (long long)frame * VL;
  const float* base = qkv + row0 * (3 * VC) + head * VHD;
  for (int i = threadIdx.x; i < MROWS * (VHD / 4); i += blockDim.x) {
    const int r = i / (VHD / 4), d = (i % (VHD / 4)) * 4;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f), k = q, v = q;
    if (r < VL) {
      q = *reinterpret_cast<const float4*>(base + (long long)r * (3 * VC) + d);
      k = *reinterpret_cast<const float4*>(base + (long long)r * (3 * VC) + VC + d);
      v = *reinterpret_cast<const float4*>(base + (long long)r * (3 * VC) + 2 * VC + d);
    }
    *reinterpret_cast<float4*>(sQ + r * ML + d) = make_float4(tf32_round(q.x), tf32_round(q.y), tf32_round(q.z), tf32_round(q.w));
    *reinterpret_cast<float4*>(sK + r * ML + d) = make_float4(tf32_round(k.x), tf32_round(k.y), tf32_round(k.z), tf32_round(k.w));
    *reinterpret_cast<float4*>(sV + r * ML + d) = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
  }
  __syncthreads();
  uint32_t xa[8][4];
  float c[4], acc[8][4];
  float* Lrow = Lout + (long long)blockIdx.x * MROWS;
  for (int rt = warp; rt < MTILES; rt += 8) {
    const int r0 = 16 * rt;
    vit_afrag(sQ, r0, g, t, xa);
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
#pragma unroll 1
    for (int jt = 0; jt < NTILES; ++jt) {
      const int j0 = 8 * jt, col = j0 + 2 * t;
      vit_score_tile(xa, sK, j0, g, t, c);
      const float s00 = col < VL ? c[0] * V_SCALE_LOG2 : -INFINITY, s01 = col + 1 < VL ? c[1] * V_SCALE_LOG2 : -INFINITY;
      const float s10 = col < VL ? c[2] * V_SCALE_LOG2 : -INFINITY, s11 = col + 1 < VL ? c[3] * V_SCALE_LOG2 : -INFINITY;
      const float n0 = fmaxf(m0, fmaxf(s00, s01)), n1 = fmaxf(m1, fmaxf(s10, s11));
      if (n0 > -INFINITY) { l0 = l0 * exp2f(m0 - n0) + exp2f(s00 - n0) + exp2f(s01 - n0); m0 = n0; }
      if (n1 > -INFINITY) { l1 = l1 * exp2f(m1 - n1) + exp2f(s10 - n1) + exp2f(s11 - n1); m1 = n1; }
    }
#pragma unroll
    for (int off = 1; off <= 2; off <<= 1) {
      const float mo0 = __shfl_xor_sync(0xffffffffu, m0, off), lo0 = __shfl_xor_sync(0xffffffffu, l0, off);
      const float mo1 = __shfl_xor_sync(0xffffffffu, m1, off), lo1 = __shfl_xor_sync(0xffffffffu, l1, off);
      const float n0 = fmaxf(m0, mo0), n1 = fmaxf(m1, mo1);
      l0 = (m0 > -INFINITY ? l0 * exp2f(m0 - n0) : 0.f) + (mo0 > -INFINITY ? lo0 * exp2f(mo0 - n0) : 0.f);
      l1 = (m1 > -INFINITY ? l1 * exp2f(m1 - n1) : 0.f) + (mo1 > -INFINITY ? lo1 * exp2f(mo1 - n1) : 0.f);
      m0 = n0; m1 = n1;
    }
    const float L0 = m0 + log2f(l0), L1 = m1 + log2f(l1);
    if (t == 0) { Lrow[r0 + g] = L0; Lrow[r0 + g + 8] = L1; }
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) acc[nd][0] = acc[nd][1] = acc[nd][2] = acc[nd][3] = 0.f;
#pragma unroll 1
    for (int jt = 0; jt < NTILES; ++jt) {
      const int j0 = 8 * jt, col = j0 + 2 * t;
      vit_score_tile(xa, sK, j0, g, t, c);
      const bool v0 = col < VL, v1 = col + 1 < VL;
      float pr[4];
      pr[0] = v0 ? exp2f(c[0] * V_SCALE_LOG2 - L0) : 0.f;
      pr[1] = v1 ? exp2f(c[1] * V_SCALE_LOG2 - L0) : 0.f;
      pr[2] = v0 ? exp2f(c[2] * V_SCALE_LOG2 - L1) : 0.f;
      pr[3] = v1 ? exp2f(c[3] * V_SCALE_LOG2 - L1) : 0.f;
      vit_accum_tile(pr, sV, j0, g, t, acc);
    }
    float* o = out + row0 * VC + head * VHD;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      if (r0 + g < VL) *reinterpret_cast<float2*>(o + (long long)(r0 + g) * VC + 8 * nd + 2 * t) = make_float2(acc[nd][0], acc[nd][1]);
      if (r0 + g + 8 < VL)
        *reinterpret_cast<float2*>(o + (long long)(r0 + g + 8) * VC + 8 * nd + 2 * t) = make_float2(acc[nd][2], acc[nd][3]);
    }
  }
}

__global__ void __launch_bounds__(256, 1)
vit_attention_bwd_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ AO, const float* __restrict__ dO,
                             float* __restrict__ dqkv, const float* __restrict__ Lin) {
  extern __shared__ __align__(16) float sm_vm[];
  float* sQ = sm_vm;
  float* sK = sQ + MROWS * ML;
  float* sV = sK + MROWS * ML;
  float* sO = sV + MROWS * ML;
  float* sL = sO + MROWS * ML;    // log2 of the row's softmax denominator (base-2 scores)
  float* sD = sL + MROWS;         // delta
  const int frame = blockIdx.x / VHEADS, head = blockIdx.x % VHEADS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long row0 = (long long)frame * VL;
  const float* base = qkv + row0 * (3 * VC) + head * VHD;
  for (int i = threadIdx.x; i < MROWS * (VHD / 4); i += blockDim.x) {
    const int r = i / (VHD / 4), d = (i % (VHD / 4)) * 4;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f), k = q, v = q, o = q;
    if (r < VL) {
      q = *reinterpret_cast<const float4*>(base + (long long)r * (3 * VC) + d);
      k = *reinterpret_cast<const float4*>(base + (long long)r * (3 * VC) + VC + d);
      v = *reinterpret_cast<const float4*>(base + (long long)r * (3 * VC) + 2 * VC + d);
      o = *reinterpret_cast<const float4*>(dO + (row0 + r) * VC + head * VHD + d);
    }
    *reinterpret_cast<float4*>(sQ + r * ML + d) = make_float4(tf32_round(q.x), tf32_round(q.y), tf32_round(q.z), tf32_round(q.w));
    *reinterpret_cast<float4*>(sK + r * ML + d) = make_float4(tf32_round(k.x), tf32_round(k.y), tf32_round(k.z), tf32_round(k.w));
    *reinterpret_cast<float4*>(sV + r * ML + d) = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
    *reinterpret_cast<float4*>(sO + r * ML + d) = make_float4(tf32_round(o.x), tf32_round(o.y), tf32_round(o.z), tf32_round(o.w));
  }
  // delta_i = dO_i . O_i on the unrounded values, a warp per row
  for (int i = warp; i < MROWS; i += 8) {
    float s = 0.f;
    if (i < VL) {
      const float* o = AO + (row0 + i) * VC + head * VHD;
      const float* d = dO + (row0 + i) * VC + head * VHD;
      s = o[lane] * d[lane] + o[lane + 32] * d[lane + 32];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) sD[i] = s;
  }
  __syncthreads();
  uint32_t xa[8][4], ya[8][4];
  float c[4], d[4];
  // ---- phase 0: L_i (kept by the tensor-core forward; recomputed here when the forward was the fp32 kernel)
  if (Lin != nullptr) {
    for (int i = threadIdx.x; i < MROWS; i += blockDim.x) sL[i] = Lin[(long long)blockIdx.x * MROWS + i];
  } else
  for (int rt = warp; rt < MTILES; rt += 8) {
    const int r0 = 16 * rt;
    vit_afrag(sQ, r0, g, t, xa);
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
#pragma unroll 1
    for (int jt = 0; jt < NTILES; ++jt) {
      const int j0 = 8 * jt, col = j0 + 2 * t;
      vit_score_tile(xa, sK, j0, g, t, c);
      const float s00 = col < VL ? c[0] * V_SCALE_LOG2 : -INFINITY, s01 = col + 1 < VL ? c[1] * V_SCALE_LOG2 : -INFINITY;
      const float s10 = col < VL ? c[2] * V_SCALE_LOG2 : -INFINITY, s11 = col + 1 < VL ? c[3] * V_SCALE_LOG2 : -INFINITY;
      const float n0 = fmaxf(m0, fmaxf(s00, s01)), n1 = fmaxf(m1, fmaxf(s10, s11));
      if (n0 > -INFINITY) { l0 = l0 * exp2f(m0 - n0) + exp2f(s00 - n0) + exp2f(s01 - n0); m0 = n0; }
      if (n1 > -INFINITY) { l1 = l1 * exp2f(m1 - n1) + exp2f(s10 - n1) + exp2f(s11 - n1); m1 = n1; }
    }
#pragma unroll
    for (int off = 1; off <= 2; off <<= 1) {      // the 4 lanes of a row hold disjoint columns
      const float mo0 = __shfl_xor_sync(0xffffffffu, m0, off), lo0 = __shfl_xor_sync(0xffffffffu, l0, off);
      const float mo1 = __shfl_xor_sync(0xffffffffu, m1, off), lo1 = __shfl_xor_sync(0xffffffffu, l1, off);
      const float n0 = fmaxf(m0, mo0), n1 = fmaxf(m1, mo1);
      l0 = (m0 > -INFINITY ? l0 * exp2f(m0 - n0) : 0.f) + (mo0 > -INFINITY ? lo0 * exp2f(mo0 - n0) : 0.f);
      l1 = (m1 > -INFINITY ? l1 * exp2f(m1 - n1) : 0.f) + (mo1 > -INFINITY ? lo1 * exp2f(mo1 - n1) : 0.f);
      m0 = n0; m1 = n1;
    }
    if (t == 0) { sL[r0 + g] = m0 + log2f(l0); sL[r0 + g + 8] = m1 + log2f(l1); }
  }
  __syncthreads();
  float* out = dqkv + row0 * (3 * VC) + head * VHD;
  float acc[8][4], acc2[8][4];
  // ---- phase 1: dQ, a warp per 16 query rows
  for (int rt = warp; rt < MTILES; rt += 8) {
    const int r0 = 16 * rt;
    vit_afrag(sQ, r0, g, t, xa);
    vit_afrag(sO, r0, g, t, ya);
    const float L0 = sL[r0 + g], L1 = sL[r0 + g + 8], D0 = sD[r0 + g], D1 = sD[r0 + g + 8];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) acc[nd][0] = acc[nd][1] = acc[nd][2] = acc[nd][3] = 0.f;
#pragma unroll 1
    for (int jt = 0; jt < NTILES; ++jt) {
      const int j0 = 8 * jt, col = j0 + 2 * t;
      vit_score_tile(xa, sK, j0, g, t, c);     // S
      vit_score_tile(ya, sV, j0, g, t, d);     // dP
      const bool v0 = col < VL, v1 = col + 1 < VL;
      float ds[4];
      ds[0] = v0 ? exp2f(c[0] * V_SCALE_LOG2 - L0) * (d[0] - D0) : 0.f;
      ds[1] = v1 ? exp2f(c[1] * V_SCALE_LOG2 - L0) * (d[1] - D0) : 0.f;
      ds[2] = v0 ? exp2f(c[2] * V_SCALE_LOG2 - L1) * (d[2] - D1) : 0.f;
      ds[3] = v1 ? exp2f(c[3] * V_SCALE_LOG2 - L1) * (d[3] - D1) : 0.f;
      vit_accum_tile(ds, sK, j0, g, t, acc);
    }
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      if (r0 + g < VL)
        *reinterpret_cast<float2*>(out + (long long)(r0 + g) * (3 * VC) + 8 * nd + 2 * t) = make_float2(acc[nd][0] * 0.125f, acc[nd][1] * 0.125f);
      if (r0 + g + 8 < VL)
        *reinterpret_cast<float2*>(out + (long long)(r0 + g + 8) * (3 * VC) + 8 * nd + 2 * t) = make_float2(acc[nd][2] * 0.125f, acc[nd][3] * 0.125f);
    }
  }
  // ---- phase 2: dK, dV, a warp per 16 key rows; the tile columns are QUERY rows
  for (int rt = warp; rt < MTILES; rt += 8) {
    const int r0 = 16 * rt;
    vit_afrag(sK, r0, g, t, xa);
    vit_afrag(sV, r0, g, t, ya);
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      acc[nd][0] = acc[nd][1] = acc[nd][2] = acc[nd][3] = 0.f;
      acc2[nd][0] = acc2[nd][1] = acc2[nd][2] = acc2[nd][3] = 0.f;
    }
#pragma unroll 1
    for (int it = 0; it < NTILES; ++it) {
      const int i0 = 8 * it, col = i0 + 2 * t;
      vit_score_tile(xa, sQ, i0, g, t, c);     // S^T: c[row = key][col = query]
      vit_score_tile(ya, sO, i0, g, t, d);     // dP^T
      const bool v0 = col < VL, v1 = col + 1 < VL;
      const float L0 = sL[col], L1 = sL[col + 1], D0 = sD[col], D1 = sD[col + 1];
      float pr[4], ds[4];
      pr[0] = v0 ? exp2f(c[0] * V_SCALE_LOG2 - L0) : 0.f;
      pr[1] = v1 ? exp2f(c[1] * V_SCALE_LOG2 - L1) : 0.f;
      pr[2] = v0 ? exp2f(c[2] * V_SCALE_LOG2 - L0) : 0.f;
      pr[3] = v1 ? exp2f(c[3] * V_SCALE_LOG2 - L1) : 0.f;
      ds[0] = pr[0] * (d[0] - D0); ds[1] = pr[1] * (d[1] - D1); ds[2] = pr[2] * (d[2] - D0); ds[3] = pr[3] * (d[3] - D1);
      vit_accum_tile(ds, sQ, i0, g, t, acc);    // dK += dS^T Q
      vit_accum_tile(pr, sO, i0, g, t, acc2);   // dV += P^T dO
    }
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
      if (r0 + g < VL) {
        float* o = out + (long long)(r0 + g) * (3 * VC) + 8 * nd + 2 * t;
        *reinterpret_cast<float2*>(o + VC) = make_float2(acc[nd][0] * 0.125f, acc[nd][1] * 0.125f);
        *reinterpret_cast<float2*>(o + 2 * VC) = make_float2(acc2[nd][0], acc2[nd][1]);
      }
      if (r0 + g + 8 < VL) {
        float* o = out + (long long)(r0 + g + 8) * (3 * VC) + 8 * nd + 2 * t;
        *reinterpret_cast<float2*>(o + VC) = make_float2(acc[nd][2] * 0.125f, acc[nd][3] * 0.125f);
        *reinterpret_cast<float2*>(o + 2 * VC) = make_float2(acc2[nd][2], acc2[nd][3]);
      }
    }
  }
}

constexpr int TV1_SEQ_MAX = 64;
size_t attn_bwd_smem(int n, int dh) { return (size_t)(4 * n * (dh + 1) + 2 * n * n) * sizeof(float); }

int tv1_alloc(spm_tv1* h, float** p, long long n) {
  SPM_CUDA(cudaMalloc(reinterpret_cast<void**>(p), (size_t)n * sizeof(float)));
  h->allocs.push_back(*p);
  return 0;
}

int tv1_transpose(cudaStream_t st, const float* in, int R, int C, float* out, int ldo) {
  dim3 grid((C + 31) / 32, (ldo + 31) / 32), block(32, 8);
  transpose_pad_kernel<<<grid, block, 0, st>>>(in, R, C, out, ldo);
  TV1_LAUNCH_CHECK();
  return 0;
}
const unsigned long long* g_seed_src = nullptr;   // spm_dropout_seed_source

int tv1_dropout(cudaStream_t st, const float* x, const float* add, float* y, long long n, float p, unsigned long long seed,
                unsigned site) {
  const long long quads = (n + 3) / 4;
  dropout_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, st>>>(x, add, y, n, p, seed, site, g_seed_src);
  TV1_LAUNCH_CHECK();
  return 0;
}
// part: CS_CHUNKS * C floats of scratch for the two-stage form (null, or a short matrix: one stage)
int tv1_colsum(cudaStream_t st, const float* a, const float* b, int R, int C, float* out, float* part = nullptr) {
  if (part != nullptr && b == nullptr && R >= 512) {
    const int rpc = ((R + CS_CHUNKS - 1) / CS_CHUNKS + 7) / 8 * 8, nch = (R + rpc - 1) / rpc;
    dim3 g1((C + 31) / 32, nch), blk(32, 8);
    colsum_part_kernel<<<g1, blk, 0, st>>>(a, R, C, rpc, part);
    TV1_LAUNCH_CHECK();
    colsum_final_kernel<<<(C + 127) / 128, 128, 0, st>>>(part, nch, C, out);
    TV1_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid((C + 31) / 32), block(32, 8);
  colsum_kernel<<<grid, block, 0, st>>>(a, b, R, C, out);
  TV1_LAUNCH_CHECK();
  return 0;
}

// out[M, N] = A[M, K] B[N, K]^T (+ bias) (act) (+ residual); planned and run in one go (the block is not a hot loop)
int tv1_gemm(spm_tv1* h, cudaStream_t st, const float* A, long long lda, const float* B, long long ldb, int M, int N, int K,
             const float* bias, int act, const float* residual, float* out) {
  GemmEpilogue e;
  e.bias = bias; e.act = act; e.residual = residual; e.ldr = N; e.out = out; e.ldo = N;
  GemmOp op;
  const char* err = "";
  if (gemm_plan(&op, h->fp32 ? GEMM_F32_SIMT : GEMM_TF32, A, lda, B, ldb, M, N, K, e, h->sms, &err) ||
      gemm_run(&op, st, &err)) {
    set_error(std::string("tv1 gemm: ") + err);
    return 1;
  }
  return 0;
}

// Backward scratch (recomputed pre-activation, gradient tensors, transposed copies) is shared by every handle of the
// process: a training step runs its blocks one after the other on one stream, and 12 encoder blocks x 3 GB of private
// scratch would be waste.  The per-handle buffers are only what a forward must keep for its backward.
struct Tv1Scratch {
  long long cap_rows = 0, cap_wide = 0, cap_d = 0, cap_i3 = 0;
  float *PRE = nullptr, *dY = nullptr, *dAO = nullptr, *dQKV = nullptr, *dHN = nullptr, *tA = nullptr, *tB = nullptr, *G3 = nullptr,
        *part = nullptr;   // CS_CHUNKS x widest row of column-sum partials
} g_scr;

int tv1_scratch(const spm_tv1* h, long long R) {
  const long long D = h->D, I3 = 3LL * h->inner, wide = std::max<long long>(std::max<long long>(I3, h->mlp), D);
  if (R <= g_scr.cap_rows && wide <= g_scr.cap_wide && D <= g_scr.cap_d && I3 <= g_scr.cap_i3) return 0;
  SPM_CUDA(cudaDeviceSynchronize());
  for (float** p : {&g_scr.PRE, &g_scr.dY, &g_scr.dAO, &g_scr.dQKV, &g_scr.dHN, &g_scr.tA, &g_scr.tB, &g_scr.G3, &g_scr.part})
    if (*p) { cudaFree(*p); *p = nullptr; }
  const long long r = std::max(R, g_scr.cap_rows), w = std::max(wide, g_scr.cap_wide), d = std::max(D, g_scr.cap_d),
                  i3 = std::max(I3, g_scr.cap_i3), rp = (r + 3) / 4 * 4;
  g_scr.cap_rows = g_scr.cap_wide = g_scr.cap_d = g_scr.cap_i3 = 0;
  auto al = [](float** p, long long n) { return cudaMalloc(reinterpret_cast<void**>(p), (size_t)n * sizeof(float)); };
  SPM_CUDA(al(&g_scr.PRE, r * w));
  SPM_CUDA(al(&g_scr.dY, r * d));
  SPM_CUDA(al(&g_scr.dAO, r * i3 / 3));
  SPM_CUDA(al(&g_scr.dQKV, r * i3));
  SPM_CUDA(al(&g_scr.dHN, r * w));      // also dFFH / dPRE [R, mlp]
  SPM_CUDA(al(&g_scr.tA, w * rp));
  SPM_CUDA(al(&g_scr.tB, w * rp));
  SPM_CUDA(al(&g_scr.G3, r * d));
  SPM_CUDA(al(&g_scr.part, (long long)CS_CHUNKS * w));
  g_scr.cap_rows = r; g_scr.cap_wide = w; g_scr.cap_d = d; g_scr.cap_i3 = i3;
  return 0;
}

int tv1_workspace(spm_tv1* h, long long R) {
  if (R <= h->cap_rows) return 0;
  SPM_CUDA(cudaDeviceSynchronize());
  for (float** p : {&h->HN, &h->QKV, &h->AO, &h->Y, &h->FFH, &h->H2, &h->Lsm})
    if (*p) { cudaFree(*p); h->allocs.erase(std::remove(h->allocs.begin(), h->allocs.end(), (void*)*p), h->allocs.end()); *p = nullptr; }
  const long long D = h->D, I = h->inner, M = h->mlp;
  SPM_TRY(tv1_alloc(h, &h->HN, R * D));
  SPM_TRY(tv1_alloc(h, &h->QKV, R * 3 * I));
  SPM_TRY(tv1_alloc(h, &h->AO, R * I));
  SPM_TRY(tv1_alloc(h, &h->Y, R * D));
  SPM_TRY(tv1_alloc(h, &h->FFH, R * M));
  if (h->vit) SPM_TRY(tv1_alloc(h, &h->H2, R * D));
  if (h->vit) SPM_TRY(tv1_alloc(h, &h->Lsm, (R / VL + 1) * (long long)VHEADS * MROWS));
  h->cap_rows = R;
  return 0;
}

int block_create(int D, int heads, int dim_head, int mlp_dim, int precision, bool vit, spm_tv1** out) {
  SPM_CHECK(out != nullptr, "block create: null argument");
  SPM_CHECK(D > 0 && D % 32 == 0 && heads > 0 && dim_head > 0 && dim_head % 32 == 0 && dim_head <= 256 && mlp_dim % 32 == 0,
            "block create: D, dim_head and mlp_dim must be multiples of 32 (dim_head <= 256)");
  SPM_CHECK(precision == 0 || precision == 1, "block create: precision 0 (tf32 products) or 1 (fp32)");
  int ndev = 0;
  SPM_CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0, "block create: no CUDA device -- this library has no CPU path");
  const char* err = "";
  SPM_CHECK(gemm_init(&err) == 0 && gemm2_init(&err) == 0, "block create: gemm_init failed");
  SPM_CHECK(k_seq_attention_init() == 0 && k_vit_attention_f32_init() == 0, "block create: cudaFuncSetAttribute failed");
  SPM_CUDA(cudaFuncSetAttribute(seq_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)attn_bwd_smem(48, 256)));
  SPM_CUDA(cudaFuncSetAttribute(vit_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VIT_ATT_BWD_SMEM));
  SPM_CUDA(cudaFuncSetAttribute(vit_attention_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VIT_ATT_BWD_MMA_SMEM));
  SPM_CUDA(cudaFuncSetAttribute(vit_attention_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VIT_ATT_FWD_MMA_SMEM));
  spm_tv1* h = new spm_tv1();
  h->D = D; h->heads = heads; h->dh = dim_head; h->inner = heads * dim_head; h->mlp = mlp_dim; h->fp32 = precision; h->vit = vit;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, dev);
  *out = h;
  return 0;
}

// weights: wqkv = [q; k; v] rows ([3*inner, D]); ViT: + bqkv, ln2
int block_load(spm_tv1* h, cudaStream_t st, const float* ln_g, const float* ln_b, const float* wq, const float* wk,
               const float* wv, const float* bqkv, const float* wout, const float* bout, const float* ln2_g, const float* ln2_b,
               const float* w0, const float* b0, const float* w3, const float* b3) {
  const long long D = h->D, I = h->inner, M = h->mlp;
  if (!h->loaded) {
    SPM_TRY(tv1_alloc(h, &h->ln_g, D)); SPM_TRY(tv1_alloc(h, &h->ln_b, D));
    SPM_TRY(tv1_alloc(h, &h->wqkv, 3 * I * D)); SPM_TRY(tv1_alloc(h, &h->wout, D * I)); SPM_TRY(tv1_alloc(h, &h->bout, D));
    SPM_TRY(tv1_alloc(h, &h->w0, M * D)); SPM_TRY(tv1_alloc(h, &h->b0, M));
    SPM_TRY(tv1_alloc(h, &h->w3, D * M)); SPM_TRY(tv1_alloc(h, &h->b3, D));
    SPM_TRY(tv1_alloc(h, &h->wqkvT, D * 3 * I)); SPM_TRY(tv1_alloc(h, &h->woutT, I * D));
    SPM_TRY(tv1_alloc(h, &h->w0T, D * M)); SPM_TRY(tv1_alloc(h, &h->w3T, M * D));
    if (h->vit) { SPM_TRY(tv1_alloc(h, &h->ln2_g, D)); SPM_TRY(tv1_alloc(h, &h->ln2_b, D)); SPM_TRY(tv1_alloc(h, &h->bqkv, 3 * I)); }
    h->loaded = true;
  }
  auto cp = [&](float* dst, const float* src, long long n) {
    return cudaMemcpyAsync(dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
  };
  SPM_CUDA(cp(h->ln_g, ln_g, D)); SPM_CUDA(cp(h->ln_b, ln_b, D));
  SPM_CUDA(cp(h->wqkv, wq, I * D)); SPM_CUDA(cp(h->wqkv + I * D, wk, I * D)); SPM_CUDA(cp(h->wqkv + 2 * I * D, wv, I * D));
  SPM_CUDA(cp(h->wout, wout, D * I)); SPM_CUDA(cp(h->bout, bout, D));
  SPM_CUDA(cp(h->w0, w0, M * D)); SPM_CUDA(cp(h->b0, b0, M)); SPM_CUDA(cp(h->w3, w3, D * M)); SPM_CUDA(cp(h->b3, b3, D));
  if (h->vit) { SPM_CUDA(cp(h->ln2_g, ln2_g, D)); SPM_CUDA(cp(h->ln2_b, ln2_b, D)); SPM_CUDA(cp(h->bqkv, bqkv, 3 * I)); }
  // transposes for the dgrad GEMMs (B operand = W^T, K-contiguous): [rows, cols] -> [cols, rows]
  SPM_TRY(tv1_transpose(st, h->wqkv, (int)(3 * I), (int)D, h->wqkvT, (int)(3 * I)));
  SPM_TRY(tv1_transpose(st, h->wout, (int)D, (int)I, h->woutT, (int)D));
  SPM_TRY(tv1_transpose(st, h->w0, (int)M, (int)D, h->w0T, (int)M));
  SPM_TRY(tv1_transpose(st, h->w3, (int)D, (int)M, h->w3T, (int)D));
  return 0;
}

int block_forward(spm_tv1* h, cudaStream_t st, const float* x, int n_seq, int seq_len, float* out) {
  const int D = h->D, I = h->inner, M = h->mlp, R = n_seq * seq_len;
  SPM_TRY(tv1_workspace(h, R));
  h->x = x; h->B = n_seq; h->n = seq_len;
  TV1_KERNEL(k_layernorm(st, x, D, R, D, h->ln_g, h->ln_b, nullptr, 0, h->HN, nullptr, D));
  SPM_TRY(tv1_gemm(h, st, h->HN, D, h->wqkv, D, R, 3 * I, D, h->vit ? h->bqkv : nullptr, ACT_NONE, nullptr, h->QKV));
  static const bool simt_attn = [] { const char* e = getenv("SPM_TRAIN_ATTN"); return e != nullptr && std::string(e) == "simt"; }();
  h->have_L = false;
  if (h->vit && !h->fp32 && !simt_attn) {      // tf32 path: tensor cores; keeps L for the backward
    vit_attention_fwd_mma_kernel<<<n_seq * VHEADS, 256, VIT_ATT_FWD_MMA_SMEM, st>>>(h->QKV, h->AO, h->Lsm);
    TV1_LAUNCH_CHECK();
    h->have_L = true;
  } else if (h->vit) TV1_KERNEL(k_vit_attention_f32(st, h->QKV, h->AO, n_seq));
  else TV1_KERNEL(k_seq_attention(st, h->QKV, h->AO, n_seq, seq_len, 1, 0, seq_len, 0, 0, h->heads, h->dh));
  // nn.Dropout sites of the block (myRes.py:961-962 to_out, :990,992 FeedForward): 0 = after to_out, 1 = after GELU, 2 = after
  // net.3; each is applied BEFORE the residual add, so with p > 0 the residual leaves the GEMM epilogue
  h->fwd_p_atte = h->p_atte; h->fwd_p_ffn = h->p_ffn; h->fwd_seed = h->seed;
  const long long nD = (long long)R * D, nM = (long long)R * M;
  if (h->p_atte > 0.f) {
    SPM_TRY(tv1_gemm(h, st, h->AO, I, h->wout, I, R, D, I, h->bout, ACT_NONE, nullptr, h->Y));
    SPM_TRY(tv1_dropout(st, h->Y, x, h->Y, nD, h->p_atte, h->seed, 0));
  } else {
    SPM_TRY(tv1_gemm(h, st, h->AO, I, h->wout, I, R, D, I, h->bout, ACT_NONE, x, h->Y));
  }
  const float* h2 = h->Y;          // input of the MLP: y itself, or ln_2(y) in the encoder block
  if (h->vit) {
    TV1_KERNEL(k_layernorm(st, h->Y, D, R, D, h->ln2_g, h->ln2_b, nullptr, 0, h->H2, nullptr, D));
    h2 = h->H2;
  }
  const int act = h->vit ? ACT_QUICKGELU : ACT_GELU_ERF;
  SPM_TRY(tv1_gemm(h, st, h2, D, h->w0, D, R, M, D, h->b0, act, nullptr, h->FFH));
  if (h->p_ffn > 0.f) {
    SPM_TRY(tv1_dropout(st, h->FFH, nullptr, h->FFH, nM, h->p_ffn, h->seed, 1));
    SPM_TRY(tv1_gemm(h, st, h->FFH, M, h->w3, M, R, D, M, h->b3, ACT_NONE, nullptr, out));
    SPM_TRY(tv1_dropout(st, out, h->Y, out, nD, h->p_ffn, h->seed, 2));
  } else {
    SPM_TRY(tv1_gemm(h, st, h->FFH, M, h->w3, M, R, D, M, h->b3, ACT_NONE, h->Y, out));
  }
  return 0;
}

// g_qkv: [3*inner, D]; g_bqkv / g_ln2_*: ViT only
int block_backward(spm_tv1* h, cudaStream_t st, const float* grad_out, float* grad_x, float* g_ln_g, float* g_ln_b,
                   float* g_qkv, float* g_bqkv, float* g_wout, float* g_bout, float* g_ln2_g, float* g_ln2_b, float* g_w0,
                   float* g_b0, float* g_w3, float* g_b3) {
  const int D = h->D, I = h->inner, M = h->mlp, n = h->n, R = h->B * h->n, Rp = (R + 3) / 4 * 4;
  SPM_TRY(tv1_scratch(h, R));
  Tv1Scratch& s = g_scr;
  float* dF = s.dHN;   // [R, mlp]: dFFH, then dPRE in place
  const long long nD = (long long)R * D, nM = (long long)R * M;
  const float pa = h->fwd_p_atte, pf = h->fwd_p_ffn;
  const float* h2 = h->vit ? h->H2 : h->Y;
  // ---- out = drop2(W3 drop1(act(PRE)) + b3) + y
  const float* g3 = grad_out;          // gradient behind the dropout of site 2 (the residual branch keeps grad_out itself)
  if (pf > 0.f) { SPM_TRY(tv1_dropout(st, grad_out, nullptr, s.G3, nD, pf, h->fwd_seed, 2)); g3 = s.G3; }
  SPM_TRY(tv1_colsum(st, g3, nullptr, R, D, g_b3, s.part));
  SPM_TRY(tv1_transpose(st, g3, R, D, s.tA, Rp));
  SPM_TRY(tv1_transpose(st, h->FFH, R, M, s.tB, Rp));
  SPM_TRY(tv1_gemm(h, st, s.tA, Rp, s.tB, Rp, D, M, Rp, nullptr, ACT_NONE, nullptr, g_w3));               // dW3 = dOut^T drop(act(PRE))
  SPM_TRY(tv1_gemm(h, st, g3, D, h->w3T, D, R, M, D, nullptr, ACT_NONE, nullptr, dF));                    // dFFH = dOut W3
  if (pf > 0.f) SPM_TRY(tv1_dropout(st, dF, nullptr, dF, nM, pf, h->fwd_seed, 1));
  SPM_TRY(tv1_gemm(h, st, h2, D, h->w0, D, R, M, D, h->b0, ACT_NONE, nullptr, s.PRE));                    // PRE recomputed
  if (h->vit) quickgelu_bwd_kernel<<<(unsigned)((nM + 255) / 256), 256, 0, st>>>(dF, s.PRE, nM);
  else gelu_bwd_kernel<<<(unsigned)((nM + 255) / 256), 256, 0, st>>>(dF, s.PRE, nM);
  TV1_LAUNCH_CHECK();
  SPM_TRY(tv1_colsum(st, dF, nullptr, R, M, g_b0, s.part));
  SPM_TRY(tv1_transpose(st, dF, R, M, s.tA, Rp));
  SPM_TRY(tv1_transpose(st, h2, R, D, s.tB, Rp));
  SPM_TRY(tv1_gemm(h, st, s.tA, Rp, s.tB, Rp, M, D, Rp, nullptr, ACT_NONE, nullptr, g_w0));               // dW0 = dPRE^T h2
  if (h->vit) {
    // h2 = ln_2(y): dY = LN-backward(dPRE W0) + dOut
    SPM_TRY(tv1_gemm(h, st, dF, M, h->w0T, M, R, D, M, nullptr, ACT_NONE, nullptr, s.PRE));               // dH2 (PRE is free again)
    SPM_TRY(tv1_colsum(st, s.PRE, nullptr, R, D, g_ln2_b, s.part));
    ln_bwd_kernel<<<(R + 7) / 8, 256, 0, st>>>(h->Y, s.PRE, h->ln2_g, grad_out, R, D, s.dY, s.tA);        // tA := yhat o dH2
    TV1_LAUNCH_CHECK();
    SPM_TRY(tv1_colsum(st, s.tA, nullptr, R, D, g_ln2_g, s.part));
  } else {
    SPM_TRY(tv1_gemm(h, st, dF, M, h->w0T, M, R, D, M, nullptr, ACT_NONE, grad_out, s.dY));               // dY = dPRE W0 + dOut
  }
  // ---- y = drop0(Wout ao + b_out) + x
  const float* gy = s.dY;              // gradient behind the dropout of site 0 (the residual branch keeps dY itself)
  if (pa > 0.f) { SPM_TRY(tv1_dropout(st, s.dY, nullptr, s.G3, nD, pa, h->fwd_seed, 0)); gy = s.G3; }
  SPM_TRY(tv1_colsum(st, gy, nullptr, R, D, g_bout, s.part));
  SPM_TRY(tv1_transpose(st, gy, R, D, s.tA, Rp));
  SPM_TRY(tv1_transpose(st, h->AO, R, I, s.tB, Rp));
  SPM_TRY(tv1_gemm(h, st, s.tA, Rp, s.tB, Rp, D, I, Rp, nullptr, ACT_NONE, nullptr, g_wout));             // dWout = dY^T ao
  SPM_TRY(tv1_gemm(h, st, gy, D, h->woutT, D, R, I, D, nullptr, ACT_NONE, nullptr, s.dAO));               // dAO = dY Wout
  // ---- attention
  if (h->vit) {
    // tf32 path: tensor cores (mma.sync); exact-fp32 path (and SPM_TRAIN_ATTN=simt, the cross-check): the FFMA kernel
    static const bool simt = [] { const char* e = getenv("SPM_TRAIN_ATTN"); return e != nullptr && std::string(e) == "simt"; }();
    if (h->fp32 || simt)
      vit_attention_bwd_kernel<<<h->B * VHEADS, 256, VIT_ATT_BWD_SMEM, st>>>(h->QKV, s.dAO, s.dQKV);
    else
      vit_attention_bwd_mma_kernel<<<h->B * VHEADS, 256, VIT_ATT_BWD_MMA_SMEM, st>>>(h->QKV, h->AO, s.dAO, s.dQKV,
                                                                                        h->have_L ? h->Lsm : nullptr);
    TV1_LAUNCH_CHECK();
  } else {
    SPM_CHECK(n <= TV1_SEQ_MAX && attn_bwd_smem(n, h->dh) <= attn_bwd_smem(48, 256), "block backward: sequence too long");
    dim3 grid(h->B, h->heads);
    seq_attention_bwd_kernel<<<grid, 256, attn_bwd_smem(n, h->dh), st>>>(h->QKV, s.dAO, s.dQKV, n, h->heads, h->dh);
    TV1_LAUNCH_CHECK();
  }
  // ---- q, k, v = W{q,k,v} LN(x) (+ b)
  if (h->vit) SPM_TRY(tv1_colsum(st, s.dQKV, nullptr, R, 3 * I, g_bqkv, s.part));
  SPM_TRY(tv1_transpose(st, s.dQKV, R, 3 * I, s.tA, Rp));
  SPM_TRY(tv1_transpose(st, h->HN, R, D, s.tB, Rp));
  SPM_TRY(tv1_gemm(h, st, s.tA, Rp, s.tB, Rp, 3 * I, D, Rp, nullptr, ACT_NONE, nullptr, g_qkv));          // [dWq; dWk; dWv]
  SPM_TRY(tv1_gemm(h, st, s.dQKV, 3 * I, h->wqkvT, 3 * I, R, D, 3 * I, nullptr, ACT_NONE, nullptr, s.dHN));   // dLN = dQKV Wqkv
  // ---- LayerNorm + the residual branch (dY)
  SPM_TRY(tv1_colsum(st, s.dHN, nullptr, R, D, g_ln_b, s.part));
  ln_bwd_kernel<<<(R + 7) / 8, 256, 0, st>>>(h->x, s.dHN, h->ln_g, s.dY, R, D, grad_x, s.tA);             // tA := xhat o dLN
  TV1_LAUNCH_CHECK();
  SPM_TRY(tv1_colsum(st, s.tA, nullptr, R, D, g_ln_g, s.part));
  return 0;
}

}  // namespace
}  // namespace spm

using namespace spm;

extern "C" {

int spm_tv1_create(int D, int heads, int dim_head, int mlp_dim, int precision, spm_tv1** out) {
  return block_create(D, heads, dim_head, mlp_dim, precision, false, out);
}

int spm_tv1_destroy(spm_tv1* h) {
  if (h == nullptr) return 0;
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return 0;
}

int spm_tv1_load_weights(spm_tv1* h, void* stream, const float* ln_g, const float* ln_b, const float* wq, const float* wk,
                         const float* wv, const float* wout, const float* bout, const float* w0, const float* b0,
                         const float* w3, const float* b3) {
  SPM_CHECK(h && ln_g && ln_b && wq && wk && wv && wout && bout && w0 && b0 && w3 && b3, "spm_tv1_load_weights: null argument");
  SPM_CHECK(!h->vit, "spm_tv1_load_weights: this handle is an encoder block (spm_vitblock_load_weights)");
  return block_load(h, (cudaStream_t)stream, ln_g, ln_b, wq, wk, wv, nullptr, wout, bout, nullptr, nullptr, w0, b0, w3, b3);
}

int spm_tv1_set_dropout(spm_tv1* h, float p_atte, float p_ffn, unsigned long long seed) {
  SPM_CHECK(h != nullptr, "spm_tv1_set_dropout: null handle");
  SPM_CHECK(p_atte >= 0.f && p_atte < 1.f && p_ffn >= 0.f && p_ffn < 1.f, "spm_tv1_set_dropout: probabilities in [0, 1)");
  h->p_atte = p_atte; h->p_ffn = p_ffn; h->seed = seed;
  return 0;
}

int spm_dropout_seed_source(const unsigned long long* device_counter) {
  g_seed_src = device_counter;
  return 0;
}

int spm_dropout(void* stream, const float* x, long long n, float p, unsigned long long seed, unsigned site, float* y) {
  SPM_CHECK(x && y && n > 0, "spm_dropout: null argument");
  SPM_CHECK(p >= 0.f && p < 1.f, "spm_dropout: p in [0, 1)");
  return tv1_dropout((cudaStream_t)stream, x, nullptr, y, n, p, seed, site);
}

int spm_tv1_forward(spm_tv1* h, void* stream, const float* x, int n_seq, int seq_len, float* out) {
  SPM_CHECK(h && x && out, "spm_tv1_forward: null argument");
  SPM_CHECK(h->loaded && !h->vit, "spm_tv1_forward: weights not loaded");
  SPM_CHECK(n_seq > 0 && seq_len > 0 && seq_len <= 48, "spm_tv1_forward: 1..48 tokens per sequence");
  return block_forward(h, (cudaStream_t)stream, x, n_seq, seq_len, out);
}

int spm_tv1_backward(spm_tv1* h, void* stream, const float* grad_out, float* grad_x, float* g_ln_g, float* g_ln_b,
                     float* g_wq, float* g_wk, float* g_wv, float* g_wout, float* g_bout, float* g_w0, float* g_b0,
                     float* g_w3, float* g_b3) {
  SPM_CHECK(h && grad_out && grad_x && g_ln_g && g_ln_b && g_wq && g_wk && g_wv && g_wout && g_bout && g_w0 && g_b0 && g_w3 &&
                g_b3, "spm_tv1_backward: null argument");
  SPM_CHECK(h->x != nullptr && !h->vit, "spm_tv1_backward: no forward to differentiate");
  SPM_CHECK(g_wk == g_wq + (long long)h->inner * h->D && g_wv == g_wk + (long long)h->inner * h->D,
            "spm_tv1_backward: the q / k / v weight gradients must be three consecutive [inner, D] blocks");
  return block_backward(h, (cudaStream_t)stream, grad_out, grad_x, g_ln_g, g_ln_b, g_wq, nullptr, g_wout, g_bout, nullptr, nullptr,
                        g_w0, g_b0, g_w3, g_b3);
}

/* ---- frame-encoder residual block (models/clip_fsar.py:622-643), 197 tokens per frame ---- */
int spm_vitblock_create(int precision, spm_tv1** out) { return block_create(VC, VHEADS, VHD, 4 * VC, precision, true, out); }

int spm_vitblock_load_weights(spm_tv1* h, void* stream, const float* ln1_g, const float* ln1_b, const float* in_proj_w,
                              const float* in_proj_b, const float* out_w, const float* out_b, const float* ln2_g,
                              const float* ln2_b, const float* fc_w, const float* fc_b, const float* proj_w, const float* proj_b) {
  SPM_CHECK(h && ln1_g && ln1_b && in_proj_w && in_proj_b && out_w && out_b && ln2_g && ln2_b && fc_w && fc_b && proj_w && proj_b,
            "spm_vitblock_load_weights: null argument");
  SPM_CHECK(h->vit, "spm_vitblock_load_weights: not an encoder-block handle");
  const long long ID = (long long)h->inner * h->D;
  return block_load(h, (cudaStream_t)stream, ln1_g, ln1_b, in_proj_w, in_proj_w + ID, in_proj_w + 2 * ID, in_proj_b, out_w, out_b,
                    ln2_g, ln2_b, fc_w, fc_b, proj_w, proj_b);
}

int spm_vitblock_forward(spm_tv1* h, void* stream, const float* x, int n_frames, float* out) {
  SPM_CHECK(h && x && out, "spm_vitblock_forward: null argument");
  SPM_CHECK(h->loaded && h->vit, "spm_vitblock_forward: weights not loaded");
  SPM_CHECK(n_frames > 0, "spm_vitblock_forward: no frames");
  return block_forward(h, (cudaStream_t)stream, x, n_frames, VL, out);
}

int spm_vitblock_backward(spm_tv1* h, void* stream, const float* grad_out, float* grad_x, float* g_ln1_g, float* g_ln1_b,
                          float* g_in_proj_w, float* g_in_proj_b, float* g_out_w, float* g_out_b, float* g_ln2_g, float* g_ln2_b,
                          float* g_fc_w, float* g_fc_b, float* g_proj_w, float* g_proj_b) {
  SPM_CHECK(h && grad_out && grad_x && g_ln1_g && g_ln1_b && g_in_proj_w && g_in_proj_b && g_out_w && g_out_b && g_ln2_g &&
                g_ln2_b && g_fc_w && g_fc_b && g_proj_w && g_proj_b, "spm_vitblock_backward: null argument");
  SPM_CHECK(h->x != nullptr && h->vit, "spm_vitblock_backward: no forward to differentiate");
  return block_backward(h, (cudaStream_t)stream, grad_out, grad_x, g_ln1_g, g_ln1_b, g_in_proj_w, g_in_proj_b, g_out_w, g_out_b,
                        g_ln2_g, g_ln2_b, g_fc_w, g_fc_b, g_proj_w, g_proj_b);
}

/* ---- nn.LayerNorm (eps 1e-5) forward / backward over rows of C ---- */
int spm_layernorm_forward(void* stream, const float* x, int rows, int C, const float* gamma, const float* beta, float* y) {
  SPM_CHECK(x && gamma && beta && y && rows > 0 && C > 0, "spm_layernorm_forward: null argument");
  TV1_KERNEL(k_layernorm((cudaStream_t)stream, x, C, rows, C, gamma, beta, nullptr, 0, y, nullptr, C));
  return 0;
}

int spm_layernorm_backward(void* stream, const float* x, const float* dy, const float* gamma, int rows, int C, float* dx,
                           float* dgamma, float* dbeta, float* workspace) {
  SPM_CHECK(x && dy && gamma && dx && dgamma && dbeta && workspace && rows > 0 && C > 0, "spm_layernorm_backward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  SPM_TRY(tv1_colsum(st, dy, nullptr, rows, C, dbeta, workspace + (long long)rows * C));
  ln_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, dy, gamma, nullptr, rows, C, dx, workspace);   // workspace [rows, C] := xhat o dy, then 64 x C partials
  TV1_LAUNCH_CHECK();
  SPM_TRY(tv1_colsum(st, workspace, nullptr, rows, C, dgamma, workspace + (long long)rows * C));
  return 0;
}

long long spm_linear_backward_workspace(int M, int N, int K) {
  const long long Mp = (M + 3) / 4 * 4;
  return (long long)M * N + (long long)N * Mp + (long long)K * Mp + (long long)K * N + (long long)CS_CHUNKS * N;
}

int spm_linear_backward(void* stream, int precision, const float* x, const float* W, const float* bias, const float* y,
                        const float* dy, int M, int N, int K, int act, float slope, float* dx, float* dW, float* db,
                        float* workspace, long long workspace_floats) {
  SPM_CHECK(x && W && dy && dW && workspace, "spm_linear_backward: null argument");
  SPM_CHECK(M > 0 && N > 0 && K > 0 && N % 32 == 0 && K % 32 == 0, "spm_linear_backward: N and K must be multiples of 32");
  SPM_CHECK(precision == 0 || precision == 1, "spm_linear_backward: precision 0 (tf32 products) or 1 (fp32)");
  SPM_CHECK(act == ACT_NONE || act == ACT_LEAKY || act == ACT_SIGMOID || act == ACT_GELU_ERF,
            "spm_linear_backward: activation none / GELU(erf) / LeakyReLU / sigmoid");
  SPM_CHECK(act == ACT_NONE || act == ACT_GELU_ERF || y != nullptr, "spm_linear_backward: the layer's output is needed");
  SPM_CHECK(workspace_floats >= spm_linear_backward_workspace(M, N, K), "spm_linear_backward: workspace too small");
  const char* err = "";
  SPM_CHECK(gemm_init(&err) == 0 && gemm2_init(&err) == 0, "spm_linear_backward: gemm_init failed");
  cudaStream_t st = (cudaStream_t)stream;
  spm_tv1 cfg;   // carries the precision and SM count into tv1_gemm
  cfg.fp32 = precision;
  SPM_TRY(device_sm_count(&cfg.sms));
  const int Mp = (M + 3) / 4 * 4;
  float* g = workspace;
  float* gT = g + (long long)M * N;
  float* xT = gT + (long long)N * Mp;
  float* WT = xT + (long long)K * Mp;
  float* part = WT + (long long)K * N;
  const float* gr = dy;
  if (act != ACT_NONE) {
    if (act == ACT_GELU_ERF) SPM_TRY(tv1_gemm(&cfg, st, x, K, W, K, M, N, K, bias, ACT_NONE, nullptr, g));   // pre-activation
    const long long nel = (long long)M * N;
    act_bwd_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, st>>>(g, dy, y, act, slope, nel);
    TV1_LAUNCH_CHECK();
    gr = g;
  }
  if (db != nullptr) SPM_TRY(tv1_colsum(st, gr, nullptr, M, N, db, part));
  SPM_TRY(tv1_transpose(st, gr, M, N, gT, Mp));
  SPM_TRY(tv1_transpose(st, x, M, K, xT, Mp));
  SPM_TRY(tv1_gemm(&cfg, st, gT, Mp, xT, Mp, N, K, Mp, nullptr, ACT_NONE, nullptr, dW));          // dW = g^T x
  if (dx != nullptr) {
    SPM_TRY(tv1_transpose(st, W, N, K, WT, N));
    SPM_TRY(tv1_gemm(&cfg, st, gr, N, WT, N, M, K, N, nullptr, ACT_NONE, nullptr, dx));           // dx = g W
  }
  return 0;
}

}  // extern "C"
