// Metric-head kernels around the head GEMMs (HSMR motion path, SPM gating / token assembly, PADM sequences,
// class prototypes).  fp32 throughout, float4-vectorised, no host synchronisation: class indices are resolved
// on the device from the float labels (the reference does torch.unique + nonzero host syncs,
// models/model_clipspm.py:133,231,277 + models/myRes.py:730-739).
#include "head_kernels.cuh"
#include "head_device.cuh"
#include "profile.cuh"

namespace spm {

#define SPM_LAUNCH_CHECK()                                   \
  do {                                                       \
    cudaError_t _e = cudaGetLastError();                     \
    if (_e != cudaSuccess) return (int)_e;                   \
    count_launch();                                          \
  } while (0)

// ------------------------------------------------------------------------------------------------------
// HSMR motion path (model_clipspm.py:169-191)
// ------------------------------------------------------------------------------------------------------
// out[(v*T+t), kk*D + c] = x[v, t+kk-1, c] (zero outside [0,T)): Conv1d(k=3,pad=1) as one GEMM with K = 3D
__global__ void temporal_im2col_kernel(const float* __restrict__ x, long long vid_stride, int V, int T, int D,
                                       float* __restrict__ out) {
  const int d4 = D / 4;
  const long long n = (long long)V * T * 3 * d4;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % d4);
  const int kk = (int)((i / d4) % 3);
  const int t = (int)((i / (3LL * d4)) % T);
  const long long v = i / (3LL * d4 * T);
  const int ts = t + kk - 1;
  float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ts >= 0 && ts < T) val = *reinterpret_cast<const float4*>(x + v * vid_stride + (long long)ts * D + c * 4);
  reinterpret_cast<float4*>(out)[i] = val;
}
int k_temporal_im2col(cudaStream_t st, const float* x, long long vid_stride, int V, int T, int D, float* out) {
  const long long n = (long long)V * T * 3 * (D / 4);
  if (n <= 0) return 0;
  temporal_im2col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, vid_stride, V, T, D, out);
  SPM_LAUNCH_CHECK();
  return 0;
}

// motion[v, c] = mean_{t<T-1} 0.5 * ((conv[v,t+1,c] - x[v,t,c]) + (conv[v,t,c] - x[v,t+1,c]))
__global__ void motion_reduce_kernel(const float* __restrict__ conv, const float* __restrict__ x, long long vid_stride,
                                     int V, int T, int D, float* __restrict__ out) {
  const int d4 = D / 4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)V * d4) return;
  const int c = (int)(i % d4);
  const long long v = i / d4;
  const float4* cv = reinterpret_cast<const float4*>(conv + v * T * D) + c;
  const float4* xv = reinterpret_cast<const float4*>(x + v * vid_stride) + c;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 c0 = cv[0], x0 = xv[0];
  for (int t = 0; t + 1 < T; ++t) {
    const float4 c1 = cv[(long long)(t + 1) * d4], x1 = xv[(long long)(t + 1) * d4];
    acc.x += 0.5f * ((c1.x - x0.x) + (c0.x - x1.x));
    acc.y += 0.5f * ((c1.y - x0.y) + (c0.y - x1.y));
    acc.z += 0.5f * ((c1.z - x0.z) + (c0.z - x1.z));
    acc.w += 0.5f * ((c1.w - x0.w) + (c0.w - x1.w));
    c0 = c1; x0 = x1;
  }
  const float inv = 1.f / (float)(T - 1);
  reinterpret_cast<float4*>(out + v * D)[c] = f4_scale(acc, inv);
}
int k_motion_reduce(cudaStream_t st, const float* conv, const float* x, long long vid_stride, int V, int T, int D,
                    float* out) {
  const long long n = (long long)V * (D / 4);
  if (n <= 0) return 0;
  motion_reduce_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(conv, x, vid_stride, V, T, D, out);
  SPM_LAUNCH_CHECK();
  return 0;
}

// mo_dist (model_clipspm.py:202-205, :341-346): per episode
//   dists[e] = mo_alpha1 * ( mean_q |new_m[q] - tok2[q]|^2 + mean_s |new_m[s] - tok2[s]|^2 )
// new_m [E*N, D]; tok2 = refined motion token = row 0 of each sequence of the `mo` se_te call (stride tok_stride)
__global__ void mo_dist_kernel(const float* __restrict__ new_m, const float* __restrict__ z_tok, long long tok_stride,
                               int S, int Q, int D, const float* __restrict__ mo_alpha1, float* __restrict__ dists) {
  __shared__ float red[32];
  const int e = blockIdx.x, N = S + Q;
  float ss = 0.f, sq = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* a = new_m + ((long long)e * N + n) * D;
    const float* b = z_tok + ((long long)e * N + n) * tok_stride;
    float acc = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      const float df = a[d] - b[d];
      acc += df * df;
    }
    if (n < S) ss += acc; else sq += acc;
  }
  ss = block_sum(ss, red);
  sq = block_sum(sq, red);
  if (threadIdx.x == 0) dists[e] = mo_alpha1[0] * (sq / (float)Q + ss / (float)S);
}
int k_mo_dist(cudaStream_t st, const float* new_m, const float* z_tok, long long tok_stride, int E, int S, int Q, int D,
              const float* mo_alpha1, float* dists) {
  mo_dist_kernel<<<E, 256, 0, st>>>(new_m, z_tok, tok_stride, S, Q, D, mo_alpha1, dists);
  SPM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// SPM tokens (model_clipspm.py:120-121, 213-216, 371-378)
//   tok_b[e, s]   = text[real_support[e,s]]                                   (support: the real class prompt)
//   tt_in[e*Q+q]  = mean(text rows of the episode's targets and supports) * mean_{t,d}(qu[e,q])
//                   (input of token_tr's FeedForward; its output is the query token of the `sem` se_te call)
// X: [E, N, T, D] with the S support videos first.  tok_b: [E*N, D] (rows e*N+s written here; rows of queries
// are written later by the token_tr GEMM epilogue).
// ------------------------------------------------------------------------------------------------------
__global__ void token_prepare_kernel(const float* __restrict__ text, int n_cls, const float* __restrict__ real_support,
                                     const float* __restrict__ real_target, const float* __restrict__ X, int S, int Q,
                                     int T, int D, float* __restrict__ tok_b, float* __restrict__ tt_in,
                                     int* __restrict__ err_flag) {
  extern __shared__ float sm_tp[];  // token[D] + red[32] + qmean[Q]
  float* token = sm_tp;
  float* red = token + D;
  float* qmean = red + 32;
  const int e = blockIdx.x, N = S + Q;
  // a class id outside the text table (the reference would raise an IndexError, model_clipspm.py:116-121): flag the
  // call (NaN logits + an error from the host entry points) and read row 0 instead of out of bounds
  if (threadIdx.x == 0) {
    bool bad = false;
    for (int q = 0; q < Q; ++q) { const float r = real_target[e * Q + q]; bad |= !(r >= 0.f && r < (float)n_cls); }
    for (int s = 0; s < S; ++s) { const float r = real_support[e * S + s]; bad |= !(r >= 0.f && r < (float)n_cls); }
    if (bad && err_flag != nullptr) atomicExch(err_flag, 2);
  }
  auto row = [n_cls](float r) { return (r >= 0.f && r < (float)n_cls) ? (long long)r : 0LL; };
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    // torch.concat([target_context_support, context_support]).mean(0): targets first (summation order)
    for (int q = 0; q < Q; ++q) acc += text[row(real_target[e * Q + q]) * D + d];
    for (int s = 0; s < S; ++s) {
      const float v = text[row(real_support[e * S + s]) * D + d];
      acc += v;
      tok_b[((long long)e * N + s) * D + d] = v;
    }
    token[d] = acc / (float)(S + Q);
  }
  for (int q = 0; q < Q; ++q) {
    const float* xq = X + ((long long)e * N + S + q) * T * D;
    float acc = 0.f;
    for (int i = threadIdx.x; i < T * D; i += blockDim.x) acc += xq[i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) qmean[q] = acc / (float)(T * D);
  }
  __syncthreads();
  for (int q = 0; q < Q; ++q)
    for (int d = threadIdx.x; d < D; d += blockDim.x)
      tt_in[((long long)e * Q + q) * D + d] = token[d] * qmean[q];
}
int k_token_prepare(cudaStream_t st, const float* text, int n_cls, const float* real_support, const float* real_target,
                    const float* X, int E, int S, int Q, int T, int D, float* tok_b, float* tt_in, int* err_flag) {
  const size_t smem = (size_t)(D + 32 + Q) * sizeof(float);
  token_prepare_kernel<<<E, 256, smem, st>>>(text, n_cls, real_support, real_target, X, S, Q, T, D, tok_b, tt_in,
                                             err_flag);
  SPM_LAUNCH_CHECK();
  return 0;
}

// se_te mixing (model_clipspm.py:302-309) for `n_calls` se_te calls sharing the same frames X:
//   seq[c, v, 0]     = tok[c, v]
//   seq[c, v, 1 + t] = alpha * tok[c, v] * gate_text[c, v] + X[v, t] * gate_vision[v, t]
__global__ void seq_build_kernel(const float* __restrict__ tok, const float* __restrict__ gt,
                                 const float* __restrict__ X, const float* __restrict__ gv, int n_calls, long long V,
                                 int T, int D, float alpha, float* __restrict__ seq) {
  const int d4 = D / 4;
  const long long n = (long long)n_calls * V * (T + 1) * d4;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % d4);
  const int tt = (int)((i / d4) % (T + 1));
  const long long cv = i / ((long long)d4 * (T + 1));  // call * V + v
  const long long v = cv % V;
  const float4 tk = reinterpret_cast<const float4*>(tok + cv * D)[c];
  float4 o = tk;
  if (tt > 0) {
    const float4 g = reinterpret_cast<const float4*>(gt + cv * D)[c];
    const float4 xv = reinterpret_cast<const float4*>(X + (v * T + (tt - 1)) * D)[c];
    const float4 gvv = reinterpret_cast<const float4*>(gv + (v * T + (tt - 1)) * D)[c];
    o.x = tk.x * g.x * alpha + xv.x * gvv.x;
    o.y = tk.y * g.y * alpha + xv.y * gvv.y;
    o.z = tk.z * g.z * alpha + xv.z * gvv.z;
    o.w = tk.w * g.w * alpha + xv.w * gvv.w;
  }
  reinterpret_cast<float4*>(seq)[i] = o;
}
int k_seq_build(cudaStream_t st, const float* tok, const float* gt, const float* X, const float* gv, int n_calls,
                int V, int T, int D, float alpha, float* seq) {
  const long long n = (long long)n_calls * V * (T + 1) * (D / 4);
  if (n <= 0) return 0;
  seq_build_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tok, gt, X, gv, n_calls, V, T, D, alpha, seq);
  SPM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// short-sequence multi-head attention, fp32 (models/myRes.py:964-982): softmax(q k^T * dh^-1/2) v
// One CTA per (sequence, head).  K (row stride dh+1: conflict-free when lanes walk different keys) and V live in
// shared memory; a warp owns a query row: lanes = keys for the scores, lanes = channels for the P.V product.
// ------------------------------------------------------------------------------------------------------
constexpr int SEQ_MAX = 64;
__global__ void __launch_bounds__(256)
seq_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int rows_per_batch, int n_groups,
                     int off0, int len0, int off1, int len1, int heads, int dh) {
  extern __shared__ float sm_sa[];
  const int seq = blockIdx.x, head = blockIdx.y;
  const int b = seq / n_groups, g = seq % n_groups;
  const int off = g == 0 ? off0 : off1, n = g == 0 ? len0 : len1;
  const long long row0 = (long long)b * rows_per_batch + off;
  const int inner = heads * dh, ld = 3 * inner;
  const int n_cap = len0 > len1 ? len0 : len1;  // the launch sizes shared memory for its longest group
  float* sK = sm_sa;                    // [n][dh+1]
  float* sV = sK + n_cap * (dh + 1);    // [n][dh]
  float* sQ = sV + n_cap * dh;          // [8 warps][dh]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < n * dh; i += blockDim.x) {
    const int j = i / dh, d = i % dh;
    const float* r = qkv + (row0 + j) * ld + head * dh + d;
    sK[j * (dh + 1) + d] = r[inner];
    sV[j * dh + d] = r[2 * inner];
  }
  __syncthreads();
  const float scale = rsqrtf((float)dh);
  float* q = sQ + warp * dh;
  for (int i = warp; i < n; i += 8) {
    for (int d = lane; d < dh; d += 32) q[d] = qkv[(row0 + i) * ld + head * dh + d];
    __syncwarp();
    float s0 = -INFINITY, s1 = -INFINITY;
    if (lane < n) {
      float acc = 0.f;
      const float* kr = sK + lane * (dh + 1);
      for (int d = 0; d < dh; ++d) acc += q[d] * kr[d];
      s0 = acc * scale;
    }
    if (lane + 32 < n) {
      float acc = 0.f;
      const float* kr = sK + (lane + 32) * (dh + 1);
      for (int d = 0; d < dh; ++d) acc += q[d] * kr[d];
      s1 = acc * scale;
    }
    float m = fmaxf(s0, s1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float p0 = (lane < n) ? expf(s0 - m) : 0.f;
    const float p1 = (lane + 32 < n) ? expf(s1 - m) : 0.f;
    const float inv = 1.f / warp_sum(p0 + p1);
    for (int d0 = 0; d0 < dh; d0 += 32) {
      float acc = 0.f;
      for (int j = 0; j < n; ++j) {
        const float p = __shfl_sync(0xffffffffu, j < 32 ? p0 : p1, j & 31);
        acc += p * sV[j * dh + d0 + lane];
      }
      out[(row0 + i) * inner + head * dh + d0 + lane] = acc * inv;
    }
    __syncwarp();
  }
}
// sized by the longest sequence of the launch (9 tokens for se_te: 27 KB, 8 CTAs per SM instead of 1 at SEQ_MAX)
static size_t seq_attention_smem(int n, int dh) { return (size_t)(n * (dh + 1) + n * dh + 8 * dh) * sizeof(float); }
int k_seq_attention_init() {
  return (int)cudaFuncSetAttribute(seq_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)seq_attention_smem(SEQ_MAX, 256));
}
int k_seq_attention(cudaStream_t st, const float* qkv, float* out, int n_batch, int rows_per_batch, int n_groups,
                    int off0, int len0, int off1, int len1, int heads, int dh) {
  if (len0 > SEQ_MAX || len1 > SEQ_MAX || dh > 256 || dh % 32 != 0) return -2;
  if (n_batch <= 0) return 0;
  dim3 grid(n_batch * n_groups, heads);
  seq_attention_kernel<<<grid, 256, seq_attention_smem(len0 > len1 ? len0 : len1, dh), st>>>(qkv, out, rows_per_batch, n_groups, off0, len0, off1,
                                                                   len1, heads, dh);
  SPM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// class prototypes and PADM sequences (model_clipspm.py:231-239, 275-287)
// z: output of the `sem` se_te call, video v frame t at z + (v*(T+1) + 1 + t)*D, videos ordered [E][S supports, Q queries]
//   su_pro[e,w,t]        = mean_{s in class w} su_real[s,t]
//   seq1[e,t, w]         = token_s[w,t] = mean over (supports of class w  U  all Q queries)
//   seq1[e,t, W + s]     = su_real[s,t]
//   seq1[e,t, W+S]       = token_q[t] = mean_w token_s[w,t]
//   seq1[e,t, W+S+1 + q] = qu_fake[q,t]
// ------------------------------------------------------------------------------------------------------
__global__ void padm_build_kernel(const float* __restrict__ z, const float* __restrict__ labels, int S, int Q, int W,
                                  int T, int D, float* __restrict__ su_pro, float* __restrict__ seq1,
                                  int* __restrict__ err_flag) {
  __shared__ int cls[256];
  const int e = blockIdx.x, t = blockIdx.y, N = S + Q, d4 = D / 4, L1 = W + S + 1 + Q;
  const int Wd = class_indices(labels + (long long)e * S, S, cls);
  if (Wd != W) {
    if (threadIdx.x == 0 && t == 0) atomicExch(err_flag, 1);
    return;
  }
  float4* srow = reinterpret_cast<float4*>(seq1 + ((long long)(e * T + t) * L1) * D);
  for (int c = threadIdx.x; c < d4; c += blockDim.x) {
    float4 qsum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < Q; ++q) {
      const float4 v = reinterpret_cast<const float4*>(z + (((long long)e * N + S + q) * (T + 1) + 1 + t) * D)[c];
      qsum = f4_add(qsum, v);
      srow[(long long)(W + S + 1 + q) * d4 + c] = v;
    }
    float4 tq = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < W; ++w) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
      for (int s = 0; s < S; ++s) {
        if (cls[s] == w) {
          acc = f4_add(acc, reinterpret_cast<const float4*>(z + (((long long)e * N + s) * (T + 1) + 1 + t) * D)[c]);
          ++cnt;
        }
      }
      reinterpret_cast<float4*>(su_pro + (((long long)e * W + w) * T + t) * D)[c] = f4_scale(acc, 1.f / (float)cnt);
      const float4 ts = f4_scale(f4_add(acc, qsum), 1.f / (float)(cnt + Q));
      srow[(long long)w * d4 + c] = ts;
      tq = f4_add(tq, ts);
    }
    srow[(long long)(W + S) * d4 + c] = f4_scale(tq, 1.f / (float)W);
    for (int s = 0; s < S; ++s)
      srow[(long long)(W + s) * d4 + c] =
          reinterpret_cast<const float4*>(z + (((long long)e * N + s) * (T + 1) + 1 + t) * D)[c];
  }
}
int k_padm_build(cudaStream_t st, const float* z, const float* labels, int E, int S, int Q, int W, int T, int D,
                 float* su_pro, float* seq1, int* err_flag) {
  if (S > 256) return -2;
  dim3 grid(E, T);
  padm_build_kernel<<<grid, 128, 0, st>>>(z, labels, S, Q, W, T, D, su_pro, seq1, err_flag);
  SPM_LAUNCH_CHECK();
  return 0;
}

// su_pro2[e,w,t] = mean_{s in class w} z1[e,t,W+s]   (model_clipspm.py:133-137 on the PADM output)
__global__ void class_mean_padm_kernel(const float* __restrict__ z1, const float* __restrict__ labels, int S, int Q,
                                       int W, int T, int D, float* __restrict__ su_pro2) {
  __shared__ int cls[256];
  const int e = blockIdx.x, t = blockIdx.y, d4 = D / 4, L1 = W + S + 1 + Q;
  class_indices(labels + (long long)e * S, S, cls);
  const float4* base = reinterpret_cast<const float4*>(z1 + ((long long)(e * T + t) * L1) * D);
  for (int c = threadIdx.x; c < d4; c += blockDim.x) {
    for (int w = 0; w < W; ++w) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
      for (int s = 0; s < S; ++s)
        if (cls[s] == w) { acc = f4_add(acc, base[(long long)(W + s) * d4 + c]); ++cnt; }
      reinterpret_cast<float4*>(su_pro2 + (((long long)e * W + w) * T + t) * D)[c] =
          f4_scale(acc, 1.f / (float)max(cnt, 1));
    }
  }
}
int k_class_mean_padm(cudaStream_t st, const float* z1, const float* labels, int E, int S, int Q, int W, int T, int D,
                      float* su_pro2) {
  dim3 grid(E, T);
  class_mean_padm_kernel<<<grid, 128, 0, st>>>(z1, labels, S, Q, W, T, D, su_pro2);
  SPM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// logits / loss / accuracy (model_clipspm.py:138-143; utils/utils.py:174-186,259-264; run/main_run.py:390-392)
//   logits[e,q,w] = -(acc[e,q,w] + d3[e,w])     acc = 0.5*class_dists_l + otam(su_pro2, qu_2), d3 = otam(su_t2, qu_t2)
//   loss[e] = sum_q CE(logits[e,q,:], y[e,q]) / tasks_per_batch + 0.001 * dists[e]
//   acc[e]  = mean_q [argmax_w logits[e,q,w] == y[e,q]]      (first maximal index, like torch.argmax)
// ------------------------------------------------------------------------------------------------------
__global__ void finalize_kernel(const float* __restrict__ accd, const float* __restrict__ d3, int Q, int W,
                                const long long* __restrict__ target, float tasks_per_batch,
                                const float* __restrict__ dists, float* __restrict__ logits, float* __restrict__ loss,
                                float* __restrict__ accuracy, int* __restrict__ pred,
                                const int* __restrict__ err_flag) {
  __shared__ float s_ce[64];
  __shared__ int s_ok[64];
  const int e = blockIdx.x;
  for (int q = threadIdx.x; q < Q; q += blockDim.x) {
    float mx = -INFINITY;
    int am = 0;
    for (int w = 0; w < W; ++w) {
      float lg = -(accd[((long long)e * Q + q) * W + w] + d3[(long long)e * W + w]);
      if (err_flag != nullptr && *err_flag != 0) lg = __int_as_float(0x7fc00000);  // W mismatch: fail loudly
      logits[((long long)e * Q + q) * W + w] = lg;
      if (lg > mx) { mx = lg; am = w; }
    }
    if (pred != nullptr) pred[(long long)e * Q + q] = am;
    if (target != nullptr) {
      float se = 0.f;
      for (int w = 0; w < W; ++w) se += expf(logits[((long long)e * Q + q) * W + w] - mx);
      const int y = (int)target[(long long)e * Q + q];
      const float ly = logits[((long long)e * Q + q) * W + min(max(y, 0), W - 1)];
      s_ce[q] = (mx + logf(se)) - ly;
      s_ok[q] = (am == y) ? 1 : 0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && target != nullptr) {
    float ce = 0.f;
    int ok = 0;
    for (int q = 0; q < Q; ++q) { ce += s_ce[q]; ok += s_ok[q]; }
    if (loss != nullptr) loss[e] = ce / tasks_per_batch + 0.001f * dists[e];
    if (accuracy != nullptr) accuracy[e] = (float)ok / (float)Q;
  }
}
int k_finalize(cudaStream_t st, const float* accd, const float* d3, int E, int Q, int W, const long long* target,
               float tasks_per_batch, const float* dists, float* logits, float* loss, float* accuracy, int* pred,
               const int* err_flag) {
  if (Q > 64) return -2;
  finalize_kernel<<<E, 64, 0, st>>>(accd, d3, Q, W, target, tasks_per_batch, dists, logits, loss, accuracy, pred,
                                    err_flag);
  SPM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// strided 4-D gather (spm_head_stage: copies one stage tensor of the head workspace out in the reference's layout)
//   out[i0, i1, i2, :n3] = src[i0*s0 + i1*s1 + i2*s2 + (0..n3)]      n3 % 4 == 0, strides in floats (% 4 == 0)
// ------------------------------------------------------------------------------------------------------
__global__ void gather4_kernel(const float* __restrict__ src, int n1, int n2, int n3, long long s0, long long s1,
                               long long s2, long long total4, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int d4 = n3 / 4;
  const int c = (int)(i % d4);
  const int i2 = (int)((i / d4) % n2);
  const int i1 = (int)((i / ((long long)d4 * n2)) % n1);
  const long long i0 = i / ((long long)d4 * n2 * n1);
  reinterpret_cast<float4*>(out)[i] = *reinterpret_cast<const float4*>(src + i0 * s0 + i1 * s1 + i2 * s2 + c * 4);
}
int k_gather4(cudaStream_t st, const float* src, int n0, int n1, int n2, int n3, long long s0, long long s1,
              long long s2, float* out) {
  if (n3 % 4 != 0 || s0 % 4 != 0 || s1 % 4 != 0 || s2 % 4 != 0) return -2;
  const long long total4 = (long long)n0 * n1 * n2 * (n3 / 4);
  if (total4 <= 0) return 0;
  gather4_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(src, n1, n2, n3, s0, s1, s2, total4, out);
  SPM_LAUNCH_CHECK();
  return 0;
}

}  // namespace spm
