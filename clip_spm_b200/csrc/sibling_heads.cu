// Sibling heads on the library's kernels.  CLIP-FSAR (models/model_clipfsar.py:325-383, evaluation
// branch): the non-GEMM kernels around the shared transformer block / OTAM kernels; STEN (models/model_sten.py:62-113 as
// shipped): the whole head.  fp32, float4-vectorised, labels resolved on the device (no host sync).
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
// context2 input sequences (model_clipfsar.py:338-348).  X [E, S+Q, T, D], supports first.
//   support video sv = e*S + s : rows sv*(T+1) + {0: text_test[real_support[sv]], 1+t: X[e,s,t]}
//   query video   qv = e*Q + q : rows E*S*(T+1) + qv*T + t = X[e,S+q,t]
// The reference appends the prompt token AFTER the frames (torch.cat([support, context], dim=1)) and keeps rows
// [:T]; Transformer_v1 has no positional term, so the token's position is free -- it goes first here, the layout
// the other kernels of this library already use for (token, frames) sequences.
// ------------------------------------------------------------------------------------------------------
__global__ void fsar_seq_build_kernel(const float* __restrict__ X, const float* __restrict__ text, int n_cls,
                                      const float* __restrict__ real_s, int E, int S, int Q, int T, int D,
                                      float* __restrict__ seq) {
  const int v = blockIdx.x, N = S + Q, e = v / N, i = v % N, d4 = D / 4;
  const float4* src = reinterpret_cast<const float4*>(X + (long long)v * T * D);
  float4* dst;
  if (i < S) {
    const long long sv = (long long)e * S + i;
    float4* base = reinterpret_cast<float4*>(seq + sv * (T + 1) * D);
    const int c = min(max((int)real_s[sv], 0), n_cls - 1);   // .long() of the float label (model_clipfsar.py:338)
    const float4* tok = reinterpret_cast<const float4*>(text + (long long)c * D);
    for (int k = threadIdx.x; k < d4; k += blockDim.x) base[k] = tok[k];
    dst = base + d4;
  } else {
    dst = reinterpret_cast<float4*>(seq + ((long long)E * S * (T + 1) + ((long long)e * Q + (i - S)) * T) * D);
  }
  for (int k = threadIdx.x; k < T * d4; k += blockDim.x) dst[k] = src[k];
}
int k_fsar_seq_build(cudaStream_t st, const float* X, const float* text, int n_cls, const float* real_s, int E, int S,
                     int Q, int T, int D, float* seq) {
  fsar_seq_build_kernel<<<E * (S + Q), 128, 0, st>>>(X, text, n_cls, real_s, E, S, Q, T, D, seq);
  SPM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// MODEL.MERGE_BEFORE (model_clipfsar.py:341-348): the class means are taken BEFORE context2 -- of the support frames and
// of the supports' prompts -- so the support batch is E*W sequences of (mean prompt, T mean frames):
//   class sequence cw = e*W + w : rows cw*(T+1) + {0: mean_s text[real_s[s]], 1+t: mean_s X[e,s,t]},  s in class w
//   query video   qv = e*Q + q : rows E*W*(T+1) + qv*T + t = X[e,S+q,t]
// grid (E*(W+Q), T+1); class w = rank among the episode's sorted distinct labels (torch.unique, :342)
// ------------------------------------------------------------------------------------------------------
__global__ void fsar_merge_seq_build_kernel(const float* __restrict__ X, const float* __restrict__ text, int n_cls,
                                            const float* __restrict__ labels, const float* __restrict__ real_s, int E,
                                            int S, int Q, int W, int T, int D, float* __restrict__ seq,
                                            int* __restrict__ err_flag) {
  __shared__ int cls[256];
  const int e = blockIdx.x / (W + Q), i = blockIdx.x % (W + Q), r = blockIdx.y, N = S + Q, d4 = D / 4;
  if (i >= W) {   // query rows are copied; (T+1)-th grid row has nothing to do
    if (r >= T) return;
    const float4* src = reinterpret_cast<const float4*>(X + (((long long)e * N + S + (i - W)) * T + r) * D);
    float4* dst = reinterpret_cast<float4*>(seq + ((long long)E * W * (T + 1) + ((long long)e * Q + (i - W)) * T + r) * D);
    for (int k = threadIdx.x; k < d4; k += blockDim.x) dst[k] = src[k];
    return;
  }
  const int Wd = class_indices(labels + (long long)e * S, S, cls);
  if (Wd != W) {
    if (threadIdx.x == 0 && r == 0 && i == 0) atomicExch(err_flag, 1);
    return;
  }
  float4* dst = reinterpret_cast<float4*>(seq + (((long long)e * W + i) * (T + 1) + r) * D);
  for (int k = threadIdx.x; k < d4; k += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
    for (int s = 0; s < S; ++s)
      if (cls[s] == i) {
        const float* row;
        if (r == 0) {
          const int c = min(max((int)real_s[(long long)e * S + s], 0), n_cls - 1);
          row = text + (long long)c * D;
        } else {
          row = X + (((long long)e * N + s) * T + (r - 1)) * D;
        }
        const float4 v = reinterpret_cast<const float4*>(row)[k];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        ++cnt;
      }
    const float inv = 1.f / (float)max(cnt, 1);
    dst[k] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}
int k_fsar_merge_seq_build(cudaStream_t st, const float* X, const float* text, int n_cls, const float* labels,
                           const float* real_s, int E, int S, int Q, int W, int T, int D, float* seq, int* err_flag) {
  if (S > 256) return -2;
  dim3 grid(E * (W + Q), T + 1);
  fsar_merge_seq_build_kernel<<<grid, 128, 0, st>>>(X, text, n_cls, labels, real_s, E, S, Q, W, T, D, seq, err_flag);
  SPM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// class prototypes (model_clipfsar.py:352-354): su_pro[e,w,t] = mean_{s in class w} z[(e*S+s)*(T+1) + 1 + t]
// ------------------------------------------------------------------------------------------------------
__global__ void fsar_class_mean_kernel(const float* __restrict__ z, const float* __restrict__ labels, int S, int W,
                                       int T, int D, float* __restrict__ su_pro, int* __restrict__ err_flag) {
  __shared__ int cls[256];
  const int e = blockIdx.x, t = blockIdx.y, d4 = D / 4;
  // rank of each label among the episode's sorted distinct labels == position in torch.unique (:352)
  const int Wd = class_indices(labels + (long long)e * S, S, cls);
  if (Wd != W) {
    if (threadIdx.x == 0 && t == 0) atomicExch(err_flag, 1);
    return;
  }
  for (int c = threadIdx.x; c < d4; c += blockDim.x) {
    for (int w = 0; w < W; ++w) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
      for (int s = 0; s < S; ++s)
        if (cls[s] == w) {
          const float4 v = reinterpret_cast<const float4*>(z + (((long long)e * S + s) * (T + 1) + 1 + t) * D)[c];
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          ++cnt;
        }
      const float r = 1.f / (float)max(cnt, 1);
      reinterpret_cast<float4*>(su_pro + (((long long)e * W + w) * T + t) * D)[c] =
          make_float4(acc.x * r, acc.y * r, acc.z * r, acc.w * r);
    }
  }
}
int k_fsar_class_mean(cudaStream_t st, const float* z, const float* labels, int E, int S, int W, int T, int D,
                      float* su_pro, int* err_flag) {
  if (S > 256) return -2;
  dim3 grid(E, T);
  fsar_class_mean_kernel<<<grid, 128, 0, st>>>(z, labels, S, W, T, D, su_pro, err_flag);
  SPM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// class_text_logits (model_clipfsar.py:329-331): cos_sim(mean_t X[v], text_train) * scale, cos_sim of myRes.py:756-765
//   out[v, c] = scale * <m, y_c> / (|m| |y_c| + 0.01),  m = mean_t X[v,t,:]
// ------------------------------------------------------------------------------------------------------
__global__ void fsar_class_logits_kernel(const float* __restrict__ X, const float* __restrict__ text, int n_cls,
                                         const float* __restrict__ scale, int T, int D, float* __restrict__ out) {
  extern __shared__ float sm_cl[];   // [D] mean feature, [32] reduction scratch
  float* m = sm_cl;
  float* red = sm_cl + D;
  const int v = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* x = X + (long long)v * T * D;
  float nn = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a = 0.f;
    for (int t = 0; t < T; ++t) a += x[(long long)t * D + d];
    a /= (float)T;
    m[d] = a;
    nn += a * a;
  }
  nn = warp_sum(nn);
  if (lane == 0) red[warp] = nn;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < nw; ++w) tot += red[w];
  const float mnorm = sqrtf(tot), sc = scale[0];
  for (int c = warp; c < n_cls; c += nw) {
    const float* y = text + (long long)c * D;
    float dot = 0.f, yy = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float yv = y[d];
      dot = fmaf(m[d], yv, dot);
      yy = fmaf(yv, yv, yy);
    }
    dot = warp_sum(dot);
    yy = warp_sum(yy);
    if (lane == 0) out[(long long)v * n_cls + c] = sc * dot / (mnorm * sqrtf(yy) + 0.01f);
  }
}
int k_fsar_class_logits(cudaStream_t st, const float* X, const float* text_train, int n_cls, const float* scale, int V,
                        int T, int D, float* out) {
  fsar_class_logits_kernel<<<V, 256, (size_t)(D + 32) * sizeof(float), st>>>(X, text_train, n_cls, scale, T, D, out);
  SPM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// loss[e] += coef * sum_v CE(class_logits[e,v,:], real[e,v])   (run/main_run.py:355-356; rows = supports then queries)
// ------------------------------------------------------------------------------------------------------
__global__ void fsar_class_ce_add_kernel(const float* __restrict__ cls, const float* __restrict__ real_s,
                                         const float* __restrict__ real_t, int S, int Q, int n_cls, float coef,
                                         float* __restrict__ loss) {
  __shared__ float s_ce[256];
  const int e = blockIdx.x, N = S + Q;
  for (int v = threadIdx.x; v < N; v += blockDim.x) {
    const float* row = cls + ((long long)e * N + v) * n_cls;
    float mx = -INFINITY;
    for (int c = 0; c < n_cls; ++c) mx = fmaxf(mx, row[c]);
    float se = 0.f;
    for (int c = 0; c < n_cls; ++c) se += expf(row[c] - mx);
    const int y = (int)(v < S ? real_s[(long long)e * S + v] : real_t[(long long)e * Q + v - S]);
    // a label outside the table has no logit: the reference's cross_entropy raises there; here the loss becomes NaN
    s_ce[v] = (y >= 0 && y < n_cls) ? (mx + logf(se)) - row[y] : __int_as_float(0x7fc00000);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ce = 0.f;
    for (int v = 0; v < N; ++v) ce += s_ce[v];   // fixed order: deterministic
    loss[e] += coef * ce;
  }
}
int k_fsar_class_ce_add(cudaStream_t st, const float* cls, const float* real_s, const float* real_t, int E, int S, int Q,
                        int n_cls, float coef, float* loss) {
  if (S + Q > 256) return -2;
  fsar_class_ce_add_kernel<<<E, 64, 0, st>>>(cls, real_s, real_t, S, Q, n_cls, coef, loss);
  SPM_LAUNCH_CHECK();
  return 0;
}

}  // namespace spm

// ======================================================================================================
// Sibling head STEN as shipped (models/model_sten.py:62-113): no learned head at all --
//   su_f, qu_f = mean over the T frames; per class (sorted distinct support labels) the mean support feature and the
//   mean prompt; logits[q, w] = softmax_w(cos_sim(qu_f, t_f))[q, w] * softmax_w(cos_sim(qu_f, su_f))[q, w]
// ======================================================================================================
namespace spm {

// out[v, :] = mean_t X[v, t, :]
__global__ void sten_frame_mean_kernel(const float* __restrict__ X, int T, int D, float* __restrict__ out) {
  const long long v = blockIdx.x;
  const float4* x = reinterpret_cast<const float4*>(X + v * T * D);
  const int d4 = D / 4;
  const float r = 1.f / (float)T;
  for (int c = threadIdx.x; c < d4; c += blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      const float4 q = x[(long long)t * d4 + c];
      a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
    }
    reinterpret_cast<float4*>(out + v * D)[c] = make_float4(a.x * r, a.y * r, a.z * r, a.w * r);
  }
}

// One CTA per episode.  m [E, S+Q, D] frame means (supports first).  proto [E, W, 2, D] scratch: class means of the
// support features and of their prompts.  negsim [E, Q, W] = -(softmax(cos qt) * softmax(cos qs)) (the finalize
// kernel negates its input).
__global__ void __launch_bounds__(256)
sten_sim_kernel(const float* __restrict__ m, const float* __restrict__ text, int n_cls,
                const float* __restrict__ labels, const float* __restrict__ real_s, int S, int Q, int W, int D,
                float* __restrict__ proto, float* __restrict__ negsim, int* __restrict__ err_flag) {
  __shared__ int cls[256];
  __shared__ float pn[64];        // [W][2] prototype norms
  __shared__ float cs[4096];      // [Q][W][2] cosines (Q*W <= 2048, checked by the launcher)
  const int e = blockIdx.x, N = S + Q, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int Wd = class_indices(labels + (long long)e * S, S, cls);   // rank among the sorted distinct labels (:97)
  if (Wd != W) {
    if (threadIdx.x == 0) atomicExch(err_flag, 1);
    return;
  }
  float* pr = proto + (long long)e * W * 2 * D;
  const float* me_ = m + (long long)e * N * D;
  // class means (:98-102), one warp per class; the norms fall out of the same pass
  for (int w = warp; w < W; w += nw) {
    float ns = 0.f, nt = 0.f;
    for (int d = lane; d < D; d += 32) {
      float a = 0.f, b = 0.f;
      int cnt = 0;
      for (int s = 0; s < S; ++s)
        if (cls[s] == w) {
          a += me_[(long long)s * D + d];
          const int c = min(max((int)real_s[(long long)e * S + s], 0), n_cls - 1);
          b += text[(long long)c * D + d];
          ++cnt;
        }
      a /= (float)cnt; b /= (float)cnt;
      pr[((long long)w * 2 + 0) * D + d] = a;
      pr[((long long)w * 2 + 1) * D + d] = b;
      ns = fmaf(a, a, ns); nt = fmaf(b, b, nt);
    }
    ns = warp_sum(ns); nt = warp_sum(nt);
    if (lane == 0) { pn[w * 2] = sqrtf(ns); pn[w * 2 + 1] = sqrtf(nt); }
  }
  __syncthreads();   // the prototypes were written by this CTA: visible to it after the barrier
  // cosines (myRes.py:756-765), one warp per (query, class)
  for (int i = warp; i < Q * W; i += nw) {
    const int q = i / W, w = i - q * W;
    const float* x = me_ + (long long)(S + q) * D;
    float ds = 0.f, dt = 0.f, xx = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float xv = x[d];
      ds = fmaf(xv, pr[((long long)w * 2 + 0) * D + d], ds);
      dt = fmaf(xv, pr[((long long)w * 2 + 1) * D + d], dt);
      xx = fmaf(xv, xv, xx);
    }
    ds = warp_sum(ds); dt = warp_sum(dt); xx = warp_sum(xx);
    if (lane == 0) {
      const float xn = sqrtf(xx);
      cs[i * 2 + 0] = ds / (xn * pn[w * 2] + 0.01f);
      cs[i * 2 + 1] = dt / (xn * pn[w * 2 + 1] + 0.01f);
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < Q; q += blockDim.x) {   // the two softmaxes over the classes, multiplied (:104-106)
    float ms = -INFINITY, mt = -INFINITY;
    for (int w = 0; w < W; ++w) { ms = fmaxf(ms, cs[(q * W + w) * 2]); mt = fmaxf(mt, cs[(q * W + w) * 2 + 1]); }
    float ss = 0.f, st = 0.f;
    for (int w = 0; w < W; ++w) { ss += expf(cs[(q * W + w) * 2] - ms); st += expf(cs[(q * W + w) * 2 + 1] - mt); }
    for (int w = 0; w < W; ++w)
      negsim[((long long)e * Q + q) * W + w] =
          -((expf(cs[(q * W + w) * 2 + 1] - mt) / st) * (expf(cs[(q * W + w) * 2] - ms) / ss));
  }
}

int k_sten_head(cudaStream_t st, const float* X, const float* text, int n_cls, const float* labels,
                const float* real_s, int E, int S, int Q, int W, int T, int D, float* frame_mean, float* proto,
                float* negsim, int* err_flag) {
  if (S > 256 || Q * W > 2048 || W > 32) return -2;
  sten_frame_mean_kernel<<<E * (S + Q), 128, 0, st>>>(X, T, D, frame_mean);
  SPM_LAUNCH_CHECK();
  sten_sim_kernel<<<E, 256, 0, st>>>(frame_mean, text, n_cls, labels, real_s, S, Q, W, D, proto, negsim, err_flag);
  SPM_LAUNCH_CHECK();
  return 0;
}

}  // namespace spm
