// Sibling head CPM2C (models/model_cpm2c.py::CLIP_CPMMC_FSAR, evaluation forward :207-312) -- the small kernels around
// the shared GEMM / transformer-block / OTAM kernels.  fp32, float4-vectorised, labels resolved on the device.
//   Z layout used below: the context2 outputs of one pass, [2][V][L][D] with L = Tp + 1 rows per sequence (row 0 = the
//   token), call 0 = sequences built on the REAL prompt of each video, call 1 = on the class token ("fake"); videos are
//   ordered [E][S supports, Q queries].  In the reference's names (:327-418): support_features = Z[0][supports],
//   support_features_contra = Z[1][supports], target_features = Z[1][queries], target_features_contra = Z[0][queries].
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

// temporal im2col with dilation (Conv1d k=3, padding = dilation): out[(v*T+t), kk*D + c] = x[v, t + (kk-1)*dil, c]
__global__ void temporal_im2col_dil_kernel(const float* __restrict__ x, int V, int T, int D, int dil, float* __restrict__ out) {
  const int d4 = D / 4;
  const long long n = (long long)V * T * 3 * d4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % d4), kk = (int)((i / d4) % 3), t = (int)((i / (3LL * d4)) % T);
  const long long v = i / (3LL * d4 * T);
  const int ts = t + (kk - 1) * dil;
  float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ts >= 0 && ts < T) val = *reinterpret_cast<const float4*>(x + (v * T + ts) * D + c * 4);
  reinterpret_cast<float4*>(out)[i] = val;
}
int k_temporal_im2col_dil(cudaStream_t st, const float* x, int V, int T, int D, int dil, float* out) {
  const long long n = (long long)V * T * 3 * (D / 4);
  if (n <= 0) return 0;
  temporal_im2col_dil_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, V, T, D, dil, out);
  SPM_LAUNCH_CHECK();
  return 0;
}

// model_cpm2c.py:187-199: motion[v,t] = 0.5 * ((conv[v,t+1] - x[v,t]) + (conv[v,t] - x[v,t+1])), t < T-1
__global__ void cpm2c_motion_diff_kernel(const float* __restrict__ conv, const float* __restrict__ x, int V, int T, int D,
                                         float* __restrict__ out) {
  const int d4 = D / 4;
  const long long n = (long long)V * (T - 1) * d4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % d4), t = (int)((i / d4) % (T - 1));
  const long long v = i / ((long long)d4 * (T - 1));
  const float4* cv = reinterpret_cast<const float4*>(conv + (v * T + t) * D) + c;
  const float4* xv = reinterpret_cast<const float4*>(x + (v * T + t) * D) + c;
  const float4 c0 = cv[0], c1 = cv[d4], x0 = xv[0], x1 = xv[d4];
  float4 o;
  o.x = 0.5f * ((c1.x - x0.x) + (c0.x - x1.x)); o.y = 0.5f * ((c1.y - x0.y) + (c0.y - x1.y));
  o.z = 0.5f * ((c1.z - x0.z) + (c0.z - x1.z)); o.w = 0.5f * ((c1.w - x0.w) + (c0.w - x1.w));
  reinterpret_cast<float4*>(out)[i] = o;
}
int k_cpm2c_motion_diff(cudaStream_t st, const float* conv, const float* x, int V, int T, int D, float* out) {
  const long long n = (long long)V * (T - 1) * (D / 4);
  if (n <= 0) return 0;
  cpm2c_motion_diff_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(conv, x, V, T, D, out);
  SPM_LAUNCH_CHECK();
  return 0;
}

// tokens of the two calls (:214-218, :349, :364): tok[0][v] = text[real label of video v], tok[1][v] = class token
__global__ void cpm2c_tokens_kernel(const float* __restrict__ text, int n_cls, const float* __restrict__ real_s,
                                    const float* __restrict__ real_t, const float* __restrict__ cls_token, int S, int Q,
                                    long long V, int D, float* __restrict__ tok, int* __restrict__ err_flag) {
  const long long v = blockIdx.x;
  const int N = S + Q, e = (int)(v / N), i = (int)(v % N);
  const float r = i < S ? real_s[(long long)e * S + i] : real_t[(long long)e * Q + (i - S)];
  const bool ok = r >= 0.f && r < (float)n_cls;
  if (!ok && threadIdx.x == 0) atomicExch(err_flag, 2);
  const float4* src = reinterpret_cast<const float4*>(text + (ok ? (long long)r : 0LL) * D);
  const float4* ct = reinterpret_cast<const float4*>(cls_token);
  for (int k = threadIdx.x; k < D / 4; k += blockDim.x) {
    reinterpret_cast<float4*>(tok + v * D)[k] = src[k];
    reinterpret_cast<float4*>(tok + (V + v) * D)[k] = ct[k];
  }
}
int k_cpm2c_tokens(cudaStream_t st, const float* text, int n_cls, const float* real_s, const float* real_t,
                   const float* cls_token, int E, int S, int Q, int D, float* tok, int* err_flag) {
  const long long V = (long long)E * (S + Q);
  cpm2c_tokens_kernel<<<(unsigned)V, 128, 0, st>>>(text, n_cls, real_s, real_t, cls_token, S, Q, V, D, tok, err_flag);
  SPM_LAUNCH_CHECK();
  return 0;
}

// class prototypes over ALL L rows (:412-416): su_pro[e,w,l] = mean_{s in class w} Z[0][e*N+s][l]
__global__ void cpm2c_class_mean_kernel(const float* __restrict__ z, const float* __restrict__ labels, int S, int Q, int W,
                                        int L, int D, float* __restrict__ su_pro, int* __restrict__ err_flag) {
  __shared__ int cls[256];
  const int e = blockIdx.x, l = blockIdx.y, N = S + Q, d4 = D / 4;
  const int Wd = class_indices(labels + (long long)e * S, S, cls);
  if (Wd != W) {
    if (threadIdx.x == 0 && l == 0) atomicExch(err_flag, 1);
    return;
  }
  for (int c = threadIdx.x; c < d4; c += blockDim.x)
    for (int w = 0; w < W; ++w) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
      for (int s = 0; s < S; ++s)
        if (cls[s] == w) { acc = f4_add(acc, reinterpret_cast<const float4*>(z + (((long long)e * N + s) * L + l) * D)[c]); ++cnt; }
      reinterpret_cast<float4*>(su_pro + (((long long)e * W + w) * L + l) * D)[c] = f4_scale(acc, 1.f / (float)max(cnt, 1));
    }
}
int k_cpm2c_class_mean(cudaStream_t st, const float* z, const float* labels, int E, int S, int Q, int W, int L, int D,
                       float* su_pro, int* err_flag) {
  if (S > 256) return -2;
  cpm2c_class_mean_kernel<<<dim3(E, L), 128, 0, st>>>(z, labels, S, Q, W, L, D, su_pro, err_flag);
  SPM_LAUNCH_CHECK();
  return 0;
}

// consistency distance (:243-271): consist[e] = beta*consist[e] + coeff * ( mean_s |Z0[s] - Z1[s]|_F^2 + mean_q |Z1[q] - Z0[q]|_F^2 )
__global__ void cpm2c_consist_kernel(const float* __restrict__ z, long long V, int S, int Q, int L, int D, float coeff,
                                     float beta, float* __restrict__ consist) {
  __shared__ float red[32];
  const int e = blockIdx.x, N = S + Q;
  const long long LD = (long long)L * D;
  float ss = 0.f, sq = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* a = z + ((long long)e * N + n) * LD;
    const float* b = a + V * LD;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < LD; i += blockDim.x) { const float df = a[i] - b[i]; acc += df * df; }
    if (n < S) ss += acc; else sq += acc;
  }
  ss = block_sum(ss, red);
  sq = block_sum(sq, red);
  if (threadIdx.x == 0) consist[e] = (beta != 0.f ? beta * consist[e] : 0.f) + coeff * (ss / (float)S + sq / (float)Q);
}
int k_cpm2c_consist(cudaStream_t st, const float* z, int E, int S, int Q, int L, int D, float coeff, float beta, float* consist) {
  cpm2c_consist_kernel<<<E, 256, 0, st>>>(z, (long long)E * (S + Q), S, Q, L, D, coeff, beta, consist);
  SPM_LAUNCH_CHECK();
  return 0;
}

// global distance (:315-325, :279-285): g[e,q,w] = beta*g + coeff * sum_{s in class w} sum_l (1 - cos(Z1[q][l], Z0[s][0]))
// cos_sim of myRes.py:756-765: x.y / (|x||y| + 0.01).  One CTA per (episode, query); warp = (l, s) pairs.
__global__ void cpm2c_global_kernel(const float* __restrict__ z, const float* __restrict__ labels, long long V, int S, int Q,
                                    int W, int L, int D, float coeff, float beta, float* __restrict__ g) {
  extern __shared__ float sm_gd[];   // [L*S] pair distances
  __shared__ int cls[256];
  const int e = blockIdx.x, q = blockIdx.y, N = S + Q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  class_indices(labels + (long long)e * S, S, cls);
  const long long LD = (long long)L * D;
  const float* tq = z + (V + (long long)e * N + S + q) * LD;          // Z1[query]
  for (int p = warp; p < L * S; p += nw) {
    const int l = p / S, s = p - l * S;
    const float* x = tq + (long long)l * D;
    const float* y = z + ((long long)e * N + s) * LD;                 // Z0[support] row 0 (its token)
    float dot = 0.f, xx = 0.f, yy = 0.f;
    for (int d = lane; d < D; d += 32) { const float a = x[d], b = y[d]; dot = fmaf(a, b, dot); xx = fmaf(a, a, xx); yy = fmaf(b, b, yy); }
    dot = warp_sum(dot); xx = warp_sum(xx); yy = warp_sum(yy);
    if (lane == 0) sm_gd[p] = 1.f - dot / (sqrtf(xx) * sqrtf(yy) + 0.01f);
  }
  __syncthreads();
  if (threadIdx.x < W) {   // fixed summation order: deterministic
    float acc = 0.f;
    for (int s = 0; s < S; ++s)
      if (cls[s] == (int)threadIdx.x)
        for (int l = 0; l < L; ++l) acc += sm_gd[l * S + s];
    float* o = g + ((long long)e * Q + q) * W + threadIdx.x;
    *o = (beta != 0.f ? beta * (*o) : 0.f) + coeff * acc;
  }
}
int k_cpm2c_global(cudaStream_t st, const float* z, const float* labels, int E, int S, int Q, int W, int L, int D,
                   float coeff, float beta, float* g) {
  if (S > 256 || W > 32 || (size_t)L * S * 4 > 40 * 1024) return -2;
  cpm2c_global_kernel<<<dim3(E, Q), 256, (size_t)L * S * sizeof(float), st>>>(z, labels, (long long)E * (S + Q), S, Q, W, L, D,
                                                                             coeff, beta, g);
  SPM_LAUNCH_CHECK();
  return 0;
}

// outputs + loss / accuracy (model_cpm2c.py:229-236, run/main_run.py:370-380):
//   logits_local = -loc, logits_global = -glob, total = l1 * local + l2 * global (the logits accuracy is taken on)
//   loss[e] = (l0 * sum_v CE(class_logits[v], real label) + l1 * sum_q CE(local) + l2 * sum_q CE(global)) / tasks_per_batch
__global__ void cpm2c_finalize_kernel(const float* __restrict__ loc, const float* __restrict__ glob, const float* __restrict__ cls,
                                      int n_cls, const float* __restrict__ real_s, const float* __restrict__ real_t, int S,
                                      int Q, int W, const long long* __restrict__ target, float l0, float l1, float l2,
                                      float tasks_per_batch, float* __restrict__ out_local, float* __restrict__ out_global,
                                      float* __restrict__ out_total, float* __restrict__ loss, float* __restrict__ accuracy,
                                      int* __restrict__ pred, const int* __restrict__ err_flag) {
  __shared__ float s_ce[64];
  __shared__ int s_ok[64];
  __shared__ float red[32];
  const int e = blockIdx.x, N = S + Q;
  const bool bad = err_flag != nullptr && *err_flag != 0;
  const float nanv = __int_as_float(0x7fc00000);
  for (int q = threadIdx.x; q < Q; q += blockDim.x) {
    float mt = -INFINITY, ml = -INFINITY, mg = -INFINITY;
    int am = 0;
    for (int w = 0; w < W; ++w) {
      const long long i = ((long long)e * Q + q) * W + w;
      const float a = bad ? nanv : -loc[i], b = bad ? nanv : -glob[i], t = l1 * a + l2 * b;
      out_local[i] = a; out_global[i] = b; out_total[i] = t;
      if (t > mt) { mt = t; am = w; }
      ml = fmaxf(ml, a); mg = fmaxf(mg, b);
    }
    if (pred != nullptr) pred[(long long)e * Q + q] = am;
    if (target != nullptr) {
      float sl = 0.f, sg = 0.f;
      for (int w = 0; w < W; ++w) {
        const long long i = ((long long)e * Q + q) * W + w;
        sl += expf(out_local[i] - ml); sg += expf(out_global[i] - mg);
      }
      const int y = (int)target[(long long)e * Q + q], yc = min(max(y, 0), W - 1);
      const long long iy = ((long long)e * Q + q) * W + yc;
      s_ce[q] = l1 * ((ml + logf(sl)) - out_local[iy]) + l2 * ((mg + logf(sg)) - out_global[iy]);
      s_ok[q] = (am == y) ? 1 : 0;
    }
  }
  // class-logit cross entropy over the episode's S + Q videos (one warp per row)
  float ce_cls = 0.f;
  if (target != nullptr && cls != nullptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int n = warp; n < N; n += nw) {
      const float* row = cls + ((long long)e * N + n) * n_cls;
      float m = -INFINITY;
      for (int c = lane; c < n_cls; c += 32) m = fmaxf(m, row[c]);
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float se = 0.f;
      for (int c = lane; c < n_cls; c += 32) se += expf(row[c] - m);
      se = warp_sum(se);
      const float r = n < S ? real_s[(long long)e * S + n] : real_t[(long long)e * Q + (n - S)];
      const int y = min(max((int)r, 0), n_cls - 1);
      if (lane == 0) ce_cls += (m + logf(se)) - row[y];
    }
  }
  else if (target != nullptr && threadIdx.x == 0 && n_cls > 0) {
    ce_cls = (float)N * logf((float)n_cls);   // USE_CLASSIFICATION off: class_logits are zeros (model_cpm2c.py:423-424)
  }
  ce_cls = block_sum(ce_cls, red);
  __syncthreads();
  if (threadIdx.x == 0 && target != nullptr) {
    float ce = 0.f;
    int ok = 0;
    for (int q = 0; q < Q; ++q) { ce += s_ce[q]; ok += s_ok[q]; }
    if (loss != nullptr) loss[e] = (l0 * ce_cls + ce) / tasks_per_batch;
    if (accuracy != nullptr) accuracy[e] = (float)ok / (float)Q;
  }
}
int k_cpm2c_finalize(cudaStream_t st, const float* loc, const float* glob, const float* cls, int n_cls, const float* real_s,
                     const float* real_t, int E, int S, int Q, int W, const long long* target, float l0, float l1, float l2,
                     float tasks_per_batch, float* out_local, float* out_global, float* out_total, float* loss,
                     float* accuracy, int* pred, const int* err_flag) {
  if (Q > 64) return -2;
  cpm2c_finalize_kernel<<<E, 128, 0, st>>>(loc, glob, cls, n_cls, real_s, real_t, S, Q, W, target, l0, l1, l2, tasks_per_batch,
                                           out_local, out_global, out_total, loss, accuracy, pred, err_flag);
  SPM_LAUNCH_CHECK();
  return 0;
}

}  // namespace spm
