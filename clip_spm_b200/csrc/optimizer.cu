// Optimiser half of the training step: what run/main_run.py does with the gradients --
//   :84-88    torch.optim.Adam(model.parameters(), lr, betas=(0.5, 0.999), weight_decay)   (classic L2: wd * p joins the gradient)
//   :76       GradScaler(enabled=USE_AMP);  :252 scaler.scale(loss).backward()
//   :207-209  scaler.step(optimizer); scaler.update(); optimizer.zero_grad()     every TASKS_PER_BATCH tasks
// as three multi-tensor kernels over ALL parameters of the model at once (one launch each, no host synchronisation:
// whether the step is skipped is decided on the device):
//   unscale_check   g *= 1 / scale in place, found_inf = any non-finite gradient           (GradScaler.unscale_)
//   adam            m, v, p update with bias correction, skipped as a whole when found_inf  (GradScaler.step -> Adam.step)
//   scaler_update   scale *= backoff on an overflow, *= growth after growth_interval clean steps   (GradScaler.update)
// HBM-bound elementwise work: 4 tensors read + 3 written per parameter element (28 bytes), coalesced, grid sized
// by 64 K-element chunks over the flattened parameter list.  State (exp_avg, exp_avg_sq, the step count) lives in the
// handle; parameters and gradients stay the caller's tensors.
#include <cstring>
#include <vector>

#include "../../include/clipspm_b200.h"
#include "api_common.cuh"
#include "profile.cuh"

struct spm_adam {
  int n = 0;
  std::vector<float*> params, m, v;
  std::vector<long long> numel;
  float** d_params = nullptr;   // device tables, one entry per tensor
  float** d_grads = nullptr;
  float** d_m = nullptr;
  float** d_v = nullptr;
  long long* d_numel = nullptr;
  int* d_chunk_tensor = nullptr;   // chunk -> tensor index
  long long* d_chunk_off = nullptr;
  int n_chunks = 0;
  float* d_step = nullptr;         // [n] steps taken PER PARAMETER (fp32 like torch's state['step']: a parameter without a
                                   // gradient is skipped and its count does not advance)
  float* d_found_inf = nullptr;    // used when no scaler state is passed (always 0)
  std::vector<float*> h_grads;
};

namespace spm {
namespace {

constexpr long long OPT_CHUNK = 65536;

__global__ void unscale_check_kernel(float* const* __restrict__ grads, const long long* __restrict__ numel,
                                     const int* __restrict__ chunk_tensor, const long long* __restrict__ chunk_off,
                                     float* __restrict__ scaler /* [scale, growth_tracker, found_inf] */) {
  const int t = chunk_tensor[blockIdx.x];
  float* g = grads[t];
  if (g == nullptr) return;   // a parameter without a gradient (torch skips it too)
  const long long o0 = chunk_off[blockIdx.x], o1 = min(o0 + OPT_CHUNK, numel[t]);
  const float inv = __fdiv_rn(1.f, scaler[0]);   // GradScaler.unscale_: inv_scale = scale.double().reciprocal().float()
  bool bad = false;
  for (long long i = o0 + threadIdx.x; i < o1; i += blockDim.x) {
    const float x = g[i] * inv;
    bad = bad || (__float_as_uint(x) & 0x7f800000u) == 0x7f800000u;   // inf / nan by bit pattern (no fast-math surprises)
    g[i] = x;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) scaler[2] = 1.f;   // benign race: every writer stores the same value
}

__global__ void adam_kernel(float* const* __restrict__ params, float* const* __restrict__ grads, float* const* __restrict__ ms,
                            float* const* __restrict__ vs, const long long* __restrict__ numel,
                            const int* __restrict__ chunk_tensor, const long long* __restrict__ chunk_off,
                            const float* __restrict__ step, const float* __restrict__ found_inf, double lr, double beta1,
                            double beta2, float eps, float weight_decay) {
  if (*found_inf != 0.f) return;   // GradScaler.step: an overflowed step is skipped as a whole
  const int t = chunk_tensor[blockIdx.x];
  const float* g = grads[t];
  if (g == nullptr) return;
  float* p = params[t];
  float* m = ms[t];
  float* v = vs[t];
  const long long o0 = chunk_off[blockIdx.x], o1 = min(o0 + OPT_CHUNK, numel[t]);
  // torch.optim.Adam (_single_tensor_adam, amsgrad=False, maximize=False): step counts from 1
  // (the scalars in double, as the Python floats they are there; the library is built with --use_fast_math, so the
  // per-element division and square root are the correctly rounded intrinsics)
  const double tt = (double)step[t] + 1.0;
  const double bc1 = 1.0 - pow(beta1, tt), bc2 = 1.0 - pow(beta2, tt);
  const float step_size = (float)(lr / bc1), sqrt_bc2 = (float)sqrt(bc2);
  // the lerp / addcmul weights as torch forms them: (1 - beta) in double, then rounded to fp32 once
  const float b2 = (float)beta2, omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
  for (long long i = o0 + threadIdx.x; i < o1; i += blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(weight_decay, pi, g[i]);
    const float mi = fmaf(omb1, gi - m[i], m[i]);        // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(omb2, gi * gi, b2 * v[i]);     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = __fdiv_rn(__fsqrt_rn(vi), sqrt_bc2) + eps;
    p[i] = pi - step_size * __fdiv_rn(mi, denom);
  }
}

// torch.optim.SGD(lr, momentum, weight_decay) (dampening 0, nesterov False; run/main_run.py:92-96): g += wd p;
// buf = g on a parameter's first step, momentum buf + g afterwards; p -= lr buf  (momentum 0: p -= lr g).  The momentum buffer is
// the handle's first state tensor.
__global__ void sgd_kernel(float* const* __restrict__ params, float* const* __restrict__ grads, float* const* __restrict__ bufs,
                           const long long* __restrict__ numel, const int* __restrict__ chunk_tensor,
                           const long long* __restrict__ chunk_off, const float* __restrict__ step,
                           const float* __restrict__ found_inf, float lr, float momentum, float weight_decay) {
  if (*found_inf != 0.f) return;
  const int t = chunk_tensor[blockIdx.x];
  const float* g = grads[t];
  if (g == nullptr) return;
  float* p = params[t];
  float* b = bufs[t];
  const bool first = step[t] == 0.f;
  const long long o0 = chunk_off[blockIdx.x], o1 = min(o0 + OPT_CHUNK, numel[t]);
  for (long long i = o0 + threadIdx.x; i < o1; i += blockDim.x) {
    const float pi = p[i];
    float gi = fmaf(weight_decay, pi, g[i]);
    if (momentum != 0.f) {
      gi = first ? gi : fmaf(momentum, b[i], gi);
      b[i] = gi;
    }
    p[i] = fmaf(-lr, gi, pi);
  }
}

__global__ void adam_bump_kernel(float* __restrict__ step, float* const* __restrict__ grads, int n,
                                 const float* __restrict__ found_inf) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && *found_inf == 0.f && grads[t] != nullptr) step[t] += 1.f;
}

// torch._amp_update_scale_: scaler = [scale, growth_tracker, found_inf]; found_inf is cleared for the next step
__global__ void scaler_update_kernel(float* __restrict__ scaler, float growth, float backoff, int growth_interval) {
  if (scaler[2] != 0.f) {
    scaler[0] *= backoff;
    scaler[1] = 0.f;
  } else {
    const float tr = scaler[1] + 1.f;
    if (tr >= (float)growth_interval) {
      const float ns = scaler[0] * growth;
      if ((__float_as_uint(ns) & 0x7f800000u) != 0x7f800000u) scaler[0] = ns;   // torch keeps the old scale when the grown one overflows
      scaler[1] = 0.f;
    } else {
      scaler[1] = tr;
    }
  }
  scaler[2] = 0.f;
}

template <typename T>
int to_device(T** dst, const std::vector<T>& src) {
  SPM_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), src.size() * sizeof(T)));
  SPM_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace
}  // namespace spm

using namespace spm;

extern "C" {

int spm_adam_create(int n_tensors, float* const* params, const long long* numel, spm_adam** out) {
  SPM_CHECK(n_tensors > 0 && params != nullptr && numel != nullptr && out != nullptr, "spm_adam_create: bad argument");
  int ndev = 0;
  SPM_CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0, "spm_adam_create: no CUDA device -- this library has no CPU path");
  spm_adam* a = new spm_adam();
  a->n = n_tensors;
  std::vector<int> chunk_tensor;
  std::vector<long long> chunk_off;
  for (int i = 0; i < n_tensors; ++i) {
    if (params[i] == nullptr || numel[i] <= 0) {
      delete a;
      set_error("spm_adam_create: null parameter or empty tensor");
      return 1;
    }
    a->params.push_back(params[i]);
    a->numel.push_back(numel[i]);
    float *m = nullptr, *v = nullptr;
    SPM_CUDA(cudaMalloc(reinterpret_cast<void**>(&m), (size_t)numel[i] * 4));
    SPM_CUDA(cudaMalloc(reinterpret_cast<void**>(&v), (size_t)numel[i] * 4));
    SPM_CUDA(cudaMemset(m, 0, (size_t)numel[i] * 4));
    SPM_CUDA(cudaMemset(v, 0, (size_t)numel[i] * 4));
    a->m.push_back(m);
    a->v.push_back(v);
    for (long long o = 0; o < numel[i]; o += OPT_CHUNK) { chunk_tensor.push_back(i); chunk_off.push_back(o); }
  }
  a->n_chunks = (int)chunk_tensor.size();
  SPM_TRY(to_device(&a->d_params, a->params));
  SPM_TRY(to_device(&a->d_m, a->m));
  SPM_TRY(to_device(&a->d_v, a->v));
  SPM_TRY(to_device(&a->d_numel, a->numel));
  SPM_TRY(to_device(&a->d_chunk_tensor, chunk_tensor));
  SPM_TRY(to_device(&a->d_chunk_off, chunk_off));
  SPM_CUDA(cudaMalloc(reinterpret_cast<void**>(&a->d_grads), (size_t)n_tensors * sizeof(float*)));
  SPM_CUDA(cudaMalloc(reinterpret_cast<void**>(&a->d_step), (size_t)n_tensors * 4));
  SPM_CUDA(cudaMalloc(reinterpret_cast<void**>(&a->d_found_inf), 4));
  SPM_CUDA(cudaMemset(a->d_step, 0, (size_t)n_tensors * 4));
  SPM_CUDA(cudaMemset(a->d_found_inf, 0, 4));
  a->h_grads.assign(n_tensors, nullptr);
  *out = a;
  return 0;
}

int spm_adam_destroy(spm_adam* a) {
  if (a == nullptr) return 0;
  for (float* p : a->m) cudaFree(p);
  for (float* p : a->v) cudaFree(p);
  cudaFree(a->d_params); cudaFree(a->d_grads); cudaFree(a->d_m); cudaFree(a->d_v); cudaFree(a->d_numel);
  cudaFree(a->d_chunk_tensor); cudaFree(a->d_chunk_off); cudaFree(a->d_step); cudaFree(a->d_found_inf);
  delete a;
  return 0;
}

int spm_adam_state(spm_adam* a, int i, float** exp_avg, float** exp_avg_sq, float** step) {
  SPM_CHECK(a != nullptr && i >= 0 && i < a->n, "spm_adam_state: bad argument");
  if (exp_avg) *exp_avg = a->m[i];
  if (exp_avg_sq) *exp_avg_sq = a->v[i];
  if (step) *step = a->d_step + i;
  return 0;
}

int spm_adam_step(spm_adam* a, void* stream, float* const* grads, double lr, double beta1, double beta2, double eps,
                  double weight_decay, float* scaler_state) {
  SPM_CHECK(a != nullptr && grads != nullptr, "spm_adam_step: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  bool same = true;
  for (int i = 0; i < a->n; ++i) same = same && a->h_grads[i] == grads[i];
  if (!same) {   // autograd may hand out new gradient tensors every iteration: refresh the table (stream-ordered)
    std::memcpy(a->h_grads.data(), grads, (size_t)a->n * sizeof(float*));
    SPM_CUDA(cudaMemcpyAsync(a->d_grads, a->h_grads.data(), (size_t)a->n * sizeof(float*), cudaMemcpyHostToDevice, st));
    SPM_CUDA(cudaStreamSynchronize(st));   // h_grads is pageable: the copy must not read it after a later overwrite
  }
  const float* found = a->d_found_inf;
  if (scaler_state != nullptr) {
    unscale_check_kernel<<<a->n_chunks, 256, 0, st>>>(a->d_grads, a->d_numel, a->d_chunk_tensor, a->d_chunk_off, scaler_state);
    SPM_CUDA(cudaGetLastError());
    count_launch();
    found = scaler_state + 2;
  }
  adam_kernel<<<a->n_chunks, 256, 0, st>>>(a->d_params, a->d_grads, a->d_m, a->d_v, a->d_numel, a->d_chunk_tensor, a->d_chunk_off,
                                          a->d_step, found, lr, beta1, beta2, (float)eps, (float)weight_decay);
  SPM_CUDA(cudaGetLastError());
  count_launch();
  adam_bump_kernel<<<(a->n + 255) / 256, 256, 0, st>>>(a->d_step, a->d_grads, a->n, found);
  SPM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int spm_sgd_step(spm_adam* a, void* stream, float* const* grads, double lr, double momentum, double weight_decay,
                 float* scaler_state) {
  SPM_CHECK(a != nullptr && grads != nullptr, "spm_sgd_step: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  bool same = true;
  for (int i = 0; i < a->n; ++i) same = same && a->h_grads[i] == grads[i];
  if (!same) {
    std::memcpy(a->h_grads.data(), grads, (size_t)a->n * sizeof(float*));
    SPM_CUDA(cudaMemcpyAsync(a->d_grads, a->h_grads.data(), (size_t)a->n * sizeof(float*), cudaMemcpyHostToDevice, st));
    SPM_CUDA(cudaStreamSynchronize(st));
  }
  const float* found = a->d_found_inf;
  if (scaler_state != nullptr) {
    unscale_check_kernel<<<a->n_chunks, 256, 0, st>>>(a->d_grads, a->d_numel, a->d_chunk_tensor, a->d_chunk_off, scaler_state);
    SPM_CUDA(cudaGetLastError());
    count_launch();
    found = scaler_state + 2;
  }
  sgd_kernel<<<a->n_chunks, 256, 0, st>>>(a->d_params, a->d_grads, a->d_m, a->d_numel, a->d_chunk_tensor, a->d_chunk_off, a->d_step,
                                         found, (float)lr, (float)momentum, (float)weight_decay);
  SPM_CUDA(cudaGetLastError());
  count_launch();
  adam_bump_kernel<<<(a->n + 255) / 256, 256, 0, st>>>(a->d_step, a->d_grads, a->n, found);
  SPM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int spm_scaler_update(void* stream, float* scaler_state, float growth_factor, float backoff_factor, int growth_interval) {
  SPM_CHECK(scaler_state != nullptr && growth_interval > 0, "spm_scaler_update: bad argument");
  scaler_update_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(scaler_state, growth_factor, backoff_factor, growth_interval);
  SPM_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // extern "C"
