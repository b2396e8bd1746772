// Soft-DTW of TA2N (models/OTAM.py): the reference's only own GPU kernels are the two numba.cuda kernels
// compute_softdtw_cuda (:34-89) and compute_softdtw_backward_cuda (:93-130) behind _SoftDTWCUDA (:134-203).  These
// are their sm_100a counterparts behind the C ABI (spm_softdtw_forward / spm_softdtw_backward).
//
//   forward   R[i,j] = D[i-1,j-1] + softmin_gamma(R[i-1,j-1], R[i-1,j], R[i,j-1]),  R[0,0] = 0, other borders +inf,
//             cells with |i-j| > bandwidth > 0 stay +inf;  out = R[N,M]          (R is [B, N+2, M+2], kept for backward)
//   backward  E[i,j] = E[i+1,j] a + E[i,j+1] b + E[i+1,j+1] c,  a = exp((R[i+1,j] - R[i,j] - D[i+1,j]) / gamma) ...
//             with R's last row/column -inf, R[N+1,M+1] = R[N,M], E[N+1,M+1] = 1      (:160-170)
//
// One CTA per problem, thread = row, anti-diagonal = one step (the dependency structure admits nothing else), but the
// neighbours never touch global memory: three rolling diagonals live in shared memory (the numba kernels re-read
// R / E from global memory behind a block barrier at every step).  The backward applies the reference's in-place
// edits of R (:160-162, :111-112) on the fly instead of modifying the caller's tensor.
#include "head_kernels.cuh"
#include "profile.cuh"

namespace spm {

namespace {
#define INF __int_as_float(0x7f800000)

__global__ void softdtw_forward_kernel(const float* __restrict__ D, int N, int M, float gamma, float bandwidth,
                                       float* __restrict__ R, float* __restrict__ out) {
  extern __shared__ float sm_sd[];   // 3 rolling diagonals of N rows
  const int b = blockIdx.x, I = threadIdx.x, i = I + 1;
  const long long rs = M + 2;
  float* Rb = R + (long long)b * (N + 2) * rs;
  const float* Db = D + (long long)b * N * M;
  for (int k = threadIdx.x; k < (N + 2) * (M + 2); k += blockDim.x) Rb[k] = k == 0 ? 0.f : INF;
  float* d0 = sm_sd;            // diagonal p (being written)
  float* d1 = sm_sd + N;        // diagonal p-1
  float* d2 = sm_sd + 2 * N;    // diagonal p-2
  if (I < N) { d0[I] = INF; d1[I] = INF; d2[I] = INF; }
  __syncthreads();
  const float inv_gamma = 1.f / gamma;
  for (int p = 0; p < N + M - 1; ++p) {
    const int J = p - I, j = J + 1;
    if (I < N) {
      float v = INF;
      if (J >= 0 && J < M && !(fabsf((float)(i - j)) > bandwidth && bandwidth > 0.f)) {
        const float diag = I == 0 ? (J == 0 ? 0.f : INF) : d2[I - 1];   // R[i-1, j-1]
        const float up = I == 0 ? INF : d1[I - 1];                       // R[i-1, j]
        const float left = d1[I];                                         // R[i, j-1]
        const float r0 = -diag * inv_gamma, r1 = -up * inv_gamma, r2 = -left * inv_gamma;
        const float rmax = fmaxf(fmaxf(r0, r1), r2);
        const float rsum = expf(r0 - rmax) + expf(r1 - rmax) + expf(r2 - rmax);
        v = Db[(long long)I * M + J] - gamma * (logf(rsum) + rmax);
        Rb[(long long)i * rs + j] = v;
      }
      d0[I] = v;
    }
    __syncthreads();
    float* t = d2; d2 = d1; d1 = d0; d0 = t;
  }
  if (threadIdx.x == 0 && out != nullptr) out[b] = Rb[(long long)N * rs + M];
}

__global__ void softdtw_backward_kernel(const float* __restrict__ D, const float* __restrict__ R, int N, int M,
                                        float gamma, float bandwidth, float* __restrict__ E) {
  extern __shared__ float sm_sd[];
  const int b = blockIdx.x, I = threadIdx.x, i = I + 1;
  const long long rs = M + 2;
  const float* Rb = R + (long long)b * (N + 2) * rs;
  const float* Db = D + (long long)b * N * M;
  float* Eb = E + (long long)b * N * M;
  float* d0 = sm_sd;
  float* d1 = sm_sd + N;        // reverse diagonal p+1
  float* d2 = sm_sd + 2 * N;    // reverse diagonal p+2
  if (I < N) { d0[I] = 0.f; d1[I] = 0.f; d2[I] = 0.f; }
  __syncthreads();
  const float inv_gamma = 1.f / gamma;
  const float r_last = Rb[(long long)N * rs + M];
  // R as the reference edits it before / during the sweep: last row and column -inf, corner = R[N,M], +inf -> -inf
  auto Rm = [&](int ii, int jj) {
    if (ii == N + 1 || jj == M + 1) return (ii == N + 1 && jj == M + 1) ? r_last : -INF;
    const float r = Rb[(long long)ii * rs + jj];
    return isinf(r) ? -INF : r;
  };
  auto Dm = [&](int ii, int jj) { return (ii <= N && jj <= M) ? Db[(long long)(ii - 1) * M + (jj - 1)] : 0.f; };
  for (int p = N + M - 2; p >= 0; --p) {
    const int J = p - I, j = J + 1;
    if (I < N) {
      float v = 0.f;
      if (J >= 0 && J < M) {
        if (!(fabsf((float)(i - j)) > bandwidth && bandwidth > 0.f)) {
          const float r = Rm(i, j);
          const float e_down = I == N - 1 ? 0.f : d1[I + 1];                            // E[i+1, j]
          const float e_right = J == M - 1 ? 0.f : d1[I];                                // E[i, j+1]
          const float e_diag = (I == N - 1 || J == M - 1) ? ((I == N - 1 && J == M - 1) ? 1.f : 0.f) : d2[I + 1];
          const float a = expf((Rm(i + 1, j) - r - Dm(i + 1, j)) * inv_gamma);
          const float bb = expf((Rm(i, j + 1) - r - Dm(i, j + 1)) * inv_gamma);
          const float c = expf((Rm(i + 1, j + 1) - r - Dm(i + 1, j + 1)) * inv_gamma);
          v = e_down * a + e_right * bb + e_diag * c;
        }
        Eb[(long long)I * M + J] = v;
      }
      d0[I] = v;
    }
    __syncthreads();
    float* t = d2; d2 = d1; d1 = d0; d0 = t;
  }
}
#undef INF
}  // namespace

int k_softdtw_forward(cudaStream_t st, const float* D, int B, int N, int M, float gamma, float bandwidth, float* R,
                      float* out) {
  if (N < 1 || M < 1 || N > 1024 || M > 1024) return -2;   // one thread per row, like the reference (:352)
  if (B <= 0) return 0;
  const int threads = ((N + 31) / 32) * 32;
  softdtw_forward_kernel<<<B, threads, (size_t)3 * N * sizeof(float), st>>>(D, N, M, gamma, bandwidth, R, out);
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}

int k_softdtw_backward(cudaStream_t st, const float* D, const float* R, int B, int N, int M, float gamma,
                       float bandwidth, float* E) {
  if (N < 1 || M < 1 || N > 1024 || M > 1024) return -2;
  if (B <= 0) return 0;
  const int threads = ((N + 31) / 32) * 32;
  softdtw_backward_kernel<<<B, threads, (size_t)3 * N * sizeof(float), st>>>(D, R, N, M, gamma, bandwidth, E);
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}

}  // namespace spm
