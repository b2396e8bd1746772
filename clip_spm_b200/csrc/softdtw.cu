// Soft-DTW of TA2N (models/OTAM.py): sm_100a counterparts of the reference's two numba.cuda kernels behind _SoftDTWCUDA
// (:134-203), exposed as spm_softdtw_forward / spm_softdtw_backward.
//
//   forward   R[i,j] = D[i-1,j-1] + softmin_gamma(R[i-1,j-1], R[i-1,j], R[i,j-1]),  R[0,0] = 0, other borders +inf,
//             cells with |i-j| > bandwidth > 0 stay +inf;  out = R[N,M]          (R is [B, N+2, M+2], kept for backward)
//   backward  E[i,j] = E[i+1,j] a + E[i,j+1] b + E[i+1,j+1] c,  a = exp((R[i+1,j] - R[i,j] - D[i+1,j]) / gamma) ...
//             with R's last row/column -inf, R[N+1,M+1] = R[N,M], E[N+1,M+1] = 1      (:160-170)
//
// Design (r02, replacing the block-per-problem / thread-per-row / barrier-per-diagonal scheme): the recurrence runs as a
// REGISTER wavefront inside one warp, the same construction as the OTAM wavefront of otam_dp.cuh --
//   * a lane owns a COLUMN; on diagonal step k it holds cell (k - column, column).  Its vertical neighbour is its own
//     previous value, the horizontal and diagonal ones arrive from lane - 1 (forward) / lane + 1 (backward) by warp
//     shuffles of the last two values.  No shared-memory diagonals, no block barrier anywhere in the sweep.
//   * M <= 32: a warp runs 32 / M problems side by side (four at TA2N's 8 x 8); their D, R (and E) tables are staged in
//     shared memory with coalesced copies -- consecutive problems are contiguous in HBM -- so every table byte crosses
//     HBM exactly once, in full lines.
//   * longer sequences: one warp per problem sweeps the table in strips of 32 columns; the column shared by two
//     strips is handed over through a per-warp shared-memory array (R for the forward, E for the backward).
// The soft-min is taken min-shifted (all exponents <= 0): same value as the reference's max-shifted form (:44-49).
// The backward applies the reference's in-place edits of R (:160-162, :111-112) on the fly; the caller's R is const.
#include "head_kernels.cuh"
#include "profile.cuh"

namespace spm {

namespace {
#define SDTW_INF __int_as_float(0x7f800000)
constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS = 4;   // warps per CTA (each warp is independent: the kernels contain no block barrier)

__device__ __forceinline__ float softmin3(float a, float b, float c, float gamma, float inv_gamma) {
  const float mn = fminf(fminf(a, b), c);
  if (mn == SDTW_INF) return SDTW_INF;
  const float s = expf((mn - a) * inv_gamma) + expf((mn - b) * inv_gamma) + expf((mn - c) * inv_gamma);
  return mn - gamma * logf(s);
}
__device__ __forceinline__ bool pruned(int I, int J, float bandwidth) {
  return bandwidth > 0.f && fabsf((float)(I - J)) > bandwidth;
}

// R as the reference's backward sees it (edits applied on the fly): last row / column -inf, the corner = R[N,M],
// +inf entries (never reached or pruned) -> -inf.  `tab` is a [N+2][M+2] table (shared or global).
__device__ __forceinline__ float r_edit(const float* tab, int ii, int jj, int N, int M, float r_last) {
  if (ii == N + 1 || jj == M + 1) return (ii == N + 1 && jj == M + 1) ? r_last : -SDTW_INF;
  const float r = tab[ii * (M + 2) + jj];
  return isinf(r) ? -SDTW_INF : r;
}
__device__ __forceinline__ float d_pad(const float* d, int ii, int jj, int N, int M) {   // 1-based, zero outside
  return (ii <= N && jj <= M) ? d[(ii - 1) * M + (jj - 1)] : 0.f;
}
__device__ __forceinline__ float e_cell(const float* rtab, const float* d, int i, int j, int N, int M, float r_last,
                                        float inv_gamma, float e_down, float e_right, float e_diag) {
  const float r = r_edit(rtab, i, j, N, M, r_last);
  const float a = expf((r_edit(rtab, i + 1, j, N, M, r_last) - r - d_pad(d, i + 1, j, N, M)) * inv_gamma);
  const float b = expf((r_edit(rtab, i, j + 1, N, M, r_last) - r - d_pad(d, i, j + 1, N, M)) * inv_gamma);
  const float c = expf((r_edit(rtab, i + 1, j + 1, N, M, r_last) - r - d_pad(d, i + 1, j + 1, N, M)) * inv_gamma);
  return e_down * a + e_right * b + e_diag * c;
}

// ---------------------------------------------------------------------------------------------------------------
// M <= 32: 32 / M problems per warp, tables in shared memory
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * WARPS)
softdtw_fwd_small_kernel(const float* __restrict__ D, int B, int N, int M, float gamma, float bandwidth,
                         float* __restrict__ R, float* __restrict__ out) {
  extern __shared__ float sm_sd[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 32 / M, tabR = (N + 2) * (M + 2), tabD = N * M;
  const long long b0 = ((long long)blockIdx.x * WARPS + warp) * per_warp;
  if (b0 >= B) return;
  const int nb = (int)min((long long)per_warp, B - b0);
  float* sR = sm_sd + warp * per_warp * (tabR + tabD);
  float* sD = sR + per_warp * tabR;
  for (int i = lane; i < nb * tabD; i += 32) sD[i] = __ldg(D + b0 * tabD + i);
  for (int i = lane; i < nb * tabR; i += 32) sR[i] = (i % tabR == 0) ? 0.f : SDTW_INF;
  __syncwarp();
  const int seg = lane / M, c = lane - seg * M;
  const bool valid = seg < nb;
  const float* myD = sD + (valid ? seg : 0) * tabD;
  float* myR = sR + (valid ? seg : 0) * tabR;
  const float inv_gamma = 1.f / gamma;
  float v1 = SDTW_INF, v2 = SDTW_INF;   // this column's last / second-to-last cell
  for (int k = 0; k < N + M - 1; ++k) {
    float left = __shfl_up_sync(FULL, v1, 1);   // R[i, j-1]   (lane - 1, one step ago)
    float diag = __shfl_up_sync(FULL, v2, 1);   // R[i-1, j-1] (lane - 1, two steps ago)
    const int I = k - c;
    float up = v1;                              // R[i-1, j]
    if (c == 0) { left = SDTW_INF; diag = I == 0 ? 0.f : SDTW_INF; }
    if (I == 0) { up = SDTW_INF; if (c != 0) diag = SDTW_INF; }
    if (valid && I >= 0 && I < N) {
      float v = SDTW_INF;
      if (!pruned(I, c, bandwidth)) v = myD[I * M + c] + softmin3(diag, up, left, gamma, inv_gamma);
      myR[(I + 1) * (M + 2) + c + 1] = v;
      v2 = v1; v1 = v;
    }
  }
  __syncwarp();
  for (int i = lane; i < nb * tabR; i += 32) R[b0 * tabR + i] = sR[i];
  if (out != nullptr && lane < nb) out[b0 + lane] = sR[lane * tabR + N * (M + 2) + M];
}

__global__ void __launch_bounds__(32 * WARPS)
softdtw_bwd_small_kernel(const float* __restrict__ D, const float* __restrict__ R, int B, int N, int M, float gamma,
                         float bandwidth, float* __restrict__ E) {
  extern __shared__ float sm_sd[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = 32 / M, tabR = (N + 2) * (M + 2), tabD = N * M;
  const long long b0 = ((long long)blockIdx.x * WARPS + warp) * per_warp;
  if (b0 >= B) return;
  const int nb = (int)min((long long)per_warp, B - b0);
  float* sR = sm_sd + warp * per_warp * (tabR + 2 * tabD);
  float* sD = sR + per_warp * tabR;
  float* sE = sD + per_warp * tabD;
  for (int i = lane; i < nb * tabD; i += 32) { sD[i] = __ldg(D + b0 * tabD + i); sE[i] = 0.f; }
  for (int i = lane; i < nb * tabR; i += 32) sR[i] = __ldg(R + b0 * tabR + i);
  __syncwarp();
  const int seg = lane / M, c = lane - seg * M;
  const bool valid = seg < nb;
  const float* myD = sD + (valid ? seg : 0) * tabD;
  const float* myR = sR + (valid ? seg : 0) * tabR;
  float* myE = sE + (valid ? seg : 0) * tabD;
  const float inv_gamma = 1.f / gamma, r_last = myR[N * (M + 2) + M];
  float v1 = 0.f, v2 = 0.f;
  for (int k = N + M - 2; k >= 0; --k) {
    float right = __shfl_down_sync(FULL, v1, 1);   // E[i, j+1]   (lane + 1, one step ago)
    float diag = __shfl_down_sync(FULL, v2, 1);    // E[i+1, j+1] (lane + 1, two steps ago)
    const int I = k - c;
    float down = v1;                               // E[i+1, j]
    if (c == M - 1) { right = 0.f; diag = I == N - 1 ? 1.f : 0.f; }
    if (I == N - 1) { down = 0.f; if (c != M - 1) diag = 0.f; }
    if (valid && I >= 0 && I < N) {
      float v = 0.f;
      if (!pruned(I, c, bandwidth)) v = e_cell(myR, myD, I + 1, c + 1, N, M, r_last, inv_gamma, down, right, diag);
      myE[I * M + c] = v;
      v2 = v1; v1 = v;
    }
  }
  __syncwarp();
  for (int i = lane; i < nb * tabD; i += 32) E[b0 * tabD + i] = sE[i];
}

// ---------------------------------------------------------------------------------------------------------------
// any N, M <= 1024: one warp per problem, strips of 32 columns, boundary column handed over through shared memory
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * WARPS)
softdtw_fwd_strip_kernel(const float* __restrict__ D, int B, int N, int M, float gamma, float bandwidth,
                         float* __restrict__ R, float* __restrict__ out) {
  extern __shared__ float sm_sd[];   // per warp: two boundary columns of N values
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * WARPS + warp;
  if (b >= B) return;
  float* bcur = sm_sd + warp * 2 * N;
  float* bnxt = bcur + N;
  const int rs = M + 2;
  float* Rb = R + b * (long long)(N + 2) * rs;
  const float* Db = D + b * (long long)N * M;
  for (int i = lane; i < (N + 2) * rs; i += 32) Rb[i] = i == 0 ? 0.f : SDTW_INF;
  __syncwarp();
  const float inv_gamma = 1.f / gamma;
  float last = SDTW_INF;
  for (int c0 = 0; c0 < M; c0 += 32) {
    const int Wd = min(32, M - c0), c = lane, J = c0 + c;
    float v1 = SDTW_INF, v2 = SDTW_INF;
    float dn = (c < Wd && c == 0) ? __ldg(Db + J) : 0.f;   // D of this lane's next cell, fetched one step ahead
    for (int k = 0; k < N + Wd - 1; ++k) {
      float left = __shfl_up_sync(FULL, v1, 1), diag = __shfl_up_sync(FULL, v2, 1);
      const int I = k - c;
      const bool active = c < Wd && I >= 0 && I < N;
      const float d = dn;
      const int In = I + 1;   // the cell this lane computes at step k + 1
      if (c < Wd && In >= 0 && In < N) dn = __ldg(Db + (long long)In * M + J);
      float up = v1;
      if (c == 0) {
        left = c0 == 0 ? SDTW_INF : bcur[min(max(I, 0), N - 1)];   // (clamped: lanes past the table are inactive)
        diag = I <= 0 ? ((I == 0 && c0 == 0) ? 0.f : SDTW_INF) : (c0 == 0 ? SDTW_INF : bcur[min(I - 1, N - 1)]);
      }
      if (I == 0) { up = SDTW_INF; if (c != 0) diag = SDTW_INF; }
      if (active) {
        float v = SDTW_INF;
        if (!pruned(I, J, bandwidth)) {
          v = d + softmin3(diag, up, left, gamma, inv_gamma);
          Rb[(long long)(I + 1) * rs + J + 1] = v;
        }
        if (c == Wd - 1) bnxt[I] = v;
        v2 = v1; v1 = v;
      }
    }
    if (c0 + Wd == M) last = __shfl_sync(FULL, v1, Wd - 1);
    __syncwarp();
    float* t = bcur; bcur = bnxt; bnxt = t;
  }
  if (lane == 0 && out != nullptr) out[b] = last;
}

__global__ void __launch_bounds__(32 * WARPS)
softdtw_bwd_strip_kernel(const float* __restrict__ D, const float* __restrict__ R, int B, int N, int M, float gamma,
                         float bandwidth, float* __restrict__ E) {
  extern __shared__ float sm_sd[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * WARPS + warp;
  if (b >= B) return;
  float* bcur = sm_sd + warp * 2 * N;   // E[., first column of the strip to the right]
  float* bnxt = bcur + N;
  const float* Rb = R + b * (long long)(N + 2) * (M + 2);
  const float* Db = D + b * (long long)N * M;
  float* Eb = E + b * (long long)N * M;
  const float inv_gamma = 1.f / gamma, r_last = Rb[(long long)N * (M + 2) + M];
  const int n_strips = (M + 31) / 32;
  for (int s = n_strips - 1; s >= 0; --s) {
    const int c0 = 32 * s, Wd = min(32, M - c0), c = lane, J = c0 + c;
    const bool last_strip = s == n_strips - 1;
    float v1 = 0.f, v2 = 0.f;
    for (int k = N + Wd - 2; k >= 0; --k) {
      float right = __shfl_down_sync(FULL, v1, 1), diag = __shfl_down_sync(FULL, v2, 1);
      const int I = k - c;
      float down = v1;
      if (c == Wd - 1) {
        if (last_strip) { right = 0.f; diag = I == N - 1 ? 1.f : 0.f; }
        else { right = bcur[min(max(I, 0), N - 1)]; diag = (I >= 0 && I + 1 < N) ? bcur[I + 1] : 0.f; }
      }
      if (I == N - 1) { down = 0.f; if (!(c == Wd - 1 && last_strip)) diag = 0.f; }
      if (c < Wd && I >= 0 && I < N) {
        float v = 0.f;
        if (!pruned(I, J, bandwidth)) v = e_cell(Rb, Db, I + 1, J + 1, N, M, r_last, inv_gamma, down, right, diag);
        Eb[(long long)I * M + J] = v;
        if (c == 0) bnxt[I] = v;
        v2 = v1; v1 = v;
      }
    }
    __syncwarp();
    float* t = bcur; bcur = bnxt; bnxt = t;
  }
}
#undef SDTW_INF

// shared memory a warp needs on the small path; 0 = use the strip kernels
size_t small_bytes_per_warp(int N, int M, bool backward) {
  if (M > 32) return 0;
  const size_t per_warp = 32 / M, tabR = (size_t)(N + 2) * (M + 2), tabD = (size_t)N * M;
  const size_t bytes = per_warp * (tabR + (backward ? 2 : 1) * tabD) * sizeof(float);
  return bytes <= 11 * 1024 ? bytes : 0;   // 4 warps stay below the 48 KB default dynamic shared-memory limit
}
}  // namespace

int k_softdtw_forward(cudaStream_t st, const float* D, int B, int N, int M, float gamma, float bandwidth, float* R,
                      float* out) {
  if (N < 1 || M < 1 || N > 1024 || M > 1024) return -2;
  if (B <= 0) return 0;
  if (const size_t per_warp_bytes = small_bytes_per_warp(N, M, false)) {
    const long long warps = ((long long)B + 32 / M - 1) / (32 / M);
    softdtw_fwd_small_kernel<<<(unsigned)((warps + WARPS - 1) / WARPS), 32 * WARPS, WARPS * per_warp_bytes, st>>>(
        D, B, N, M, gamma, bandwidth, R, out);
  } else {
    softdtw_fwd_strip_kernel<<<(unsigned)((B + WARPS - 1) / WARPS), 32 * WARPS, (size_t)WARPS * 2 * N * sizeof(float), st>>>(
        D, B, N, M, gamma, bandwidth, R, out);
  }
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}

int k_softdtw_backward(cudaStream_t st, const float* D, const float* R, int B, int N, int M, float gamma,
                       float bandwidth, float* E) {
  if (N < 1 || M < 1 || N > 1024 || M > 1024) return -2;
  if (B <= 0) return 0;
  if (const size_t per_warp_bytes = small_bytes_per_warp(N, M, true)) {
    const long long warps = ((long long)B + 32 / M - 1) / (32 / M);
    softdtw_bwd_small_kernel<<<(unsigned)((warps + WARPS - 1) / WARPS), 32 * WARPS, WARPS * per_warp_bytes, st>>>(
        D, R, B, N, M, gamma, bandwidth, E);
  } else {
    softdtw_bwd_strip_kernel<<<(unsigned)((B + WARPS - 1) / WARPS), 32 * WARPS, (size_t)WARPS * 2 * N * sizeof(float), st>>>(
        D, R, B, N, M, gamma, bandwidth, E);
  }
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}

}  // namespace spm
