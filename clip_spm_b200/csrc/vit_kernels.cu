// Bandwidth-bound kernels around the frame-encoder GEMMs: weight packing, patch im2col, LayerNorm.
// All are coalesced, 128-bit vectorised, one pass over their input.
#include "kernels.cuh"
#include "gemm.cuh"
#include "profile.cuh"

namespace spm {

#define SPM_LAUNCH_CHECK()                                   \
  do {                                                       \
    cudaError_t _e = cudaGetLastError();                     \
    if (_e != cudaSuccess) return (int)_e;                   \
    count_launch();                                          \
  } while (0)

// ------------------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __float2bfloat16_rn(in[i]);
}
int k_cast_bf16(cudaStream_t st, const float* in, __nv_bfloat16* out, long long n) {
  int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  cast_bf16_kernel<<<blocks, 256, 0, st>>>(in, out, n);
  SPM_LAUNCH_CHECK();
  return 0;
}

__global__ void transpose_cast_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < R && c < C) ? in[(long long)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < C) out[(long long)c * R + r] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}
int k_transpose_cast_bf16(cudaStream_t st, const float* in, __nv_bfloat16* out, int R, int C) {
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  transpose_cast_kernel<<<grid, block, 0, st>>>(in, out, R, C);
  SPM_LAUNCH_CHECK();
  return 0;
}

__global__ void repack_conv1d_kernel(const float* __restrict__ in, float* __restrict__ out, int O, int I) {
  const long long n = (long long)O * I * 3;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // out index: o*(3I) + kk*I + c
  const int c = (int)(i % I);
  const int kk = (int)((i / I) % 3);
  const int o = (int)(i / (3LL * I));
  out[i] = in[((long long)o * I + c) * 3 + kk];
}
int k_repack_conv1d(cudaStream_t st, const float* in, float* out, int O, int I) {
  const long long n = (long long)O * I * 3;
  repack_conv1d_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(in, out, O, I);
  SPM_LAUNCH_CHECK();
  return 0;
}

__global__ void add_vec_kernel(const float* a, const float* b, float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}
int k_add_vec(cudaStream_t st, const float* a, const float* b, float* out, int n) {
  add_vec_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, b, out, n);
  SPM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// patch im2col: one thread moves 8 horizontally adjacent pixels (32 B read, 16 B bf16 write).
// Thread order follows image memory order, so reads are perfectly coalesced; the writes land as 16-byte
// pieces (pairs of threads complete 32-byte sectors).
// ------------------------------------------------------------------------------------------------------
__global__ void patch_im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out,
                                    long long n_chunks) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_chunks) return;
  const int x8 = (int)(i % 28);
  const int y = (int)((i / 28) % 224);
  const int c = (int)((i / (28 * 224)) % 3);
  const long long f = i / (28LL * 224 * 3);
  const float4* p = reinterpret_cast<const float4*>(img + i * 8);
  const float4 a = __ldg(p), b = __ldg(p + 1);
  const int px = x8 >> 1, kx0 = (x8 & 1) * 8, py = y >> 4, ky = y & 15;
  const long long row = f * 196 + py * 14 + px;
  __nv_bfloat162 v0 = __floats2bfloat162_rn(a.x, a.y), v1 = __floats2bfloat162_rn(a.z, a.w);
  __nv_bfloat162 v2 = __floats2bfloat162_rn(b.x, b.y), v3 = __floats2bfloat162_rn(b.z, b.w);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&v0); u.y = *reinterpret_cast<uint32_t*>(&v1);
  u.z = *reinterpret_cast<uint32_t*>(&v2); u.w = *reinterpret_cast<uint32_t*>(&v3);
  *reinterpret_cast<uint4*>(out + row * 768 + c * 256 + ky * 16 + kx0) = u;
}
int k_patch_im2col(cudaStream_t st, const float* images, __nv_bfloat16* patches, int n_frames) {
  const long long n_chunks = (long long)n_frames * 3 * 224 * 28;
  patch_im2col_kernel<<<(unsigned)((n_chunks + 255) / 256), 256, 0, st>>>(images, patches, n_chunks);
  SPM_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (C/128 float4 per lane), exact two-pass
// mean / variance in fp32 like the reference.
// ------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ in, long long in_stride, int rows, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ cls_row, int cls_period,
                 float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, long long out_stride, int reverse,
                 float* __restrict__ stats_out) {
  constexpr int C = NV * 128;
  int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  if (reverse) row = rows - 1 - row;
  const float* src = in + (long long)row * in_stride;
  if (cls_period > 0 && (row % cls_period) == 0) src = cls_row;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = *reinterpret_cast<const float4*>(src + (i * 32 + lane) * 4);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.f / C) + 1e-5f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + (long long)row * out_stride + col) = o;
    if (out_bf16 != nullptr) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(o.x, o.y), p1 = __floats2bfloat162_rn(o.z, o.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(out_bf16 + (long long)row * out_stride + col) = u;
    }
    v[i] = o;
  }
  if (stats_out != nullptr) {
    // {sum y, sum y^2} of the OUTPUT row in partial-sum slot 0 (the other slots zero): statistics of the LayerNorm folded
    // into the GEMM that reads this row next (GemmEpilogue::ln_stats_in)
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      t1 += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      t2 = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, t2))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    if (lane < LN_FOLD_SLOTS)
      *reinterpret_cast<float2*>(stats_out + ((long long)row * LN_FOLD_SLOTS + lane) * 2) =
          lane == 0 ? make_float2(t1, t2) : make_float2(0.f, 0.f);
  }
}

// LayerNorm folded into the GEMM that follows it: W'[n,k] = bf16(W[n,k] * gamma[k]), colsum[n] = sum_k W'[n,k] (of the
// ROUNDED weights, so that acc - mean * colsum == sum_k (x_k - mean) W'[n,k] exactly), bias'[n] = bias[n] + sum_k beta[k] W[n,k]
__global__ void __launch_bounds__(256)
fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
               const float* __restrict__ bias, int N, int K, __nv_bfloat16* __restrict__ Wf, float* __restrict__ colsum,
               float* __restrict__ bias2, int center) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float c = 0.f, b = 0.f, wm = 0.f;
  if (center) {   // centred rows: sum_k W'[n,k] = 0 in fp32, so the mean term of the LayerNorm drops out of the product
    for (int k = lane; k < K; k += 32) wm += W[(long long)n * K + k] * gamma[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wm += __shfl_xor_sync(0xffffffffu, wm, o);
    wm /= (float)K;
  }
  for (int k = lane; k < K; k += 32) {
    const float w = W[(long long)n * K + k];
    const __nv_bfloat16 wf = __float2bfloat16_rn(w * gamma[k] - wm);
    Wf[(long long)n * K + k] = wf;
    c += __bfloat162float(wf);
    b = fmaf(beta[k], w, b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c += __shfl_xor_sync(0xffffffffu, c, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) { colsum[n] = c; bias2[n] = bias[n] + b; }
}
int k_fold_ln(cudaStream_t st, const float* W, const float* gamma, const float* beta, const float* bias, int N, int K,
              __nv_bfloat16* Wf, float* colsum, float* bias2, int center) {
  fold_ln_kernel<<<(N + 7) / 8, 256, 0, st>>>(W, gamma, beta, bias, N, K, Wf, colsum, bias2, center);
  SPM_LAUNCH_CHECK();
  return 0;
}

// same LayerNorm on a bf16 input row (the bf16 residual stream of SPM_PRECISION_BF16_RESID): 8 elements per 16-byte load,
// statistics in fp32 exactly as above
template <int NV8>
__global__ void __launch_bounds__(256)
layernorm_bf16in_kernel(const __nv_bfloat16* __restrict__ in, long long in_stride, int rows, const float* __restrict__ gamma,
                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ out_bf16, long long out_stride,
                        int reverse) {
  constexpr int C = NV8 * 256;
  int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  if (reverse) row = rows - 1 - row;
  const __nv_bfloat16* src = in + (long long)row * in_stride;
  float v[NV8][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + (i * 32 + lane) * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
      v[i][2 * e] = f.x; v[i][2 * e + 1] = f.y;
      s += f.x + f.y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV8; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float d = v[i][e] - mean; q += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.f / C) + 1e-5f);
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    const int col = (i * 32 + lane) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col)), b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat162 p = __floats2bfloat162_rn((v[i][2 * e] - mean) * rstd * g[2 * e] + b[2 * e],
                                               (v[i][2 * e + 1] - mean) * rstd * g[2 * e + 1] + b[2 * e + 1]);
      o[e] = *reinterpret_cast<uint32_t*>(&p);
    }
    *reinterpret_cast<uint4*>(out_bf16 + (long long)row * out_stride + col) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

int k_layernorm_bf16in(cudaStream_t st, const __nv_bfloat16* in, long long in_stride, int rows, int C, const float* gamma,
                       const float* beta, __nv_bfloat16* out_bf16, long long out_stride, int reverse) {
  if (rows <= 0) return 0;
  if (C != 768) return -1;
  layernorm_bf16in_kernel<3><<<(rows + 7) / 8, 256, 0, st>>>(in, in_stride, rows, gamma, beta, out_bf16, out_stride, reverse);
  SPM_LAUNCH_CHECK();
  return 0;
}

int k_layernorm(cudaStream_t st, const float* in, long long in_stride, int rows, int C, const float* gamma,
                const float* beta, const float* cls_row, int cls_period, float* out_f32, __nv_bfloat16* out_bf16,
                long long out_stride, int reverse, float* stats_out) {
  const int blocks = (rows + 7) / 8;
  if (rows <= 0) return 0;
#define SPM_LN(NV)                                                                                             \
  layernorm_kernel<NV><<<blocks, 256, 0, st>>>(in, in_stride, rows, gamma, beta, cls_row, cls_period, out_f32, \
                                               out_bf16, out_stride, reverse, stats_out)
  switch (C) {
    case 512: SPM_LN(4); break;
    case 768: SPM_LN(6); break;
    case 1024: SPM_LN(8); break;
    case 2048: SPM_LN(16); break;
    default: return -1;
  }
#undef SPM_LN
  SPM_LAUNCH_CHECK();
  return 0;
}

}  // namespace spm
