// fp32 path of SPM_PRECISION_FP32 (parity mode, north_star tolerance 1e-4): CUDA-core FFMA kernels with exactly the
// reference's fp32 arithmetic (there is no fp32 tensor-core MMA on Blackwell; SURVEY.md section 7 "hard parts").
//   sgemm_f32_kernel          out[orow(m), n] = act(sum_k A[m,k] B[n,k] + bias[n]) (+ residual)  -- same GemmEpilogue
//                             contract as the tcgen05 GEMM (row maps, positional residual, activations)
//   vit_attention_f32_kernel  softmax(q k^T / 8) v, 197 tokens x 64 dims per (frame, head), fp32 in / out
// Throughput is irrelevant here (tens of TFLOP/s); the bf16 tcgen05 path is the product path that bench.py measures.
#include "gemm.cuh"
#include "gemm_epilogue.cuh"
#include "kernels.cuh"
#include "profile.cuh"

namespace spm {

namespace {
constexpr int SBM = 64, SBN = 64, SBK = 16;

__global__ void __launch_bounds__(256)
sgemm_f32_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb, int M, int N,
                 int K, GemmEpilogue ep) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;  // loader: row 0..63, k offset 0,4,8,12
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += SBK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (m0 + lr < M && k0 + lk < K) a = *reinterpret_cast<const float4*>(A + (long long)(m0 + lr) * lda + k0 + lk);
    if (n0 + lr < N && k0 + lk < K) b = *reinterpret_cast<const float4*>(B + (long long)(n0 + lr) * ldb + k0 + lk);
    As[lk + 0][lr] = a.x; As[lk + 1][lr] = a.y; As[lk + 2][lr] = a.z; As[lk + 3][lr] = a.w;
    Bs[lk + 0][lr] = b.x; Bs[lk + 1][lr] = b.y; Bs[lk + 2][lr] = b.z; Bs[lk + 3][lr] = b.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const long long orow = out_row(ep, m), rrow = res_row(ep, m);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (ep.bias != nullptr) v += ep.bias[n];
      if (ep.act == ACT_QUICKGELU) v = v / (1.f + expf(-1.702f * v));  // exact form in the parity mode
      else if (ep.act == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
      else v = apply_act(v, ep.act, ep.slope);
      if (ep.residual != nullptr) v += ep.residual[rrow * ep.ldr + n];
      reinterpret_cast<float*>(ep.out)[orow * ep.ldo + n] = v;
    }
  }
}

constexpr int AL = 197, AHD = 64, AHEADS = 12, AC = 768;
constexpr int ATT32_SMEM = (AL * (AHD + 1) + AL * AHD + 8 * 4 * AHD) * 4;

// A warp handles FOUR query rows per pass: the lane's 7 key columns x 4 rows form a register tile, so a step of the dot
// products costs 4 broadcast + 7 strided shared-memory reads per 28 FMAs (one row at a time paid 2 reads per FMA).
__global__ void __launch_bounds__(256)
vit_attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out) {
  extern __shared__ float sm_a32[];
  float* sK = sm_a32;                  // [197][65]
  float* sV = sK + AL * (AHD + 1);     // [197][64]
  float* sQ = sV + AL * AHD;           // [8 warps][4 rows][64]
  const int frame = blockIdx.x / AHEADS, head = blockIdx.x % AHEADS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + (long long)frame * AL * (3 * AC) + head * AHD;
  for (int i = threadIdx.x; i < AL * AHD; i += blockDim.x) {
    const int r = i / AHD, d = i % AHD;
    sK[r * (AHD + 1) + d] = base[(long long)r * (3 * AC) + AC + d];
    sV[r * AHD + d] = base[(long long)r * (3 * AC) + 2 * AC + d];
  }
  __syncthreads();
  float* q = sQ + warp * 4 * AHD;
  int col[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) col[j] = min(lane + 32 * j, AL - 1) * (AHD + 1);
  for (int g = warp; g < (AL + 3) / 4; g += 8) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = min(4 * g + r, AL - 1);
      q[r * AHD + lane] = base[(long long)row * (3 * AC) + lane];
      q[r * AHD + lane + 32] = base[(long long)row * (3 * AC) + lane + 32];
    }
    __syncwarp();
    float s[4][7];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 7; ++j) s[r][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < AHD; ++d) {
      float a[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = q[r * AHD + d];
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const float k = sK[col[j] + d];
#pragma unroll
        for (int r = 0; r < 4; ++r) s[r][j] = fmaf(a[r], k, s[r][j]);
      }
    }
    float inv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        s[r][j] = (lane + 32 * j < AL) ? s[r][j] * 0.125f : -INFINITY;
        mx = fmaxf(mx, s[r][j]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        s[r][j] = (lane + 32 * j < AL) ? expf(s[r][j] - mx) : 0.f;
        l += s[r][j];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
      inv[r] = 1.f / l;
    }
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int nsrc = j < 6 ? 32 : AL - 192;
      for (int src = 0; src < nsrc; ++src) {
        const int key = src + 32 * j;
        const float v0 = sV[key * AHD + lane], v1 = sV[key * AHD + lane + 32];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float p = __shfl_sync(0xffffffffu, s[r][j], src);
          a0[r] = fmaf(p, v0, a0[r]);
          a1[r] = fmaf(p, v1, a1[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (4 * g + r < AL) {
        float* o = out + ((long long)frame * AL + 4 * g + r) * AC + head * AHD;
        o[lane] = a0[r] * inv[r];
        o[lane + 32] = a1[r] * inv[r];
      }
    __syncwarp();
  }
}

// fp32 patch im2col: images [F,3,224,224] -> patches [F*196, 768] fp32, column = c*256 + ky*16 + kx
__global__ void patch_im2col_f32_kernel(const float* __restrict__ img, float* __restrict__ out, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int x4 = (int)(i % 56);
  const int y = (int)((i / 56) % 224);
  const int c = (int)((i / (56 * 224)) % 3);
  const long long f = i / (56LL * 224 * 3);
  const float4 v = __ldg(reinterpret_cast<const float4*>(img) + i);
  const int px = x4 >> 2, kx0 = (x4 & 3) * 4, py = y >> 4, ky = y & 15;
  *reinterpret_cast<float4*>(out + (f * 196 + py * 14 + px) * 768 + c * 256 + ky * 16 + kx0) = v;
}
}  // namespace

int sgemm_f32_run(const GemmOp* op, cudaStream_t st) {
  dim3 grid((op->N + SBN - 1) / SBN, (op->M + SBM - 1) / SBM);
  sgemm_f32_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(op->simt_a), op->simt_lda,
                                         static_cast<const float*>(op->simt_b), op->simt_ldb, op->M, op->N, op->K, op->ep);
  return (int)cudaGetLastError();
}

int k_vit_attention_f32_init() {
  return (int)cudaFuncSetAttribute(vit_attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT32_SMEM);
}
int k_vit_attention_f32(cudaStream_t st, const float* qkv, float* out, int n_frames) {
  if (n_frames <= 0) return 0;
  vit_attention_f32_kernel<<<n_frames * AHEADS, 256, ATT32_SMEM, st>>>(qkv, out);
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}
int k_patch_im2col_f32(cudaStream_t st, const float* images, float* patches, int n_frames) {
  const long long n4 = (long long)n_frames * 3 * 224 * 56;
  patch_im2col_f32_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(images, patches, n4);
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}

}  // namespace spm
