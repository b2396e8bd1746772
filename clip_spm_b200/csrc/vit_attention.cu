// Frame-encoder self-attention: softmax(Q K^T / sqrt(64)) V, 197 tokens x 64 dims per (frame, head).
// Restates the attention inside nn.MultiheadAttention(768, 12) of models/clip_fsar.py:626,638.
//
// One CTA (4 warps) per (frame, head).  Q, K and V of that head (3 x 197 x 64 bf16 = 74 KB) are staged once in
// padded shared memory (row stride 144 B: conflict-free ldmatrix); every warp then owns whole 16-row query tiles:
// S = Q K^T for all 200 keys stays in registers (no online-softmax rescaling needed at this sequence length),
// row softmax with quad shuffles, P re-used in place as the A operand of P V.  The O tile is staged through the
// warp's own (dead) Q rows so the global store is 128-byte coalesced.
// Tensor-core path: mma.sync.m16n8k16 bf16 (4 % of the encoder FLOPs; the GEMMs carry the tcgen05 path).
#include "kernels.cuh"
#include "profile.cuh"

namespace spm {

namespace {
constexpr int L = 197;        // tokens per frame
constexpr int LP = 208;       // padded to 13 tiles of 16
constexpr int HD = 64;        // head dim
constexpr int RS = 72;        // smem row stride in elements (144 B)
constexpr int NT = 25;        // key tiles of 8 actually multiplied (200 >= 197)
constexpr int C = 768, C3 = 2304, HEADS = 12;
constexpr int SMEM_BYTES = 3 * LP * RS * 2;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
}  // namespace

__global__ void __launch_bounds__(128)
vit_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sK = sQ + LP * RS;
  __nv_bfloat16* sV = sK + LP * RS;
  const int frame = blockIdx.x / HEADS, head = blockIdx.x % HEADS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* base = qkv + (long long)frame * L * C3 + head * HD;

  // ---- stage Q, K, V (rows >= 197 are zero) : 3 x 208 rows x 8 chunks of 16 B
  for (int i = threadIdx.x; i < 3 * LP * 8; i += 128) {
    const int mat = i / (LP * 8);
    const int r = (i / 8) % LP;
    const int ch = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < L) v = __ldg(reinterpret_cast<const uint4*>(base + (long long)r * C3 + mat * C + ch * 8));
    *reinterpret_cast<uint4*>(sQ + (mat * LP + r) * RS + ch * 8) = v;
  }
  __syncthreads();

  const uint32_t sQ_u = (uint32_t)__cvta_generic_to_shared(sQ);
  const uint32_t sK_u = (uint32_t)__cvta_generic_to_shared(sK);
  const uint32_t sV_u = (uint32_t)__cvta_generic_to_shared(sV);
  const float scale_log2 = 0.125f * 1.4426950408889634f;

  for (int mt = warp; mt < LP / 16; mt += 4) {
    // ---- Q fragments (16 rows x 64 dims = 4 k-steps)
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t addr = sQ_u + (uint32_t)(((mt * 16 + (lane & 15)) * RS + ks * 16 + (lane >> 4) * 8) * 2);
      ldsm_x4(addr, qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    }
    // ---- S = Q K^T
    float s[NT + 1][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t b0, b1, b2, b3;
        const uint32_t addr = sK_u + (uint32_t)(((nt * 8 + (lane & 7)) * RS + kp * 32 + (lane >> 3) * 8) * 2);
        ldsm_x4(addr, b0, b1, b2, b3);
        mma_bf16(s[nt], qa[2 * kp], b0, b1);
        mma_bf16(s[nt], qa[2 * kp + 1], b2, b3);
      }
    }
    // ---- mask padded keys (197..199 live in tile 24), row max
    {
      const int k0 = 24 * 8 + 2 * (lane & 3);
      if (k0 >= L) { s[24][0] = -INFINITY; s[24][2] = -INFINITY; }
      if (k0 + 1 >= L) { s[24][1] = -INFINITY; s[24][3] = -INFINITY; }
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float mb0 = m0 * scale_log2, mb1 = m1 * scale_log2;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = exp2f(s[nt][0] * scale_log2 - mb0);
      s[nt][1] = exp2f(s[nt][1] * scale_log2 - mb0);
      s[nt][2] = exp2f(s[nt][2] * scale_log2 - mb1);
      s[nt][3] = exp2f(s[nt][3] * scale_log2 - mb1);
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
    s[NT][0] = s[NT][1] = s[NT][2] = s[NT][3] = 0.f;  // keys 200..207: pure padding
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // ---- O = P V
    float o[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < LP / 16; ++kt) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * kt][0], s[2 * kt][1]);
      pa[1] = pack_bf16(s[2 * kt][2], s[2 * kt][3]);
      pa[2] = pack_bf16(s[2 * kt + 1][0], s[2 * kt + 1][1]);
      pa[3] = pack_bf16(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b0, b1, b2, b3;
        const uint32_t addr =
            sV_u + (uint32_t)(((kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * RS + dp * 16 + (lane >> 4) * 8) * 2);
        ldsm_x4_t(addr, b0, b1, b2, b3);
        mma_bf16(o[2 * dp], pa, b0, b1);
        mma_bf16(o[2 * dp + 1], pa, b2, b3);
      }
    }
    // ---- normalise, stage in this warp's own Q rows, coalesced store
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    __syncwarp();
    {
      const int r0 = mt * 16 + (lane >> 2), cc = 2 * (lane & 3);
#pragma unroll
      for (int dn = 0; dn < 8; ++dn) {
        *reinterpret_cast<uint32_t*>(sQ + r0 * RS + dn * 8 + cc) = pack_bf16(o[dn][0] * inv0, o[dn][1] * inv0);
        *reinterpret_cast<uint32_t*>(sQ + (r0 + 8) * RS + dn * 8 + cc) = pack_bf16(o[dn][2] * inv1, o[dn][3] * inv1);
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane;  // 16 rows x 8 chunks
      const int r = mt * 16 + (idx >> 3), ch = idx & 7;
      if (r < L) {
        const uint4 v = *reinterpret_cast<const uint4*>(sQ + r * RS + ch * 8);
        *reinterpret_cast<uint4*>(out + ((long long)frame * L + r) * C + head * HD + ch * 8) = v;
      }
    }
  }
}

int k_vit_attention_init() {
  return (int)cudaFuncSetAttribute(vit_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
}

int k_vit_attention(cudaStream_t st, const __nv_bfloat16* qkv, __nv_bfloat16* out, int n_frames) {
  if (n_frames <= 0) return 0;
  vit_attention_kernel<<<n_frames * HEADS, 128, SMEM_BYTES, st>>>(qkv, out);
  cudaError_t e = cudaGetLastError();
  count_launch();
  return (int)e;
}

}  // namespace spm
