// Epilogue of the tcgen05 GEMMs, shared by the 1-CTA (gemm_tcgen05.cu) and 2-CTA (gemm2_tcgen05.cu) kernels:
// one warp drains its 32 accumulator rows (TMEM lane quarter) for the column groups it owns.
#pragma once
#include "gemm.cuh"
#include "ptx.cuh"

namespace spm {

// TMA prefetch of one box into L2 (no shared-memory destination, no barrier): used by the producer warps to run a
// few k-blocks ahead of the smem pipeline so the real loads hit L2 instead of paying HBM latency.
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1)
               : "memory");
}
constexpr int GEMM_L2_PREFETCH_KB = 8;  // k-blocks of A prefetched ahead of the TMA load pointer

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// x * sigmoid(1.702 x) = x * (0.5 + 0.5 * tanh(0.851 x)): one MUFU op (tanh.approx, rel. error ~2^-11, far below
// the bf16 rounding of the stored activation) instead of ex2 + rcp
__device__ __forceinline__ float quick_gelu_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  return x * fmaf(0.5f, t, 0.5f);
}

__device__ __forceinline__ float apply_act(float x, int act, float slope) {
  switch (act) {
    case ACT_QUICKGELU: return quick_gelu_fast(x);
    case ACT_GELU_ERF: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    case ACT_LEAKY: return x > 0.f ? x : slope * x;
    case ACT_SIGMOID: return __fdividef(1.f, 1.f + __expf(-x));
    case ACT_RELU: return fmaxf(x, 0.f);
    default: return x;
  }
}

__device__ __forceinline__ long long out_row(const GemmEpilogue& ep, int m) {
  if (ep.out_row_group > 0)
    return (long long)(m / ep.out_row_group) * ep.out_group_stride + (m % ep.out_row_group) + ep.out_row_off;
  return m;
}
__device__ __forceinline__ long long res_row(const GemmEpilogue& ep, int m) {
  if (ep.res_row_mod > 0) return (m % ep.res_row_mod) + ep.res_row_off;
  return out_row(ep, m);
}

// bf16 residual of a convolution tile (ResNet bottleneck identity): the epilogue loads it with plain 16-byte loads at the
// start of every 64-column group, which exposes one HBM latency per group (4 per 256-wide tile: ~6 of the ~8 us a
// K = 64 / 128 tile takes).  Called BEFORE the wait for the accumulator: every lane asks L2 for the 128-byte lines of its
// row that this warp will read, so the loads later hit L2 while the tile's MMAs are still running.
template <int BN>
__device__ __forceinline__ void prefetch_residual_bf16(const GemmEpilogue& ep, int m_base, int n0, int M, int N, int lane,
                                                       int half) {
  const int m = m_base + lane;
  if (ep.residual_bf16 == nullptr || m >= M) return;
  const __nv_bfloat16* row = reinterpret_cast<const __nv_bfloat16*>(ep.residual_bf16) + res_row(ep, m) * ep.ldr;
#pragma unroll
  for (int g = half * 2; g < BN / 32; g += 4) {
    const int col0 = n0 + g * 32;
    if (col0 < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + col0));
  }
}

// taddr: TMEM address of (lane quarter base, first accumulator column of the tile); stg_u: this warp's 4 KB staging
// tile (32 rows x 128 B, 16-byte units XOR-swizzled by row & 7); half: which of the two warps of the lane quarter.
// BORDER: compile the zero-border handling of the convolution path in (the plain-GEMM instantiations stay free of it)
template <int BN, bool BORDER = false>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmEpilogue& ep, uint32_t stg_u, uint32_t taddr, int m_base,
                                                   int n0, int M, int N, int lane, int half) {
  const int rr = lane >> 3, uu = lane & 7;  // read-back mapping: row i*4 + rr, 16-byte unit uu
  const int halves = ep.out_bf16 ? 2 : 1;   // 32-column accumulator chunks per 128-byte output group
#pragma unroll 1
  for (int g = half * halves; g < BN / 32; g += 2 * halves) {
    const int col0 = n0 + g * 32;  // first output column of this 128-byte store group
    if (col0 >= N) break;          // warp-uniform
    // (a) residual tile: coalesced 16-byte loads, issued before the TMEM round trip so their latency overlaps it
    float4 res[8];
    if (ep.residual_bf16 != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int m = m_base + i * 4 + rr;
        if (m < M && col0 + uu * 8 < N)
          res[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const __nv_bfloat16*>(ep.residual_bf16) +
                                                    res_row(ep, m) * ep.ldr + col0 + uu * 8);
      }
    } else if (ep.residual != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int m = m_base + i * 4 + rr;
        if (m < M) res[i] = *reinterpret_cast<const float4*>(ep.residual + res_row(ep, m) * ep.ldr + col0 + uu * 4);
      }
    }
    // (b) accumulator row (one per thread) -> bias / activation -> swizzled staging tile [32 rows][128 B]
#pragma unroll 1
    for (int h = 0; h < halves; ++h) {
      const int c0 = col0 + h * 32;
      if (c0 >= N) break;
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr + (uint32_t)((g + h) * 32), r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (ep.bias != nullptr) {
        const float4* bp = reinterpret_cast<const float4*>(ep.bias + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(bp + j);
          v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
      }
      if (ep.act == ACT_QUICKGELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = quick_gelu_fast(v[j]);
      } else if (ep.act == ACT_RELU) {
        // the activation is selected ONCE per group, outside the element loop: a per-element switch costs ~10
        // instructions and a branch per value, which made the epilogue (one warp per scheduler, no latency hiding)
        // the bottleneck of the small-N / small-K convolution GEMMs (7.7 us per 128 x 64 tile)
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (ep.act == ACT_LEAKY) {
        const float slope = ep.slope;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : slope * v[j];
      } else if (ep.act == ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __fdividef(1.f, 1.f + __expf(-v[j]));
      } else if (ep.act == ACT_GELU_ERF) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.f + erff(v[j] * 0.70710678118654752f));
      }
      if (BORDER && ep.border_w2 > 0) {  // zero-bordered image rows stay zero (thread == pixel row)
        const int pr = (m_base + lane) % ep.border_h2w2, py = pr / ep.border_w2, px = pr - py * ep.border_w2;
        if (py == 0 || px == 0 || px == ep.border_w2 - 1 || py == ep.border_h2w2 / ep.border_w2 - 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
      }
      const uint32_t srow = stg_u + (uint32_t)(lane * 128);
      if (ep.out_bf16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = pack2_bf16(v[8 * j + 0], v[8 * j + 1]);
          u.y = pack2_bf16(v[8 * j + 2], v[8 * j + 3]);
          u.z = pack2_bf16(v[8 * j + 4], v[8 * j + 5]);
          u.w = pack2_bf16(v[8 * j + 6], v[8 * j + 7]);
          st_shared_v4(srow + (uint32_t)((((h * 4 + j) ^ (lane & 7))) * 16), u);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 u;
          u.x = __float_as_uint(v[4 * j]); u.y = __float_as_uint(v[4 * j + 1]);
          u.z = __float_as_uint(v[4 * j + 2]); u.w = __float_as_uint(v[4 * j + 3]);
          st_shared_v4(srow + (uint32_t)(((j ^ (lane & 7))) * 16), u);
        }
      }
    }
    __syncwarp();
    // (c) read back row-contiguous: 8 lanes cover one row's 128 bytes -> full-line coalesced global stores
    const bool col_ok = ep.out_bf16 ? (col0 + uu * 8 < N) : (col0 + uu * 4 < N);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = i * 4 + rr;
      uint4 d = ld_shared_v4(stg_u + (uint32_t)(rl * 128 + ((uu ^ (rl & 7)) * 16)));
      if (m_base + rl < M && col_ok) {
        const long long obase = out_row(ep, m_base + rl) * ep.ldo;
        if (ep.out_bf16) {
          if (ep.residual_bf16 != nullptr) {
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(&res[i]);
            uint32_t* dw = reinterpret_cast<uint32_t*>(&d);
            // packed bf16x2 add / max: rn_bf16(a + b) of two bf16 values is what the fp32 add + re-rounding gave, at
            // 2 instructions per pair instead of ~10 (this read-back is on the critical path of the K = 64 convolutions)
            const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
            for (int w2 = 0; w2 < 4; ++w2) {
              __nv_bfloat162 sum = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&dw[w2]),
                                           *reinterpret_cast<const __nv_bfloat162*>(&rw[w2]));
              if (ep.relu_after_residual) sum = __hmax2(sum, zero2);
              dw[w2] = *reinterpret_cast<uint32_t*>(&sum);
            }
          }
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + obase + col0 + uu * 8) = d;
        } else {
          if (ep.residual != nullptr) {
            d.x = __float_as_uint(__uint_as_float(d.x) + res[i].x);
            d.y = __float_as_uint(__uint_as_float(d.y) + res[i].y);
            d.z = __float_as_uint(__uint_as_float(d.z) + res[i].z);
            d.w = __float_as_uint(__uint_as_float(d.w) + res[i].w);
          }
          *reinterpret_cast<uint4*>(reinterpret_cast<float*>(ep.out) + obase + col0 + uu * 4) = d;
        }
      }
    }
    __syncwarp();  // staging tile is reused by the next group
  }
}

// Specialisation for the two hot consumer GEMMs of the ViT encoder (QKV: bias; fc: bias + QuickGELU): bf16 output,
// identity rows, no residual, N a multiple of BN.  Same arithmetic and the same staging scheme as gemm_epilogue_tile,
// but every option is a compile-time constant: no residual registers, no activation dispatch, no row-map arithmetic
// (the general epilogue sits at the 168-register cap; an A/B run showed the encoder GEMMs lose 4 % to ~10 extra
// registers' worth of code in it).
// LNF: the A operand was the raw (not normalised) bf16 residual stream and the weights carry gamma; the row's LayerNorm
// is applied here, out = act(rstd * acc - rstd * mean * colsum[n] + bias[n]) (see GemmEpilogue::ln_stats_in).
template <int BN, int ACT, int LNF = 0>
__device__ __forceinline__ void gemm_epilogue_tile_bf16_bias(const GemmEpilogue& ep, uint32_t stg_u, uint32_t taddr,
                                                             int m_base, int n0, int M, int lane, int half) {
  const int rr = lane >> 3, uu = lane & 7;  // read-back mapping: row i*4 + rr, 16-byte unit uu
  __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(ep.out);
  const uint32_t srow = stg_u + (uint32_t)(lane * 128);
  float rstd = 1.f, nmr = 0.f;   // this thread's accumulator row: 1/std and -mean/std
  if (LNF) {
    const int m = min(m_base + lane, M - 1);
    const float4* sp = reinterpret_cast<const float4*>(ep.ln_stats_in + (long long)m * (2 * LN_FOLD_SLOTS));
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < LN_FOLD_SLOTS / 2; ++j) {   // fixed order: deterministic
      const float4 a = __ldg(sp + j);
      s1 += a.x; s2 += a.y; s1 += a.z; s2 += a.w;
    }
    const float mean = s1 * (1.f / LN_FOLD_C);
    rstd = rsqrtf(fmaxf(s2 * (1.f / LN_FOLD_C) - mean * mean, 0.f) + 1e-5f);
    nmr = -rstd * mean;
  }
#pragma unroll 1
  for (int g = half * 2; g < BN / 32; g += 4) {
    const int col0 = n0 + g * 32;  // first output column of this 128-byte store group (64 bf16)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(taddr + (uint32_t)((g + h) * 32), r);
      tmem_ld_wait();
      const float4* bp = reinterpret_cast<const float4*>(ep.bias + col0 + h * 32);
      float v[32];
      if (LNF == 2) {   // weight rows centred (zero column sums): the mean term vanishes
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(bp + j);
          v[4 * j + 0] = fmaf(rstd, __uint_as_float(r[4 * j + 0]), b.x);
          v[4 * j + 1] = fmaf(rstd, __uint_as_float(r[4 * j + 1]), b.y);
          v[4 * j + 2] = fmaf(rstd, __uint_as_float(r[4 * j + 2]), b.z);
          v[4 * j + 3] = fmaf(rstd, __uint_as_float(r[4 * j + 3]), b.w);
        }
      } else if (LNF) {
        const float4* cp = reinterpret_cast<const float4*>(ep.ln_colsum + col0 + h * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(bp + j), c = __ldg(cp + j);
          v[4 * j + 0] = fmaf(rstd, __uint_as_float(r[4 * j + 0]), fmaf(nmr, c.x, b.x));
          v[4 * j + 1] = fmaf(rstd, __uint_as_float(r[4 * j + 1]), fmaf(nmr, c.y, b.y));
          v[4 * j + 2] = fmaf(rstd, __uint_as_float(r[4 * j + 2]), fmaf(nmr, c.z, b.z));
          v[4 * j + 3] = fmaf(rstd, __uint_as_float(r[4 * j + 3]), fmaf(nmr, c.w, b.w));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(bp + j);
          v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b.x;
          v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
          v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
          v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
        }
      }
      if (ACT == ACT_QUICKGELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = quick_gelu_fast(v[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack2_bf16(v[8 * j + 0], v[8 * j + 1]);
        u.y = pack2_bf16(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack2_bf16(v[8 * j + 4], v[8 * j + 5]);
        u.w = pack2_bf16(v[8 * j + 6], v[8 * j + 7]);
        st_shared_v4(srow + (uint32_t)((((h * 4 + j) ^ (lane & 7))) * 16), u);
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = i * 4 + rr;
      const uint4 d = ld_shared_v4(stg_u + (uint32_t)(rl * 128 + ((uu ^ (rl & 7)) * 16)));
      if (m_base + rl < M)
        *reinterpret_cast<uint4*>(outp + (long long)(m_base + rl) * ep.ldo + col0 + uu * 8) = d;
    }
    __syncwarp();  // staging tile is reused by the next group
  }
}

// Specialisation for the convolutions of the RN50 tower (rn50.cu): bf16 output with identity rows, bias, zero-bordered
// images, and either an optional ReLU (RESID = false) or the bottleneck tail out = relu(acc + bias + identity) with a
// bf16 identity tensor (RESID = true).  The short-K convolutions (K = 64 .. 576) are bound by this epilogue, not by their
// MMAs: ncu showed ~7 k cycles per 32 x 64 group in the general epilogue (~1000 warp instructions: per-row row-map
// arithmetic, a border test with two integer divisions per 32-column chunk, activation dispatch), one warp per scheduler
// with nothing to overlap.  Here the border test is done once per tile, rows are addressed incrementally, the bias is
// requested before the TMEM wait, ReLU runs on packed bf16x2, and every option is a compile-time constant.
template <int BN, bool RESID>
__device__ __forceinline__ void gemm_epilogue_tile_conv(const GemmEpilogue& ep, uint32_t stg_u, uint32_t taddr, int m_base,
                                                        int n0, int M, int N, int lane, int half) {
  const int rr = lane >> 3, uu = lane & 7;  // read-back mapping: row i*4 + rr, 16-byte unit uu
  __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(ep.out);
  const __nv_bfloat16* resp = reinterpret_cast<const __nv_bfloat16*>(ep.residual_bf16);
  // zero-bordered image rows stay zero (thread == pixel row m_base + lane): one test per tile
  const int pr = (m_base + lane) % ep.border_h2w2, py = pr / ep.border_w2, px = pr - py * ep.border_w2;
  const bool border = py == 0 || px == 0 || px == ep.border_w2 - 1 || py == ep.border_h2w2 / ep.border_w2 - 1;
  const bool relu_now = !RESID && ep.act == ACT_RELU;
  const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
  const uint32_t srow = stg_u + (uint32_t)(lane * 128);
  const int m_rd = m_base + rr;             // first row of this lane in the read-back mapping (then +4 per step)
#pragma unroll 1
  for (int g = half * 2; g < BN / 32; g += 4) {
    const int col0 = n0 + g * 32;           // first output column of this 128-byte store group (64 bf16)
    if (col0 >= N) break;                   // warp-uniform
    const bool col_ok = col0 + uu * 8 < N;
    uint4 res[8];
    if (RESID) {
      const __nv_bfloat16* rp = resp + (long long)m_rd * ep.ldr + col0 + uu * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        res[i] = make_uint4(0u, 0u, 0u, 0u);
        if (m_rd + i * 4 < M && col_ok) res[i] = *reinterpret_cast<const uint4*>(rp + (long long)(i * 4) * ep.ldr);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c0 = col0 + h * 32;
      if (c0 < N) {                         // warp-uniform
        float4 b[8];
        const float4* bp = reinterpret_cast<const float4*>(ep.bias + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = __ldg(bp + j);   // in flight while the accumulator chunk arrives
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)((g + h) * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t u[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 bb = b[2 * j + (e >> 1)];
            const float lo = __uint_as_float(r[8 * j + 2 * e]) + ((e & 1) ? bb.z : bb.x);
            const float hi = __uint_as_float(r[8 * j + 2 * e + 1]) + ((e & 1) ? bb.w : bb.y);
            __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
            if (relu_now) p = __hmax2(p, zero2);
            u[e] = border ? 0u : *reinterpret_cast<uint32_t*>(&p);
          }
          st_shared_v4(srow + (uint32_t)((((h * 4 + j) ^ (lane & 7))) * 16), make_uint4(u[0], u[1], u[2], u[3]));
        }
      }
    }
    __syncwarp();
    __nv_bfloat16* op = outp + (long long)m_rd * ep.ldo + col0 + uu * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = i * 4 + rr;
      uint4 d = ld_shared_v4(stg_u + (uint32_t)(rl * 128 + ((uu ^ (rl & 7)) * 16)));
      if (RESID) {
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(&res[i]);
        uint32_t* dw = reinterpret_cast<uint32_t*>(&d);
#pragma unroll
        for (int w2 = 0; w2 < 4; ++w2) {
          __nv_bfloat162 sum = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&dw[w2]),
                                       *reinterpret_cast<const __nv_bfloat162*>(&rw[w2]));
          sum = __hmax2(sum, zero2);
          dw[w2] = *reinterpret_cast<uint32_t*>(&sum);
        }
      }
      if (m_rd + i * 4 < M && col_ok) *reinterpret_cast<uint4*>(op + (long long)(i * 4) * ep.ldo) = d;
    }
    __syncwarp();  // staging tile is reused by the next group
  }
}
// the convolution epilogue applies when the plan has exactly the options it hard-codes
__device__ __forceinline__ bool conv_epilogue_applies(const GemmEpilogue& ep) {
  return ep.out_bf16 && ep.bias != nullptr && ep.residual == nullptr && ep.border_w2 > 0 && ep.out_row_group == 0 &&
         ep.res_row_mod == 0 && (ep.act == ACT_NONE || ep.act == ACT_RELU) &&
         (ep.residual_bf16 == nullptr || (ep.relu_after_residual && ep.act == ACT_NONE));
}

// Variant for fp32 output + fp32 residual (identity rows), used by the 2-CTA kernel: the residual box of this warp's
// 32 rows x 32 columns has been TMA-loaded (SWIZZLE_128B: 16-byte unit u of row r sits at unit u ^ (r & 7), the same
// permutation the staging tile uses) into `buf_u`; the thread owning accumulator row `lane` adds its row in place,
// then the tile is read back row-contiguously and stored.  No residual registers, and the load was issued one column
// group earlier, so its HBM latency is hidden behind the previous group's work.
// LNF: additionally accumulates this thread's row sums (s1 += sum x, s2 += sum x^2 over the 32 final values) and writes a
// bf16 copy of the rows (GemmEpilogue::out2_bf16) -- the operand and the statistics of the LayerNorm folded into the next GEMM.
template <bool LNF = false>
__device__ __forceinline__ void gemm_epilogue_group_restma(const GemmEpilogue& ep, uint32_t buf_u, uint32_t taddr_cols,
                                                           int m_base, int col0, int M, int lane, float& s1, float& s2) {
  uint32_t r[32];
  tmem_ld_32x32b_x32(taddr_cols, r);
  tmem_ld_wait();
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (ep.bias != nullptr) {
    const float4* bp = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(bp + j);
      v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
    }
  }
  const uint32_t srow = buf_u + (uint32_t)(lane * 128);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t a = srow + (uint32_t)((j ^ (lane & 7)) * 16);
    uint4 u = ld_shared_v4(a);
    u.x = __float_as_uint(__uint_as_float(u.x) + v[4 * j]);
    u.y = __float_as_uint(__uint_as_float(u.y) + v[4 * j + 1]);
    u.z = __float_as_uint(__uint_as_float(u.z) + v[4 * j + 2]);
    u.w = __float_as_uint(__uint_as_float(u.w) + v[4 * j + 3]);
    st_shared_v4(a, u);
    if (LNF) {
      const float x0 = __uint_as_float(u.x), x1 = __uint_as_float(u.y), x2 = __uint_as_float(u.z), x3 = __uint_as_float(u.w);
      s1 += (x0 + x1) + (x2 + x3);
      s2 = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, s2))));
    }
  }
  __syncwarp();
  const int rr = lane >> 3, uu = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = i * 4 + rr;
    const uint4 d = ld_shared_v4(buf_u + (uint32_t)(rl * 128 + ((uu ^ (rl & 7)) * 16)));
    if (m_base + rl < M) {
      *reinterpret_cast<uint4*>(reinterpret_cast<float*>(ep.out) + (long long)(m_base + rl) * ep.ldo + col0 + uu * 4) = d;
      if (LNF) {
        uint2 p;
        p.x = pack2_bf16(__uint_as_float(d.x), __uint_as_float(d.y));
        p.y = pack2_bf16(__uint_as_float(d.z), __uint_as_float(d.w));
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(ep.out2_bf16) + (long long)(m_base + rl) * ep.ldo + col0 +
                                  uu * 4) = p;
      }
    }
  }
  __syncwarp();  // every lane has read the buffer: it may be refilled
}

}  // namespace spm
