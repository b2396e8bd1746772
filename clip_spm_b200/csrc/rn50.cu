// CLIP ModifiedResNet-50 frame encoder (models/clip_fsar.py:502-608 Bottleneck / ModifiedResNet, :396-500
// AttentionPool2d) for sm_100a.
//
// Activations are NHWC bf16 images with a 1-pixel ZERO BORDER ([F, H+2, W+2, C]), so every convolution is a GEMM on
// the tcgen05 kernel of gemm_tcgen05.cu over the matrix [F*(H+2)*(W+2), C] exactly as it lies in memory:
//   1x1 conv            A = that matrix, B = folded weight [Cout, Cin]
//   3x3 conv (pad 1)    IMPLICIT GEMM: tap (dy,dx) of the window is the same matrix shifted by the constant row offset
//                       dy*(W+2)+dx, so k-block kb just loads TMA box (channel block, m0 + offset(tap)); no im2col is
//                       ever materialised and the 9 re-reads hit L2.  B = folded weight [Cout, 9*Cpad]
//   stem conv1 (s=2)    A = im2col straight from the fp32 NCHW image, K = 27 padded to 32
// Border rows are computed like any other row and written as zeros by the epilogue (GemmEpilogue::border_*), which
// keeps every tensor a valid zero-padded input for the next 3x3 convolution.
// BatchNorm (eval, running statistics) is folded into the weights and a per-channel bias at load time; ReLU, the
// bottleneck's residual add and the final ReLU run in the GEMM epilogue.  AvgPool2d(2) (stem, anti-aliased strides,
// downsample branches) is a vectorised NHWC kernel.  The attention pool uses only the mean-token query
// (clip_fsar.py:481-499): k/v projections as one fused GEMM over 50 tokens, q projection on the F mean tokens, a
// one-warp-per-(frame, head) softmax, then c_proj.
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <cuda_bf16.h>

#include "api_common.cuh"
#include "gemm.cuh"
#include "profile.cuh"
#include "rn50.cuh"

namespace spm {

namespace {
// Frames per launch (r02), set separately for the FRONT (stem, layer1, layer2: 3364 .. 900 pixel rows per frame) and the
// BACK (layer3, layer4, attention pool: 256 / 81 rows per frame; it runs over one or several front sub-chunks at once).
// Measured on the config-4 shape (tools/rn50_throughput.py, 8 episodes = 1280 frames per call): 64 / 64 frames 131
// episodes/s -> 216 / 216 233 -> 320 / 320 240; small FRONT chunks that would keep layer1's tensors inside the 126 MB L2
// lose more to the fixed cost of ~30 extra launches than they gain (27: 190, 72: 219, 108: 226 episodes/s), and the
// BACK needs hundreds of frames for its GEMMs to fill 74 CTA pairs for more than one round (at 64 frames layer4's
// 512-channel convolutions were 164 tiles of 128 x 128 = 1.1 rounds).
// SPM_RN50_FRONT_CHUNK / SPM_RN50_BACK_CHUNK override (back is rounded to a multiple of front).
constexpr int RN_FRONT_DEFAULT = 320, RN_BACK_DEFAULT = 320;
constexpr int N_FRONT_BLOCKS = 7;   // layer1 (3) + layer2 (4)
constexpr int EMB = 2048, HEADS = 32, HD = 64, OUT_DIM = 1024, NTOK = 50;
constexpr long long FRAME_ELEMS = 3LL * 224 * 224;
constexpr long long SCRATCH_PER_FRAME = 58LL * 58 * 256;   // largest padded activation: 58x58x256 (> 114x114x64)
constexpr long long MID_PER_FRAME = 30LL * 30 * 512;       // layer2 output = layer3 input (and the back's largest tensor)
constexpr long long COL_PER_FRAME = 114LL * 114 * 32;      // stem conv1 im2col over the padded 114x114 grid

#define RN_LAUNCH_CHECK()                                                         \
  do {                                                                            \
    cudaError_t _e = cudaGetLastError();                                          \
    if (_e != cudaSuccess) { set_error(std::string("rn50 kernel launch: ") + cudaGetErrorString(_e)); return 1; } \
    count_launch();                                                               \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// load-time folding:  w [Cout, Cin, kh, kw] fp32 + BN -> wout [Cout, Kpad] bf16 (k = (ky*kw + kx)*Cin + c), bias
// ---------------------------------------------------------------------------------------------------------
// column k = tap * cpad + c (cpad >= Cin: channels of a tap padded with zero columns), zero beyond kh*kw*cpad
__global__ void fold_conv_kernel(const float* __restrict__ w, const float* __restrict__ g, const float* __restrict__ b,
                                 const float* __restrict__ mean, const float* __restrict__ var, int Cout, int Cin,
                                 int kh, int kw, int cpad, int Kpad, __nv_bfloat16* __restrict__ wout,
                                 float* __restrict__ bias, int pair) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * Kpad) return;
  const int o = (int)(i / Kpad), k = (int)(i % Kpad);
  const float scale = g[o] * rsqrtf(var[o] + 1e-5f);
  float v = 0.f;
  int c = k % cpad, tap = k / cpad;
  if (pair) {
    // pixel-pair layout of gemm_plan_conv3x3 (Cin == 32): k-block kb = k / 64 holds (dy = kb / 2; kb even: taps x-1 | x,
    // kb odd: tap x+1 | zeros), 32 channels per tap
    const int kb = k >> 6, w64 = k & 63, hf = w64 >> 5, kx = (kb & 1) ? (hf == 0 ? 2 : -1) : hf;
    c = w64 & 31;
    tap = kx < 0 ? kh * kw : (kb >> 1) * kw + kx;
  }
  if (tap < kh * kw && c < Cin) {
    const int ky = tap / kw, kx = tap % kw;
    v = w[(((long long)o * Cin + c) * kh + ky) * kw + kx] * scale;
  }
  wout[i] = __float2bfloat16_rn(v);
  if (k == 0) bias[o] = b[o] - mean[o] * scale;
}
__global__ void cast_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}

// ---------------------------------------------------------------------------------------------------------
// activation-side kernels (NHWC bf16, 8 channels = 16 bytes per thread)
// ---------------------------------------------------------------------------------------------------------
// stem conv1: 3x3, stride 2, pad 1 on the fp32 NCHW image -> rows of the zero-bordered 114x114 output grid,
// [F*114*114, 32] (27 taps, 5 zero columns; border rows all zero)
__global__ void stem_im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, long long rows) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int ox = (int)(r % 114) - 1, oy = (int)((r / 114) % 114) - 1;
  const long long f = r / (114 * 114);
  const bool border = ox < 0 || oy < 0 || ox >= 112 || oy >= 112;
  const float* base = img + f * FRAME_ELEMS;
  __align__(16) __nv_bfloat16 v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = __float2bfloat16_rn(0.f);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
      if (!border && iy >= 0 && iy < 224 && ix >= 0 && ix < 224) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          v[(ky * 3 + kx) * 3 + c] = __float2bfloat16_rn(__ldg(base + ((long long)c * 224 + iy) * 224 + ix));
      }
    }
  }
  uint4* o = reinterpret_cast<uint4*>(out + r * 32);
  const uint4* s = reinterpret_cast<const uint4*>(v);
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = s[j];
}

__device__ __forceinline__ float2 bf2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
__device__ __forceinline__ uint32_t pk(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}
// AvgPool2d(2) on zero-bordered images: [F,H+2,W+2,C] -> [F,H/2+2,W/2+2,C] (border of the output written as zeros)
__global__ void avgpool2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int H, int W,
                                int C, long long n_units) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_units) return;
  const int c8 = C / 8, Ho = H / 2, Wo = W / 2, Wp = W + 2, Hp = H + 2;
  const int cu = (int)(i % c8);
  const long long orow = i / c8;
  const int ox = (int)(orow % (Wo + 2)) - 1, oy = (int)((orow / (Wo + 2)) % (Ho + 2)) - 1;
  const long long f = orow / ((long long)(Wo + 2) * (Ho + 2));
  if (ox < 0 || oy < 0 || ox >= Wo || oy >= Ho) {
    reinterpret_cast<uint4*>(out)[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((f * Hp + 2 * oy + dy + 1) * Wp + 2 * ox + dx + 1) * C) + cu);
      const float2 a = bf2(v.x), b = bf2(v.y), c = bf2(v.z), d = bf2(v.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
  uint4 o;
  o.x = pk(acc[0] * 0.25f, acc[1] * 0.25f); o.y = pk(acc[2] * 0.25f, acc[3] * 0.25f);
  o.z = pk(acc[4] * 0.25f, acc[5] * 0.25f); o.w = pk(acc[6] * 0.25f, acc[7] * 0.25f);
  reinterpret_cast<uint4*>(out)[i] = o;
}

// attention-pool tokens (clip_fsar.py:407-410): tok[f,0] = mean_s x[f,s] + pos[0]; tok[f,1+s] = x[f,s] + pos[1+s]
// grid (F, EMB / 1024): a thread owns 8 channels (16 bytes) of one frame; the 49 spatial rows stream through it once
__global__ void __launch_bounds__(128)
attnpool_tokens_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ pos,
                       __nv_bfloat16* __restrict__ tok) {
  const int f = blockIdx.x, c = (blockIdx.y * 128 + threadIdx.x) * 8;
  float mean[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int s = 0; s < 49; ++s) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((long long)f * 81 + (s / 7 + 1) * 9 + (s % 7 + 1)) * EMB + c));
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + (1 + s) * EMB + c));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + (1 + s) * EMB + c + 4));
    const float2 a = bf2(v.x), b = bf2(v.y), cc = bf2(v.z), d = bf2(v.w);
    mean[0] += a.x; mean[1] += a.y; mean[2] += b.x; mean[3] += b.y;
    mean[4] += cc.x; mean[5] += cc.y; mean[6] += d.x; mean[7] += d.y;
    uint4 o;
    o.x = pk(a.x + p0.x, a.y + p0.y); o.y = pk(b.x + p0.z, b.y + p0.w);
    o.z = pk(cc.x + p1.x, cc.y + p1.y); o.w = pk(d.x + p1.z, d.y + p1.w);
    *reinterpret_cast<uint4*>(tok + ((long long)f * NTOK + 1 + s) * EMB + c) = o;
  }
  const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + c)), p1 = __ldg(reinterpret_cast<const float4*>(pos + c + 4));
  const float k = 1.f / 49.f;
  uint4 o;
  o.x = pk(mean[0] * k + p0.x, mean[1] * k + p0.y); o.y = pk(mean[2] * k + p0.z, mean[3] * k + p0.w);
  o.z = pk(mean[4] * k + p1.x, mean[5] * k + p1.y); o.w = pk(mean[6] * k + p1.z, mean[7] * k + p1.w);
  *reinterpret_cast<uint4*>(tok + (long long)f * NTOK * EMB + c) = o;
}

// one warp per (frame, head): q [F, 2048], kv [F*50, 4096] (k | v) -> o [F, 2048]
__global__ void attnpool_attend_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                                       __nv_bfloat16* __restrict__ o, int n_pairs) {
  const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pair >= n_pairs) return;
  const int f = pair / HEADS, hh = pair % HEADS;
  const __nv_bfloat16* qp = q + (long long)f * EMB + hh * HD;
  float s[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = lane + 32 * r;
    if (key < NTOK) {
      const __nv_bfloat16* kp = kv + ((long long)f * NTOK + key) * (2 * EMB) + hh * HD;
      float acc = 0.f;
      for (int d = 0; d < HD; ++d) acc += __bfloat162float(qp[d]) * __bfloat162float(kp[d]);
      s[r] = acc * 0.125f;
    }
  }
  float m = fmaxf(s[0], s[1]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  const float p0 = expf(s[0] - m), p1 = (lane + 32 < NTOK) ? expf(s[1] - m) : 0.f;
  float l = p0 + p1;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
  float a0 = 0.f, a1 = 0.f;
  for (int key = 0; key < NTOK; ++key) {
    const float p = __shfl_sync(0xffffffffu, key < 32 ? p0 : p1, key & 31);
    const __nv_bfloat16* vp = kv + ((long long)f * NTOK + key) * (2 * EMB) + EMB + hh * HD;
    a0 += p * __bfloat162float(vp[lane]);
    a1 += p * __bfloat162float(vp[lane + 32]);
  }
  const float inv = 1.f / l;
  o[(long long)f * EMB + hh * HD + lane] = __float2bfloat16_rn(a0 * inv);
  o[(long long)f * EMB + hh * HD + lane + 32] = __float2bfloat16_rn(a1 * inv);
}

struct Conv {
  __nv_bfloat16* w = nullptr;
  float* bias = nullptr;
  int cout = 0, cin = 0, ksz = 1, kpad = 0;
  bool implicit3x3 = false;  // weights laid out [Cout, 9 * cpad] for the shifted-TMA implicit GEMM
};
struct Block {
  Conv c1, c2, c3, down;
  int inpl, planes, stride;
  bool has_down;
};
struct Plan {
  std::vector<GemmOp> ops;
};
}  // namespace

struct Rn50 {
  int sms = 148;
  int front = RN_FRONT_DEFAULT, back = RN_BACK_DEFAULT;
  __nv_bfloat16* mid = nullptr;   // [back, 30, 30, 512] layer2 outputs of the current back chunk
  std::vector<void*> allocs;
  Conv stem[3];
  std::vector<Block> blocks;
  float* pos = nullptr;
  __nv_bfloat16 *q_w = nullptr, *kv_w = nullptr, *c_w = nullptr;
  float *q_b = nullptr, *kv_b = nullptr, *c_b = nullptr;
  // workspace (RN_CHUNK frames)
  __nv_bfloat16 *xa = nullptr, *xb = nullptr, *t1 = nullptr, *t2 = nullptr, *t2p = nullptr, *xd = nullptr,
                *idn = nullptr, *col = nullptr, *tok = nullptr, *qbuf = nullptr, *kvbuf = nullptr, *obuf = nullptr;
  std::map<int, std::unique_ptr<Plan>> plans;        // front, by frame count
  std::map<int, std::unique_ptr<Plan>> back_plans;   // back, by frame count
};

namespace {
template <class T>
int ralloc(Rn50* r, T** p, long long n) {
  SPM_CUDA(cudaMalloc(reinterpret_cast<void**>(p), (size_t)std::max<long long>(n, 4) * sizeof(T)));
  r->allocs.push_back(*p);
  return 0;
}

int load_conv(Rn50* r, cudaStream_t st, const WeightGetter& get, const std::string& wname, const std::string& bn,
              int cout, int cin, int ksz, Conv* c, bool implicit3x3 = false) {
  const float *w, *g, *b, *mean, *var;
  SPM_TRY(get(wname, (long long)cout * cin * ksz * ksz, &w));
  SPM_TRY(get(bn + "weight", cout, &g));
  SPM_TRY(get(bn + "bias", cout, &b));
  SPM_TRY(get(bn + "running_mean", cout, &mean));
  SPM_TRY(get(bn + "running_var", cout, &var));
  c->cout = cout; c->cin = cin; c->ksz = ksz; c->implicit3x3 = implicit3x3;
  const int cpad = implicit3x3 ? (cin + 63) / 64 * 64 : cin;
  const int pair = (implicit3x3 && cin == 32 && gemm_conv_pair_supported()) ? 1 : 0;
  c->kpad = pair ? 6 * 64 : implicit3x3 ? 9 * cpad : (cin * ksz * ksz + 31) / 32 * 32;
  SPM_TRY(ralloc(r, &c->w, (long long)cout * c->kpad));
  SPM_TRY(ralloc(r, &c->bias, cout));
  const long long n = (long long)cout * c->kpad;
  fold_conv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, g, b, mean, var, cout, cin, ksz, ksz, cpad, c->kpad,
                                                                 c->w, c->bias, pair);
  RN_LAUNCH_CHECK();
  return 0;
}
int load_linear(Rn50* r, cudaStream_t st, const WeightGetter& get, const std::string& name, int nout, int nin,
                __nv_bfloat16* wdst, float* bdst) {
  const float *w, *b;
  SPM_TRY(get(name + ".weight", (long long)nout * nin, &w));
  SPM_TRY(get(name + ".bias", nout, &b));
  const long long n = (long long)nout * nin;
  cast_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, wdst, n);
  RN_LAUNCH_CHECK();
  SPM_CUDA(cudaMemcpyAsync(bdst, b, (size_t)nout * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// rows = F * H2 * H2 pixels of zero-bordered H2 x H2 images
int plan_conv(Rn50* r, Plan* pl, const Conv& c, const __nv_bfloat16* A, long long rows, int H2, __nv_bfloat16* out,
              bool relu, const __nv_bfloat16* residual) {
  GemmEpilogue ep;
  ep.bias = c.bias; ep.out = out; ep.ldo = c.cout; ep.out_bf16 = 1;
  ep.border_w2 = H2; ep.border_h2w2 = H2 * H2;
  if (residual != nullptr) {
    ep.residual_bf16 = residual; ep.ldr = c.cout; ep.relu_after_residual = 1;
  } else if (relu) {
    ep.act = ACT_RELU;
  }
  GemmOp op;
  const char* err = "";
  const int rc = c.implicit3x3 ? gemm_plan_conv3x3(&op, A, c.cin, (int)rows, H2, c.w, c.cout, ep, r->sms, &err)
                               : gemm_plan(&op, GEMM_BF16, A, c.kpad, c.w, c.kpad, (int)rows, c.cout, c.kpad, ep, r->sms, &err);
  if (rc) {
    set_error(std::string("rn50 gemm_plan: ") + err);
    return 1;
  }
  pl->ops.push_back(op);
  return 0;
}

// GEMM lists in execution order.  The non-GEMM kernels are replayed in the same order by run_front / run_back.
int plan_blocks(Rn50* r, Plan* pl, int F, int b_begin, int b_end, __nv_bfloat16*& x, __nv_bfloat16*& y, int& H) {
  for (int bi = b_begin; bi < b_end; ++bi) {
    const Block& b = r->blocks[bi];
    const int Ho = H / b.stride;
    const long long rows_in = (long long)F * (H + 2) * (H + 2), rows_out = (long long)F * (Ho + 2) * (Ho + 2);
    SPM_TRY(plan_conv(r, pl, b.c1, x, rows_in, H + 2, r->t1, true, nullptr));
    SPM_TRY(plan_conv(r, pl, b.c2, r->t1, rows_in, H + 2, r->t2, true, nullptr));   // implicit 3x3
    const __nv_bfloat16* c3_in = b.stride > 1 ? r->t2p : r->t2;
    const __nv_bfloat16* idn = x;
    if (b.has_down) {
      SPM_TRY(plan_conv(r, pl, b.down, b.stride > 1 ? r->xd : x, rows_out, Ho + 2, r->idn, false, nullptr));
      idn = r->idn;
    }
    SPM_TRY(plan_conv(r, pl, b.c3, c3_in, rows_out, Ho + 2, y, false, idn));
    std::swap(x, y);
    H = Ho;
  }
  return 0;
}

// stem + layer1 + layer2 for F frames; the last convolution's output pointer is patched per call (-> r->mid + offset)
int build_front_plan(Rn50* r, int F, Plan* pl) {
  const long long R114 = (long long)F * 114 * 114;
  SPM_TRY(plan_conv(r, pl, r->stem[0], r->col, R114, 114, r->t1, true, nullptr));   // explicit im2col (from the image)
  SPM_TRY(plan_conv(r, pl, r->stem[1], r->t1, R114, 114, r->t2, true, nullptr));    // implicit 3x3
  SPM_TRY(plan_conv(r, pl, r->stem[2], r->t2, R114, 114, r->t1, true, nullptr));    // implicit 3x3
  __nv_bfloat16 *x = r->xa, *y = r->xb;
  int H = 56;
  return plan_blocks(r, pl, F, 0, N_FRONT_BLOCKS, x, y, H);
}

// layer3 + layer4 + attention pool for F frames, reading r->mid
int build_back_plan(Rn50* r, int F, Plan* pl) {
  __nv_bfloat16 *x = r->mid, *y = r->xb;
  int H = 28;
  SPM_TRY(plan_blocks(r, pl, F, N_FRONT_BLOCKS, (int)r->blocks.size(), x, y, H));
  // attention pool: kv over all 50 tokens, q over the mean tokens (row stride 50*2048), c_proj
  const char* err = "";
  {
    GemmEpilogue ep;
    ep.bias = r->kv_b; ep.out = r->kvbuf; ep.ldo = 2 * EMB; ep.out_bf16 = 1;
    GemmOp op;
    if (gemm_plan(&op, GEMM_BF16, r->tok, EMB, r->kv_w, EMB, F * NTOK, 2 * EMB, EMB, ep, r->sms, &err)) { set_error(err); return 1; }
    pl->ops.push_back(op);
  }
  {
    GemmEpilogue ep;
    ep.bias = r->q_b; ep.out = r->qbuf; ep.ldo = EMB; ep.out_bf16 = 1;
    GemmOp op;
    if (gemm_plan(&op, GEMM_BF16, r->tok, (long long)NTOK * EMB, r->q_w, EMB, F, EMB, EMB, ep, r->sms, &err)) { set_error(err); return 1; }
    pl->ops.push_back(op);
  }
  {
    GemmEpilogue ep;
    ep.bias = r->c_b; ep.out = r->qbuf /* patched per call */; ep.ldo = OUT_DIM;
    GemmOp op;
    if (gemm_plan(&op, GEMM_BF16, r->obuf, EMB, r->c_w, EMB, F, OUT_DIM, EMB, ep, r->sms, &err)) { set_error(err); return 1; }
    pl->ops.push_back(op);
  }
  return 0;
}

int run_gemm(const GemmOp& op, cudaStream_t st) {
  const char* err = "";
  if (gemm_run(&op, st, &err)) { set_error(std::string("rn50 gemm_run: ") + err); return 1; }
  return 0;
}
int avgpool(cudaStream_t st, const __nv_bfloat16* in, __nv_bfloat16* out, int F, int H, int C) {
  const long long n = (long long)F * (H / 2 + 2) * (H / 2 + 2) * (C / 8);  // all pixels of the padded output
  avgpool2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, H, H, C, n);
  RN_LAUNCH_CHECK();
  return 0;
}

// the residual blocks [b_begin, b_end) in plan order
int run_blocks(Rn50* r, cudaStream_t st, const std::vector<GemmOp>& ops, size_t& g, int F, int b_begin, int b_end,
               __nv_bfloat16*& x, __nv_bfloat16*& y, int& H, __nv_bfloat16* last_out) {
  for (int bi = b_begin; bi < b_end; ++bi) {
    const Block& b = r->blocks[bi];
    SPM_TRY(run_gemm(ops[g++], st));                                   // conv1 1x1 -> t1
    SPM_TRY(run_gemm(ops[g++], st));                                   // conv2 3x3 (implicit) -> t2
    if (b.stride > 1) SPM_TRY(avgpool(st, r->t2, r->t2p, F, H, b.planes));
    if (b.has_down) {
      if (b.stride > 1) SPM_TRY(avgpool(st, x, r->xd, F, H, b.inpl));
      SPM_TRY(run_gemm(ops[g++], st));                                 // downsample conv+bn -> idn
    }
    if (bi == b_end - 1 && last_out != nullptr) {                      // conv3 + bn3 + identity + relu -> caller's buffer
      GemmOp c3 = ops[g++];
      c3.ep.out = last_out;
      SPM_TRY(run_gemm(c3, st));
    } else {
      SPM_TRY(run_gemm(ops[g++], st));                                 // conv3 + bn3 + identity + relu -> y
    }
    std::swap(x, y);
    H /= b.stride;
  }
  return 0;
}

// stem + layer1 + layer2 of F frames -> out [F, 30, 30, 512] (zero-bordered)
int run_front(Rn50* r, cudaStream_t st, const float* images, int F, __nv_bfloat16* out) {
  auto it = r->plans.find(F);
  if (it == r->plans.end()) {
    std::unique_ptr<Plan> pl(new Plan());
    SPM_TRY(build_front_plan(r, F, pl.get()));
    it = r->plans.emplace(F, std::move(pl)).first;
  }
  const std::vector<GemmOp>& ops = it->second->ops;
  size_t g = 0;
  // ---- stem (clip_fsar.py:594-599); every tensor below is a zero-bordered image
  const long long R114 = (long long)F * 114 * 114;
  stem_im2col_kernel<<<(unsigned)((R114 + 127) / 128), 128, 0, st>>>(images, r->col, R114);
  RN_LAUNCH_CHECK();
  SPM_TRY(run_gemm(ops[g++], st));                         // conv1+bn1+relu -> t1 [.,32]
  SPM_TRY(run_gemm(ops[g++], st));                         // conv2 (implicit 3x3) -> t2 [.,32]
  SPM_TRY(run_gemm(ops[g++], st));                         // conv3 (implicit 3x3) -> t1 [.,64]
  SPM_TRY(avgpool(st, r->t1, r->xa, F, 112, 64));          // -> x [F,58,58,64]
  // ---- layer1, layer2 (clip_fsar.py:534-547)
  __nv_bfloat16 *x = r->xa, *y = r->xb;
  int H = 56;
  return run_blocks(r, st, ops, g, F, 0, N_FRONT_BLOCKS, x, y, H, out);
}

// layer3 + layer4 + attention pool of the F frames in r->mid -> feats_out [F, 1024]
int run_back(Rn50* r, cudaStream_t st, int F, float* feats_out) {
  auto it = r->back_plans.find(F);
  if (it == r->back_plans.end()) {
    std::unique_ptr<Plan> pl(new Plan());
    SPM_TRY(build_back_plan(r, F, pl.get()));
    it = r->back_plans.emplace(F, std::move(pl)).first;
  }
  const std::vector<GemmOp>& ops = it->second->ops;
  size_t g = 0;
  __nv_bfloat16 *x = r->mid, *y = r->xb;
  int H = 28;
  SPM_TRY(run_blocks(r, st, ops, g, F, N_FRONT_BLOCKS, (int)r->blocks.size(), x, y, H, nullptr));
  // ---- attention pool
  attnpool_tokens_kernel<<<dim3(F, EMB / 1024), 128, 0, st>>>(x, r->pos, r->tok);
  RN_LAUNCH_CHECK();
  SPM_TRY(run_gemm(ops[g++], st));  // k | v
  SPM_TRY(run_gemm(ops[g++], st));  // q (mean token)
  attnpool_attend_kernel<<<(F * HEADS + 3) / 4, 128, 0, st>>>(r->qbuf, r->kvbuf, r->obuf, F * HEADS);
  RN_LAUNCH_CHECK();
  GemmOp fin = ops[g++];
  fin.ep.out = feats_out;
  SPM_TRY(run_gemm(fin, st));
  return 0;
}
}  // namespace

int rn50_create(Rn50** out, cudaStream_t st, int sms, const WeightGetter& get) {
  std::unique_ptr<Rn50> r(new Rn50());
  r->sms = sms;
  const std::string p = "backbone.";
  const float* probe;
  if (get(p + "conv1.weight", 32LL * 3 * 9, &probe) != 0) {
    // head-only use (no backbone tensors supplied): keep an empty encoder; encoding frames reports it
    set_error("");
    *out = r.release();
    return 0;
  }
  SPM_TRY(load_conv(r.get(), st, get, p + "conv1.weight", p + "bn1.", 32, 3, 3, &r->stem[0]));
  SPM_TRY(load_conv(r.get(), st, get, p + "conv2.weight", p + "bn2.", 32, 32, 3, &r->stem[1], true));
  SPM_TRY(load_conv(r.get(), st, get, p + "conv3.weight", p + "bn3.", 64, 32, 3, &r->stem[2], true));
  int inpl = 64;
  const int nblk[4] = {3, 4, 6, 3};
  for (int li = 0; li < 4; ++li) {
    const int planes = 64 << li;
    for (int bi = 0; bi < nblk[li]; ++bi) {
      Block b;
      b.inpl = inpl; b.planes = planes; b.stride = (li > 0 && bi == 0) ? 2 : 1;
      b.has_down = b.stride > 1 || inpl != planes * 4;
      const std::string bp = p + "layer" + std::to_string(li + 1) + "." + std::to_string(bi) + ".";
      SPM_TRY(load_conv(r.get(), st, get, bp + "conv1.weight", bp + "bn1.", planes, inpl, 1, &b.c1));
      SPM_TRY(load_conv(r.get(), st, get, bp + "conv2.weight", bp + "bn2.", planes, planes, 3, &b.c2, true));
      SPM_TRY(load_conv(r.get(), st, get, bp + "conv3.weight", bp + "bn3.", planes * 4, planes, 1, &b.c3));
      if (b.has_down)
        SPM_TRY(load_conv(r.get(), st, get, bp + "downsample.0.weight", bp + "downsample.1.", planes * 4, inpl, 1, &b.down));
      r->blocks.push_back(b);
      inpl = planes * 4;
    }
  }
  const std::string ap = p + "attnpool.";
  const float* pos;
  SPM_TRY(get(ap + "positional_embedding", (long long)NTOK * EMB, &pos));
  SPM_TRY(ralloc(r.get(), &r->pos, (long long)NTOK * EMB));
  SPM_CUDA(cudaMemcpyAsync(r->pos, pos, (size_t)NTOK * EMB * 4, cudaMemcpyDeviceToDevice, st));
  SPM_TRY(ralloc(r.get(), &r->q_w, (long long)EMB * EMB));
  SPM_TRY(ralloc(r.get(), &r->kv_w, 2LL * EMB * EMB));
  SPM_TRY(ralloc(r.get(), &r->c_w, (long long)OUT_DIM * EMB));
  SPM_TRY(ralloc(r.get(), &r->q_b, EMB));
  SPM_TRY(ralloc(r.get(), &r->kv_b, 2 * EMB));
  SPM_TRY(ralloc(r.get(), &r->c_b, OUT_DIM));
  SPM_TRY(load_linear(r.get(), st, get, ap + "q_proj", EMB, EMB, r->q_w, r->q_b));
  SPM_TRY(load_linear(r.get(), st, get, ap + "k_proj", EMB, EMB, r->kv_w, r->kv_b));
  SPM_TRY(load_linear(r.get(), st, get, ap + "v_proj", EMB, EMB, r->kv_w + (long long)EMB * EMB, r->kv_b + EMB));
  SPM_TRY(load_linear(r.get(), st, get, ap + "c_proj", OUT_DIM, EMB, r->c_w, r->c_b));
  // workspace: the buffers are shared by the front (sized per front sub-chunk) and the back (per back chunk)
  if (const char* e = getenv("SPM_RN50_FRONT_CHUNK")) r->front = std::max(1, atoi(e));
  if (const char* e = getenv("SPM_RN50_BACK_CHUNK")) r->back = std::max(1, atoi(e));
  r->back = std::max(r->front, r->back / r->front * r->front);
  const long long S = std::max(SCRATCH_PER_FRAME * r->front, MID_PER_FRAME * r->back);
  SPM_TRY(ralloc(r.get(), &r->mid, MID_PER_FRAME * r->back));
  SPM_TRY(ralloc(r.get(), &r->xa, S));
  SPM_TRY(ralloc(r.get(), &r->xb, S));
  SPM_TRY(ralloc(r.get(), &r->t1, S));
  SPM_TRY(ralloc(r.get(), &r->t2, S));
  SPM_TRY(ralloc(r.get(), &r->t2p, S));
  SPM_TRY(ralloc(r.get(), &r->xd, S));
  SPM_TRY(ralloc(r.get(), &r->idn, S));
  SPM_TRY(ralloc(r.get(), &r->col, COL_PER_FRAME * r->front));
  SPM_TRY(ralloc(r.get(), &r->tok, (long long)r->back * NTOK * EMB));
  SPM_TRY(ralloc(r.get(), &r->qbuf, (long long)r->back * EMB));
  SPM_TRY(ralloc(r.get(), &r->kvbuf, (long long)r->back * NTOK * 2 * EMB));
  SPM_TRY(ralloc(r.get(), &r->obuf, (long long)r->back * EMB));
  *out = r.release();
  return 0;
}

int rn50_encode(Rn50* r, cudaStream_t st, const float* images, int n_frames, float* feats_out) {
  SPM_CHECK(r != nullptr && !r->blocks.empty(), "RN50 encoder: backbone weights were not loaded");
  for (int b0 = 0; b0 < n_frames; b0 += r->back) {
    const int FB = std::min(r->back, n_frames - b0);
    for (int f0 = 0; f0 < FB; f0 += r->front) {
      const int F = std::min(r->front, FB - f0);
      SPM_TRY(run_front(r, st, images + (long long)(b0 + f0) * FRAME_ELEMS, F, r->mid + (long long)f0 * MID_PER_FRAME));
    }
    SPM_TRY(run_back(r, st, FB, feats_out + (long long)b0 * OUT_DIM));
  }
  return 0;
}

void rn50_destroy(Rn50* r) {
  if (r == nullptr) return;
  for (void* p : r->allocs) cudaFree(p);
  delete r;
}

}  // namespace spm
