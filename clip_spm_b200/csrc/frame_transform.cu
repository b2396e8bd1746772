// Evaluation-time frame transform on the GPU: decoded RGB uint8 frames [F, H, W, 3] ->
// Resize(256) -> CenterCrop(224) -> ToTensor, bit for bit what the reference's data loader produces
// (video_reader.py:83-111,265-272; videotransforms/functional.py:24-73 -- its Resize() ends in PIL BILINEAR;
// video_transforms.py:204-247 CenterCrop; torchvision ToTensor = uint8 / 255 in fp32).
// The resampling is Pillow's ImagingResample for 8-bit images (libImaging/Resample.c): separable triangle filter
// widened by the down-scaling factor, double-precision coefficients rounded to 22-bit fixed point, horizontal pass
// into an 8-bit intermediate, vertical pass, accumulators started at 2^21 and `>> 22` clipped to [0, 255].
//
// One fused kernel: only the 224 x 224 crop is ever computed.  A CTA owns a 32 x 32 output tile: it resamples the
// input rows its vertical taps need horizontally into shared memory (packed RGBx words), then resamples those
// vertically and writes either fp32 CHW (the ToTensor layout, [F,3,224,224]) or directly the bf16 patch matrix the
// ViT patch-embedding GEMM consumes ([F*196, 768], column c*256 + ky*16 + kx) -- the fp32 image never exists then.
// HBM-bound byte work: input bytes are read once (plus tile-border rows), the output is written once in 16-byte pieces.
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/clipspm_b200.h"
#include "api_common.cuh"
#include "kernels.cuh"
#include "profile.cuh"

namespace spm {
namespace {
constexpr int CROP = 224, RESIZE = 256, PBITS = 22, TILE = 32;

struct Taps {          // per cropped output coordinate: first input coordinate, tap count, ksize fixed-point weights
  std::vector<int> mn, cnt, k;
  int ksize = 1;
};

// Resample.c precompute_coeffs + normalize_coeffs_8bpc (bilinear filter, support 1) for outputs [first, first+count)
void bilinear_taps(int in_size, int out_size, int first, Taps* t, int count = CROP) {
  const int CROP = count;   // (the evaluation transform needs the crop window only; the training transform the whole extent)
  t->mn.assign(CROP, 0);
  t->cnt.assign(CROP, 1);
  if (in_size == out_size) {  // the pass is skipped by Pillow: identity taps give the same bytes
    t->ksize = 1;
    t->k.assign(CROP, 1 << PBITS);
    for (int i = 0; i < CROP; ++i) t->mn[i] = first + i;
    return;
  }
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const int ksize = (int)std::ceil(support) * 2 + 1;
  const double ss = 1.0 / filterscale;
  t->ksize = ksize;
  t->k.assign((size_t)CROP * ksize, 0);
  std::vector<double> w(ksize);
  for (int i = 0; i < CROP; ++i) {
    const int xx = first + i;
    const double center = (xx + 0.5) * scale;
    int lo = (int)(center - support + 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5);
    if (hi > in_size) hi = in_size;
    const int n = hi - lo;
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
      double a = (x + lo - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      const double v = a < 1.0 ? 1.0 - a : 0.0;
      w[x] = v;
      ww += v;
    }
    for (int x = 0; x < n; ++x) {
      if (ww != 0.0) w[x] /= ww;
      const double f = w[x] * (double)(1 << PBITS);
      t->k[(size_t)i * ksize + x] = w[x] < 0 ? (int)(-0.5 + f) : (int)(0.5 + f);
    }
    t->mn[i] = lo;
    t->cnt[i] = n;
  }
}

struct Geometry {
  int oh, ow, y1, x1;
};

// functional.py:45-52,66-73 (resize target) and video_transforms.py:244-245 (crop origin; Python round = half to even)
int geometry(int H, int W, Geometry* g) {
  if ((W <= H && W == RESIZE) || (H <= W && H == RESIZE)) {
    g->oh = H; g->ow = W;
  } else if (W < H) {
    g->ow = RESIZE; g->oh = (int)((double)(RESIZE * (long long)H) / (double)W);
  } else {
    g->oh = RESIZE; g->ow = (int)((double)(RESIZE * (long long)W) / (double)H);
  }
  if (g->oh < CROP || g->ow < CROP) {
    set_error("frame transform: resized frame is smaller than the 224 x 224 crop");
    return 1;
  }
  g->y1 = (int)std::nearbyint((g->oh - CROP) / 2.0);
  g->x1 = (int)std::nearbyint((g->ow - CROP) / 2.0);
  return 0;
}

struct DevTables {
  int *hmin, *hcnt, *hk, *vmin, *vcnt, *vk;
  int hks, vks, max_rows;
  int ow, oh;   // full-extent tables (training transform): the resized size the tables cover; 0 for crop-window tables
};

std::mutex g_mu;
std::map<std::pair<int, int>, DevTables> g_tables;
std::map<std::pair<int, int>, DevTables> g_tables_full;

int upload(const std::vector<int>& v, int** dst) {
  SPM_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), v.size() * sizeof(int)));
  SPM_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

int get_tables(int H, int W, DevTables* out) {
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_tables.find({H, W});
  if (it != g_tables.end()) { *out = it->second; return 0; }
  Geometry g;
  SPM_TRY(geometry(H, W, &g));
  Taps th, tv;
  bilinear_taps(W, g.ow, g.x1, &th);
  bilinear_taps(H, g.oh, g.y1, &tv);
  DevTables d{};
  d.hks = th.ksize; d.vks = tv.ksize;
  d.max_rows = 0;
  for (int y0 = 0; y0 < CROP; y0 += TILE) {
    const int y1 = std::min(CROP, y0 + TILE) - 1;
    d.max_rows = std::max(d.max_rows, tv.mn[y1] + tv.cnt[y1] - tv.mn[y0]);
  }
  if ((size_t)d.max_rows * TILE * 4 > 200 * 1024) {
    set_error("frame transform: down-scaling factor too large for the shared-memory tile");
    return 1;
  }
  if (g_tables.size() >= 256) {  // bounded cache: drop everything once no kernel can still be reading a table
    SPM_CUDA(cudaDeviceSynchronize());
    for (auto& kv : g_tables) {
      cudaFree(kv.second.hmin); cudaFree(kv.second.hcnt); cudaFree(kv.second.hk);
      cudaFree(kv.second.vmin); cudaFree(kv.second.vcnt); cudaFree(kv.second.vk);
    }
    g_tables.clear();
  }
  SPM_TRY(upload(th.mn, &d.hmin)); SPM_TRY(upload(th.cnt, &d.hcnt)); SPM_TRY(upload(th.k, &d.hk));
  SPM_TRY(upload(tv.mn, &d.vmin)); SPM_TRY(upload(tv.cnt, &d.vcnt)); SPM_TRY(upload(tv.k, &d.vk));
  g_tables[{H, W}] = d;
  *out = d;
  return 0;
}

// Tables over the WHOLE resized frame: the training transform crops at a per-clip random origin (and may mirror the clip), so
// the taps of any 224-wide window must be at hand
int get_tables_full(int H, int W, DevTables* out) {
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_tables_full.find({H, W});
  if (it != g_tables_full.end()) { *out = it->second; return 0; }
  Geometry g;
  SPM_TRY(geometry(H, W, &g));
  Taps th, tv;
  bilinear_taps(W, g.ow, 0, &th, g.ow);
  bilinear_taps(H, g.oh, 0, &tv, g.oh);
  DevTables d{};
  d.hks = th.ksize; d.vks = tv.ksize; d.ow = g.ow; d.oh = g.oh;
  d.max_rows = 0;
  for (int y0 = 0; y0 < g.oh; ++y0) {      // a 32-row output tile may start at any row
    const int y1 = std::min(g.oh, y0 + TILE) - 1;
    d.max_rows = std::max(d.max_rows, tv.mn[y1] + tv.cnt[y1] - tv.mn[y0]);
  }
  if ((size_t)d.max_rows * TILE * 4 > 200 * 1024) {
    set_error("frame transform: down-scaling factor too large for the shared-memory tile");
    return 1;
  }
  if (g_tables_full.size() >= 256) {
    SPM_CUDA(cudaDeviceSynchronize());
    for (auto& kv : g_tables_full) {
      cudaFree(kv.second.hmin); cudaFree(kv.second.hcnt); cudaFree(kv.second.hk);
      cudaFree(kv.second.vmin); cudaFree(kv.second.vcnt); cudaFree(kv.second.vk);
    }
    g_tables_full.clear();
  }
  SPM_TRY(upload(th.mn, &d.hmin)); SPM_TRY(upload(th.cnt, &d.hcnt)); SPM_TRY(upload(th.k, &d.hk));
  SPM_TRY(upload(tv.mn, &d.vmin)); SPM_TRY(upload(tv.cnt, &d.vcnt)); SPM_TRY(upload(tv.k, &d.vk));
  g_tables_full[{H, W}] = d;
  *out = d;
  return 0;
}

__device__ __forceinline__ int clip8(int acc) { return min(max(acc >> PBITS, 0), 255); }

// MODE 0: fp32 [F,3,224,224]; MODE 1: bf16 ViT patch rows [F*196, 768]
// AUG (training transform, video_reader.py:97-103: Resize -> RandomHorizontalFlip -> RandomCrop): the tables cover the whole
// resized frame and aug[f] = {crop y1, crop x1, flip} picks the window (and mirrors it) per frame
template <int MODE, bool AUG = false>
__global__ void __launch_bounds__(256)
frame_transform_kernel(const uint8_t* __restrict__ frames, int H, int W, DevTables t, float* __restrict__ out_f32,
                       __nv_bfloat16* __restrict__ out_patch, const int* __restrict__ aug = nullptr) {
  extern __shared__ uint32_t hbuf[];  // [rows][TILE] packed R | G<<8 | B<<16 after the horizontal pass
  const int x0 = blockIdx.x * TILE, y0 = blockIdx.y * TILE;
  const long long f = blockIdx.z;
  const uint8_t* src = frames + f * (long long)H * W * 3;
  int oy = 0, ox = 0, flip = 0;     // table index of output (y, x): row oy + y, column ox + x (mirrored: ow - 1 - (ox + x))
  if (AUG) {
    oy = min(max(__ldg(aug + 3 * f), 0), t.oh - CROP);
    ox = min(max(__ldg(aug + 3 * f + 1), 0), t.ow - CROP);
    flip = __ldg(aug + 3 * f + 2) != 0;
  }
  const int rlo = __ldg(t.vmin + oy + y0);
  const int ylast = min(CROP, y0 + TILE) - 1;
  const int nrows = __ldg(t.vmin + oy + ylast) + __ldg(t.vcnt + oy + ylast) - rlo;
  // (A) horizontal pass of the needed input rows for the tile's 32 columns (a warp = one row: contiguous bytes)
  {
    const int col = threadIdx.x & 31;
    const int tx = AUG ? (flip ? t.ow - 1 - (ox + x0 + col) : ox + x0 + col) : x0 + col;
    const int mn = __ldg(t.hmin + tx), cnt = __ldg(t.hcnt + tx);
    const int* kp = t.hk + tx * t.hks;
    for (int r = threadIdx.x >> 5; r < nrows; r += 8) {
      const uint8_t* p = src + ((long long)(rlo + r) * W + mn) * 3;
      int a0 = 1 << (PBITS - 1), a1 = a0, a2 = a0;
      for (int j = 0; j < cnt; ++j) {
        const int k = __ldg(kp + j);
        a0 += (int)p[3 * j] * k; a1 += (int)p[3 * j + 1] * k; a2 += (int)p[3 * j + 2] * k;
      }
      hbuf[r * TILE + col] = (uint32_t)clip8(a0) | ((uint32_t)clip8(a1) << 8) | ((uint32_t)clip8(a2) << 16);
    }
  }
  __syncthreads();
  // (B) vertical pass; a thread produces PIX consecutive pixels of one channel -> one 16-byte store
  constexpr int PIX = MODE == 0 ? 4 : 8, GROUPS = TILE / PIX;
  for (int item = threadIdx.x; item < TILE * GROUPS * 3; item += 256) {
    const int xg = item % GROUPS, c = (item / GROUPS) % 3, ry = item / (GROUPS * 3);
    const int y = y0 + ry;
    if (y >= CROP) break;
    const int r0 = __ldg(t.vmin + oy + y) - rlo, cnt = __ldg(t.vcnt + oy + y);
    const int* kp = t.vk + (oy + y) * t.vks;
    int acc[PIX];
#pragma unroll
    for (int i = 0; i < PIX; ++i) acc[i] = 1 << (PBITS - 1);
    for (int j = 0; j < cnt; ++j) {
      const int k = __ldg(kp + j);
      const uint32_t* row = hbuf + (r0 + j) * TILE + xg * PIX;
#pragma unroll
      for (int i = 0; i < PIX; ++i) acc[i] += (int)((row[i] >> (8 * c)) & 0xffu) * k;
    }
    float v[PIX];
#pragma unroll
    for (int i = 0; i < PIX; ++i) v[i] = __fdiv_rn((float)clip8(acc[i]), 255.f);  // ToTensor: exact fp32 division
    const int x = x0 + xg * PIX;
    if (MODE == 0) {
      *reinterpret_cast<float4*>(out_f32 + ((f * 3 + c) * CROP + y) * CROP + x) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      const long long prow = f * 196 + (y >> 4) * 14 + (x >> 4);
      __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0], v[1]), b1 = __floats2bfloat162_rn(v[2], v[3]);
      __nv_bfloat162 b2 = __floats2bfloat162_rn(v[PIX - 4], v[PIX - 3]), b3 = __floats2bfloat162_rn(v[PIX - 2], v[PIX - 1]);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&b0); u.y = *reinterpret_cast<uint32_t*>(&b1);
      u.z = *reinterpret_cast<uint32_t*>(&b2); u.w = *reinterpret_cast<uint32_t*>(&b3);
      *reinterpret_cast<uint4*>(out_patch + prow * 768 + c * 256 + (y & 15) * 16 + (x & 15)) = u;
    }
  }
}
}  // namespace

// frames [F,H,W,3] uint8 (device) -> out_f32 [F,3,224,224] (if non-null) or out_patch [F*196,768] bf16
int k_frame_transform(cudaStream_t st, const uint8_t* frames, int n_frames, int H, int W, float* out_f32,
                      __nv_bfloat16* out_patch) {
  if (n_frames <= 0) return 0;
  if (H < 1 || W < 1 || (long long)H * W > (1LL << 26)) { set_error("frame transform: bad frame size"); return 1; }
  DevTables t;
  SPM_TRY(get_tables(H, W, &t));
  static bool attr_set = false;
  if (!attr_set) {
    SPM_CUDA(cudaFuncSetAttribute(frame_transform_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SPM_CUDA(cudaFuncSetAttribute(frame_transform_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const size_t smem = (size_t)t.max_rows * TILE * 4;
  for (int f0 = 0; f0 < n_frames; f0 += 65535) {
    const int nf = std::min(65535, n_frames - f0);
    const dim3 grid(CROP / TILE, CROP / TILE, nf);
    const uint8_t* src = frames + (long long)f0 * H * W * 3;
    if (out_f32 != nullptr)
      frame_transform_kernel<0><<<grid, 256, smem, st>>>(src, H, W, t, out_f32 + (long long)f0 * 3 * CROP * CROP, nullptr);
    else
      frame_transform_kernel<1><<<grid, 256, smem, st>>>(src, H, W, t, nullptr, out_patch + (long long)f0 * 196 * 768);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("frame transform launch: ") + cudaGetErrorString(e)); return 1; }
    count_launch();
  }
  return 0;
}
// training transform: frames [F,H,W,3] uint8, aug [F,3] int32 {y1, x1, flip} (device) -> out_f32 [F,3,224,224]
int k_frame_transform_train(cudaStream_t st, const uint8_t* frames, int n_frames, int H, int W, const int* aug, float* out_f32) {
  if (n_frames <= 0) return 0;
  if (H < 1 || W < 1 || (long long)H * W > (1LL << 26)) { set_error("frame transform: bad frame size"); return 1; }
  DevTables t;
  SPM_TRY(get_tables_full(H, W, &t));
  static bool attr_set = false;
  if (!attr_set) {
    SPM_CUDA(cudaFuncSetAttribute(frame_transform_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const size_t smem = (size_t)t.max_rows * TILE * 4;
  for (int f0 = 0; f0 < n_frames; f0 += 65535) {
    const int nf = std::min(65535, n_frames - f0);
    const dim3 grid(CROP / TILE, CROP / TILE, nf);
    frame_transform_kernel<0, true><<<grid, 256, smem, st>>>(frames + (long long)f0 * H * W * 3, H, W, t,
                                                             out_f32 + (long long)f0 * 3 * CROP * CROP, nullptr, aug + 3LL * f0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("frame transform launch: ") + cudaGetErrorString(e)); return 1; }
    count_launch();
  }
  return 0;
}
}  // namespace spm

extern "C" {

int spm_transform_frames_train(void* stream, const uint8_t* frames, int n_frames, int H, int W, const int32_t* aug,
                               float* images_out) {
  if (n_frames <= 0) return 0;
  SPM_CHECK(frames && aug && images_out, "spm_transform_frames_train: null argument");
  return spm::k_frame_transform_train((cudaStream_t)stream, frames, n_frames, H, W, aug, images_out);
}

int spm_frame_geometry(int H, int W, int* resized_h, int* resized_w, int* crop_y, int* crop_x) {
  spm::Geometry g;
  SPM_TRY(spm::geometry(H, W, &g));
  if (resized_h) *resized_h = g.oh;
  if (resized_w) *resized_w = g.ow;
  if (crop_y) *crop_y = g.y1;
  if (crop_x) *crop_x = g.x1;
  return 0;
}

int spm_transform_frames(void* stream, const uint8_t* frames, int n_frames, int H, int W, float* images_out) {
  if (n_frames <= 0) return 0;
  SPM_CHECK(frames && images_out, "spm_transform_frames: null argument");
  return spm::k_frame_transform((cudaStream_t)stream, frames, n_frames, H, W, images_out, nullptr);
}

}  // extern "C"
