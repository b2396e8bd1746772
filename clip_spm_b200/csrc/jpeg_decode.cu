// JPEG decode on the GPU (the head of the episode input pipeline): the frames the reference reads
// with PIL (video_reader.py:227-230 `Image.open(path).load()`, one file per frame) are decoded here from the file bytes
// to the uint8 RGB frames [n, H, W, 3] that frame_transform.cu (Resize / CenterCrop / ToTensor) consumes -- bit-identical
// to PIL / libjpeg-turbo (jpeg_core.cuh restates its integer arithmetic).
//
//   host    parse_jpeg (jpeg_parse.h) per image on a small thread pool: headers, tables, byte un-stuffing, restart cuts
//   K1      entropy decode: ONE THREAD PER (image, restart interval) walks its bit stream sequentially (Huffman coding is
//           serial by nature); a sweep holds thousands of frames, so the grid still has thousands of threads.  Writes the
//           quantised coefficients (int16, natural order) of every block
//   K2      dequantisation + 8x8 integer IDCT, one thread per block, coalesced 8-byte row stores into the sample planes
//   K3      fancy chroma upsampling + YCbCr -> RGB, one thread per output pixel
// HBM-bound byte work after K1: coefficients are written once and read once (2 x 2 bytes per sample), planes once each.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "../../include/clipspm_b200.h"
#include "api_common.cuh"
#include "jpeg_parse.h"
#include "profile.cuh"

namespace spm {
namespace {

using namespace jpeg;

__global__ void jpeg_entropy_kernel(const ImageDesc* __restrict__ descs, const uint8_t* __restrict__ stream,
                                    const int32_t* __restrict__ iv_start, const int32_t* __restrict__ iv_image,
                                    int n_intervals_total, int16_t* __restrict__ coef) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_intervals_total) return;
  const int img = iv_image[t];
  if (img < 0) return;   // the slot that only carries a scan's length
  const ImageDesc& d = descs[img];
  const int k = t - (int)d.interval_off;
  const uint8_t* base = stream + d.data_off;
  const int32_t* ivs = iv_start + d.interval_off;
  const long long scan_len = ivs[d.n_intervals];   // one extra entry: the length of the image's scan
  BitReader br;
  br.init(base + ivs[k], base + (k + 1 < d.n_intervals ? (long long)ivs[k + 1] : scan_len));
  const int total = d.mcus_x * d.mcus_y;
  const int m0 = d.restart_interval > 0 ? k * d.restart_interval : 0;
  const int m1 = d.restart_interval > 0 ? min(total, m0 + d.restart_interval) : total;
  int pred[MAX_COMP] = {0, 0, 0};
  for (int m = m0; m < m1; ++m) {
    const int my = m / d.mcus_x, mx = m - my * d.mcus_x;
    for (int c = 0; c < d.n_comp; ++c)
      for (int v = 0; v < d.vs[c]; ++v)
        for (int h = 0; h < d.hs[c]; ++h) {
          const int bx = mx * d.hs[c] + h, by = my * d.vs[c] + v;
          decode_block(br, d.dc[d.td[c]], d.ac[d.ta[c]], pred[c],
                       coef + d.coef_off[c] + ((long long)by * d.blocks_x[c] + bx) * 64);
        }
  }
}

// grid (blocks per image, images): all images of a call share the geometry (checked on the host)
__global__ void jpeg_idct_kernel(const ImageDesc* __restrict__ descs, const int16_t* __restrict__ coef,
                                 uint8_t* __restrict__ planes, int blocks_per_image) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= blocks_per_image) return;
  const ImageDesc& d = descs[blockIdx.y];
  int c = 0, rem = b;
  while (c + 1 < d.n_comp && rem >= d.blocks_x[c] * d.blocks_y[c]) { rem -= d.blocks_x[c] * d.blocks_y[c]; ++c; }
  const int by = rem / d.blocks_x[c], bx = rem - by * d.blocks_x[c];
  int16_t cf[64];
  const uint4* src = reinterpret_cast<const uint4*>(coef + d.coef_off[c] + (long long)rem * 64);
#pragma unroll
  for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(cf)[i] = src[i];
  const int stride = d.blocks_x[c] * 8;
  uint8_t px[64];
  idct_block(cf, d.quant[d.tq[c]], px, 8);
  uint8_t* dst = planes + d.plane_off[c] + (long long)(by * 8) * stride + bx * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) *reinterpret_cast<uint2*>(dst + (long long)r * stride) = reinterpret_cast<const uint2*>(px)[r];
}

// grid (ceil(W*H / 256), images): out [n, H, W, 3]
__global__ void jpeg_color_kernel(const ImageDesc* __restrict__ descs, const uint8_t* __restrict__ planes,
                                  uint8_t* __restrict__ out) {
  const ImageDesc& d = descs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.width * d.height) return;
  const int y = i / d.width, x = i - y * d.width;
  const int sy = d.blocks_x[0] * 8, sc = d.blocks_x[1] * 8;
  const uint8_t* py = planes + d.plane_off[0];
  const uint8_t* pb = planes + d.plane_off[1];
  const uint8_t* pr = planes + d.plane_off[2];
  const int32_t Y = py[(long long)y * sy + x];
  int32_t cb, cr;
  if (d.hs[0] == 2 && d.vs[0] == 2) {
    const int cols = (d.width + 1) / 2, rows = (d.height + 1) / 2;
    cb = upsample_h2v2(pb, sc, cols, rows, x, y);
    cr = upsample_h2v2(pr, sc, cols, rows, x, y);
  } else if (d.hs[0] == 2) {
    const int cols = (d.width + 1) / 2;
    cb = upsample_h2v1(pb + (long long)y * sc, cols, x);
    cr = upsample_h2v1(pr + (long long)y * sc, cols, x);
  } else {
    cb = pb[(long long)y * sc + x];
    cr = pr[(long long)y * sc + x];
  }
  ycc_to_rgb(Y, cb, cr, out + ((long long)blockIdx.y * d.height * d.width + i) * 3);
}

#define JPEG_LAUNCH_CHECK()                                                                        \
  do {                                                                                             \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess) { set_error(std::string("jpeg kernel launch: ") + cudaGetErrorString(_e)); return 1; } \
    count_launch();                                                                                \
  } while (0)

template <class F>
void parallel_for(int n, F f) {
  const int nt = std::max(1, std::min<int>({n, 16, (int)std::thread::hardware_concurrency()}));
  if (nt == 1) { for (int i = 0; i < n; ++i) f(i); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&] { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) f(i); });
  for (auto& x : th) x.join();
}

}  // namespace
}  // namespace spm

using namespace spm;

extern "C" {

int spm_jpeg_info(const uint8_t* jpeg_host, long long n_bytes, int* height, int* width, int* h_samp, int* v_samp) {
  SPM_CHECK(jpeg_host != nullptr && n_bytes > 0, "spm_jpeg_info: null argument");
  jpeg::Parsed p;
  if (jpeg::parse_jpeg(jpeg_host, (size_t)n_bytes, &p)) { set_error("spm_jpeg_info: " + p.err); return 1; }
  if (height) *height = p.d.height;
  if (width) *width = p.d.width;
  if (h_samp) *h_samp = p.d.hs[0];
  if (v_samp) *v_samp = p.d.vs[0];
  return 0;
}

int spm_jpeg_decode(void* stream, int n_images, const uint8_t* const* jpeg_host, const int64_t* jpeg_bytes, int H, int W,
                    uint8_t* frames_out) {
  SPM_CHECK(jpeg_host != nullptr && jpeg_bytes != nullptr && frames_out != nullptr, "spm_jpeg_decode: null argument");
  if (n_images <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<jpeg::Parsed> parsed((size_t)n_images);
  std::atomic<int> bad(-1);
  parallel_for(n_images, [&](int i) {
    if (jpeg::parse_jpeg(jpeg_host[i], (size_t)jpeg_bytes[i], &parsed[i])) { int e = -1; bad.compare_exchange_strong(e, i); }
  });
  if (bad.load() >= 0) {
    set_error("spm_jpeg_decode: image " + std::to_string(bad.load()) + ": " + parsed[bad.load()].err);
    return 1;
  }
  // ---- pack: one geometry for the whole call (the frames of a dataset share it), per-image tables and streams
  const jpeg::ImageDesc& g = parsed[0].d;
  SPM_CHECK(g.height == H && g.width == W, "spm_jpeg_decode: image size differs from the stated H x W");
  long long coef_per_image = 0, plane_per_image = 0;
  int blocks_per_image = 0;
  long long coef_off[jpeg::MAX_COMP], plane_off[jpeg::MAX_COMP];
  for (int c = 0; c < g.n_comp; ++c) {
    coef_off[c] = coef_per_image; plane_off[c] = plane_per_image;
    const int nb = g.blocks_x[c] * g.blocks_y[c];
    blocks_per_image += nb;
    coef_per_image += (long long)nb * 64;
    plane_per_image += (long long)nb * 64;
  }
  std::vector<jpeg::ImageDesc> descs((size_t)n_images);
  std::vector<int32_t> iv_start, iv_image;
  long long stream_bytes = 0;
  for (int i = 0; i < n_images; ++i) {
    jpeg::ImageDesc d = parsed[i].d;
    if (d.height != H || d.width != W || d.hs[0] != g.hs[0] || d.vs[0] != g.vs[0]) {
      set_error("spm_jpeg_decode: image " + std::to_string(i) + " differs in size or chroma subsampling from image 0");
      return 1;
    }
    d.data_off = stream_bytes;
    d.interval_off = (long long)iv_start.size();
    for (int c = 0; c < d.n_comp; ++c) {
      d.coef_off[c] = (long long)i * coef_per_image + coef_off[c];
      d.plane_off[c] = (long long)i * plane_per_image + plane_off[c];
    }
    for (int k = 0; k < d.n_intervals; ++k) { iv_start.push_back(parsed[i].interval_start[(size_t)k]); iv_image.push_back(i); }
    iv_start.push_back((int32_t)parsed[i].scan.size());   // the extra entry: scan length
    iv_image.push_back(-1);
    stream_bytes += ((long long)parsed[i].scan.size() + 15) / 16 * 16;
    descs[(size_t)i] = d;
  }
  std::vector<uint8_t> packed((size_t)stream_bytes, 0);
  parallel_for(n_images, [&](int i) {
    std::memcpy(packed.data() + descs[(size_t)i].data_off, parsed[(size_t)i].scan.data(), parsed[(size_t)i].scan.size());
  });
  // ---- device buffers (stream-ordered pool: no device-wide sync)
  jpeg::ImageDesc* d_desc = nullptr;
  uint8_t *d_stream = nullptr, *d_planes = nullptr;
  int32_t *d_ivs = nullptr, *d_ivi = nullptr;
  int16_t* d_coef = nullptr;
  const size_t n_slots = iv_start.size();
  SPM_CUDA(cudaMallocAsync((void**)&d_desc, descs.size() * sizeof(jpeg::ImageDesc), st));
  SPM_CUDA(cudaMallocAsync((void**)&d_stream, (size_t)std::max<long long>(stream_bytes, 16), st));
  SPM_CUDA(cudaMallocAsync((void**)&d_ivs, n_slots * 4, st));
  SPM_CUDA(cudaMallocAsync((void**)&d_ivi, n_slots * 4, st));
  SPM_CUDA(cudaMallocAsync((void**)&d_coef, (size_t)n_images * coef_per_image * 2, st));
  SPM_CUDA(cudaMallocAsync((void**)&d_planes, (size_t)n_images * plane_per_image, st));
  SPM_CUDA(cudaMemcpyAsync(d_desc, descs.data(), descs.size() * sizeof(jpeg::ImageDesc), cudaMemcpyHostToDevice, st));
  SPM_CUDA(cudaMemcpyAsync(d_stream, packed.data(), (size_t)stream_bytes, cudaMemcpyHostToDevice, st));
  SPM_CUDA(cudaMemcpyAsync(d_ivs, iv_start.data(), n_slots * 4, cudaMemcpyHostToDevice, st));
  SPM_CUDA(cudaMemcpyAsync(d_ivi, iv_image.data(), n_slots * 4, cudaMemcpyHostToDevice, st));
  SPM_CUDA(cudaMemsetAsync(d_coef, 0, (size_t)n_images * coef_per_image * 2, st));
  jpeg_entropy_kernel<<<(unsigned)((n_slots + 63) / 64), 64, 0, st>>>(d_desc, d_stream, d_ivs, d_ivi, (int)n_slots, d_coef);
  JPEG_LAUNCH_CHECK();
  jpeg_idct_kernel<<<dim3((unsigned)((blocks_per_image + 127) / 128), (unsigned)n_images), 128, 0, st>>>(d_desc, d_coef, d_planes,
                                                                                                      blocks_per_image);
  JPEG_LAUNCH_CHECK();
  jpeg_color_kernel<<<dim3((unsigned)(((long long)H * W + 255) / 256), (unsigned)n_images), 256, 0, st>>>(d_desc, d_planes,
                                                                                                       frames_out);
  JPEG_LAUNCH_CHECK();
  cudaFreeAsync(d_desc, st); cudaFreeAsync(d_stream, st); cudaFreeAsync(d_ivs, st); cudaFreeAsync(d_ivi, st);
  cudaFreeAsync(d_coef, st); cudaFreeAsync(d_planes, st);
  // the host vectors above are pageable: their copies must have been consumed before they go out of scope
  SPM_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
