// TEST INFRASTRUCTURE (oracle/): the JPEG arithmetic of clip_spm_b200/csrc/jpeg_core.cuh + jpeg_parse.h compiled for the
// CPU and run sequentially, so that the restatement of libjpeg's integer algorithms can be pinned against PIL -- the
// decoder the reference calls (video_reader.py:227-230) -- on a box without a GPU.  Built by oracle/Makefile into
// oracle/_ref/libjpeg_check.so; only tests/ load it.  The product path (csrc/jpeg_decode.cu) never does.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../clip_spm_b200/csrc/jpeg_parse.h"

using namespace spm::jpeg;

extern "C" int jpeg_check_info(const unsigned char* data, long long n, int* H, int* W, int* hs, int* vs, char* err, int err_cap) {
  Parsed p;
  if (parse_jpeg(data, (size_t)n, &p)) { std::snprintf(err, (size_t)err_cap, "%s", p.err.c_str()); return 1; }
  *H = p.d.height; *W = p.d.width; *hs = p.d.hs[0]; *vs = p.d.vs[0];
  return 0;
}

// out: [H, W, 3] uint8 RGB
extern "C" int jpeg_check_decode(const unsigned char* data, long long n, unsigned char* out, char* err, int err_cap) {
  Parsed p;
  if (parse_jpeg(data, (size_t)n, &p)) { std::snprintf(err, (size_t)err_cap, "%s", p.err.c_str()); return 1; }
  ImageDesc& d = p.d;
  std::vector<std::vector<int16_t>> coef(3);
  std::vector<std::vector<uint8_t>> plane(3);
  for (int c = 0; c < 3; ++c) {
    coef[c].assign((size_t)d.blocks_x[c] * d.blocks_y[c] * 64, 0);
    plane[c].assign((size_t)d.blocks_x[c] * d.blocks_y[c] * 64, 0);
  }
  const int total = d.mcus_x * d.mcus_y;
  for (int k = 0; k < d.n_intervals; ++k) {
    BitReader br;
    const uint8_t* base = p.scan.data();
    br.init(base + p.interval_start[(size_t)k],
            base + (k + 1 < d.n_intervals ? (size_t)p.interval_start[(size_t)k + 1] : p.scan.size()));
    const int m0 = d.restart_interval > 0 ? k * d.restart_interval : 0;
    const int m1 = d.restart_interval > 0 ? (m0 + d.restart_interval < total ? m0 + d.restart_interval : total) : total;
    int pred[3] = {0, 0, 0};
    for (int m = m0; m < m1; ++m) {
      const int my = m / d.mcus_x, mx = m - my * d.mcus_x;
      for (int c = 0; c < 3; ++c)
        for (int v = 0; v < d.vs[c]; ++v)
          for (int h = 0; h < d.hs[c]; ++h) {
            const int bx = mx * d.hs[c] + h, by = my * d.vs[c] + v;
            decode_block(br, d.dc[d.td[c]], d.ac[d.ta[c]], pred[c], coef[c].data() + ((size_t)by * d.blocks_x[c] + bx) * 64);
          }
    }
  }
  for (int c = 0; c < 3; ++c) {
    const int stride = d.blocks_x[c] * 8;
    for (int by = 0; by < d.blocks_y[c]; ++by)
      for (int bx = 0; bx < d.blocks_x[c]; ++bx)
        idct_block(coef[c].data() + ((size_t)by * d.blocks_x[c] + bx) * 64, d.quant[d.tq[c]],
                   plane[c].data() + (size_t)(by * 8) * stride + bx * 8, stride);
  }
  const int sy = d.blocks_x[0] * 8, sc = d.blocks_x[1] * 8;
  for (int y = 0; y < d.height; ++y)
    for (int x = 0; x < d.width; ++x) {
      const int Y = plane[0][(size_t)y * sy + x];
      int cb, cr;
      if (d.hs[0] == 2 && d.vs[0] == 2) {
        const int cols = (d.width + 1) / 2, rows = (d.height + 1) / 2;
        cb = upsample_h2v2(plane[1].data(), sc, cols, rows, x, y);
        cr = upsample_h2v2(plane[2].data(), sc, cols, rows, x, y);
      } else if (d.hs[0] == 2) {
        const int cols = (d.width + 1) / 2;
        cb = upsample_h2v1(plane[1].data() + (size_t)y * sc, cols, x);
        cr = upsample_h2v1(plane[2].data() + (size_t)y * sc, cols, x);
      } else {
        cb = plane[1][(size_t)y * sc + x];
        cr = plane[2][(size_t)y * sc + x];
      }
      ycc_to_rgb(Y, cb, cr, out + ((size_t)y * d.width + x) * 3);
    }
  return 0;
}
