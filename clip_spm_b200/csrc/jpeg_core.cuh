// Baseline-JPEG decoding arithmetic shared by the device kernels (jpeg_decode.cu) and the host-side check build
// (oracle/jpeg_check.cu compiles the same functions for the CPU to compare them with PIL without a GPU).
//
// What the reference does: video_reader.py:227-230 `Image.open(path).load()` -- PIL hands the file to libjpeg(-turbo)
// with its defaults: Huffman entropy decoding, ISLOW integer IDCT, "fancy" (triangle) chroma upsampling, fixed-point
// YCbCr -> RGB.  Every step below is integer arithmetic restated from the published libjpeg algorithm descriptions
// (ITU T.81 Annex F for the entropy coder; the Loeffler-Ligtenberg-Moschytz 13-bit fixed-point IDCT of jidctint.c; the
// 3/4-1/4 triangle filter of jdsample.c; the 16-bit fixed-point colour tables of jdcolor.c), so the decoded RGB bytes are
// identical to PIL's, bit for bit (tests/test_jpeg_*: compared with PIL on the box).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SPM_HD __host__ __device__ __forceinline__
#else
#define SPM_HD inline
#endif

namespace spm {
namespace jpeg {

constexpr int MAX_COMP = 3;

// zigzag position -> natural (row-major) position inside an 8x8 block (T.81 Figure A.6)
SPM_HD int zigzag_to_natural(int k) {
  constexpr uint8_t Z[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  return Z[k];
}

// One Huffman table in canonical form (T.81 Annex C / F.2.2.3): codes of length l are the integers
// mincode[l] .. maxcode[l], their symbols huffval[valptr[l] + code - mincode[l]]; look[] resolves codes of <= 8 bits in
// one step (high byte = code length, low byte = symbol; 0 = longer code).
struct HuffTable {
  int32_t maxcode[18];   // maxcode[l] = -1 when no code has length l; maxcode[17] = sentinel
  int32_t mincode[17];
  int32_t valptr[17];
  uint8_t huffval[256];
  uint16_t look[256];
};

// Per-image decode parameters, filled by the host parser (jpeg_decode.cu: parse_jpeg)
struct ImageDesc {
  int32_t width, height;
  int32_t n_comp;                      // 3 (YCbCr)
  int32_t hs[MAX_COMP], vs[MAX_COMP];  // sampling factors (luma 1x1, 2x1 or 2x2; chroma 1x1)
  int32_t tq[MAX_COMP];                // quantisation table of each component
  int32_t td[MAX_COMP], ta[MAX_COMP];  // DC / AC Huffman table of each component
  int32_t mcus_x, mcus_y;              // MCU grid
  int32_t restart_interval;            // MCUs per restart interval (0 = none)
  int32_t n_intervals;                 // entropy-coded segments (1 when restart_interval == 0)
  int64_t data_off;                    // byte offset of the unstuffed scan data in the packed stream buffer
  int64_t interval_off;                // index of this image's first entry in the interval-start table
  int64_t coef_off[MAX_COMP];          // int16 offset of each component's coefficient blocks [by][bx][64]
  int32_t blocks_x[MAX_COMP], blocks_y[MAX_COMP];   // padded block grid of each component (whole MCUs)
  int64_t plane_off[MAX_COMP];         // byte offset of each component's sample plane [blocks_y*8][blocks_x*8]
  uint16_t quant[4][64];               // quantisation tables, NATURAL order
  HuffTable dc[2], ac[2];
};

// ------------------------------------------------------------------------------------------------------------
// entropy decoding (T.81 F.2.2): MSB-first bit reader over the unstuffed byte stream
// ------------------------------------------------------------------------------------------------------------
struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t buf;   // the next `cnt` bits are the LOW cnt bits of buf, most significant first
  int cnt;
  SPM_HD void init(const uint8_t* begin, const uint8_t* e) { p = begin; end = e; buf = 0; cnt = 0; }
  SPM_HD void fill() {   // keep at least 32 bits available (zeros past the end, like libjpeg's padding with a warning)
    while (cnt <= 32) {
      const uint64_t b = p < end ? *p : 0;
      ++p;
      buf = (buf << 8) | b;
      cnt += 8;
    }
  }
  SPM_HD uint32_t peek(int n) const { return (uint32_t)((buf >> (cnt - n)) & ((1u << n) - 1u)); }
  SPM_HD void skip(int n) { cnt -= n; }
  SPM_HD uint32_t get(int n) { const uint32_t v = peek(n); cnt -= n; return v; }
};

SPM_HD int huff_decode(BitReader& br, const HuffTable& t) {
  br.fill();
  const uint16_t e = t.look[br.peek(8)];
  if (e != 0) { br.skip(e >> 8); return e & 255; }
  int l = 9;
  int32_t code = (int32_t)br.peek(9);
  while (l <= 16 && code > t.maxcode[l]) { ++l; code = (int32_t)br.peek(l); }
  if (l > 16) { br.skip(16); return 0; }   // corrupt data: libjpeg substitutes a zero symbol
  br.skip(l);
  return t.huffval[t.valptr[l] + code - t.mincode[l]];
}

// F.2.2.1 EXTEND: the s-bit magnitude category value -> signed coefficient
SPM_HD int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

// One 8x8 block: DC difference + run-length coded AC coefficients -> coef[64] (natural order, NOT dequantised).
// `coef` must be zero on entry.
SPM_HD void decode_block(BitReader& br, const HuffTable& dc, const HuffTable& ac, int& pred, int16_t* coef) {
  int s = huff_decode(br, dc);
  if (s) { br.fill(); pred += extend((int)br.get(s), s); }
  coef[0] = (int16_t)pred;
  for (int k = 1; k < 64;) {
    const int rs = huff_decode(br, ac), r = rs >> 4;
    s = rs & 15;
    if (s) {
      k += r;
      br.fill();
      const int v = extend((int)br.get(s), s);
      if (k < 64) coef[zigzag_to_natural(k)] = (int16_t)v;
      ++k;
    } else {
      if (r != 15) break;   // EOB
      k += 16;              // ZRL
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// inverse DCT: the 13-bit fixed-point LLM algorithm ("ISLOW"), two passes with 2 extra bits kept between them
// ------------------------------------------------------------------------------------------------------------
SPM_HD int32_t descale(int32_t x, int n) { return (x + (1 << (n - 1))) >> n; }
// libjpeg's post-IDCT range table: the low 10 bits as a signed value, + 128, clamped to [0, 255]
SPM_HD uint8_t idct_range(int32_t v) {
  int32_t s = v & 1023;
  if (s >= 512) s -= 1024;
  s += 128;
  return (uint8_t)(s < 0 ? 0 : (s > 255 ? 255 : s));
}

SPM_HD void idct_1d(const int32_t in[8], int32_t out[8], int shift_even_dc, int descale_bits) {
  constexpr int32_t F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633,
                    F_1_501 = 12299, F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995,
                    F_3_072 = 25172;
  // even part
  int32_t z2 = in[2], z3 = in[6];
  int32_t z1 = (z2 + z3) * F_0_541;
  int32_t tmp2 = z1 + z3 * (-F_1_847);
  int32_t tmp3 = z1 + z2 * F_0_765;
  z2 = in[0]; z3 = in[4];
  int32_t tmp0 = (z2 + z3) * (1 << shift_even_dc);
  int32_t tmp1 = (z2 - z3) * (1 << shift_even_dc);
  const int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  // odd part
  tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int32_t z4 = tmp1 + tmp3;
  const int32_t z5 = (z3 + z4) * F_1_175;
  tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
  z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
  z3 += z5; z4 += z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  out[0] = descale(tmp10 + tmp3, descale_bits); out[7] = descale(tmp10 - tmp3, descale_bits);
  out[1] = descale(tmp11 + tmp2, descale_bits); out[6] = descale(tmp11 - tmp2, descale_bits);
  out[2] = descale(tmp12 + tmp1, descale_bits); out[5] = descale(tmp12 - tmp1, descale_bits);
  out[3] = descale(tmp13 + tmp0, descale_bits); out[4] = descale(tmp13 - tmp0, descale_bits);
}

// coef[64] (natural order) * quant[64] -> 8x8 samples written to out[r * stride + c]
SPM_HD void idct_block(const int16_t* coef, const uint16_t* quant, uint8_t* out, int stride) {
  int32_t ws[64];
  for (int c = 0; c < 8; ++c) {   // pass 1: columns, results scaled up by 2^PASS1_BITS (2)
    int32_t in[8], o[8];
    for (int r = 0; r < 8; ++r) in[r] = (int32_t)coef[r * 8 + c] * (int32_t)quant[r * 8 + c];
    idct_1d(in, o, 13, 13 - 2);
    for (int r = 0; r < 8; ++r) ws[r * 8 + c] = o[r];
  }
  for (int r = 0; r < 8; ++r) {   // pass 2: rows, remove the 2 + 3 extra bits, centre on 128
    int32_t in[8], o[8];
    for (int c = 0; c < 8; ++c) in[c] = ws[r * 8 + c];
    idct_1d(in, o, 13, 13 + 2 + 3);
    for (int c = 0; c < 8; ++c) out[r * stride + c] = idct_range(o[c]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// chroma upsampling ("fancy": 3/4 nearer + 1/4 further sample in each axis) and colour conversion
// ------------------------------------------------------------------------------------------------------------
// vertical blend of a 2:1 subsampled plane for output row y: 3 * nearer row + further row; rows are clamped to the
// last real subsampled row (libjpeg replicates the edge rows), column x is a subsampled column
SPM_HD int32_t v_blend(const uint8_t* plane, int stride, int rows, int y, int x) {
  const int r0 = y >> 1;
  int r1 = (y & 1) ? r0 + 1 : r0 - 1;
  r1 = r1 < 0 ? 0 : (r1 > rows - 1 ? rows - 1 : r1);
  return 3 * (int32_t)plane[r0 * stride + x] + (int32_t)plane[r1 * stride + x];
}
// h2v2 fancy upsampling of one output sample (x, y); cols / rows = real extent of the subsampled plane
SPM_HD int32_t upsample_h2v2(const uint8_t* plane, int stride, int cols, int rows, int x, int y) {
  const int c = x >> 1;
  const int32_t cur = v_blend(plane, stride, rows, y, c);
  if (x & 1) {
    if (c == cols - 1) return (cur * 4 + 7) >> 4;
    return (cur * 3 + v_blend(plane, stride, rows, y, c + 1) + 7) >> 4;
  }
  if (c == 0) return (cur * 4 + 8) >> 4;
  return (cur * 3 + v_blend(plane, stride, rows, y, c - 1) + 8) >> 4;
}
// h2v1 fancy upsampling (4:2:2)
SPM_HD int32_t upsample_h2v1(const uint8_t* row, int cols, int x) {
  const int c = x >> 1;
  const int32_t cur = row[c];
  if (x & 1) return c == cols - 1 ? cur : (3 * cur + (int32_t)row[c + 1] + 2) >> 2;
  return c == 0 ? cur : (3 * cur + (int32_t)row[c - 1] + 1) >> 2;
}

SPM_HD uint8_t clamp255(int32_t v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
// 16-bit fixed-point YCbCr -> RGB (the table entries of jdcolor.c evaluated on the fly; >> is arithmetic)
SPM_HD void ycc_to_rgb(int32_t y, int32_t cb, int32_t cr, uint8_t* rgb) {
  const int32_t xb = cb - 128, xr = cr - 128;
  const int32_t cr_r = (91881 * xr + 32768) >> 16;
  const int32_t cb_b = (116130 * xb + 32768) >> 16;
  const int32_t g = (-22554 * xb + 32768 + (-46802) * xr) >> 16;
  rgb[0] = clamp255(y + cr_r);
  rgb[1] = clamp255(y + g);
  rgb[2] = clamp255(y + cb_b);
}

}  // namespace jpeg
}  // namespace spm
