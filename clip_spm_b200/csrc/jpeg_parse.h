// Host-side JPEG container parsing (ITU T.81 Annex B): markers, quantisation / Huffman tables, frame and scan headers,
// and removal of byte stuffing / restart markers from the entropy-coded segment, so that the device decodes a plain bit
// stream.  Baseline (SOF0) and extended-sequential Huffman (SOF1) 8-bit YCbCr images with one interleaved scan -- what
// `ffmpeg`-extracted frame dumps (the reference's datasets, video_reader.py:227-230) contain.  Plain C++: shared by the
// library (jpeg_decode.cu) and the host check build (oracle/jpeg_check.cu).
#pragma once
#include <cstring>
#include <string>
#include <vector>

#include "jpeg_core.cuh"

namespace spm {
namespace jpeg {

struct Parsed {
  ImageDesc d;
  std::vector<uint8_t> scan;             // entropy-coded data with FF00 -> FF and the RSTn markers removed
  std::vector<int32_t> interval_start;   // byte offset in `scan` where each restart interval begins (first = 0)
  std::string err;
};

inline void build_huff(const uint8_t counts[16], const uint8_t* symbols, int n_symbols, HuffTable* t) {
  // T.81 Annex C: canonical code assignment by increasing length
  std::memset(t, 0, sizeof(*t));
  int32_t code = 0;
  int k = 0;
  for (int l = 1; l <= 16; ++l) {
    t->valptr[l] = k;
    t->mincode[l] = code;
    for (int i = 0; i < counts[l - 1] && k < n_symbols; ++i, ++k, ++code) {
      t->huffval[k] = symbols[k];
      if (l <= 8) {   // every 8-bit prefix that starts with this code resolves in one lookup
        const int shift = 8 - l;
        for (int f = 0; f < (1 << shift); ++f) t->look[(code << shift) | f] = (uint16_t)((l << 8) | symbols[k]);
      }
    }
    t->maxcode[l] = counts[l - 1] ? code - 1 : -1;
    code <<= 1;
  }
  t->maxcode[17] = 0x7fffffff;
}

inline int fail(Parsed* out, const char* msg) { out->err = msg; return 1; }

// Returns 0 on success; on failure out->err says why (unsupported feature or corrupt file).
inline int parse_jpeg(const uint8_t* p, size_t n, Parsed* out) {
  ImageDesc& d = out->d;
  std::memset(&d, 0, sizeof(d));
  out->scan.clear();
  out->interval_start.clear();
  if (n < 4 || p[0] != 0xFF || p[1] != 0xD8) return fail(out, "not a JPEG file (no SOI marker)");
  size_t pos = 2;
  int comp_id[MAX_COMP] = {0, 0, 0};
  bool have_sof = false, have_sos = false;
  bool have_q[4] = {false, false, false, false}, have_dc[2] = {false, false}, have_ac[2] = {false, false};
  while (pos + 4 <= n && !have_sos) {
    if (p[pos] != 0xFF) return fail(out, "corrupt JPEG (marker expected)");
    while (pos < n && p[pos] == 0xFF) ++pos;   // fill bytes
    if (pos >= n) break;
    const int m = p[pos++];
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;   // standalone markers
    if (m == 0xD9) break;
    if (pos + 2 > n) return fail(out, "corrupt JPEG (truncated segment)");
    const size_t len = ((size_t)p[pos] << 8) | p[pos + 1];
    if (len < 2 || pos + len > n) return fail(out, "corrupt JPEG (bad segment length)");
    const uint8_t* s = p + pos + 2;
    const size_t sl = len - 2;
    if (m == 0xDB) {                                     // DQT
      size_t i = 0;
      while (i < sl) {
        const int pq = s[i] >> 4, tq = s[i] & 15;
        ++i;
        if (tq > 3 || i + (pq ? 128 : 64) > sl) return fail(out, "corrupt JPEG (DQT)");
        for (int k = 0; k < 64; ++k) {
          const int v = pq ? ((s[i] << 8) | s[i + 1]) : s[i];
          i += pq ? 2 : 1;
          d.quant[tq][zigzag_to_natural(k)] = (uint16_t)v;
        }
        have_q[tq] = true;
      }
    } else if (m == 0xC0 || m == 0xC1) {                 // SOF0 / SOF1
      if (sl < 6 || s[0] != 8) return fail(out, "unsupported JPEG (sample precision is not 8 bits)");
      d.height = (s[1] << 8) | s[2];
      d.width = (s[3] << 8) | s[4];
      d.n_comp = s[5];
      if (d.n_comp != 3) return fail(out, "unsupported JPEG (not a 3-component YCbCr image)");
      if (sl < 6 + 3 * (size_t)d.n_comp || d.width < 3 || d.height < 3) return fail(out, "corrupt JPEG (SOF)");
      for (int c = 0; c < d.n_comp; ++c) {
        comp_id[c] = s[6 + 3 * c];
        d.hs[c] = s[7 + 3 * c] >> 4;
        d.vs[c] = s[7 + 3 * c] & 15;
        d.tq[c] = s[8 + 3 * c];
        if (d.tq[c] > 3) return fail(out, "corrupt JPEG (quantisation table index)");
      }
      const bool luma_ok = (d.hs[0] == 1 && d.vs[0] == 1) || (d.hs[0] == 2 && d.vs[0] == 1) || (d.hs[0] == 2 && d.vs[0] == 2);
      if (!luma_ok || d.hs[1] != 1 || d.vs[1] != 1 || d.hs[2] != 1 || d.vs[2] != 1)
        return fail(out, "unsupported JPEG (chroma subsampling other than 4:4:4, 4:2:2, 4:2:0)");
      have_sof = true;
    } else if (m == 0xC2 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
      return fail(out, "unsupported JPEG (progressive, lossless or arithmetic-coded)");
    } else if (m == 0xC4) {                              // DHT
      size_t i = 0;
      while (i < sl) {
        if (i + 17 > sl) return fail(out, "corrupt JPEG (DHT)");
        const int tc = s[i] >> 4, th = s[i] & 15;
        int total = 0;
        for (int k = 0; k < 16; ++k) total += s[i + 1 + k];
        if (tc > 1 || th > 1 || total > 256 || i + 17 + total > sl)
          return fail(out, "unsupported JPEG (more than two Huffman tables per class, or a corrupt DHT)");
        build_huff(s + i + 1, s + i + 17, total, tc == 0 ? &d.dc[th] : &d.ac[th]);
        (tc == 0 ? have_dc : have_ac)[th] = true;
        i += 17 + total;
      }
    } else if (m == 0xDD) {                              // DRI
      if (sl < 2) return fail(out, "corrupt JPEG (DRI)");
      d.restart_interval = (s[0] << 8) | s[1];
    } else if (m == 0xEE) {                              // Adobe APP14: colour transform flag
      if (sl >= 12 && std::memcmp(s, "Adobe", 5) == 0 && s[11] != 1)
        return fail(out, "unsupported JPEG (Adobe colour transform other than YCbCr)");
    } else if (m == 0xDA) {                              // SOS
      if (!have_sof) return fail(out, "corrupt JPEG (SOS before SOF)");
      if (sl < 1 || s[0] != d.n_comp || sl < 1 + 2 * (size_t)d.n_comp + 3)
        return fail(out, "unsupported JPEG (non-interleaved scans)");
      for (int c = 0; c < d.n_comp; ++c) {
        if (s[1 + 2 * c] != comp_id[c]) return fail(out, "unsupported JPEG (scan component order)");
        d.td[c] = s[2 + 2 * c] >> 4;
        d.ta[c] = s[2 + 2 * c] & 15;
        if (d.td[c] > 1 || d.ta[c] > 1 || !have_dc[d.td[c]] || !have_ac[d.ta[c]] || !have_q[d.tq[c]])
          return fail(out, "corrupt JPEG (scan refers to a table that was not defined)");
      }
      have_sos = true;
    }
    pos += len;
  }
  if (!have_sos) return fail(out, "corrupt JPEG (no scan)");
  // ---- geometry
  const int hmax = d.hs[0], vmax = d.vs[0];
  d.mcus_x = (d.width + 8 * hmax - 1) / (8 * hmax);
  d.mcus_y = (d.height + 8 * vmax - 1) / (8 * vmax);
  for (int c = 0; c < d.n_comp; ++c) {
    d.blocks_x[c] = d.mcus_x * d.hs[c];
    d.blocks_y[c] = d.mcus_y * d.vs[c];
  }
  // ---- entropy-coded segment: remove stuffing, cut at restart markers
  out->scan.reserve(n - pos);
  out->interval_start.push_back(0);
  while (pos < n) {
    const uint8_t b = p[pos++];
    if (b != 0xFF) { out->scan.push_back(b); continue; }
    if (pos >= n) break;
    const uint8_t m = p[pos];
    if (m == 0x00) { out->scan.push_back(0xFF); ++pos; }
    else if (m >= 0xD0 && m <= 0xD7) { out->interval_start.push_back((int32_t)out->scan.size()); ++pos; }
    else if (m == 0xFF) { /* fill byte: the next iteration looks at the second FF */ }
    else break;   // EOI or another segment: end of the scan
  }
  const long long total_mcus = (long long)d.mcus_x * d.mcus_y;
  const long long want = d.restart_interval > 0 ? (total_mcus + d.restart_interval - 1) / d.restart_interval : 1;
  if ((long long)out->interval_start.size() < want) return fail(out, "corrupt JPEG (missing restart markers)");
  out->interval_start.resize((size_t)want);
  d.n_intervals = (int32_t)want;
  return 0;
}

}  // namespace jpeg
}  // namespace spm
