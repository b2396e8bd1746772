// Soft-min DP of OTAM_cum_dist_v2 (models/myRes.py:821-855) as a warp-level anti-diagonal wavefront; shared by the
// streaming kernel (otam.cu) and the tensor-core batch kernel (otam_mma.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>

namespace spm {
namespace otam_dp {
constexpr float LBDA = 0.5f;
constexpr float INV_LBDA = 2.0f;

// soft-min with lambda = 0.5 is taken in the min-shifted form: -l*log(sum exp(-x/l)) = min - l*log(sum exp(-(x-min)/l)).
// Mathematically identical to the reference's expression (myRes.py:838-853), better conditioned (all exponents <= 0,
// the sum lies in [1, 3]) -- which also makes the fast ex2/lg2 units accurate to ~1e-6 here.

// A warp runs 32 / (T + 2) DPs side by side: lane = seg * (T + 2) + m, m = padded column.  The neighbours come from
// lane - 1 by a full-width __shfl_up; a segment's column 0 is the constant 0 and never looks at what it receives
// from the segment before it, so segments need no power-of-two width (T = 8: three DPs per warp).
// dw: the [T][T] distance table of this lane's (query, class) pair; dir 0 walks it as dist[l][j], dir 1 transposed
// (dist[j][l]) -- l = row of the DP (0..T-1), j = unpadded column (0..T-1).  The lane of column T+1 returns
// C[T-1, T+1]; the value returned by the other lanes is meaningless.
//
// Branch-free on purpose: the DPs of a warp differ in direction and every diagonal mixes the four cell kinds
// (column 0, top row, three-neighbour edge columns 1 / T+1, two-neighbour interior).  Written with branches the
// warp executed each kind one after the other and both directions one after the other (158 instructions per
// diagonal step in ncu); here every lane runs the same ~30: the interior's missing vertical neighbour is +inf
// (its exponential is exactly 0), the soft-min is taken in sorted form (the minimum's exponential is exactly 1, so
// two ex2 serve both the two- and the three-neighbour case), and the distance index advances by a per-lane stride.
__device__ __forceinline__ int otam_dps_per_warp(int T) { return 32 / (T + 2); }

__device__ __forceinline__ float otam_wavefront(int T, int m, bool valid, const float* __restrict__ dw, int dir) {
  constexpr float K_EX2 = INV_LBDA * 1.4426950408889634f;   // exp(x / lambda) = ex2(x * K_EX2)
  constexpr float K_LG2 = LBDA * 0.6931471805599453f;       // lambda * ln(s) = K_LG2 * lg2(s)
  const bool has_d = m >= 1 && m <= T;    // columns 0 and T+1 are the zero padding
  const bool edge = m == 1 || m == T + 1;
  const bool col = valid && m >= 1 && m <= T + 1;   // column 0 is never written (stays 0)
  // cell (l, m) adds dist element (l, m-1); l = k - m on diagonal k
  const int stride = dir ? 1 : T;
  int idx = dir ? (m - 1) * T - m : -m * T + m - 1;
  float v1 = 0.f, v2 = 0.f;               // this lane's last / second-to-last computed cells
  for (int k = 0; k <= 2 * T; ++k, idx += stride) {
    const float left = __shfl_up_sync(0xffffffffu, v1, 1);   // C[l,   m-1]
    const float diag = __shfl_up_sync(0xffffffffu, v2, 1);   // C[l-1, m-1]
    const int l = k - m;
    const bool active = col && l >= 0 && l < T;
    const float d = (active && has_d) ? dw[idx] : 0.f;
    const float up = edge ? v1 : __int_as_float(0x7f800000);      // C[l-1, m] only in the edge columns
    const float lo = fminf(diag, left), hi = fmaxf(diag, left);
    const float mn = fminf(lo, up), mx = fmaxf(hi, up), md = fmaxf(lo, fminf(hi, up));
    const float s = 1.f + exp2f((mn - md) * K_EX2) + exp2f((mn - mx) * K_EX2);
    const float soft = mn - K_LG2 * __log2f(s);
    const float c = d + (l == 0 ? left : soft);                   // top row: plain prefix sum
    v2 = active ? v1 : v2;
    v1 = active ? c : v1;
  }
  return v1;
}

// The same recurrence in the exponent domain (r02).  With E = exp(-C / lambda) the soft-min becomes LINEAR:
//   C[l,m] = d[l,m] - lambda * log(sum_prev exp(-C_prev / lambda))   <=>   E[l,m] = exp(-d[l,m] / lambda) * sum_prev E_prev
// (top row: C = d + C_left <=> E = e * E_left), so the dependent chain of a diagonal step is shuffle -> add -> multiply
// instead of shuffle -> min/max sort -> two ex2 -> lg2 -> fma (~45 instead of ~270 cycles per step measured in the fused
// kernel); the exponentials exp(-d / lambda) do not depend on the chain and are taken one step ahead.  One lg2 at the
// end gives C[T-1, T+1] back.  Range: d = 1 - cos lies in (0, 2), a path has at most l + m + 1 cells and a cell at most
// 2^(l+m) paths, so E spans 2^(-5.8 (l+m)) .. 2^(l+m); scaled by g(l,m) = 4^(l+m) -- exact powers of two folded into the
// recurrence as the constants 4 (left / up neighbour) and 16 (diagonal neighbour) -- it stays inside 2^(+-121) for
// l + m <= 2T <= 32: no overflow, no underflow, no per-step rescaling for T <= 16.  Longer sequences keep the form above.
constexpr int OTAM_EXP_MAX_T = 16;

// PRE: the table already holds exp(-d / lambda) (otam_exp_of_dist applied by whoever wrote it), so a step is one LDS away
// from its factor; otherwise the exponential is taken here, one step ahead of its use.
__device__ __forceinline__ float otam_exp_of_dist(float d) { return exp2f(-d * (INV_LBDA * 1.4426950408889634f)); }

template <bool PRE>
__device__ __forceinline__ float otam_wavefront_exp(int T, int m, bool valid, const float* __restrict__ dw, int dir) {
  constexpr float K_EX2 = INV_LBDA * 1.4426950408889634f;
  constexpr float K_LG2 = LBDA * 0.6931471805599453f;
  constexpr float G1 = 4.f, G2 = 16.f;
  const bool has_d = m >= 1 && m <= T;
  const bool edge = m == 1 || m == T + 1;
  const bool col = valid && m >= 1 && m <= T + 1;
  const bool col0 = m == 0;
  const int stride = dir ? 1 : T;
  int idx = dir ? (m - 1) * T - m : -m * T + m - 1;
  // column 0 is the constant C = 0, i.e. E * g = 4^l: its lane carries 4^(k-1), 4^(k-2) into iteration k
  float v1 = col0 ? 1.f / G1 : 0.f, v2 = col0 ? 1.f / G2 : 0.f;
  // the exponential of iteration 0 (l = -m < 0 for every column that has one)
  float e = 1.f;
  for (int k = 0; k <= 2 * T; ++k) {
    const float left = __shfl_up_sync(0xffffffffu, v1, 1);   // E[l,   m-1]
    const float diag = __shfl_up_sync(0xffffffffu, v2, 1);   // E[l-1, m-1]
    const int l = k - m;
    const bool active = col && l >= 0 && l < T;
    // next iteration's exponential, off the chain
    idx += stride;
    const bool act_n = col && has_d && l + 1 >= 0 && l + 1 < T;
    const float e_next = act_n ? (PRE ? dw[idx] : exp2f(-dw[idx] * K_EX2)) : 1.f;
    const float up = edge ? v1 : 0.f;                        // E[l-1, m] only in the edge columns
    const float sum = l == 0 ? G1 * left : fmaf(G2, diag, G1 * (left + up));
    const float c = e * sum;
    v2 = (active || col0) ? v1 : v2;
    v1 = col0 ? v1 * G1 : (active ? c : v1);
    e = e_next;
  }
  return K_LG2 * ((float)(4 * T) - __log2f(v1));   // -lambda * ln(E), E = v1 / 4^(2T)
}

// NW independent wavefronts per lane in one loop (tables of exp(-d / lambda), i.e. PRE): a diagonal step is a chain of
// ~10 dependent fixed-latency instructions behind two shuffles, and a warp that is alone on its scheduler spends most of
// the step waiting on them (ncu source page of the fused kernel, r02: 190 cycles per step, 70 % `wait` / `short_sb`);
// the other chains fill those slots.
template <int NW>
__device__ __forceinline__ void otam_wavefront_exp_pre_n(int T, int m, const bool (&valid)[NW],
                                                         const float* const (&dw)[NW], const int (&dir)[NW],
                                                         float (&res)[NW]) {
  constexpr float K_LG2 = LBDA * 0.6931471805599453f;
  constexpr float G1 = 4.f, G2 = 16.f;
  const bool has_d = m >= 1 && m <= T;
  const bool edge = m == 1 || m == T + 1;
  const bool col0 = m == 0;
  bool col[NW];
  int stride[NW], idx[NW];
  float v1[NW], v2[NW], e[NW];
#pragma unroll
  for (int u = 0; u < NW; ++u) {
    col[u] = valid[u] && m >= 1 && m <= T + 1;
    stride[u] = dir[u] ? 1 : T;
    idx[u] = dir[u] ? (m - 1) * T - m : -m * T + m - 1;
    v1[u] = col0 ? 1.f / G1 : 0.f;
    v2[u] = col0 ? 1.f / G2 : 0.f;
    e[u] = 1.f;
  }
  for (int k = 0; k <= 2 * T; ++k) {
    const int l = k - m;
    const bool in_rows = l >= 0 && l < T, next_in_rows = l + 1 >= 0 && l + 1 < T, top = l == 0;
#pragma unroll
    for (int u = 0; u < NW; ++u) {
      const float left = __shfl_up_sync(0xffffffffu, v1[u], 1);
      const float diag = __shfl_up_sync(0xffffffffu, v2[u], 1);
      const bool active = col[u] && in_rows;
      idx[u] += stride[u];
      const float e_next = (col[u] && has_d && next_in_rows) ? dw[u][idx[u]] : 1.f;
      const float up = edge ? v1[u] : 0.f;
      const float sum = top ? G1 * left : fmaf(G2, diag, G1 * (left + up));
      const float c = e[u] * sum;
      v2[u] = (active || col0) ? v1[u] : v2[u];
      v1[u] = col0 ? v1[u] * G1 : (active ? c : v1[u]);
      e[u] = e_next;
    }
  }
#pragma unroll
  for (int u = 0; u < NW; ++u) res[u] = K_LG2 * ((float)(4 * T) - __log2f(v1[u]));
}

// the wavefront the kernels call: exponent domain up to T = 16 unless `force_log` (SPM_OTAM_DP=log, the cross-check)
__device__ __forceinline__ float otam_wavefront_auto(int T, int m, bool valid, const float* __restrict__ dw, int dir,
                                                     int force_log) {
  if (T <= OTAM_EXP_MAX_T && !force_log) return otam_wavefront_exp<false>(T, m, valid, dw, dir);
  return otam_wavefront(T, m, valid, dw, dir);
}
// tables written through otam_table_value(): exp(-d / lambda) when the exponent-domain form will read them
__device__ __forceinline__ bool otam_exp_mode(int T, int force_log) { return T <= OTAM_EXP_MAX_T && !force_log; }
__device__ __forceinline__ float otam_table_value(float d, bool exp_mode) { return exp_mode ? otam_exp_of_dist(d) : d; }
__device__ __forceinline__ float otam_wavefront_table(int T, int m, bool valid, const float* __restrict__ dw, int dir,
                                                      bool exp_mode) {
  if (exp_mode) return otam_wavefront_exp<true>(T, m, valid, dw, dir);
  return otam_wavefront(T, m, valid, dw, dir);
}

// host side: SPM_OTAM_DP=log keeps the log-domain wavefront for every T (cross-check of the exponent-domain form)
inline int otam_dp_force_log() {
  static const int v = [] { const char* e = getenv("SPM_OTAM_DP"); return (e != nullptr && strcmp(e, "log") == 0) ? 1 : 0; }();
  return v;
}
}  // namespace otam_dp
}  // namespace spm
