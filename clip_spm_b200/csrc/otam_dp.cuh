// Soft-min DP of OTAM_cum_dist_v2 (models/myRes.py:821-855) as a warp-level anti-diagonal wavefront; shared by the
// streaming kernel (otam.cu) and the tensor-core batch kernel (otam_mma.cu).
#pragma once
#include <cuda_runtime.h>

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
}  // namespace otam_dp
}  // namespace spm
