/* ORACLE -- TEST INFRASTRUCTURE ONLY (plain C restatement, never linked into the product library).
 *
 * cos_sim + OTAM soft-DTW of the metric tail, restated from the reference in scalar C with double accumulation:
 *   models/myRes.py:756-765        cos_sim(x, y) = x.y / (|x||y| + 0.01)
 *   models/myRes.py:821-855        OTAM_cum_dist_v2 (zero-padded [T, T+2] grid, lambda = 0.5, asymmetric recurrences)
 *   models/model_clipspm.py:348-362  otam_distance = OTAM(1 - sim) + OTAM((1 - sim)^T)   (unless SINGLE_DIRECT)
 * Pinned by tests/test_oracle_cpu.py against oracle/clipspm_oracle.py, which is itself pinned against the executed
 * reference (oracle/pin_against_reference.py).  Built by oracle/Makefile into oracle/_ref/libotam_ref.so. */
#include <math.h>
#include <stdlib.h>

static double softmin2(double a, double b, double l) { return -l * log(exp(-a / l) + exp(-b / l)); }
static double softmin3(double a, double b, double c, double l) {
  return -l * log(exp(-a / l) + exp(-b / l) + exp(-c / l));
}

/* d: [L][M] row-major distances; returns cum[L-1][M+1] of the padded grid (myRes.py:821-855) */
static double otam_cum_dist_v2(const double* d, int L, int M, int transposed) {
  const double lb = 0.5;
  const int M2 = M + 2;
  double* c = (double*)calloc((size_t)L * M2, sizeof(double));
#define DD(l, m) (((m) == 0 || (m) == M + 1) ? 0.0 : (transposed ? d[((m)-1) * L + (l)] : d[(l)*M + ((m)-1)]))
  for (int m = 1; m < M2; ++m) c[m] = DD(0, m) + c[m - 1];                       /* top row: prefix sum (:832-835) */
  for (int l = 1; l < L; ++l) {
    c[l * M2 + 1] = DD(l, 1) + softmin3(c[(l - 1) * M2], c[(l - 1) * M2 + 1], c[l * M2], lb);          /* :838-842 */
    for (int m = 2; m < M2 - 1; ++m)
      c[l * M2 + m] = DD(l, m) + softmin2(c[(l - 1) * M2 + m - 1], c[l * M2 + m - 1], lb);             /* :845-847 */
    c[l * M2 + M2 - 1] = DD(l, M2 - 1) + softmin3(c[(l - 1) * M2 + M2 - 2], c[(l - 1) * M2 + M2 - 1],
                                                  c[l * M2 + M2 - 2], lb);                             /* :850-853 */
  }
#undef DD
  const double r = c[(L - 1) * M2 + M2 - 1];
  free(c);
  return r;
}

/* support [W][T][D], target [Q][T][D] (float) -> out [Q][W] */
void otam_distance_ref(const float* support, const float* target, int W, int Q, int T, int D, int single_direct,
                       float* out) {
  double* dist = (double*)malloc((size_t)T * T * sizeof(double));
  for (int q = 0; q < Q; ++q) {
    for (int w = 0; w < W; ++w) {
      for (int a = 0; a < T; ++a) {
        const float* x = target + ((size_t)q * T + a) * D;
        double xn = 0;
        for (int k = 0; k < D; ++k) xn += (double)x[k] * x[k];
        for (int b = 0; b < T; ++b) {
          const float* y = support + ((size_t)w * T + b) * D;
          double yn = 0, dot = 0;
          for (int k = 0; k < D; ++k) { yn += (double)y[k] * y[k]; dot += (double)x[k] * y[k]; }
          dist[a * T + b] = 1.0 - dot / (sqrt(xn) * sqrt(yn) + 0.01);
        }
      }
      double r = otam_cum_dist_v2(dist, T, T, 0);
      if (!single_direct) r += otam_cum_dist_v2(dist, T, T, 1);
      out[q * W + w] = (float)r;
    }
  }
  free(dist);
}
