/* CPU oracle, C restatement of EdgeFEM's per-tet assembly loop.  TEST INFRASTRUCTURE ONLY: loaded by
 * tests/ and by bench.py's cpu_baseline leg through ctypes (oracle/edgefem_oracle_c.py); the product
 * path never links or calls it.
 *
 * Follows the reference line by line (paths relative to /root/reference):
 *   gradients_and_volume   src/edge_basis.cpp:14-26   (Eigen 3x3 inverse = cofactors / det)
 *   whitney_curl_curl      src/edge_basis.cpp:48-63
 *   whitney_mass           src/edge_basis.cpp:66-86, lambda_int :28-30
 *   triplet loop           src/assemble_maxwell.cpp:114-204 (no PML; materials resolved per tet by the caller)
 * The triplets are returned unsummed in the reference's emission order (tet-major, i-major, j-minor);
 * summation (Eigen setFromTriplets) is done by the caller.
 */
#include <math.h>
#include <stdint.h>

static const int EP[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

static void cross3(const double *a, const double *b, double *c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static void gradients_and_volume(const double v[4][3], double g[4][3], double *V) {
  double b0[3], b1[3], b2[3], c12[3], c20[3], c01[3];
  for (int k = 0; k < 3; ++k) {
    b0[k] = v[0][k] - v[3][k];
    b1[k] = v[1][k] - v[3][k];
    b2[k] = v[2][k] - v[3][k];
  }
  cross3(b1, b2, c12);
  cross3(b2, b0, c20);
  cross3(b0, b1, c01);
  const double det = dot3(b0, c12), inv = 1.0 / det;
  for (int k = 0; k < 3; ++k) {
    g[0][k] = c12[k] * inv;
    g[1][k] = c20[k] * inv;
    g[2][k] = c01[k] * inv;
    g[3][k] = -g[0][k] - g[1][k] - g[2][k];
  }
  *V = fabs(det) / 6.0;
}

void efo_element_matrices(const double v[4][3], double K[6][6], double M[6][6]) {
  double g[4][3], V, c[6][3];
  gradients_and_volume(v, g, &V);
  for (int i = 0; i < 6; ++i) {
    cross3(g[EP[i][0]], g[EP[i][1]], c[i]);
    for (int k = 0; k < 3; ++k) c[i][k] *= 2.0;
  }
  for (int i = 0; i < 6; ++i) {
    const int a = EP[i][0], b = EP[i][1];
    for (int j = 0; j < 6; ++j) {
      const int cc = EP[j][0], d = EP[j][1];
      K[i][j] = V * dot3(c[i], c[j]);
      double t = 0.0;
      t += dot3(g[b], g[d]) * ((a == cc) ? V / 10.0 : V / 20.0);
      t -= dot3(g[b], g[cc]) * ((a == d) ? V / 10.0 : V / 20.0);
      t -= dot3(g[a], g[d]) * ((b == cc) ? V / 10.0 : V / 20.0);
      t += dot3(g[a], g[cc]) * ((b == d) ? V / 10.0 : V / 20.0);
      M[i][j] = t;
    }
  }
}

/* val = K/mu - k0^2 eps M, times s_i s_j  (src/assemble_maxwell.cpp:184-203), complex as (re,im) pairs */
void efo_volume_triplets(int64_t n_tet, const double *xyz, const int32_t *tet_nodes, const int32_t *tet_edges, const int32_t *tet_orient,
                         const double *eps_c128, const double *mu_c128, double omega, int32_t *rows, int32_t *cols, double *vals_c128) {
  const double c0 = 299792458.0, k0 = omega / c0, k0sq = k0 * k0;
  for (int64_t t = 0; t < n_tet; ++t) {
    double v[4][3], K[6][6], M[6][6];
    for (int i = 0; i < 4; ++i)
      for (int k = 0; k < 3; ++k) v[i][k] = xyz[3 * (int64_t)tet_nodes[4 * t + i] + k];
    efo_element_matrices(v, K, M);
    const double er = eps_c128[2 * t], ei = eps_c128[2 * t + 1], mr = mu_c128[2 * t], mi = mu_c128[2 * t + 1];
    const double den = mr * mr + mi * mi;
    const double ir = mr / den, ii = -mi / den; /* 1/mu */
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        const double s = (double)(tet_orient[6 * t + i] * tet_orient[6 * t + j]);
        const int64_t o = 36 * t + 6 * i + j;
        rows[o] = tet_edges[6 * t + i];
        cols[o] = tet_edges[6 * t + j];
        vals_c128[2 * o] = (K[i][j] * ir - k0sq * er * M[i][j]) * s;
        vals_c128[2 * o + 1] = (K[i][j] * ii - k0sq * ei * M[i][j]) * s;
      }
  }
}
