// field post-processing of the solution vector, on the device.
//   efb_huygens_eval   tangential E, H at the centroid of every triangle of a Huygens surface from the parent tet's six
//                      edge DOFs (Whitney interpolation + piecewise-constant curl): src/post/huygens_surface.cpp:59-135,
//                      src/edge_basis.cpp:132-192 (compute_barycentric, evaluate_edge_field), :33-46 (whitney_edge_curls)
//   efb_stratton_chu   far field E_theta, E_phi of the Love currents over a set of directions:
//                      src/post/ntf.cpp:86-203 (stratton_chu_2d / _3d share one kernel)
// One thread per triangle; one CTA per direction with a fixed-order block reduction (deterministic).
#include "common.cuh"

namespace efb {
namespace {

struct V3 { double x, y, z; };
struct C3 { c128 x, y, z; };
__device__ __forceinline__ V3 v3(double4 p) { return {p.x, p.y, p.z}; }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 scale(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ c128 cdotr(C3 a, V3 b) {  // plain projection sum a_i b_i
  return cmake(a.x.x * b.x + a.y.x * b.y + a.z.x * b.z, a.x.y * b.x + a.y.y * b.y + a.z.y * b.z);
}
// plain (non-conjugating) cross products real x complex and complex x real (ntf.cpp:16-30)
__device__ __forceinline__ C3 cross_rc(V3 a, C3 b) {
  return {csub(cscale(a.y, b.z), cscale(a.z, b.y)), csub(cscale(a.z, b.x), cscale(a.x, b.z)), csub(cscale(a.x, b.y), cscale(a.y, b.x))};
}
__device__ __forceinline__ C3 cross_cr(C3 a, V3 b) {
  return {csub(cscale(b.z, a.y), cscale(b.y, a.z)), csub(cscale(b.x, a.z), cscale(b.z, a.x)), csub(cscale(b.y, a.x), cscale(b.x, a.y))};
}

// gradients of the barycentric functions and |V| like gradients_and_volume (src/edge_basis.cpp:14-26)
__device__ void tet_gradients(const V3 (&X)[4], V3 (&g)[4], double &vol) {
  const V3 a = sub(X[0], X[3]), b = sub(X[1], X[3]), c = sub(X[2], X[3]);
  const V3 bc = cross(b, c), ca = cross(c, a), ab = cross(a, b);
  const double det = dot(a, bc);
  g[0] = scale(bc, 1.0 / det);
  g[1] = scale(ca, 1.0 / det);
  g[2] = scale(ab, 1.0 / det);
  g[3] = {-(g[0].x + g[1].x + g[2].x), -(g[0].y + g[1].y + g[2].y), -(g[0].z + g[1].z + g[2].z)};
  vol = fabs(det) / 6.0;
}

__global__ void k_huygens(const double4 *__restrict__ xyz, const int4 *__restrict__ tet_nodes, const uint8_t *__restrict__ tet_sign,
                          const c128 *__restrict__ x, int n_tri, const int32_t *__restrict__ tri_nodes, const int32_t *__restrict__ tri_tet,
                          const int32_t *__restrict__ tri_tet_edges, c128 inv_jwmu, double *__restrict__ r_out, double *__restrict__ n_out,
                          c128 *__restrict__ E_out, c128 *__restrict__ H_out, double *__restrict__ area_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tri) return;
  const V3 v0 = v3(xyz[tri_nodes[3 * i]]), v1 = v3(xyz[tri_nodes[3 * i + 1]]), v2 = v3(xyz[tri_nodes[3 * i + 2]]);
  const V3 cen = {(v0.x + v1.x + v2.x) / 3.0, (v0.y + v1.y + v2.y) / 3.0, (v0.z + v1.z + v2.z) / 3.0};
  const V3 cr = cross(sub(v1, v0), sub(v2, v0));
  const double nrm = sqrt(dot(cr, cr));
  const double area = 0.5 * nrm;
  V3 n = nrm > 0.0 ? scale(cr, 1.0 / nrm) : cr;
  const int t = tri_tet[i];
  const int4 tn = tet_nodes[t];
  V3 X[4] = {v3(xyz[tn.x]), v3(xyz[tn.y]), v3(xyz[tn.z]), v3(xyz[tn.w])};
  const V3 tc = {(X[0].x + X[1].x + X[2].x + X[3].x) / 4.0, (X[0].y + X[1].y + X[2].y + X[3].y) / 4.0, (X[0].z + X[1].z + X[2].z + X[3].z) / 4.0};
  if (dot(n, sub(cen, tc)) < 0.0) n = scale(n, -1.0);  // outward: away from the parent tet
  V3 g[4];
  double vol;
  tet_gradients(X, g, vol);
  // barycentric coordinates of the centroid: lambda_i = 1/4-type affine functions, lambda_i(p) = g_i . (p - X_3) (+1 for i=3 sum rule)
  const V3 d = sub(cen, X[3]);
  double lam[4];
  lam[0] = dot(g[0], d);
  lam[1] = dot(g[1], d);
  lam[2] = dot(g[2], d);
  lam[3] = 1.0 - lam[0] - lam[1] - lam[2];
  const unsigned sg = tet_sign[t];
  constexpr int PA[6] = {0, 0, 0, 1, 1, 2}, PB[6] = {1, 2, 3, 2, 3, 3};
  C3 E = {cmake(0, 0), cmake(0, 0), cmake(0, 0)}, curl = E;
#pragma unroll
  for (int e = 0; e < 6; ++e) {
    const int a = PA[e], b = PB[e];
    c128 cf = x[tri_tet_edges[6 * i + e]];
    if ((sg >> e) & 1u) cf = cneg(cf);
    const V3 W = {lam[a] * g[b].x - lam[b] * g[a].x, lam[a] * g[b].y - lam[b] * g[a].y, lam[a] * g[b].z - lam[b] * g[a].z};
    const V3 cu = scale(cross(g[a], g[b]), 2.0);
    E.x = cadd(E.x, cscale(W.x, cf)); E.y = cadd(E.y, cscale(W.y, cf)); E.z = cadd(E.z, cscale(W.z, cf));
    curl.x = cadd(curl.x, cscale(cu.x, cf)); curl.y = cadd(curl.y, cscale(cu.y, cf)); curl.z = cadd(curl.z, cscale(cu.z, cf));
  }
  C3 H = {cmul(curl.x, inv_jwmu), cmul(curl.y, inv_jwmu), cmul(curl.z, inv_jwmu)};
  // X_tan = X - (X.dot(n)) n with Eigen's dot, which CONJUGATES its left operand (huygens_surface.cpp:123-126):
  // the subtracted amplitude is conj(X . n)
  const c128 en = cconj(cdotr(E, n)), hn = cconj(cdotr(H, n));
  E.x = csub(E.x, cscale(n.x, en)); E.y = csub(E.y, cscale(n.y, en)); E.z = csub(E.z, cscale(n.z, en));
  H.x = csub(H.x, cscale(n.x, hn)); H.y = csub(H.y, cscale(n.y, hn)); H.z = csub(H.z, cscale(n.z, hn));
  r_out[3 * i] = cen.x; r_out[3 * i + 1] = cen.y; r_out[3 * i + 2] = cen.z;
  n_out[3 * i] = n.x; n_out[3 * i + 1] = n.y; n_out[3 * i + 2] = n.z;
  E_out[3 * i] = E.x; E_out[3 * i + 1] = E.y; E_out[3 * i + 2] = E.z;
  H_out[3 * i] = H.x; H_out[3 * i + 1] = H.y; H_out[3 * i + 2] = H.z;
  area_out[i] = area;
}

constexpr double Z0_FREE = 376.730313668;  // literal at ntf.cpp:11
constexpr int SC_THREADS = 256;

// grid = directions; CTA sums the surface samples in a fixed order
__global__ void __launch_bounds__(SC_THREADS)
k_stratton_chu(int n_s, const double *__restrict__ r, const double *__restrict__ n, const c128 *__restrict__ E, const c128 *__restrict__ H,
               const double *__restrict__ area, const double *__restrict__ theta, const double *__restrict__ phi, double k0,
               c128 *__restrict__ e_theta, c128 *__restrict__ e_phi) {
  const int dir = blockIdx.x;
  const double th = theta[dir], ph = phi[dir];
  const double st = sin(th), ct = cos(th), sp = sin(ph), cp = cos(ph);
  const V3 rhat = {st * cp, st * sp, ct}, th_hat = {ct * cp, ct * sp, -st}, ph_hat = {-sp, cp, 0.0};
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int s = threadIdx.x; s < n_s; s += SC_THREADS) {
    const V3 ns = {n[3 * s], n[3 * s + 1], n[3 * s + 2]}, rs = {r[3 * s], r[3 * s + 1], r[3 * s + 2]};
    const C3 Es = {E[3 * s], E[3 * s + 1], E[3 * s + 2]}, Hs = {H[3 * s], H[3 * s + 1], H[3 * s + 2]};
    const C3 J = cross_rc(ns, Hs);
    C3 M = cross_rc(ns, Es);
    M = {cneg(M.x), cneg(M.y), cneg(M.z)};
    double sn, cs;
    sincos(-k0 * dot(rhat, rs), &sn, &cs);
    const c128 phase = cmake(cs, sn);
    const C3 t1 = cross_cr(cross_rc(rhat, J), rhat), t2 = cross_rc(rhat, M);
    const C3 term = {csub(cscale(Z0_FREE, t1.x), t2.x), csub(cscale(Z0_FREE, t1.y), t2.y), csub(cscale(Z0_FREE, t1.z), t2.z)};
    // (j k0 / 4 pi) * term * phase * area
    const c128 f = cmul(cmake(0.0, k0 / (4.0 * M_PI)), cscale(area[s], phase));
    const c128 ex = cmul(f, term.x), ey = cmul(f, term.y), ez = cmul(f, term.z);
    acc[0] += ex.x; acc[1] += ex.y; acc[2] += ey.x; acc[3] += ey.y; acc[4] += ez.x; acc[5] += ez.y;
  }
  __shared__ double red[SC_THREADS][6];
#pragma unroll
  for (int k = 0; k < 6; ++k) red[threadIdx.x][k] = acc[k];
  __syncthreads();
  for (int s = SC_THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
#pragma unroll
      for (int k = 0; k < 6; ++k) red[threadIdx.x][k] += red[threadIdx.x + s][k];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const C3 Ef = {cmake(red[0][0], red[0][1]), cmake(red[0][2], red[0][3]), cmake(red[0][4], red[0][5])};
    e_theta[dir] = cdotr(Ef, th_hat);
    e_phi[dir] = cdotr(Ef, ph_hat);
  }
}

struct Tmp {
  void *p = nullptr;
  Ctx *c = nullptr;
  int alloc(Ctx *ctx, size_t bytes) {
    c = ctx;
    unsigned char *q = nullptr;
    int rc = dev_alloc(ctx, &q, bytes);
    p = q;
    return rc;
  }
  ~Tmp() {
    if (!p) return;
    cudaStreamSynchronize(c->stream);
    dfree(p);
  }
};

}  // namespace
}  // namespace efb

using namespace efb;

extern "C" {

int efb_huygens_eval(efb_mesh *mesh_, efb_system *sys_, int32_t rhs, const double *x_host_c128, int32_t n_tri, const int32_t *tri_nodes,
                     const int32_t *tri_tet, const int32_t *tri_tet_edges, double omega, const double *mu_r_c128, double *r_out, double *n_out,
                     double *E_tan_c128, double *H_tan_c128, double *area_out) {
  Mesh *M = (Mesh *)mesh_;
  System *S = (System *)sys_;
  if (!M) return fail(nullptr, EFB_ERR_INVALID, "efb_huygens_eval: NULL mesh");
  Ctx *c = M->ctx;
  if ((S == nullptr) == (x_host_c128 == nullptr)) return fail(c, EFB_ERR_INVALID, "efb_huygens_eval: pass either a system (+rhs) or a host solution vector");
  if (S) {
    EFB_WHOLE_ONLY(S, "efb_huygens_eval");
    if (S->mesh != M || rhs < 0 || rhs >= S->n_sys) return fail(c, EFB_ERR_INVALID, "efb_huygens_eval: system does not belong to the mesh / bad rhs");
  }
  if (n_tri <= 0 || !tri_nodes || !tri_tet || !tri_tet_edges || !mu_r_c128 || !r_out || !n_out || !E_tan_c128 ||
      !H_tan_c128 || !area_out || omega == 0.0)
    return fail(c, EFB_ERR_INVALID, "efb_huygens_eval: bad arguments");
  for (int i = 0; i < n_tri; ++i) {
    if (tri_tet[i] < 0 || tri_tet[i] >= M->n_tet) return fail(c, EFB_ERR_INVALID, "efb_huygens_eval: tri_tet[%d] out of range", i);
    for (int k = 0; k < 3; ++k)
      if (tri_nodes[3 * i + k] < 0 || tri_nodes[3 * i + k] >= M->n_node) return fail(c, EFB_ERR_INVALID, "efb_huygens_eval: tri_nodes out of range");
    for (int k = 0; k < 6; ++k)
      if (tri_tet_edges[6 * i + k] < 0 || tri_tet_edges[6 * i + k] >= M->m) return fail(c, EFB_ERR_INVALID, "efb_huygens_eval: edge id out of range");
  }
  EFB_CUDA(c, cudaSetDevice(c->device));
  constexpr double mu0 = 4.0e-7 * M_PI;
  const c128 jwmu = cmul(cmake(0.0, omega * mu0), cmake(mu_r_c128[0], mu_r_c128[1]));
  const c128 inv = cdiv(cmake(1.0, 0.0), jwmu);
  Tmp tn, tt, te, o_r, o_n, o_E, o_H, o_a, xs;
  int rc;
  const c128 *d_x = S ? S->d_x + (size_t)rhs * S->m : nullptr;
  if (!S) {
    if ((rc = xs.alloc(c, (size_t)M->m * 16))) return rc;
    EFB_CUDA(c, cudaMemcpyAsync(xs.p, x_host_c128, (size_t)M->m * 16, cudaMemcpyHostToDevice, c->stream));
    d_x = (const c128 *)xs.p;
  }
  if ((rc = tn.alloc(c, (size_t)n_tri * 12)) || (rc = tt.alloc(c, (size_t)n_tri * 4)) || (rc = te.alloc(c, (size_t)n_tri * 24)) ||
      (rc = o_r.alloc(c, (size_t)n_tri * 24)) || (rc = o_n.alloc(c, (size_t)n_tri * 24)) || (rc = o_E.alloc(c, (size_t)n_tri * 48)) ||
      (rc = o_H.alloc(c, (size_t)n_tri * 48)) || (rc = o_a.alloc(c, (size_t)n_tri * 8)))
    return rc;
  cudaStream_t st = c->stream;
  EFB_CUDA(c, cudaMemcpyAsync(tn.p, tri_nodes, (size_t)n_tri * 12, cudaMemcpyHostToDevice, st));
  EFB_CUDA(c, cudaMemcpyAsync(tt.p, tri_tet, (size_t)n_tri * 4, cudaMemcpyHostToDevice, st));
  EFB_CUDA(c, cudaMemcpyAsync(te.p, tri_tet_edges, (size_t)n_tri * 24, cudaMemcpyHostToDevice, st));
  {
    Timed tm(c);
    k_huygens<<<(n_tri + 127) / 128, 128, 0, st>>>(M->d_xyz, M->d_tet_nodes, M->d_tet_sign, d_x, n_tri, (const int32_t *)tn.p,
                                                  (const int32_t *)tt.p, (const int32_t *)te.p, inv, (double *)o_r.p, (double *)o_n.p, (c128 *)o_E.p,
                                                  (c128 *)o_H.p, (double *)o_a.p);
    EFB_CHECK_LAUNCH(c);
  }
  EFB_CUDA(c, cudaMemcpyAsync(r_out, o_r.p, (size_t)n_tri * 24, cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaMemcpyAsync(n_out, o_n.p, (size_t)n_tri * 24, cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaMemcpyAsync(E_tan_c128, o_E.p, (size_t)n_tri * 48, cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaMemcpyAsync(H_tan_c128, o_H.p, (size_t)n_tri * 48, cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaMemcpyAsync(area_out, o_a.p, (size_t)n_tri * 8, cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaStreamSynchronize(st));
  return EFB_OK;
}

int efb_stratton_chu(efb_ctx *ctx_, int32_t n_s, const double *r, const double *n, const double *E_c128, const double *H_c128, const double *area,
                     int32_t n_dir, const double *theta, const double *phi, double k0, double *e_theta_c128, double *e_phi_c128) {
  Ctx *c = (Ctx *)ctx_;
  if (!c || n_s < 0 || n_dir <= 0 || (n_s > 0 && (!r || !n || !E_c128 || !H_c128 || !area)) || !theta || !phi || !e_theta_c128 || !e_phi_c128)
    return fail(c, EFB_ERR_INVALID, "efb_stratton_chu: bad arguments");
  EFB_CUDA(c, cudaSetDevice(c->device));
  Tmp dr, dn, dE, dH, da, dt, dp, oe, op;
  int rc;
  const size_t ns = (size_t)std::max(n_s, 1);
  if ((rc = dr.alloc(c, ns * 24)) || (rc = dn.alloc(c, ns * 24)) || (rc = dE.alloc(c, ns * 48)) || (rc = dH.alloc(c, ns * 48)) ||
      (rc = da.alloc(c, ns * 8)) || (rc = dt.alloc(c, (size_t)n_dir * 8)) || (rc = dp.alloc(c, (size_t)n_dir * 8)) ||
      (rc = oe.alloc(c, (size_t)n_dir * 16)) || (rc = op.alloc(c, (size_t)n_dir * 16)))
    return rc;
  cudaStream_t st = c->stream;
  if (n_s > 0) {
    EFB_CUDA(c, cudaMemcpyAsync(dr.p, r, (size_t)n_s * 24, cudaMemcpyHostToDevice, st));
    EFB_CUDA(c, cudaMemcpyAsync(dn.p, n, (size_t)n_s * 24, cudaMemcpyHostToDevice, st));
    EFB_CUDA(c, cudaMemcpyAsync(dE.p, E_c128, (size_t)n_s * 48, cudaMemcpyHostToDevice, st));
    EFB_CUDA(c, cudaMemcpyAsync(dH.p, H_c128, (size_t)n_s * 48, cudaMemcpyHostToDevice, st));
    EFB_CUDA(c, cudaMemcpyAsync(da.p, area, (size_t)n_s * 8, cudaMemcpyHostToDevice, st));
  }
  EFB_CUDA(c, cudaMemcpyAsync(dt.p, theta, (size_t)n_dir * 8, cudaMemcpyHostToDevice, st));
  EFB_CUDA(c, cudaMemcpyAsync(dp.p, phi, (size_t)n_dir * 8, cudaMemcpyHostToDevice, st));
  {
    Timed tm(c);
    k_stratton_chu<<<n_dir, SC_THREADS, 0, st>>>(n_s, (const double *)dr.p, (const double *)dn.p, (const c128 *)dE.p, (const c128 *)dH.p,
                                                 (const double *)da.p, (const double *)dt.p, (const double *)dp.p, k0, (c128 *)oe.p, (c128 *)op.p);
    EFB_CHECK_LAUNCH(c);
  }
  EFB_CUDA(c, cudaMemcpyAsync(e_theta_c128, oe.p, (size_t)n_dir * 16, cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaMemcpyAsync(e_phi_c128, op.p, (size_t)n_dir * 16, cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaStreamSynchronize(st));
  return EFB_OK;
}

}  // extern "C"
