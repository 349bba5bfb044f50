// SURVEY 8f-f3 host side: Whitney field evaluation helpers, Huygens-surface extraction (adjacency on the host, field
// evaluation on the GPU) and the near-to-far-field transformation (direction x sample loop on the GPU).
#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <stdexcept>
#include <unordered_map>

#include "edgefem/edge_basis.hpp"
#include "edgefem/post/ntf.hpp"
#include "host_internal.hpp"

namespace edgefem {

namespace {
const int kPairs[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

void gradients(const std::array<Vector3d, 4> &v, Vector3d g[4], double &V) {  // src/edge_basis.cpp:14-26
  const Vector3d b0 = v[0] - v[3], b1 = v[1] - v[3], b2 = v[2] - v[3];
  const Vector3d c12 = b1.cross(b2), c20 = b2.cross(b0), c01 = b0.cross(b1);
  const double det = b0.dot(c12), inv = 1.0 / det;
  g[0] = c12 * inv;
  g[1] = c20 * inv;
  g[2] = c01 * inv;
  g[3] = -g[0] - g[1] - g[2];
  V = std::abs(det) / 6.0;
}
}  // namespace

std::array<Vector3d, 6> whitney_edge_curls(const std::array<Vector3d, 4> &v) {  // edge_basis.cpp:33-46
  Vector3d g[4];
  double V;
  gradients(v, g, V);
  std::array<Vector3d, 6> c;
  for (int i = 0; i < 6; ++i) c[i] = g[kPairs[i][0]].cross(g[kPairs[i][1]]) * 2.0;
  return c;
}

std::array<double, 4> compute_barycentric(const std::array<Vector3d, 4> &v, const Vector3d &p) {  // edge_basis.cpp:132-151
  Vector3d g[4];
  double V;
  gradients(v, g, V);  // rows of T^-1, T = [v0-v3, v1-v3, v2-v3]
  const Vector3d d = p - v[3];
  std::array<double, 4> l;
  l[0] = g[0].dot(d);
  l[1] = g[1].dot(d);
  l[2] = g[2].dot(d);
  l[3] = 1.0 - l[0] - l[1] - l[2];
  return l;
}

Vector3d compute_grad_lambda(const std::array<Vector3d, 4> &v, int i) {  // edge_basis.cpp:153-159
  Vector3d g[4];
  double V;
  gradients(v, g, V);
  return g[i];
}

std::array<std::complex<double>, 3> evaluate_edge_field(const std::array<Vector3d, 4> &vertices, const std::array<int, 6> &edge_orient,
                                                        const std::array<std::complex<double>, 6> &edge_dofs, const Vector3d &point) {
  // edge_basis.cpp:161-190
  const auto lam = compute_barycentric(vertices, point);
  Vector3d g[4];
  double V;
  gradients(vertices, g, V);
  std::array<std::complex<double>, 3> E{};
  for (int e = 0; e < 6; ++e) {
    const int a = kPairs[e][0], b = kPairs[e][1];
    const Vector3d W = g[b] * lam[a] - g[a] * lam[b];
    const std::complex<double> cf = edge_dofs[e] * (double)edge_orient[e];
    for (int k = 0; k < 3; ++k) E[k] += cf * W[k];
  }
  return E;
}

HuygensSurfaceData extract_huygens_surface(const Mesh &mesh, const VecC &solution, int surface_tag, double omega, std::complex<double> mu_r) {
  // src/post/huygens_surface.cpp:29-149.  Face -> parent tet: the reference packs the sorted node-id triple in 21-bit
  // fields of a uint64 and lets later tets overwrite earlier ones; an exact triple key keeps the "last tet wins" rule
  // without the 2^21 id limit.
  if ((size_t)solution.size() != mesh.edges.size()) throw std::runtime_error("extract_huygens_surface: solution size does not match the mesh edges");
  static const int face_nodes[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
  auto key_of = [](std::int64_t a, std::int64_t b, std::int64_t c) {
    std::array<std::int64_t, 3> k{a, b, c};
    std::sort(k.begin(), k.end());
    return k;
  };
  // only faces of the tagged triangles are looked up: index them first, then scan the tets once
  std::map<std::array<std::int64_t, 3>, int> parent;
  std::vector<int> sel;
  for (size_t i = 0; i < mesh.tris.size(); ++i)
    if (mesh.tris[i].phys == surface_tag) {
      sel.push_back((int)i);
      parent[key_of(mesh.tris[i].conn[0], mesh.tris[i].conn[1], mesh.tris[i].conn[2])] = -1;
    }
  for (size_t t = 0; t < mesh.tets.size(); ++t)
    for (int f = 0; f < 4; ++f) {
      const auto &c = mesh.tets[t].conn;
      auto it = parent.find(key_of(c[face_nodes[f][0]], c[face_nodes[f][1]], c[face_nodes[f][2]]));
      if (it != parent.end()) it->second = (int)t;  // later tets overwrite: last one wins
    }
  std::vector<int32_t> tri_nodes, tri_tet, tri_edges;
  for (int i : sel) {
    const Element &tri = mesh.tris[i];
    const int i0 = mesh.nodeIndex.at(tri.conn[0]), i1 = mesh.nodeIndex.at(tri.conn[1]), i2 = mesh.nodeIndex.at(tri.conn[2]);
    const Vector3d e1 = mesh.nodes[i1].xyz - mesh.nodes[i0].xyz, e2 = mesh.nodes[i2].xyz - mesh.nodes[i0].xyz;
    if (0.5 * e1.cross(e2).norm() < 1e-30) continue;
    const int t = parent.at(key_of(tri.conn[0], tri.conn[1], tri.conn[2]));
    if (t < 0) continue;  // orphan triangle
    tri_nodes.insert(tri_nodes.end(), {i0, i1, i2});
    tri_tet.push_back(t);
    for (int e = 0; e < 6; ++e) tri_edges.push_back(mesh.tets[t].edges[e]);
  }
  const int n = (int)tri_tet.size();
  if (n == 0) throw std::runtime_error("extract_huygens_surface: no triangles found with surface_tag = " + std::to_string(surface_tag));
  std::vector<double> r(3 * (size_t)n), nn(3 * (size_t)n), area(n);
  std::vector<std::complex<double>> E(3 * (size_t)n), H(3 * (size_t)n);
  const double mu[2] = {mu_r.real(), mu_r.imag()};
  detail::check(efb_huygens_eval(detail::device_mesh_handle(mesh), nullptr, 0, reinterpret_cast<const double *>(solution.data()), n, tri_nodes.data(),
                                 tri_tet.data(), tri_edges.data(), omega, mu, r.data(), nn.data(), reinterpret_cast<double *>(E.data()),
                                 reinterpret_cast<double *>(H.data()), area.data()),
                "efb_huygens_eval");
  HuygensSurfaceData d;
  d.r.resize(n); d.n.resize(n); d.E_tan.resize(n); d.H_tan.resize(n); d.area = area;
  for (int i = 0; i < n; ++i) {
    d.r[i] = Vector3d(r[3 * i], r[3 * i + 1], r[3 * i + 2]);
    d.n[i] = Vector3d(nn[3 * i], nn[3 * i + 1], nn[3 * i + 2]);
    d.E_tan[i] = {E[3 * i], E[3 * i + 1], E[3 * i + 2]};
    d.H_tan[i] = {H[3 * i], H[3 * i + 1], H[3 * i + 2]};
  }
  return d;
}

// ------------------------------------------------------------------ near-to-far field (src/post/ntf.cpp)
namespace {
constexpr double Z0 = 376.730313668;

void far_field(const std::vector<Vector3d> &r, const std::vector<Vector3d> &n, const std::vector<Vector3cd> &E, const std::vector<Vector3cd> &H,
               const std::vector<double> &area, const std::vector<double> &th, const std::vector<double> &ph, double k0,
               std::vector<std::complex<double>> &et, std::vector<std::complex<double>> &ep) {
  const size_t ns = r.size();
  if (n.size() != ns || E.size() != ns || H.size() != ns || area.size() != ns) throw std::invalid_argument("stratton_chu: surface arrays differ in length");
  std::vector<double> rr(3 * ns), nn(3 * ns);
  std::vector<std::complex<double>> EE(3 * ns), HH(3 * ns);
  for (size_t i = 0; i < ns; ++i)
    for (int k = 0; k < 3; ++k) {
      rr[3 * i + k] = r[i][k];
      nn[3 * i + k] = n[i][k];
      EE[3 * i + k] = E[i][k];
      HH[3 * i + k] = H[i][k];
    }
  et.assign(th.size(), 0.0);
  ep.assign(th.size(), 0.0);
  if (th.empty()) return;
  detail::check(efb_stratton_chu(detail::device_ctx(), (int32_t)ns, rr.data(), nn.data(), reinterpret_cast<const double *>(EE.data()),
                                 reinterpret_cast<const double *>(HH.data()), area.data(), (int32_t)th.size(), th.data(), ph.data(), k0,
                                 reinterpret_cast<double *>(et.data()), reinterpret_cast<double *>(ep.data())),
                "efb_stratton_chu");
}
}  // namespace

std::vector<NTFPoint2D> stratton_chu_2d(const std::vector<Vector3d> &r, const std::vector<Vector3d> &n, const std::vector<Vector3cd> &E,
                                        const std::vector<Vector3cd> &H, const std::vector<double> &area, const std::vector<double> &theta_rad,
                                        double phi_rad, double k0) {  // ntf.cpp:86-133
  std::vector<double> ph(theta_rad.size(), phi_rad);
  std::vector<std::complex<double>> et, ep;
  far_field(r, n, E, H, area, theta_rad, ph, k0, et, ep);
  std::vector<NTFPoint2D> out(theta_rad.size());
  for (size_t i = 0; i < out.size(); ++i) out[i] = {theta_rad[i] * 180.0 / M_PI, et[i], ep[i]};
  return out;
}

FFPattern3D stratton_chu_3d(const std::vector<Vector3d> &r, const std::vector<Vector3d> &n, const std::vector<Vector3cd> &E,
                            const std::vector<Vector3cd> &H, const std::vector<double> &area, const std::vector<double> &theta_rad,
                            const std::vector<double> &phi_rad, double k0) {  // ntf.cpp:135-203
  const int Nth = (int)theta_rad.size(), Nph = (int)phi_rad.size();
  FFPattern3D p;
  p.theta_grid.resize(Nth, Nph);
  p.phi_grid.resize(Nth, Nph);
  p.E_theta.resize(Nth, Nph);
  p.E_phi.resize(Nth, Nph);
  std::vector<double> th((size_t)Nth * Nph), ph((size_t)Nth * Nph);
  for (int i = 0; i < Nth; ++i)
    for (int j = 0; j < Nph; ++j) {
      p.theta_grid(i, j) = theta_rad[i];
      p.phi_grid(i, j) = phi_rad[j];
      th[(size_t)i * Nph + j] = theta_rad[i];
      ph[(size_t)i * Nph + j] = phi_rad[j];
    }
  std::vector<std::complex<double>> et, ep;
  far_field(r, n, E, H, area, th, ph, k0, et, ep);
  for (int i = 0; i < Nth; ++i)
    for (int j = 0; j < Nph; ++j) {
      p.E_theta(i, j) = et[(size_t)i * Nph + j];
      p.E_phi(i, j) = ep[(size_t)i * Nph + j];
    }
  return p;
}

MatrixXd FFPattern3D::power_pattern() const {  // ntf.cpp:54-62
  MatrixXd w(E_theta.rows(), E_theta.cols());
  for (int i = 0; i < E_theta.rows(); ++i)
    for (int j = 0; j < E_theta.cols(); ++j) w(i, j) = std::norm(E_theta(i, j)) + std::norm(E_phi(i, j));
  return w;
}

MatrixXd FFPattern3D::total_magnitude() const {  // ntf.cpp:42-52
  MatrixXd w = power_pattern();
  for (int i = 0; i < w.rows(); ++i)
    for (int j = 0; j < w.cols(); ++j) w(i, j) = std::sqrt(w(i, j));
  return w;
}

namespace {
double max_coeff(const MatrixXd &a, int *mi = nullptr, int *mj = nullptr) {  // Eigen visits column-major, first maximum wins
  double best = -std::numeric_limits<double>::infinity();
  for (int j = 0; j < a.cols(); ++j)
    for (int i = 0; i < a.rows(); ++i)
      if (a(i, j) > best) {
        best = a(i, j);
        if (mi) *mi = i;
        if (mj) *mj = j;
      }
  return best;
}
}  // namespace

MatrixXd FFPattern3D::pattern_dB() const {  // ntf.cpp:64-80
  MatrixXd pwr = power_pattern();
  const double mx = max_coeff(pwr);
  MatrixXd dB(pwr.rows(), pwr.cols());
  for (int i = 0; i < pwr.rows(); ++i)
    for (int j = 0; j < pwr.cols(); ++j) {
      if (mx < std::numeric_limits<double>::epsilon()) {
        dB(i, j) = -200.0;
      } else {
        const double ratio = pwr(i, j) / mx;
        dB(i, j) = ratio > 1e-20 ? 10.0 * std::log10(ratio) : -200.0;
      }
    }
  return dB;
}

double compute_directivity(const FFPattern3D &pattern) {  // ntf.cpp:209-262
  const MatrixXd pwr = pattern.power_pattern();
  const int Nth = pattern.theta_grid.rows(), Nph = pattern.theta_grid.cols();
  if (Nth == 0 || Nph == 0) return 0.0;
  const double U_max = max_coeff(pwr) / (2.0 * Z0);
  if (U_max < std::numeric_limits<double>::epsilon()) return 0.0;
  if (Nth < 2 || Nph < 2) return 1.0;
  const double dtheta = (pattern.theta_grid(Nth - 1, 0) - pattern.theta_grid(0, 0)) / (Nth - 1);
  const double dphi = (pattern.phi_grid(0, Nph - 1) - pattern.phi_grid(0, 0)) / (Nph - 1);
  double P = 0.0;
  for (int i = 0; i < Nth; ++i) {
    const double s = std::sin(pattern.theta_grid(i, 0)), wt = (i == 0 || i == Nth - 1) ? 0.5 : 1.0;
    for (int j = 0; j < Nph; ++j) P += pwr(i, j) / (2.0 * Z0) * s * wt * ((j == 0 || j == Nph - 1) ? 0.5 : 1.0);
  }
  P *= dtheta * dphi;
  if (P < std::numeric_limits<double>::epsilon()) return 0.0;
  return 4.0 * M_PI * U_max / P;
}

double compute_max_gain(const FFPattern3D &pattern, double efficiency) { return efficiency * compute_directivity(pattern); }

std::pair<double, double> compute_hpbw(const FFPattern3D &pattern) {  // ntf.cpp:268-318
  const MatrixXd pwr = pattern.power_pattern();
  const int Nth = pattern.theta_grid.rows(), Nph = pattern.theta_grid.cols();
  if (Nth == 0 || Nph == 0) return {0.0, 0.0};
  int mi = 0, mj = 0;
  const double mx = max_coeff(pwr, &mi, &mj);
  if (mx < std::numeric_limits<double>::epsilon()) return {0.0, 0.0};
  const double half = mx / 2.0;
  int lo = mi, hi = mi;
  while (lo > 0 && pwr(lo, mj) > half) --lo;
  while (hi < Nth - 1 && pwr(hi, mj) > half) ++hi;
  const double e_plane = (pattern.theta_grid(hi, mj) - pattern.theta_grid(lo, mj)) * 180.0 / M_PI;
  lo = hi = mj;
  while (lo > 0 && pwr(mi, lo) > half) --lo;
  while (hi < Nph - 1 && pwr(mi, hi) > half) ++hi;
  const double h_plane = (pattern.phi_grid(mi, hi) - pattern.phi_grid(mi, lo)) * std::sin(pattern.theta_grid(mi, mj)) * 180.0 / M_PI;
  return {e_plane, h_plane};
}

}  // namespace edgefem
