// Host-side set-up of the hot path's inputs: Gmsh v2 reader, global edge numbering, PEC set,
// host element matrices, wave/lumped port builders, periodic pair matching.
// Behaviour follows the reference (cited per function); the code is written for this project.
#include <algorithm>
#include <cmath>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

#include "edgefem/bc.hpp"
#include "edgefem/edge_basis.hpp"
#include "edgefem/mesh.hpp"
#include "edgefem/periodic.hpp"
#include "edgefem/ports/lumped_port.hpp"
#include "edgefem/ports/wave_port.hpp"
#include "host_internal.hpp"

namespace edgefem {

// ------------------------------------------------------------------ mesh
// Reference: src/mesh_gmsh.cpp:15-75.  Sections other than $Nodes/$Elements are skipped; the
// first tag of an element is its physical tag; element types 2 (Tri3), 1 (Line2), 4 (Tet4).
Mesh load_gmsh_v2(const std::string &path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("Failed to open mesh file: " + path);
  Mesh mesh;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line == "$Nodes") {
      long long n = 0;
      in >> n;
      mesh.nodes.reserve((size_t)n);
      for (long long i = 0; i < n; ++i) {
        std::int64_t id;
        double x, y, z;
        in >> id >> x >> y >> z;
        mesh.nodeIndex[id] = (int)mesh.nodes.size();
        mesh.nodes.push_back(Node{id, Vector3d{x, y, z}});
      }
    } else if (line == "$Elements") {
      long long cnt = 0;
      in >> cnt;
      for (long long i = 0; i < cnt; ++i) {
        std::int64_t id;
        int type = 0, ntags = 0;
        in >> id >> type >> ntags;
        Element e;
        e.id = id;
        e.type = static_cast<ElemType>(type);
        for (int t = 0; t < ntags; ++t) {
          int tag;
          in >> tag;
          if (t == 0) e.phys = tag;
        }
        if (type == 2) {
          for (int k = 0; k < 3; ++k) in >> e.conn[k];
          mesh.tris.push_back(e);
        } else if (type == 1) {
          BoundaryLine bl;
          bl.phys = e.phys;
          in >> bl.n0 >> bl.n1;
          mesh.boundary_lines.push_back(bl);
        } else if (type == 4) {
          for (int k = 0; k < 4; ++k) in >> e.conn[k];
          mesh.tets.push_back(e);
        } else {
          std::string rest;
          std::getline(in, rest);
        }
      }
    }
  }
  build_edges(mesh);
  return mesh;
}

// Reference: src/mesh_gmsh.cpp:104-146 -- first-seen numbering, tets then tris.
void build_edges(Mesh &mesh) {
  static const int tet_pairs[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
  static const int tri_pairs[3][2] = {{0, 1}, {1, 2}, {2, 0}};
  mesh.edges.clear();
  mesh.edgeIndex.clear();
  mesh.edgeIndex.reserve(mesh.tets.size() * 7 / 5 + mesh.tris.size() + 16);
  auto visit = [&](Element &el, const int (*pairs)[2], int ne) {
    for (int e = 0; e < ne; ++e) {
      const std::int64_t a = el.conn[pairs[e][0]], b = el.conn[pairs[e][1]];
      const std::uint64_t key = make_edge_key(a, b);
      auto ins = mesh.edgeIndex.emplace(key, (int)mesh.edges.size());
      if (ins.second) mesh.edges.push_back(Edge{std::min(a, b), std::max(a, b)});
      el.edges[e] = ins.first->second;
      el.edge_orient[e] = (a < b) ? 1 : -1;
    }
  };
  for (auto &t : mesh.tets) visit(t, tet_pairs, 6);
  for (auto &t : mesh.tris) visit(t, tri_pairs, 3);
}

Mesh mesh_from_arrays(const std::vector<double> &xyz, const std::vector<std::int64_t> &tet_conn, const std::vector<int> &tet_phys,
                      const std::vector<std::int64_t> &tri_conn, const std::vector<int> &tri_phys,
                      const std::vector<std::int64_t> &node_ids) {
  if (xyz.size() % 3 || tet_conn.size() % 4 || tri_conn.size() % 3 || tet_phys.size() != tet_conn.size() / 4 ||
      tri_phys.size() != tri_conn.size() / 3 || (!node_ids.empty() && node_ids.size() != xyz.size() / 3))
    throw std::invalid_argument("mesh_from_arrays: inconsistent array sizes");
  Mesh mesh;
  const size_t nn = xyz.size() / 3;
  mesh.nodes.reserve(nn);
  mesh.nodeIndex.reserve(nn);
  for (size_t i = 0; i < nn; ++i) {
    const std::int64_t id = node_ids.empty() ? (std::int64_t)i + 1 : node_ids[i];
    mesh.nodeIndex[id] = (int)i;
    mesh.nodes.push_back(Node{id, Vector3d{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}});
  }
  mesh.tets.resize(tet_phys.size());
  for (size_t t = 0; t < tet_phys.size(); ++t) {
    Element &e = mesh.tets[t];
    e.id = (std::int64_t)t + 1;
    e.type = ElemType::Tet4;
    e.phys = tet_phys[t];
    for (int k = 0; k < 4; ++k) e.conn[k] = tet_conn[4 * t + k];
  }
  mesh.tris.resize(tri_phys.size());
  for (size_t t = 0; t < tri_phys.size(); ++t) {
    Element &e = mesh.tris[t];
    e.id = (std::int64_t)(tet_phys.size() + t) + 1;
    e.type = ElemType::Tri3;
    e.phys = tri_phys[t];
    for (int k = 0; k < 3; ++k) e.conn[k] = tri_conn[3 * t + k];
  }
  for (const auto &e : mesh.tets)
    for (int k = 0; k < 4; ++k)
      if (!mesh.nodeIndex.count(e.conn[k])) throw std::invalid_argument("mesh_from_arrays: tet references an unknown node id");
  for (const auto &e : mesh.tris)
    for (int k = 0; k < 3; ++k)
      if (!mesh.nodeIndex.count(e.conn[k])) throw std::invalid_argument("mesh_from_arrays: tri references an unknown node id");
  build_edges(mesh);
  return mesh;
}

// ------------------------------------------------------------------ bc  (src/bc.cpp:47-109)
PhysicalTagInfo list_physical_tags(const Mesh &mesh) {
  PhysicalTagInfo info;
  for (const auto &t : mesh.tets) info.volume_tags.insert(t.phys);
  for (const auto &t : mesh.tris) info.surface_tags.insert(t.phys);
  return info;
}

bool has_surface_tag(const Mesh &mesh, int tag) {
  for (const auto &t : mesh.tris)
    if (t.phys == tag) return true;
  return false;
}

bool has_volume_tag(const Mesh &mesh, int tag) {
  for (const auto &t : mesh.tets)
    if (t.phys == tag) return true;
  return false;
}

BC build_edge_pec(const Mesh &mesh, int pec_tag) {
  BC bc;
  for (const auto &tri : mesh.tris)
    if (tri.phys == pec_tag)
      for (int k = 0; k < 3; ++k) bc.dirichlet_edges.insert(tri.edges[k]);
  if (bc.dirichlet_edges.empty()) {
    std::ostringstream oss;
    oss << "WARNING: build_edge_pec() found no edges for PEC tag " << pec_tag << ".\n  Available surface tags: ";
    auto tags = list_physical_tags(mesh);
    if (tags.surface_tags.empty()) oss << "(none)";
    bool first = true;
    for (int t : tags.surface_tags) {
      oss << (first ? "" : ", ") << t;
      first = false;
    }
    std::cerr << oss.str() << "\n";
  }
  return bc;
}

// ------------------------------------------------------------------ host element matrices
namespace {
const int kTetPairs[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
const int kTriPairs[3][2] = {{0, 1}, {1, 2}, {2, 0}};

// src/edge_basis.cpp:14-26: gradients = columns of B^-T (cofactors / det), V = |det|/6
void tet_gradients(const std::array<Vector3d, 4> &v, Vector3d g[4], double &V) {
  const Vector3d b0 = v[0] - v[3], b1 = v[1] - v[3], b2 = v[2] - v[3];
  const Vector3d c12 = b1.cross(b2), c20 = b2.cross(b0), c01 = b0.cross(b1);
  const double det = b0.dot(c12);
  const double inv = 1.0 / det;
  g[0] = c12 * inv;
  g[1] = c20 * inv;
  g[2] = c01 * inv;
  g[3] = -g[0] - g[1] - g[2];
  V = std::abs(det) / 6.0;
}
} // namespace

Matrix6d whitney_curl_curl_matrix(const std::array<Vector3d, 4> &v) {  // src/edge_basis.cpp:48-63
  Vector3d g[4];
  double V;
  tet_gradients(v, g, V);
  Vector3d c[6];
  for (int i = 0; i < 6; ++i) c[i] = 2.0 * g[kTetPairs[i][0]].cross(g[kTetPairs[i][1]]);
  Matrix6d K{};
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) K[i][j] = V * c[i].dot(c[j]);
  return K;
}

Matrix6d whitney_mass_matrix(const std::array<Vector3d, 4> &v) {  // src/edge_basis.cpp:66-86
  Vector3d g[4];
  double V;
  tet_gradients(v, g, V);
  auto I = [V](int i, int j) { return i == j ? V / 10.0 : V / 20.0; };
  Matrix6d M{};
  for (int i = 0; i < 6; ++i) {
    const int a = kTetPairs[i][0], b = kTetPairs[i][1];
    for (int j = 0; j < 6; ++j) {
      const int c = kTetPairs[j][0], d = kTetPairs[j][1];
      double t = 0.0;
      t += g[b].dot(g[d]) * I(a, c);
      t -= g[b].dot(g[c]) * I(a, d);
      t -= g[a].dot(g[d]) * I(b, c);
      t += g[a].dot(g[c]) * I(b, d);
      M[i][j] = t;
    }
  }
  return M;
}

namespace {
struct TriGeom {
  bool ok = false;
  Vector3d n_hat, g[3];
  double area = 0.0, area2 = 0.0;
};
// in-plane barycentric gradients n x edge / 2A (src/edge_basis.cpp:96-110, src/ports/lumped_port.cpp:41-62)
TriGeom tri_geometry(const std::array<Vector3d, 3> &v) {
  TriGeom t;
  const Vector3d nrm = (v[1] - v[0]).cross(v[2] - v[0]);
  t.area2 = nrm.norm();
  if (t.area2 < 1e-30) return t;
  t.ok = true;
  t.n_hat = nrm / t.area2;
  t.area = t.area2 / 2.0;
  t.g[0] = t.n_hat.cross(v[2] - v[1]) / t.area2;
  t.g[1] = t.n_hat.cross(v[0] - v[2]) / t.area2;
  t.g[2] = t.n_hat.cross(v[1] - v[0]) / t.area2;
  return t;
}
std::array<Vector3d, 3> tri_vertices(const Mesh &mesh, const Element &tri) {
  std::array<Vector3d, 3> v;
  for (int k = 0; k < 3; ++k) v[k] = mesh.nodes[mesh.nodeIndex.at(tri.conn[k])].xyz;
  return v;
}
} // namespace

Matrix3d triangle_whitney_mass_matrix(const std::array<Vector3d, 3> &v) {  // src/edge_basis.cpp:89-130
  Matrix3d M{};
  const TriGeom t = tri_geometry(v);
  if (!t.ok) return M;
  auto I = [&t](int i, int j) { return i == j ? t.area / 6.0 : t.area / 12.0; };
  for (int i = 0; i < 3; ++i) {
    const int a = kTriPairs[i][0], b = kTriPairs[i][1];
    for (int j = 0; j < 3; ++j) {
      const int c = kTriPairs[j][0], d = kTriPairs[j][1];
      double s = 0.0;
      s += t.g[b].dot(t.g[d]) * I(a, c);
      s -= t.g[b].dot(t.g[c]) * I(a, d);
      s -= t.g[a].dot(t.g[d]) * I(b, c);
      s += t.g[a].dot(t.g[c]) * I(b, d);
      M[i][j] = s;
    }
  }
  return M;
}

// ------------------------------------------------------------------ dense symmetric eigen-solver
namespace detail {
// Generalised symmetric-definite problem K v = lambda M v (dense, row-major n x n):
// Cholesky M = L L^T, C = L^-1 K L^-T, cyclic Jacobi on C, v = L^-T y (so v^T M v = 1).
// Eigenvalues ascending.  Returns false if M is not positive definite.
bool sym_gen_eig(std::vector<double> K, std::vector<double> M, int n, std::vector<double> &evals, std::vector<double> &evecs) {
  auto at = [n](std::vector<double> &A, int i, int j) -> double & { return A[(size_t)i * n + j]; };
  // Cholesky (lower) in place of M
  for (int j = 0; j < n; ++j) {
    double d = at(M, j, j);
    for (int k = 0; k < j; ++k) d -= at(M, j, k) * at(M, j, k);
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    at(M, j, j) = d;
    for (int i = j + 1; i < n; ++i) {
      double s = at(M, i, j);
      for (int k = 0; k < j; ++k) s -= at(M, i, k) * at(M, j, k);
      at(M, i, j) = s / d;
    }
  }
  // C = L^-1 K L^-T : forward-substitute columns, then rows
  for (int c = 0; c < n; ++c)  // K <- L^-1 K
    for (int i = 0; i < n; ++i) {
      double s = at(K, i, c);
      for (int k = 0; k < i; ++k) s -= at(M, i, k) * at(K, k, c);
      at(K, i, c) = s / at(M, i, i);
    }
  for (int r = 0; r < n; ++r)  // K <- K L^-T  (solve X L^T = K row-wise)
    for (int j = 0; j < n; ++j) {
      double s = at(K, r, j);
      for (int k = 0; k < j; ++k) s -= at(K, r, k) * at(M, j, k);
      at(K, r, j) = s / at(M, j, j);
    }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      const double s = 0.5 * (at(K, i, j) + at(K, j, i));
      at(K, i, j) = s;
      at(K, j, i) = s;
    }
  // cyclic Jacobi
  std::vector<double> Y((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) at(Y, i, i) = 1.0;
  double fro = 0.0;
  for (double x : K) fro += x * x;
  fro = std::sqrt(fro);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) off += at(K, i, j) * at(K, i, j);
    if (std::sqrt(2.0 * off) <= 1e-15 * fro) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = at(K, p, q);
        if (apq == 0.0) continue;
        const double app = at(K, p, p), aqq = at(K, q, q);
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {  // columns p,q
          const double kp = at(K, k, p), kq = at(K, k, q);
          at(K, k, p) = c * kp - s * kq;
          at(K, k, q) = s * kp + c * kq;
        }
        for (int k = 0; k < n; ++k) {  // rows p,q
          const double pk = at(K, p, k), qk = at(K, q, k);
          at(K, p, k) = c * pk - s * qk;
          at(K, q, k) = s * pk + c * qk;
        }
        for (int k = 0; k < n; ++k) {
          const double yp = at(Y, k, p), yq = at(Y, k, q);
          at(Y, k, p) = c * yp - s * yq;
          at(Y, k, q) = s * yp + c * yq;
        }
      }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return at(K, a, a) < at(K, b, b); });
  evals.resize(n);
  evecs.assign((size_t)n * n, 0.0);  // column j = eigenvector j, stored row-major [i*n + j]
  for (int j = 0; j < n; ++j) {
    const int src = order[j];
    evals[j] = at(K, src, src);
    // back-substitute L^T v = y
    std::vector<double> y(n);
    for (int i = 0; i < n; ++i) y[i] = at(Y, i, src);
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      for (int k = i + 1; k < n; ++k) s -= at(M, k, i) * y[k];
      y[i] = s / at(M, i, i);
    }
    for (int i = 0; i < n; ++i) evecs[(size_t)i * n + j] = y[i];
  }
  return true;
}
} // namespace detail

// ------------------------------------------------------------------ wave ports
PortMode solve_te10_mode(const RectWaveguidePort &port, double freq) {  // src/ports/port_eigensolve.cpp:45-63
  constexpr double c0 = 299792458.0, eta0 = 376.730313668;  // literal at port_eigensolve.cpp:15
  constexpr double mu0 = 4.0 * M_PI * 1e-7;
  constexpr double eps0 = 1.0 / (mu0 * c0 * c0);
  PortMode mode{};
  mode.pol = ModePolarization::TE;
  mode.fc = c0 / (2.0 * port.a);
  const double k = 2.0 * M_PI * freq / c0, kc = M_PI / port.a;
  mode.kc = kc;
  mode.omega = 2.0 * M_PI * freq;
  mode.mu = mu0;
  mode.eps = eps0;
  const double beta = std::sqrt(std::max(0.0, k * k - kc * kc));
  mode.beta = beta;
  mode.Z0 = beta > 0 ? eta0 * k / beta : 0.0;
  return mode;
}

VectorXd solve_port_mode_2d(const Mesh &mesh, int surface_tag, const std::unordered_set<int> &pec_edges, double target_kc_sq,
                            double &kc_sq_out) {  // src/ports/wave_port.cpp:410-514
  const int m = (int)mesh.edges.size();
  kc_sq_out = 0.0;
  std::vector<int> port_edges;
  std::unordered_map<int, int> local;
  for (const auto &tri : mesh.tris) {
    if (tri.phys != surface_tag) continue;
    for (int le = 0; le < 3; ++le) {
      const int ge = tri.edges[le];
      if (pec_edges.count(ge) || local.count(ge)) continue;
      local[ge] = (int)port_edges.size();
      port_edges.push_back(ge);
    }
  }
  const int n = (int)port_edges.size();
  VectorXd v_full = VectorXd::Zero(m);
  if (n == 0) return v_full;
  std::vector<double> K((size_t)n * n, 0.0), M((size_t)n * n, 0.0);
  for (const auto &tri : mesh.tris) {
    if (tri.phys != surface_tag) continue;
    const auto v = tri_vertices(mesh, tri);
    const TriGeom t = tri_geometry(v);
    if (!t.ok) continue;
    const Matrix3d Ml = triangle_whitney_mass_matrix(v);
    double curl[3];
    for (int le = 0; le < 3; ++le) curl[le] = 2.0 * t.g[kTriPairs[le][0]].cross(t.g[kTriPairs[le][1]]).dot(t.n_hat);
    for (int i = 0; i < 3; ++i) {
      auto ii = local.find(tri.edges[i]);
      if (ii == local.end()) continue;
      for (int j = 0; j < 3; ++j) {
        auto jj = local.find(tri.edges[j]);
        if (jj == local.end()) continue;
        const double sign = tri.edge_orient[i] * tri.edge_orient[j];
        K[(size_t)ii->second * n + jj->second] += sign * curl[i] * curl[j] * t.area;
        M[(size_t)ii->second * n + jj->second] += sign * Ml[i][j];
      }
    }
  }
  std::vector<double> ev, V;
  if (!detail::sym_gen_eig(K, M, n, ev, V)) return v_full;
  const double kc_min = 1e-6 * std::max(target_kc_sq, 1.0);
  int best = -1;
  double best_dist = std::numeric_limits<double>::max();
  for (int i = 0; i < n; ++i) {
    if (ev[i] < kc_min) continue;
    const double d = std::abs(ev[i] - target_kc_sq);
    if (d < best_dist) {
      best_dist = d;
      best = i;
    }
  }
  if (best < 0) return v_full;
  kc_sq_out = ev[best];
  // The eigenvector sign is solver-defined in the reference (Eigen GSAES).  Fixed here so results
  // are reproducible: the component of largest magnitude is made positive (first such on ties).
  int imax = 0;
  for (int i = 1; i < n; ++i)
    if (std::abs(V[(size_t)i * n + best]) > std::abs(V[(size_t)imax * n + best]) * (1.0 + 1e-12)) imax = i;
  const double sgn = V[(size_t)imax * n + best] < 0 ? -1.0 : 1.0;
  for (int i = 0; i < n; ++i) v_full[port_edges[i]] = sgn * V[(size_t)i * n + best];
  return v_full;
}

WavePort build_wave_port_2d(const Mesh &mesh, int surface_tag, const PortMode &mode, const std::unordered_set<int> &pec_edges,
                            double target_kc_sq) {  // src/ports/wave_port.cpp:516-546
  WavePort port;
  port.surface_tag = surface_tag;
  port.mode = mode;
  double kc_sq = 0.0;
  const VectorXd v = solve_port_mode_2d(mesh, surface_tag, pec_edges, target_kc_sq, kc_sq);
  if (kc_sq > 0.0) port.mode.kc = std::sqrt(kc_sq);
  std::set<int> edge_set;
  for (const auto &tri : mesh.tris)
    if (tri.phys == surface_tag)
      for (int le = 0; le < 3; ++le) edge_set.insert(tri.edges[le]);
  port.edges.assign(edge_set.begin(), edge_set.end());
  port.weights = VectorXcd::Zero(port.edges.size());
  for (size_t k = 0; k < port.edges.size(); ++k) port.weights[k] = v[port.edges[k]];
  return port;
}

SparseMatrix<double> assemble_port_surface_mass(const Mesh &mesh, int surface_tag,
                                                const std::unordered_set<int> &dirichlet_edges) {  // wave_port.cpp:549-585
  const int m = (int)mesh.edges.size();
  std::map<std::pair<int, int>, double> acc;
  for (const auto &tri : mesh.tris) {
    if (tri.phys != surface_tag) continue;
    const Matrix3d Ml = triangle_whitney_mass_matrix(tri_vertices(mesh, tri));
    for (int i = 0; i < 3; ++i) {
      const int gi = tri.edges[i];
      if (dirichlet_edges.count(gi)) continue;
      for (int j = 0; j < 3; ++j) {
        const int gj = tri.edges[j];
        if (dirichlet_edges.count(gj)) continue;
        acc[{gi, gj}] += tri.edge_orient[i] * tri.edge_orient[j] * Ml[i][j];
      }
    }
  }
  SparseMatrix<double> Ms(m, m);
  auto &rp = Ms.rowptr();
  auto &ci = Ms.colidx();
  auto &va = Ms.values();
  for (const auto &kv : acc) rp[kv.first.first + 1]++;
  for (int r = 0; r < m; ++r) rp[r + 1] += rp[r];
  ci.reserve(acc.size());
  va.reserve(acc.size());
  for (const auto &kv : acc) {  // std::map iterates (row, col) ascending => CSR order
    ci.push_back(kv.first.second);
    va.push_back(kv.second);
  }
  return Ms;
}

// ------------------------------------------------------------------ lumped port (src/ports/lumped_port.cpp:21-160)
WavePort build_lumped_port(const Mesh &mesh, const LumpedPortConfig &config) {
  std::set<int> edge_set;  // sorted: deterministic edge order (the reference's order is unordered_set-defined)
  for (const auto &tri : mesh.tris)
    if (tri.phys == config.surface_tag)
      for (int e = 0; e < 3; ++e) edge_set.insert(tri.edges[e]);
  if (edge_set.empty())
    throw std::runtime_error("build_lumped_port: no triangles found with surface_tag = " + std::to_string(config.surface_tag));
  std::vector<int> edges(edge_set.begin(), edge_set.end());
  std::unordered_map<int, size_t> idx;
  for (size_t i = 0; i < edges.size(); ++i) idx[edges[i]] = i;
  const Vector3d e_dir = config.e_direction.normalized();
  VectorXcd w = VectorXcd::Zero(edges.size());
  if (config.weight_mode == LumpedPortWeightMode::SurfaceIntegral) {
    for (const auto &tri : mesh.tris) {
      if (tri.phys != config.surface_tag) continue;
      const TriGeom t = tri_geometry(tri_vertices(mesh, tri));
      if (!t.ok) continue;
      for (int le = 0; le < 3; ++le) {
        const int li = kTriPairs[le][0], lj = kTriPairs[le][1];
        const double integral = (t.area / 3.0) * (t.g[lj] - t.g[li]).dot(e_dir);
        w[idx.at(tri.edges[le])] += cplx(tri.edge_orient[le] * integral, 0.0);
      }
    }
  } else {
    for (size_t i = 0; i < edges.size(); ++i) {
      const Edge &ed = mesh.edges[edges[i]];
      const Vector3d p0 = mesh.nodes[mesh.nodeIndex.at(ed.n0)].xyz, p1 = mesh.nodes[mesh.nodeIndex.at(ed.n1)].xyz;
      w[i] = cplx((p1 - p0).dot(e_dir), 0.0);
    }
  }
  const double nrm = w.norm();
  if (nrm > 1e-15) w *= std::sqrt(config.z0) / nrm;
  WavePort port;
  port.surface_tag = config.surface_tag;
  port.edges = edges;
  port.weights = w;
  port.mode.Z0 = config.z0;
  port.mode.beta = 0.0;
  port.mode.kc = 0.0;
  port.mode.fc = 0.0;
  return port;
}

// ------------------------------------------------------------------ periodic (src/periodic.cpp:48-225)
namespace {
Vector3d edge_centroid(const Mesh &mesh, int e) {
  const Edge &ed = mesh.edges[e];
  return 0.5 * (mesh.nodes[mesh.nodeIndex.at(ed.n0)].xyz + mesh.nodes[mesh.nodeIndex.at(ed.n1)].xyz);
}
std::vector<int> surface_edges_sorted(const Mesh &mesh, int tag) {
  std::set<int> s;
  for (const auto &tri : mesh.tris)
    if (tri.phys == tag)
      for (int i = 0; i < 3; ++i) s.insert(tri.edges[i]);
  return std::vector<int>(s.begin(), s.end());
}
// orientation recorded by the LAST tagged triangle containing the edge (periodic.cpp:135-146)
std::unordered_map<int, int> orient_table(const Mesh &mesh, int tag) {
  std::unordered_map<int, int> tab;
  for (const auto &tri : mesh.tris) {
    if (tri.phys != tag) continue;
    for (int i = 0; i < 3; ++i) {
      bool first_in_tri = true;
      for (int k = 0; k < i; ++k) first_in_tri &= (tri.edges[k] != tri.edges[i]);
      if (first_in_tri) tab[tri.edges[i]] = tri.edge_orient[i];
    }
  }
  return tab;
}
} // namespace

PeriodicBC build_periodic_pairs(const Mesh &mesh, int master_tag, int slave_tag, const Vector3d &period_vector, double tolerance) {
  PeriodicBC pbc;
  pbc.period_vector = period_vector;
  pbc.phase_shift = {1.0, 0.0};
  const std::vector<int> masters = surface_edges_sorted(mesh, master_tag), slaves = surface_edges_sorted(mesh, slave_tag);
  if (masters.empty()) throw std::runtime_error("No edges found on master surface with tag " + std::to_string(master_tag));
  if (slaves.empty()) throw std::runtime_error("No edges found on slave surface with tag " + std::to_string(slave_tag));
  // spatial grid over slave centroids (cell = tolerance-independent bucket of the bbox), exact distance test
  std::vector<Vector3d> sc(slaves.size());
  for (size_t i = 0; i < slaves.size(); ++i) sc[i] = edge_centroid(mesh, slaves[i]);
  std::vector<size_t> by_x(slaves.size());
  for (size_t i = 0; i < by_x.size(); ++i) by_x[i] = i;
  std::sort(by_x.begin(), by_x.end(), [&](size_t a, size_t b) { return sc[a].x() < sc[b].x(); });
  const auto mo = orient_table(mesh, master_tag), so = orient_table(mesh, slave_tag);
  std::unordered_set<int> matched;
  for (int me : masters) {
    const Vector3d want = edge_centroid(mesh, me) + period_vector;
    // candidates with |x - want.x| < tolerance via binary search on the x-sorted list
    size_t lo = std::lower_bound(by_x.begin(), by_x.end(), want.x() - tolerance,
                                 [&](size_t a, double x) { return sc[a].x() < x; }) - by_x.begin();
    int found = -1;
    for (size_t k = lo; k < by_x.size() && sc[by_x[k]].x() <= want.x() + tolerance; ++k)
      if ((sc[by_x[k]] - want).norm() < tolerance) {
        found = slaves[by_x[k]];
        break;
      }
    if (found < 0) throw std::runtime_error("Could not find matching slave edge for master edge " + std::to_string(me));
    if (matched.count(found)) throw std::runtime_error("Slave edge " + std::to_string(found) + " matched to multiple master edges");
    matched.insert(found);
    PeriodicPair pr;
    pr.master_edge = me;
    pr.slave_edge = found;
    auto im = mo.find(me);
    auto is = so.find(found);
    pr.master_orient = (im != mo.end() && im->second != 0) ? im->second : 1;
    pr.slave_orient = (is != so.end() && is->second != 0) ? is->second : 1;
    pr.translation = period_vector;
    pbc.pairs.push_back(pr);
  }
  return pbc;
}

bool validate_periodic_bc(const Mesh &mesh, const PeriodicBC &pbc) {
  if (pbc.pairs.empty()) return false;
  const int ne = (int)mesh.edges.size();
  std::unordered_set<int> ms, ss;
  for (const auto &p : pbc.pairs) {
    if (p.master_edge < 0 || p.master_edge >= ne || p.slave_edge < 0 || p.slave_edge >= ne) return false;
    if ((p.master_orient != 1 && p.master_orient != -1) || (p.slave_orient != 1 && p.slave_orient != -1)) return false;
    if (!ms.insert(p.master_edge).second || !ss.insert(p.slave_edge).second) return false;
  }
  return true;
}

void set_floquet_phase(PeriodicBC &pbc, const Vector2d &k) {
  const double arg = k.x() * pbc.period_vector.x() + k.y() * pbc.period_vector.y();
  pbc.phase_shift = {std::cos(arg), std::sin(arg)};
}

std::complex<double> floquet_phase_from_angle(const Vector3d &L, double theta, double phi, double k0) {
  const Vector3d k(k0 * std::sin(theta) * std::cos(phi), k0 * std::sin(theta) * std::sin(phi), -k0 * std::cos(theta));
  const double arg = k.dot(L);
  return {std::cos(arg), std::sin(arg)};
}

int count_surface_edges(const Mesh &mesh, int surface_tag) { return (int)surface_edges_sorted(mesh, surface_tag).size(); }

} // namespace edgefem
