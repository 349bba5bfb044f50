// SURVEY 8f rows f1 / f4 on the host: nodal 2-D port eigenmodes, port-face extraction, modal line-integral weights,
// and the Touchstone writers.  These run once per sweep (ports) or once per result (files); the per-frequency path
// stays on the device.  Each function cites the reference lines it restates.
#include <algorithm>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>

#include "edgefem/io/touchstone.hpp"
#include "edgefem/ports/wave_port.hpp"
#include "host_internal.hpp"

namespace edgefem {

namespace {
constexpr double c0 = 299792458.0;
constexpr double mu0 = 4.0 * M_PI * 1e-7;
constexpr double eps0 = 1.0 / (mu0 * c0 * c0);

// the reference's ensure_sqrt is the principal root (src/ports/port_eigensolve.cpp:19-21)
cplx ensure_sqrt(cplx v) { return std::sqrt(v); }

// P1 shape-function gradients of a triangle in the xy-plane: rows = nodes, columns = d/dx, d/dy; also |area|
void tri_shape(const Mesh &mesh, const Element &tri, double G[3][2], double &area) {
  double x[3], y[3];
  for (int k = 0; k < 3; ++k) {
    const auto &n = mesh.nodes.at(mesh.nodeIndex.at(tri.conn[k]));
    x[k] = n.xyz.x();
    y[k] = n.xyz.y();
  }
  const double det = (x[1] - x[0]) * (y[2] - y[0]) - (x[2] - x[0]) * (y[1] - y[0]);
  area = 0.5 * std::abs(det);
  // inverse of [[1,x0,y0],[1,x1,y1],[1,x2,y2]], columns 1..2
  G[0][0] = (y[1] - y[2]) / det; G[1][0] = (y[2] - y[0]) / det; G[2][0] = (y[0] - y[1]) / det;
  G[0][1] = (x[2] - x[1]) / det; G[1][1] = (x[0] - x[2]) / det; G[2][1] = (x[1] - x[0]) / det;
}
}  // namespace

SParams2 straight_waveguide_sparams(const RectWaveguidePort &port, double length, double freq) {  // port_eigensolve.cpp:67-88
  const double k = 2.0 * M_PI * freq / c0, kc = M_PI / port.a;
  SParams2 s{};
  if (k <= kc) {
    s.s11 = 1.0; s.s22 = 1.0; s.s21 = 0.0; s.s12 = 0.0;
  } else {
    const double beta = std::sqrt(k * k - kc * kc);
    const cplx phase = std::exp(cplx(0.0, -beta * length));
    s.s11 = 0.0; s.s22 = 0.0; s.s21 = phase; s.s12 = phase;
  }
  return s;
}

std::vector<PortMode> solve_port_eigens(const Mesh &mesh, int num_modes, double omega, cplx eps_r, cplx mu_r,
                                        ModePolarization pol) {  // port_eigensolve.cpp:97-275
  if (mesh.tris.empty()) throw std::runtime_error("Port eigensolver requires a 2D mesh.");
  if (omega == 0.0) throw std::runtime_error("Port eigensolver requires non-zero frequency.");
  const int nn = (int)mesh.nodes.size();
  std::vector<char> pec(nn, 0);
  if (pol == ModePolarization::TM)
    for (const auto &b : mesh.boundary_lines)
      if (b.phys == 1) {
        pec.at(mesh.nodeIndex.at(b.n0)) = 1;
        pec.at(mesh.nodeIndex.at(b.n1)) = 1;
      }
  std::vector<int> dof(nn, -1);
  int nd = 0;
  for (int i = 0; i < nn; ++i)
    if (!pec[i]) dof[i] = nd++;
  std::vector<double> A((size_t)nd * nd, 0.0), B((size_t)nd * nd, 0.0);
  for (const auto &tri : mesh.tris) {
    double G[3][2], area;
    tri_shape(mesh, tri, G, area);
    for (int i = 0; i < 3; ++i) {
      const int di = dof[mesh.nodeIndex.at(tri.conn[i])];
      if (di < 0) continue;
      for (int j = 0; j < 3; ++j) {
        const int dj = dof[mesh.nodeIndex.at(tri.conn[j])];
        if (dj < 0) continue;
        A[(size_t)di * nd + dj] += area * (G[i][0] * G[j][0] + G[i][1] * G[j][1]);
        B[(size_t)di * nd + dj] += (area / 12.0) * (i == j ? 2.0 : 1.0);
      }
    }
  }
  std::vector<double> evals, evecs;
  if (nd == 0 || !detail::sym_gen_eig(A, B, nd, evals, evecs)) throw std::runtime_error("Eigenvalue computation failed.");
  const cplx eps = eps0 * eps_r, mu = mu0 * mu_r;
  const cplx k = omega * ensure_sqrt(mu * eps);
  const cplx j(0.0, 1.0);
  std::vector<PortMode> modes;
  // The reference skips kc^2 < 1e-12 (absolute).  The constant TE null mode comes out of a dense eigen-solver at
  // ~1e-16 * lambda_max (1e-11..1e-9 here), on either side of that constant; the cut is therefore taken relative to
  // the spectrum so the null mode is always skipped (documented deviation).
  double lam_max = 0.0;
  for (int i = 0; i < nd; ++i)
    if (std::isfinite(evals[i])) lam_max = std::max(lam_max, std::abs(evals[i]));
  const double null_cut = std::max(1e-12, 1e-9 * lam_max);
  for (int i = 0; i < nd; ++i) {
    const double kc2 = evals[i];
    if (!std::isfinite(kc2) || kc2 < null_cut) continue;
    const double kc = std::sqrt(kc2);
    PortMode mode;
    mode.pol = pol;
    mode.fc = kc * c0 / (2.0 * M_PI);
    mode.kc = kc;
    mode.omega = omega;
    mode.eps = eps;
    mode.mu = mu;
    const cplx beta = ensure_sqrt(k * k - kc2);
    mode.beta = beta;
    if (pol == ModePolarization::TE) mode.Z0 = (beta == cplx(0.0)) ? cplx(0.0) : (omega * mu / beta);
    else mode.Z0 = (beta == cplx(0.0)) ? cplx(0.0) : (beta / (omega * eps));
    VectorXcd field(nn, cplx(0.0));
    int imax = 0;
    for (int n = 0; n < nn; ++n)
      if (dof[n] >= 0) {
        field[n] = evecs[(size_t)dof[n] * nd + i];  // sym_gen_eig stores eigenvector i in column i (row-major n x n)
        if (std::abs(field[n]) > std::abs(field[imax])) imax = n;
      }
    if (field[imax].real() < 0.0) field *= -1.0;  // deterministic sign (the reference's is the eigen-solver's)
    cplx power = 0.0;
    for (const auto &tri : mesh.tris) {
      double G[3][2], area;
      tri_shape(mesh, tri, G, area);
      cplx gx = 0.0, gy = 0.0;
      for (int q = 0; q < 3; ++q) {
        const cplx v = field[mesh.nodeIndex.at(tri.conn[q])];
        gx += G[q][0] * v;
        gy += G[q][1] * v;
      }
      cplx Ex, Ey, Hx, Hy;
      if (pol == ModePolarization::TE) {
        const cplx fe = j * omega * mu / (kc * kc), fh = beta / (kc * kc);
        Ex = -fe * gy; Ey = fe * gx; Hx = fh * gx; Hy = fh * gy;
      } else {
        const cplx fe = -beta / (kc * kc), fh = 1.0 / (j * omega * mu);
        Ex = fe * gx; Ey = fe * gy; Hx = -fh * gy; Hy = fh * gx;
      }
      power += 0.5 * area * (Ex * std::conj(Hy) - Ey * std::conj(Hx));
    }
    double pr = power.real();
    if (pr <= 0.0) pr = std::abs(power);
    if (pr <= 0.0) continue;
    field *= 1.0 / std::sqrt(pr);
    mode.field = field;
    modes.push_back(std::move(mode));
    if ((int)modes.size() >= num_modes) break;
  }
  std::sort(modes.begin(), modes.end(), [](const PortMode &a, const PortMode &b) { return a.fc < b.fc; });
  if ((int)modes.size() > num_modes) modes.resize(num_modes);
  return modes;
}

PortSurfaceMesh extract_surface_mesh(const Mesh &volume_mesh, int surface_tag) {  // wave_port.cpp:54-112
  PortSurfaceMesh surface;
  for (size_t ti = 0; ti < volume_mesh.tris.size(); ++ti) {
    const auto &tri = volume_mesh.tris[ti];
    if (tri.phys != surface_tag) continue;
    Element t2;
    t2.type = ElemType::Tri3;
    t2.phys = tri.phys;
    for (int k = 0; k < 3; ++k) {
      const auto id = tri.conn[k];
      if (!surface.mesh.nodeIndex.count(id)) {
        const auto &node = volume_mesh.nodes.at(volume_mesh.nodeIndex.at(id));
        Node n;
        n.id = node.id;
        n.xyz = Vector3d(node.xyz.x(), node.xyz.y(), 0.0);
        surface.mesh.nodeIndex[n.id] = (int)surface.mesh.nodes.size();
        surface.mesh.nodes.push_back(n);
      }
      t2.conn[k] = id;
    }
    surface.mesh.tris.push_back(t2);
    surface.volume_tri_indices.push_back((int)ti);
  }
  // boundary lines = edges seen once; the reference iterates an unordered_map (implementation-defined order),
  // here ascending (n0, n1)
  std::map<std::pair<std::int64_t, std::int64_t>, int> cnt;
  for (const auto &tri : surface.mesh.tris)
    for (int e = 0; e < 3; ++e) {
      const auto a = tri.conn[e], b = tri.conn[(e + 1) % 3];
      cnt[{std::min(a, b), std::max(a, b)}] += 1;
    }
  for (const auto &kv : cnt)
    if (kv.second == 1) {
      BoundaryLine bl;
      bl.n0 = kv.first.first;
      bl.n1 = kv.first.second;
      bl.phys = 1;
      surface.mesh.boundary_lines.push_back(bl);
    }
  return surface;
}

WavePort build_wave_port(const Mesh &volume_mesh, const PortSurfaceMesh &surface, const PortMode &mode) {  // wave_port.cpp:114-212
  if (mode.field.size() != surface.mesh.nodes.size()) throw std::runtime_error("Port mode field size does not match surface mesh nodes");
  std::map<int, cplx> accum;
  std::map<int, int> counts;
  double normal_accum = 0.0;
  const cplx j(0.0, 1.0);
  for (size_t ti = 0; ti < surface.mesh.tris.size(); ++ti) {
    const auto &t2 = surface.mesh.tris[ti];
    const auto &t3 = volume_mesh.tris.at(surface.volume_tri_indices.at(ti));
    double G[3][2], area;
    tri_shape(surface.mesh, t2, G, area);
    cplx gx = 0.0, gy = 0.0;
    for (int q = 0; q < 3; ++q) {
      const cplx v = mode.field[surface.mesh.nodeIndex.at(t2.conn[q])];
      gx += G[q][0] * v;
      gy += G[q][1] * v;
    }
    if (mode.kc == 0.0) throw std::runtime_error("Port mode has zero cutoff wavenumber");
    cplx Ex, Ey;
    if (mode.pol == ModePolarization::TE) {
      const cplx f = j * mode.omega * mode.mu / (mode.kc * mode.kc);
      Ex = f * gy;
      Ey = -f * gx;
    } else {
      const cplx f = -mode.beta / (mode.kc * mode.kc);
      Ex = f * gx;
      Ey = f * gy;
    }
    const auto &n0 = volume_mesh.nodes.at(volume_mesh.nodeIndex.at(t3.conn[0])).xyz;
    const auto &n1 = volume_mesh.nodes.at(volume_mesh.nodeIndex.at(t3.conn[1])).xyz;
    const auto &n2 = volume_mesh.nodes.at(volume_mesh.nodeIndex.at(t3.conn[2])).xyz;
    normal_accum += (n1 - n0).cross(n2 - n0).z();
    for (int e = 0; e < 3; ++e) {
      const int ei = t3.edges[e];
      const auto &edge = volume_mesh.edges.at(ei);
      const auto &pa = volume_mesh.nodes.at(volume_mesh.nodeIndex.at(edge.n0)).xyz;
      const auto &pb = volume_mesh.nodes.at(volume_mesh.nodeIndex.at(edge.n1)).xyz;
      const Vector3d ev = pb - pa;
      accum[ei] += Ex * ev.x() + Ey * ev.y();
      counts[ei] += 1;
    }
  }
  const double sign = normal_accum >= 0.0 ? 1.0 : -1.0;
  WavePort port;
  port.surface_tag = surface.mesh.tris.empty() ? 0 : surface.mesh.tris.front().phys;
  port.mode = mode;
  port.weights.resize(accum.size());
  int idx = 0;
  for (const auto &kv : accum) {  // std::map: ascending edge index, like the reference's sort
    port.edges.push_back(kv.first);
    port.weights[idx++] = sign * kv.second / (double)counts[kv.first];
  }
  return port;
}

void populate_te10_field(const PortSurfaceMesh &surface, const RectWaveguidePort &port, PortMode &mode) {  // wave_port.cpp:214-262
  const size_t nn = surface.mesh.nodes.size();
  mode.field.resize(nn);
  double x_min = std::numeric_limits<double>::max();
  for (const auto &n : surface.mesh.nodes) x_min = std::min(x_min, n.xyz.x());
  for (size_t i = 0; i < nn; ++i) mode.field[i] = std::cos(M_PI * (surface.mesh.nodes[i].xyz.x() - x_min) / port.a);
  const double A_sq = 4.0 * std::pow(mode.kc, 4) * port.a / (mode.omega * mode.mu.real() * mode.beta.real() * M_PI * M_PI * port.b);
  mode.field *= std::sqrt(A_sq);
}

WavePort build_wave_port_from_eigenvector(const Mesh &volume_mesh, const PortSurfaceMesh &surface, const VectorXd &eigenvector,
                                          const PortMode &mode, const std::unordered_set<int> &pec_edges) {  // wave_port.cpp:264-309
  WavePort port;
  port.surface_tag = surface.mesh.tris.empty() ? 0 : surface.mesh.tris.front().phys;
  port.mode = mode;
  std::vector<int> edges;
  for (size_t ti = 0; ti < surface.mesh.tris.size(); ++ti) {
    const auto &t3 = volume_mesh.tris.at(surface.volume_tri_indices.at(ti));
    for (int e = 0; e < 3; ++e) edges.push_back(t3.edges[e]);
  }
  std::sort(edges.begin(), edges.end());
  edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
  port.edges = edges;
  port.weights.resize(edges.size());
  double norm_sq = 0.0;
  for (size_t i = 0; i < edges.size(); ++i) {
    const int e = edges[i];
    if (pec_edges.count(e)) {
      port.weights[i] = cplx(0.0, 0.0);
    } else {
      port.weights[i] = cplx(0.0, eigenvector[(size_t)e]);
      norm_sq += eigenvector[(size_t)e] * eigenvector[(size_t)e];
    }
  }
  if (norm_sq > 1e-15 && mode.Z0.real() > 1e-15) port.weights *= std::sqrt(std::sqrt(mode.Z0.real()) / norm_sq);
  return port;
}

// ------------------------------------------------------------------ Touchstone (src/io/touchstone.cpp)
namespace {
const char *format_string(TouchstoneFormat f) { return f == TouchstoneFormat::MA ? "MA" : (f == TouchstoneFormat::DB ? "DB" : "RI"); }

void write_value(std::ostream &os, cplx v, TouchstoneFormat f) {  // touchstone.cpp:24-44
  if (f == TouchstoneFormat::RI) {
    os << v.real() << ' ' << v.imag();
  } else if (f == TouchstoneFormat::MA) {
    os << std::abs(v) << ' ' << std::arg(v) * 180.0 / M_PI;
  } else {
    os << 20.0 * std::log10(std::max(std::abs(v), 1e-20)) << ' ' << std::arg(v) * 180.0 / M_PI;
  }
}
}  // namespace

void write_touchstone(const std::string &path, const std::vector<double> &freq, const std::vector<SParams2> &data) {  // touchstone.cpp:50-62
  std::ofstream ofs(path);
  ofs << "# Hz S RI R 50\n";
  ofs << std::setprecision(12);
  for (size_t i = 0; i < freq.size(); ++i) {
    const auto &s = data[i];
    ofs << freq[i] << ' ' << s.s11.real() << ' ' << s.s11.imag() << ' ' << s.s21.real() << ' ' << s.s21.imag() << ' ' << s.s12.real() << ' '
        << s.s12.imag() << ' ' << s.s22.real() << ' ' << s.s22.imag() << '\n';
  }
}

void write_touchstone_nport(const std::string &path, const std::vector<double> &freq, const std::vector<MatrixXcd> &S,
                            const TouchstoneOptions &opts) {  // touchstone.cpp:65-139
  if (freq.empty() || S.empty()) throw std::runtime_error("Empty frequency or S-parameter data");
  const int np = S[0].rows();
  if (np < 1 || np > 99) throw std::runtime_error("Invalid number of ports: " + std::to_string(np));
  for (size_t i = 0; i < S.size(); ++i)
    if (S[i].rows() != np || S[i].cols() != np) throw std::runtime_error("S-matrix dimension mismatch at frequency " + std::to_string(i));
  if (freq.size() != S.size()) throw std::runtime_error("Frequency and S-matrix count mismatch");
  std::ofstream ofs(path);
  if (!ofs) throw std::runtime_error("Cannot open file for writing: " + path);
  ofs << std::setprecision(12);
  ofs << "! Touchstone file generated by EdgeFEM\n";
  ofs << "! Number of ports: " << np << "\n";
  ofs << "# Hz S " << format_string(opts.format) << " R " << opts.z0 << "\n";
  for (size_t fi = 0; fi < freq.size(); ++fi) {
    ofs << freq[fi];
    int count = 0;
    for (int i = 0; i < np; ++i)
      for (int j = 0; j < np; ++j) {
        if (np > 2 && count > 0 && count % 4 == 0) ofs << '\n';  // continuation line, 4 complex values per line
        ofs << ' ';
        write_value(ofs, S[fi](i, j), opts.format);
        ++count;
      }
    ofs << '\n';
  }
}

std::string touchstone_extension(int num_ports) {  // touchstone.cpp:141-149
  if (num_ports < 1 || num_ports > 99) throw std::runtime_error("Invalid number of ports: " + std::to_string(num_ports));
  std::ostringstream ss;
  ss << ".s" << num_ports << "p";
  return ss.str();
}

}  // namespace edgefem
