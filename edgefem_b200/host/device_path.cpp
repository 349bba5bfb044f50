// The GPU-backed implementations of EdgeFEM's hot-path functions: assemble_maxwell,
// solve_linear, calculate_sparams*, normalize_port_weights, assemble_maxwell_km, frequency_sweep.
// Host code only marshals inputs and calls the C-ABI (include/edgefem_b200.h); all element
// integration, boundary terms, Krylov iterations and projections run on the device.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <limits>
#include <mutex>
#include <set>

#include "edgefem/maxwell.hpp"
#include "edgefem/sweep.hpp"
#include "host_internal.hpp"

namespace edgefem {
namespace {

constexpr double c0 = 299792458.0;
constexpr double mu0 = 4.0 * M_PI * 1e-7;
constexpr double eps0 = 1.0 / (mu0 * c0 * c0);
constexpr double eta0 = mu0 * c0;
const cplx kNaN(std::numeric_limits<double>::quiet_NaN(), std::numeric_limits<double>::quiet_NaN());

// ------------------------------------------------------------------ RAII over the C-ABI
struct DeviceSystem {
  efb_system *h = nullptr;
  int m = 0, n_matrix = 0, n_rhs = 0;
  std::int64_t nnz = 0;
  ~DeviceSystem() {
    if (h) efb_system_destroy(h);
  }
};

struct DeviceMesh {
  efb_mesh *h = nullptr;
  std::vector<int32_t> slot_tags;
  long long upload_bytes = 0;
  // fingerprint
  const Mesh *addr = nullptr;
  size_t n_nodes = 0, n_tets = 0, n_edges = 0;
  std::uint64_t hash = 0;
  // Per-mesh reuse for drivers that are called once per frequency (calculate_sparams_eigenmode in a Python loop,
  // python/edgefem/designs/waveguide.py:394-433): the device system of the last call (pattern, maps, solver structures,
  // cluster plan; every value is rewritten by efb_assemble_volume) and the host-side port surface mass matrices.
  struct PooledSystem {
    std::uint64_t key = 0;
    std::unique_ptr<DeviceSystem> sys;
  };
  std::vector<PooledSystem> pool;  // at most 2, small systems only
  struct PortMass {
    int tag = 0;
    std::uint64_t dir_hash = 0;
    SparseMatrix<double> Ms;
  };
  std::vector<PortMass> port_mass;  // at most 8
  struct PortTag {
    int surface_tag = 0;
    bool found = false;
    int tag = 0;
  };
  std::vector<PortTag> port_tags;
  std::uint64_t tri_hash = 0;  // of mesh.tris when the two port caches were filled (the device mesh itself does not use them)
  ~DeviceMesh() {
    pool.clear();  // systems refer to the mesh
    if (h) efb_mesh_destroy(h);
  }
};

struct DevicePort {
  efb_port *h = nullptr;
  ~DevicePort() {
    if (h) efb_port_destroy(h);
  }
};

std::mutex g_mu;
efb_ctx *g_ctx = nullptr;
std::vector<std::shared_ptr<DeviceMesh>> g_mesh_cache;  // most recent first, at most 2
long long g_bytes_h2d = 0, g_bytes_d2h = 0;

// EDGEFEM_B200_TRACE=1: wall-clock stage log of the drivers on stderr
struct StageTrace {
  bool on;
  std::chrono::steady_clock::time_point t;
  explicit StageTrace() : on(std::getenv("EDGEFEM_B200_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
  void mark(const char *what) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    std::cerr << "[edgefem-b200 trace] " << what << ": " << std::chrono::duration<double, std::milli>(n - t).count() << " ms\n";
    t = n;
  }
};

std::uint64_t mesh_fingerprint(const Mesh &mesh) {
  std::uint64_t h = 1469598103934665603ull;
  auto mix = [&h](std::uint64_t v) {
    h ^= v;
    h *= 1099511628211ull;
  };
  // EVERY input of the device mesh goes into the key (coordinates, connectivity, tags, edge ids, orientations,
  // edge end nodes): a mesh edited in place, or a new one at the same address with the same counts, must miss.
  // O(N) and cheap next to the marshalling + upload it guards (20 M tets: ~0.3 s against 3 s).
  for (const Element &e : mesh.tets) {
    for (int k = 0; k < 4; ++k) mix((std::uint64_t)e.conn[k]);
    mix((std::uint64_t)(std::int64_t)e.phys);
    for (int k = 0; k < 6; ++k) mix(((std::uint64_t)(std::uint32_t)e.edges[k] << 2) | (std::uint64_t)(e.edge_orient[k] > 0 ? 1 : 0));
  }
  for (const Node &n : mesh.nodes) {
    std::uint64_t b;
    mix((std::uint64_t)n.id);
    for (int a = 0; a < 3; ++a) {
      const double d = n.xyz[a];
      memcpy(&b, &d, 8);
      mix(b);
    }
  }
  for (const Edge &e : mesh.edges) {
    mix((std::uint64_t)e.n0);
    mix((std::uint64_t)e.n1);
  }
  return h;
}

std::shared_ptr<DeviceMesh> device_mesh_for(const Mesh &mesh) {
  efb_ctx *ctx = detail::device_ctx();
  const std::uint64_t fp = mesh_fingerprint(mesh);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto &c : g_mesh_cache)
      if (c->addr == &mesh && c->n_nodes == mesh.nodes.size() && c->n_tets == mesh.tets.size() && c->n_edges == mesh.edges.size() && c->hash == fp)
        return c;
  }
  if (mesh.edges.empty()) throw std::runtime_error("mesh has no edges (was it loaded with load_gmsh_v2 / build_edges?)");
  const size_t nn = mesh.nodes.size(), nt = mesh.tets.size(), ne = mesh.edges.size();
  std::vector<double> xyz(3 * nn);
  for (size_t i = 0; i < nn; ++i)
    for (int a = 0; a < 3; ++a) xyz[3 * i + a] = mesh.nodes[i].xyz[a];
  std::vector<int32_t> tn(4 * nt), te(6 * nt), tp(nt), en(2 * ne);
  std::vector<int8_t> to(6 * nt);
  for (size_t t = 0; t < nt; ++t) {
    const Element &e = mesh.tets[t];
    for (int k = 0; k < 4; ++k) tn[4 * t + k] = mesh.nodeIndex.at(e.conn[k]);
    for (int k = 0; k < 6; ++k) {
      te[6 * t + k] = e.edges[k];
      to[6 * t + k] = (int8_t)e.edge_orient[k];
    }
    tp[t] = e.phys;
  }
  for (size_t e = 0; e < ne; ++e) {
    en[2 * e] = mesh.nodeIndex.at(mesh.edges[e].n0);
    en[2 * e + 1] = mesh.nodeIndex.at(mesh.edges[e].n1);
  }
  efb_mesh_desc d{};
  d.n_node = (int32_t)nn;
  d.xyz = xyz.data();
  d.n_tet = (int32_t)nt;
  d.tet_nodes = tn.data();
  d.tet_edges = te.data();
  d.tet_orient = to.data();
  d.tet_phys = tp.data();
  d.n_edge = (int32_t)ne;
  d.edge_nodes = en.data();
  auto dm = std::make_shared<DeviceMesh>();
  detail::check(efb_mesh_create(ctx, &d, &dm->h), "efb_mesh_create");
  dm->slot_tags.resize(efb_mesh_num_slots(dm->h));
  detail::check(efb_mesh_get_slot_tags(dm->h, dm->slot_tags.data()), "efb_mesh_get_slot_tags");
  dm->upload_bytes = (long long)(xyz.size() * 8 + tn.size() * 4 + te.size() * 4 + to.size() + tp.size() * 4 + en.size() * 4);
  dm->addr = &mesh;
  dm->n_nodes = nn;
  dm->n_tets = nt;
  dm->n_edges = ne;
  dm->hash = fp;
  std::lock_guard<std::mutex> lk(g_mu);
  g_bytes_h2d += dm->upload_bytes;
  g_mesh_cache.insert(g_mesh_cache.begin(), dm);
  if (g_mesh_cache.size() > 2) g_mesh_cache.pop_back();
  return dm;
}


std::uint64_t fnv_bytes(std::uint64_t h, const void *ptr, size_t bytes) {
  const unsigned char *b = (const unsigned char *)ptr;
  size_t i = 0;
  for (; i + 8 <= bytes; i += 8) {
    std::uint64_t w;
    std::memcpy(&w, b + i, 8);
    h = (h ^ w) * 1099511628211ull;
  }
  for (; i < bytes; ++i) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}

constexpr double kPoolMaxBytes = 512e6;

// surface triangles (connectivity, tags, edge ids) are not part of the device-mesh fingerprint; the port caches hang on them
void sync_port_caches(DeviceMesh &dm, const Mesh &mesh) {
  std::uint64_t h = 1469598103934665603ull;
  for (const Element &t : mesh.tris) {
    const std::int64_t rec[7] = {t.conn[0], t.conn[1], t.conn[2], (std::int64_t)t.phys, (std::int64_t)t.edges[0], (std::int64_t)t.edges[1],
                                 (std::int64_t)t.edges[2]};
    h = fnv_bytes(h, rec, sizeof rec);
  }
  std::lock_guard<std::mutex> lk(g_mu);
  if (dm.tri_hash != h) {
    dm.port_mass.clear();
    dm.port_tags.clear();
    dm.tri_hash = h;
  }
}

std::unique_ptr<DeviceSystem> pool_take(DeviceMesh &dm, std::uint64_t key, int n_matrix, int n_rhs) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (size_t i = 0; i < dm.pool.size(); ++i)
    if (dm.pool[i].key == key && dm.pool[i].sys->n_matrix == n_matrix && dm.pool[i].sys->n_rhs == n_rhs) {
      auto s = std::move(dm.pool[i].sys);
      dm.pool.erase(dm.pool.begin() + (long)i);
      return s;
    }
  return nullptr;
}

void pool_give(DeviceMesh &dm, std::uint64_t key, std::unique_ptr<DeviceSystem> sys) {
  if (!sys || (double)sys->nnz * 16.0 * sys->n_matrix + (double)sys->m * 160.0 * sys->n_matrix * sys->n_rhs > kPoolMaxBytes) return;
  std::lock_guard<std::mutex> lk(g_mu);
  DeviceMesh::PooledSystem e;
  e.key = key;
  e.sys = std::move(sys);
  dm.pool.insert(dm.pool.begin(), std::move(e));
  if (dm.pool.size() > 2) dm.pool.pop_back();
}

}  // namespace

namespace detail {
efb_mesh *device_mesh_handle(const Mesh &mesh) { return device_mesh_for(mesh)->h; }
}  // namespace detail

namespace {

std::vector<uint8_t> dirichlet_flags(const Mesh &mesh, const BC &bc) {
  std::vector<uint8_t> f(mesh.edges.size(), 0);
  for (int e : bc.dirichlet_edges)
    if (e >= 0 && (size_t)e < f.size()) f[e] = 1;
  return f;
}

bool port_valid(const WavePort &p) {  // src/assemble_maxwell.cpp:214-217
  return p.weights.size() == p.edges.size() && p.mode.Z0 != cplx(0.0);
}

// complex-symmetric admittance block w w^H/Z0 needs all weights to share one complex phase
bool weights_phase_aligned(const WavePort &p) {
  cplx ref(0.0);
  double mx = 0.0;
  for (size_t k = 0; k < p.weights.size(); ++k)
    if (std::abs(p.weights[k]) > mx) {
      mx = std::abs(p.weights[k]);
      ref = p.weights[k];
    }
  if (mx == 0.0) return true;
  const cplx u = std::conj(ref) / mx;
  for (size_t k = 0; k < p.weights.size(); ++k)
    if (std::abs((p.weights[k] * u).imag()) > 1e-12 * mx) return false;
  return true;
}

std::vector<int32_t> abc_edge_list(const Mesh &mesh, const MaxwellParams &p, const std::vector<uint8_t> &dir) {
  std::set<int> s;  // src/assemble_maxwell.cpp:249-260
  for (const auto &tri : mesh.tris) {
    if (!p.abc_surface_tags.empty() && !p.abc_surface_tags.count(tri.phys)) continue;
    for (int e = 0; e < 3; ++e)
      if (!dir[tri.edges[e]]) s.insert(tri.edges[e]);
  }
  return std::vector<int32_t>(s.begin(), s.end());
}

// port-ABC diagonal coefficient (src/assemble_maxwell.cpp:274-309); false => nothing to add
bool port_abc_coeff(const MaxwellParams &p, const WavePort &port, double omega, cplx &out) {
  if (port.mode.Z0 == cplx(0.0)) return false;
  const double k0 = omega / c0, kc = port.mode.kc;
  const double beta_sq = k0 * k0 - kc * kc;
  if (!(beta_sq > 0)) return false;
  const double beta = std::sqrt(beta_sq), z0r = std::real(port.mode.Z0);
  switch (p.port_abc_type) {
    case PortABCType::Beta: out = cplx(0.0, beta); break;
    case PortABCType::BetaNorm: out = cplx(0.0, beta / k0); break;
    case PortABCType::ImpedanceMatch: out = cplx(0.0, beta * std::sqrt(z0r / eta0)); break;
    case PortABCType::ModalAdmittance: out = cplx(0.0, omega * eps0 / z0r); break;
    case PortABCType::None: out = cplx(0.0, 0.0); break;
  }
  return true;
}

// ------------------------------------------------------------------ materials marshalling
struct MaterialTable {
  std::vector<cplx> eps, mu;
  std::vector<efb_model> em, mm;
  std::vector<efb_pole> poles;
  std::vector<efb_pml> pml;
  bool opaque = false;  // a user-defined model without describe(): host evaluation per frequency
  efb_materials view() {
    efb_materials m{};
    m.n_slots = (int32_t)eps.size();
    m.eps_static_c128 = reinterpret_cast<const double *>(eps.data());
    m.mu_static_c128 = reinterpret_cast<const double *>(mu.data());
    m.eps_models = em.data();
    m.mu_models = mm.data();
    m.n_poles = (int32_t)poles.size();
    m.poles = poles.empty() ? nullptr : poles.data();
    m.pml = pml.data();
    return m;
  }
};

// `host_omega` >= 0: evaluate every dispersive model on the host at that omega (static table)
MaterialTable build_materials(const MaxwellParams &p, const std::vector<int32_t> &tags, bool dispersive, double host_omega = -1.0) {
  MaterialTable T;
  const size_t n = tags.size();
  T.eps.resize(n);
  T.mu.resize(n);
  T.em.assign(n, efb_model{});
  T.mm.assign(n, efb_model{});
  T.pml.assign(n, efb_pml{});
  auto put_model = [&](const materials::DispersiveMaterial &mdl, efb_model &dst) {
    const materials::ModelDescription d = mdl.describe();
    if (d.kind == 0) return false;
    dst.kind = d.kind;
    dst.p0 = d.p0;
    dst.p1 = d.p1;
    dst.p2 = d.p2;
    dst.pole_begin = (int32_t)T.poles.size();
    dst.n_poles = (int32_t)d.poles.size();
    for (const auto &q : d.poles) T.poles.push_back(efb_pole{q[0], q[1], q[2]});
    return true;
  };
  for (size_t s = 0; s < n; ++s) {
    const int tag = tags[s];
    T.eps[s] = p.get_eps_r(tag);
    T.mu[s] = p.get_mu_r(tag);
    if (dispersive) {
      auto ie = p.eps_models.find(tag);
      if (ie != p.eps_models.end() && ie->second) {
        if (host_omega >= 0.0) T.eps[s] = ie->second->eval_eps(host_omega);
        else if (!put_model(*ie->second, T.em[s])) T.opaque = true;
      }
      auto im = p.mu_models.find(tag);
      if (im != p.mu_models.end() && im->second) {
        // eval_mu may be overridden by user models: always resolved on the host side
        if (host_omega >= 0.0) T.mu[s] = im->second->eval_mu(host_omega);
        else T.opaque = true;
      }
      auto it = p.pml_tensor_regions.find(tag);
      if (it != p.pml_tensor_regions.end()) {  // tensor spec wins over pml_regions (assemble_maxwell.cpp:127-171)
        efb_pml &q = T.pml[s];
        q.kind = EFB_PML_TENSOR;
        q.enforce_heuristics = p.enforce_pml_heuristics ? 1 : 0;
        for (int a = 0; a < 3; ++a) {
          q.sigma[a] = it->second.sigma_max[a];
          q.thickness[a] = it->second.thickness[a];
        }
        q.grading_order = it->second.grading_order;
      } else if (p.pml_regions.count(tag)) {
        T.pml[s].kind = EFB_PML_UNIFORM;
        T.pml[s].sigma[0] = p.pml_sigma;
      }
    }
  }
  return T;
}

std::vector<PMLDiagnostic> pml_diagnostics(const MaxwellParams &p) {  // src/assemble_maxwell.cpp:91-112
  std::vector<PMLDiagnostic> out;
  const double w = std::abs(p.omega);
  for (const auto &kv : p.pml_tensor_regions) {
    PMLDiagnostic d;
    d.region_tag = kv.first;
    d.sigma_max = kv.second.sigma_max;
    d.thickness = kv.second.thickness;
    for (int a = 0; a < 3; ++a) {
      const double s = kv.second.sigma_max[a], t = kv.second.thickness[a];
      d.reflection_est[a] = (s <= 0.0 || t <= 0.0 || w <= 0.0) ? 1.0 : std::exp(-2.0 * (s * t / (kv.second.grading_order + 1.0)) / w);
    }
    out.push_back(d);
  }
  return out;
}

// ------------------------------------------------------------------ system building blocks
std::unique_ptr<DeviceSystem> make_system(DeviceMesh &dm, const std::vector<int32_t> &xr, const std::vector<int32_t> &xc, int n_matrix,
                                          int n_rhs) {
  auto s = std::make_unique<DeviceSystem>();
  detail::check(efb_system_create(dm.h, (int64_t)xr.size(), xr.data(), xc.data(), n_matrix, n_rhs, &s->h), "efb_system_create");
  detail::check(efb_system_dims(s->h, &s->m, &s->nnz, &s->n_matrix, &s->n_rhs), "efb_system_dims");
  g_bytes_h2d += (long long)xr.size() * 8;
  return s;
}

void assemble_volume(DeviceSystem &sys, const DeviceMesh &dm, const MaxwellParams &p, const std::vector<double> &omegas, int first, int mode) {
  const bool dispersive = (mode == 0);
  MaterialTable T = build_materials(p, dm.slot_tags, dispersive);
  if (!T.opaque) {
    efb_materials mv = T.view();
    detail::check(efb_assemble_volume(sys.h, first, (int32_t)omegas.size(), omegas.data(), &mv, mode), "efb_assemble_volume");
    return;
  }
  for (size_t f = 0; f < omegas.size(); ++f) {  // user-defined model objects: evaluate on the host per frequency
    MaterialTable Tf = build_materials(p, dm.slot_tags, true, omegas[f]);
    efb_materials mv = Tf.view();
    detail::check(efb_assemble_volume(sys.h, first + (int)f, 1, &omegas[f], &mv, mode), "efb_assemble_volume");
  }
}

std::unique_ptr<DevicePort> make_port(DeviceSystem &sys, const WavePort &port, const SparseMatrix<double> *Ms) {
  std::vector<int32_t> e(port.edges.begin(), port.edges.end());
  std::vector<int32_t> r, c;
  const double *vals = nullptr;
  if (Ms) {
    const auto &rp = Ms->rowptr();
    r.reserve(Ms->nonZeros());
    for (int i = 0; i < Ms->rows(); ++i)
      for (int k = rp[i]; k < rp[i + 1]; ++k) r.push_back(i);
    c.assign(Ms->colidx().begin(), Ms->colidx().end());
    vals = Ms->values().data();
  }
  auto dp = std::make_unique<DevicePort>();
  detail::check(efb_port_create(sys.h, (int32_t)e.size(), e.data(), reinterpret_cast<const double *>(port.weights.data()), (int64_t)r.size(),
                                r.data(), c.data(), vals, &dp->h),
                "efb_port_create");
  g_bytes_h2d += (long long)e.size() * 20 + (long long)r.size() * 16;
  return dp;
}

struct SolveOutcome {
  std::vector<efb_solve_result> res;
  std::string method;
};

std::string method_name(int method, int precond) {
  std::string s = "B200:";
  s += (method == EFB_METHOD_COCG) ? "COCG" : "BiCGSTAB";
  s += (precond == EFB_PRECOND_AUX) ? "+AUX" : (precond == EFB_PRECOND_JACOBI ? "+Jacobi" : "");
  return s;
}

SolveOutcome solve_on_device(DeviceSystem &sys, int first, int count, const SolveOptions &opt, bool symmetric) {
  efb_solve_opts o{};
  o.method = EFB_METHOD_AUTO;
  o.precond = opt.use_ilut ? EFB_PRECOND_AUX : EFB_PRECOND_JACOBI;
  o.tolerance = opt.use_direct ? std::min(opt.tolerance, 1e-12) : opt.tolerance;
  o.max_iterations = opt.max_iterations;
  o.check_every = 0;
  o.symmetric_hint = symmetric ? 1 : 0;
  o.zero_initial_guess = 1;
  o.max_restarts = 3;
  SolveOutcome out;
  out.res.resize((size_t)count * sys.n_rhs);
  detail::check(efb_solve(sys.h, first, count, &o, out.res.data()), "efb_solve");
  auto all_ok = [&]() {
    for (auto &r : out.res)
      if (!r.converged && !(opt.use_direct && r.residual <= opt.tolerance)) return false;
    return true;
  };
  out.method = method_name(out.res[0].method, out.res[0].precond);
  if (!all_ok() && opt.auto_fallback) {
    // The reference falls back to SparseLU (src/solver.cpp:55-80, method "<failed>->SparseLU").  Here: every matrix with an
    // unconverged right-hand side is re-solved by the dense LU on the device (efb_solve_direct) when it is small enough for
    // it; larger symmetric systems get a second Krylov attempt with the general method, continuing from the current iterate.
    bool direct_done = false;
    for (int f = 0; f < count; ++f) {
      bool bad = false;
      for (int k = 0; k < sys.n_rhs; ++k) {
        const efb_solve_result &r = out.res[(size_t)f * sys.n_rhs + k];
        bad = bad || (!r.converged && !(opt.use_direct && r.residual <= opt.tolerance));
      }
      if (!bad) continue;
      std::vector<efb_solve_result> dr((size_t)sys.n_rhs);
      const int rc = efb_solve_direct(sys.h, first + f, 1, dr.data());
      if (rc == EFB_ERR_LIMIT) break;  // too large for the dense factorisation
      detail::check(rc, "efb_solve_direct");
      for (int k = 0; k < sys.n_rhs; ++k) {
        dr[k].iters += out.res[(size_t)f * sys.n_rhs + k].iters;
        out.res[(size_t)f * sys.n_rhs + k] = dr[k];
      }
      direct_done = true;
    }
    if (direct_done) {
      out.method += "->B200:DenseLU";
    } else if (symmetric) {
      std::vector<efb_solve_result> first_try = out.res;
      o.method = EFB_METHOD_BICGSTAB;
      o.zero_initial_guess = 0;
      detail::check(efb_solve(sys.h, first, count, &o, out.res.data()), "efb_solve");
      for (size_t i = 0; i < out.res.size(); ++i) out.res[i].iters += first_try[i].iters;
      out.method += "->" + method_name(out.res[0].method, out.res[0].precond);
    }
  }
  if (opt.use_direct)
    for (auto &r : out.res)
      if (!r.converged && r.residual <= opt.tolerance) r.converged = 1;
  return out;
}

void warn_not_converged(const efb_solve_result &r, const std::string &method, const std::string &context) {
  std::cerr << "WARNING: Solver did not converge in " << context << " (method=" << method << ", iters=" << r.iters
            << ", residual=" << r.residual << "). S-parameters may be unreliable." << std::endl;
}

SpMatC download_matrix(DeviceSystem &sys, int matrix) {
  SpMatC A(sys.m, sys.m);
  A.rowptr().resize((size_t)sys.m + 1);
  A.colidx().resize((size_t)sys.nnz);
  A.values().resize((size_t)sys.nnz);
  static_assert(sizeof(int) == sizeof(int32_t), "int32 CSR");
  detail::check(efb_system_get_pattern(sys.h, A.rowptr().data(), A.colidx().data()), "efb_system_get_pattern");
  detail::check(efb_system_get_values(sys.h, matrix, reinterpret_cast<double *>(A.values().data())), "efb_system_get_values");
  g_bytes_d2h += (long long)sys.nnz * 20;
  return A;
}

VecC download_vec(DeviceSystem &sys, int idx, bool rhs) {
  VecC v(sys.m);
  detail::check((rhs ? efb_rhs_get : efb_x_get)(sys.h, idx, reinterpret_cast<double *>(v.data())), rhs ? "efb_rhs_get" : "efb_x_get");
  g_bytes_d2h += (long long)sys.m * 16;
  return v;
}

// Standard (lumped / analytic wave-port) system: volume + w w^H / Z0 blocks + ABC + port ABC,
// for a list of frequencies, one right-hand side per port (src/assemble_maxwell.cpp:46-350).
struct StdSystem {
  std::shared_ptr<DeviceMesh> dm;
  std::unique_ptr<DeviceSystem> sys;
  std::vector<std::unique_ptr<DevicePort>> dports;  // null for invalid ports
  std::vector<uint8_t> dir;
  bool symmetric = true;
};

void standard_extras(const Mesh &mesh, const MaxwellParams &p, const std::vector<WavePort> &ports, const std::vector<uint8_t> &dir,
                     bool with_port_abc, std::vector<int32_t> &xr, std::vector<int32_t> &xc, std::vector<int32_t> &abc_edges) {
  for (size_t e = 0; e < dir.size(); ++e)
    if (dir[e]) {
      xr.push_back((int32_t)e);
      xc.push_back((int32_t)e);
    }
  for (const auto &port : ports) {
    if (!port_valid(port)) continue;
    std::vector<int32_t> fe;
    for (int e : port.edges)
      if (!dir.at(e)) fe.push_back(e);
    for (int a : fe)
      for (int b : fe) {
        xr.push_back(a);
        xc.push_back(b);
      }
  }
  if (p.use_abc) {
    abc_edges = abc_edge_list(mesh, p, dir);
    for (int e : abc_edges) {
      xr.push_back(e);
      xc.push_back(e);
    }
  }
  if (with_port_abc && p.use_port_abc && p.port_abc_type != PortABCType::None)
    for (const auto &port : ports) {
      if (port.mode.Z0 == cplx(0.0)) continue;
      for (int e : port.edges)
        if (!dir.at(e)) {
          xr.push_back(e);
          xc.push_back(e);
        }
    }
}

// pairs that take part in the elimination: both edges free (src/assemble_maxwell.cpp:534-537)
struct ActivePairs {
  std::vector<int32_t> master, slave;
  std::vector<cplx> phase;
};
ActivePairs active_pairs(const PeriodicBC &pbc, const std::vector<uint8_t> &dir) {
  ActivePairs a;
  for (const auto &pr : pbc.pairs) {
    if (dir.at(pr.master_edge) || dir.at(pr.slave_edge)) continue;
    a.master.push_back(pr.master_edge);
    a.slave.push_back(pr.slave_edge);
    a.phase.push_back(pbc.phase_shift * static_cast<double>(pr.master_orient * pr.slave_orient));
  }
  return a;
}

StdSystem build_standard(const Mesh &mesh, const MaxwellParams &p, const BC &bc, const std::vector<WavePort> &ports,
                         const std::vector<double> &omegas, int n_rhs, int extra_matrices = 0, const PeriodicBC *pbc = nullptr) {
  StdSystem S;
  S.dm = device_mesh_for(mesh);
  S.dir = dirichlet_flags(mesh, bc);
  std::vector<int32_t> xr, xc, abc_edges;
  standard_extras(mesh, p, ports, S.dir, true, xr, xc, abc_edges);
  if (pbc && !pbc->pairs.empty()) {
    const ActivePairs ap = active_pairs(*pbc, S.dir);
    int64_t n_more = 0;
    detail::check(efb_periodic_extra(S.dm->h, (int64_t)xr.size(), xr.data(), xc.data(), (int32_t)ap.master.size(), ap.master.data(),
                                     ap.slave.data(), &n_more, nullptr, nullptr), "efb_periodic_extra");
    const size_t base = xr.size();
    xr.resize(base + (size_t)n_more);
    xc.resize(base + (size_t)n_more);
    detail::check(efb_periodic_extra(S.dm->h, (int64_t)base, xr.data(), xc.data(), (int32_t)ap.master.size(), ap.master.data(),
                                     ap.slave.data(), &n_more, xr.data() + base, xc.data() + base), "efb_periodic_extra");
  }
  const int F = (int)omegas.size();
  S.sys = make_system(*S.dm, xr, xc, F + extra_matrices, std::max(1, n_rhs));
  detail::check(efb_system_set_dirichlet(S.sys->h, S.dir.data()), "efb_system_set_dirichlet");
  g_bytes_h2d += (long long)S.dir.size();
  assemble_volume(*S.sys, *S.dm, p, omegas, 0, 0);
  S.dports.resize(ports.size());
  for (size_t i = 0; i < ports.size(); ++i) {
    if (!port_valid(ports[i])) continue;
    S.symmetric = S.symmetric && weights_phase_aligned(ports[i]);
    S.dports[i] = make_port(*S.sys, ports[i], nullptr);
    std::vector<cplx> coef(F, cplx(1.0) / ports[i].mode.Z0);
    detail::check(efb_port_add_block(S.sys->h, S.dports[i]->h, 0, F, reinterpret_cast<const double *>(coef.data())), "efb_port_add_block");
  }
  if (p.use_abc && !abc_edges.empty()) {
    std::vector<cplx> coef(F);
    for (int f = 0; f < F; ++f) coef[f] = cplx(0.0, omegas[f] / c0);
    detail::check(efb_add_diag(S.sys->h, 0, F, (int32_t)abc_edges.size(), abc_edges.data(), reinterpret_cast<const double *>(coef.data())), "efb_add_diag");
  }
  if (p.use_port_abc && p.port_abc_type != PortABCType::None)
    for (const auto &port : ports) {
      std::vector<int32_t> e(port.edges.begin(), port.edges.end());
      if (e.empty()) continue;
      for (int f = 0; f < F; ++f) {  // beta^2 > 0 is a per-frequency condition
        cplx cf;
        if (!port_abc_coeff(p, port, omegas[f], cf)) continue;
        detail::check(efb_add_diag(S.sys->h, f, 1, (int32_t)e.size(), e.data(), reinterpret_cast<const double *>(&cf)), "efb_add_diag");
      }
    }
  return S;
}

void add_source(StdSystem &S, const std::vector<WavePort> &ports, int port_idx, int rhs) {  // src/assemble_maxwell.cpp:233-241
  if (port_idx < 0 || port_idx >= (int)ports.size() || !S.dports[port_idx]) return;
  const cplx scale = 2.0 / std::sqrt(ports[port_idx].mode.Z0);
  detail::check(efb_port_rhs_weights(S.sys->h, S.dports[port_idx]->h, rhs, reinterpret_cast<const double *>(&scale)), "efb_port_rhs_weights");
}

// V_j = sum conj(w_jk) x(edge_jk) over non-PEC edges.  Ports that were skipped as invalid still
// project (the reference does not filter them in the extraction loop): build a weights-only port.
cplx project_weights(StdSystem &S, const std::vector<WavePort> &ports, int j, int rhs) {
  cplx v(0.0);
  if (S.dports[j]) {
    detail::check(efb_port_project_weights(S.sys->h, S.dports[j]->h, rhs, reinterpret_cast<double *>(&v)), "efb_port_project_weights");
  } else if (ports[j].weights.size() == ports[j].edges.size() && !ports[j].edges.empty()) {
    auto tmp = make_port(*S.sys, ports[j], nullptr);
    detail::check(efb_port_project_weights(S.sys->h, tmp->h, rhs, reinterpret_cast<double *>(&v)), "efb_port_project_weights");
  }
  g_bytes_d2h += 16;
  return v;
}

void fill_sparams_column(MatrixXcd &Smat, StdSystem &S, const std::vector<WavePort> &ports, int active, int rhs) {
  const cplx vinc = std::sqrt(ports[active].mode.Z0);
  for (int j = 0; j < (int)ports.size(); ++j) {
    const cplx vj = project_weights(S, ports, j, rhs);
    Smat(j, active) = (j == active) ? (vj - vinc) / vinc : vj / vinc;
  }
}

// beta_i = sqrt(eps mu k0^2 - kc^2) with the material of the first tet (file order) that has a
// face on the port surface (src/assemble_maxwell.cpp:652-700).  Returns the tag per port (or a
// flag that no tet was found => eps*mu = 1).
struct PortRegion {
  bool found = false;
  int tag = 0;
};
std::vector<PortRegion> port_regions(const Mesh &mesh, const std::vector<WavePort> &ports) {
  static const int faces[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
  std::vector<PortRegion> out(ports.size());
  for (size_t i = 0; i < ports.size(); ++i) {
    std::set<std::array<std::int64_t, 3>> pf;
    for (const auto &tri : mesh.tris) {
      if (tri.phys != ports[i].surface_tag) continue;
      std::array<std::int64_t, 3> f{tri.conn[0], tri.conn[1], tri.conn[2]};
      std::sort(f.begin(), f.end());
      pf.insert(f);
    }
    if (pf.empty()) continue;
    for (const auto &tet : mesh.tets) {
      for (const auto &tf : faces) {
        std::array<std::int64_t, 3> f{tet.conn[tf[0]], tet.conn[tf[1]], tet.conn[tf[2]]};
        std::sort(f.begin(), f.end());
        if (pf.count(f)) {
          out[i].found = true;
          out[i].tag = tet.phys;
          break;
        }
      }
      if (out[i].found) break;
    }
  }
  return out;
}


std::vector<PortRegion> cached_port_regions(DeviceMesh &dm, const Mesh &mesh, const std::vector<WavePort> &ports) {
  std::vector<PortRegion> out(ports.size());
  std::vector<WavePort> miss;
  std::vector<size_t> miss_at;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t i = 0; i < ports.size(); ++i) {
      bool hit = false;
      for (const auto &t : dm.port_tags)
        if (t.surface_tag == ports[i].surface_tag) {
          out[i].found = t.found;
          out[i].tag = t.tag;
          hit = true;
          break;
        }
      if (!hit) {
        WavePort w;
        w.surface_tag = ports[i].surface_tag;
        miss.push_back(w);
        miss_at.push_back(i);
      }
    }
  }
  if (miss.empty()) return out;
  const std::vector<PortRegion> r = port_regions(mesh, miss);
  std::lock_guard<std::mutex> lk(g_mu);
  for (size_t k = 0; k < miss.size(); ++k) {
    out[miss_at[k]] = r[k];
    DeviceMesh::PortTag t;
    t.surface_tag = miss[k].surface_tag;
    t.found = r[k].found;
    t.tag = r[k].tag;
    dm.port_tags.push_back(t);
  }
  return out;
}

SparseMatrix<double> cached_port_mass(DeviceMesh &dm, const Mesh &mesh, int surface_tag, const BC &bc, std::uint64_t dir_hash) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (const auto &e : dm.port_mass)
      if (e.tag == surface_tag && e.dir_hash == dir_hash) return e.Ms;
  }
  DeviceMesh::PortMass e;
  e.tag = surface_tag;
  e.dir_hash = dir_hash;
  e.Ms = assemble_port_surface_mass(mesh, surface_tag, bc.dirichlet_edges);
  std::lock_guard<std::mutex> lk(g_mu);
  dm.port_mass.insert(dm.port_mass.begin(), e);
  if (dm.port_mass.size() > 8) dm.port_mass.pop_back();
  return e.Ms;
}

} // namespace

// ------------------------------------------------------------------ detail
namespace detail {

efb_ctx *device_ctx() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_ctx) return g_ctx;
  int dev = 0;
  if (const char *e = std::getenv("EDGEFEM_B200_DEVICE")) dev = std::atoi(e);
  else if (const char *l = std::getenv("LOCAL_RANK")) dev = std::atoi(l) % std::max(1, efb_device_count());
  int rc = efb_ctx_create(dev, &g_ctx);
  if (rc != EFB_OK) {
    g_ctx = nullptr;
    throw std::runtime_error(std::string("edgefem-b200: cannot create a CUDA context: ") + efb_last_error(nullptr));
  }
  return g_ctx;
}

void check(int rc, const char *what) {
  if (rc == EFB_OK) return;
  const char *msg = efb_last_error(g_ctx);
  throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + (msg ? msg : "?"));
}

void clear_device_cache() {
  std::lock_guard<std::mutex> lk(g_mu);
  g_mesh_cache.clear();
  efb_clear_caches();  // derived structures shared between systems (cluster-split plans)
}

long long launch_count() { return g_ctx ? (long long)efb_launch_count(g_ctx) : 0; }

} // namespace detail

// ------------------------------------------------------------------ assemble_maxwell
MaxwellAssembly assemble_maxwell(const Mesh &mesh, const MaxwellParams &p, const BC &bc, const std::vector<WavePort> &ports,
                                 int active_port_idx) {
  StdSystem S = build_standard(mesh, p, bc, ports, {p.omega}, 1);
  add_source(S, ports, active_port_idx, 0);
  MaxwellAssembly out;
  out.A = download_matrix(*S.sys, 0);
  out.b = download_vec(*S.sys, 0, true);
  out.diagnostics = pml_diagnostics(p);
  return out;
}

// ------------------------------------------------------------------ solve_linear
SolveResult solve_linear(const SpMatC &A, const VecC &b, const SolveOptions &opt) {
  SolveResult res;
  if (A.rows() != A.cols() || (size_t)A.rows() != b.size()) throw std::invalid_argument("solve_linear: dimension mismatch");
  if (opt.verbose)
    std::cerr << "Solver: Starting B200 Krylov solve (N=" << A.rows() << ", nnz=" << A.nonZeros() << ")\n";
  efb_ctx *ctx = detail::device_ctx();
  DeviceSystem sys;
  detail::check(efb_system_create_csr(ctx, A.rows(), A.nonZeros(), A.rowptr().data(), A.colidx().data(),
                                      reinterpret_cast<const double *>(A.values().data()), 1, 1, &sys.h),
                "efb_system_create_csr");
  detail::check(efb_system_dims(sys.h, &sys.m, &sys.nnz, &sys.n_matrix, &sys.n_rhs), "efb_system_dims");
  detail::check(efb_rhs_set(sys.h, 0, reinterpret_cast<const double *>(b.data())), "efb_rhs_set");
  g_bytes_h2d += A.nonZeros() * 20 + (long long)b.size() * 16;
  // complex symmetric?  (structure + values, tolerance relative to the entry)
  bool symmetric = true;
  {
    const auto &rp = A.rowptr();
    const auto &ci = A.colidx();
    const auto &va = A.values();
    double amax = 0.0;
    for (const auto &v : va) amax = std::max(amax, std::abs(v));
    // assembled matrices are symmetric only up to the rounding of two different summation orders
    for (int i = 0; i < A.rows() && symmetric; ++i)
      for (int k = rp[i]; k < rp[i + 1]; ++k) {
        const int j = ci[k];
        if (j == i) continue;  // both triangles: an entry stored only below the diagonal must be seen too
        const cplx t = A.coeff(j, i);
        if (std::abs(t - va[k]) > 1e-10 * std::max(std::max(std::abs(t), std::abs(va[k])), 1e-3 * amax)) {
          symmetric = false;
          break;
        }
      }
  }
  SolveOptions o = opt;
  SolveOutcome out = solve_on_device(sys, 0, 1, o, symmetric);
  res.method = out.method;
  res.iters = out.res[0].iters;
  res.residual = out.res[0].residual;
  res.converged = out.res[0].converged != 0;
  if (!res.converged) res.error_message = "Solver did not converge within max iterations";
  res.x = download_vec(sys, 0, false);
  if (opt.verbose)
    std::cerr << "Solver: " << res.method << " " << (res.converged ? "CONVERGED" : "FAILED") << " in " << res.iters
              << " iterations, residual=" << std::scientific << res.residual << std::fixed << "\n";
  if (opt.progress_callback) opt.progress_callback(res.iters, res.residual);
  return res;
}

// ------------------------------------------------------------------ calculate_sparams
MatrixXcd calculate_sparams(const Mesh &mesh, const MaxwellParams &p, const BC &bc, const std::vector<WavePort> &ports,
                            const SolveOptions &opts) {
  const int P = (int)ports.size();
  MatrixXcd Smat(P, P);
  if (P == 0) return Smat;
  StdSystem S = build_standard(mesh, p, bc, ports, {p.omega}, P);
  for (int i = 0; i < P; ++i) add_source(S, ports, i, i);
  SolveOutcome out = solve_on_device(*S.sys, 0, 1, opts, S.symmetric);
  for (int i = 0; i < P; ++i) {
    if (!out.res[i].converged) {
      warn_not_converged(out.res[i], out.method, "calculate_sparams (active port " + std::to_string(i) + ")");
      for (int j = 0; j < P; ++j) Smat(j, i) = kNaN;
      continue;
    }
    fill_sparams_column(Smat, S, ports, i, i);
  }
  return Smat;
}

// ------------------------------------------------------------------ periodic (Bloch) path
namespace {
void apply_periodic(StdSystem &S, const PeriodicBC &pbc, int n_matrix) {
  const ActivePairs ap = active_pairs(pbc, S.dir);
  if (ap.master.empty()) return;
  detail::check(efb_apply_periodic(S.sys->h, 0, n_matrix, (int32_t)ap.master.size(), ap.master.data(), ap.slave.data(),
                                   reinterpret_cast<const double *>(ap.phase.data())), "efb_apply_periodic");
  if (std::abs(pbc.phase_shift.imag()) > 1e-14 * std::abs(pbc.phase_shift)) S.symmetric = false;  // T A T^H is not symmetric for complex phi
}

// Eigen's sparseView() (src/assemble_maxwell.cpp:569) keeps only numerically non-zero entries
SpMatC prune_zeros(const SpMatC &A) {
  SpMatC B(A.rows(), A.cols());
  auto &rp = B.rowptr();
  auto &ci = B.colidx();
  auto &va = B.values();
  for (int i = 0; i < A.rows(); ++i) {
    for (int k = A.rowptr()[i]; k < A.rowptr()[i + 1]; ++k)
      if (A.values()[k] != cplx(0.0)) {
        ci.push_back(A.colidx()[k]);
        va.push_back(A.values()[k]);
      }
    rp[i + 1] = (int)ci.size();
  }
  return B;
}
} // namespace

MaxwellAssembly assemble_maxwell_periodic(const Mesh &mesh, const MaxwellParams &p, const BC &bc, const PeriodicBC &pbc,
                                          const std::vector<WavePort> &ports, int active_port_idx) {
  if (pbc.pairs.empty()) return assemble_maxwell(mesh, p, bc, ports, active_port_idx);  // :504-506
  StdSystem S = build_standard(mesh, p, bc, ports, {p.omega}, 1, 0, &pbc);
  add_source(S, ports, active_port_idx, 0);
  apply_periodic(S, pbc, 1);
  MaxwellAssembly out;
  out.A = prune_zeros(download_matrix(*S.sys, 0));
  out.b = download_vec(*S.sys, 0, true);
  out.diagnostics = pml_diagnostics(p);
  return out;
}

MatrixXcd calculate_sparams_periodic(const Mesh &mesh, const MaxwellParams &p, const BC &bc, const PeriodicBC &pbc,
                                     const std::vector<WavePort> &ports) {
  const int P = (int)ports.size();  // src/assemble_maxwell.cpp:576-635
  MatrixXcd Smat(P, P);
  if (P == 0) return Smat;
  StdSystem S = build_standard(mesh, p, bc, ports, {p.omega}, P, 0, &pbc);
  for (int i = 0; i < P; ++i) add_source(S, ports, i, i);
  apply_periodic(S, pbc, 1);
  SolveOptions defaults;  // the reference passes {} here (:593)
  SolveOutcome out = solve_on_device(*S.sys, 0, 1, defaults, S.symmetric);
  // slave recovery x[s] = phi * o * x[m] for EVERY pair (:583-613)
  std::vector<int32_t> dst, src;
  std::vector<cplx> ph;
  for (const auto &pr : pbc.pairs) {
    dst.push_back(pr.slave_edge);
    src.push_back(pr.master_edge);
    ph.push_back(pbc.phase_shift * static_cast<double>(pr.master_orient * pr.slave_orient));
  }
  for (int i = 0; i < P; ++i) {
    if (!out.res[i].converged) {
      warn_not_converged(out.res[i], out.method, "calculate_sparams_periodic (active port " + std::to_string(i) + ")");
      for (int j = 0; j < P; ++j) Smat(j, i) = kNaN;
      continue;
    }
    detail::check(efb_x_recover(S.sys->h, i, (int32_t)dst.size(), dst.data(), src.data(), reinterpret_cast<const double *>(ph.data())), "efb_x_recover");
    fill_sparams_column(Smat, S, ports, i, i);
  }
  return Smat;
}

// ------------------------------------------------------------------ normalize_port_weights
void normalize_port_weights(const Mesh &mesh, const MaxwellParams &p, const BC &bc, std::vector<WavePort> &ports, const SolveOptions &opts) {
  if (ports.empty()) return;  // src/assemble_maxwell.cpp:396-494
  std::vector<int> valid;
  for (size_t i = 0; i < ports.size(); ++i)
    if (port_valid(ports[i])) valid.push_back((int)i);
  if (valid.empty()) return;
  std::vector<WavePort> none;
  StdSystem S = build_standard(mesh, p, bc, none, {p.omega}, (int)valid.size());
  std::vector<std::unique_ptr<DevicePort>> dp(valid.size());
  bool sym = true;
  const cplx one(1.0, 0.0);
  for (size_t k = 0; k < valid.size(); ++k) {
    dp[k] = make_port(*S.sys, ports[valid[k]], nullptr);
    detail::check(efb_port_rhs_weights(S.sys->h, dp[k]->h, (int)k, reinterpret_cast<const double *>(&one)), "efb_port_rhs_weights");
  }
  SolveOutcome out = solve_on_device(*S.sys, 0, 1, opts, sym);
  for (size_t k = 0; k < valid.size(); ++k) {
    WavePort &port = ports[valid[k]];
    if (!out.res[k].converged) {
      warn_not_converged(out.res[k], out.method, "normalize_port_weights");
      std::cerr << "  Skipping normalization for this port due to solver failure." << std::endl;
      continue;
    }
    if (port.weights.norm() < 1e-15) {
      std::cerr << "WARNING: Port weight vector norm is near-zero. This indicates a port formulation issue." << std::endl;
      continue;
    }
    cplx wAw(0.0);
    detail::check(efb_port_project_weights(S.sys->h, dp[k]->h, (int)k, reinterpret_cast<double *>(&wAw)), "efb_port_project_weights");
    int free_count = 0;
    for (int e : port.edges) free_count += S.dir.at(e) ? 0 : 1;
    const double mag = std::abs(wAw), target = std::real(port.mode.Z0);
    std::cerr << "  Port normalization: free_edges=" << free_count << ", wAinvw_exact=" << mag << ", target=Z0=" << target
              << ", ratio=" << mag / target << std::endl;
    if (mag > 1e-15 && target > 1e-15) {
      const double alpha = std::sqrt(target / mag);
      std::cerr << "  Applying scale factor alpha=" << alpha << std::endl;
      port.weights *= alpha;
    }
  }
}

// ------------------------------------------------------------------ eigenmode S-parameters
std::vector<MatrixXcd> calculate_sparams_eigenmode_sweep(const Mesh &mesh, const MaxwellParams &p, const BC &bc,
                                                         const std::vector<WavePort> &ports, const std::vector<double> &frequencies,
                                                         BatchStats *stats) {
  const int P = (int)ports.size();
  const int F = (int)frequencies.size();
  std::vector<MatrixXcd> result(F, MatrixXcd(P, P));
  if (P == 0 || F == 0) return result;
  const long long h2d0 = g_bytes_h2d, d2h0 = g_bytes_d2h, launches0 = detail::launch_count();
  StageTrace tr;
  auto dm = device_mesh_for(mesh);
  tr.mark("device mesh (flatten + upload + incidence lists + geometry)");
  const std::vector<uint8_t> dir = dirichlet_flags(mesh, bc);
  std::vector<int32_t> xr, xc;
  for (size_t e = 0; e < dir.size(); ++e)
    if (dir[e]) {
      xr.push_back((int32_t)e);
      xc.push_back((int32_t)e);
    }
  // M_s entries lie inside the volume pattern for a port face of a tet; listed anyway so a
  // tri-only port edge cannot fall outside (coeffRef would insert it, assemble_maxwell.cpp:743)
  const std::uint64_t dir_hash = fnv_bytes(1469598103934665603ull, dir.data(), dir.size());
  sync_port_caches(*dm, mesh);
  std::vector<SparseMatrix<double>> Ms(P);
  for (int i = 0; i < P; ++i) {
    Ms[i] = cached_port_mass(*dm, mesh, ports[i].surface_tag, bc, dir_hash);
    const auto &rp = Ms[i].rowptr();
    for (int r = 0; r < Ms[i].rows(); ++r)
      for (int k = rp[r]; k < rp[r + 1]; ++k) {
        xr.push_back(r);
        xc.push_back(Ms[i].colidx()[k]);
      }
  }
  const std::vector<PortRegion> regions = cached_port_regions(*dm, mesh, ports);
  std::uint64_t sys_key = fnv_bytes(dir_hash, xr.data(), xr.size() * sizeof(int32_t));
  sys_key = fnv_bytes(sys_key, xc.data(), xc.size() * sizeof(int32_t));
  tr.mark("port surface mass + port regions (host)");
  // frequencies are processed in device batches that fit comfortably in HBM
  // upper bound of nnz without building the pattern: 36 triplets per tet + extras
  const double nnz_bound = 36.0 * (double)mesh.tets.size() + (double)xr.size();
  const double bytes_per_matrix = nnz_bound * 16.0 + (double)mesh.edges.size() * 16.0 * P * 10.0;
  const int max_batch = (int)std::max(1.0, std::min((double)F, 24e9 / bytes_per_matrix));
  double device_ms = 0.0;
  if (stats) {
    stats->iterations.assign((size_t)F * P, 0);
    stats->residuals.assign((size_t)F * P, 0.0);
    stats->converged.assign((size_t)F * P, 0);
  }
  efb_ctx *ctx = detail::device_ctx();
  for (int f0 = 0; f0 < F; f0 += max_batch) {
    const int nb = std::min(max_batch, F - f0);
    std::vector<double> omegas(nb);
    for (int f = 0; f < nb; ++f) omegas[f] = 2.0 * M_PI * frequencies[f0 + f];
    // per (frequency, port) propagation constants; an evanescent port voids that frequency
    std::vector<cplx> betas((size_t)nb * P);
    std::vector<char> ok(nb, 1);
    for (int f = 0; f < nb; ++f) {
      const double k0 = omegas[f] / c0;
      for (int i = 0; i < P; ++i) {
        cplx eps_mu(1.0, 0.0);
        if (regions[i].found) eps_mu = p.get_eps_r(regions[i].tag, omegas[f]) * p.get_mu_r(regions[i].tag, omegas[f]);
        const cplx beta_sq = eps_mu * k0 * k0 - ports[i].mode.kc * ports[i].mode.kc;
        if (std::real(beta_sq) <= 0) {
          const double fc = ports[i].mode.kc * c0 / (2.0 * M_PI * std::sqrt(std::abs(std::real(eps_mu))));
          std::cerr << "WARNING: Frequency " << frequencies[f0 + f] / 1e9 << " GHz is below cutoff " << fc / 1e9 << " GHz for port " << i
                    << " — mode is evanescent, S-parameters may be meaningless." << std::endl;
          ok[f] = 0;
          betas[(size_t)f * P + i] = cplx(0.0);
        } else {
          betas[(size_t)f * P + i] = std::sqrt(beta_sq);
        }
      }
    }
    detail::check(efb_timer_start(ctx), "efb_timer_start");
    // the system of the previous call on this mesh with the same pattern, Dirichlet set and shape is reused: every value
    // and right-hand side is rewritten by efb_assemble_volume, the solve starts from x = 0
    auto sys = pool_take(*dm, sys_key, nb, P);
    if (!sys) {
      sys = make_system(*dm, xr, xc, nb, P);
      detail::check(efb_system_set_dirichlet(sys->h, dir.data()), "efb_system_set_dirichlet");
      g_bytes_h2d += (long long)dir.size();
    }
    tr.mark("system (pattern + maps + allocations, or pooled)");
    MaxwellParams pf = p;
    assemble_volume(*sys, *dm, pf, omegas, 0, 0);
    tr.mark("volume assembly launch");
    std::vector<std::unique_ptr<DevicePort>> dp(P);
    for (int i = 0; i < P; ++i) {
      dp[i] = make_port(*sys, ports[i], &Ms[i]);
      detail::check(efb_port_normalize_mass(dp[i]->h, nullptr), "efb_port_normalize_mass");
      std::vector<cplx> coef(nb);
      for (int f = 0; f < nb; ++f) coef[f] = cplx(0.0, p.port_abc_scale) * betas[(size_t)f * P + i];
      detail::check(efb_port_add_mass(sys->h, dp[i]->h, 0, nb, reinterpret_cast<const double *>(coef.data())), "efb_port_add_mass");
    }
    for (int a = 0; a < P; ++a) {  // b[f,a] = 2 j scale beta_a(f) M_s,a e_a : one launch per port
      std::vector<int32_t> idx(nb);
      std::vector<cplx> coef(nb);
      for (int f = 0; f < nb; ++f) {
        idx[f] = f * P + a;
        coef[f] = 2.0 * cplx(0.0, p.port_abc_scale) * betas[(size_t)f * P + a];
      }
      detail::check(efb_port_rhs_batch(sys->h, dp[a]->h, nb, idx.data(), reinterpret_cast<const double *>(coef.data()), 1), "efb_port_rhs_batch");
    }
    tr.mark("port terms + right-hand sides");
    SolveOptions defaults;  // the reference ignores caller options here: solve_linear(A, b, {}) (assemble_maxwell.cpp:755)
    SolveOutcome out = solve_on_device(*sys, 0, nb, defaults, true);
    tr.mark("solve");
    // V[j][f,a] = e_j^H M_s,j x[f,a] for every solution: one launch per port
    std::vector<std::vector<cplx>> V(P, std::vector<cplx>((size_t)nb * P));
    {
      std::vector<int32_t> all((size_t)nb * P);
      for (size_t i = 0; i < all.size(); ++i) all[i] = (int32_t)i;
      for (int j = 0; j < P; ++j)
        detail::check(efb_port_project_batch(sys->h, dp[j]->h, (int32_t)all.size(), all.data(), reinterpret_cast<double *>(V[j].data()), 1),
                      "efb_port_project_batch");
      g_bytes_d2h += (long long)P * (long long)all.size() * 16;
    }
    for (int f = 0; f < nb; ++f) {
      MatrixXcd &Smat = result[f0 + f];
      for (int a = 0; a < P; ++a) {
        const efb_solve_result &r = out.res[(size_t)f * P + a];
        if (stats) {
          stats->iterations[(size_t)(f0 + f) * P + a] = r.iters;
          stats->residuals[(size_t)(f0 + f) * P + a] = r.residual;
          stats->converged[(size_t)(f0 + f) * P + a] = (char)r.converged;
        }
        if (!ok[f]) {
          for (int j = 0; j < P; ++j) Smat(j, a) = kNaN;  // the reference returns an uninitialised S here (:687-698)
          continue;
        }
        if (!r.converged) {
          warn_not_converged(r, out.method, "calculate_sparams_eigenmode (active port " + std::to_string(a) + ")");
          for (int j = 0; j < P; ++j) Smat(j, a) = kNaN;
          continue;
        }
        for (int j = 0; j < P; ++j) {
          const cplx vj = V[j][(size_t)f * P + a];
          Smat(j, a) = (j == a) ? vj - cplx(1.0) : vj;
        }
      }
    }
    double ms = 0.0;
    detail::check(efb_timer_stop(ctx, &ms), "efb_timer_stop");
    device_ms += ms;
    tr.mark("projection + S");
    dp.clear();
    pool_give(*dm, sys_key, std::move(sys));
    tr.mark("release");
  }
  if (stats) {
    stats->device_ms = device_ms;
    stats->kernel_launches = detail::launch_count() - launches0;
    stats->h2d_bytes = g_bytes_h2d - h2d0;
    stats->d2h_bytes = g_bytes_d2h - d2h0;
  }
  return result;
}

MatrixXcd calculate_sparams_eigenmode(const Mesh &mesh, const MaxwellParams &p, const BC &bc, const std::vector<WavePort> &ports) {
  const int P = (int)ports.size();
  if (P == 0) return MatrixXcd(0, 0);
  return calculate_sparams_eigenmode_sweep(mesh, p, bc, ports, {p.omega / (2.0 * M_PI)}, nullptr)[0];
}

// ------------------------------------------------------------------ K/M + sweep
SpMatC KMMatrices::combine(double omega) const {  // src/sweep.cpp:82-88 (identical patterns here)
  const double k0 = omega / c0, k0sq = k0 * k0;
  SpMatC A = K;
  if (K.nonZeros() != M.nonZeros()) throw std::runtime_error("KMMatrices::combine: K and M patterns differ");
  auto &v = A.values();
  const auto &mv = M.values();
  for (size_t i = 0; i < v.size(); ++i) v[i] = v[i] - k0sq * mv[i];
  return A;
}

KMMatrices assemble_maxwell_km(const Mesh &mesh, const MaxwellParams &p, const BC &bc) {  // src/sweep.cpp:90-172
  auto dm = device_mesh_for(mesh);
  const std::vector<uint8_t> dir = dirichlet_flags(mesh, bc);
  std::vector<int32_t> xr, xc;
  for (size_t e = 0; e < dir.size(); ++e)
    if (dir[e]) {
      xr.push_back((int32_t)e);
      xc.push_back((int32_t)e);
    }
  auto sys = make_system(*dm, xr, xc, 2, 1);
  detail::check(efb_system_set_dirichlet(sys->h, dir.data()), "efb_system_set_dirichlet");
  assemble_volume(*sys, *dm, p, {0.0}, 0, 1);
  assemble_volume(*sys, *dm, p, {0.0}, 1, 2);
  KMMatrices km;
  km.K = download_matrix(*sys, 0);
  km.M = download_matrix(*sys, 1);
  return km;
}

SweepResult frequency_sweep(const Mesh &mesh, const MaxwellParams &p, const BC &bc, const std::vector<WavePort> &ports,
                            const std::vector<double> &frequencies, const SolveOptions &opts) {
  SweepResult result;  // src/sweep.cpp:174-349
  result.frequencies = frequencies;
  const int P = (int)ports.size(), F = (int)frequencies.size();
  if (P == 0 || F == 0) return result;
  const bool has_pml = !p.pml_regions.empty() || !p.pml_tensor_regions.empty();
  const bool has_disp = !p.eps_models.empty() || !p.mu_models.empty();
  const bool fast = !has_pml && !has_disp;
  if (opts.verbose)
    std::cerr << "FrequencySweep: " << (fast ? "K/M separation" : "per-frequency assembly") << " on device (" << F << " points, " << P
              << " ports)\n";
  std::vector<double> omegas(F);
  for (int f = 0; f < F; ++f) omegas[f] = 2.0 * M_PI * frequencies[f];
  MaxwellParams q = p;
  if (fast) q.use_port_abc = false;  // the reference's fast path never applies the port ABC (sweep.cpp:255-263)
  // the volume term of all F matrices is assembled in one launch; with static materials the
  // K/M pair is assembled once and combined per frequency (K6) exactly like the reference
  StdSystem S;
  if (fast) {
    S.dm = device_mesh_for(mesh);
    S.dir = dirichlet_flags(mesh, bc);
    std::vector<int32_t> xr, xc, abc_edges;
    standard_extras(mesh, q, ports, S.dir, false, xr, xc, abc_edges);
    S.sys = make_system(*S.dm, xr, xc, F + 2, P);
    detail::check(efb_system_set_dirichlet(S.sys->h, S.dir.data()), "efb_system_set_dirichlet");
    assemble_volume(*S.sys, *S.dm, q, {0.0}, F, 1);
    assemble_volume(*S.sys, *S.dm, q, {0.0}, F + 1, 2);
    std::vector<double> k0sq(F);
    for (int f = 0; f < F; ++f) k0sq[f] = (omegas[f] / c0) * (omegas[f] / c0);
    detail::check(efb_combine_km(S.sys->h, 0, F, k0sq.data(), F, F + 1), "efb_combine_km");
    S.dports.resize(P);
    for (int i = 0; i < P; ++i) {
      if (!port_valid(ports[i])) continue;
      S.symmetric = S.symmetric && weights_phase_aligned(ports[i]);
      S.dports[i] = make_port(*S.sys, ports[i], nullptr);
      std::vector<cplx> coef(F, cplx(1.0) / ports[i].mode.Z0);
      detail::check(efb_port_add_block(S.sys->h, S.dports[i]->h, 0, F, reinterpret_cast<const double *>(coef.data())), "efb_port_add_block");
    }
    if (q.use_abc && !abc_edges.empty()) {
      std::vector<cplx> coef(F);
      for (int f = 0; f < F; ++f) coef[f] = cplx(0.0, omegas[f] / c0);
      detail::check(efb_add_diag(S.sys->h, 0, F, (int32_t)abc_edges.size(), abc_edges.data(), reinterpret_cast<const double *>(coef.data())), "efb_add_diag");
    }
  } else {
    S = build_standard(mesh, q, bc, ports, omegas, P);
  }
  for (int f = 0; f < F; ++f)
    for (int a = 0; a < P; ++a) {
      if (S.dports[a]) {
        add_source(S, ports, a, f * P + a);
      } else if (ports[a].weights.size() == ports[a].edges.size() && ports[a].mode.Z0 != cplx(0.0)) {
        add_source(S, ports, a, f * P + a);
      }
    }
  SolveOutcome out = solve_on_device(*S.sys, 0, F, opts, S.symmetric);
  result.S_matrices.assign(F, MatrixXcd(P, P));
  for (int f = 0; f < F; ++f)
    for (int a = 0; a < P; ++a) {
      const efb_solve_result &r = out.res[(size_t)f * P + a];
      if (!r.converged) {
        for (int j = 0; j < P; ++j) result.S_matrices[f](j, a) = kNaN;
        if (!fast) warn_not_converged(r, out.method, "calculate_sparams (active port " + std::to_string(a) + ")");
        continue;
      }
      fill_sparams_column(result.S_matrices[f], S, ports, a, f * P + a);
    }
  return result;
}

} // namespace edgefem
