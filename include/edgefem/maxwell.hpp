// Maxwell assembly + S-parameter drivers (reference: include/edgefem/maxwell.hpp,
// src/assemble_maxwell.cpp).  Same structs, names, argument order and defaults; every function
// below runs its element integration, boundary terms, solve and projection on the GPU through
// the C-ABI in include/edgefem_b200.h.
#pragma once
#include <complex>
#include <memory>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "edgefem/bc.hpp"
#include "edgefem/fem.hpp"
#include "edgefem/materials/dispersive.hpp"
#include "edgefem/mesh.hpp"
#include "edgefem/periodic.hpp"
#include "edgefem/ports/port_eigensolve.hpp"
#include "edgefem/ports/wave_port.hpp"
#include "edgefem/solver.hpp"

namespace edgefem {

struct PMLRegionSpec {
  Vector3d sigma_max = Vector3d::Zero();
  Vector3d thickness = Vector3d::Zero();
  double grading_order = 3.0;
};

struct PMLDiagnostic {
  int region_tag = 0;
  Vector3d sigma_max = Vector3d::Zero();
  Vector3d thickness = Vector3d::Zero();
  Vector3d reflection_est = Vector3d::Ones();
};

enum class PortABCType { None, Beta, BetaNorm, ImpedanceMatch, ModalAdmittance };

struct MaxwellParams {
  double omega = 0.0;
  std::complex<double> eps_r = 1.0;
  std::complex<double> mu_r = 1.0;
  std::unordered_map<int, std::complex<double>> eps_r_regions;
  std::unordered_map<int, std::complex<double>> mu_r_regions;
  std::unordered_map<int, std::shared_ptr<materials::DispersiveMaterial>> eps_models;
  std::unordered_map<int, std::shared_ptr<materials::DispersiveMaterial>> mu_models;
  double pml_sigma = 0.0;
  std::unordered_set<int> pml_regions;
  std::unordered_map<int, PMLRegionSpec> pml_tensor_regions;
  bool enforce_pml_heuristics = true;
  bool use_abc = false;
  std::unordered_set<int> abc_surface_tags;
  bool use_port_abc = false;
  PortABCType port_abc_type = PortABCType::Beta;
  double port_weight_scale = 1.0;
  double port_abc_scale = 1.0;
  bool use_eigenmode_excitation = false;

  std::complex<double> get_eps_r(int phys_tag) const {
    auto it = eps_r_regions.find(phys_tag);
    return (it != eps_r_regions.end()) ? it->second : eps_r;
  }
  std::complex<double> get_eps_r(int phys_tag, double w) const {
    auto m = eps_models.find(phys_tag);
    if (m != eps_models.end() && m->second) return m->second->eval_eps(w);
    return get_eps_r(phys_tag);
  }
  std::complex<double> get_mu_r(int phys_tag) const {
    auto it = mu_r_regions.find(phys_tag);
    return (it != mu_r_regions.end()) ? it->second : mu_r;
  }
  std::complex<double> get_mu_r(int phys_tag, double w) const {
    auto m = mu_models.find(phys_tag);
    if (m != mu_models.end() && m->second) return m->second->eval_mu(w);
    return get_mu_r(phys_tag);
  }
};

struct MaxwellAssembly {
  SpMatC A;
  VecC b;
  std::vector<PMLDiagnostic> diagnostics;
};

MaxwellAssembly assemble_maxwell(const Mesh &mesh, const MaxwellParams &p, const BC &bc,
                                 const std::vector<WavePort> &ports, int active_port_idx = -1);

MaxwellAssembly assemble_maxwell_periodic(const Mesh &mesh, const MaxwellParams &p, const BC &bc,
                                          const PeriodicBC &pbc, const std::vector<WavePort> &ports,
                                          int active_port_idx = -1);

MatrixXcd calculate_sparams(const Mesh &mesh, const MaxwellParams &p, const BC &bc,
                            const std::vector<WavePort> &ports, const SolveOptions &opts = SolveOptions());

void normalize_port_weights(const Mesh &mesh, const MaxwellParams &p, const BC &bc, std::vector<WavePort> &ports,
                            const SolveOptions &opts = SolveOptions());

MatrixXcd calculate_sparams_periodic(const Mesh &mesh, const MaxwellParams &p, const BC &bc, const PeriodicBC &pbc,
                                     const std::vector<WavePort> &ports);

MatrixXcd calculate_sparams_eigenmode(const Mesh &mesh, const MaxwellParams &p, const BC &bc,
                                      const std::vector<WavePort> &ports);

// ---- extension (not in the reference): the same eigenmode S-parameter computation for a whole
// list of frequencies, assembled and solved as ONE device batch (p.omega is ignored).
// Returns one P x P matrix per frequency; solver statistics are optional.
struct BatchStats {
  std::vector<int> iterations;   // per (frequency, active port)
  std::vector<double> residuals; // true relative residuals
  std::vector<char> converged;
  double device_ms = 0.0;        // CUDA-event time of the device work
  long long kernel_launches = 0;
  long long h2d_bytes = 0, d2h_bytes = 0;
};
std::vector<MatrixXcd> calculate_sparams_eigenmode_sweep(const Mesh &mesh, const MaxwellParams &p, const BC &bc,
                                                         const std::vector<WavePort> &ports,
                                                         const std::vector<double> &frequencies,
                                                         BatchStats *stats = nullptr);

} // namespace edgefem
