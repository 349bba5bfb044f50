// Port mode description + analytic TE10 (reference: include/edgefem/ports/port_eigensolve.hpp,
// src/ports/port_eigensolve.cpp:45-63).  The nodal 2-D Laplace eigen-solve (solve_port_eigens)
// is host set-up that the reference keeps dense; it is a "next" row.
#pragma once
#include <complex>
#include <vector>

#include "edgefem/linalg.hpp"
#include "edgefem/mesh.hpp"

namespace edgefem {

struct RectWaveguidePort {
  double a;
  double b;
};

enum class ModePolarization { TE, TM };

struct PortMode {
  ModePolarization pol = ModePolarization::TE;
  double fc = 0.0;
  double kc = 0.0;
  double omega = 0.0;
  std::complex<double> eps = 0.0;
  std::complex<double> mu = 0.0;
  std::complex<double> beta = 0.0;
  std::complex<double> Z0 = 0.0;
  VectorXcd field;
};

PortMode solve_te10_mode(const RectWaveguidePort &port, double freq);

/// Legacy 2-port container (include/edgefem/ports/port_eigensolve.hpp:36-42 of the reference).
struct SParams2 {
  std::complex<double> s11, s21, s12, s22;
};

/// Nodal (P1) scalar Helmholtz eigenmodes of a 2-D port cross-section mesh (tris in the xy-plane):
/// -lap(phi) = kc^2 phi with natural boundary (TE) or phi = 0 on boundary_lines of phys 1 (TM).  Modes with
/// kc^2 < 1e-12 are skipped, fields are scaled to unit modal power, at most num_modes ascending by cutoff.
/// Mirrors src/ports/port_eigensolve.cpp:97-275.  Dense solve.  The eigenvector SIGN is the solver's choice in the
/// reference; here the field is signed so that its largest-magnitude sample is positive.
std::vector<PortMode> solve_port_eigens(const Mesh &mesh, int num_modes, double omega, std::complex<double> eps_r,
                                        std::complex<double> mu_r, ModePolarization pol);

/// Analytic S-parameters of a straight rectangular guide (src/ports/port_eigensolve.cpp:67-88).
SParams2 straight_waveguide_sparams(const RectWaveguidePort &port, double length, double freq);

} // namespace edgefem
