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

} // namespace edgefem
