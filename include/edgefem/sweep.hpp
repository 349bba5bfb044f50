// K/M pre-assembly + frequency sweep (reference: include/edgefem/sweep.hpp, src/sweep.cpp:82-349).
// The toy two-resonator sweep of the reference header is out of scope.
#pragma once
#include <complex>
#include <vector>

#include "edgefem/fem.hpp"
#include "edgefem/maxwell.hpp"
#include "edgefem/mesh.hpp"
#include "edgefem/ports/wave_port.hpp"
#include "edgefem/solver.hpp"

namespace edgefem {

struct KMMatrices {
  SpMatC K;
  SpMatC M;
  SpMatC combine(double omega) const; ///< A = K - k0^2 M  (same pattern for K and M here)
};

KMMatrices assemble_maxwell_km(const Mesh &mesh, const MaxwellParams &p, const BC &bc);

struct SweepResult {
  std::vector<double> frequencies;
  std::vector<MatrixXcd> S_matrices;
};

SweepResult frequency_sweep(const Mesh &mesh, const MaxwellParams &p, const BC &bc, const std::vector<WavePort> &ports,
                            const std::vector<double> &frequencies, const SolveOptions &opts = SolveOptions());

} // namespace edgefem
