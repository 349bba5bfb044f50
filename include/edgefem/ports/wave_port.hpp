// Wave ports (reference: include/edgefem/ports/wave_port.hpp, src/ports/wave_port.cpp:410-585).
#pragma once
#include <unordered_set>
#include <vector>

#include "edgefem/linalg.hpp"
#include "edgefem/mesh.hpp"
#include "edgefem/ports/port_eigensolve.hpp"

namespace edgefem {

struct WavePort {
  int surface_tag = 0;
  PortMode mode;
  std::vector<int> edges;
  VectorXcd weights;
};

/// 2-D discrete TE port mode on the port face: K_s e = kc^2 M_s e on the free port edges,
/// eigenvector closest to target_kc_sq (gradient null space skipped).  Global edge indexing.
VectorXd solve_port_mode_2d(const Mesh &mesh, int surface_tag, const std::unordered_set<int> &pec_edges,
                            double target_kc_sq, double &kc_sq_out);

WavePort build_wave_port_2d(const Mesh &mesh, int surface_tag, const PortMode &mode,
                            const std::unordered_set<int> &pec_edges, double target_kc_sq);

/// Port surface mass matrix M_s (real, m x m, PEC rows/cols omitted).
SparseMatrix<double> assemble_port_surface_mass(const Mesh &mesh, int surface_tag,
                                                const std::unordered_set<int> &dirichlet_edges);

} // namespace edgefem
