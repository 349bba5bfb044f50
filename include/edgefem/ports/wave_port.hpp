// Wave ports (reference: include/edgefem/ports/wave_port.hpp, src/ports/wave_port.cpp:410-585).
#pragma once
#include <unordered_set>
#include <vector>

#include "edgefem/linalg.hpp"
#include "edgefem/mesh.hpp"
#include "edgefem/ports/port_eigensolve.hpp"

namespace edgefem {

/// Port face as a stand-alone 2-D mesh (z dropped), src/ports/wave_port.cpp:54-112.
struct PortSurfaceMesh {
  Mesh mesh;
  std::vector<int> volume_tri_indices;  // index into the volume mesh's tris
};

PortSurfaceMesh extract_surface_mesh(const Mesh &volume_mesh, int surface_tag);

struct WavePort {
  int surface_tag = 0;
  PortMode mode;
  std::vector<int> edges;
  VectorXcd weights;
};

/// 2-D discrete TE port mode on the port face: K_s e = kc^2 M_s e on the free port edges,
/// eigenvector closest to target_kc_sq (gradient null space skipped).  Global edge indexing.
/// Edge weights = line integrals of the modal tangential E-field along the port edges (src/ports/wave_port.cpp:114-212).
WavePort build_wave_port(const Mesh &volume_mesh, const PortSurfaceMesh &surface, const PortMode &mode);

/// Analytic TE10 scalar field cos(pi (x - x_min)/a), scaled to unit power (src/ports/wave_port.cpp:214-262).
void populate_te10_field(const PortSurfaceMesh &surface, const RectWaveguidePort &port, PortMode &mode);

/// Port whose weights are j * (3-D eigenvector restricted to the port edges), ||w||^2 = sqrt(Re Z0)
/// (src/ports/wave_port.cpp:264-309).
WavePort build_wave_port_from_eigenvector(const Mesh &volume_mesh, const PortSurfaceMesh &surface, const VectorXd &eigenvector,
                                          const PortMode &mode, const std::unordered_set<int> &pec_edges);

VectorXd solve_port_mode_2d(const Mesh &mesh, int surface_tag, const std::unordered_set<int> &pec_edges,
                            double target_kc_sq, double &kc_sq_out);

WavePort build_wave_port_2d(const Mesh &mesh, int surface_tag, const PortMode &mode,
                            const std::unordered_set<int> &pec_edges, double target_kc_sq);

/// Port surface mass matrix M_s (real, m x m, PEC rows/cols omitted).
SparseMatrix<double> assemble_port_surface_mass(const Mesh &mesh, int surface_tag,
                                                const std::unordered_set<int> &dirichlet_edges);

} // namespace edgefem
