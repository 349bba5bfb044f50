// Periodic (Bloch/Floquet) boundary pairs (reference: include/edgefem/periodic.hpp, src/periodic.cpp).
#pragma once
#include <complex>
#include <vector>

#include "edgefem/linalg.hpp"
#include "edgefem/mesh.hpp"

namespace edgefem {

struct PeriodicPair {
  int master_edge;
  int slave_edge;
  int master_orient;
  int slave_orient;
  Vector3d translation;
};

struct PeriodicBC {
  std::vector<PeriodicPair> pairs;
  Vector3d period_vector;
  std::complex<double> phase_shift{1.0, 0.0};
};

PeriodicBC build_periodic_pairs(const Mesh &mesh, int master_tag, int slave_tag, const Vector3d &period_vector,
                                double tolerance = 1e-9);
bool validate_periodic_bc(const Mesh &mesh, const PeriodicBC &pbc);
void set_floquet_phase(PeriodicBC &pbc, const Vector2d &k_transverse);
std::complex<double> floquet_phase_from_angle(const Vector3d &period_vector, double theta, double phi, double k0);
int count_surface_edges(const Mesh &mesh, int surface_tag);

} // namespace edgefem
