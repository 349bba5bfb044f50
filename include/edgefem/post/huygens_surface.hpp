// Huygens-surface extraction from an edge-element solution.
// Mirrors include/edgefem/post/huygens_surface.hpp:15-39 and src/post/huygens_surface.cpp:29-149 of the reference; the
// per-triangle field evaluation runs on the GPU (efb_huygens_eval).
#pragma once
#include <complex>
#include <vector>

#include <array>

#include "edgefem/fem.hpp"
#include "edgefem/linalg.hpp"
#include "edgefem/mesh.hpp"

namespace edgefem {

using Vector3cd = std::array<std::complex<double>, 3>;

struct HuygensSurfaceData {
  std::vector<Vector3d> r;       // triangle centroids
  std::vector<Vector3d> n;       // outward normals (away from the parent tet)
  std::vector<Vector3cd> E_tan;  // tangential E
  std::vector<Vector3cd> H_tan;  // tangential H
  std::vector<double> area;      // triangle areas
};

/// Tangential E and H at the centroid of every triangle tagged surface_tag, from the parent tet's Whitney interpolation
/// and curl.  Parent tet = the LAST tet (highest index) that owns the face, like the reference's map overwrite; triangles
/// with area < 1e-30 or without a parent tet are skipped; throws if nothing is left.
HuygensSurfaceData extract_huygens_surface(const Mesh &mesh, const VecC &solution, int surface_tag, double omega,
                                           std::complex<double> mu_r = 1.0);

} // namespace edgefem
