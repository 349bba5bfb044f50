// Host-side Whitney element matrices (reference: include/edgefem/edge_basis.hpp, src/edge_basis.cpp).
// The volume matrices are computed on the GPU inside assemble_maxwell; these host versions serve
// the port set-up (small, once per port) and user code that calls them directly.
#pragma once
#include <array>

#include "edgefem/linalg.hpp"

namespace edgefem {

using Matrix6d = std::array<std::array<double, 6>, 6>;
using Matrix3d = std::array<std::array<double, 3>, 3>;

Matrix6d whitney_curl_curl_matrix(const std::array<Vector3d, 4> &v);
Matrix6d whitney_mass_matrix(const std::array<Vector3d, 4> &v);
Matrix3d triangle_whitney_mass_matrix(const std::array<Vector3d, 3> &v);

} // namespace edgefem
