// Host-side Whitney element matrices (reference: include/edgefem/edge_basis.hpp, src/edge_basis.cpp).
// The volume matrices are computed on the GPU inside assemble_maxwell; these host versions serve
// the port set-up (small, once per port) and user code that calls them directly.
#pragma once
#include <array>
#include <complex>

#include "edgefem/linalg.hpp"

namespace edgefem {

using Matrix6d = std::array<std::array<double, 6>, 6>;
using Matrix3d = std::array<std::array<double, 3>, 3>;

Matrix6d whitney_curl_curl_matrix(const std::array<Vector3d, 4> &v);
Matrix6d whitney_mass_matrix(const std::array<Vector3d, 4> &v);
Matrix3d triangle_whitney_mass_matrix(const std::array<Vector3d, 3> &v);

// ---- field evaluation (src/edge_basis.cpp:33-46,132-192) -------------------------------------------
/// rows = curl of the six Whitney functions, 2 grad(lambda_a) x grad(lambda_b) (constant per tet)
std::array<Vector3d, 6> whitney_edge_curls(const std::array<Vector3d, 4> &v);
/// barycentric coordinates of a point w.r.t. the tet's vertices
std::array<double, 4> compute_barycentric(const std::array<Vector3d, 4> &v, const Vector3d &p);
Vector3d compute_grad_lambda(const std::array<Vector3d, 4> &v, int i);
/// E(p) = sum_e dof_e * orient_e * (lambda_a grad lambda_b - lambda_b grad lambda_a)
std::array<std::complex<double>, 3> evaluate_edge_field(const std::array<Vector3d, 4> &vertices, const std::array<int, 6> &edge_orient,
                                                        const std::array<std::complex<double>, 6> &edge_dofs, const Vector3d &point);

} // namespace edgefem
