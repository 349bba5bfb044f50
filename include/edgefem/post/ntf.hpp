// Near-to-far-field transformation (Stratton-Chu over a Huygens surface).
// Mirrors include/edgefem/post/ntf.hpp:11-100 and src/post/ntf.cpp:40-262 of the reference; the direction x sample
// double loop runs on the GPU (efb_stratton_chu).  The CSV / VTK pattern writers are not provided.
#pragma once
#include <complex>
#include <utility>
#include <vector>

#include "edgefem/linalg.hpp"
#include "edgefem/post/huygens_surface.hpp"

namespace edgefem {

struct NTFPoint2D {
  double theta_deg;
  std::complex<double> e_theta;
  std::complex<double> e_phi;
};

struct FFPattern3D {  // Ntheta x Nphi, theta along rows
  MatrixXd theta_grid, phi_grid;
  MatrixXcd E_theta, E_phi;
  MatrixXd total_magnitude() const;
  MatrixXd power_pattern() const;
  MatrixXd pattern_dB() const;
};

std::vector<NTFPoint2D> stratton_chu_2d(const std::vector<Vector3d> &r, const std::vector<Vector3d> &n, const std::vector<Vector3cd> &E,
                                        const std::vector<Vector3cd> &H, const std::vector<double> &area, const std::vector<double> &theta_rad,
                                        double phi_rad, double k0);

FFPattern3D stratton_chu_3d(const std::vector<Vector3d> &r, const std::vector<Vector3d> &n, const std::vector<Vector3cd> &E,
                            const std::vector<Vector3cd> &H, const std::vector<double> &area, const std::vector<double> &theta_rad,
                            const std::vector<double> &phi_rad, double k0);

double compute_directivity(const FFPattern3D &pattern);
double compute_max_gain(const FFPattern3D &pattern, double efficiency = 1.0);
std::pair<double, double> compute_hpbw(const FFPattern3D &pattern);

} // namespace edgefem
