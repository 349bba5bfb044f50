// solve_linear (reference: include/edgefem/solver.hpp, src/solver.cpp:35-193).
// Same option/result structs.  The arithmetic runs on the GPU (hand-written COCG / BiCGSTAB with
// Jacobi or auxiliary-space preconditioning) instead of Eigen's BiCGSTAB+IncompleteLUT / SparseLU;
// SolveResult::method names what actually ran (e.g. "B200:COCG+AUX"), residual is the TRUE
// relative residual ||b - A x|| / ||b||.  There is no CPU fallback: auto_fallback re-runs the
// GPU solve with the robust general method (BiCGSTAB) when the first attempt fails.
#pragma once
#include <functional>
#include <string>

#include "edgefem/fem.hpp"

namespace edgefem {

using SolverProgressCallback = std::function<void(int iteration, double residual)>;

struct SolveOptions {
  bool use_bicgstab = true;
  bool use_direct = false; ///< request direct-solver accuracy: iterate to min(tolerance, 1e-12)
  double tolerance = 1e-10;
  int max_iterations = 10000;
  bool use_ilut = true;            ///< true: strongest available preconditioner; false: Jacobi
  double ilut_fill_factor = 10.0;  ///< accepted for source compatibility (no ILUT on the GPU)
  double ilut_drop_tolerance = 1e-4;
  bool auto_fallback = true;
  bool verbose = false;
  SolverProgressCallback progress_callback = nullptr;
  int progress_interval = 100;
};

struct SolveResult {
  std::string method;
  int iters = 0;
  double residual = 0.0;
  bool converged = false;
  std::string error_message;
  VecC x;
};

SolveResult solve_linear(const SpMatC &A, const VecC &b, const SolveOptions &opt = SolveOptions());

} // namespace edgefem
