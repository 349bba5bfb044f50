// Touchstone writers -- the wire format after the S-parameter sweep.
// Mirrors include/edgefem/io/touchstone.hpp:13-47 and src/io/touchstone.cpp:48-149 of the reference
// (12 significant digits; N-port data row-major S(i,j), 4 complex values per line for N > 2).
#pragma once
#include <string>
#include <vector>

#include "edgefem/linalg.hpp"
#include "edgefem/ports/port_eigensolve.hpp"

namespace edgefem {

enum class TouchstoneFormat { RI, MA, DB };

struct TouchstoneOptions {
  TouchstoneFormat format = TouchstoneFormat::RI;
  double z0 = 50.0;
};

void write_touchstone(const std::string &path, const std::vector<double> &freq, const std::vector<SParams2> &data);

void write_touchstone_nport(const std::string &path, const std::vector<double> &freq, const std::vector<MatrixXcd> &S_matrices,
                            const TouchstoneOptions &opts = TouchstoneOptions());

std::string touchstone_extension(int num_ports);

} // namespace edgefem
