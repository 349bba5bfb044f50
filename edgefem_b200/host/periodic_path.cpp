// Periodic (Bloch) assembly + S-parameters on the device (reference: src/assemble_maxwell.cpp:496-635).
#include <stdexcept>

#include "edgefem/maxwell.hpp"
#include "host_internal.hpp"

namespace edgefem {

MaxwellAssembly assemble_maxwell_periodic(const Mesh &, const MaxwellParams &, const BC &, const PeriodicBC &, const std::vector<WavePort> &, int) {
  throw std::runtime_error("assemble_maxwell_periodic: not implemented yet in the B200 build");
}

MatrixXcd calculate_sparams_periodic(const Mesh &, const MaxwellParams &, const BC &, const PeriodicBC &, const std::vector<WavePort> &) {
  throw std::runtime_error("calculate_sparams_periodic: not implemented yet in the B200 build");
}

} // namespace edgefem
