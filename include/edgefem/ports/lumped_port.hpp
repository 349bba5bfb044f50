// Lumped ports (reference: include/edgefem/ports/lumped_port.hpp, src/ports/lumped_port.cpp:21-160).
#pragma once
#include "edgefem/linalg.hpp"
#include "edgefem/mesh.hpp"
#include "edgefem/ports/wave_port.hpp"

namespace edgefem {

enum class LumpedPortWeightMode { Projection, SurfaceIntegral };

struct LumpedPortConfig {
  int surface_tag = 0;
  double z0 = 50.0;
  Vector3d e_direction = {0, 0, 1};
  LumpedPortWeightMode weight_mode = LumpedPortWeightMode::SurfaceIntegral;
};

WavePort build_lumped_port(const Mesh &mesh, const LumpedPortConfig &config);

} // namespace edgefem
