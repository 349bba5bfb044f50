// Internal helpers shared by the host translation units (not installed).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "edgefem/maxwell.hpp"
#include "edgefem_b200.h"

namespace edgefem {
namespace detail {

bool sym_gen_eig(std::vector<double> K, std::vector<double> M, int n, std::vector<double> &evals, std::vector<double> &evecs);

efb_ctx *device_ctx();                 // lazily created; throws std::runtime_error without a GPU
void check(int rc, const char *what);  // throws with efb_last_error text
void clear_device_cache();
long long launch_count();
efb_mesh *device_mesh_handle(const Mesh &mesh);  // cached device copy of the mesh (valid until the cache drops it)

} // namespace detail
} // namespace edgefem
