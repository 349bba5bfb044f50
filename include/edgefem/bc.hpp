// PEC boundary-condition set (reference: include/edgefem/bc.hpp, src/bc.cpp:47-109).
#pragma once
#include <set>
#include <unordered_set>

#include "edgefem/mesh.hpp"

namespace edgefem {

struct BC {
  std::unordered_set<int> dirichlet_nodes;
  std::unordered_set<int> dirichlet_edges;
};

BC build_edge_pec(const Mesh &mesh, int pec_tag);

struct PhysicalTagInfo {
  std::set<int> volume_tags;
  std::set<int> surface_tags;
};
PhysicalTagInfo list_physical_tags(const Mesh &mesh);
bool has_surface_tag(const Mesh &mesh, int tag);
bool has_volume_tag(const Mesh &mesh, int tag);

} // namespace edgefem
