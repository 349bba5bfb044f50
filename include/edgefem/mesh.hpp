// Mesh container + Gmsh v2 reader + global edge numbering.
// Mirrors include/edgefem/mesh.hpp:13-60 and src/mesh_gmsh.cpp:15-146 of the reference.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "edgefem/linalg.hpp"

namespace edgefem {

struct Node {
  std::int64_t id;
  Vector3d xyz;
};

struct Edge {
  std::int64_t n0;
  std::int64_t n1;
};

struct BoundaryLine {
  std::int64_t n0;
  std::int64_t n1;
  int phys = 0;
};

enum class ElemType : int { Tri3 = 2, Tet4 = 4 };

struct Element {
  std::int64_t id = 0;
  ElemType type = ElemType::Tet4;
  std::array<std::int64_t, 4> conn{};
  int phys = 0;
  std::array<int, 6> edges{};
  std::array<int, 6> edge_orient{};
};

struct Mesh {
  std::vector<Node> nodes;
  std::vector<Element> tets;
  std::vector<Element> tris;
  std::vector<BoundaryLine> boundary_lines;
  std::vector<Edge> edges;
  std::unordered_map<std::uint64_t, int> edgeIndex;
  std::unordered_map<std::int64_t, int> nodeIndex;
};

Mesh load_gmsh_v2(const std::string &path);

inline std::uint64_t make_edge_key(std::int64_t a, std::int64_t b) {
  if (a > b) std::swap(a, b);
  return (static_cast<std::uint64_t>(a) << 32) ^ static_cast<std::uint64_t>(b);
}

// ---- extensions (not in the reference) ------------------------------------------------
/// Global edge numbering exactly as load_gmsh_v2 applies it: first-seen order over tets
/// (local pairs 01,02,03,12,13,23) then tris (01,12,20); orient=+1 iff conn[a] < conn[b].
/// Fills Element::edges / edge_orient, Mesh::edges and Mesh::edgeIndex.
void build_edges(Mesh &mesh);

/// Build a Mesh from flat arrays (node ids 1..n when node_ids is empty) and number its edges.
Mesh mesh_from_arrays(const std::vector<double> &xyz, const std::vector<std::int64_t> &tet_conn,
                      const std::vector<int> &tet_phys, const std::vector<std::int64_t> &tri_conn,
                      const std::vector<int> &tri_phys, const std::vector<std::int64_t> &node_ids = {});

} // namespace edgefem
