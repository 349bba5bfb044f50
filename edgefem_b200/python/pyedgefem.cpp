// pybind11 module `pyedgefem` for the B200 build: the same flat names, keyword arguments and
// defaults as the reference's python/pyedgefem.cpp for the frequency-domain hot path
// (mesh + BC + ports + assemble_maxwell / solve_linear / calculate_sparams* / frequency_sweep /
// periodic + the materials submodule).  Out-of-scope subsystems of the reference module
// (far field, coupling, exporters, scalar stub) are not provided.  Eigen return types become
// numpy arrays via the casters below.
#include <pybind11/complex.h>
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "edgefem/bc.hpp"
#include "edgefem/edge_basis.hpp"
#include "edgefem/io/touchstone.hpp"
#include "edgefem/maxwell.hpp"
#include "edgefem/mesh.hpp"
#include "edgefem/periodic.hpp"
#include "edgefem/post/ntf.hpp"
#include "edgefem/ports/lumped_port.hpp"
#include "edgefem/ports/wave_port.hpp"
#include "edgefem/solver.hpp"
#include "edgefem/sweep.hpp"
#include "host_internal.hpp"

namespace py = pybind11;
using namespace edgefem;

namespace pybind11 {
namespace detail {
template <>
struct type_caster<Vector3d> {
  PYBIND11_TYPE_CASTER(Vector3d, const_name("Vector3d"));
  bool load(handle src, bool) {
    if (!src) return false;
    try {
      auto seq = py::cast<std::vector<double>>(src);
      if (seq.size() != 3) return false;
      value = Vector3d(seq[0], seq[1], seq[2]);
      return true;
    } catch (...) {
      return false;
    }
  }
  static handle cast(const Vector3d &v, return_value_policy, handle) {
    py::array_t<double> a(3);
    auto r = a.mutable_unchecked<1>();
    for (int i = 0; i < 3; ++i) r(i) = v[i];
    return a.release();
  }
};
template <>
struct type_caster<Vector2d> {
  PYBIND11_TYPE_CASTER(Vector2d, const_name("Vector2d"));
  bool load(handle src, bool) {
    try {
      auto seq = py::cast<std::vector<double>>(src);
      if (seq.size() != 2) return false;
      value = Vector2d(seq[0], seq[1]);
      return true;
    } catch (...) {
      return false;
    }
  }
  static handle cast(const Vector2d &v, return_value_policy, handle) { return py::make_tuple(v.x(), v.y()).release(); }
};
template <>
struct type_caster<MatrixXcd> {
  PYBIND11_TYPE_CASTER(MatrixXcd, const_name("numpy.ndarray[complex128[m, n]]"));
  bool load(handle src, bool) {
    auto a = py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast>::ensure(src);
    if (!a || a.ndim() != 2) return false;
    value.resize((int)a.shape(0), (int)a.shape(1));
    auto r = a.unchecked<2>();
    for (int i = 0; i < value.rows(); ++i)
      for (int j = 0; j < value.cols(); ++j) value(i, j) = r(i, j);
    return true;
  }
  static handle cast(const MatrixXcd &M, return_value_policy, handle) {
    py::array_t<std::complex<double>> a({M.rows(), M.cols()});
    auto r = a.mutable_unchecked<2>();
    for (int i = 0; i < M.rows(); ++i)
      for (int j = 0; j < M.cols(); ++j) r(i, j) = M(i, j);
    return a.release();
  }
};
template <>
struct type_caster<VectorXd> {
  PYBIND11_TYPE_CASTER(VectorXd, const_name("numpy.ndarray[float64[n]]"));
  bool load(handle src, bool) {
    auto a = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(src);
    if (!a || a.ndim() != 1) return false;
    value.resize((size_t)a.shape(0));
    auto r = a.unchecked<1>();
    for (py::ssize_t i = 0; i < a.shape(0); ++i) value[(size_t)i] = r(i);
    return true;
  }
  static handle cast(const VectorXd &v, return_value_policy, handle) {
    py::array_t<double> a((py::ssize_t)v.size());
    auto r = a.mutable_unchecked<1>();
    for (size_t i = 0; i < v.size(); ++i) r((py::ssize_t)i) = v[i];
    return a.release();
  }
};
} // namespace detail
} // namespace pybind11

namespace {

py::array_t<std::complex<double>> vec_to_numpy(const VecC &v) {
  py::array_t<std::complex<double>> a((py::ssize_t)v.size());
  std::memcpy(a.mutable_data(), v.data(), v.size() * sizeof(cplx));
  return a;
}

VecC vec_from_numpy(const py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> &a) {
  if (a.ndim() != 1) throw std::invalid_argument("VecC: expected a 1-D array");
  VecC v((size_t)a.shape(0));
  std::memcpy(v.data(), a.data(), v.size() * sizeof(cplx));
  return v;
}

template <typename T>
py::tuple csr_arrays(const SparseMatrix<T> &A) {
  py::array_t<int> rp((py::ssize_t)A.rowptr().size()), ci((py::ssize_t)A.colidx().size());
  py::array_t<T> va((py::ssize_t)A.values().size());
  std::memcpy(rp.mutable_data(), A.rowptr().data(), A.rowptr().size() * sizeof(int));
  if (!A.colidx().empty()) std::memcpy(ci.mutable_data(), A.colidx().data(), A.colidx().size() * sizeof(int));
  if (!A.values().empty()) std::memcpy(va.mutable_data(), A.values().data(), A.values().size() * sizeof(T));
  return py::make_tuple(rp, ci, va);
}

template <typename T>
py::array_t<T> to_dense(const SparseMatrix<T> &A) {
  py::array_t<T> d({A.rows(), A.cols()});
  std::fill(d.mutable_data(), d.mutable_data() + (size_t)A.rows() * A.cols(), T(0));
  auto r = d.template mutable_unchecked<2>();
  for (int i = 0; i < A.rows(); ++i)
    for (int k = A.rowptr()[i]; k < A.rowptr()[i + 1]; ++k) r(i, A.colidx()[k]) = A.values()[k];
  return d;
}

} // namespace

// ---- f3 helpers: [n,3] numpy arrays <-> vectors of 3-vectors
static py::array_t<double> v3_to_numpy(const std::vector<Vector3d> &v) {
  py::array_t<double> a({(py::ssize_t)v.size(), (py::ssize_t)3});
  auto r = a.mutable_unchecked<2>();
  for (size_t i = 0; i < v.size(); ++i)
    for (int k = 0; k < 3; ++k) r(i, k) = v[i][k];
  return a;
}
static py::array_t<std::complex<double>> c3_to_numpy(const std::vector<Vector3cd> &v) {
  py::array_t<std::complex<double>> a({(py::ssize_t)v.size(), (py::ssize_t)3});
  auto r = a.mutable_unchecked<2>();
  for (size_t i = 0; i < v.size(); ++i)
    for (int k = 0; k < 3; ++k) r(i, k) = v[i][k];
  return a;
}
static std::vector<Vector3d> v3_from_numpy(py::array_t<double, py::array::c_style | py::array::forcecast> a) {
  if (a.ndim() != 2 || a.shape(1) != 3) throw std::invalid_argument("expected an [n,3] float array");
  std::vector<Vector3d> v((size_t)a.shape(0));
  auto r = a.unchecked<2>();
  for (size_t i = 0; i < v.size(); ++i) v[i] = Vector3d(r(i, 0), r(i, 1), r(i, 2));
  return v;
}
static std::vector<Vector3cd> c3_from_numpy(py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> a) {
  if (a.ndim() != 2 || a.shape(1) != 3) throw std::invalid_argument("expected an [n,3] complex array");
  std::vector<Vector3cd> v((size_t)a.shape(0));
  auto r = a.unchecked<2>();
  for (size_t i = 0; i < v.size(); ++i) v[i] = {r(i, 0), r(i, 1), r(i, 2)};
  return v;
}
static py::array_t<double> md_to_numpy(const MatrixXd &M) {
  py::array_t<double> a({(py::ssize_t)M.rows(), (py::ssize_t)M.cols()});
  auto r = a.mutable_unchecked<2>();
  for (int i = 0; i < M.rows(); ++i)
    for (int j = 0; j < M.cols(); ++j) r(i, j) = M(i, j);
  return a;
}

PYBIND11_MODULE(pyedgefem, m) {
  m.doc() = "EdgeFEM frequency-domain hot path on NVIDIA B200 (sm_100a): drop-in subset of the reference pyedgefem module";
  m.attr("backend") = "b200";

  // ---------------------------------------------------------------- mesh
  py::class_<Node>(m, "Node").def(py::init<>()).def_readonly("id", &Node::id).def_readonly("xyz", &Node::xyz);
  py::enum_<ElemType>(m, "ElemType").value("Tri3", ElemType::Tri3).value("Tet4", ElemType::Tet4);
  py::class_<Element>(m, "Element")
      .def(py::init<>())
      .def_readonly("id", &Element::id)
      .def_readonly("type", &Element::type)
      .def_readonly("conn", &Element::conn)
      .def_readonly("phys", &Element::phys)
      .def_readonly("edges", &Element::edges)
      .def_readonly("edge_orient", &Element::edge_orient);
  py::class_<Edge>(m, "Edge").def(py::init<>()).def_readonly("n0", &Edge::n0).def_readonly("n1", &Edge::n1);
  py::class_<Mesh>(m, "Mesh", "Represents a 3D mesh.")
      .def(py::init<>())
      .def_readonly("nodes", &Mesh::nodes)
      .def_readonly("tets", &Mesh::tets)
      .def_readonly("tris", &Mesh::tris)
      .def_readonly("edges", &Mesh::edges)
      .def_readonly("nodeIndex", &Mesh::nodeIndex)
      .def("num_nodes", [](const Mesh &me) { return me.nodes.size(); })
      .def("num_tets", [](const Mesh &me) { return me.tets.size(); })
      .def("num_tris", [](const Mesh &me) { return me.tris.size(); })
      .def("num_edges", [](const Mesh &me) { return me.edges.size(); })
      // bulk numpy views (extension; the per-object lists above are O(N) Python objects)
      .def("boundary_lines_array",
           [](const Mesh &me) {  // [l,3] = (n0, n1, phys); filled by extract_surface_mesh
             py::array_t<std::int64_t> a({(py::ssize_t)me.boundary_lines.size(), (py::ssize_t)3});
             auto r = a.mutable_unchecked<2>();
             for (size_t i = 0; i < me.boundary_lines.size(); ++i) {
               r(i, 0) = me.boundary_lines[i].n0;
               r(i, 1) = me.boundary_lines[i].n1;
               r(i, 2) = me.boundary_lines[i].phys;
             }
             return a;
           })
      .def("tet_edges_array",
           [](const Mesh &me) {
             py::array_t<int> a({(py::ssize_t)me.tets.size(), (py::ssize_t)6});
             auto r = a.mutable_unchecked<2>();
             for (size_t t = 0; t < me.tets.size(); ++t)
               for (int k = 0; k < 6; ++k) r(t, k) = me.tets[t].edges[k];
             return a;
           })
      .def("tet_orient_array",
           [](const Mesh &me) {
             py::array_t<int> a({(py::ssize_t)me.tets.size(), (py::ssize_t)6});
             auto r = a.mutable_unchecked<2>();
             for (size_t t = 0; t < me.tets.size(); ++t)
               for (int k = 0; k < 6; ++k) r(t, k) = me.tets[t].edge_orient[k];
             return a;
           })
      .def("tri_edges_array",
           [](const Mesh &me) {
             py::array_t<int> a({(py::ssize_t)me.tris.size(), (py::ssize_t)3});
             auto r = a.mutable_unchecked<2>();
             for (size_t t = 0; t < me.tris.size(); ++t)
               for (int k = 0; k < 3; ++k) r(t, k) = me.tris[t].edges[k];
             return a;
           })
      .def("tri_orient_array",
           [](const Mesh &me) {
             py::array_t<int> a({(py::ssize_t)me.tris.size(), (py::ssize_t)3});
             auto r = a.mutable_unchecked<2>();
             for (size_t t = 0; t < me.tris.size(); ++t)
               for (int k = 0; k < 3; ++k) r(t, k) = me.tris[t].edge_orient[k];
             return a;
           })
      .def("xyz_array",
           [](const Mesh &me) {
             py::array_t<double> a({(py::ssize_t)me.nodes.size(), (py::ssize_t)3});
             auto r = a.mutable_unchecked<2>();
             for (size_t i = 0; i < me.nodes.size(); ++i)
               for (int k = 0; k < 3; ++k) r(i, k) = me.nodes[i].xyz[k];
             return a;
           })
      .def("tet_nodes_array",
           [](const Mesh &me) {  // node INDICES (not ids)
             py::array_t<int> a({(py::ssize_t)me.tets.size(), (py::ssize_t)4});
             auto r = a.mutable_unchecked<2>();
             for (size_t t = 0; t < me.tets.size(); ++t)
               for (int k = 0; k < 4; ++k) r(t, k) = me.nodeIndex.at(me.tets[t].conn[k]);
             return a;
           })
      .def("tet_phys_array",
           [](const Mesh &me) {
             py::array_t<int> a((py::ssize_t)me.tets.size());
             auto r = a.mutable_unchecked<1>();
             for (size_t t = 0; t < me.tets.size(); ++t) r(t) = me.tets[t].phys;
             return a;
           })
      .def("edge_nodes_array",
           [](const Mesh &me) {  // node INDICES of each global edge
             py::array_t<int> a({(py::ssize_t)me.edges.size(), (py::ssize_t)2});
             auto r = a.mutable_unchecked<2>();
             for (size_t e = 0; e < me.edges.size(); ++e) {
               r(e, 0) = me.nodeIndex.at(me.edges[e].n0);
               r(e, 1) = me.nodeIndex.at(me.edges[e].n1);
             }
             return a;
           })
      .def("edges_array", [](const Mesh &me) {
        py::array_t<std::int64_t> a({(py::ssize_t)me.edges.size(), (py::ssize_t)2});
        auto r = a.mutable_unchecked<2>();
        for (size_t e = 0; e < me.edges.size(); ++e) {
          r(e, 0) = me.edges[e].n0;
          r(e, 1) = me.edges[e].n1;
        }
        return a;
      });
  m.def("load_gmsh", &load_gmsh_v2, "Loads a mesh from a Gmsh v2 .msh file.");
  m.def(
      "mesh_from_arrays",
      [](py::array_t<double, py::array::c_style | py::array::forcecast> xyz,
         py::array_t<std::int64_t, py::array::c_style | py::array::forcecast> tets,
         py::array_t<int, py::array::c_style | py::array::forcecast> tet_phys,
         py::array_t<std::int64_t, py::array::c_style | py::array::forcecast> tris,
         py::array_t<int, py::array::c_style | py::array::forcecast> tri_phys, py::object node_ids) {
        std::vector<std::int64_t> ids;
        if (!node_ids.is_none()) ids = py::cast<std::vector<std::int64_t>>(node_ids);
        return mesh_from_arrays(std::vector<double>(xyz.data(), xyz.data() + xyz.size()),
                                std::vector<std::int64_t>(tets.data(), tets.data() + tets.size()),
                                std::vector<int>(tet_phys.data(), tet_phys.data() + tet_phys.size()),
                                std::vector<std::int64_t>(tris.data(), tris.data() + tris.size()),
                                std::vector<int>(tri_phys.data(), tri_phys.data() + tri_phys.size()), ids);
      },
      "Build a mesh from arrays (extension): numbers edges exactly like load_gmsh.", py::arg("xyz"), py::arg("tets"),
      py::arg("tet_phys"), py::arg("tris"), py::arg("tri_phys"), py::arg("node_ids") = py::none());

  // ---------------------------------------------------------------- BC
  py::class_<BC>(m, "BC", "Boundary condition definitions.")
      .def(py::init<>())
      .def_property_readonly("dirichlet_nodes", [](const BC &b) { return std::vector<int>(b.dirichlet_nodes.begin(), b.dirichlet_nodes.end()); })
      .def_property_readonly("dirichlet_edges", [](const BC &b) { return std::vector<int>(b.dirichlet_edges.begin(), b.dirichlet_edges.end()); })
      .def("add_pec_edge", [](BC &b, int edge) { b.dirichlet_edges.insert(edge); }, py::arg("edge"))
      .def("merge",
           [](BC &b, const BC &o) {
             b.dirichlet_edges.insert(o.dirichlet_edges.begin(), o.dirichlet_edges.end());
             b.dirichlet_nodes.insert(o.dirichlet_nodes.begin(), o.dirichlet_nodes.end());
           },
           py::arg("other"));
  m.def("build_edge_pec", &build_edge_pec, py::arg("mesh"), py::arg("pec_tag"));
  py::class_<PhysicalTagInfo>(m, "PhysicalTagInfo")
      .def_readonly("volume_tags", &PhysicalTagInfo::volume_tags)
      .def_readonly("surface_tags", &PhysicalTagInfo::surface_tags);
  m.def("list_physical_tags", &list_physical_tags, py::arg("mesh"));
  m.def("has_surface_tag", &has_surface_tag, py::arg("mesh"), py::arg("tag"));
  m.def("has_volume_tag", &has_volume_tag, py::arg("mesh"), py::arg("tag"));

  // ---------------------------------------------------------------- opaque linear algebra
  py::class_<SpMatC>(m, "SpMatC", "Sparse matrix with complex coefficients.")
      .def(py::init<>())
      .def_property_readonly("shape", [](const SpMatC &A) { return py::make_tuple(A.rows(), A.cols()); })
      .def("coeff", &SpMatC::coeff, py::arg("i"), py::arg("j"))
      .def("to_dense", [](const SpMatC &A) { return to_dense(A); })
      .def("nnz", [](const SpMatC &A) { return A.nonZeros(); })
      .def("to_csr", [](const SpMatC &A) { return csr_arrays(A); }, "(rowptr, colidx, values) numpy arrays (extension)")
      .def_static(
          "from_csr",
          [](int n, py::array_t<int, py::array::c_style | py::array::forcecast> rp,
             py::array_t<int, py::array::c_style | py::array::forcecast> ci,
             py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> va) {
            SpMatC A(n, n);
            A.rowptr().assign(rp.data(), rp.data() + rp.size());
            A.colidx().assign(ci.data(), ci.data() + ci.size());
            A.values().assign(va.data(), va.data() + va.size());
            if ((int)A.rowptr().size() != n + 1 || A.colidx().size() != A.values().size() || A.rowptr().back() != (int)A.colidx().size())
              throw std::invalid_argument("SpMatC.from_csr: inconsistent arrays");
            return A;
          },
          py::arg("n"), py::arg("rowptr"), py::arg("colidx"), py::arg("values"));
  py::class_<SparseMatrix<double>>(m, "SpMatD", "Sparse matrix with real coefficients.")
      .def_property_readonly("shape", [](const SparseMatrix<double> &A) { return py::make_tuple(A.rows(), A.cols()); })
      .def("coeff", &SparseMatrix<double>::coeff)
      .def("to_dense", [](const SparseMatrix<double> &A) { return to_dense(A); })
      .def("nnz", [](const SparseMatrix<double> &A) { return A.nonZeros(); })
      .def("to_csr", [](const SparseMatrix<double> &A) { return csr_arrays(A); });
  py::class_<VecC>(m, "VecC", "Vector with complex coefficients.")
      .def(py::init<>())
      .def(py::init([](py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> a) { return vec_from_numpy(a); }))
      .def("__len__", [](const VecC &v) { return v.size(); })
      .def("__getitem__",
           [](const VecC &v, py::ssize_t i) {
             if (i < 0) i += (py::ssize_t)v.size();
             if (i < 0 || (size_t)i >= v.size()) throw py::index_error();
             return v[(size_t)i];
           })
      .def("to_numpy", [](const VecC &v) { return vec_to_numpy(v); });
  py::implicitly_convertible<py::array, VecC>();

  // ---------------------------------------------------------------- Maxwell params / assembly
  py::enum_<PortABCType>(m, "PortABCType")
      .value("None_", PortABCType::None)
      .value("Beta", PortABCType::Beta)
      .value("BetaNorm", PortABCType::BetaNorm)
      .value("ImpedanceMatch", PortABCType::ImpedanceMatch)
      .value("ModalAdmittance", PortABCType::ModalAdmittance);
  py::class_<PMLRegionSpec>(m, "PMLRegionSpec")
      .def(py::init<>())
      .def_readwrite("sigma_max", &PMLRegionSpec::sigma_max)
      .def_readwrite("thickness", &PMLRegionSpec::thickness)
      .def_readwrite("grading_order", &PMLRegionSpec::grading_order);
  py::class_<PMLDiagnostic>(m, "PMLDiagnostic")
      .def_readonly("region_tag", &PMLDiagnostic::region_tag)
      .def_readonly("sigma_max", &PMLDiagnostic::sigma_max)
      .def_readonly("thickness", &PMLDiagnostic::thickness)
      .def_readonly("reflection_est", &PMLDiagnostic::reflection_est);
  using DispPtr = std::shared_ptr<materials::DispersiveMaterial>;
  py::class_<MaxwellParams>(m, "MaxwellParams", "Parameters for Maxwell's equations.")
      .def(py::init<>())
      .def_readwrite("omega", &MaxwellParams::omega)
      .def_readwrite("eps_r", &MaxwellParams::eps_r)
      .def_readwrite("mu_r", &MaxwellParams::mu_r)
      .def_property("eps_r_regions", [](const MaxwellParams &p) { return p.eps_r_regions; },
                    [](MaxwellParams &p, const std::unordered_map<int, cplx> &v) { p.eps_r_regions = v; })
      .def("set_eps_r_region", [](MaxwellParams &p, int tag, cplx v) { p.eps_r_regions[tag] = v; }, py::arg("phys_tag"), py::arg("eps_r"))
      .def_property("mu_r_regions", [](const MaxwellParams &p) { return p.mu_r_regions; },
                    [](MaxwellParams &p, const std::unordered_map<int, cplx> &v) { p.mu_r_regions = v; })
      .def("set_mu_r_region", [](MaxwellParams &p, int tag, cplx v) { p.mu_r_regions[tag] = v; }, py::arg("phys_tag"), py::arg("mu_r"))
      .def_property("eps_models", [](const MaxwellParams &p) { return p.eps_models; },
                    [](MaxwellParams &p, const std::unordered_map<int, DispPtr> &v) { p.eps_models = v; })
      .def("set_eps_model", [](MaxwellParams &p, int tag, DispPtr mdl) { p.eps_models[tag] = mdl; }, py::arg("phys_tag"), py::arg("model"))
      .def_property("mu_models", [](const MaxwellParams &p) { return p.mu_models; },
                    [](MaxwellParams &p, const std::unordered_map<int, DispPtr> &v) { p.mu_models = v; })
      .def("set_mu_model", [](MaxwellParams &p, int tag, DispPtr mdl) { p.mu_models[tag] = mdl; }, py::arg("phys_tag"), py::arg("model"))
      .def("get_eps_r", [](const MaxwellParams &p, int tag) { return p.get_eps_r(tag); }, py::arg("phys_tag"))
      .def("get_eps_r_at_freq", [](const MaxwellParams &p, int tag, double w) { return p.get_eps_r(tag, w); }, py::arg("phys_tag"), py::arg("omega"))
      .def("get_mu_r", [](const MaxwellParams &p, int tag) { return p.get_mu_r(tag); }, py::arg("phys_tag"))
      .def("get_mu_r_at_freq", [](const MaxwellParams &p, int tag, double w) { return p.get_mu_r(tag, w); }, py::arg("phys_tag"), py::arg("omega"))
      .def_readwrite("pml_sigma", &MaxwellParams::pml_sigma)
      .def_property("pml_regions", [](const MaxwellParams &p) { return p.pml_regions; },
                    [](MaxwellParams &p, const std::unordered_set<int> &v) { p.pml_regions = v; })
      .def_property("pml_tensor_regions", [](const MaxwellParams &p) { return p.pml_tensor_regions; },
                    [](MaxwellParams &p, const std::unordered_map<int, PMLRegionSpec> &v) { p.pml_tensor_regions = v; })
      .def_readwrite("enforce_pml_heuristics", &MaxwellParams::enforce_pml_heuristics)
      .def_readwrite("use_abc", &MaxwellParams::use_abc)
      .def_property("abc_surface_tags", [](const MaxwellParams &p) { return p.abc_surface_tags; },
                    [](MaxwellParams &p, const std::unordered_set<int> &v) { p.abc_surface_tags = v; })
      .def_readwrite("use_port_abc", &MaxwellParams::use_port_abc)
      .def_readwrite("port_abc_type", &MaxwellParams::port_abc_type)
      .def_readwrite("port_weight_scale", &MaxwellParams::port_weight_scale)
      .def_readwrite("port_abc_scale", &MaxwellParams::port_abc_scale)
      .def_readwrite("use_eigenmode_excitation", &MaxwellParams::use_eigenmode_excitation);
  py::class_<MaxwellAssembly>(m, "MaxwellAssembly", "Assembled Maxwell system.")
      .def_property_readonly("A", [](MaxwellAssembly &a) -> SpMatC & { return a.A; }, py::return_value_policy::reference_internal)
      .def_property_readonly("b", [](MaxwellAssembly &a) -> VecC & { return a.b; }, py::return_value_policy::reference_internal)
      .def_readonly("diagnostics", &MaxwellAssembly::diagnostics);

  // ---------------------------------------------------------------- ports
  py::enum_<ModePolarization>(m, "ModePolarization").value("TE", ModePolarization::TE).value("TM", ModePolarization::TM);
  py::class_<PortMode>(m, "PortMode")
      .def(py::init<>())
      .def_readwrite("pol", &PortMode::pol)
      .def_readwrite("fc", &PortMode::fc)
      .def_readwrite("kc", &PortMode::kc)
      .def_readwrite("omega", &PortMode::omega)
      .def_readwrite("eps", &PortMode::eps)
      .def_readwrite("mu", &PortMode::mu)
      .def_readwrite("beta", &PortMode::beta)
      .def_readwrite("Z0", &PortMode::Z0)
      .def_property("field", [](const PortMode &pm) { return vec_to_numpy(pm.field); },
                    [](PortMode &pm, py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> a) { pm.field = vec_from_numpy(a); });
  py::class_<RectWaveguidePort>(m, "RectWaveguidePort")
      .def(py::init<>())
      .def(py::init([](double a, double b) { return RectWaveguidePort{a, b}; }), py::arg("a"), py::arg("b"))
      .def_readwrite("a", &RectWaveguidePort::a)
      .def_readwrite("b", &RectWaveguidePort::b);
  m.def("solve_te10_mode", &solve_te10_mode, py::arg("port"), py::arg("freq"));
  py::class_<WavePort>(m, "WavePort", "Wave port definition with modal data.")
      .def(py::init<>())
      .def_readwrite("surface_tag", &WavePort::surface_tag)
      .def_readwrite("mode", &WavePort::mode)
      .def_readwrite("edges", &WavePort::edges)
      .def_property("weights", [](const WavePort &w) { return vec_to_numpy(w.weights); },
                    [](WavePort &w, py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast> a) { w.weights = vec_from_numpy(a); });
  py::enum_<LumpedPortWeightMode>(m, "LumpedPortWeightMode")
      .value("Projection", LumpedPortWeightMode::Projection)
      .value("SurfaceIntegral", LumpedPortWeightMode::SurfaceIntegral)
      .export_values();
  py::class_<LumpedPortConfig>(m, "LumpedPortConfig")
      .def(py::init<>())
      .def_readwrite("surface_tag", &LumpedPortConfig::surface_tag)
      .def_readwrite("z0", &LumpedPortConfig::z0)
      .def_readwrite("e_direction", &LumpedPortConfig::e_direction)
      .def_readwrite("weight_mode", &LumpedPortConfig::weight_mode);
  m.def("build_lumped_port", &build_lumped_port, py::arg("mesh"), py::arg("config"));
  // ---- nodal port eigenmodes and modal line-integral ports (reference python/pyedgefem.cpp:392-420,488-494,595-598)
  m.def("solve_port_eigens", &solve_port_eigens, "Solve 2D eigenmode problem on port cross-section.", py::arg("mesh"), py::arg("num_modes"),
        py::arg("omega"), py::arg("eps_r"), py::arg("mu_r"), py::arg("pol"));
  py::class_<SParams2>(m, "SParams2", "2-port S-parameter data.")
      .def(py::init<>())
      .def_readwrite("s11", &SParams2::s11)
      .def_readwrite("s21", &SParams2::s21)
      .def_readwrite("s12", &SParams2::s12)
      .def_readwrite("s22", &SParams2::s22);
  m.def("straight_waveguide_sparams", &straight_waveguide_sparams, py::arg("port"), py::arg("length"), py::arg("freq"));
  py::class_<PortSurfaceMesh>(m, "PortSurfaceMesh", "Surface mesh extracted from a volume port.")
      .def(py::init<>())
      .def_property_readonly("mesh", [](PortSurfaceMesh &s) -> Mesh & { return s.mesh; }, py::return_value_policy::reference_internal)
      .def_readonly("volume_tri_indices", &PortSurfaceMesh::volume_tri_indices);
  m.def("extract_surface_mesh", &extract_surface_mesh, "Extract a surface mesh for a port.", py::arg("mesh"), py::arg("surface_tag"));
  m.def("build_wave_port", &build_wave_port, "Project a modal field onto port edges.", py::arg("volume_mesh"), py::arg("surface"),
        py::arg("mode"));
  m.def(
      "populate_te10_field",
      [](const PortSurfaceMesh &surface, const RectWaveguidePort &port, PortMode &mode) { populate_te10_field(surface, port, mode); },
      "Populate TE10 mode field on a surface mesh using analytical formula.", py::arg("surface"), py::arg("port"), py::arg("mode"));
  m.def(
      "build_wave_port_from_eigenvector",
      [](const Mesh &mesh, const PortSurfaceMesh &surface, py::array_t<double, py::array::c_style | py::array::forcecast> ev,
         const PortMode &mode, const std::unordered_set<int> &pec) {
        VectorXd v((size_t)ev.size());
        for (py::ssize_t i = 0; i < ev.size(); ++i) v[(size_t)i] = ev.data()[i];
        return build_wave_port_from_eigenvector(mesh, surface, v, mode, pec);
      },
      "Build wave port using 3D FEM eigenvector as weights.", py::arg("mesh"), py::arg("surface"), py::arg("eigenvector"), py::arg("mode"),
      py::arg("pec_edges"));
  // ---- field post-processing (reference python/pyedgefem.cpp:451-482, 975-1040)
  using CArr = py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast>;
  using DArr = py::array_t<double, py::array::c_style | py::array::forcecast>;
  m.def(
      "evaluate_edge_field",
      [](const std::array<Vector3d, 4> &vertices, const std::array<int, 6> &edge_orient, CArr dofs, const Vector3d &point) {
        if (dofs.size() != 6) throw std::invalid_argument("edge_dofs must hold 6 values");
        std::array<std::complex<double>, 6> d;
        for (int i = 0; i < 6; ++i) d[i] = dofs.data()[i];
        const auto E = evaluate_edge_field(vertices, edge_orient, d, point);
        py::array_t<std::complex<double>> a(3);
        for (int k = 0; k < 3; ++k) a.mutable_data()[k] = E[k];
        return a;
      },
      "Evaluate E-field at a point inside a tetrahedron from edge DOFs.", py::arg("vertices"), py::arg("edge_orient"), py::arg("edge_dofs"),
      py::arg("point"));
  m.def(
      "compute_barycentric",
      [](const std::array<Vector3d, 4> &v, const Vector3d &p) {
        const auto l = compute_barycentric(v, p);
        return py::make_tuple(l[0], l[1], l[2], l[3]);
      },
      py::arg("vertices"), py::arg("point"));
  m.def("whitney_edge_curls", [](const std::array<Vector3d, 4> &v) {
    const auto c = whitney_edge_curls(v);
    return v3_to_numpy(std::vector<Vector3d>(c.begin(), c.end()));
  }, py::arg("vertices"));
  py::class_<HuygensSurfaceData>(m, "HuygensSurfaceData", "Extracted Huygens surface data for near-to-far field transformation.")
      .def(py::init<>())
      .def_property_readonly("r", [](const HuygensSurfaceData &d) { return v3_to_numpy(d.r); }, "Triangle centroids [n,3]")
      .def_property_readonly("n", [](const HuygensSurfaceData &d) { return v3_to_numpy(d.n); }, "Outward normals [n,3]")
      .def_property_readonly("E_tan", [](const HuygensSurfaceData &d) { return c3_to_numpy(d.E_tan); }, "Tangential E-field [n,3]")
      .def_property_readonly("H_tan", [](const HuygensSurfaceData &d) { return c3_to_numpy(d.H_tan); }, "Tangential H-field [n,3]")
      .def_readonly("area", &HuygensSurfaceData::area, "Triangle areas");
  m.def(
      "extract_huygens_surface",
      [](const Mesh &mesh, const VecC &x, int tag, double omega, std::complex<double> mu_r) { return extract_huygens_surface(mesh, x, tag, omega, mu_r); },
      "Extract Huygens surface fields from FEM solution.", py::arg("mesh"), py::arg("solution"), py::arg("surface_tag"), py::arg("omega"),
      py::arg("mu_r") = std::complex<double>(1.0, 0.0));
  py::class_<NTFPoint2D>(m, "NTFPoint2D")
      .def(py::init<>())
      .def_readwrite("theta_deg", &NTFPoint2D::theta_deg)
      .def_readwrite("e_theta", &NTFPoint2D::e_theta)
      .def_readwrite("e_phi", &NTFPoint2D::e_phi);
  m.def(
      "stratton_chu_2d",
      [](DArr r, DArr n, CArr E, CArr H, const std::vector<double> &area, const std::vector<double> &theta, double phi, double k0) {
        return stratton_chu_2d(v3_from_numpy(r), v3_from_numpy(n), c3_from_numpy(E), c3_from_numpy(H), area, theta, phi, k0);
      },
      py::arg("r"), py::arg("n"), py::arg("E"), py::arg("H"), py::arg("area"), py::arg("theta_rad"), py::arg("phi_rad"), py::arg("k0"));
  py::class_<FFPattern3D>(m, "FFPattern3D", "3D far-field pattern over a (theta, phi) grid.")
      .def(py::init<>())
      .def_property_readonly("theta_grid", [](const FFPattern3D &p) { return md_to_numpy(p.theta_grid); })
      .def_property_readonly("phi_grid", [](const FFPattern3D &p) { return md_to_numpy(p.phi_grid); })
      .def_readonly("E_theta", &FFPattern3D::E_theta)
      .def_readonly("E_phi", &FFPattern3D::E_phi)
      .def("total_magnitude", [](const FFPattern3D &p) { return md_to_numpy(p.total_magnitude()); })
      .def("power_pattern", [](const FFPattern3D &p) { return md_to_numpy(p.power_pattern()); })
      .def("pattern_dB", [](const FFPattern3D &p) { return md_to_numpy(p.pattern_dB()); });
  m.def(
      "stratton_chu_3d",
      [](DArr r, DArr n, CArr E, CArr H, const std::vector<double> &area, const std::vector<double> &theta, const std::vector<double> &phi,
         double k0) { return stratton_chu_3d(v3_from_numpy(r), v3_from_numpy(n), c3_from_numpy(E), c3_from_numpy(H), area, theta, phi, k0); },
      py::arg("r"), py::arg("n"), py::arg("E"), py::arg("H"), py::arg("area"), py::arg("theta_rad"), py::arg("phi_rad"), py::arg("k0"));
  m.def("compute_directivity", &compute_directivity, py::arg("pattern"));
  m.def("compute_max_gain", &compute_max_gain, py::arg("pattern"), py::arg("efficiency") = 1.0);
  m.def("compute_hpbw", &compute_hpbw, py::arg("pattern"));
  // ---- Touchstone writers (reference python/pyedgefem.cpp:630-656)
  m.def("write_touchstone", &write_touchstone, "Writes S-parameters to a Touchstone file.", py::arg("path"), py::arg("freq"), py::arg("data"));
  py::enum_<TouchstoneFormat>(m, "TouchstoneFormat")
      .value("RI", TouchstoneFormat::RI, "Real/Imaginary format")
      .value("MA", TouchstoneFormat::MA, "Magnitude/Angle format")
      .value("DB", TouchstoneFormat::DB, "dB/Angle format");
  py::class_<TouchstoneOptions>(m, "TouchstoneOptions", "Options for Touchstone export.")
      .def(py::init<>())
      .def_readwrite("format", &TouchstoneOptions::format)
      .def_readwrite("z0", &TouchstoneOptions::z0);
  m.def(
      "write_touchstone_nport",
      [](const std::string &path, const std::vector<double> &freq,
         const std::vector<py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast>> &mats, const TouchstoneOptions &opts) {
        std::vector<MatrixXcd> S;
        for (const auto &a : mats) {
          if (a.ndim() != 2) throw std::runtime_error("S matrices must be 2-D");
          MatrixXcd M((int)a.shape(0), (int)a.shape(1));
          for (int i = 0; i < M.rows(); ++i)
            for (int j = 0; j < M.cols(); ++j) M(i, j) = a.at(i, j);
          S.push_back(std::move(M));
        }
        write_touchstone_nport(path, freq, S, opts);
      },
      "Writes N-port S-parameters to a Touchstone file.", py::arg("path"), py::arg("freq"), py::arg("S_matrices"),
      py::arg("opts") = TouchstoneOptions());
  m.def("touchstone_extension", &touchstone_extension, py::arg("num_ports"));
  m.def("build_wave_port_2d", &build_wave_port_2d, py::arg("mesh"), py::arg("surface_tag"), py::arg("mode"), py::arg("pec_edges"),
        py::arg("target_kc_sq"));
  m.def(
      "solve_port_mode_2d",
      [](const Mesh &mesh, int surface_tag, const std::unordered_set<int> &pec, double target) {
        double kc = 0.0;
        VectorXd v = solve_port_mode_2d(mesh, surface_tag, pec, target, kc);
        return py::make_tuple(v, kc);
      },
      py::arg("mesh"), py::arg("surface_tag"), py::arg("pec_edges"), py::arg("target_kc_sq"));
  m.def("assemble_port_surface_mass", &assemble_port_surface_mass, py::arg("mesh"), py::arg("surface_tag"), py::arg("dirichlet_edges"));

  // ---------------------------------------------------------------- solver
  py::class_<SolveOptions>(m, "SolveOptions", "Options for the linear solver.")
      .def(py::init<>())
      .def_readwrite("use_bicgstab", &SolveOptions::use_bicgstab)
      .def_readwrite("use_direct", &SolveOptions::use_direct)
      .def_readwrite("tolerance", &SolveOptions::tolerance)
      .def_readwrite("max_iterations", &SolveOptions::max_iterations)
      .def_readwrite("use_ilut", &SolveOptions::use_ilut)
      .def_readwrite("ilut_fill_factor", &SolveOptions::ilut_fill_factor)
      .def_readwrite("ilut_drop_tolerance", &SolveOptions::ilut_drop_tolerance)
      .def_readwrite("auto_fallback", &SolveOptions::auto_fallback)
      .def_readwrite("verbose", &SolveOptions::verbose)
      .def_readwrite("progress_interval", &SolveOptions::progress_interval);
  py::class_<SolveResult>(m, "SolveResult", "Results from a linear solve.")
      .def_readonly("method", &SolveResult::method)
      .def_readonly("iters", &SolveResult::iters)
      .def_readonly("residual", &SolveResult::residual)
      .def_readonly("converged", &SolveResult::converged)
      .def_readonly("error_message", &SolveResult::error_message)
      .def_property_readonly("x", [](SolveResult &r) -> VecC & { return r.x; }, py::return_value_policy::reference_internal);
  m.def("solve_linear", &solve_linear, py::arg("A"), py::arg("b"), py::arg("options") = SolveOptions());

  m.def("assemble_maxwell", &assemble_maxwell, py::arg("mesh"), py::arg("params"), py::arg("bc"), py::arg("ports") = std::vector<WavePort>(),
        py::arg("active_port_idx") = -1);
  m.def("calculate_sparams", &calculate_sparams, py::arg("mesh"), py::arg("params"), py::arg("bc"), py::arg("ports"),
        py::arg("solver_options") = SolveOptions());
  m.def(
      "normalize_port_weights",
      [](const Mesh &mesh, const MaxwellParams &p, const BC &bc, std::vector<WavePort> ports, const SolveOptions &opts) {
        normalize_port_weights(mesh, p, bc, ports, opts);
        return ports;
      },
      py::arg("mesh"), py::arg("params"), py::arg("bc"), py::arg("ports"), py::arg("solver_options") = SolveOptions());
  m.def("calculate_sparams_eigenmode", &calculate_sparams_eigenmode, py::arg("mesh"), py::arg("params"), py::arg("bc"), py::arg("ports"));
  py::class_<BatchStats>(m, "BatchStats")
      .def_readonly("iterations", &BatchStats::iterations)
      .def_readonly("residuals", &BatchStats::residuals)
      .def_property_readonly("converged", [](const BatchStats &s) { return std::vector<int>(s.converged.begin(), s.converged.end()); })
      .def_readonly("device_ms", &BatchStats::device_ms)
      .def_readonly("kernel_launches", &BatchStats::kernel_launches)
      .def_readonly("h2d_bytes", &BatchStats::h2d_bytes)
      .def_readonly("d2h_bytes", &BatchStats::d2h_bytes);
  m.def(
      "calculate_sparams_eigenmode_sweep",
      [](const Mesh &mesh, const MaxwellParams &p, const BC &bc, const std::vector<WavePort> &ports, const std::vector<double> &freqs) {
        BatchStats st;
        std::vector<MatrixXcd> S;
        {
          py::gil_scoped_release rel;
          S = calculate_sparams_eigenmode_sweep(mesh, p, bc, ports, freqs, &st);
        }
        return py::make_tuple(S, st);
      },
      "Eigenmode S-parameters for a list of frequencies as one device batch (extension). Returns (list of S, BatchStats).", py::arg("mesh"),
      py::arg("params"), py::arg("bc"), py::arg("ports"), py::arg("frequencies"));

  py::class_<KMMatrices>(m, "KMMatrices")
      .def_property_readonly("K", [](KMMatrices &k) -> SpMatC & { return k.K; }, py::return_value_policy::reference_internal)
      .def_property_readonly("M", [](KMMatrices &k) -> SpMatC & { return k.M; }, py::return_value_policy::reference_internal)
      .def("combine", &KMMatrices::combine, py::arg("omega"));
  py::class_<SweepResult>(m, "SweepResult")
      .def_readonly("frequencies", &SweepResult::frequencies)
      .def_readonly("S_matrices", &SweepResult::S_matrices);
  m.def("frequency_sweep", &frequency_sweep, py::arg("mesh"), py::arg("params"), py::arg("bc"), py::arg("ports"), py::arg("frequencies"),
        py::arg("solver_options") = SolveOptions());
  m.def("assemble_maxwell_km", &assemble_maxwell_km, py::arg("mesh"), py::arg("params"), py::arg("bc"));

  // ---------------------------------------------------------------- periodic
  py::class_<PeriodicPair>(m, "PeriodicPair")
      .def(py::init<>())
      .def_readwrite("master_edge", &PeriodicPair::master_edge)
      .def_readwrite("slave_edge", &PeriodicPair::slave_edge)
      .def_readwrite("master_orient", &PeriodicPair::master_orient)
      .def_readwrite("slave_orient", &PeriodicPair::slave_orient)
      .def_readwrite("translation", &PeriodicPair::translation);
  py::class_<PeriodicBC>(m, "PeriodicBC")
      .def(py::init<>())
      .def_readwrite("pairs", &PeriodicBC::pairs)
      .def_readwrite("period_vector", &PeriodicBC::period_vector)
      .def_readwrite("phase_shift", &PeriodicBC::phase_shift);
  m.def("build_periodic_pairs", &build_periodic_pairs, py::arg("mesh"), py::arg("master_tag"), py::arg("slave_tag"), py::arg("period_vector"),
        py::arg("tolerance") = 1e-9);
  m.def("validate_periodic_bc", &validate_periodic_bc, py::arg("mesh"), py::arg("pbc"));
  m.def("set_floquet_phase", [](PeriodicBC &pbc, const Vector2d &k) { set_floquet_phase(pbc, k); }, py::arg("pbc"), py::arg("k_transverse"));
  m.def("floquet_phase_from_angle", &floquet_phase_from_angle, py::arg("period_vector"), py::arg("theta"), py::arg("phi"), py::arg("k0"));
  m.def("count_surface_edges", &count_surface_edges, py::arg("mesh"), py::arg("surface_tag"));
  m.def("assemble_maxwell_periodic", &assemble_maxwell_periodic, py::arg("mesh"), py::arg("params"), py::arg("bc"), py::arg("pbc"),
        py::arg("ports") = std::vector<WavePort>(), py::arg("active_port_idx") = -1);
  m.def("calculate_sparams_periodic", &calculate_sparams_periodic, py::arg("mesh"), py::arg("params"), py::arg("bc"), py::arg("pbc"), py::arg("ports"));

  // ---------------------------------------------------------------- element matrices (host)
  m.def("whitney_curl_curl_matrix", [](const std::array<Vector3d, 4> &v) { return whitney_curl_curl_matrix(v); });
  m.def("whitney_mass_matrix", [](const std::array<Vector3d, 4> &v) { return whitney_mass_matrix(v); });
  m.def("triangle_whitney_mass_matrix", [](const std::array<Vector3d, 3> &v) { return triangle_whitney_mass_matrix(v); });

  // ---------------------------------------------------------------- B200 runtime helpers (extension)
  m.def("b200_clear_cache", &edgefem::detail::clear_device_cache, "Drop cached device meshes (forces re-upload on the next call).");
  m.def("b200_launch_count", &edgefem::detail::launch_count, "Number of CUDA kernels launched by this process so far.");
  m.def("b200_device_count", []() { return efb_device_count(); });

  // ---------------------------------------------------------------- materials submodule
  py::module_ mat = m.def_submodule("materials", "Dispersive material models for frequency-dependent permittivity.");
  py::class_<materials::DispersiveMaterial, DispPtr>(mat, "DispersiveMaterial")
      .def("eval_eps", &materials::DispersiveMaterial::eval_eps, py::arg("omega"))
      .def("eval_mu", &materials::DispersiveMaterial::eval_mu, py::arg("omega"));
  py::class_<materials::DebyeMaterial, materials::DispersiveMaterial, std::shared_ptr<materials::DebyeMaterial>>(mat, "DebyeMaterial")
      .def(py::init<double, double, double>(), py::arg("eps_static"), py::arg("eps_inf"), py::arg("tau"))
      .def_property_readonly("eps_static", &materials::DebyeMaterial::eps_static)
      .def_property_readonly("eps_inf", &materials::DebyeMaterial::eps_inf)
      .def_property_readonly("tau", &materials::DebyeMaterial::tau);
  py::class_<materials::LorentzMaterial, materials::DispersiveMaterial, std::shared_ptr<materials::LorentzMaterial>>(mat, "LorentzMaterial")
      .def(py::init<>())
      .def(py::init<double>(), py::arg("eps_inf"))
      .def("add_pole", &materials::LorentzMaterial::add_pole, py::arg("delta_eps"), py::arg("omega0"), py::arg("gamma"))
      .def_property("eps_inf", &materials::LorentzMaterial::eps_inf, &materials::LorentzMaterial::set_eps_inf)
      .def_property_readonly("num_poles", &materials::LorentzMaterial::num_poles);
  py::class_<materials::DrudeMaterial, materials::DispersiveMaterial, std::shared_ptr<materials::DrudeMaterial>>(mat, "DrudeMaterial")
      .def(py::init<double, double>(), py::arg("omega_p"), py::arg("gamma"))
      .def_property_readonly("omega_p", &materials::DrudeMaterial::omega_p)
      .def_property_readonly("gamma", &materials::DrudeMaterial::gamma);
  py::class_<materials::DrudeLorentzMaterial, materials::DispersiveMaterial, std::shared_ptr<materials::DrudeLorentzMaterial>>(mat, "DrudeLorentzMaterial")
      .def(py::init<double, double, double>(), py::arg("eps_inf"), py::arg("omega_p"), py::arg("gamma_d"))
      .def("add_lorentz_pole", &materials::DrudeLorentzMaterial::add_lorentz_pole, py::arg("delta_eps"), py::arg("omega0"), py::arg("gamma"))
      .def_property_readonly("eps_inf", &materials::DrudeLorentzMaterial::eps_inf)
      .def_property_readonly("omega_p", &materials::DrudeLorentzMaterial::omega_p)
      .def_property_readonly("gamma_d", &materials::DrudeLorentzMaterial::gamma_d);
}
