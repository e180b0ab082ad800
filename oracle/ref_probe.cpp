// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Small pybind11 module linked against the UNMODIFIED reference objects
// (oracle/Makefile target `ref`). It exposes reference internals that
// trimem.core does not bind, so that golden vectors can be generated for them:
//   * the half-edge arrays of the (shim) OpenMesh kernel        -> numbering parity
//   * per-vertex vertex_properties (src/mesh_properties.cpp:12)  -> per-vertex parity
//   * edge_vertex_properties (src/flip_utils.cpp:19), is_flip_ok, flip
//   * flip_serial's loop (src/flips.cpp:22-80) with the random numbers injected,
//     because the reference seeds its mt19937 from the clock (flips.cpp:20).
// The loop below calls the reference's own functions in the reference's order; only
// the source of (edge index, uniform variate) differs.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <cmath>
#include <omp.h>

#include "defs.h"
#include "energy.h"
#include "flip_utils.h"
#include "mesh_properties.h"
#include "mesh_repulsion.h"
#include "mesh_tether.h"

namespace py = pybind11;
using namespace trimem;

static py::array_t<double> props_row(const VertexProperties& p) {
    py::array_t<double> a(6);
    double* o = a.mutable_data();
    o[0] = p.area; o[1] = p.volume; o[2] = p.curvature;
    o[3] = p.bending; o[4] = p.tethering; o[5] = p.repulsion;
    return a;
}

PYBIND11_MODULE(refprobe, m) {
    m.def("halfedge_arrays", [](TriMesh& mesh) {
        int nh = (int)mesh.n_halfedges(), nv = (int)mesh.n_vertices(), nf = (int)mesh.n_faces();
        py::array_t<int> to(nh), next(nh), face(nh), vhe(nv), fhe(nf);
        for (int h = 0; h < nh; h++) {
            auto hh = mesh.halfedge_handle((unsigned)h);
            to.mutable_at(h) = mesh.to_vertex_handle(hh).idx();
            next.mutable_at(h) = mesh.next_halfedge_handle(hh).idx();
            face.mutable_at(h) = mesh.face_handle(hh).idx();
        }
        for (int v = 0; v < nv; v++) vhe.mutable_at(v) = mesh.halfedge_handle(mesh.vertex_handle(v)).idx();
        for (int f = 0; f < nf; f++) fhe.mutable_at(f) = mesh.halfedge_handle(mesh.face_handle(f)).idx();
        return py::make_tuple(to, next, face, vhe, fhe);
    });

    m.def("vertex_properties", [](TriMesh& mesh, EnergyManager& estore) {
        int nv = (int)mesh.n_vertices();
        py::array_t<double> out({nv, 6});
        auto o = out.mutable_unchecked<2>();
        for (int v = 0; v < nv; v++) {
            auto p = vertex_properties(mesh, *estore.bonds, *estore.repulse, mesh.vertex_handle(v));
            o(v, 0) = p.area; o(v, 1) = p.volume; o(v, 2) = p.curvature;
            o(v, 3) = p.bending; o(v, 4) = p.tethering; o(v, 5) = p.repulsion;
        }
        return out;
    });

    m.def("edge_vertex_properties", [](TriMesh& mesh, EnergyManager& estore, int e) {
        return props_row(edge_vertex_properties(mesh, mesh.edge_handle((unsigned)e), *estore.bonds, *estore.repulse));
    });
    m.def("is_flip_ok", [](TriMesh& mesh, int e) { return mesh.is_flip_ok(mesh.edge_handle((unsigned)e)); });
    m.def("flip", [](TriMesh& mesh, int e) { mesh.flip(mesh.edge_handle((unsigned)e)); });

    // flip_serial (src/flips.cpp:22-80) with injected proposals/variates.
    // Returns (accepted flag per proposal, energy after each proposal).
    m.def("flip_sweep", [](TriMesh& mesh, EnergyManager& estore, py::array_t<int> edges, py::array_t<double> u) {
        int n = (int)edges.size();
        py::array_t<int> acc(n);
        py::array_t<double> ener(n);
        VertexProperties props = estore.properties(mesh);
        real e0 = estore.energy(props);
        for (int i = 0; i < n; i++) {
            acc.mutable_at(i) = 0;
            auto eh = mesh.edge_handle((unsigned)edges.at(i));
            if (mesh.is_flip_ok(eh) and !mesh.is_boundary(eh)) {
                auto oprops = edge_vertex_properties(mesh, eh, *(estore.bonds), *(estore.repulse));
                props -= oprops;
                mesh.flip(eh);
                auto nprops = edge_vertex_properties(mesh, eh, *(estore.bonds), *(estore.repulse));
                props += nprops;
                real en = estore.energy(props);
                real de = en - e0;
                real alpha = de < 0.0 ? 1.0 : std::exp(-de);
                if (u.at(i) <= alpha) {
                    e0 = en;
                    acc.mutable_at(i) = 1;
                } else {
                    mesh.flip(eh);
                    props -= nprops;
                    props += oprops;
                }
            }
            ener.mutable_at(i) = e0;
        }
        return py::make_tuple(acc, ener, props_row(props));
    });
}
