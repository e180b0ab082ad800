// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Minimal half-edge triangle mesh exposing the OpenMesh API subset used by the
// reference sources (27 methods), so that the UNMODIFIED
// reference sources under /root/reference/src compile against it. Semantics
// restated from OpenMesh's published behaviour:
//   * ArrayKernel numbering: edge k <-> half-edges 2k (to b) / 2k+1 (to a),
//     opposite(h) = h^1, edges numbered by first appearance in add_face order,
//     halfedge_handle(face) = v[n-1]->v[0].
//   * default circulators rotate clockwise: h -> next(opposite(h)), starting at
//     halfedge_handle(v); adjust_outgoing_halfedge prefers a boundary half-edge.
//   * TriConnectivity::is_flip_ok / flip; PolyMeshT::calc_sector_* /
//     calc_dihedral_angle.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../Geometry/VectorT.hh"
#include "Handles.hh"

namespace OpenMesh {

struct DefaultTraits {
    typedef Vec3f Point;
    typedef Vec3f Normal;
};

template <class T> inline T sane_aarg(T a) {
    if (a < -1) a = -1; else if (a > 1) a = 1;
    return a;
}
template <class T> inline T angle(T c, T s) {
    c = sane_aarg(c);
    return (T)(s >= 0 ? std::acos(c) : -std::acos(c));
}

template <class Traits>
class TriMesh_ArrayKernelT {
public:
    typedef typename Traits::Point Point;
    typedef typename Traits::Normal Normal;
    typedef typename Point::value_type Scalar;
    typedef TriMesh_ArrayKernelT Self;

    struct HE { int to = -1, next = -1, prev = -1, face = -1; };
    std::vector<Point> points_;
    std::vector<int> vhe_;
    std::vector<HE> he_;
    std::vector<int> fhe_;

    // ---- smart handles
    struct SmartHalfedgeHandle : HalfedgeHandle {
        const Self* m;
        SmartHalfedgeHandle(int i, const Self* m_) : HalfedgeHandle(i), m(m_) {}
        bool is_boundary() const { return m->he_[idx_].face < 0; }
    };
    struct SmartVertexHandle : VertexHandle {
        const Self* m;
        SmartVertexHandle(int i, const Self* m_) : VertexHandle(i), m(m_) {}
    };

    // ---- counts / handles
    size_t n_vertices() const { return points_.size(); }
    size_t n_edges() const { return he_.size() / 2; }
    size_t n_halfedges() const { return he_.size(); }
    size_t n_faces() const { return fhe_.size(); }
    VertexHandle vertex_handle(unsigned i) const { return VertexHandle((int)i); }
    HalfedgeHandle halfedge_handle(unsigned i) const { return HalfedgeHandle((int)i); }
    EdgeHandle edge_handle(unsigned i) const { return EdgeHandle((int)i); }
    FaceHandle face_handle(unsigned i) const { return FaceHandle((int)i); }
    HalfedgeHandle halfedge_handle(VertexHandle v) const { return HalfedgeHandle(vhe_[v.idx()]); }
    HalfedgeHandle halfedge_handle(FaceHandle f) const { return HalfedgeHandle(fhe_[f.idx()]); }
    HalfedgeHandle halfedge_handle(EdgeHandle e, unsigned i) const { return HalfedgeHandle(2 * e.idx() + (int)i); }
    EdgeHandle edge_handle(HalfedgeHandle h) const { return EdgeHandle(h.idx() >> 1); }
    FaceHandle face_handle(HalfedgeHandle h) const { return FaceHandle(he_[h.idx()].face); }
    VertexHandle to_vertex_handle(HalfedgeHandle h) const { return VertexHandle(he_[h.idx()].to); }
    VertexHandle from_vertex_handle(HalfedgeHandle h) const { return to_vertex_handle(opposite_halfedge_handle(h)); }
    HalfedgeHandle next_halfedge_handle(HalfedgeHandle h) const { return HalfedgeHandle(he_[h.idx()].next); }
    HalfedgeHandle prev_halfedge_handle(HalfedgeHandle h) const { return HalfedgeHandle(he_[h.idx()].prev); }
    HalfedgeHandle opposite_halfedge_handle(HalfedgeHandle h) const { return HalfedgeHandle(h.idx() ^ 1); }
    HalfedgeHandle cw_rotated_halfedge_handle(HalfedgeHandle h) const {
        return next_halfedge_handle(opposite_halfedge_handle(h));
    }
    bool is_boundary(HalfedgeHandle h) const { return he_[h.idx()].face < 0; }
    bool is_boundary(EdgeHandle e) const {
        return is_boundary(halfedge_handle(e, 0)) || is_boundary(halfedge_handle(e, 1));
    }
    bool is_boundary(VertexHandle v) const {
        int h = vhe_[v.idx()];
        return !(h >= 0 && he_[h].face >= 0);
    }
    Point& point(VertexHandle v) { return points_[v.idx()]; }
    const Point& point(VertexHandle v) const { return points_[v.idx()]; }

    // ---- circulators (default = clockwise, as OpenMesh)
    struct VOHIter {
        const Self* m;
        int start, cur, lap;
        bool is_valid() const { return cur >= 0 && lap == 0; }
        SmartHalfedgeHandle operator*() const { return SmartHalfedgeHandle(cur, m); }
        VOHIter& operator++() {
            cur = m->cw_rotated_halfedge_handle(HalfedgeHandle(cur)).idx();
            if (cur == start) ++lap;
            return *this;
        }
        bool operator!=(const VOHIter& o) const { return cur != o.cur || lap != o.lap; }
    };
    struct VVIter {
        VOHIter it;
        bool is_valid() const { return it.is_valid(); }
        SmartVertexHandle operator*() const { return SmartVertexHandle(it.m->he_[it.cur].to, it.m); }
        VVIter& operator++() { ++it; return *this; }
        bool operator!=(const VVIter& o) const { return it != o.it; }
    };
    template <class It> struct Range {
        It b, e;
        It begin() const { return b; }
        It end() const { return e; }
    };
    VOHIter voh_iter(VertexHandle v) const { int h = vhe_[v.idx()]; return VOHIter{this, h, h, 0}; }
    Range<VOHIter> voh_range(VertexHandle v) const {
        int h = vhe_[v.idx()];
        return {VOHIter{this, h, h, 0}, VOHIter{this, h, h, h >= 0 ? 1 : 0}};
    }
    Range<VVIter> vv_range(VertexHandle v) const {
        int h = vhe_[v.idx()];
        return {VVIter{VOHIter{this, h, h, 0}}, VVIter{VOHIter{this, h, h, h >= 0 ? 1 : 0}}};
    }
    struct FVIter {
        const Self* m;
        int cur;
        struct Arrow {
            VertexHandle v;
            const VertexHandle* operator->() const { return &v; }
        };
        Arrow operator->() const { return Arrow{VertexHandle(m->he_[cur].to)}; }
        VertexHandle operator*() const { return VertexHandle(m->he_[cur].to); }
        FVIter& operator++() { cur = m->he_[cur].next; return *this; }
    };
    FVIter fv_iter(FaceHandle f) const { return FVIter{this, fhe_[f.idx()]}; }

    // ---- construction (PolyConnectivity::add_face semantics)
    VertexHandle add_vertex(const Point& p) {
        points_.push_back(p);
        vhe_.push_back(-1);
        return VertexHandle((int)points_.size() - 1);
    }
    HalfedgeHandle find_halfedge(VertexHandle a, VertexHandle b) const {
        for (auto it = voh_iter(a); it.is_valid(); ++it)
            if (he_[(*it).idx()].to == b.idx()) return *it;
        return HalfedgeHandle();
    }
    void set_next(int h, int n) { he_[h].next = n; he_[n].prev = h; }
    void adjust_outgoing_halfedge(VertexHandle v) {
        for (auto it = voh_iter(v); it.is_valid(); ++it)
            if (is_boundary(*it)) { vhe_[v.idx()] = (*it).idx(); break; }
    }
    FaceHandle add_face(const std::vector<VertexHandle>& vhs) {
        size_t n = vhs.size();
        assert(n == 3);
        struct ED { int h; bool is_new, needs_adjust; };
        std::vector<ED> ed(n);
        std::vector<std::pair<int, int>> nc;
        for (size_t i = 0, ii = 1; i < n; ++i, ++ii, ii %= n) {
            if (!is_boundary(vhs[i])) { std::cerr << "add_face: complex vertex\n"; return FaceHandle(); }
            ed[i].h = find_halfedge(vhs[i], vhs[ii]).idx();
            ed[i].is_new = ed[i].h < 0;
            ed[i].needs_adjust = false;
            if (!ed[i].is_new && !is_boundary(HalfedgeHandle(ed[i].h))) {
                std::cerr << "add_face: complex edge\n";
                return FaceHandle();
            }
        }
        for (size_t i = 0, ii = 1; i < n; ++i, ++ii, ii %= n) {
            if (!ed[i].is_new && !ed[ii].is_new) {
                int inner_prev = ed[i].h, inner_next = ed[ii].h;
                if (he_[inner_prev].next != inner_next) {
                    int outer_prev = inner_next ^ 1;
                    int boundary_prev = outer_prev;
                    do boundary_prev = he_[boundary_prev].next ^ 1;
                    while (he_[boundary_prev].face >= 0);
                    int boundary_next = he_[boundary_prev].next;
                    if (boundary_prev == inner_prev) {
                        std::cerr << "add_face: patch re-linking failed\n";
                        return FaceHandle();
                    }
                    int patch_start = he_[inner_prev].next, patch_end = he_[inner_next].prev;
                    nc.push_back({boundary_prev, patch_start});
                    nc.push_back({patch_end, boundary_next});
                    nc.push_back({inner_prev, inner_next});
                }
            }
        }
        for (size_t i = 0, ii = 1; i < n; ++i, ++ii, ii %= n)
            if (ed[i].is_new) {
                int h = (int)he_.size();
                he_.push_back(HE());
                he_.push_back(HE());
                he_[h].to = vhs[ii].idx();
                he_[h + 1].to = vhs[i].idx();
                ed[i].h = h;
            }
        int fh = (int)fhe_.size();
        fhe_.push_back(ed[n - 1].h);
        for (size_t i = 0, ii = 1; i < n; ++i, ++ii, ii %= n) {
            int vh = vhs[ii].idx();
            int inner_prev = ed[i].h, inner_next = ed[ii].h;
            int id = 0;
            if (ed[i].is_new) id |= 1;
            if (ed[ii].is_new) id |= 2;
            if (id) {
                int outer_prev = inner_next ^ 1, outer_next = inner_prev ^ 1;
                switch (id) {
                    case 1: {
                        int bp = he_[inner_next].prev;
                        nc.push_back({bp, outer_next});
                        vhe_[vh] = outer_next;
                        break;
                    }
                    case 2: {
                        int bn = he_[inner_prev].next;
                        nc.push_back({outer_prev, bn});
                        vhe_[vh] = bn;
                        break;
                    }
                    case 3:
                        if (vhe_[vh] < 0) {
                            vhe_[vh] = outer_next;
                            nc.push_back({outer_prev, outer_next});
                        } else {
                            int bn = vhe_[vh];
                            int bp = he_[bn].prev;
                            nc.push_back({bp, outer_next});
                            nc.push_back({outer_prev, bn});
                        }
                        break;
                }
                nc.push_back({inner_prev, inner_next});
            } else
                ed[ii].needs_adjust = (vhe_[vh] == inner_next);
            he_[ed[i].h].face = fh;
        }
        for (auto& p : nc) set_next(p.first, p.second);
        for (size_t i = 0; i < n; ++i)
            if (ed[i].needs_adjust) adjust_outgoing_halfedge(vhs[i]);
        return FaceHandle(fh);
    }

    // ---- flips (TriConnectivity semantics)
    bool is_flip_ok(EdgeHandle e) const {
        if (is_boundary(e)) return false;
        HalfedgeHandle hh = halfedge_handle(e, 0), oh = halfedge_handle(e, 1);
        VertexHandle ah = to_vertex_handle(next_halfedge_handle(hh));
        VertexHandle bh = to_vertex_handle(next_halfedge_handle(oh));
        if (ah == bh) return false;
        for (auto it = voh_iter(ah); it.is_valid(); ++it)
            if (he_[(*it).idx()].to == bh.idx()) return false;
        return true;
    }
    void flip(EdgeHandle e) {
        int a0 = 2 * e.idx(), b0 = a0 + 1;
        int a1 = he_[a0].next, a2 = he_[a1].next, b1 = he_[b0].next, b2 = he_[b1].next;
        int va0 = he_[a0].to, va1 = he_[a1].to, vb0 = he_[b0].to, vb1 = he_[b1].to;
        int fa = he_[a0].face, fb = he_[b0].face;
        he_[a0].to = va1;
        he_[b0].to = vb1;
        set_next(a0, a2); set_next(a2, b1); set_next(b1, a0);
        set_next(b0, b2); set_next(b2, a1); set_next(a1, b0);
        he_[a1].face = fb;
        he_[b1].face = fa;
        fhe_[fa] = a0;
        fhe_[fb] = b0;
        if (vhe_[va0] == b0) vhe_[va0] = a1;
        if (vhe_[vb0] == a0) vhe_[vb0] = b1;
    }

    // ---- geometry (PolyMeshT semantics)
    Normal calc_edge_vector(HalfedgeHandle h) const {
        return point(to_vertex_handle(h)) - point(from_vertex_handle(h));
    }
    void calc_edge_vector(HalfedgeHandle h, Normal& e) const { e = calc_edge_vector(h); }
    void calc_sector_vectors(HalfedgeHandle in, Normal& v0, Normal& v1) const {
        calc_edge_vector(next_halfedge_handle(in), v0);
        calc_edge_vector(opposite_halfedge_handle(in), v1);
    }
    void calc_sector_normal(HalfedgeHandle in, Normal& n) const {
        Normal v0, v1;
        calc_sector_vectors(in, v0, v1);
        n = cross(v0, v1);
    }
    Scalar calc_sector_angle(HalfedgeHandle in) const {
        Normal v0, v1;
        calc_sector_vectors(in, v0, v1);
        Scalar denom = norm(v0) * norm(v1);
        if (denom == Scalar(0)) return 0;
        Scalar cos_a = dot(v0, v1) / denom;
        if (is_boundary(in)) {
            Normal fn;
            calc_sector_normal(opposite_halfedge_handle(in), fn);
            Scalar sign_a = dot(cross(v0, v1), fn);
            return angle(cos_a, sign_a);
        }
        return std::acos(sane_aarg(cos_a));
    }
    Scalar calc_dihedral_angle(HalfedgeHandle h) const {
        if (is_boundary(edge_handle(h))) return 0;
        Normal n0, n1, he;
        calc_sector_normal(h, n0);
        calc_sector_normal(opposite_halfedge_handle(h), n1);
        calc_edge_vector(h, he);
        Scalar denom = norm(n0) * norm(n1);
        if (denom == Scalar(0)) return 0;
        Scalar da_cos = dot(n0, n1) / denom;
        Scalar da_sin_sign = dot(cross(n0, n1), he);
        return angle(da_cos, da_sin_sign);
    }
};

}  // namespace OpenMesh
