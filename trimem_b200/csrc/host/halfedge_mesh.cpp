#include "halfedge_mesh.h"
#include <algorithm>
#include <atomic>
#include <thread>
#if defined(__linux__)
#include <sched.h>
#endif
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <utility>

#include "../../../include/trimem_b200.h"

namespace trimem_b200 {

namespace { int g_device = 0; }
void set_default_device(int device) { g_device = device; }
int default_device() { return g_device; }

HalfedgeMesh::~HalfedgeMesh() {
    if (dev_) tm_mesh_destroy(dev_);
    if (xyz_) tm_host_free(xyz_);
}

void HalfedgeMesh::set_points(const double* xyz, int n) {
    if (xyz_) tm_host_free(xyz_);
    nv_ = n;
    xyz_ = static_cast<double*>(tm_host_alloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1)));
    if (!xyz_) throw std::bad_alloc();
    if (n > 0) std::memcpy(xyz_, xyz, sizeof(double) * 3 * (size_t)n);
    v_he_.assign(n, -1);
    he_to_.clear(); he_next_.clear(); he_prev_.clear(); he_face_.clear(); f_he_.clear();
    ++topo_version_;
    resident_ = false;
}

void HalfedgeMesh::update_points(const double* xyz) {
    if (nv_ <= 0) throw std::runtime_error("TriMesh is empty");
    tm_mesh* dm = device(false);
    if (tm_sync(dm)) throw std::runtime_error(tm_last_error());      // (an upload from xyz_ may still be in flight)
    // A 16 MB copy takes one core ~1 ms, more than the upload and the sweeps behind it together: up to four threads,
    // and streaming stores -- lines left dirty in several cores' caches make the upload's DMA three times slower
    // (0.45 -> 1.03 ms for the energy() behind a plain four-thread memcpy, scripts/set_points_timing.py).
    int cores = 4;
#if defined(__linux__)
    { cpu_set_t set; if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set); }
#endif
    static const int nt_env = [] { const char* e = getenv("TM_SET_POINTS_THREADS"); return e ? atoi(e) : 0; }();
    int nt = nt_env > 0 ? nt_env : (cores >= 6 ? 4 : (cores >= 3 ? 2 : 1));
    if (nt > 16) nt = 16;
    // pieces of whole rows, a multiple of 8 rows each (streaming stores want 16-byte aligned destinations); piece p is
    // copied by thread p % nt, and uploaded by this thread as soon as it is in place: the upload runs beside the copy
    const int rows_per = std::max(8, ((nv_ + 4 * nt - 1) / (4 * nt) + 7) & ~7);
    const int pieces = nv_ < (1 << 18) ? 1 : (nv_ + rows_per - 1) / rows_per;     // (below ~6 MB the threads cost more than they copy)
    auto copy = [this, xyz](size_t a, size_t b) {           // doubles [a, b), a even (xyz_ is page-aligned)
        double* d = xyz_;
        size_t k = a;
#if defined(__SSE2__)
        for (; k + 2 <= b; k += 2) _mm_stream_pd(d + k, _mm_loadu_pd(xyz + k));
        _mm_sfence();
#endif
        for (; k < b; k++) d[k] = xyz[k];
    };
    if (pieces <= 1) {
        copy(0, 3 * (size_t)nv_);
        if (tm_mesh_set_positions(dm, xyz_)) throw std::runtime_error(tm_last_error());
    } else {
        std::vector<std::atomic<int>> done(pieces);
        for (auto& f : done) f.store(0, std::memory_order_relaxed);
        auto rows = [&](int p, int& first, int& count) { first = p * rows_per; count = std::min(rows_per, nv_ - first); };
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&, t] {
                for (int p = t; p < pieces; p += nt) {
                    int first, count;
                    rows(p, first, count);
                    copy(3 * (size_t)first, 3 * (size_t)(first + count));
                    done[p].store(1, std::memory_order_release);
                }
            });
        int rc = 0;
        for (int p = 0; p < pieces; p++) {
            while (!done[p].load(std::memory_order_acquire)) std::this_thread::yield();
            int first, count;
            rows(p, first, count);
            if (!rc) rc = tm_mesh_set_positions_rows(dm, xyz_, first, count, p == pieces - 1);
        }
        for (auto& t : th) t.join();
        if (rc) throw std::runtime_error(tm_last_error());
    }
    resident_ = true;
}

int HalfedgeMesh::find_halfedge(int a, int b) const {
    int h0 = v_he_[a];
    if (h0 < 0) return -1;
    int h = h0;
    do {
        if (he_to_[h] == b) return h;
        h = he_next_[h ^ 1];
    } while (h != h0);
    return -1;
}

void HalfedgeMesh::adjust_outgoing_halfedge(int v) {
    int h0 = v_he_[v];
    if (h0 < 0) return;
    int h = h0;
    do {
        if (he_face_[h] < 0) { v_he_[v] = h; return; }
        h = he_next_[h ^ 1];
    } while (h != h0);
}

// PolyConnectivity::add_face restricted to triangles. Order of operations matters: it
// fixes the edge numbering, the stored outgoing half-edge of every vertex and the
// boundary-loop links, all of which are visible through handles / summation order.
int HalfedgeMesh::add_face(int va, int vb, int vc) {
    const int v[3] = {va, vb, vc};
    int eh[3];
    bool is_new[3], needs_adjust[3] = {false, false, false};
    std::pair<int, int> relink[18];
    int n_relink = 0;

    for (int i = 0; i < 3; i++) {
        int ii = (i + 1) % 3;
        if (!is_boundary_vertex(v[i])) return -1;                 // complex vertex
        eh[i] = find_halfedge(v[i], v[ii]);
        is_new[i] = eh[i] < 0;
        if (!is_new[i] && he_face_[eh[i]] >= 0) return -1;        // complex edge
    }

    // re-link patches when two consecutive existing boundary half-edges are not adjacent
    for (int i = 0; i < 3; i++) {
        int ii = (i + 1) % 3;
        if (is_new[i] || is_new[ii]) continue;
        int inner_prev = eh[i], inner_next = eh[ii];
        if (he_next_[inner_prev] == inner_next) continue;
        int outer_prev = inner_next ^ 1;
        int boundary_prev = outer_prev;
        do boundary_prev = he_next_[boundary_prev] ^ 1;
        while (he_face_[boundary_prev] >= 0);
        int boundary_next = he_next_[boundary_prev];
        if (boundary_prev == inner_prev) return -1;               // patch re-linking failed
        int patch_start = he_next_[inner_prev], patch_end = he_prev_[inner_next];
        relink[n_relink++] = {boundary_prev, patch_start};
        relink[n_relink++] = {patch_end, boundary_next};
        relink[n_relink++] = {inner_prev, inner_next};
    }

    // create the missing edges (edge ids by first appearance)
    for (int i = 0; i < 3; i++) {
        if (!is_new[i]) continue;
        int ii = (i + 1) % 3;
        int h = (int)he_to_.size();
        he_to_.push_back(v[ii]); he_to_.push_back(v[i]);
        he_next_.push_back(-1); he_next_.push_back(-1);
        he_prev_.push_back(-1); he_prev_.push_back(-1);
        he_face_.push_back(-1); he_face_.push_back(-1);
        eh[i] = h;
    }

    int fh = (int)f_he_.size();
    f_he_.push_back(eh[2]);

    for (int i = 0; i < 3; i++) {
        int ii = (i + 1) % 3;
        int vh = v[ii];
        int inner_prev = eh[i], inner_next = eh[ii];
        int id = (is_new[i] ? 1 : 0) | (is_new[ii] ? 2 : 0);
        if (id) {
            int outer_prev = inner_next ^ 1, outer_next = inner_prev ^ 1;
            if (id == 1) {                          // prev is new, next is old
                int bp = he_prev_[inner_next];
                relink[n_relink++] = {bp, outer_next};
                v_he_[vh] = outer_next;
            } else if (id == 2) {                   // next is new, prev is old
                int bn = he_next_[inner_prev];
                relink[n_relink++] = {outer_prev, bn};
                v_he_[vh] = bn;
            } else if (v_he_[vh] < 0) {             // both new, isolated vertex
                v_he_[vh] = outer_next;
                relink[n_relink++] = {outer_prev, outer_next};
            } else {                                // both new, vertex already on a boundary
                int bn = v_he_[vh];
                int bp = he_prev_[bn];
                relink[n_relink++] = {bp, outer_next};
                relink[n_relink++] = {outer_prev, bn};
            }
            relink[n_relink++] = {inner_prev, inner_next};
        } else {
            needs_adjust[ii] = (v_he_[vh] == inner_next);
        }
        he_face_[eh[i]] = fh;
    }

    for (int k = 0; k < n_relink; k++) set_next(relink[k].first, relink[k].second);
    for (int i = 0; i < 3; i++)
        if (needs_adjust[i]) adjust_outgoing_halfedge(v[i]);
    ++topo_version_;
    return fh;
}

bool HalfedgeMesh::is_flip_ok(int e) const {
    if (is_boundary_edge(e)) return false;
    int hh = 2 * e, oh = hh + 1;
    int ah = he_to_[he_next_[hh]];
    int bh = he_to_[he_next_[oh]];
    if (ah == bh) return false;
    // the flipped edge (ah, bh) must not exist yet
    int h0 = v_he_[ah], h = h0;
    if (h0 >= 0) {
        do {
            if (he_to_[h] == bh) return false;
            h = he_next_[h ^ 1];
        } while (h != h0);
    }
    return true;
}

void HalfedgeMesh::flip(int e) {
    int a0 = 2 * e, b0 = a0 + 1;
    int a1 = he_next_[a0], a2 = he_next_[a1];
    int b1 = he_next_[b0], b2 = he_next_[b1];
    int va0 = he_to_[a0], va1 = he_to_[a1];
    int vb0 = he_to_[b0], vb1 = he_to_[b1];
    int fa = he_face_[a0], fb = he_face_[b0];

    he_to_[a0] = va1;
    he_to_[b0] = vb1;
    set_next(a0, a2); set_next(a2, b1); set_next(b1, a0);
    set_next(b0, b2); set_next(b2, a1); set_next(a1, b0);
    he_face_[a1] = fb;
    he_face_[b1] = fa;
    f_he_[fa] = a0;
    f_he_[fb] = b0;
    if (v_he_[va0] == b0) v_he_[va0] = a1;
    if (v_he_[vb0] == a0) v_he_[vb0] = b1;
    ++topo_version_;
}

tm_mesh* HalfedgeMesh::device(bool upload_positions) {
    if (nv_ <= 0) throw std::runtime_error("TriMesh is empty");
    if (!dev_) {
        if (tm_mesh_create(g_device, &dev_)) { dev_ = nullptr; throw std::runtime_error(tm_last_error()); }
    }
    if (dev_topo_version_ != topo_version_) {
        if (tm_mesh_set_topology(dev_, nv_, n_halfedges(), he_to_.data(), he_next_.data(),
                                 he_face_.data(), v_he_.data()))
            throw std::runtime_error(tm_last_error());
        dev_topo_version_ = topo_version_;
    }
    if (upload_positions && !resident_ && tm_mesh_set_positions(dev_, xyz_))
        throw std::runtime_error(tm_last_error());
    return dev_;
}

}  // namespace trimem_b200
