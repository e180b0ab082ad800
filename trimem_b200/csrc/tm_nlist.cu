// Neighbour lists on the device: cell grid (one-pass counting sort on the cell
// key) and Verlet rows (count -> scan -> fill pair search over 27 cells).
//
// Replaces, in the reference (all serial CPU code there):
//   make_nlist                        src/nlists/nlist.cpp:10-53
//   CellList ctor, compute_cell_pairs src/nlists/cell_list.h:43-129
//   CellList::distance_matrix         src/nlists/cell_list.h:221-312
//   CellList::point_distances         src/nlists/cell_list.h:131-219   (tm_neigh.cuh)
//   VerletList ctor / queries         src/nlists/verlet_list.h:18-101
//
// The reference's quirks are kept: the bounding box
// starts from {DBL_MAX,0,0}/{-1,0,0}, cell edge = dim/int(dim/rlist) per
// dimension, cell coordinates by C truncation, `<=` at build vs `<` at query,
// distances evaluated without FMA contraction so cut-off decisions are
// bit-identical to the x86-64 build of the reference.
#include <cfloat>
#include <algorithm>
#include <vector>

#include "tm_common.cuh"
#include "tm_neigh.cuh"

namespace {

constexpr int NT = 256;

// ---- cell grid ------------------------------------------------------------------
struct Grid {
    double box_min[3], cell[3];
    int shape[3], strides[3];
};

__device__ __forceinline__ int cell_of_point(const Grid& g, d3 p) {
    // `id += int((point[k] - box_min[k]) / r_list[k]) * strides[k]`  (cell_list.h:71-78)
    int id = (int)((p.x - g.box_min[0]) / g.cell[0]) * g.strides[0];
    id += (int)((p.y - g.box_min[1]) / g.cell[1]) * g.strides[1];
    id += (int)((p.z - g.box_min[2]) / g.cell[2]) * g.strides[2];
    return id;
}

__global__ void k_cell_count(Grid g, const P4* __restrict__ pos, int V, int* __restrict__ cell_of,
                             int* __restrict__ count) {
    int i = blockIdx.x * NT + threadIdx.x;
    if (i >= V) return;
    int c = cell_of_point(g, ldpos(pos, i));
    cell_of[i] = c;
    atomicAdd(count + c, 1);
}

__global__ void k_cell_fill(int V, const int* __restrict__ cell_of, const int* __restrict__ cell_start,
                            int* __restrict__ cursor, int* __restrict__ items) {
    int i = blockIdx.x * NT + threadIdx.x;
    if (i >= V) return;
    int c = cell_of[i];
    int slot = atomicAdd(cursor + c, 1);
    items[cell_start[c] + slot] = i;
}

// ---- pair search -------------------------------------------------------------------
// Visits every j != i in the 27 cells around the BUILD-TIME cell of i that is not
// topologically excluded and lies within `cut` (<= when INCLUSIVE, else <).
template <bool INCLUSIVE, class F>
__device__ __forceinline__ void for_each_cell_pair(const MeshView& mv, const NlView& nl, int i, int cell,
                                                   double cut, F f) {
    int kc = cell / nl.strides[2];
    int rem = cell % nl.strides[2];
    int jc = rem / nl.strides[1];
    int ic = rem % nl.strides[1];
    d3 xi = ldpos(mv.pos, i);
    for (int a = -1; a < 2; a++) {
        int ii = ic + a;
        if (ii < 0 || ii >= nl.shape[0]) continue;
        for (int b = -1; b < 2; b++) {
            int jj = jc + b;
            if (jj < 0 || jj >= nl.shape[1]) continue;
            for (int g = -1; g < 2; g++) {
                int kk = kc + g;
                if (kk < 0 || kk >= nl.shape[2]) continue;
                int oc = ii * nl.strides[0] + jj * nl.strides[1] + kk * nl.strides[2];
                int cb = nl.cell_start[oc], ce = nl.cell_start[oc + 1];
                for (int k = cb; k < ce; k++) {
                    int j = nl.cell_items[k];
                    if (j == i) continue;
                    double l = dist_nofma(sub_nofma(xi, ldpos(mv.pos, j)));
                    bool in = INCLUSIVE ? (l <= cut) : (l < cut);
                    if (in && !excluded(mv, i, j, nl.excl)) f(j, l);
                }
            }
        }
    }
}

__global__ void k_verlet_count(MeshView mv, NlView nl, const int* __restrict__ cell_of, double rlist,
                               int* __restrict__ cnt) {
    int i = blockIdx.x * NT + threadIdx.x;
    if (i >= mv.V) return;
    int n = 0;
    for_each_cell_pair<true>(mv, nl, i, cell_of[i], rlist, [&](int, double) { ++n; });
    cnt[i] = n;
}

__global__ void k_verlet_fill(MeshView mv, NlView nl, const int* __restrict__ cell_of, double rlist,
                              const int* __restrict__ row_off, int* __restrict__ row_j) {
    int i = blockIdx.x * NT + threadIdx.x;
    if (i >= mv.V) return;
    int o = row_off[i];
    for_each_cell_pair<true>(mv, nl, i, cell_of[i], rlist, [&](int j, double) { row_j[o++] = j; });
}

// distance_matrix of a cell list: pairs i<j with dist <= rdist (cell_list.h:221-312)
__global__ void k_cell_dm_count(MeshView mv, NlView nl, const int* __restrict__ cell_of, double rdist,
                                int* __restrict__ cnt) {
    int i = blockIdx.x * NT + threadIdx.x;
    if (i >= mv.V) return;
    int n = 0;
    for_each_cell_pair<true>(mv, nl, i, cell_of[i], rdist, [&](int j, double) { if (j > i) ++n; });
    cnt[i] = n;
}

__global__ void k_cell_dm_fill(MeshView mv, NlView nl, const int* __restrict__ cell_of, double rdist,
                               const int* __restrict__ off, double* __restrict__ dist,
                               int* __restrict__ idx, int* __restrict__ jdx) {
    int i = blockIdx.x * NT + threadIdx.x;
    if (i >= mv.V) return;
    int o = off[i], b = o;
    for_each_cell_pair<true>(mv, nl, i, cell_of[i], rdist, [&](int j, double l) {
        if (j > i) { dist[o] = l; idx[o] = i; jdx[o] = j; ++o; }
    });
    // ascending j inside the row
    for (int a = b + 1; a < o; a++) {
        int v = jdx[a]; double dv = dist[a]; int q = a - 1;
        while (q >= b && jdx[q] > v) { jdx[q + 1] = jdx[q]; dist[q + 1] = dist[q]; --q; }
        jdx[q + 1] = v; dist[q + 1] = dv;
    }
}

// distance_matrix of a verlet list: both directions, dist < dmax (verlet_list.h:37-72)
__global__ void k_verlet_dm_count(MeshView mv, NlView nl, double dmax, int* __restrict__ cnt) {
    int i = blockIdx.x * NT + threadIdx.x;
    if (i >= mv.V) return;
    d3 xi = ldpos(mv.pos, i);
    int n = 0;
    for (int k = nl.row_off[i]; k < nl.row_off[i + 1]; k++)
        if (dist_nofma(sub_nofma(xi, ldpos(mv.pos, nl.row_j[k]))) < dmax) ++n;
    cnt[i] = n;
}

__global__ void k_verlet_dm_fill(MeshView mv, NlView nl, double dmax, const int* __restrict__ off,
                                 double* __restrict__ dist, int* __restrict__ idx, int* __restrict__ jdx) {
    int i = blockIdx.x * NT + threadIdx.x;
    if (i >= mv.V) return;
    d3 xi = ldpos(mv.pos, i);
    int o = off[i];
    for (int k = nl.row_off[i]; k < nl.row_off[i + 1]; k++) {
        int j = nl.row_j[k];
        double l = dist_nofma(sub_nofma(xi, ldpos(mv.pos, j)));
        if (l < dmax) { dist[o] = l; idx[o] = i; jdx[o] = j; ++o; }
    }
}

// point_distances for one vertex (single thread; a diagnostic/query entry, not the hot path)
__global__ void k_point_distances(MeshView mv, NlView nl, int pid, double dmax, int cap,
                                  int* __restrict__ n_out, double* __restrict__ dxyz, int* __restrict__ jdx) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int n = 0;
    d3 xi = ldpos(mv.pos, pid);
    for_each_neighbour(mv, nl, pid, xi, dmax, [&](int j, d3 d, double) {
        if (n < cap) { dxyz[3 * n] = d.x; dxyz[3 * n + 1] = d.y; dxyz[3 * n + 2] = d.z; jdx[n] = j; }
        ++n;
    });
    *n_out = n;
}

template <class T>
struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) { TM_CUDA(cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T))); return 0; }
};

int build_cells(tm_mesh* m, tm_nlist* nl, int** cell_of_out) {
    const int V = m->V;
    // bounding box (cell_list.h:45-56, including its initial values)
    double hb[6];
    if (tm_bbox(m, reinterpret_cast<const double*>(m->pos), 4, V, hb)) return 1;
    const double init_min[3] = {DBL_MAX, 0.0, 0.0}, init_max[3] = {-1.0, 0.0, 0.0};
    long long ncell = 1;
    for (int k = 0; k < 3; k++) {
        nl->box_min[k] = hb[k] < init_min[k] ? hb[k] : init_min[k];
        nl->box_max[k] = hb[3 + k] > init_max[k] ? hb[3 + k] : init_max[k];
        // (cell_list.h:59-65)
        double dim = nl->box_max[k] - nl->box_min[k] + nl->box_eps;
        int nk = (int)(dim / nl->rlist);
        nl->cell[k] = (nk > 0) ? dim / nk : dim;
        nl->shape[k] = (nk > 0) ? nk : 1;
        ncell *= nl->shape[k];
    }
    nl->strides[0] = 1; nl->strides[1] = nl->shape[0]; nl->strides[2] = nl->shape[0] * nl->shape[1];
    TM_CHECK(ncell < (1LL << 30), "cell grid of %d x %d x %d cells is too large for the dense cell table "
             "(rlist=%g is tiny compared with the bounding box)", nl->shape[0], nl->shape[1], nl->shape[2], nl->rlist);
    nl->ncell = ncell;

    TM_CUDA(cudaMalloc(&nl->cell_start, sizeof(int) * (size_t)(ncell + 1)));
    TM_CUDA(cudaMalloc(&nl->cell_items, sizeof(int) * (size_t)V));
    DevBuf<int> count;
    if (count.alloc((size_t)ncell + 1)) return 1;
    int* cell_of = nullptr;
    TM_CUDA(cudaMalloc(&cell_of, sizeof(int) * (size_t)V));
    *cell_of_out = cell_of;
    TM_CUDA(cudaMemsetAsync(count.p, 0, sizeof(int) * (size_t)(ncell + 1), m->stream));
    Grid g;
    for (int k = 0; k < 3; k++) {
        g.box_min[k] = nl->box_min[k]; g.cell[k] = nl->cell[k];
        g.shape[k] = nl->shape[k]; g.strides[k] = nl->strides[k];
    }
    int vb = (V + NT - 1) / NT;
    k_cell_count<<<vb, NT, 0, m->stream>>>(g, m->pos, V, cell_of, count.p); TM_LAUNCHED();
    if (tm_exclusive_scan(m, count.p, nl->cell_start, (int)ncell, nullptr)) return 1;
    TM_CUDA(cudaMemsetAsync(count.p, 0, sizeof(int) * (size_t)(ncell + 1), m->stream));
    k_cell_fill<<<vb, NT, 0, m->stream>>>(V, cell_of, nl->cell_start, count.p, nl->cell_items); TM_LAUNCHED();
    if (tm_sort_segments(m, ncell, nl->cell_start, nl->cell_items)) return 1;
    TM_CUDA(cudaGetLastError());
    TM_CUDA(cudaStreamSynchronize(m->stream));   // `count` is released on return
    return 0;
}

}  // namespace

extern "C" {

int tm_nlist_create(tm_mesh* m, int kind, double rlist, int exclusion_level, double box_eps, tm_nlist** out) {
    TM_CHECK(out != nullptr, "null output pointer");
    *out = nullptr;
    // argument errors first, with the reference's messages (nlist.cpp:32,49,52)
    TM_CHECK(kind == TM_NLIST_CELL || kind == TM_NLIST_VERLET, "Unknown neighbour search algorithm.");
    TM_CHECK(exclusion_level >= 0 && exclusion_level <= 2, "Unsupported exclusion level");
    TM_CHECK(m != nullptr && m->V > 0 && m->he_to_h, "mesh has no topology (call tm_mesh_set_topology first)");
    TM_CHECK(rlist > 0.0, "rlist must be positive (got %g)", rlist);
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_ensure_layout(m)) return 1;

    tm_nlist* nl = new tm_nlist();
    nl->device = m->device;
    nl->kind = kind; nl->excl = exclusion_level; nl->rlist = rlist; nl->box_eps = box_eps; nl->V = m->V;
    nl->layout_epoch = m->layout_epoch;
    int* cell_of = nullptr;
    int rc = build_cells(m, nl, &cell_of);
    if (rc == 0 && kind == TM_NLIST_VERLET) {
        // the verlet list is the symmetrised distance_matrix(rlist) of a temporary cell list
        // (verlet_list.h:18-35): count -> scan -> fill, rows ascending
        NlView cv = nl->view();
        cv.kind = TM_NLIST_CELL;
        int vb = (m->V + NT - 1) / NT;
        long long total = 0;
        rc = (cudaMalloc(&nl->row_off, sizeof(int) * ((size_t)m->V + 1)) != cudaSuccess);
        if (rc) tm_set_error("out of device memory for verlet rows");
        if (!rc) {
            k_verlet_count<<<vb, NT, 0, m->stream>>>(m->view(), cv, cell_of, rlist, nl->row_off); TM_LAUNCHED();
            rc = tm_exclusive_scan(m, nl->row_off, nl->row_off, m->V, &total);
        }
        if (!rc) {
            nl->n_entries = total;
            rc = (cudaMalloc(&nl->row_j, sizeof(int) * (size_t)(total > 0 ? total : 1)) != cudaSuccess);
            if (rc) tm_set_error("out of device memory for %lld verlet entries", total);
        }
        if (!rc) {
            k_verlet_fill<<<vb, NT, 0, m->stream>>>(m->view(), cv, cell_of, rlist, nl->row_off, nl->row_j); TM_LAUNCHED();
            rc = tm_sort_segments(m, m->V, nl->row_off, nl->row_j);
            if (!rc) rc = (cudaStreamSynchronize(m->stream) != cudaSuccess);
            if (rc) tm_set_error("verlet build failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
        // a verlet list does not keep the grid
        cudaFree(nl->cell_start); nl->cell_start = nullptr;
        cudaFree(nl->cell_items); nl->cell_items = nullptr;
        cudaFree(cell_of); cell_of = nullptr;
    }
    if (rc) { cudaFree(cell_of); tm_nlist_destroy(nl); return 1; }
    if (kind == TM_NLIST_CELL) nl->cell_of = cell_of;   // build-time cells, for distance_matrix
    *out = nl;
    return 0;
}

int tm_nlist_destroy(tm_nlist* nl) {
    if (!nl) return 0;
    cudaSetDevice(nl->device);
    cudaFree(nl->cell_start); cudaFree(nl->cell_items); cudaFree(nl->row_off); cudaFree(nl->row_j);
    cudaFree(nl->cell_of);
    delete nl;
    return 0;
}

int tm_nlist_info(tm_nlist* nl, long long* n_entries, int shape[3], double box[6], double cell[3]) {
    TM_CHECK(nl != nullptr, "null neighbour list");
    if (n_entries) *n_entries = nl->n_entries;
    for (int k = 0; k < 3; k++) {
        if (shape) shape[k] = nl->shape[k];
        if (box) { box[k] = nl->box_min[k]; box[3 + k] = nl->box_max[k]; }
        if (cell) cell[k] = nl->cell[k];
    }
    return 0;
}

int tm_nlist_distance_matrix(tm_nlist* nl, tm_mesh* m, double rdist, long long* n, double** dist,
                             int** idx, int** jdx) {
    TM_CHECK(nl && m && n && dist && idx && jdx, "null argument");
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_ensure_layout(m) || tm_check_nlist(m, nl)) return 1;
    const int V = m->V, vb = (V + NT - 1) / NT;
    DevBuf<int> off;
    if (off.alloc((size_t)V + 1)) return 1;
    NlView v = nl->view();
    long long total = 0;
    if (nl->kind == TM_NLIST_CELL) {
        k_cell_dm_count<<<vb, NT, 0, m->stream>>>(m->view(), v, nl->cell_of, rdist, off.p); TM_LAUNCHED();
    } else {
        k_verlet_dm_count<<<vb, NT, 0, m->stream>>>(m->view(), v, rdist, off.p); TM_LAUNCHED();
    }
    if (tm_exclusive_scan(m, off.p, off.p, V, &total)) return 1;
    DevBuf<double> dd;
    DevBuf<int> di, dj;
    if (dd.alloc(total) || di.alloc(total) || dj.alloc(total)) return 1;
    if (nl->kind == TM_NLIST_CELL) {
        k_cell_dm_fill<<<vb, NT, 0, m->stream>>>(m->view(), v, nl->cell_of, rdist, off.p, dd.p, di.p, dj.p);
    } else {
        k_verlet_dm_fill<<<vb, NT, 0, m->stream>>>(m->view(), v, rdist, off.p, dd.p, di.p, dj.p);
    }
    TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    size_t cnt = (size_t)(total > 0 ? total : 1);
    *dist = (double*)malloc(cnt * sizeof(double));
    *idx = (int*)malloc(cnt * sizeof(int));
    *jdx = (int*)malloc(cnt * sizeof(int));
    TM_CHECK(*dist && *idx && *jdx, "out of host memory for %lld pairs", total);
    TM_CUDA(cudaMemcpyAsync(*dist, dd.p, total * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaMemcpyAsync(*idx, di.p, total * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaMemcpyAsync(*jdx, dj.p, total * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    std::vector<int> d2h(V);
    TM_CUDA(cudaMemcpyAsync(d2h.data(), m->d2h, sizeof(int) * V, cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    // device numbering -> host numbering (a query entry, not the hot path): map the ids, keep
    // i < j for cell lists (the distance is symmetric bit for bit), sort by (i, j)
    std::vector<long long> order((size_t)total);
    for (long long k = 0; k < total; k++) {
        int a = d2h[(*idx)[k]], b = d2h[(*jdx)[k]];
        if (nl->kind == TM_NLIST_CELL && a > b) std::swap(a, b);
        (*idx)[k] = a; (*jdx)[k] = b;
        order[(size_t)k] = k;
    }
    const int* pi = *idx;
    const int* pj = *jdx;
    std::sort(order.begin(), order.end(), [pi, pj](long long x, long long y) {
        return pi[x] != pi[y] ? pi[x] < pi[y] : pj[x] < pj[y];
    });
    std::vector<double> td((size_t)total);
    std::vector<int> ti((size_t)total), tj((size_t)total);
    for (long long k = 0; k < total; k++) {
        td[(size_t)k] = (*dist)[order[(size_t)k]]; ti[(size_t)k] = pi[order[(size_t)k]]; tj[(size_t)k] = pj[order[(size_t)k]];
    }
    if (total > 0) {
        memcpy(*dist, td.data(), sizeof(double) * (size_t)total);
        memcpy(*idx, ti.data(), sizeof(int) * (size_t)total);
        memcpy(*jdx, tj.data(), sizeof(int) * (size_t)total);
    }
    *n = total;
    return 0;
}

int tm_nlist_point_distances(tm_nlist* nl, tm_mesh* m, int pid, double rdist, int* n, double** dxyz, int** jdx) {
    TM_CHECK(nl && m && n && dxyz && jdx, "null argument");
    TM_CHECK(pid >= 0 && pid < m->V, "vertex index %d out of range", pid);
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_ensure_layout(m) || tm_check_nlist(m, nl)) return 1;
    NlView v = nl->view();
    int hpid = pid;                      // host id -> device slot
    TM_CUDA(cudaMemcpyAsync(&pid, m->h2d + hpid, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    int cap = 4096;
    for (int attempt = 0; attempt < 2; attempt++) {
        DevBuf<double> dd;
        DevBuf<int> dj, dn;
        if (dd.alloc((size_t)cap * 3) || dj.alloc(cap) || dn.alloc(1)) return 1;
        k_point_distances<<<1, 32, 0, m->stream>>>(m->view(), v, pid, rdist, cap, dn.p, dd.p, dj.p); TM_LAUNCHED();
        int cnt = 0;
        TM_CUDA(cudaMemcpyAsync(&cnt, dn.p, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        TM_CUDA(cudaStreamSynchronize(m->stream));
        if (cnt > cap) { cap = cnt; continue; }
        *dxyz = (double*)malloc(sizeof(double) * 3 * (size_t)(cnt > 0 ? cnt : 1));
        *jdx = (int*)malloc(sizeof(int) * (size_t)(cnt > 0 ? cnt : 1));
        TM_CHECK(*dxyz && *jdx, "out of host memory");
        TM_CUDA(cudaMemcpy(*dxyz, dd.p, sizeof(double) * 3 * (size_t)cnt, cudaMemcpyDeviceToHost));
        TM_CUDA(cudaMemcpy(*jdx, dj.p, sizeof(int) * (size_t)cnt, cudaMemcpyDeviceToHost));
        // neighbour ids back to host numbering, ascending
        std::vector<int> d2h(m->V), order(cnt);
        TM_CUDA(cudaMemcpy(d2h.data(), m->d2h, sizeof(int) * m->V, cudaMemcpyDeviceToHost));
        for (int k = 0; k < cnt; k++) { (*jdx)[k] = d2h[(*jdx)[k]]; order[k] = k; }
        const int* pj = *jdx;
        std::sort(order.begin(), order.end(), [pj](int x, int y) { return pj[x] < pj[y]; });
        std::vector<double> td(3 * (size_t)cnt);
        std::vector<int> tj(cnt);
        for (int k = 0; k < cnt; k++) {
            tj[k] = pj[order[k]];
            for (int c = 0; c < 3; c++) td[3 * (size_t)k + c] = (*dxyz)[3 * (size_t)order[k] + c];
        }
        if (cnt > 0) {
            memcpy(*dxyz, td.data(), sizeof(double) * 3 * (size_t)cnt);
            memcpy(*jdx, tj.data(), sizeof(int) * (size_t)cnt);
        }
        *n = cnt;
        return 0;
    }
    tm_set_error("point_distances: result did not fit");
    return 1;
}

}  // extern "C"
