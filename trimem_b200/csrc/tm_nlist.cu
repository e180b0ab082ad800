// Neighbour lists on the device: cell grid (one-pass counting sort on the cell
// key) and Verlet rows (single-pass pair search over 27 cells into fixed-width,
// column-major rows; exclusion by binary search in sorted two-ring rows).
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

#include <cmath>

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
    if (c < 0) return;                              // outside a clipped grid
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

// ---- verlet build ---------------------------------------------------------------------
// The list is the symmetrised distance_matrix(rlist) of a cell list built from the same
// positions (verlet_list.h:18-35): all pairs with dist <= rlist that are not topologically
// excluded. The pair SET does not depend on the cell geometry (any grid with cells >= rlist
// finds the same pairs), so the build uses its own tight grid; distances use the reference's
// non-FMA arithmetic so that the <= decision is bit-identical.
struct VGrid {
    double lo[3], inv[3];
    int n[3];
    // clip: only the vertices inside [lo, hi] are binned (a partitioned mesh bins what its owned rows can reach)
    double hi[3];
    int clip;
};
__device__ __forceinline__ void vcell(const VGrid& g, d3 p, int c[3]) {
    c[0] = min(g.n[0] - 1, max(0, (int)((p.x - g.lo[0]) * g.inv[0])));
    c[1] = min(g.n[1] - 1, max(0, (int)((p.y - g.lo[1]) * g.inv[1])));
    c[2] = min(g.n[2] - 1, max(0, (int)((p.z - g.lo[2]) * g.inv[2])));
}
__global__ void k_vcell_count(VGrid g, const P4* __restrict__ pos, int V, int* __restrict__ cell_of,
                              int* __restrict__ count) {
    int i = blockIdx.x * NT + threadIdx.x;
    if (i >= V) return;
    int c[3];
    const d3 p = ldpos(pos, i);
    if (g.clip && (p.x < g.lo[0] || p.x > g.hi[0] || p.y < g.lo[1] || p.y > g.hi[1] || p.z < g.lo[2] || p.z > g.hi[2])) {
        cell_of[i] = -1;
        return;
    }
    vcell(g, p, c);
    const int id = (c[2] * g.n[1] + c[1]) * g.n[0] + c[0];
    cell_of[i] = id;
    atomicAdd(count + id, 1);
}

// cell-ordered copy of the positions for the search: rec[k] = (x, y, z of items[k], id)
__global__ void k_vgather(int V, const int* __restrict__ n_items, const int* __restrict__ items, const P4* __restrict__ pos,
                          P4* __restrict__ rec) {
    const int k = blockIdx.x * NT + threadIdx.x;
    if (k >= V || k >= *n_items) return;           // (n_items: the binned vertices, = V unless the grid is clipped)
    const int j = items[k];
    const d3 p = ldpos(pos, j);
    P4 r;
    r.x = p.x; r.y = p.y; r.z = p.z;
    r.w = __hiloint2double(0, j);
    rec[k] = r;
}

// The search, shaped for the memory system and the issue slots (first version: one thread per vertex over
// 27 cells with an index -> record chain, a square root per candidate and a binary search per hit, 776 us
// at 655,362 vertices, plus two sort kernels; now ~400 us in all):
//  * the candidates of a cell are contiguous records (k_vgather), x is the fastest cell index, so the 27
//    cells are 9 runs read four records per trip with the next trip's in flight (one memory level);
//  * `dist <= rlist` is decided on the squared distance, accumulated exactly like the reference's norm
//    (no FMA): s2max is the largest double whose correctly rounded square root is <= rlist (found on the
//    host), so the decision is bit-identical and the 150 square roots per vertex disappear;
//  * candidates inside the radius are queued per thread in shared memory, the exclusion test compares four
//    of them per pass with the vertex's two-ring (or ring) row held in shared memory, and the survivors go
//    to the vertex's row in scan order.
constexpr int VS_THREADS = 128, VS_QCAP = 64, VS_XCAP = 32;
__global__ void __launch_bounds__(VS_THREADS, 4)
k_verlet_search2(MeshView mv, VGrid g, const int* __restrict__ cell_start, const P4* __restrict__ rec,
                 int excl, double s2max, int cap, int* __restrict__ row_j, int* __restrict__ row_cnt,
                 int* __restrict__ maxcnt, SweepRange R) {
    __shared__ int q_s[VS_QCAP][VS_THREADS];
    __shared__ int x_s[VS_XCAP][VS_THREADS];
    const int tid = threadIdx.x;
    const int t = blockIdx.x * VS_THREADS + tid;
    const int i = R.first + t;
    int n = 0;
    if (t < R.count) {
        const d3 xi = ldpos(mv.pos, i);
        int c[3];
        vcell(g, xi, c);
        int xb = 0, xe = 0;
        const int* xsrc = nullptr;
        if (excl == 2) { xb = mv.ex2_off[i]; xe = mv.ex2_off[i + 1]; xsrc = mv.ex2_v; }
        else if (excl == 1) { xb = mv.ring_off[i]; xe = mv.ring_off[i + 1]; xsrc = mv.ring_v; }
        const int xlen = xe - xb;
        const bool xfast = xlen <= VS_XCAP;
        if (xfast)
            for (int s = 0; s < xlen; s++) x_s[s][tid] = __ldg(xsrc + xb + s) & TM_RING_MASK;
        int* row = row_j + tm_row_base(i, cap);
        // queue layout: [0, filt) candidates that passed the exclusion test, [filt, cnt) not yet tested
        int cnt = 0, filt = 0;
        bool spilled = false;
        auto filter = [&]() {
            int keep = filt;
            for (int k = filt; k < cnt; k += 4) {
                int j[4];
                bool ex[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { j[u] = (k + u < cnt) ? q_s[k + u][tid] : -1; ex[u] = false; }
                if (xfast) {
                    for (int s = 0; s < xlen; s++) {
                        const int v = x_s[s][tid];
#pragma unroll
                        for (int u = 0; u < 4; u++) ex[u] |= (v == j[u]);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (j[u] < 0) continue;
                        if (excl == 2) ex[u] = in_sorted(xsrc, xb, xe, j[u]);
                        else for (int s = xb; s < xe; s++) ex[u] |= ((__ldg(xsrc + s) & TM_RING_MASK) == j[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (j[u] >= 0 && !ex[u]) { q_s[keep][tid] = j[u]; ++keep; }
            }
            filt = cnt = keep;
        };
        auto make_room = [&]() {
            filter();
            if (filt == VS_QCAP) {                 // a row longer than the queue: continue in global memory
                for (int k = 0; k < filt; k++) { if (n < cap) row[n * 32] = q_s[k][tid]; ++n; }
                filt = cnt = 0;
                spilled = true;
            }
        };
        auto flush = make_room;
        const int x0 = max(c[0] - 1, 0), x1 = min(c[0] + 1, g.n[0] - 1);
        for (int dz = -1; dz < 2; dz++) {
            const int cz = c[2] + dz;
            if (cz < 0 || cz >= g.n[2]) continue;
            for (int dy = -1; dy < 2; dy++) {
                const int cy = c[1] + dy;
                if (cy < 0 || cy >= g.n[1]) continue;
                const int base = (cz * g.n[1] + cy) * g.n[0];
                const int kb = __ldg(cell_start + base + x0), ke = __ldg(cell_start + base + x1 + 1);
                auto test = [&](const P4& x) {
                    const d3 d = sub_nofma(xi, mk3(x.x, x.y, x.z));
                    double s2 = __dmul_rn(d.x, d.x);
                    s2 = __dadd_rn(s2, __dmul_rn(d.y, d.y));
                    s2 = __dadd_rn(s2, __dmul_rn(d.z, d.z));
                    if (s2 <= s2max) {
                        const int j = __double2loint(x.w);
                        if (j != i) {
                            if (cnt == VS_QCAP) flush();
                            q_s[cnt][tid] = j;
                            ++cnt;
                        }
                    }
                };
                int k = kb;
                // whole groups of four without per-candidate bounds logic (two groups in flight), then the tail
#pragma unroll 2
                for (; k + 4 <= ke; k += 4) {
                    P4 x[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) x[u] = ldrec(rec, k + u);
#pragma unroll
                    for (int u = 0; u < 4; u++) test(x[u]);
                }
                for (; k < ke; k++) test(ldrec(rec, k));
            }
        }
        filter();
        if (!spilled) {
            // the whole row sits in the queue: ascending ids (= ascending along the space-filling order,
            // which keeps the pair sweep's gathers on few cache lines) by insertion in shared memory
            for (int a = 1; a < filt; a++) {
                const int v = q_s[a][tid];
                int b = a - 1;
                while (b >= 0 && q_s[b][tid] > v) { q_s[b + 1][tid] = q_s[b][tid]; --b; }
                q_s[b + 1][tid] = v;
            }
            for (int k = 0; k < filt; k++) { if (n < cap) row[n * 32] = q_s[k][tid]; ++n; }
        } else {
            for (int k = 0; k < filt; k++) { if (n < cap) row[n * 32] = q_s[k][tid]; ++n; }
            const int m = n < cap ? n : cap;
            for (int a = 1; a < m; a++) {
                const int v = row[a * 32];
                int b = a - 1;
                while (b >= 0 && row[b * 32] > v) { row[(b + 1) * 32] = row[b * 32]; --b; }
                row[(b + 1) * 32] = v;
            }
        }
        row_cnt[i] = n < cap ? n : cap;
    }
    // widest row of the build: the host re-runs with wider rows if it did not fit
    int w = n;
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if ((threadIdx.x & 31) == 0 && w > 0) atomicMax(maxcnt, w);
}

// positions at build time, and the largest displacement since (max is order-free: deterministic)
__global__ void k_save_positions(int V, const P4* __restrict__ pos, double* __restrict__ x3) {
    const int i = blockIdx.x * NT + threadIdx.x;
    if (i >= V) return;
    const d3 p = ldpos(pos, i);
    x3[3 * (size_t)i] = p.x; x3[3 * (size_t)i + 1] = p.y; x3[3 * (size_t)i + 2] = p.z;
}
__global__ void k_max_displacement2(int V, const P4* __restrict__ pos, const double* __restrict__ x3,
                                    unsigned long long* __restrict__ out) {
    double mx = 0.0;
    for (int i = blockIdx.x * NT + threadIdx.x; i < V; i += gridDim.x * NT) {
        const d3 p = ldpos(pos, i);
        const d3 d = mk3(p.x - x3[3 * (size_t)i], p.y - x3[3 * (size_t)i + 1], p.z - x3[3 * (size_t)i + 2]);
        mx = fmax(mx, dot3(d, d));
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // non-negative doubles order like their bit patterns
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(mx));
}

// number of directed entries: sum of row_cnt (fixed-order, integers)
__global__ void k_sum_int(const int* __restrict__ v, int n, unsigned long long* __restrict__ out) {
    unsigned long long s = 0;
    for (int i = blockIdx.x * NT + threadIdx.x; i < n; i += gridDim.x * NT) s += (unsigned long long)v[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
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
    for (int k = 0; k < nl.row_cnt[i]; k++)
        if (dist_nofma(sub_nofma(xi, ldpos(mv.pos, nl.row_j[tm_row_base(i, nl.row_cap) + k * 32]))) < dmax) ++n;
    cnt[i] = n;
}

__global__ void k_verlet_dm_fill(MeshView mv, NlView nl, double dmax, const int* __restrict__ off,
                                 double* __restrict__ dist, int* __restrict__ idx, int* __restrict__ jdx) {
    int i = blockIdx.x * NT + threadIdx.x;
    if (i >= mv.V) return;
    d3 xi = ldpos(mv.pos, i);
    int o = off[i];
    for (int k = 0; k < nl.row_cnt[i]; k++) {
        int j = nl.row_j[tm_row_base(i, nl.row_cap) + k * 32];
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
struct PoolBuf {
    T* p = nullptr;
    int dev;
    explicit PoolBuf(int d) : dev(d) {}
    ~PoolBuf() { tm_pool_free(dev, p); }
    int alloc(size_t n) { p = (T*)tm_pool_alloc(dev, (n ? n : 1) * sizeof(T)); return p ? 0 : 1; }
};

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

    nl->cell_start = (int*)tm_pool_alloc(m->device, sizeof(int) * (size_t)(ncell + 1));
    nl->cell_items = (int*)tm_pool_alloc(m->device, sizeof(int) * (size_t)V);
    nl->cell_rec = (P4*)tm_pool_alloc(m->device, sizeof(P4) * (size_t)V);
    TM_CHECK(nl->cell_rec, "out of device memory for the cell list");
    PoolBuf<int> count(m->device);
    int* cell_of = (int*)tm_pool_alloc(m->device, sizeof(int) * (size_t)V);
    *cell_of_out = cell_of;
    TM_CHECK(nl->cell_start && nl->cell_items && cell_of && !count.alloc((size_t)ncell + 1),
             "out of device memory for a cell grid of %lld cells", ncell);
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

// Verlet rows: tight grid -> counting sort of the vertices by cell -> single search pass.
int build_verlet(tm_mesh* m, tm_nlist* nl) {
    const int V = m->V, vb = (V + NT - 1) / NT, dev = m->device;
    if (nl->excl == 2 && tm_ensure_ring2(m)) return 1;
    double hb[6];
    if (tm_bbox(m, reinterpret_cast<const double*>(m->pos), 4, V, hb)) return 1;
    // The search grid. A partitioned mesh builds the rows of the vertices it owns, so its grid covers their bounding
    // box widened by rlist -- everything a row can list lies inside -- and the rest of the mesh is not binned: at
    // 10,485,762 vertices on 8 ranks that is an eighth of the 134 M cells to scan and a quarter of the vertices to
    // sort and gather, which is what kept the rebuild from scaling (round 1: 9.5 ms on one GPU, 2.8 ms on eight).
    const SweepRange Rg = tm_owned_range(m);
    double gb[6] = {hb[0], hb[1], hb[2], hb[3], hb[4], hb[5]};
    const bool clip = Rg.count > 0 && Rg.count < V;
    if (clip) {
        double ob[6];
        if (tm_bbox(m, reinterpret_cast<const double*>(m->pos + Rg.first), 4, Rg.count, ob)) return 1;
        for (int k = 0; k < 3; k++) {
            gb[k] = ob[k] - nl->rlist > hb[k] ? ob[k] - nl->rlist : hb[k];
            gb[3 + k] = ob[3 + k] + nl->rlist < hb[3 + k] ? ob[3 + k] + nl->rlist : hb[3 + k];
        }
    }
    VGrid g;
    g.clip = clip ? 1 : 0;
    long long ncell = 1;
    for (int k = 0; k < 3; k++) {
        const double dim = gb[3 + k] - gb[k];
        int nk = (int)(dim / nl->rlist);
        if (nk < 1) nk = 1;
        g.lo[k] = gb[k];
        g.hi[k] = gb[3 + k];
        g.n[k] = nk;
        g.inv[k] = dim > 0.0 ? nk / dim : 0.0;      // cell edge dim/nk >= rlist
        ncell *= nk;
    }
    {   // what tm_nlist_info reports is the geometry of the reference's temporary cell list
        // (cell_list.h:45-65), whatever grid the search itself uses
        const double init_min[3] = {DBL_MAX, 0.0, 0.0}, init_max[3] = {-1.0, 0.0, 0.0};
        for (int k = 0; k < 3; k++) {
            nl->box_min[k] = hb[k] < init_min[k] ? hb[k] : init_min[k];
            nl->box_max[k] = hb[3 + k] > init_max[k] ? hb[3 + k] : init_max[k];
            const double dim = nl->box_max[k] - nl->box_min[k] + nl->box_eps;
            const int nk = (int)(dim / nl->rlist);
            nl->cell[k] = (nk > 0) ? dim / nk : dim;
            nl->shape[k] = (nk > 0) ? nk : 1;
        }
    }
    TM_CHECK(ncell < (1LL << 30), "cell grid of %d x %d x %d cells is too large for the dense cell table "
             "(rlist=%g is tiny compared with the bounding box)", g.n[0], g.n[1], g.n[2], nl->rlist);
    nl->ncell = ncell;
    nl->strides[0] = 1; nl->strides[1] = nl->shape[0]; nl->strides[2] = nl->shape[0] * nl->shape[1];

    PoolBuf<int> count(dev), start(dev), items(dev), cell_of(dev), maxcnt(dev);
    TM_CHECK(!count.alloc((size_t)ncell + 1) && !start.alloc((size_t)ncell + 1) && !items.alloc(V) &&
             !cell_of.alloc(V) && !maxcnt.alloc(4), "out of device memory for a cell grid of %lld cells", ncell);
    TM_CUDA(cudaMemsetAsync(count.p, 0, sizeof(int) * (size_t)(ncell + 1), m->stream));
    TM_CUDA(cudaMemsetAsync(maxcnt.p, 0, sizeof(int) * 4, m->stream));
    k_vcell_count<<<vb, NT, 0, m->stream>>>(g, m->pos, V, cell_of.p, count.p); TM_LAUNCHED();
    if (tm_exclusive_scan(m, count.p, start.p, (int)ncell, nullptr)) return 1;
    TM_CUDA(cudaMemsetAsync(count.p, 0, sizeof(int) * (size_t)(ncell + 1), m->stream));
    k_cell_fill<<<vb, NT, 0, m->stream>>>(V, cell_of.p, start.p, count.p, items.p); TM_LAUNCHED();
    // (the order inside a cell is whatever the atomics made it: every row is sorted before it is stored)

    // largest squared distance whose correctly rounded square root is still <= rlist
    double s2max = nl->rlist * nl->rlist;
    while (std::sqrt(s2max) > nl->rlist) s2max = std::nextafter(s2max, 0.0);
    while (std::sqrt(std::nextafter(s2max, HUGE_VAL)) <= nl->rlist) s2max = std::nextafter(s2max, HUGE_VAL);
    PoolBuf<P4> rec(dev);
    TM_CHECK(!rec.alloc(V), "out of device memory for the search records");
    k_vgather<<<vb, NT, 0, m->stream>>>(V, start.p + ncell, items.p, m->pos, rec.p); TM_LAUNCHED();

    nl->row_cnt = (int*)tm_pool_alloc(dev, sizeof(int) * (size_t)V);
    TM_CHECK(nl->row_cnt, "out of device memory for verlet rows");
    // a partitioned mesh builds the rows of the vertices it owns (the only rows its sweeps read); the
    // cell grid above covers every vertex, so owned rows list non-owned neighbours as they should
    const SweepRange R = tm_owned_range(m);
    nl->rows_first = R.first; nl->rows_count = R.count;
    if (R.count < V) TM_CUDA(cudaMemsetAsync(nl->row_cnt, 0, sizeof(int) * (size_t)V, m->stream));
    int cap = m->verlet_cap_hint > 0 ? m->verlet_cap_hint : 32;
    for (int attempt = 0; attempt < 3; attempt++) {
        nl->row_cap = cap;
        nl->row_j = (int*)tm_pool_alloc(dev, sizeof(int) * (size_t)cap * 32 * ((V + 31) / 32));
        TM_CHECK(nl->row_j, "out of device memory for verlet rows of width %d", cap);
        if (R.count > 0)
            k_verlet_search2<<<(R.count + VS_THREADS - 1) / VS_THREADS, VS_THREADS, 0, m->stream>>>(
                m->view(), g, start.p, rec.p, nl->excl, s2max, cap, nl->row_j, nl->row_cnt, maxcnt.p, R);
        TM_LAUNCHED();
        TM_CUDA(cudaGetLastError());
        int widest = 0;
        TM_CUDA(cudaMemcpyAsync(&widest, maxcnt.p, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        TM_CUDA(cudaStreamSynchronize(m->stream));
        if (widest <= cap) {
            // next build of this mesh starts from the width that fitted (plus slack)
            m->verlet_cap_hint = ((widest + widest / 4 + 7) / 8) * 8;
            if (m->verlet_cap_hint < 8) m->verlet_cap_hint = 8;
            break;
        }
        TM_CHECK(attempt < 2, "verlet rows of %d entries did not fit", widest);
        tm_pool_free(dev, nl->row_j); nl->row_j = nullptr;
        cap = ((widest + 7) / 8) * 8;
    }
    nl->x_build = (double*)tm_pool_alloc(dev, sizeof(double) * 3 * (size_t)V);
    TM_CHECK(nl->x_build, "out of device memory for the list's build positions");
    k_save_positions<<<vb, NT, 0, m->stream>>>(V, m->pos, nl->x_build); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    nl->n_entries = -1;                 // counted on demand (tm_nlist_info)
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
    // partitioned mesh: the list is built from every rank's current positions (collective)
    if (tm_part_make_current(m)) return 1;

    tm_nlist* nl = new tm_nlist();
    nl->device = m->device;
    nl->kind = kind; nl->excl = exclusion_level; nl->rlist = rlist; nl->box_eps = box_eps; nl->V = m->V;
    nl->layout_epoch = m->layout_epoch;
    static std::atomic<unsigned long long> g_serial{0};      // (lists are created from several host threads)
    nl->serial = ++g_serial;
    if (kind == TM_NLIST_VERLET) {
        if (build_verlet(m, nl)) { tm_nlist_destroy(nl); return 1; }
        *out = nl;
        return 0;
    }
    int* cell_of = nullptr;
    if (build_cells(m, nl, &cell_of)) { tm_pool_free(m->device, cell_of); tm_nlist_destroy(nl); return 1; }
    nl->cell_of = cell_of;   // build-time cells, for distance_matrix
    *out = nl;
    return 0;
}

int tm_nlist_destroy(tm_nlist* nl) {
    if (!nl) return 0;
    cudaSetDevice(nl->device);
    tm_pool_free(nl->device, nl->cell_start); tm_pool_free(nl->device, nl->cell_items);
    tm_pool_free(nl->device, nl->cell_rec);
    tm_pool_free(nl->device, nl->row_cnt); tm_pool_free(nl->device, nl->row_j);
    tm_pool_free(nl->device, nl->cell_of);
    tm_pool_free(nl->device, nl->x_build);
    delete nl;
    return 0;
}

// Skin criterion of a verlet list (extension; the reference rebuilds on a fixed cadence,
// mc/evaluators.py:181-182): the frozen rows stay complete for a cut-off dmax as long as no vertex has moved
// further than (rlist - dmax) / 2 since the build. Returns the largest displacement on the mesh's
// current device positions; needs_rebuild = 2 * displacement > rlist - dmax.
}  // extern "C"

// asynchronous: out (device, 8 bytes) <- largest squared displacement of the mesh's current device positions from
// the list's build positions; 1 if the list keeps none (cell lists)
int tm_nlist_measure_displacement(tm_nlist* nl, tm_mesh* m, double* out) {
    if (!nl || nl->kind != TM_NLIST_VERLET || !nl->x_build || nl->layout_epoch != m->layout_epoch) return 1;
    if (cudaMemsetAsync(out, 0, 8, m->stream) != cudaSuccess) return 1;
    k_max_displacement2<<<296, NT, 0, m->stream>>>(m->V, m->pos, nl->x_build, reinterpret_cast<unsigned long long*>(out));
    TM_LAUNCHED();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" {

int tm_nlist_displacement(tm_nlist* nl, tm_mesh* m, double dmax, double* displacement, int* needs_rebuild) {
    TM_CHECK(nl != nullptr && m != nullptr, "null argument");
    TM_CHECK(nl->kind == TM_NLIST_VERLET && nl->x_build, "the skin criterion applies to verlet lists");
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_ensure_layout(m) || tm_check_nlist(m, nl) || tm_part_make_current(m)) return 1;
    unsigned long long* d = (unsigned long long*)tm_pool_alloc(nl->device, 8);
    TM_CHECK(d, "out of device memory");
    unsigned long long h = 0;
    cudaMemsetAsync(d, 0, 8, m->stream);
    k_max_displacement2<<<296, NT, 0, m->stream>>>(m->V, m->pos, nl->x_build, d); TM_LAUNCHED();
    cudaError_t e = cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    tm_pool_free(nl->device, d);
    TM_CHECK(e == cudaSuccess, "tm_nlist_displacement failed: %s", cudaGetErrorString(e));
    double d2;
    memcpy(&d2, &h, 8);
    const double disp = sqrt(d2);
    if (displacement) *displacement = disp;
    if (needs_rebuild) *needs_rebuild = (2.0 * disp > nl->rlist - dmax) ? 1 : 0;
    return 0;
}

int tm_nlist_info(tm_nlist* nl, long long* n_entries, int shape[3], double box[6], double cell[3]) {
    TM_CHECK(nl != nullptr, "null neighbour list");
    if (n_entries && nl->n_entries < 0 && nl->kind == TM_NLIST_VERLET) {
        TM_CUDA(cudaSetDevice(nl->device));
        unsigned long long* d = (unsigned long long*)tm_pool_alloc(nl->device, 8);
        TM_CHECK(d, "out of device memory");
        unsigned long long h = 0;
        cudaMemset(d, 0, 8);
        k_sum_int<<<256, NT>>>(nl->row_cnt, nl->V, d); TM_LAUNCHED();
        cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        tm_pool_free(nl->device, d);
        TM_CHECK(e == cudaSuccess, "tm_nlist_info failed: %s", cudaGetErrorString(e));
        nl->n_entries = (long long)h;
    }
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
