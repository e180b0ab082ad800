// Neighbour-list traversal shared by the energy sweeps, the list queries and the flips.
#pragma once
#include "tm_common.cuh"

// Is j within `excl` rings of i on the current connectivity? (cell_list.h:167-201)
// CSR variant: reads the one-ring table (valid whenever the table is in sync with the
// half-edge arrays, i.e. everywhere except inside a flip sweep).
__device__ __forceinline__ bool excluded(const MeshView& mv, int i, int j, int excl) {
    if (excl <= 0) return false;
    int b = mv.ring_off[i], e = mv.ring_off[i + 1];
    for (int s = b; s < e; s++) {
        int v = mv.ring_v[s] & TM_RING_MASK;
        if (v == j) return true;
        if (excl > 1) {
            int b2 = mv.ring_off[v], e2 = mv.ring_off[v + 1];
            for (int t = b2; t < e2; t++)
                if ((mv.ring_v[t] & TM_RING_MASK) == j) return true;
        }
    }
    return false;
}

// is j in the sorted row [b, e)?
__device__ __forceinline__ bool in_sorted(const int* __restrict__ v, int b, int e, int j) {
    while (b < e) {
        const int mid = (b + e) >> 1;
        const int x = __ldg(v + mid);
        if (x == j) return true;
        if (x < j) b = mid + 1; else e = mid;
    }
    return false;
}

// Live variant: walks the half-edge arrays (h -> next(opposite(h))) exactly like the
// reference's vv_range circulators; used while flips mutate the connectivity.
__device__ __forceinline__ bool excluded_live(const MeshView& mv, int i, int j, int excl) {
    if (excl <= 0) return false;
    const int h0 = mv.v_he[i];
    if (h0 < 0) return false;
    int h = h0;
    do {
        const int v = mv.he_to[h];
        if (v == j) return true;
        if (excl > 1) {
            const int g0 = mv.v_he[v];
            if (g0 >= 0) {
                int g = g0;
                do {
                    if (mv.he_to[g] == j) return true;
                    g = mv.he_next[g ^ 1];
                } while (g != g0);
            }
        }
        h = mv.he_next[h ^ 1];
    } while (h != h0);
    return false;
}

// Calls f(j, d = x_i - x_j, l = |d|) for every neighbour-list candidate with l < dmax,
// with the semantics of CellList/VerletList::point_distances (cell_list.h:131-219,
// verlet_list.h:74-101): cell lists use build-time occupancy, the query vertex's live
// position and live exclusions; verlet lists filter the frozen rows.
template <bool LIVE, class F>
__device__ __forceinline__ void for_each_neighbour_t(const MeshView& mv, const NlView& nl, int i,
                                                     d3 xi, double dmax, F f) {
    if (nl.kind == TM_NLIST_VERLET) {
        const int n = nl.row_cnt[i];
        for (int k = 0; k < n; k++) {
            int j = __ldg(nl.row_j + tm_row_base(i, nl.row_cap) + k * 32);
            d3 d = sub_nofma(xi, ldpos(mv.pos, j));
            double l = dist_nofma(d);
            if (l < dmax) f(j, d, l);
        }
    } else if (nl.kind == TM_NLIST_CELL) {
        int c[3];
        c[0] = (int)((xi.x - nl.box_min[0]) / nl.cell[0]);
        c[1] = (int)((xi.y - nl.box_min[1]) / nl.cell[1]);
        c[2] = (int)((xi.z - nl.box_min[2]) / nl.cell[2]);
        for (int a = -1; a < 2; a++) {
            int ii = c[0] + a;
            if (ii < 0 || ii >= nl.shape[0]) continue;
            for (int b = -1; b < 2; b++) {
                int jj = c[1] + b;
                if (jj < 0 || jj >= nl.shape[1]) continue;
                for (int g = -1; g < 2; g++) {
                    int kk = c[2] + g;
                    if (kk < 0 || kk >= nl.shape[2]) continue;
                    int cell = ii * nl.strides[0] + jj * nl.strides[1] + kk * nl.strides[2];
                    int cb = nl.cell_start[cell], ce = nl.cell_start[cell + 1];
                    for (int k = cb; k < ce; k++) {
                        int j = nl.cell_items[k];
                        if (j == i) continue;
                        d3 d = sub_nofma(xi, ldpos(mv.pos, j));
                        double l = dist_nofma(d);
                        // the reference tests the exclusion first; the selected set is the same
                        if (l < dmax) {
                            bool ex = LIVE ? excluded_live(mv, i, j, nl.excl) : excluded(mv, i, j, nl.excl);
                            if (!ex) f(j, d, l);
                        }
                    }
                }
            }
        }
    }
}

template <class F>
__device__ __forceinline__ void for_each_neighbour(const MeshView& mv, const NlView& nl, int i,
                                                   d3 xi, double dmax, F f) {
    for_each_neighbour_t<false>(mv, nl, i, xi, dmax, f);
}


// Sweep variant: calls f(j, d = x_i - x_j, l2 = |d|^2) for every candidate with l2 < dmax2. FMA
// arithmetic and no square root: the caller re-tests l < lc1 on the rounded length, and a pair
// whose length rounds differently from the reference's non-FMA value lies within an ulp of the
// cut-off, where the potential and all its derivatives vanish (exp(l/(l-lc1)) -> 0), so the
// sums are unaffected. Cell lists keep the reference's candidate and exclusion semantics
// (cell_list.h:131-219): build-time occupancy, the query vertex's live position, live exclusions.
// REC: fetch the whole (x,y,z,rho) record through the read-only path (kernels that do not write pos).
template <bool LIVE, bool REC, class F>
__device__ __forceinline__ void for_each_pair_fast(const MeshView& mv, const NlView& nl, int i,
                                                   d3 xi, double dmax2, F f) {
    if (nl.kind == TM_NLIST_VERLET) {
        const int e = nl.row_cnt[i];
        if (e <= 0) return;
        // software pipeline of depth 1: the next neighbour's record is in flight while the
        // current pair is evaluated
        int jn = __ldg(nl.row_j + tm_row_base(i, nl.row_cap));
        d3 xn;
        if (REC) { const P4 r = ldrec(mv.pos, jn); xn = mk3(r.x, r.y, r.z); } else xn = ldpos(mv.pos, jn);
        for (int k = 0; k < e; k++) {
            const int j = jn;
            const d3 xj = xn;
            if (k + 1 < e) {
                jn = __ldg(nl.row_j + tm_row_base(i, nl.row_cap) + (k + 1) * 32);
                if (REC) { const P4 r = ldrec(mv.pos, jn); xn = mk3(r.x, r.y, r.z); } else xn = ldpos(mv.pos, jn);
            }
            const d3 d = xi - xj;
            const double l2 = dot3(d, d);
            if (l2 < dmax2) f(j, d, l2);
        }
    } else if (nl.kind == TM_NLIST_CELL && LIVE) {
        // Flip kernels (one thread evaluates one vertex, latency is everything): the candidates of a cell are
        // contiguous records in cell_rec (position and id, gathered at the current positions: tm_ensure_cellrec)
        // and x is the fastest cell index, so the 27 cells are 9 runs read four records per trip -- independent
        // loads -- instead of 150 times cell_items -> pos, two dependent loads per candidate. Same candidate set
        // and semantics as below; the order of the sum is that of the cell-list pair sweep (k_pairs).
        int c[3];
        c[0] = (int)((xi.x - nl.box_min[0]) / nl.cell[0]);
        c[1] = (int)((xi.y - nl.box_min[1]) / nl.cell[1]);
        c[2] = (int)((xi.z - nl.box_min[2]) / nl.cell[2]);
        const int x0 = max(c[0] - 1, 0), x1 = min(c[0] + 1, nl.shape[0] - 1);
        if (x0 > x1) return;
        for (int b = -1; b < 2; b++) {
            const int jj = c[1] + b;
            if (jj < 0 || jj >= nl.shape[1]) continue;
            for (int g = -1; g < 2; g++) {
                const int kk = c[2] + g;
                if (kk < 0 || kk >= nl.shape[2]) continue;
                const int base = jj * nl.strides[1] + kk * nl.strides[2];          // strides[0] == 1
                const int kb = __ldg(nl.cell_start + base + x0), ke = __ldg(nl.cell_start + base + x1 + 1);
                for (int k = kb; k < ke; k += 4) {
                    P4 x[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) x[u] = ldrec(nl.cell_rec, min(k + u, ke - 1));
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int j = __double2loint(x[u].w);
                        const d3 d = xi - mk3(x[u].x, x[u].y, x[u].z);
                        const double l2 = dot3(d, d);
                        if (k + u < ke && j != i && l2 < dmax2 && !excluded_live(mv, i, j, nl.excl)) f(j, d, l2);
                    }
                }
            }
        }
    } else if (nl.kind == TM_NLIST_CELL) {
        int c[3];
        c[0] = (int)((xi.x - nl.box_min[0]) / nl.cell[0]);
        c[1] = (int)((xi.y - nl.box_min[1]) / nl.cell[1]);
        c[2] = (int)((xi.z - nl.box_min[2]) / nl.cell[2]);
        for (int a = -1; a < 2; a++) {
            const int ii = c[0] + a;
            if (ii < 0 || ii >= nl.shape[0]) continue;
            for (int b = -1; b < 2; b++) {
                const int jj = c[1] + b;
                if (jj < 0 || jj >= nl.shape[1]) continue;
                for (int g = -1; g < 2; g++) {
                    const int kk = c[2] + g;
                    if (kk < 0 || kk >= nl.shape[2]) continue;
                    const int cell = ii * nl.strides[0] + jj * nl.strides[1] + kk * nl.strides[2];
                    const int cb = nl.cell_start[cell], ce = nl.cell_start[cell + 1];
                    for (int k = cb; k < ce; k++) {
                        const int j = nl.cell_items[k];
                        if (j == i) continue;
                        d3 xj;
                        if (REC) { const P4 r = ldrec(mv.pos, j); xj = mk3(r.x, r.y, r.z); } else xj = ldpos(mv.pos, j);
                        const d3 d = xi - xj;
                        const double l2 = dot3(d, d);
                        if (l2 < dmax2) {
                            const bool ex = LIVE ? excluded_live(mv, i, j, nl.excl) : excluded(mv, i, j, nl.excl);
                            if (!ex) f(j, d, l2);
                        }
                    }
                }
            }
        }
    }
}
