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
        int b = nl.row_off[i], e = nl.row_off[i + 1];
        for (int k = b; k < e; k++) {
            int j = __ldg(nl.row_j + k);
            d3 d = sub_nofma(xi, ldpos(mv.x3, j));
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
                        d3 d = sub_nofma(xi, ldpos(mv.x3, j));
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
