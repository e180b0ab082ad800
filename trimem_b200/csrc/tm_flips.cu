// Edge-flip sweeps on the device connectivity.
//
// Replaces, in the reference:
//   flip_serial / flip_parallel_batches      src/flips.cpp:22-165
//   edge_vertex_properties, flip_patch,
//   test_patch (+ OmpGuard try-locks)        src/flip_utils.cpp:19-114, src/omp_guard.cpp
//   TriConnectivity::is_flip_ok / flip       OpenMesh
//
// TM_FLIP_SERIAL: one persistent warp walks the proposals strictly in order (same
// arithmetic and the same order of updates as flip_serial); lanes 0-3 evaluate the
// four vertex patches of a proposal concurrently.
//
// TM_FLIP_BATCH: rounds of conflict-free proposals. Every remaining proposal claims
// the edges of its patch (star + link edges of its 4 vertices = flip_patch) with
// atomicMin(owner[edge], proposal index); a proposal that owns its whole patch is
// selected -- the index is a random priority because proposals are drawn at random,
// and conflicting proposals are therefore always processed in proposal order, like
// the reference's lock-or-skip rounds but without dropping proposals. Selected
// proposals evaluate their patch before/after in parallel (4 threads each), the
// flips are applied in parallel (patches are disjoint), and the Metropolis test
// runs on the device as an ordered scan over the round, because acceptance depends
// on the global energy through the non-linear penalties (the reference's
// `omp critical`, flips.cpp:142-160). Rejected flips are flipped back.
#include <climits>

#include "tm_common.cuh"
#include "tm_neigh.cuh"
#include "tm_vertex.cuh"

namespace {

constexpr int MAXV = 64;          // largest valence the flip kernels handle
constexpr int NT = 128;

struct Topo {                      // mutable connectivity
    int* he_to;
    int* he_next;
    int* he_face;
    int* v_he;
};

struct FlipWS {
    int cap = 0, E = 0;
    int *edges = nullptr, *owner = nullptr, *rem_a = nullptr, *rem_b = nullptr, *sel = nullptr;
    int *flips = nullptr, *counters = nullptr;   // counters: [0] n_rem_next [1] n_flips [2] n_acc [3] err
    int *selidx = nullptr;                       // proposals selected and flippable this round (counters[5] of them)
    int *patch = nullptr;                        // [PATCH_ROWS][cap]: the patch edges k_batch_mark visited, for k_batch_select
    double *u = nullptr, *before = nullptr, *after = nullptr, *state = nullptr;   // state: props[6], e0
};

// ---- connectivity primitives (volatile: other threads mutate these arrays) ---------
__device__ __forceinline__ int vld(const int* p) { return *(const volatile int*)p; }

__device__ bool d_is_flip_ok(const Topo& t, int e) {
    const int hh = 2 * e, oh = hh + 1;
    if (vld(t.he_face + hh) < 0 || vld(t.he_face + oh) < 0) return false;
    const int ah = vld(t.he_to + vld(t.he_next + hh));
    const int bh = vld(t.he_to + vld(t.he_next + oh));
    if (ah == bh) return false;
    const int h0 = vld(t.v_he + ah);
    if (h0 >= 0) {
        int h = h0, guard = 0;
        do {
            if (vld(t.he_to + h) == bh) return false;
            h = vld(t.he_next + (h ^ 1));
        } while (h != h0 && ++guard < 4096);
    }
    return true;
}

__device__ void d_flip(const Topo& t, int e) {
    const int a0 = 2 * e, b0 = a0 + 1;
    const int a1 = vld(t.he_next + a0), a2 = vld(t.he_next + a1);
    const int b1 = vld(t.he_next + b0), b2 = vld(t.he_next + b1);
    const int va0 = vld(t.he_to + a0), va1 = vld(t.he_to + a1);
    const int vb0 = vld(t.he_to + b0), vb1 = vld(t.he_to + b1);
    const int fa = vld(t.he_face + a0), fb = vld(t.he_face + b0);
    t.he_to[a0] = va1;
    t.he_to[b0] = vb1;
    t.he_next[a0] = a2; t.he_next[a2] = b1; t.he_next[b1] = a0;
    t.he_next[b0] = b2; t.he_next[b2] = a1; t.he_next[a1] = b0;
    t.he_face[a1] = fb;
    t.he_face[b1] = fa;
    if (vld(t.v_he + va0) == b0) t.v_he[va0] = a1;
    if (vld(t.v_he + vb0) == a0) t.v_he[vb0] = b1;
}

// the 4 vertices of edge_vertex_properties, in its order (flip_utils.cpp:27-38)
__device__ __forceinline__ int patch_vertex(const Topo& t, int e, int q) {
    const int heh = 2 * e + (q >> 1);
    return (q & 1) ? vld(t.he_to + vld(t.he_next + heh)) : vld(t.he_to + heh);
}

// clockwise one-ring of v from the live half-edge arrays; returns the valence or -1
__device__ int collect_ring(const Topo& t, int v, int* rw) {
    const int h0 = vld(t.v_he + v);
    if (h0 < 0) return 0;
    int h = h0, d = 0;
    do {
        if (d >= MAXV) return -1;
        int w = vld(t.he_to + h);
        if (vld(t.he_face + h) >= 0) w |= TM_RING_FACE;
        if (vld(t.he_face + (h ^ 1)) >= 0) w |= TM_RING_OFACE;
        rw[d++] = w;
        h = vld(t.he_next + (h ^ 1));
    } while (h != h0);
    return d;
}

__device__ bool live_vertex_props(const MeshView& mv, const Topo& t, const NlView& nl, const DevParams& P,
                                  int v, double out[6]) {
    int rw[MAXV];
    const int d = collect_ring(t, v, rw);
    if (d < 0) return false;
    vertex_props<true, true>(mv, nl, P, v, d, [&rw](int s) { return rw[s]; }, out);
    return true;
}

// The same evaluation by a whole warp (all 32 lanes call it with the same v; every lane receives the six values).
// One thread per patch vertex is pure latency -- ring walk, ring records, ~150 cell-list candidates, each a chain of
// dependent loads: ~150 us per evaluation at 40,962 vertices, profiles/r02u -- so the lanes split the work: the ring
// is walked once (every lane follows the same chain; lane s keeps the word of slot s), lane s evaluates slot s with
// its neighbours' spokes taken by shuffle, the list candidates are dealt out lane by lane, and the sums are fixed
// butterflies. Per-slot and per-pair arithmetic is that of vertex_props / for_each_pair_fast<LIVE>; only the order
// of the sums differs (tree instead of ring order). Valence > 32: lane 0 evaluates serially.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ bool warp_live_vertex_props(const MeshView& mv, const Topo& t, const NlView& nl, const DevParams& P,
                                       int v, double out[6]) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // ring of v: every lane walks (same addresses: one transaction per step), lane s keeps slot s
    int d = 0, word = 0;
    bool big = false;
    {
        const int h0 = vld(t.v_he + v);
        if (h0 >= 0) {
            int h = h0;
            do {
                if (d >= 32) { big = true; break; }
                int w = vld(t.he_to + h);
                if (vld(t.he_face + h) >= 0) w |= TM_RING_FACE;
                if (vld(t.he_face + (h ^ 1)) >= 0) w |= TM_RING_OFACE;
                if (lane == d) word = w;
                ++d;
                h = vld(t.he_next + (h ^ 1));
            } while (h != h0);
        }
    }
    if (big) {
        int ok = 1;
        if (lane == 0) ok = live_vertex_props(mv, t, nl, P, v, out) ? 1 : 0;
        for (int q = 0; q < 6; q++) out[q] = __shfl_sync(FULL, out[q], 0);
        return __shfl_sync(FULL, ok, 0) != 0;
    }
    const d3 xi = ldpos(mv.pos, v);
    double area = 0, vol = 0, curv = 0, teth = 0;
    if (d > 0) {
        const bool on = lane < d;
        const d3 e = on ? ldpos(mv.pos, word & TM_RING_MASK) - xi : mk3(0, 0, 0);
        const int prv = lane == 0 ? d - 1 : lane - 1, nxt = lane + 1 >= d ? 0 : lane + 1;
        const d3 e_p = mk3(__shfl_sync(FULL, e.x, prv), __shfl_sync(FULL, e.y, prv), __shfl_sync(FULL, e.z, prv));
        const d3 e_n = mk3(__shfl_sync(FULL, e.x, nxt), __shfl_sync(FULL, e.y, nxt), __shfl_sync(FULL, e.z, nxt));
        if (on && (word & TM_RING_FACE)) {                     // slot `lane`, as vertex_props::step
            const d3 n = cross3(e, e_p), n_n = cross3(e_n, e);
            const double l = sqrt_nn(dot3(e, e));
            const double a_f = 0.5 * sqrt_nn(dot3(n, n));
            area = a_f;
            vol = dot3(xi, n);
            if (word & TM_RING_OFACE) curv = l * tm_math::atan2_fast(-l * dot3(e_n, n), dot3(n, n_n));
            if (P.bond_type == TM_BOND_EDGE) { if (l > P.lc0 || l < P.lc1) teth = tether_edge(l, P); }
            else { const double q = a_f / P.a0 - 1.0; teth = q * q; }
        }
        area = warp_sum(area); vol = warp_sum(vol); curv = warp_sum(curv); teth = warp_sum(teth);
    }
    area = area / 3.0;
    vol = vol / 18.0;
    curv = curv / 4.0;
    teth = teth / (P.bond_type == TM_BOND_EDGE ? 2.0 : 3.0);
    double rep = 0.0;
    if (nl.kind >= 0 && P.rep_lc1 > 0.0) {
        const double lc1 = P.rep_lc1, dmax2 = lc1 * lc1;
        const int rr = P.rep_r;
        double sl;
        if (nl.kind == TM_NLIST_VERLET) {
            const int n = nl.row_cnt[v];
            const int* col = nl.row_j + tm_row_base(v, nl.row_cap);
            for (int k = lane; k < n; k += 32) {
                const d3 dv = xi - ldpos(mv.pos, __ldg(col + k * 32));
                const double l2 = dot3(dv, dv);
                if (l2 < dmax2) rep += repulsion_pair<false>(l2, lc1, rr, sl);
            }
        } else {
            // cell list: build-time occupancy, live position of v, live exclusions (cell_list.h:131-219); the 27
            // cells are 9 runs of contiguous records (tm_ensure_cellrec), dealt out to the lanes
            int c[3];
            c[0] = (int)((xi.x - nl.box_min[0]) / nl.cell[0]);
            c[1] = (int)((xi.y - nl.box_min[1]) / nl.cell[1]);
            c[2] = (int)((xi.z - nl.box_min[2]) / nl.cell[2]);
            const int x0 = max(c[0] - 1, 0), x1 = min(c[0] + 1, nl.shape[0] - 1);
            // lane r < 9 fetches the bounds of run r; every lane then walks all runs with its own stride
            int kb = 0, ke = 0;
            if (lane < 9 && x0 <= x1) {
                const int jj = c[1] + lane / 3 - 1, kk = c[2] + lane % 3 - 1;
                if (jj >= 0 && jj < nl.shape[1] && kk >= 0 && kk < nl.shape[2]) {
                    const int base = jj * nl.strides[1] + kk * nl.strides[2];          // strides[0] == 1
                    kb = __ldg(nl.cell_start + base + x0);
                    ke = __ldg(nl.cell_start + base + x1 + 1);
                }
            }
            for (int r = 0; r < 9; r++) {
                const int b = __shfl_sync(FULL, kb, r), e = __shfl_sync(FULL, ke, r);
                for (int k = b + lane; k < e; k += 32) {
                    const P4 x = ldrec(nl.cell_rec, k);
                    const int j = __double2loint(x.w);
                    const d3 dv = xi - mk3(x.x, x.y, x.z);
                    const double l2 = dot3(dv, dv);
                    if (j != v && l2 < dmax2 && !excluded_live(mv, v, j, nl.excl)) rep += repulsion_pair<false>(l2, lc1, rr, sl);
                }
            }
        }
        rep = warp_sum(rep);
    }
    out[0] = area; out[1] = vol; out[2] = curv;
    out[3] = 2.0 * curv * curv / area;
    out[4] = teth; out[5] = rep;
    return true;
}

// ---- TM_FLIP_SERIAL ------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_flip_serial(MeshView mv, Topo t, NlView nl, DevParams P, int n, const int* __restrict__ edges,
              const double* __restrict__ u, double* __restrict__ state, int* __restrict__ flips,
              int* __restrict__ counters) {
    // four warps, one per patch vertex (warp_live_vertex_props); thread 0 keeps the running state and takes the
    // decisions, in the reference's order of operations (flips.cpp:22-80)
    __shared__ double s_vp[4][6];
    __shared__ int s_ok, s_err;
    const int tid = threadIdx.x, wq = tid >> 5, lane = tid & 31;
    double props[6] = {0, 0, 0, 0, 0, 0};
    double e0 = 0.0;
    int nflips = 0, nacc = 0;
    if (tid == 0) {
        for (int q = 0; q < 6; q++) props[q] = state[q];
        e0 = energy_of(P, props);
        s_err = 0;
    }
    __syncthreads();
    auto evaluate = [&](int e, double (&sum)[6]) {
        double vp[6];
        const bool fine = warp_live_vertex_props(mv, t, nl, P, patch_vertex(t, e, wq), vp);
        if (lane == 0) {
            for (int q = 0; q < 6; q++) s_vp[wq][q] = vp[q];
            if (!fine) s_err = 1;
        }
        __syncthreads();
        if (tid == 0)                       // props{0} += vp(v0) += vp(v1) ... in the reference's order
            for (int q = 0; q < 6; q++) sum[q] = ((s_vp[0][q] + s_vp[1][q]) + s_vp[2][q]) + s_vp[3][q];
        __syncthreads();
    };
    for (int p = 0; p < n; p++) {
        const int e = edges[p];
        if (tid == 0) s_ok = d_is_flip_ok(t, e) ? 1 : 0;     // also false on boundary edges
        __syncthreads();
        const int ok = s_ok;
        __syncthreads();
        if (!ok) continue;
        double oprops[6] = {0, 0, 0, 0, 0, 0}, nprops[6] = {0, 0, 0, 0, 0, 0};
        evaluate(e, oprops);
        if (tid == 0) { d_flip(t, e); __threadfence_block(); }
        __syncthreads();
        evaluate(e, nprops);
        if (tid == 0) {
            // flips.cpp:49-76
            for (int q = 0; q < 6; q++) { props[q] -= oprops[q]; props[q] += nprops[q]; }
            const double en = energy_of(P, props);
            const double de = en - e0;
            const double alpha = de < 0.0 ? 1.0 : exp(-de);
            flips[nflips++] = e;
            if (u[p] <= alpha) {
                e0 = en;
                ++nacc;
            } else {
                d_flip(t, e);
                flips[nflips++] = e;
                for (int q = 0; q < 6; q++) { props[q] -= nprops[q]; props[q] += oprops[q]; }
                __threadfence_block();
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        for (int q = 0; q < 6; q++) state[q] = props[q];
        state[6] = e0;
        counters[1] = nflips; counters[2] = nacc; counters[3] = s_err;
    }
}

// ---- TM_FLIP_BATCH ---------------------------------------------------------------------
// visit every edge of flip_patch(e) (flip_utils.cpp:43-68)
template <class F>
__device__ __forceinline__ void for_each_patch_edge(const Topo& t, int e, F f) {
    for (int q = 0; q < 4; q++) {
        const int v = patch_vertex(t, e, q);
        const int h0 = vld(t.v_he + v);
        if (h0 < 0) continue;
        int h = h0, guard = 0;
        do {
            f(h >> 1);
            f(vld(t.he_next + h) >> 1);
            h = vld(t.he_next + (h ^ 1));
        } while (h != h0 && ++guard < 4096);
    }
}

// Four threads per remaining proposal, one per patch vertex (the four ring walks are chains of dependent half-edge
// loads: side by side they take a quarter of the time). The edges visited are kept -- rows 0..3 of `patch` hold the
// four vertices' visit counts, rows 4 + 16 q + n the edges, slot-major so that a warp's stores coalesce -- for
// k_batch_select, which then reads a list instead of repeating the walks; a vertex with more than PATCH_Q visits
// (valence > 8) is walked again there.
constexpr int PATCH_Q = 16;
constexpr int PATCH_ROWS = 4 + 4 * PATCH_Q;
__global__ void k_batch_mark(Topo t, int nrem, const int* __restrict__ rem, const int* __restrict__ edges,
                             int* __restrict__ owner, int* __restrict__ patch, int cap) {
    const int gid = blockIdx.x * NT + threadIdx.x;
    const int r = gid >> 2, q = gid & 3;
    if (r >= nrem) return;
    const int p = rem[r];
    const int v = patch_vertex(t, edges[p], q);
    int n = 0;
    const int h0 = vld(t.v_he + v);
    if (h0 >= 0) {
        int h = h0, guard = 0;
        do {
            const int e1 = h >> 1, e2 = vld(t.he_next + h) >> 1;
            atomicMin(owner + e1, p);
            atomicMin(owner + e2, p);
            if (n + 1 < PATCH_Q) {
                patch[(size_t)(4 + PATCH_Q * q + n) * cap + r] = e1;
                patch[(size_t)(4 + PATCH_Q * q + n + 1) * cap + r] = e2;
            }
            n += 2;
            h = vld(t.he_next + (h ^ 1));
        } while (h != h0 && ++guard < 4096);
    }
    patch[(size_t)q * cap + r] = n;
}

// sel[p]: 0 = not selected this round, 1 = selected and flippable, 2 = selected, not flippable
__global__ void k_batch_select(Topo t, int nrem, const int* __restrict__ rem, const int* __restrict__ edges,
                               const int* __restrict__ owner, int* __restrict__ sel, const int* __restrict__ patch, int cap,
                               int* __restrict__ nsel, int* __restrict__ selidx) {
    const int r = blockIdx.x * NT + threadIdx.x;
    if (r >= nrem) return;
    const int p = rem[r];
    const int e = edges[p];
    bool mine = true, listed = true;
    for (int q = 0; q < 4; q++) {
        const int n = patch[(size_t)q * cap + r];
        if (n > PATCH_Q) { listed = false; break; }
        for (int k = 0; k < n; k++) mine &= (owner[patch[(size_t)(4 + PATCH_Q * q + k) * cap + r]] == p);
    }
    if (!listed) {
        mine = true;
        for_each_patch_edge(t, e, [&](int ed) { if (owner[ed] != p) mine = false; });
    }
    const int sv = !mine ? 0 : (d_is_flip_ok(t, e) ? 1 : 2);
    sel[p] = sv;
    if (sv == 1) selidx[atomicAdd(nsel, 1)] = p;          // (any order: the evaluations are independent)
}

// One block per selected proposal (compacted by k_batch_select into selidx, *nsel of them; the grid is sized for
// every remaining proposal and the surplus blocks leave at once), one warp per patch vertex.
__global__ void __launch_bounds__(128)
k_batch_props(MeshView mv, Topo t, NlView nl, DevParams P, const int* __restrict__ nsel, const int* __restrict__ selidx,
              const int* __restrict__ edges, double* __restrict__ out, int* __restrict__ counters) {
    if ((int)blockIdx.x >= *nsel) return;
    __shared__ double s_vp[4][6];
    const int p = selidx[blockIdx.x];
    const int wq = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double vp[6];
    const bool fine = warp_live_vertex_props(mv, t, nl, P, patch_vertex(t, edges[p], wq), vp);
    if (lane == 0) {
        for (int c = 0; c < 6; c++) s_vp[wq][c] = vp[c];
        if (!fine) counters[3] = 1;
    }
    __syncthreads();
    // the 4 vertices summed in the reference's order
    if (threadIdx.x < 6) {
        const int c = threadIdx.x;
        out[6 * (size_t)p + c] = ((s_vp[0][c] + s_vp[1][c]) + s_vp[2][c]) + s_vp[3][c];
    }
}

__global__ void k_batch_apply(Topo t, int nrem, const int* __restrict__ rem, const int* __restrict__ edges,
                              const int* __restrict__ sel, int want) {
    const int r = blockIdx.x * NT + threadIdx.x;
    if (r >= nrem) return;
    const int p = rem[r];
    if (sel[p] == want) d_flip(t, edges[p]);
}

// trimem_energy (kernel.h:110-122) with the reference sums' reciprocals multiplied in: three independent short
// chains instead of three IEEE divisions on the ordered scan's critical path (the value differs from energy_of's
// in the last bit; the batch sweep has no decision-for-decision counterpart in the reference to stay aligned with)
__device__ __forceinline__ double energy_mul(const DevParams& P, const double* s) {
    const double da = fma(s[0], P.inv_ref_area, -1.0), dv = fma(s[1], P.inv_ref_volume, -1.0),
                 dc = fma(s[2], P.inv_ref_curvature, -1.0);
    double e = P.kappa_a * da * da;
    e = fma(P.kappa_v * dv, dv, e);
    e = fma(P.kappa_c * dc, dc, e);
    e = fma(P.kappa_t, s[4], e);
    e = fma(P.kappa_r, s[5], e);
    return fma(P.kappa_b, s[3], e);
}

// Ordered Metropolis scan over the round; also compacts the not-yet-processed proposals for the next round and
// appends the applied / reverted flips to the replay list.
//
// One block. The decisions are inherently sequential (every acceptance changes the global sums the next energy
// difference is taken against) but everything around them is not: per tile of MT remaining proposals all threads
// fetch their proposal's data at once, the carried-over proposals are compacted into rem_next and the selected ones
// -- a few percent of a round -- into shared memory, both in proposal order (ballot scans); thread 0 then walks the
// compacted tile: ~25 dependent fp64 operations and no global load per decision. (Round 2 measured the earlier
// one-warp version, which paid a chain of dependent global loads per 32 proposals whether selected or not, at
// 0.8 ms for 12,288 remaining proposals: 8 of the sweep's 18 ms. profiles/r02t_pflip_launches.csv)
constexpr int MT = 512;

// exclusive prefix count of `flag` over the block in thread order; total returned to every thread
__device__ __forceinline__ int block_rank(bool flag, int* s_warp /* MT/32 + 1 */, int& total) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned b = __ballot_sync(FULL, flag);
    __syncthreads();                                   // s_warp may still be read from the previous call
    if (lane == 0) s_warp[w] = __popc(b);
    __syncthreads();
    int base = 0, tot = 0;
    for (int k = 0; k < MT / 32; k++) {
        const int c = s_warp[k];
        base += k < w ? c : 0;
        tot += c;
    }
    total = tot;
    return base + __popc(b & ((1u << lane) - 1));
}

__global__ void __launch_bounds__(MT)
k_batch_metropolis(DevParams P, int nrem, const int* __restrict__ rem, int* __restrict__ rem_next,
                   const double* __restrict__ u, const int* __restrict__ edges,
                   int* __restrict__ sel, const double* __restrict__ before,
                   const double* __restrict__ after, double* __restrict__ state,
                   int* __restrict__ flips, int* __restrict__ counters) {
    if (blockIdx.x != 0) return;
    __shared__ double s_d[6][MT];
    __shared__ double s_lu[MT];
    __shared__ int s_ed[MT], s_p[MT];
    __shared__ int s_warp[MT / 32 + 1];
    const int tid = threadIdx.x;
    int nflips = counters[1], nnext = 0;
    // replay list of this round: first every applied flip, in proposal order ...
    for (int r0 = 0; r0 < nrem; r0 += MT) {
        const int r = r0 + tid;
        const int p = r < nrem ? rem[r] : -1;
        const bool applied = p >= 0 && sel[p] == 1;
        int tot;
        const int pos = block_rank(applied, s_warp, tot);
        if (applied) flips[nflips + pos] = edges[p];
        nflips += tot;
    }
    // ... then the decisions (thread 0 carries the running state)
    double props[6] = {0, 0, 0, 0, 0, 0};
    double e0 = 0.0;
    int nacc = 0;
    if (tid == 0) {
        for (int q = 0; q < 6; q++) props[q] = state[q];
        e0 = state[6];
        nacc = counters[2];
    }
    for (int r0 = 0; r0 < nrem; r0 += MT) {
        const int r = r0 + tid;
        const int p = r < nrem ? rem[r] : -1;
        const int s = p >= 0 ? sel[p] : -1;
        // not selected this round: carried over, order kept
        int tot0, tot1;
        const int pos0 = block_rank(s == 0, s_warp, tot0);
        if (s == 0) rem_next[nnext + pos0] = p;
        nnext += tot0;
        const int pos1 = block_rank(s == 1, s_warp, tot1);
        if (s == 1) {
            // flips.cpp:133-160: dprops = before - after
            for (int q = 0; q < 6; q++) s_d[q][pos1] = before[6 * (size_t)p + q] - after[6 * (size_t)p + q];
            s_lu[pos1] = log(u[p]);                      // accept iff log(u) <= -dE: no exponential in the chain
            s_ed[pos1] = edges[p];
            s_p[pos1] = p;
        }
        __syncthreads();
        if (tid == 0) {
            for (int k = 0; k < tot1; k++) {
                double dk[6];
                for (int q = 0; q < 6; q++) { dk[q] = s_d[q][k]; props[q] -= dk[q]; }    // props -= dprops
                const double en = energy_mul(P, props);
                const double de = en - e0;
                if (de < 0.0 || s_lu[k] <= -de) {
                    e0 = en;
                    ++nacc;
                } else {
                    for (int q = 0; q < 6; q++) props[q] += dk[q];
                    sel[s_p[k]] = 3;                     // rejected: flip back
                    flips[nflips++] = s_ed[k];           // ... then every reverted flip
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        for (int q = 0; q < 6; q++) state[q] = props[q];
        state[6] = e0;
        counters[0] = nnext; counters[1] = nflips; counters[2] = nacc;
    }
}

__global__ void k_iota(int n, int* a) {
    const int i = blockIdx.x * NT + threadIdx.x;
    if (i < n) a[i] = i;
}

__global__ void k_init_state(DevParams P, double* state) {
    if (threadIdx.x == 0 && blockIdx.x == 0) state[6] = energy_mul(P, state);
}

__global__ void k_evp(MeshView mv, Topo t, NlView nl, DevParams P, int e, double* out, int* err) {
    const int lane = threadIdx.x;
    double vp[6] = {0, 0, 0, 0, 0, 0};
    if (lane < 4 && !live_vertex_props(mv, t, nl, P, patch_vertex(t, e, lane), vp)) *err = 1;
    for (int q = 0; q < 6; q++) {
        double s = 0.0;
        for (int l = 0; l < 4; l++) s += __shfl_sync(0xffffffffu, vp[q], l);
        if (lane == 0) out[q] = s;
    }
}

__global__ void k_flip_ok(Topo t, int e, int* out) {
    if (threadIdx.x == 0) *out = d_is_flip_ok(t, e) ? 1 : 0;
}
__global__ void k_flip_one(Topo t, int e) {
    if (threadIdx.x == 0) d_flip(t, e);
}

Topo topo_of(tm_mesh* m) { return Topo{m->he_to, m->he_next, m->he_face, m->v_he}; }

int ensure_ws(tm_mesh* m, int n) {
    FlipWS* ws = (FlipWS*)m->flipws;
    if (!ws) { ws = new FlipWS(); m->flipws = ws; }
    const int E = m->HE / 2;
    if (ws->E != E) {
        cudaFree(ws->owner); ws->owner = nullptr;
        TM_CUDA(cudaMalloc(&ws->owner, sizeof(int) * (size_t)(E > 0 ? E : 1)));
        ws->E = E;
    }
    if (ws->cap < n) {
        cudaFree(ws->edges); cudaFree(ws->rem_a); cudaFree(ws->rem_b); cudaFree(ws->sel); cudaFree(ws->flips);
        cudaFree(ws->u); cudaFree(ws->before); cudaFree(ws->after); cudaFree(ws->patch); cudaFree(ws->selidx);
        ws->cap = 0;
        size_t c = (size_t)n;
        TM_CUDA(cudaMalloc(&ws->edges, sizeof(int) * c));
        TM_CUDA(cudaMalloc(&ws->rem_a, sizeof(int) * c));
        TM_CUDA(cudaMalloc(&ws->rem_b, sizeof(int) * c));
        TM_CUDA(cudaMalloc(&ws->sel, sizeof(int) * c));
        TM_CUDA(cudaMalloc(&ws->flips, sizeof(int) * 2 * c));
        TM_CUDA(cudaMalloc(&ws->u, sizeof(double) * c));
        TM_CUDA(cudaMalloc(&ws->before, sizeof(double) * 6 * c));
        TM_CUDA(cudaMalloc(&ws->after, sizeof(double) * 6 * c));
        TM_CUDA(cudaMalloc(&ws->patch, sizeof(int) * PATCH_ROWS * c));
        TM_CUDA(cudaMalloc(&ws->selidx, sizeof(int) * c));
        ws->cap = n;
    }
    if (!ws->counters) TM_CUDA(cudaMalloc(&ws->counters, sizeof(int) * 8));
    if (!ws->state) TM_CUDA(cudaMalloc(&ws->state, sizeof(double) * 8));
    return 0;
}

int check(tm_mesh* m, const tm_energy_params* p) {
    TM_CHECK(m != nullptr && m->V > 0 && m->he_to_h, "mesh has no topology (call tm_mesh_set_topology first)");
    TM_CHECK(p != nullptr, "null parameter block");
    TM_CHECK(p->bond_type == TM_BOND_EDGE || p->bond_type == TM_BOND_AREA, "Unknown bond potential");
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_join_pairs(m)) return 1;
    return tm_ensure_layout(m);
}

}  // namespace

int tm_flips_release(tm_mesh* m) {
    FlipWS* ws = (FlipWS*)m->flipws;
    if (!ws) return 0;
    cudaFree(ws->edges); cudaFree(ws->owner); cudaFree(ws->rem_a); cudaFree(ws->rem_b); cudaFree(ws->sel);
    cudaFree(ws->flips); cudaFree(ws->counters); cudaFree(ws->u); cudaFree(ws->before); cudaFree(ws->after);
    cudaFree(ws->state); cudaFree(ws->patch); cudaFree(ws->selidx);
    delete ws;
    m->flipws = nullptr;
    return 0;
}

extern "C" {

// `mutated` is set as soon as a kernel that may change the device connectivity has been enqueued
static int flip_sweep_body(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, int mode, int n,
                           const int* edges, const double* u, int* n_accepted, int* n_flips, int* flipped,
                           double props_io[6], bool* mutated) {
    if (check(m, p)) return 1;
    TM_CHECK(mode == TM_FLIP_SERIAL || mode == TM_FLIP_BATCH, "unknown flip mode %d", mode);
    TM_CHECK(n >= 0 && props_io && n_accepted, "bad arguments");
    *n_accepted = 0;
    if (n_flips) *n_flips = 0;
    if (n == 0) return 0;
    TM_CHECK(edges && u, "null proposal arrays");
    const int E = m->HE / 2;
    for (int k = 0; k < n; k++) TM_CHECK(edges[k] >= 0 && edges[k] < E, "proposal %d: edge %d out of range", k, edges[k]);
    // cell lists evaluate exclusions on the live connectivity: keep the ring table usable for
    // everything outside the sweep, and walk half-edges inside it
    if (tm_check_nlist(m, nl) || ensure_ws(m, n)) return 1;
    TM_CHECK(!nl || nl->kind != TM_NLIST_VERLET || nl->rows_count == m->V,
             "flip sweeps read the rows of arbitrary vertices: this verlet list holds a partition's slice only "
             "(use a cell list, or a list built before tm_part_create)");
    // partitioned mesh: the sweep runs replicated on every rank over the whole mesh (same proposals, same
    // decisions, same connectivity everywhere), so every rank first needs every position (collective)
    if (tm_part_make_current(m)) return 1;
    if (nl && tm_ensure_cellrec(m, nl->view())) return 1;      // the live evaluations scan the cell-ordered records
    FlipWS* ws = (FlipWS*)m->flipws;
    cudaStream_t st = m->stream;
    const DevParams P = tm_to_dev(*p);
    const NlView nv = nl ? nl->view() : tm_no_nlist();
    const MeshView mv = m->view();
    const Topo t = topo_of(m);

    TM_CUDA(cudaMemcpyAsync(ws->edges, edges, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    TM_CUDA(cudaMemcpyAsync(ws->u, u, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    TM_CUDA(cudaMemcpyAsync(ws->state, props_io, sizeof(double) * 6, cudaMemcpyHostToDevice, st));
    TM_CUDA(cudaMemsetAsync(ws->counters, 0, sizeof(int) * 8, st));
    int hc[4] = {0, 0, 0, 0};
    *mutated = true;

    if (mode == TM_FLIP_SERIAL) {
        k_flip_serial<<<1, 128, 0, st>>>(mv, t, nv, P, n, ws->edges, ws->u, ws->state, ws->flips, ws->counters);
        TM_LAUNCHED();
    } else {
        k_init_state<<<1, 32, 0, st>>>(P, ws->state); TM_LAUNCHED();
        k_iota<<<(n + NT - 1) / NT, NT, 0, st>>>(n, ws->rem_a); TM_LAUNCHED();
        int nrem = n;
        int *rem = ws->rem_a, *rem_next = ws->rem_b;
        while (nrem > 0) {
            const int nb = (nrem + NT - 1) / NT;
            TM_CUDA(cudaMemsetAsync(ws->owner, 0x7f, sizeof(int) * (size_t)ws->E, st));
            k_batch_mark<<<(4 * nrem + NT - 1) / NT, NT, 0, st>>>(t, nrem, rem, ws->edges, ws->owner, ws->patch, ws->cap); TM_LAUNCHED();
            TM_CUDA(cudaMemsetAsync(ws->counters + 5, 0, sizeof(int), st));
            k_batch_select<<<nb, NT, 0, st>>>(t, nrem, rem, ws->edges, ws->owner, ws->sel, ws->patch, ws->cap,
                                             ws->counters + 5, ws->selidx); TM_LAUNCHED();
            k_batch_props<<<nrem, 128, 0, st>>>(mv, t, nv, P, ws->counters + 5, ws->selidx, ws->edges, ws->before, ws->counters); TM_LAUNCHED();
            k_batch_apply<<<nb, NT, 0, st>>>(t, nrem, rem, ws->edges, ws->sel, 1); TM_LAUNCHED();
            k_batch_props<<<nrem, 128, 0, st>>>(mv, t, nv, P, ws->counters + 5, ws->selidx, ws->edges, ws->after, ws->counters); TM_LAUNCHED();
            k_batch_metropolis<<<1, MT, 0, st>>>(P, nrem, rem, rem_next, ws->u, ws->edges, ws->sel, ws->before,
                                                 ws->after, ws->state, ws->flips, ws->counters); TM_LAUNCHED();
            k_batch_apply<<<nb, NT, 0, st>>>(t, nrem, rem, ws->edges, ws->sel, 3); TM_LAUNCHED();
            TM_CUDA(cudaMemcpyAsync(hc, ws->counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
            TM_CUDA(cudaStreamSynchronize(st));
            TM_CHECK(hc[0] < nrem, "flip sweep made no progress (%d proposals left)", nrem);
            nrem = hc[0];
            int* tmp = rem; rem = rem_next; rem_next = tmp;
        }
    }
    TM_CUDA(cudaGetLastError());
    TM_CUDA(cudaMemcpyAsync(hc, ws->counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
    TM_CUDA(cudaMemcpyAsync(props_io, ws->state, sizeof(double) * 6, cudaMemcpyDeviceToHost, st));
    TM_CUDA(cudaStreamSynchronize(st));
    TM_CHECK(hc[3] == 0, "flip sweep met a vertex of valence > %d", MAXV);
    if (flipped && hc[1] > 0)
        TM_CUDA(cudaMemcpy(flipped, ws->flips, sizeof(int) * hc[1], cudaMemcpyDeviceToHost));
    if (n_flips) *n_flips = hc[1];
    *n_accepted = hc[2];
    return 0;
}

int tm_flip_sweep(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, int mode, int n,
                  const int* edges, const double* u, int* n_accepted, int* n_flips, int* flipped,
                  double props_io[6]) {
    bool mutated = false;
    const int rc = flip_sweep_body(m, nl, p, mode, n, edges, u, n_accepted, n_flips, flipped, props_io, &mutated);
    if (mutated) {
        // Success or failure: the connectivity on the device is no longer what the derived tables describe.
        // (After a FAILED sweep it is a half-finished one: the caller must upload its own connectivity again
        // with tm_mesh_set_topology before the next evaluation -- the host classes do, energy_manager.cpp.)
        cudaStreamSynchronize(m->stream);
        m->rings_valid = false;                 // the ring table is rebuilt lazily,
        m->raw_stale = true;                    // the host-numbered copies on request,
        m->pos_epoch++;                         // and cached per-vertex properties are void
        m->topo_epoch++;
    }
    return rc;
}

int tm_edge_vertex_properties(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, int edge, double out[6]) {
    if (check(m, p)) return 1;
    TM_CHECK(edge >= 0 && edge < m->HE / 2, "edge %d out of range", edge);
    if (tm_check_nlist(m, nl) || ensure_ws(m, 1)) return 1;
    TM_CHECK(!nl || nl->kind != TM_NLIST_VERLET || nl->rows_count == m->V,
             "this verlet list holds a partition's slice only");
    FlipWS* ws = (FlipWS*)m->flipws;
    if (nl && tm_ensure_cellrec(m, nl->view())) return 1;
    TM_CUDA(cudaMemsetAsync(ws->counters, 0, sizeof(int) * 8, m->stream));
    k_evp<<<1, 32, 0, m->stream>>>(m->view(), topo_of(m), nl ? nl->view() : tm_no_nlist(), tm_to_dev(*p), edge,
                                   ws->state, ws->counters + 3);
    TM_LAUNCHED();
    int err = 0;
    TM_CUDA(cudaMemcpyAsync(out, ws->state, sizeof(double) * 6, cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaMemcpyAsync(&err, ws->counters + 3, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    TM_CHECK(err == 0, "vertex of valence > %d", MAXV);
    return 0;
}

int tm_mesh_is_flip_ok(tm_mesh* m, int edge, int* ok) {
    TM_CHECK(m != nullptr && m->V > 0 && m->he_to_h && ok, "bad arguments");
    TM_CHECK(edge >= 0 && edge < m->HE / 2, "edge %d out of range", edge);
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_ensure_layout(m) || ensure_ws(m, 1)) return 1;
    FlipWS* ws = (FlipWS*)m->flipws;
    k_flip_ok<<<1, 32, 0, m->stream>>>(topo_of(m), edge, ws->counters + 4); TM_LAUNCHED();
    TM_CUDA(cudaMemcpyAsync(ok, ws->counters + 4, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

int tm_mesh_flip(tm_mesh* m, int edge) {
    TM_CHECK(m != nullptr && m->V > 0 && m->he_to_h, "bad arguments");
    TM_CHECK(edge >= 0 && edge < m->HE / 2, "edge %d out of range", edge);
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_ensure_layout(m)) return 1;
    k_flip_one<<<1, 32, 0, m->stream>>>(topo_of(m), edge); TM_LAUNCHED();
    TM_CUDA(cudaStreamSynchronize(m->stream));
    m->rings_valid = false;
    m->raw_stale = true;
    m->pos_epoch++;
    m->topo_epoch++;
    return 0;
}

}  // extern "C"
