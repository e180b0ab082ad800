// Property sweep (K1+K2) and gradient sweep (K3+K4) of the Helfrich functional.
//
// Replaces, in the reference:
//   vertex_properties / vertex_properties_grad   src/mesh_properties.cpp:12-148
//   edge/area/volume/dihedral primitives + grads src/mesh_util.h:15-185
//   tether potentials                            src/mesh_tether.h:29-121
//   surface repulsion                            src/mesh_repulsion.h:28-79
//   penalty terms, parallel_for/reduction        src/kernel.h:19-137,243-262
//   EnergyManager::properties/energy/gradient    src/energy.cpp:63-124
//
// Formulation. One thread owns one vertex i and walks its
// one-ring j_0..j_{d-1} (clockwise). With e_s = x_{j_s} - x_i,
//   n_s   = e_s x e_{s-1}            un-normalised normal of face (i, j_s, j_{s-1})
//   p_s   = e_s . e_{s-1}
//   theta_s = atan2(-|e_s| (e_{s+1} . n_s), n_s . n_{s+1})   dihedral of spoke s
// everything the reference obtains from acos/tan of sector angles follows
// algebraically (cot = dot/|cross|), so a sweep needs one atan2, two rsqrt per
// ring slot and no other transcendental. The gradient is a gather: vertex i
// only writes g_i (no fp64 atomics) and needs the ring positions plus
// rho_v = c_v / a_v of itself and of its ring from the property sweep.
//
// Memory access. A thread fetches its whole ring as one 32-byte R8 row (coalesced
// across the warp), then issues all <= 8 position gathers back to back before any
// arithmetic: two dependent memory levels instead of three, and 6-8 loads in flight
// per thread instead of one. The ring loop is fully unrolled over the 8 slots of the
// row so the ring data sit in registers under static indices; vertices of valence
// > 8 (or < 3) take the generic CSR walk. Neighbour-list rows are consumed four
// entries at a time for the same reason, from fixed-width column-major rows (coalesced index loads).
//
// Kernels. The ring part (area, volume, curvature, tether; and their gradients) and the
// pair part (surface repulsion over the neighbour list) have different shapes -- the first
// is register-hungry straight-line fp64, the second a long gather loop -- so they are
// separate kernels, each at the occupancy that suits it. k_pairs evaluates the pair
// potential and, in the same pass, the pair force (they share the distance, the reciprocal
// and the exponential); the force is kept per vertex and the gradient sweep adds it, so an
// energy() followed by gradient() at the same positions evaluates every pair once.
//
// Reduction. Each warp reduces its six sums with shuffles and writes one partial
// per quantity; no block barrier, fence or atomic in the sweep. A small second
// kernel adds the warp partials in a fixed order (fixed chunks, fixed trees), so
// results are bit-reproducible run to run.
#include "tm_common.cuh"
#include "tm_neigh.cuh"
#include "tm_vertex.cuh"

namespace {

// ---- shared pieces -----------------------------------------------------------------
__device__ __forceinline__ int row_valence(const R8& r) {
    int d = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) d += (r.v[k] != TM_RING_PAD) ? 1 : 0;
    return d;
}

// slot s of a row held in registers (select chain: no reload of the row for a rare slot)
__device__ __forceinline__ int row_slot(const R8& r, int s) {
    int j = r.v[0];
#pragma unroll
    for (int k = 1; k < 8; k++) j = s == k ? r.v[k] : j;
    return j;
}

__device__ __forceinline__ R8 ldrow(const R8* __restrict__ rows, int i) {
    R8 r;
    const int4 a = __ldg(reinterpret_cast<const int4*>(rows + i));
    const int4 b = __ldg(reinterpret_cast<const int4*>(rows + i) + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}

// Verlet rows, four entries per trip: indices, then all four records, then the arithmetic.
// f(d = x_i - x_j, l2) for l2 < dmax2, in row order (fixed summation order).
template <bool REC, class F>
__device__ __forceinline__ void verlet_pairs(const P4* pos, const NlView& nl, int i, d3 xi, double dmax2, F f) {
    const int e = nl.row_cnt[i];
    const int* col = nl.row_j + tm_row_base(i, nl.row_cap);
    for (int k = 0; k < e; k += 4) {
        int j[4];
        j[0] = __ldg(col + k * 32);
#pragma unroll
        for (int u = 1; u < 4; u++) j[u] = (k + u < e) ? __ldg(col + (k + u) * 32) : j[0];
        d3 x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (REC) { const P4 r = ldrec(pos, j[u]); x[u] = mk3(r.x, r.y, r.z); }
            else x[u] = ldpos(pos, j[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (k + u < e) {
                const d3 dv = xi - x[u];
                const double l2 = dot3(dv, dv);
                if (l2 < dmax2) f(dv, l2);
            }
        }
    }
}

// ---- K1: per-vertex properties ------------------------------------------------------
// one ring slot of vertex_properties (mesh_properties.cpp:19-38): spoke e_cur, face
// (i, j_s, j_{s-1}) with normal n_cur, next spoke / normal for the dihedral of the spoke
template <bool FULL>
__device__ __forceinline__ void props_slot(const DevParams& P, d3 xi, int word, d3 e_cur, d3 n_cur, d3 e_next,
                                           d3 n_next, double& area, double& vol, double& curv, double& teth) {
    if (word & TM_RING_FACE) {                // `not he.is_boundary()`
        const double l = sqrt_nn(dot3(e_cur, e_cur));
        const double a_f = 0.5 * sqrt_nn(dot3(n_cur, n_cur));
        area += a_f;
        vol += dot3(xi, n_cur);               // = (x_j x x_k) . x_i
        if (word & TM_RING_OFACE) {           // dihedral is 0 on boundary edges
            const double th = tm_math::atan2_fast(-l * dot3(e_next, n_cur), dot3(n_cur, n_next));
            curv = fma(l, th, curv);
        }
        if (FULL) {
            if (P.bond_type == TM_BOND_EDGE) { if (l > P.lc0 || l < P.lc1) teth += tether_edge(l, P); }
            else { const double q = a_f / P.a0 - 1.0; teth += q * q; }
        }
    }
}

// ---- closed fans: the fast path of both ring sweeps -----------------------------------------------------
// What bounds these sweeps on B200 (profiles/r02): an fp64 instruction holds a scheduler's
// issue port for two cycles and every other instruction for one, and with 4-5 resident warps per scheduler the
// sweeps run at ~70 % of that issue bound whatever the memory system does (shared-memory tiles, cp.async
// pipelines and persistent grids were all measured: same time per instruction, more instructions). So the walk is
// written for INSTRUCTION COUNT: ring records straight from global memory into registers (one 256-bit load per
// neighbour, no staging, no barrier, no index arithmetic), the whole ring as one basic block instantiated per
// valence (no loop, no register rotation, all chains of the ring available to the scheduler at once), and a
// mask-free instantiation for warps whose closed fans all have the same valence (every warp of a regular mesh).
//
// Slot s owns spoke s and face s+1 (between spokes s and s+1; face d is face 0): with the un-normalised normal
// n_s handed over by slot s-1 it evaluates
//   n_{s+1} = e_{s+1} x e_s, |n_{s+1}|                          (area, volume, AREA bonds)
//   l_s, 1/l_s, theta_s = atan2(-l_s (e_{s+1} . n_s), n_s . n_{s+1})   (curvature, EDGE bonds)
// and hands tau_s = theta_s / l_s to the gradient sweep through the slot-major aux array (coalesced 8-byte
// store), so that sweep needs no arc tangent and a single reciprocal square root per face.
#ifndef TM_GRAD_GROUP
#define TM_GRAD_GROUP 3
#endif
#ifndef TM_GRAD_TAU_EARLY
#define TM_GRAD_TAU_EARLY 1
#endif
#ifndef TM_PROPS_GROUP
#define TM_PROPS_GROUP 3
#endif
struct PropsAcc {
    double area2, curv, teth;   // sum |n_f|, sum l theta, tether sum
    d3 nsum;                    // sum n_f
    unsigned tmask;             // slots whose EDGE bond is off the flat bottom
};

template <bool FULL>
struct PropsWalk {
    const DevParams& P;
    const P4* pos;
    const R8& row;
    d3 xi;
    int d;
    double* aux;                // tau array + i, or nullptr
    int vs;
    PropsAcc& A;

    // spoke to vertex j. Coherent load (rho (.w) of other vertices is being written by this very sweep and never
    // read by it), 256 bits in one instruction.
    __device__ __forceinline__ d3 spoke(int j) const {
        double x, y, z, w;
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(pos + j));
        return mk3(x, y, z) - xi;
    }
    // what one slot adds: spoke e_c with normal n_c of its face and squared length l2_c, next spoke e_n
    struct SlotOut { d3 n_n; double nl, lth, tau, l; };
    __device__ __forceinline__ SlotOut eval(const d3 e_c, const d3 n_c, const double l2_c, const d3 e_n) const {
        SlotOut o;
        o.n_n = cross3(e_n, e_c);
        const double nn = dot3(o.n_n, o.n_n);
        // (the 1e-300 offset is invisible for a proper face and makes a degenerate one come out as |n| = 0)
        const double in_n = rsqrt_pos(nn + 1.0e-300);
        const double il = rsqrt_pos(l2_c);
        o.nl = __dmul_rn(nn, in_n);                       // |n_{s+1}| = 2 A_f
        o.l = l2_c * il;
        const double th = tm_math::atan2_fast(-o.l * dot3(e_n, n_c), dot3(n_c, o.n_n));
        o.lth = __dmul_rn(o.l, th);
        o.tau = th * il;
        return o;
    }
    // (explicitly rounded sums: the FULL and the geometry-only instantiation must produce the same rho and
    // slot quantities bit for bit, whatever else the compiler would like to contract into them)
    template <bool MASK>
    __device__ __forceinline__ void add(const SlotOut& o, int s, bool valid) const {
        if (!MASK) valid = true;
        A.area2 = __dadd_rn(A.area2, valid ? o.nl : 0.0);
        A.curv = __dadd_rn(A.curv, valid ? o.lth : 0.0);
        A.nsum.x = __dadd_rn(A.nsum.x, valid ? o.n_n.x : 0.0);
        A.nsum.y = __dadd_rn(A.nsum.y, valid ? o.n_n.y : 0.0);
        A.nsum.z = __dadd_rn(A.nsum.z, valid ? o.n_n.z : 0.0);
        if (aux && valid) aux[s * vs] = o.tau;
        if (FULL) {
            // EDGE bonds leave the flat bottom on a fraction of a percent of the edges: those slots are
            // flagged and evaluated after the walk (one copy of the exp/pow code, outside the block)
            if (P.bond_type == TM_BOND_EDGE) A.tmask |= (valid && (o.l > P.lc0 || o.l < P.lc1)) ? (1u << s) : 0u;
            else { const double q = fma(0.5 * o.nl, P.inv_a0, -1.0); A.teth = __dadd_rn(A.teth, valid ? __dmul_rn(q, q) : 0.0); }
        }
    }
    // The whole ring in one basic block. DMAX = largest valence among the closed fans of the warp; MASK = the warp
    // has fans of lower valence too: such a lane evaluates the surplus slots on slot 0's record and leaves them out
    // of the sums, and finds its last slot (d-1, successor 0) through selects. Without MASK every index is static.
    template <int DMAX, bool MASK>
    __device__ __forceinline__ void block() const {
        int j[DMAX];
#pragma unroll
        for (int s = 0; s < DMAX; s++) j[s] = (!MASK || s < 3 || s < d) ? row.v[s] : row.v[0];
        int j_l = j[DMAX - 1], j_l2 = j[DMAX - 2];
        if (MASK) {
            j_l = row.v[2]; j_l2 = row.v[1];
#pragma unroll
            for (int s = 3; s < DMAX; s++) { j_l2 = s < d ? row.v[s - 1] : j_l2; j_l = s < d ? row.v[s] : j_l; }
        }
        const d3 e_l = spoke(j_l), e_l2 = spoke(j_l2);
        d3 e_c = spoke(j[0]);
        d3 n_c = cross3(e_c, e_l);
        {
            const SlotOut o = eval(e_l, cross3(e_l, e_l2), dot3(e_l, e_l), e_c);        // last slot; its n_n = n_0: face 0
            add<false>(o, MASK ? d - 1 : DMAX - 1, true);
        }
#pragma unroll
        for (int s = 0; s + 1 < DMAX; s++) {
            // (groups of TM_PROPS_GROUP slots between fences: keeps the scheduler from hoisting every load of the
            // ring to the top and spilling what it then has to hold)
            if (s > 0 && s % TM_PROPS_GROUP == 0) asm volatile("" ::: "memory");
            const d3 e_n = spoke(j[s + 1]);
            const SlotOut o = eval(e_c, n_c, dot3(e_c, e_c), e_n);
            add<MASK>(o, s, s + 1 < 3 || s + 1 < d);
            e_c = e_n;
            n_c = o.n_n;
        }
    }
};

// dmax: largest closed-fan valence in the warp; uniform: all of the warp's closed fans have it
template <bool FULL>
__device__ __forceinline__ void ring_props_closed(const MeshView& mv, const DevParams& P, const R8& row, int d, int dmax,
                                                  bool uniform, int i, d3 xi, bool write_aux, double& area, double& vol,
                                                  double& curv, double& teth) {
    PropsAcc A;
    A.area2 = 0.0; A.curv = 0.0; A.teth = 0.0; A.nsum = mk3(0, 0, 0); A.tmask = 0u;
    const PropsWalk<FULL> W{P, mv.pos, row, xi, d, write_aux ? mv.aux + i : nullptr, mv.aux_vs, A};
#define TM_WALK(D) do { if (uniform) W.template block<D, false>(); else W.template block<D, true>(); } while (0)
    switch (dmax) {                            // warp-uniform
        case 3: case 4: W.template block<4, true>(); break;
        case 5: TM_WALK(5); break;
        case 6: TM_WALK(6); break;
        case 7: TM_WALK(7); break;
        default: TM_WALK(8); break;
    }
#undef TM_WALK
    if (FULL) {
        unsigned tmask = A.tmask;
        while (tmask) {                            // ascending slot order; rare: the spoke is fetched again
            const int s = __ffs(tmask) - 1;
            tmask &= tmask - 1;
            const d3 e = ldpos(mv.pos, mv.ring8[i].v[s]) - xi;
            const double l2 = dot3(e, e);
            A.teth += tether_edge(l2 * rsqrt_pos(l2), P);
        }
    }
    area = 0.5 * A.area2;                      // sum of face areas
    vol = dot3(xi, A.nsum);                    // sum over faces of (x_j x x_k) . x_i
    curv = A.curv;
    teth = A.teth;
}

template <bool FULL>
__device__ __forceinline__ void ring_props_generic(const MeshView& mv, const DevParams& P, int i, d3 xi,
                                                   double& area, double& vol, double& curv, double& teth) {
    const int beg = mv.ring_off[i], d = mv.ring_off[i + 1] - beg;
    if (d <= 0) return;
    const int* ring = mv.ring_v + beg;
    int w_cur = ring[0];
    d3 e_cur = ldpos(mv.pos, w_cur & TM_RING_MASK) - xi;
    d3 n_cur = cross3(e_cur, ldpos(mv.pos, ring[d - 1] & TM_RING_MASK) - xi);
    for (int s = 0; s < d; s++) {
        const int w_next = ring[s + 1 == d ? 0 : s + 1];
        const d3 e_next = ldpos(mv.pos, w_next & TM_RING_MASK) - xi;
        const d3 n_next = cross3(e_next, e_cur);
        props_slot<FULL>(P, xi, w_cur, e_cur, n_cur, e_next, n_next, area, vol, curv, teth);
        w_cur = w_next; e_cur = e_next; n_cur = n_next;
    }
}

// FULL = also tether and repulsion (what energy()/properties() need); the property
// sweep inside gradient() only needs area, volume, curvature and rho.
//
// Two kernels: k_props walks the closed fans of the fixed-width rows (every vertex of a closed mesh of
// valence 3..8) and carries no slow path; k_props_generic runs over the short list of the remaining
// vertices (tm_rings.cu), only when there are any. Ghost sweeps of a partitioned mesh (explicit vertex
// list, nothing reduced) take the generic walk inline (GEN_INLINE) instead.
template <bool FULL, bool WRITE_VPROPS>
__device__ __forceinline__ void props_finish(const DevParams& P, int i, double area, double vol, double curv, double teth,
                                             P4* pos_rw, double* __restrict__ vprops, double (&acc)[6]) {
    area = area * (1.0 / 3.0);
    vol = vol * (1.0 / 18.0);                  // /6 per face, /3 multiplicity
    curv = curv * 0.25;                        // 0.5 l theta per half-edge, /2 multiplicity
    if (FULL) teth = teth * (P.bond_type == TM_BOND_EDGE ? 0.5 : (1.0 / 3.0));
    const double rho = area > 0.0 ? curv * rcp_pos(area) : curv / area;   // rho_i = c_i / a_i
    acc[0] = area; acc[1] = vol; acc[2] = curv;
    acc[3] = 2.0 * curv * rho;
    acc[4] = teth;
    pos_rw[i].w = rho;                         // rides in the position record
    if (WRITE_VPROPS) {
        double* o = vprops + 6 * (size_t)i;
#pragma unroll
        for (int q = 0; q < 5; q++) o[q] = acc[q];
    }
}

// One partial per warp and quantity. The six sums of a warp are reduced TOGETHER: at every level of the
// butterfly a lane keeps half of the quantities it still carries and hands the other half to its partner, so
// the tree takes 3 + 2 + 1 + 1 + 1 shuffles instead of 6 x 5 (the serial chain of shuffle latencies was a tenth
// of the sweep). Lane q of the first eight ends up with quantity q; fixed tree, bit-reproducible.
__device__ __forceinline__ void props_warp_partials(const double (&acc)[6], int nq, double* __restrict__ partials,
                                                    int ncols, int col) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    // level 16: lanes with bit 4 clear keep quantities {0,1,2,3}' low half ... written out for 8 slots (6 used)
    double v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = q < 6 ? acc[q] : 0.0;
    double w4[4];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const double send = hi ? v[k] : v[k + 4];
            const double got = __shfl_xor_sync(full, send, 16);
            w4[k] = (hi ? v[k + 4] : v[k]) + got;            // hi lanes carry quantities 4..7, lo lanes 0..3
        }
    }
    double w2[2];
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double send = hi ? w4[k] : w4[k + 2];
            const double got = __shfl_xor_sync(full, send, 8);
            w2[k] = (hi ? w4[k + 2] : w4[k]) + got;
        }
    }
    double w1;
    {
        const bool hi = lane & 4;
        const double send = hi ? w2[0] : w2[1];
        const double got = __shfl_xor_sync(full, send, 4);
        w1 = (hi ? w2[1] : w2[0]) + got;
    }
    w1 += __shfl_xor_sync(full, w1, 2);
    w1 += __shfl_xor_sync(full, w1, 1);
    // lane l (l & 3 == 0) holds quantity q = 4 * bit4 + 2 * bit3 + bit2
    const int q = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    if ((lane & 3) == 0 && q < nq && col < ncols) partials[(size_t)q * ncols + col] = w1;
}

// partials: [6][ncols] warp partials; this launch owns the columns [0, number of its warps). Quantity 5
// (repulsion) belongs to k_pairs and is written here (as 0) only when no pair sweep runs (nq == 6).
// LIST: the vertices of R.vlist (ghost sweeps of a partitioned mesh: rho only, nothing reduced, no slot
// quantities, generic vertices inline).
template <bool FULL, bool WRITE_VPROPS, bool LIST>
__global__ void __launch_bounds__(TM_PROPS_THREADS, TM_PROPS_MINB)
k_props(MeshView mv, DevParams P, P4* pos_rw, double* __restrict__ vprops,
        double* __restrict__ partials, int ncols, int nq, SweepRange R) {
    const int t = blockIdx.x * TM_PROPS_THREADS + threadIdx.x;
    const bool live = t < R.count;
    const int i = LIST ? (live ? R.vlist[t] : 0) : R.first + t;
    R8 row;
    row.v[0] = TM_RING_GENERIC;
    double xx = 0, xy = 0, xz = 0, xw = 0;
    if (live) {
        row = ldrow(mv.ring8, i);
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(xx), "=d"(xy), "=d"(xz), "=d"(xw) : "l"(mv.pos + i));
    }
    const bool closed = row.v[0] != TM_RING_GENERIC;
    const int d = closed ? row_valence(row) : 0;
    const int dmax = __reduce_max_sync(0xffffffffu, d);
    const bool uniform = __all_sync(0xffffffffu, !closed || d == dmax);
    double acc[6] = {0, 0, 0, 0, 0, 0};   // area volume curvature bending tether repulsion
    if (live) {
        const d3 xi = mk3(xx, xy, xz);
        double area = 0, vol = 0, curv = 0, teth = 0;
        if (closed) {
            ring_props_closed<FULL>(mv, P, row, d, dmax, uniform, i, xi, !LIST, area, vol, curv, teth);
            props_finish<FULL, WRITE_VPROPS>(P, i, area, vol, curv, teth, pos_rw, vprops, acc);
        } else if (LIST) {
            ring_props_generic<FULL>(mv, P, i, xi, area, vol, curv, teth);
            props_finish<FULL, WRITE_VPROPS>(P, i, area, vol, curv, teth, pos_rw, vprops, acc);
        }
    }
    if (!LIST) props_warp_partials(acc, nq, partials, ncols, t >> 5);
}

// the vertices of `list` (ascending) that fall into [R.first, R.first + R.count): columns col0.. of the partials
template <bool FULL, bool WRITE_VPROPS>
__global__ void __launch_bounds__(128)
k_props_generic(MeshView mv, DevParams P, P4* pos_rw, double* __restrict__ vprops, double* __restrict__ partials,
                int ncols, int col0, int nq, const int* __restrict__ list, int n, SweepRange R) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    if (t < n) {
        const int i = list[t];
        if (i >= R.first && i < R.first + R.count) {
            const d3 xi = ldpos(mv.pos, i);
            double area = 0, vol = 0, curv = 0, teth = 0;
            ring_props_generic<FULL>(mv, P, i, xi, area, vol, curv, teth);
            props_finish<FULL, WRITE_VPROPS>(P, i, area, vol, curv, teth, pos_rw, vprops, acc);
        }
    }
    props_warp_partials(acc, nq, partials, ncols, col0 + (t >> 5));
}

// ---- K1b: pair sweep: surface repulsion energy and force per vertex ------------------------
// (mesh_repulsion.h:28-79 over NeighbourList::point_distances). force[i] = sum_j slope_ij (x_i - x_j);
// the gradient sweep adds 2 kappa_r force[i] (mesh_properties.cpp:139-146).
//
// The inner loop is written for instruction count (the kernel is fp64-issue-bound): branch-free
// evaluation with the inactive lanes masked at the end (a warp executes an instruction whenever
// one lane needs it, so skipping buys nothing and the reconvergence bookkeeping costs), the
// exponential's coefficients and all loop invariants in registers, one 256-bit read-only load
// per neighbour.
struct PairConst {
    double lc1, lc1sq, rd;      // cut-off, its square, (double) r
    unsigned dl_hi_floor;       // high word of -(lc1 * 2^-20): |l - lc1| is kept above that (see pair_eval)
};

// phi = exp(l/(l-lc1)) / l^r and slope = -(lc1/(l-lc1)^2 + r/l) / l^(r+1) * exp(l/(l-lc1)) for l2 < lc1^2,
// both 0 for inactive entries. R2: r == 2 (the default exponent) without the power loop.
// `tab` = the 2^(j/64) table in shared memory (tm_math::exp_tab_clamped). Guards, all outside the fp64 pipe:
//  * l - lc1 is negative for an active entry except in the rounding corner l2 < lc1^2 <= l^2, and arbitrarily
//    small; its high word is raised (unsigned max: negative doubles order like their bit patterns) to that of
//    -(lc1 2^-20), which keeps t = l / (l - lc1) above ~ -2^20 -- inside the range of the exponential's
//    integer arithmetic -- and changes no entry whose exp(t) is not far below the smallest double anyway;
//  * the exponential clamps its power of two at 2^-1021 (integer max) instead of comparing t with -708:
//    such an entry contributes ~1e-308 where the reference's exp underflows to 0.
template <bool R2>
__device__ __forceinline__ void pair_eval(const PairConst& C, const double* tab, int rr, double l2, bool active,
                                          double& phi, double& sl) {
    const double l2s = active ? l2 : C.lc1sq * 0.25;          // any harmless in-range value
    const double il = rsqrt_pos(l2s);
    const double l = l2s * il;
    double dl = l - C.lc1;
    dl = __hiloint2double((int)max((unsigned)__double2hiint(dl), C.dl_hi_floor), __double2loint(dl));
    const double w = rcp_pos(dl);
    // exp(l * w), argument <= 0
    double ex = tm_math::exp_tab_clamped(l * w, tab);
    ex = active ? ex : 0.0;
    const double ilr = R2 ? il * il : powi(il, rr);
    phi = ex * ilr;
    sl = -fma(C.lc1 * w, w, R2 ? il + il : C.rd * il) * (phi * il);
}

// cell-ordered copy of the position records for the cell-list scan: rec[k] = (x, y, z of cell_items[k], id)
__global__ void k_cell_gather(int V, const int* __restrict__ items, const P4* __restrict__ pos, P4* __restrict__ rec) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= V) return;
    const int j = items[k];
    const d3 p = ldpos(pos, j);
    P4 r;
    r.x = p.x; r.y = p.y; r.z = p.z;
    r.w = __hiloint2double(0, j);
    rec[k] = r;
}

#ifndef TM_PAIR_NT
#define TM_PAIR_NT 128
#endif
#ifndef TM_PAIR_MINB
#define TM_PAIR_MINB 6
#endif
#ifndef TM_PAIR_BATCH
#define TM_PAIR_BATCH 3
#endif
constexpr int PAIR_THREADS = TM_PAIR_NT;
constexpr int PAIR_QCAP = 56;                   // queued candidates per thread before an early flush (cell lists)
constexpr int PAIR_XCAP = 32;                   // exclusion-row entries per thread kept in shared memory
constexpr int PB = TM_PAIR_BATCH;
#ifndef TM_PAIR_INTERLEAVE
#define TM_PAIR_INTERLEAVE 0
#endif

#ifndef TM_PAIR_CELL_MINB
#define TM_PAIR_CELL_MINB 3
#endif
// CELL: the list is a cell list (a separate instantiation, so that the verlet sweep keeps its own register
// allocation: 126 registers, no spills)
template <bool WITH_E, bool WITH_F, bool WRITE_VPROPS, bool R2, bool CELL>
__global__ void __launch_bounds__(PAIR_THREADS, CELL ? TM_PAIR_CELL_MINB : TM_PAIR_MINB)
k_pairs(MeshView mv, NlView nl, double lc1, int rr, double* __restrict__ force, double* __restrict__ vprops,
        double* __restrict__ partials, int nwarps, SweepRange R, int blk0) {
    __shared__ double s_exp2[64];
    if (threadIdx.x < 64) s_exp2[threadIdx.x] = tm_math::EXP2_64[threadIdx.x];
    __syncthreads();
    const int t = (blockIdx.x + blk0) * PAIR_THREADS + threadIdx.x;
    const int i = R.first + t;
    double rep = 0.0;
    if (t < R.count) {
        PairConst C;
        C.lc1 = lc1; C.lc1sq = lc1 * lc1; C.rd = (double)rr;
        C.dl_hi_floor = (unsigned)__double2hiint(-lc1) - (20u << 20);
        const P4 ri = ldrec(mv.pos, i);
        const d3 xi = mk3(ri.x, ri.y, ri.z);
        d3 gr = mk3(0, 0, 0);
        // Evaluation of a row of neighbour ids, entry k at col[k * stride]: a software pipeline, PB entries
        // per stage: while stage k is evaluated the records of stage k+1 are in flight and the indices of
        // stage k+2 have been requested, so the two dependent memory levels (index -> record, ~1000 cycles
        // each from a cold L2) overlap the ~80 instructions per entry instead of preceding them (100 -> 66 us
        // at 655,362 vertices, fp64 pipe 48 -> 70 % busy; profiles/r01g). Slots past the row's end read the
        // vertex's own record (always valid) and are masked in the evaluation.
        auto run_row = [&](const int* col, const int e, const int stride, const bool in_smem) {
            auto idx = [&](int k) { return in_smem ? col[k * stride] : __ldg(col + k * stride); };
            int jn[PB];
            P4 xn[PB];
#pragma unroll
            for (int u = 0; u < PB; u++) jn[u] = (u < e) ? idx(u) : i;
#pragma unroll
            for (int u = 0; u < PB; u++) xn[u] = ldrec(mv.pos, jn[u]);
#pragma unroll
            for (int u = 0; u < PB; u++) jn[u] = (PB + u < e) ? idx(PB + u) : i;
            for (int k = 0; k < e; k += PB) {
#if TM_PAIR_INTERLEAVE
                // Entry u's record is consumed (difference vector, squared distance) and its registers are reloaded
                // with the record of entry u of the NEXT stage at once: 71 instead of 80 instructions per entry, and
                // slower (68.7 against 6x us, profiles/r02r): the loads of a stage no longer go out together.
#pragma unroll
                for (int u = 0; u < PB; u++) {
                    const d3 dv = xi - mk3(xn[u].x, xn[u].y, xn[u].z);
                    const double l2 = dot3(dv, dv);
                    xn[u] = ldrec(mv.pos, jn[u]);
                    jn[u] = (k + 2 * PB + u < e) ? idx(k + 2 * PB + u) : i;
                    double phi, sl;
                    pair_eval<R2>(C, s_exp2, rr, l2, (k + u < e) && (l2 < C.lc1sq), phi, sl);
                    if (WITH_E) rep += phi;
                    if (WITH_F) axpy3(gr, sl, dv);
                }
#else
                P4 x[PB];
#pragma unroll
                for (int u = 0; u < PB; u++) x[u] = xn[u];
                if (k + PB < e) {
#pragma unroll
                    for (int u = 0; u < PB; u++) xn[u] = ldrec(mv.pos, jn[u]);
#pragma unroll
                    for (int u = 0; u < PB; u++) jn[u] = (k + 2 * PB + u < e) ? idx(k + 2 * PB + u) : i;
                }
#pragma unroll
                for (int u = 0; u < PB; u++) {
                    const d3 dv = xi - mk3(x[u].x, x[u].y, x[u].z);
                    const double l2 = dot3(dv, dv);
                    double phi, sl;
                    pair_eval<R2>(C, s_exp2, rr, l2, (k + u < e) && (l2 < C.lc1sq), phi, sl);
                    if (WITH_E) rep += phi;
                    if (WITH_F) axpy3(gr, sl, dv);
                }
#endif
            }
        };
        if (!CELL) {
            run_row(nl.row_j + tm_row_base(i, nl.row_cap), nl.row_cnt[i], 32, false);
        } else {
            // Cell list (CellList::point_distances, cell_list.h:131-219: build-time occupancy, the query
            // vertex's live cell, exclusions on the current connectivity): ~150 candidates per vertex, ~37
            // inside the cut-off, ~19 left after the exclusion test. Three phases per thread, each shaped
            // for what bounds it:
            //  A  scan: the candidates of a cell are CONTIGUOUS records (cell_rec, gathered once per sweep at
            //     the current positions, the vertex id in the fourth word) and x is the fastest cell index,
            //     so the 27 cells are 9 runs read four records per trip, the next trip's in flight: one
            //     memory level, no index -> record dependency. Candidates inside the cut-off are queued.
            //  B  exclusion: the vertex's sorted two-ring row (level 2) or ring row (level 1) sits in shared
            //     memory and is compared against four queued candidates per pass: no global latency.
            //  C  evaluation: the surviving queue is a verlet row (all lanes busy, scan order kept).
            extern __shared__ int pair_smem[];
            int* q = pair_smem + threadIdx.x;                                   // [PAIR_QCAP][PAIR_THREADS]
            int* xrow = pair_smem + PAIR_QCAP * PAIR_THREADS + threadIdx.x;     // [PAIR_XCAP][PAIR_THREADS]
            int xb = 0, xe = 0;
            const int* xsrc = nullptr;
            if (nl.excl == 2) { xb = mv.ex2_off[i]; xe = mv.ex2_off[i + 1]; xsrc = mv.ex2_v; }
            else if (nl.excl == 1) { xb = mv.ring_off[i]; xe = mv.ring_off[i + 1]; xsrc = mv.ring_v; }
            const int xlen = xe - xb;
            const bool xfast = xlen <= PAIR_XCAP;
            if (xfast)
                for (int s = 0; s < xlen; s++) xrow[s * PAIR_THREADS] = __ldg(xsrc + xb + s) & TM_RING_MASK;
            int cnt = 0;
            auto flush = [&]() {
                // B: drop the excluded candidates, order kept
                int keep = 0;
                for (int k = 0; k < cnt; k += 4) {
                    int j[4];
                    bool ex[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { j[u] = (k + u < cnt) ? q[(k + u) * PAIR_THREADS] : -1; ex[u] = false; }
                    if (xfast) {
                        for (int s = 0; s < xlen; s++) {
                            const int v = xrow[s * PAIR_THREADS];
#pragma unroll
                            for (int u = 0; u < 4; u++) ex[u] |= (v == j[u]);
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (j[u] < 0) continue;
                            if (nl.excl == 2) ex[u] = in_sorted(xsrc, xb, xe, j[u]);
                            else for (int s = xb; s < xe; s++) ex[u] |= ((__ldg(xsrc + s) & TM_RING_MASK) == j[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (j[u] >= 0 && !ex[u]) { q[keep * PAIR_THREADS] = j[u]; ++keep; }
                }
                // C
                run_row(q, keep, PAIR_THREADS, true);
                cnt = 0;
            };
            // A
            int c[3];
            c[0] = (int)((xi.x - nl.box_min[0]) / nl.cell[0]);
            c[1] = (int)((xi.y - nl.box_min[1]) / nl.cell[1]);
            c[2] = (int)((xi.z - nl.box_min[2]) / nl.cell[2]);
            const int x0 = max(c[0] - 1, 0), x1 = min(c[0] + 1, nl.shape[0] - 1);
            if (x0 <= x1) {
                for (int b = -1; b < 2; b++) {
                    const int jj = c[1] + b;
                    if (jj < 0 || jj >= nl.shape[1]) continue;
                    for (int g = -1; g < 2; g++) {
                        const int kk = c[2] + g;
                        if (kk < 0 || kk >= nl.shape[2]) continue;
                        const int base = jj * nl.strides[1] + kk * nl.strides[2];      // strides[0] == 1
                        const int kb = __ldg(nl.cell_start + base + x0), ke = __ldg(nl.cell_start + base + x1 + 1);
                        if (kb >= ke) continue;
                        P4 xn[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) xn[u] = ldrec(nl.cell_rec, min(kb + u, ke - 1));
                        for (int k = kb; k < ke; k += 4) {
                            P4 x[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) x[u] = xn[u];
                            if (k + 4 < ke) {
#pragma unroll
                                for (int u = 0; u < 4; u++) xn[u] = ldrec(nl.cell_rec, min(k + 4 + u, ke - 1));
                            }
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const d3 d = xi - mk3(x[u].x, x[u].y, x[u].z);
                                const double l2 = dot3(d, d);
                                const int j = __double2loint(x[u].w);
                                if (k + u < ke && j != i && l2 < C.lc1sq) {
                                    if (cnt == PAIR_QCAP) flush();
                                    q[cnt * PAIR_THREADS] = j;
                                    ++cnt;
                                }
                            }
                        }
                    }
                }
            }
            flush();
        }
        if (WITH_F) {
            double* o = force + 3 * (size_t)i;
            o[0] = gr.x; o[1] = gr.y; o[2] = gr.z;
        }
        if (WRITE_VPROPS) vprops[6 * (size_t)i + 5] = rep;
    }
    if (WITH_E) {
        const int lane = threadIdx.x & 31, w = t >> 5;
        double s = rep;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && w < nwarps) partials[(size_t)5 * nwarps + w] = s;
    }
}

// ---- K2: fixed-order sum of the warp partials ---------------------------------------------
// grid of nb <= TM_REDUCE_BLOCKS blocks; block b owns the contiguous chunk [b*per, (b+1)*per) of
// warp partials; the last block to finish adds the nb block sums in index order.
constexpr int RT = 256;
__device__ __forceinline__ void block_sum6(double (&v)[6], double* out /* 6, valid in thread 0 */) {
    __shared__ double sm[6][RT / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 6; q++) {
        double t = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) sm[q][w] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 6; q++) {
            double t = 0.0;
            for (int k = 0; k < RT / 32; k++) t += sm[q][k];
            out[q] = t;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RT)
k_reduce_props(const double* __restrict__ partials, int nwarps, int per, int qlo, int qhi, int slo, int shi,
               double* __restrict__ block_sums, unsigned* __restrict__ ticket, double* __restrict__ sums) {
    // quantities [qlo, qhi) are added up (the others count as 0), sums[slo, shi) are stored
    const int lo = blockIdx.x * per, hi = min(lo + per, nwarps);
    double tot[6] = {0, 0, 0, 0, 0, 0};
    // up to 4 independent rows of loads in flight per thread, accumulated in index order
    for (int w0 = lo + threadIdx.x; w0 < hi; w0 += 4 * RT) {
        double v[4][6];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int w = w0 + u * RT;
#pragma unroll
            for (int q = 0; q < 6; q++) v[u][q] = (w < hi && q >= qlo && q < qhi) ? __ldcg(partials + (size_t)q * nwarps + w) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int q = 0; q < 6; q++) tot[q] += v[u][q];
    }
    __shared__ double bs[6];
    __shared__ bool last;
    block_sum6(tot, bs);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 6; q++) block_sums[q * TM_REDUCE_BLOCKS + blockIdx.x] = bs[q];
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < 6 * 32) {
        __threadfence();
        // warp q adds the block sums of quantity q: lanes stride over the blocks, then a fixed tree
        const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
        double t = 0.0;
        for (unsigned b = lane; b < gridDim.x; b += 32) t += __ldcg(block_sums + q * TM_REDUCE_BLOCKS + b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0 && q >= slo && q < shi) sums[q] = t;
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// ---- K3 + K4: gather gradient ----------------------------------------------------
// chain-rule factors of the global penalties (kernel.h:27-71,124-137) and the constants of a sweep
struct GradK {
    double fA, fV, fC, kb, kt;
};

// ring state of one slot of the gradient walk: spoke e_s, |e_s|^2, rho of j_s and, with the
// previous spoke, p_s = e_s . e_{s-1}, n_s = e_s x e_{s-1}, 1/|n_s|
struct GSlot {
    d3 e, n;
    double l2, rho, p, in;
};

__device__ __forceinline__ void gslot_point(GSlot& t, d3 e, double rho) {
    t.e = e; t.rho = rho; t.l2 = dot3(e, e);
}
__device__ __forceinline__ void gslot_pair(GSlot& t, const GSlot& before) {
    t.p = dot3(t.e, before.e);
    t.n = cross3(t.e, before.e);
    t.in = rsqrt_pos(dot3(t.n, t.n));
}

// Boundary spoke (i -> j has a face, j -> i is a boundary half-edge). The reference's dihedral_angle is 0
// there, but its vertex_properties_grad still adds l * dihedral_angle_grad<1> (mesh_properties.cpp:88-98,
// mesh_util.h:135-185), evaluated on the boundary loop as OpenMesh defines sectors of
// boundary half-edges: n1 = the normal of the sector (j -> i, i -> p) -- which is the wrapped-around
// normal e_p x e_j of the ring walk -- and a04 = the signed angle at j between i and m, the boundary
// neighbour of j on the other side, the sign taken against the face (j, m, q). Returns cot(a04).
// Parity with the reference demands the same term; it only ever runs for boundary vertices of open meshes.
__device__ __noinline__ double boundary_spoke_cot(const MeshView& mv, d3 xi, d3 xj, int j) {
    const int b = mv.ring_off[j], d = mv.ring_off[j + 1] - b;
    int sm = -1;
    for (int t = 0; t < d; t++) {
        const int w = mv.ring_v[b + t];
        if ((w & TM_RING_FACE) && !(w & TM_RING_OFACE)) sm = t;       // j -> m has a face, m -> j is boundary
    }
    if (sm < 0 || d < 2) return 0.0;
    const int m = mv.ring_v[b + sm] & TM_RING_MASK;
    const int q = mv.ring_v[b + (sm == 0 ? d - 1 : sm - 1)] & TM_RING_MASK;   // face (j, m, q)
    const d3 xm = ldpos(mv.pos, m), xq = ldpos(mv.pos, q);
    const d3 v0 = xi - xj, v1 = xm - xj;
    const d3 cr = cross3(v0, v1);
    const d3 fn = cross3(xq - xm, xj - xm);
    const double sn = sqrt_nn(dot3(cr, cr));
    if (!(sn > 0.0)) return 0.0;
    const double c = dot3(v0, v1) / sn;                                // |cot|, signed by the cosine
    return dot3(cr, fn) >= 0.0 ? c : -c;
}

// The boundary-spoke terms of vertex i, evaluated apart from the ring walk (rare, and a call inside the
// walk costs the closed-mesh path ~8 % in spills): for every slot with a face whose opposite half-edge is
// on the boundary, W [cot(a03) n0^ + cot(a04) n1^] with W = fC/2 + kb (rho_i + rho_j).
__device__ __noinline__ void boundary_spoke_terms(const MeshView& mv, double fC, double kb, int i, d3 xi, double rho_i,
                                                  d3& g) {
    const int b = mv.ring_off[i], d = mv.ring_off[i + 1] - b;
    for (int s = 0; s < d; s++) {
        const int w = mv.ring_v[b + s];
        if (!(w & TM_RING_FACE) || (w & TM_RING_OFACE)) continue;
        const int j = w & TM_RING_MASK;
        const P4 rj = ldrec(mv.pos, j);
        const d3 xj = mk3(rj.x, rj.y, rj.z);
        const d3 e_s = xj - xi;
        const d3 e_p = ldpos(mv.pos, mv.ring_v[b + (s == 0 ? d - 1 : s - 1)] & TM_RING_MASK) - xi;
        const d3 e_n = ldpos(mv.pos, mv.ring_v[b + (s + 1 == d ? 0 : s + 1)] & TM_RING_MASK) - xi;
        const d3 n_cur = cross3(e_s, e_p), n_nxt = cross3(e_n, e_s);
        const double W = 0.5 * fC + kb * (rho_i + rj.w);
        const double nn = dot3(n_cur, n_cur), mm = dot3(n_nxt, n_nxt);
        if (nn > 0.0) axpy3(g, W * (dot3(e_s, e_s) - dot3(e_s, e_p)) / nn, n_cur);
        if (mm > 0.0) axpy3(g, W * boundary_spoke_cot(mv, xi, xj, j) / sqrt_nn(mm), n_nxt);
    }
}
// contribution of ring slot s (mesh_properties.cpp:73-134) to g_i
__device__ __forceinline__ void grad_slot(const DevParams& P, const GradK& K, double rho_i, int word,
                                          const GSlot& prv, const GSlot& cur, const GSlot& nxt,
                                          d3& g, d3& nsum, d3& esum) {
    if (!(word & TM_RING_FACE)) return;
    const double il = rsqrt_pos(cur.l2);       // 1/l_s
    const double l = cur.l2 * il;
    const double inn_cur = cur.in * cur.in;
    double cE = 0.0, cEp = 0.0, cN = 0.0, cNn = 0.0;

    // spoke edge (i, j_s): owners i and j_s   (mesh_properties.cpp:88-98)
    const double W = 0.5 * K.fC + K.kb * (rho_i + cur.rho);
    if (word & TM_RING_OFACE) {
        const double th = tm_math::atan2_fast(-l * dot3(nxt.e, cur.n), dot3(cur.n, nxt.n));
        cE -= W * th * il;                                   // theta * dl/dx_i
        cN += W * (cur.l2 - cur.p) * inn_cur;                // l * dtheta/dx_i
        cNn += W * (cur.l2 - nxt.p) * (nxt.in * nxt.in);
    }
    // face (i, j_s, j_{s-1}): owners i, j_s, j_{s-1}   (:103-118)
    double U = K.fA - (2.0 / 3.0) * K.kb * (rho_i * rho_i + cur.rho * cur.rho + prv.rho * prv.rho);
    if (P.bond_type == TM_BOND_AREA) {
        const double a_f = 0.5 * rcp_pos(cur.in);
        U += K.kt * 2.0 * (a_f / P.a0 - 1.0) / P.a0;
    }
    const double h = 0.5 * U * cur.in;                       // dA_f/dx_i in the e basis
    cEp += h * (cur.p - cur.l2);
    cE -= h * (prv.l2 - cur.p);
    // link edge (j_s, j_{s-1}) seen from its apex i: owners j_s, j_{s-1}   (:121-129)
    const double Wl = 0.5 * K.fC + K.kb * (cur.rho + prv.rho);
    cN -= Wl * (cur.l2 + prv.l2 - 2.0 * cur.p) * inn_cur;
    // tether bond (:132-133)
    if (P.bond_type == TM_BOND_EDGE) {
        if (l > P.lc0 || l < P.lc1) cE -= K.kt * tether_edge_fac(l, P) * il;
    }
    axpy3(g, cE, cur.e);
    axpy3(g, cEp, prv.e);
    axpy3(g, cN, cur.n);
    axpy3(g, cNn, nxt.n);
    nsum = nsum + cur.n;
    esum = esum + (prv.e - cur.e);
}

// Closed fan of valence d = 3..8: the fast path of the gradient sweep (mesh_properties.cpp:73-134).
//
// Collecting the reference's five blocks per basis vector of the fan (e_s spokes, n_s = e_s x e_{s-1}
// face normals; p_s = e_s . e_{s-1}, a_s = |e_s|^2 - p_s, b_s = |e_{s-1}|^2 - p_s):
//   g_i = sum_s E_s e_s + C_s n_s
//   C_s = kb/|n_s|^2 [(rho_i - rho_{s-1}) a_s + (rho_i - rho_s) b_s] + fV/6
//         (the spoke terms of both edges of face s and its link-edge term; the kappa_c factors cancel, the
//          volume gradient sum_f n_f / 6 rides along)
//   E_s = -W_s tau_s - h_s b_s - h_{s+1} a_{s+1} - kt tether'(l_s)/l_s,   W_s = fC/2 + kb (rho_i + rho_s),
//   h_s = U_s / (2 |n_s|),  U_s = fA - 2/3 kb (rho_i^2 + rho_s^2 + rho_{s-1}^2) [+ AREA bond]
// tau_s = theta_s / l_s and 1/|n_s| come from the property sweep (aux arrays), so a slot costs ~40 fp64
// operations and no square root, reciprocal or arc tangent. Ring records (x, y, z, rho) staged
// global -> shared with cp.async; slot loop unrolled over the 8 slots of the row.
struct GradC {
    double W0, U0h, c3, fV6, kb, kt;
};

struct GradWalk {
    const DevParams& P;
    const GradC& K;
    const P4* pos;
    const R8& row;
    const double* tau;          // aux + i: tau of slot s at tau[s * vs]
    int vs;
    d3 xi;
    double rho_i;
    int d;
    d3& g;
    unsigned& tmask;

    struct Pt { d3 e; double l2, dr, r2; };        // spoke, |e|^2, rho_i - rho, rho^2
    __device__ __forceinline__ Pt point(int j) const {
        const P4 r = ldrec(pos, j);
        Pt t;
        t.e = mk3(r.x, r.y, r.z) - xi;
        t.l2 = dot3(t.e, t.e);
        t.dr = rho_i - r.w;
        t.r2 = r.w * r.w;
        return t;
    }
    // face between spokes pp (s-1) and pc (s): adds C n_s, the face's share of both spokes and spoke s's own term;
    // `valid` masks the surplus slots of a lane of lower valence
    template <bool MASK>
    __device__ __forceinline__ void face(const Pt& pp, const Pt& pc, double ta, bool valid, int s) const {
        if (!MASK) valid = true;
        const double p = dot3(pc.e, pp.e);
        const d3 n = cross3(pc.e, pp.e);
        const double inv = rsqrt_pos(dot3(n, n) + 1.0e-300);          // 1/|n_s|
        const double a = pc.l2 - p, b = pp.l2 - p;
        const double C = fma(K.kb * inv * inv, fma(pp.dr, a, pc.dr * b), K.fV6);
        double Uh = fma(-K.c3, pc.r2 + pp.r2, K.U0h);
        if (P.bond_type == TM_BOND_AREA) {
            const double a_f = 0.5 * rcp_pos(inv);
            Uh = fma(K.kt * P.inv_a0, fma(a_f, P.inv_a0, -1.0), Uh);
        }
        const double h = Uh * inv;
        // W_s = fC/2 + kb (rho_i + rho_s) = W0 - kb (rho_i - rho_s) with W0 = fC/2 + 2 kb rho_i
        const double E_c = fma(-h, b, -fma(-K.kb, pc.dr, K.W0) * ta);
        if (P.bond_type == TM_BOND_EDGE) tmask |= (valid && (pc.l2 > P.lc0sq || pc.l2 < P.lc1sq)) ? (1u << s) : 0u;
        axpy3(g, valid ? C : 0.0, n);
        axpy3(g, valid ? E_c : 0.0, pc.e);
        axpy3(g, valid ? -h * a : 0.0, pp.e);
    }
    // the whole ring in one basic block (see PropsWalk::block): face 0 lies between the last spoke and spoke 0.
    // The slots' tau (coalesced, straight from DRAM: nobody has touched them since the property sweep wrote them)
    // are all requested before the first face, so the walk pays that latency once and not once per group.
    template <int DMAX, bool MASK>
    __device__ __forceinline__ void block() const {
#if TM_GRAD_TAU_EARLY
        double ta[DMAX];
#pragma unroll
        for (int s = 0; s < DMAX; s++) ta[s] = (!MASK || s < 3 || s < d) ? __ldg(tau + s * vs) : 0.0;
#endif
        int j_l = row.v[DMAX - 1];
        if (MASK) {
            j_l = row.v[2];
#pragma unroll
            for (int s = 3; s < DMAX; s++) j_l = s < d ? row.v[s] : j_l;
        }
        Pt pp = point(j_l);
#if TM_GRAD_TAU_EARLY
        asm volatile("" ::: "memory");
#endif
#pragma unroll
        for (int s = 0; s < DMAX; s++) {
            if (s > 0 && s % TM_GRAD_GROUP == 0) asm volatile("" ::: "memory");
            const bool valid = !MASK || s < 3 || s < d;
            const Pt pc = point(valid ? row.v[s] : row.v[0]);
#if TM_GRAD_TAU_EARLY
            const double tas = ta[s];
#else
            const double tas = valid ? __ldg(tau + s * vs) : 0.0;
#endif
            face<MASK>(pp, pc, tas, valid, s);
            pp = pc;
        }
    }
};

__device__ __forceinline__ d3 ring_grad_closed(const MeshView& mv, const DevParams& P, const GradC& K, const R8& row,
                                               int d, int dmax, bool uniform, int i, d3 xi, double rho_i) {
    d3 g = mk3(0, 0, 0);
    unsigned tmask = 0;
    const GradWalk W{P, K, mv.pos, row, mv.aux + i, mv.aux_vs, xi, rho_i, d, g, tmask};
#define TM_WALK(D) do { if (uniform) W.template block<D, false>(); else W.template block<D, true>(); } while (0)
    switch (dmax) {                            // warp-uniform
        case 3: case 4: W.template block<4, true>(); break;
        case 5: TM_WALK(5); break;
        case 6: TM_WALK(6); break;
        case 7: TM_WALK(7); break;
        default: TM_WALK(8); break;
    }
#undef TM_WALK
    // EDGE bonds off the flat bottom (mesh_properties.cpp:132-133): rare, evaluated outside the walk
    while (tmask) {
        const int s = __ffs(tmask) - 1;
        tmask &= tmask - 1;
        const d3 e = ldpos(mv.pos, row_slot(row, s)) - xi;
        const double l2 = dot3(e, e);
        const double il = rsqrt_pos(l2);
        axpy3(g, -K.kt * tether_edge_fac(l2 * il, P) * il, e);
    }
    return g;
}

__device__ __forceinline__ void ring_grad_generic(const MeshView& mv, const DevParams& P, const GradK& K, int i,
                                                  d3 xi, double rho_i, d3& g, d3& nsum, d3& esum) {
    const int beg = mv.ring_off[i], d = mv.ring_off[i + 1] - beg;
    if (d <= 0) return;
    const int* ring = mv.ring_v + beg;
    auto point = [&](GSlot& t, int word) {
        const P4 q = ldrec(mv.pos, word & TM_RING_MASK);
        gslot_point(t, mk3(q.x, q.y, q.z) - xi, q.w);
    };
    GSlot prv, cur;
    int w_cur = ring[0];
    point(prv, ring[d - 1]);
    point(cur, w_cur);
    gslot_pair(cur, prv);
    for (int s = 0; s < d; s++) {
        const int w_next = ring[s + 1 == d ? 0 : s + 1];
        GSlot nxt;
        point(nxt, w_next);
        gslot_pair(nxt, cur);
        grad_slot(P, K, rho_i, w_cur, prv, cur, nxt, g, nsum, esum);
        prv = cur; cur = nxt; w_cur = w_next;
    }
}

// what follows the ring walk of vertex i: repulsion, the gradient row, and the velocity-Verlet update fused
// into the sweep (tm_integrate.cu)
template <int UPD>
// fpre: the vertex's pair force, already in registers (the closed-fan sweep requests it with the ring records)
__device__ __forceinline__ void grad_finish(const DevParams& P, int i, const P4& ri, d3 g, const double* __restrict__ force,
                                            double* __restrict__ grad, const VVFuse& U, const d3* fpre = nullptr,
                                            const double* upre = nullptr /* momenta [0..2], previous gradient [3..5] */) {
    // repulsion (mesh_properties.cpp:139-146): 2 kappa_r * pair force of k_pairs
    if (force) {
        if (fpre) {
            axpy3(g, 2.0 * P.kappa_r, *fpre);
        } else {
            const double* f = force + 3 * (size_t)i;
            axpy3(g, 2.0 * P.kappa_r, mk3(__ldg(f), __ldg(f + 1), __ldg(f + 2)));
        }
    }

    double* o = grad + 3 * (size_t)i;
    o[0] = g.x; o[1] = g.y; o[2] = g.z;

    if (UPD > 0) {
        // velocity Verlet in the epilogue (same explicitly rounded operations as k_vv_update, i.e. numpy's
        // order): p += 0.5 (a + an) dt ; x_next = x + (p dt + 0.5 an dt^2) / m, with a = -g_prev, an = -g
        const double an[3] = {-g.x, -g.y, -g.z};
        const double xs[3] = {ri.x, ri.y, ri.z};
        double xn[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double a = -(upre ? upre[3 + k] : U.g_prev[3 * (size_t)i + k]);
            const double pk = __dadd_rn(upre ? upre[k] : U.mom[3 * (size_t)i + k], __dmul_rn(__dmul_rn(0.5, __dadd_rn(a, an[k])), U.dt));
            U.mom[3 * (size_t)i + k] = pk;
            if (UPD > 1) {
                const double num = __dadd_rn(__dmul_rn(pk, U.dt), __dmul_rn(__dmul_rn(0.5, an[k]), U.dt2));
                xn[k] = __dadd_rn(xs[k], __ddiv_rn(num, U.mass));
            }
        }
        if (UPD > 1) {
            P4* q = U.pos_next + i;
            *reinterpret_cast<double2*>(q) = make_double2(xn[0], xn[1]);
            *(reinterpret_cast<double2*>(q) + 1) = make_double2(xn[2], ri.w);
        }
    }
}

// chain-rule factors of the global penalties (kernel.h:27-71,124-137)
__device__ __forceinline__ void penalty_factors(const DevParams& P, const double* __restrict__ sums, double& fA, double& fV,
                                                double& fC) {
    fA = 2.0 * P.kappa_a * P.inv_ref_area * fma(sums[0], P.inv_ref_area, -1.0);
    fV = 2.0 * P.kappa_v * P.inv_ref_volume * fma(sums[1], P.inv_ref_volume, -1.0);
    fC = 2.0 * P.kappa_c * P.inv_ref_curvature * fma(sums[2], P.inv_ref_curvature, -1.0);
}

// Closed fans only (vertices without a fixed-width row are left to k_grad_generic). Reads only: every ring
// neighbour costs one (x, y, z, rho) record, every slot one tau of the property sweep.
template <int UPD>
__global__ void __launch_bounds__(TM_GRAD_THREADS, TM_GRAD_MINB)
k_grad(MeshView mv, DevParams P, const double* __restrict__ sums, const double* __restrict__ force,
       double* __restrict__ grad, SweepRange R, VVFuse U, int blk0) {
    const int t = (blockIdx.x + blk0) * TM_GRAD_THREADS + threadIdx.x;
    const bool live = t < R.count;
    const int i = R.first + t;
    R8 row;
    row.v[0] = TM_RING_GENERIC;
    P4 ri;
    ri.x = ri.y = ri.z = ri.w = 0.0;
    d3 fr = mk3(0, 0, 0);
    double up[6] = {0, 0, 0, 0, 0, 0};        // the fused update's momenta and previous gradient, requested with the row too
    if (live) {
        row = ldrow(mv.ring8, i);
        ri = ldrec(mv.pos, i);
        // the epilogue's operands. The pair force is loaded now (three coalesced doubles: waiting for it behind the
        // walk was a tenth of a warp's life, profiles/r02x), and so are the fused update's.
        if (force) { const double* f = force + 3 * (size_t)i; fr = mk3(__ldg(f), __ldg(f + 1), __ldg(f + 2)); }
        if (UPD > 0) {
#pragma unroll
            for (int k = 0; k < 3; k++) { up[k] = U.mom[3 * (size_t)i + k]; up[3 + k] = __ldg(U.g_prev + 3 * (size_t)i + k); }
        }
    }
    // (the three sums are requested with the row, not behind it)
    double fA, fV, fC;
    penalty_factors(P, sums, fA, fV, fC);
    const bool closed = row.v[0] != TM_RING_GENERIC;
    const int d = closed ? row_valence(row) : 0;
    const int dmax = __reduce_max_sync(0xffffffffu, d);
    const bool uniform = __all_sync(0xffffffffu, !closed || d == dmax);
    if (!closed) return;
    GradC K;
    K.kb = P.kappa_b; K.kt = P.kappa_t;
    K.W0 = fma(2.0 * K.kb, ri.w, 0.5 * fC);       // W_s = W0 - kb (rho_i - rho_s)
    K.c3 = K.kb * (1.0 / 3.0);
    K.U0h = fma(-K.c3, ri.w * ri.w, 0.5 * fA);
    K.fV6 = fV * (1.0 / 6.0);
    const d3 g = ring_grad_closed(mv, P, K, row, d, dmax, uniform, i, mk3(ri.x, ri.y, ri.z), ri.w);
    grad_finish<UPD>(P, i, ri, g, force, grad, U, &fr, UPD > 0 ? up : nullptr);
}

// The vertices of `list` inside [R.first, R.first + R.count): valence > 8 or < 3, and the boundary vertices
// of open meshes, including the reference's boundary-spoke terms (boundary_spoke_terms above).
template <int UPD>
__global__ void __launch_bounds__(128)
k_grad_generic(MeshView mv, DevParams P, const double* __restrict__ sums, const double* __restrict__ force,
               double* __restrict__ grad, const int* __restrict__ list, int n, SweepRange R, VVFuse U) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (t >= n) return;
    const int i = list[t];
    if (i < R.first || i >= R.first + R.count) return;
    const P4 ri = ldrec(mv.pos, i);
    const d3 xi = mk3(ri.x, ri.y, ri.z);
    GradK K;
    penalty_factors(P, sums, K.fA, K.fV, K.fC);
    K.kb = P.kappa_b; K.kt = P.kappa_t;
    d3 g = mk3(0, 0, 0), nsum = mk3(0, 0, 0), esum = mk3(0, 0, 0);
    ring_grad_generic(mv, P, K, i, xi, ri.w, g, nsum, esum);
    // volume: sum_f (x_j x x_k)/6 = (sum_f n_f + x_i x sum_f (e_{s-1} - e_s))/6
    axpy3(g, K.fV / 6.0, nsum + cross3(xi, esum));
    boundary_spoke_terms(mv, K.fC, K.kb, i, xi, ri.w, g);
    grad_finish<UPD>(P, i, ri, g, force, grad, U);
}

}  // namespace

// does the pair force stored in m->frep belong to the current positions, list and potential?
static bool pair_force_valid(const tm_mesh* m, const NlView& nl, const tm_energy_params& p) {
    return m->frep_epoch == m->pos_epoch && m->frep_lc1 == p.rep_lc1 && m->frep_r == p.rep_r &&
           m->frep_list == nl.serial;
}

bool tm_pair_force_valid(const tm_mesh* m, const NlView& nl, const tm_energy_params& p) {
    return nl.kind >= 0 && pair_force_valid(m, nl, p);
}

// cell lists: the candidates' records (position, id) at the current positions, in cell order
static int ensure_cellrec(tm_mesh* m, const NlView& nl, cudaStream_t st) {
    if (nl.kind == TM_NLIST_CELL && !(m->cellrec_epoch == m->pos_epoch && m->cellrec_list == nl.serial)) {
        k_cell_gather<<<(m->V + 255) / 256, 256, 0, st>>>(m->V, nl.cell_items, m->pos, const_cast<P4*>(nl.cell_rec));
        TM_LAUNCHED();
        TM_CUDA(cudaGetLastError());
        m->cellrec_epoch = m->pos_epoch; m->cellrec_list = nl.serial;
    }
    return 0;
}
int tm_ensure_cellrec(tm_mesh* m, const NlView& nl) { return tm_join_pairs(m) || ensure_cellrec(m, nl, m->stream); }

// pair sweep: energy partials (quantity 5) and/or per-vertex force
// beside = on the mesh's side streams, concurrently with whatever the caller launches next on the
// main stream (the ring property sweep); tm_join_pairs before the first consumer.
//
// A sweep beside the main stream over a verlet list is cut into K ranges of tiles, one launch and one side stream
// each, and adds up its own repulsion sum (own_sum). The ring gradient of a range then follows that range's pair
// sweep on the range's stream (launch_grad_kernels), so the gradient needs neither the whole pair sweep nor a
// device-side protocol: the kernel boundary is the fence, and while one range's sweep drains, the blocks of the
// other streams fill the multiprocessors. Measured at 655,362 vertices (profiles/r02ao): E+G 138.0 -> 122.9 us,
// a gradient at new positions 127.0 -> 118.1 us with K = 2 (K = 4: 124.3 / 118.7, K = 8: 123.9 / 120.3).
// TM_PAIR_SPLIT=K overrides (0 or 1: one sweep on one side stream, the reduction waits for it).
static int pair_split(const tm_mesh* m, const NlView& nl, int nb) {
    static const int K = [] {
        const char* e = getenv("TM_PAIR_SPLIT");
        const int k = e ? atoi(e) : 2;
        return k < 2 ? 0 : (k > TM_PAIR_STREAMS ? TM_PAIR_STREAMS : k);
    }();
    // (small meshes: a kernel of a few microseconds gains nothing from being cut in two.) The ranks of a partitioned
    // mesh must decide alike -- a split sweep exchanges its sums in two steps -- so they go by the mesh, not by what
    // each of them owns.
    const long long tiles = m->part ? (long long)((m->V + PAIR_THREADS - 1) / PAIR_THREADS) / tm_part_nranks(m) : nb;
    return (nl.kind == TM_NLIST_VERLET && tiles >= 1024 && PAIR_THREADS == TM_GRAD_THREADS) ? K : 0;
}

bool tm_pair_split_active(const tm_mesh* m, const NlView& nl) {
    return nl.kind >= 0 && pair_split(m, nl, (tm_owned_range(m).count + PAIR_THREADS - 1) / PAIR_THREADS) > 0;
}

static int launch_pairs(tm_mesh* m, const NlView& nl, const tm_energy_params& p, bool with_e, bool write_vprops,
                        int nwarps, bool beside = false) {
    if (!m->frep) TM_CUDA(cudaMalloc(&m->frep, sizeof(double) * 3 * (size_t)m->V));
    // cell lists test exclusion level 2 against the sorted two-ring rows (rebuilt when the connectivity changed)
    if (nl.kind == TM_NLIST_CELL && nl.excl == 2 && tm_ensure_ring2(m)) return 1;
    const SweepRange R = tm_owned_range(m);
    const int nb = (R.count + PAIR_THREADS - 1) / PAIR_THREADS;
    MeshView mv = m->view();
    cudaStream_t st = m->stream;
    const int K = (beside && !write_vprops) ? pair_split(m, nl, nb) : 0;
    if (beside) {
        if (tm_join_pairs(m)) return 1;
        TM_CUDA(cudaEventRecord(m->ev_fork, m->stream));
        if (K <= 0) TM_CUDA(cudaStreamWaitEvent(m->stream2, m->ev_fork, 0));
        st = m->stream2;
    }
    const size_t qbytes = nl.kind == TM_NLIST_CELL ? sizeof(int) * (PAIR_QCAP + PAIR_XCAP) * PAIR_THREADS : 0;
    if (ensure_cellrec(m, nl, st)) return 1;
    int gnb = nb, blk0 = 0;
#define TM_PAIRS2(E, F, W, R2, CL) k_pairs<E, F, W, R2, CL><<<gnb, PAIR_THREADS, qbytes, st>>>( \
        mv, nl, p.rep_lc1, p.rep_r, m->frep, m->vprops, m->partials, nwarps, R, blk0)
#define TM_PAIRS1(E, F, W, R2) do { if (nl.kind == TM_NLIST_CELL) TM_PAIRS2(E, F, W, R2, true); else TM_PAIRS2(E, F, W, R2, false); } while (0)
#define TM_PAIRS(E, F, W, R2) do { if (nb <= 0) break; if (K <= 0) { TM_PAIRS1(E, F, W, R2); break; } \
        const int per = (nb + K - 1) / K; \
        for (int k = 0; k < K; k++) { \
            blk0 = k * per; gnb = nb - blk0 < per ? nb - blk0 : per; \
            if (gnb <= 0) break; \
            st = m->sp_stream[k]; \
            TM_CUDA(cudaStreamWaitEvent(st, m->ev_fork, 0)); \
            TM_PAIRS1(E, F, W, R2); \
            TM_CUDA(cudaEventRecord(m->sp_ev[k], st)); \
            TM_LAUNCHED(); \
        } } while (0)
    const bool r2 = p.rep_r == 2;
    if (write_vprops) { if (r2) TM_PAIRS(true, true, true, true); else TM_PAIRS(true, true, true, false); }
    else if (with_e)  { if (r2) TM_PAIRS(true, true, false, true); else TM_PAIRS(true, true, false, false); }
    else              { if (r2) TM_PAIRS(false, true, false, true); else TM_PAIRS(false, true, false, false); }
#undef TM_PAIRS
#undef TM_PAIRS1
#undef TM_PAIRS2
    if (K <= 0) TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    if (beside) {
        if (K <= 0) TM_CUDA(cudaEventRecord(m->ev_join, m->stream2));
        m->pairs_inflight = true;
        m->split = K;
    }
    m->frep_epoch = m->pos_epoch; m->frep_lc1 = p.rep_lc1; m->frep_r = p.rep_r;
    m->frep_list = nl.serial;
    return 0;
}

// fixed-order sum of the warp partials of quantities [qlo, qhi) -> sums[slo, shi) on the main stream
static int launch_reduce(tm_mesh* m, int ncols, int qlo, int qhi, int slo, int shi) {
    int rb = (ncols + 4 * RT - 1) / (4 * RT);
    rb = rb < 1 ? 1 : (rb > TM_REDUCE_BLOCKS ? TM_REDUCE_BLOCKS : rb);
    const int per = (ncols + rb - 1) / rb;
    k_reduce_props<<<rb, RT, 0, m->stream>>>(m->partials, ncols, per, qlo, qhi, slo, shi, m->partials + 6 * (size_t)ncols,
                                            m->ticket, tm_part_local_sums(m));
    TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    return 0;
}

// The main stream waits for what the side streams carry. A split pair sweep that evaluated the energy left the
// repulsion column of the warp partials to be added up: that reduction goes behind the wait (rep_cols > 0).
static int join_side_streams(tm_mesh* m) {
    if (m->split > 0) { for (int k = 0; k < m->split; k++) TM_CUDA(cudaStreamWaitEvent(m->stream, m->sp_ev[k], 0)); }
    else TM_CUDA(cudaStreamWaitEvent(m->stream, m->ev_join, 0));
    if (m->rep_cols > 0) {
        const int ncols = m->rep_cols;
        m->rep_cols = 0;
        if (launch_reduce(m, ncols, 5, 6, 5, 6) || tm_part_exchange_sums(m, 5, 6)) return 1;
    }
    return 0;
}

int tm_join_pairs(tm_mesh* m) {
    if (!m->pairs_inflight) return 0;
    if (join_side_streams(m)) return 1;
    m->split = 0;
    m->pairs_inflight = false;
    return 0;
}

// warp partials of the largest sweep (whole mesh + its generic vertices) and the slot quantities the property
// sweep hands to the gradient sweep; sized for the whole mesh so that a partitioned mesh never reallocates
// (cudaFree synchronises the device) while a peer waits on the device
int tm_props_workspace(tm_mesh* m) {
    const size_t need = 12 * (size_t)((m->V + 31) / 32) + 6 * TM_REDUCE_BLOCKS;     // (n_generic <= V)
    if (m->partials_n < need) {
        if (m->partials) cudaFree(m->partials);
        m->partials = nullptr; m->partials_n = 0;
        TM_CUDA(cudaMalloc(&m->partials, need * sizeof(double)));
        m->partials_n = need;
    }
    if (!m->aux) TM_CUDA(cudaMalloc(&m->aux, sizeof(double) * 8 * (size_t)((m->V + 31) & ~31)));
    return 0;
}

int tm_launch_props(tm_mesh* m, const NlView& nl, const tm_energy_params& p, bool full, bool write_vprops) {
    if (tm_ensure_layout(m) || tm_part_refresh(m, nl)) return 1;
    const SweepRange R = tm_owned_range(m);
    const int nb = (R.count + TM_PROPS_THREADS - 1) / TM_PROPS_THREADS;
    const int nwarps = (R.count + 31) / 32;
    // warp partials: the closed-fan sweep's columns, then those of the generic-vertex sweep
    const int ngen = m->n_generic, gen_warps = (ngen + 31) / 32;
    const int ncols = nwarps + gen_warps;
    if (tm_props_workspace(m)) return 1;
    DevParams P = tm_to_dev(p);
    MeshView mv = m->view();
    // the repulsion sum is wanted whenever lc1 > 0, whatever kappa_r is (mesh_properties.cpp:46)
    const bool pairs = full && nl.kind >= 0 && p.rep_lc1 > 0.0;
    // quantities the ring sweep reduces: geometry + bending (4) for the gradient's sweep, all but
    // the repulsion when a pair sweep follows, else all six (repulsion = 0)
    const int nq = !full ? 4 : (pairs ? 5 : 6);
    // The pair sweep runs beside the ring sweep: with the energy (full), or force-only for the gradient
    // sweep that follows a geometry-only property sweep (joined in tm_launch_grad).
    const bool pairs_for_grad = !full && nl.kind >= 0 && p.rep_lc1 > 0.0 && p.kappa_r != 0.0 && !pair_force_valid(m, nl, p);
    if (write_vprops) {
        if (!m->vprops) TM_CUDA(cudaMalloc(&m->vprops, sizeof(double) * 6 * (size_t)m->V));
        TM_CUDA(cudaMemsetAsync(m->vprops, 0, sizeof(double) * 6 * (size_t)m->V, m->stream));
    }
    if (pairs && launch_pairs(m, nl, p, true, write_vprops, ncols, true)) return 1;
    // Nothing on the main stream waits for a split pair sweep here: the reduction below leaves its column (the
    // repulsion sum) out, and the first consumer of sums[5] or of the pair force joins and adds it up (tm_launch_grad,
    // fetch_sums, the trajectory's statistics). Otherwise the reduction adds the pair sweep's warp partials too.
    const bool own = pairs && m->split > 0;
    if (pairs_for_grad && launch_pairs(m, nl, p, false, false, 0, true)) return 1;
    if (nb <= 0) {
        // a rank of a tiny partitioned mesh may own nothing: it still takes part in the reduction / exchange
    } else if (write_vprops) {
        k_props<true, true, false><<<nb, TM_PROPS_THREADS, 0, m->stream>>>(mv, P, m->pos, m->vprops, m->partials, ncols, nq, R);
    } else if (full) {
        k_props<true, false, false><<<nb, TM_PROPS_THREADS, 0, m->stream>>>(mv, P, m->pos, nullptr, m->partials, ncols, nq, R);
    } else {
        k_props<false, false, false><<<nb, TM_PROPS_THREADS, 0, m->stream>>>(mv, P, m->pos, nullptr, m->partials, ncols, nq, R);
    }
    TM_LAUNCHED();
    if (ngen > 0) {
        // vertices without a fixed-width row; their columns carry all six quantities (repulsion = 0)
        const int gb = (ngen + 127) / 128;
        if (write_vprops) k_props_generic<true, true><<<gb, 128, 0, m->stream>>>(mv, P, m->pos, m->vprops, m->partials, ncols, nwarps, 6, m->gen_list, ngen, R);
        else if (full) k_props_generic<true, false><<<gb, 128, 0, m->stream>>>(mv, P, m->pos, nullptr, m->partials, ncols, nwarps, 6, m->gen_list, ngen, R);
        else k_props_generic<false, false><<<gb, 128, 0, m->stream>>>(mv, P, m->pos, nullptr, m->partials, ncols, nwarps, 6, m->gen_list, ngen, R);
        TM_LAUNCHED();
    }
    // partitioned mesh: rho of the non-owned ring neighbours of owned vertices (the gradient
    // sweep reads it) from a second, list-driven launch that reduces nothing
    const SweepRange G = tm_ghost_range(m);
    if (G.count > 0) {
        k_props<false, false, true><<<(G.count + TM_PROPS_THREADS - 1) / TM_PROPS_THREADS, TM_PROPS_THREADS, 0, m->stream>>>(
            mv, P, m->pos, nullptr, nullptr, 0, 0, G);
        TM_LAUNCHED();
    }
    if (pairs && !own && tm_join_pairs(m)) return 1;
    // own: sums[0..4] now, sums[5] from the pair sweep's column behind the join (join_side_streams)
    if (launch_reduce(m, ncols, 0, full ? (own ? 5 : 6) : 4, 0, own ? 5 : 6)) return 1;
    if (own) m->rep_cols = ncols;
    if (tm_part_exchange_sums(m, 0, full ? (own ? 5 : 6) : 4)) return 1;
    m->geo_epoch = m->pos_epoch;        // sums[0..2], rho and the slot quantities now belong to the current positions
    m->sums_full = full;                // a geometry-only sweep leaves sums[4..5] = 0
    return 0;
}

int tm_launch_pairs(tm_mesh* m, const NlView& nl, const tm_energy_params& p) {
    if (tm_ensure_layout(m)) return 1;
    if (!(nl.kind >= 0 && p.rep_lc1 > 0.0)) return 0;
    return launch_pairs(m, nl, p, false, false, 0);
}

// the ring gradient of the vertices of R on stream st: the closed-fan sweep, then the generic vertices if any
static int launch_grad_kernels(tm_mesh* m, const DevParams& P, const double* frc, const SweepRange& R, const VVFuse& U,
                               cudaStream_t st) {
    const int nb = (R.count + TM_GRAD_THREADS - 1) / TM_GRAD_THREADS;
    const MeshView mv = m->view();
    if (m->pairs_inflight && m->split > 0 && st == m->stream) {
        // every range's ring gradient behind that range's pair sweep, on the range's stream
        const int K = m->split, per = (nb + K - 1) / K;
        TM_CUDA(cudaEventRecord(m->ev_sums, m->stream));
        // (the main stream meanwhile waits for the pair sweeps -- the events as recorded so far -- and adds up the
        // repulsion column: a kernel of a few microseconds beside the ring gradients, which do not wait for it)
        if (join_side_streams(m)) return 1;
        for (int k = 0; k < K; k++) {
            const int blk0 = k * per, gnb = nb - blk0 < per ? nb - blk0 : per;
            if (gnb <= 0) break;
            cudaStream_t sk = m->sp_stream[k];
            TM_CUDA(cudaStreamWaitEvent(sk, m->ev_sums, 0));
            if (U.mode == 2) k_grad<2><<<gnb, TM_GRAD_THREADS, 0, sk>>>(mv, P, m->sums, frc, m->grad, R, U, blk0);
            else if (U.mode == 1) k_grad<1><<<gnb, TM_GRAD_THREADS, 0, sk>>>(mv, P, m->sums, frc, m->grad, R, U, blk0);
            else k_grad<0><<<gnb, TM_GRAD_THREADS, 0, sk>>>(mv, P, m->sums, frc, m->grad, R, U, blk0);
            TM_CUDA(cudaEventRecord(m->sp_ev[k], sk));
            if (k > 0) TM_LAUNCHED();
        }
        if (tm_join_pairs(m)) return 1;     // what follows on the main stream sees the finished gradient
    }
    else if (U.mode == 2) k_grad<2><<<nb, TM_GRAD_THREADS, 0, st>>>(mv, P, m->sums, frc, m->grad, R, U, 0);
    else if (U.mode == 1) k_grad<1><<<nb, TM_GRAD_THREADS, 0, st>>>(mv, P, m->sums, frc, m->grad, R, U, 0);
    else k_grad<0><<<nb, TM_GRAD_THREADS, 0, st>>>(mv, P, m->sums, frc, m->grad, R, U, 0);
    TM_LAUNCHED();
    if (m->n_generic > 0) {
        const int gb = (m->n_generic + 127) / 128;
        if (U.mode == 2) k_grad_generic<2><<<gb, 128, 0, st>>>(mv, P, m->sums, frc, m->grad, m->gen_list, m->n_generic, R, U);
        else if (U.mode == 1) k_grad_generic<1><<<gb, 128, 0, st>>>(mv, P, m->sums, frc, m->grad, m->gen_list, m->n_generic, R, U);
        else k_grad_generic<0><<<gb, 128, 0, st>>>(mv, P, m->sums, frc, m->grad, m->gen_list, m->n_generic, R, U);
        TM_LAUNCHED();
    }
    TM_CUDA(cudaGetLastError());
    return 0;
}

int tm_launch_grad(tm_mesh* m, const NlView& nl, const tm_energy_params& p, const VVFuse* fuse) {
    if (tm_ensure_layout(m) || tm_part_refresh(m, nl)) return 1;
    const bool pairs = nl.kind >= 0 && p.rep_lc1 > 0.0 && p.kappa_r != 0.0;
    if (pairs && !pair_force_valid(m, nl, p) && (tm_join_pairs(m) || launch_pairs(m, nl, p, false, false, 0))) return 1;
    // a split pair sweep still in flight: every range's ring gradient follows on that range's stream, joined behind it
    if (!(pairs && m->split > 0) && tm_join_pairs(m)) return 1;
    const SweepRange R = tm_owned_range(m);
    VVFuse U;
    memset(&U, 0, sizeof(U));
    if (fuse) U = *fuse;
    if (R.count <= 0) { m->grad_epoch = m->pos_epoch; return 0; }       // a rank of a tiny partitioned mesh may own nothing
    if (launch_grad_kernels(m, tm_to_dev(p), pairs ? m->frep : nullptr, R, U, m->stream)) return 1;
    m->grad_epoch = m->pos_epoch;
    return 0;
}

__global__ void k_export_grad(int V, const double* __restrict__ in, const int* __restrict__ d2h, double* __restrict__ out) {
    const int d = blockIdx.x * 256 + threadIdx.x;
    if (d >= V) return;
    const double* s = in + 3 * (size_t)d;
    double* o = out + 3 * (size_t)d2h[d];
    o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
}

int tm_launch_grad_speculative(tm_mesh* m, const NlView& nl, const tm_energy_params& p, cudaStream_t st, double* grad_host) {
    const bool pairs = nl.kind >= 0 && p.rep_lc1 > 0.0 && p.kappa_r != 0.0;
    SweepRange R;
    R.first = 0; R.count = m->V; R.vlist = nullptr;
    VVFuse U;
    memset(&U, 0, sizeof(U));
    if (launch_grad_kernels(m, tm_to_dev(p), pairs ? m->frep : nullptr, R, U, st)) return 1;
    k_export_grad<<<(m->V + 255) / 256, 256, 0, st>>>(m->V, m->grad, m->d2h, m->xout3); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    TM_CUDA(cudaMemcpyAsync(grad_host, m->xout3, sizeof(double) * 3 * (size_t)m->V, cudaMemcpyDeviceToHost, st));
    return 0;
}

int tm_preload_energy() {
    TM_PRELOAD((k_props<true, true, false>)); TM_PRELOAD((k_props<true, false, false>)); TM_PRELOAD((k_props<false, false, false>));
    TM_PRELOAD((k_props<false, false, true>));
    TM_PRELOAD((k_props_generic<true, true>)); TM_PRELOAD((k_props_generic<true, false>)); TM_PRELOAD((k_props_generic<false, false>));
    TM_PRELOAD(k_grad_generic<0>); TM_PRELOAD(k_grad_generic<1>); TM_PRELOAD(k_grad_generic<2>);
#define TM_PL(CL) \
    TM_PRELOAD((k_pairs<true, true, true, true, CL>)); TM_PRELOAD((k_pairs<true, true, true, false, CL>)); \
    TM_PRELOAD((k_pairs<true, true, false, true, CL>)); TM_PRELOAD((k_pairs<true, true, false, false, CL>)); \
    TM_PRELOAD((k_pairs<false, true, false, true, CL>)); TM_PRELOAD((k_pairs<false, true, false, false, CL>))
    TM_PL(false); TM_PL(true);
#undef TM_PL
    TM_PRELOAD(k_cell_gather); TM_PRELOAD(k_export_grad);
    TM_PRELOAD(k_reduce_props); TM_PRELOAD(k_grad<0>); TM_PRELOAD(k_grad<1>); TM_PRELOAD(k_grad<2>);
    return 0;
}
