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
// one-ring j_0..j_{d-1} (clockwise CSR row). With e_s = x_{j_s} - x_i,
//   n_s   = e_s x e_{s-1}            un-normalised normal of face (i, j_s, j_{s-1})
//   p_s   = e_s . e_{s-1}
//   theta_s = atan2(-|e_s| (e_{s+1} . n_s), n_s . n_{s+1})   dihedral of spoke s
// everything the reference obtains from acos/tan of sector angles follows
// algebraically (cot = dot/|cross|), so a sweep needs one atan2, two rsqrt per
// ring slot and no other transcendental. The gradient is a gather: vertex i
// only writes g_i (no fp64 atomics) and needs the ring positions plus
// rho_v = c_v / a_v of itself and of its ring from the property sweep.
// Global sums use warp shuffles, a fixed block tree and a fixed-order last-block
// pass, so results are bit-reproducible run to run.
#include "tm_common.cuh"
#include "tm_neigh.cuh"
#include "tm_vertex.cuh"

namespace {

// ---- reductions ---------------------------------------------------------------
template <int NT, int NQ>
__device__ __forceinline__ void block_reduce(double (&v)[NQ], double* out /* NQ, thread 0 */) {
    __shared__ double sm[NQ][NT / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        double t = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) sm[q][w] = t;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            double t = lane < NT / 32 ? sm[q][lane] : 0.0;
#pragma unroll
            for (int o = NT / 64; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) out[q] = t;
        }
    }
    __syncthreads();
}

// ---- K1 + K2: per-vertex properties and their global sums ----------------------
// FULL = also tether and repulsion (what energy()/properties() need); the property
// sweep inside gradient() only needs area, volume, curvature and rho.
template <bool FULL, bool WRITE_VPROPS>
__global__ void __launch_bounds__(TM_PROPS_THREADS)
k_props(MeshView mv, NlView nl, DevParams P, P4* pos_rw,
        double* __restrict__ vprops, double* __restrict__ partials,
        unsigned* __restrict__ ticket, double* __restrict__ sums) {
    int i = blockIdx.x * TM_PROPS_THREADS + threadIdx.x;
    double acc[6] = {0, 0, 0, 0, 0, 0};   // area volume curvature bending tether repulsion
    if (i < mv.V) {
        const int beg = mv.ring_off[i], d = mv.ring_off[i + 1] - beg;
        const int* ring = mv.ring_v + beg;
        vertex_props<FULL, false>(mv, nl, P, i, d, [ring](int s) { return __ldg(ring + s); }, acc);
        pos_rw[i].w = acc[2] / acc[0];        // rho_i = c_i / a_i rides in the position record
        if (WRITE_VPROPS) {
            double* o = vprops + 6 * (size_t)i;
#pragma unroll
            for (int q = 0; q < 6; q++) o[q] = acc[q];
        }
    }

    // block partials -> last block sums all partials in a fixed order
    __shared__ double bsum[6];
    __shared__ bool last;
    block_reduce<TM_PROPS_THREADS, 6>(acc, bsum);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 6; q++) partials[(size_t)q * gridDim.x + blockIdx.x] = bsum[q];
        __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double tot[6] = {0, 0, 0, 0, 0, 0};
        for (unsigned b = threadIdx.x; b < gridDim.x; b += TM_PROPS_THREADS) {
#pragma unroll
            for (int q = 0; q < 6; q++) tot[q] += __ldcg(partials + (size_t)q * gridDim.x + b);
        }
        block_reduce<TM_PROPS_THREADS, 6>(tot, bsum);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int q = 0; q < 6; q++) sums[q] = bsum[q];
            *ticket = 0;
        }
    }
}

// ---- K3 + K4: gather gradient ----------------------------------------------------
// Reads only: every ring / list neighbour costs one 256-bit read-only load of its
// (x, y, z, rho) record.
__global__ void __launch_bounds__(TM_GRAD_THREADS)
k_grad(MeshView mv, NlView nl, DevParams P, const double* __restrict__ sums, double* __restrict__ grad) {
    int i = blockIdx.x * TM_GRAD_THREADS + threadIdx.x;
    if (i >= mv.V) return;

    // chain-rule factors of the global penalties (kernel.h:27-71,124-137)
    const double fA = 2.0 * P.kappa_a / P.ref_area * (sums[0] / P.ref_area - 1.0);
    const double fV = 2.0 * P.kappa_v / P.ref_volume * (sums[1] / P.ref_volume - 1.0);
    const double fC = 2.0 * P.kappa_c / P.ref_curvature * (sums[2] / P.ref_curvature - 1.0);
    const double kb = P.kappa_b, kt = P.kappa_t;

    const P4 ri = ldrec(mv.pos, i);
    const d3 xi = mk3(ri.x, ri.y, ri.z);
    d3 g = mk3(0, 0, 0);
    int beg = mv.ring_off[i], d = mv.ring_off[i + 1] - beg;
    if (d > 0) {
        const int* ring = mv.ring_v + beg;
        const double rho_i = ri.w;
        d3 nsum = mk3(0, 0, 0);      // sum of face normals  -> volume gradient
        d3 esum = mk3(0, 0, 0);      // sum of (e_{s-1} - e_s) over faces (0 for closed fans)

        int r_prev = __ldg(ring + d - 1), r_cur = __ldg(ring);
        const P4 q_prev = ldrec(mv.pos, r_prev & TM_RING_MASK), q_cur = ldrec(mv.pos, r_cur & TM_RING_MASK);
        d3 e_prev = mk3(q_prev.x, q_prev.y, q_prev.z) - xi;
        d3 e_cur = mk3(q_cur.x, q_cur.y, q_cur.z) - xi;
        double rho_prev = q_prev.w, rho_cur = q_cur.w;
        double l2_prev = dot3(e_prev, e_prev), l2_cur = dot3(e_cur, e_cur);
        double p_cur = dot3(e_cur, e_prev);
        d3 n_cur = cross3(e_cur, e_prev);
        double nn_cur = dot3(n_cur, n_cur);
        double in_cur = rsqrt_pos(nn_cur);            // 1/|n_s|

        for (int s = 0; s < d; s++) {
            int sn = s + 1 == d ? 0 : s + 1;
            int r_next = __ldg(ring + sn);
            const P4 q_next = ldrec(mv.pos, r_next & TM_RING_MASK);
            d3 e_next = mk3(q_next.x, q_next.y, q_next.z) - xi;
            double rho_next = q_next.w;
            double l2_next = dot3(e_next, e_next);
            double p_next = dot3(e_next, e_cur);
            d3 n_next = cross3(e_next, e_cur);
            double nn_next = dot3(n_next, n_next);
            double in_next = rsqrt_pos(nn_next);

            if (r_cur & TM_RING_FACE) {
                const double il = rsqrt_pos(l2_cur);   // 1/l_s
                const double l = l2_cur * il;
                const double inn_cur = in_cur * in_cur;
                double cE = 0.0, cEp = 0.0, cN = 0.0, cNn = 0.0;

                // spoke edge (i, j_s): owners i and j_s   (mesh_properties.cpp:88-98)
                const double W = 0.5 * fC + kb * (rho_i + rho_cur);
                if (r_cur & TM_RING_OFACE) {
                    const double th = atan2(-l * dot3(e_next, n_cur), dot3(n_cur, n_next));
                    cE -= W * th * il;                                   // theta * dl/dx_i
                    cN += W * (l2_cur - p_cur) * inn_cur;                // l * dtheta/dx_i
                    cNn += W * (l2_cur - p_next) * (in_next * in_next);
                }
                // face (i, j_s, j_{s-1}): owners i, j_s, j_{s-1}   (:103-118)
                double U = fA - (2.0 / 3.0) * kb * (rho_i * rho_i + rho_cur * rho_cur + rho_prev * rho_prev);
                if (P.bond_type == TM_BOND_AREA) {
                    const double a_f = 0.5 * nn_cur * in_cur;
                    U += kt * 2.0 * (a_f / P.a0 - 1.0) / P.a0;
                }
                const double h = 0.5 * U * in_cur;                       // dA_f/dx_i in the e basis
                cEp += h * (p_cur - l2_cur);
                cE -= h * (l2_prev - p_cur);
                // link edge (j_s, j_{s-1}) seen from its apex i: owners j_s, j_{s-1}   (:121-129)
                const double Wl = 0.5 * fC + kb * (rho_cur + rho_prev);
                cN -= Wl * (l2_cur + l2_prev - 2.0 * p_cur) * inn_cur;
                // tether bond (:132-133)
                if (P.bond_type == TM_BOND_EDGE) {
                    if (l > P.lc0 || l < P.lc1) cE -= kt * tether_edge_fac(l, P) * il;
                }
                axpy3(g, cE, e_cur);
                axpy3(g, cEp, e_prev);
                axpy3(g, cN, n_cur);
                axpy3(g, cNn, n_next);
                nsum = nsum + n_cur;
                esum = esum + (e_prev - e_cur);
            }
            r_prev = r_cur; r_cur = r_next;
            e_prev = e_cur; e_cur = e_next;
            rho_prev = rho_cur; rho_cur = rho_next;
            l2_prev = l2_cur; l2_cur = l2_next;
            p_cur = p_next; n_cur = n_next; nn_cur = nn_next; in_cur = in_next;
        }
        // volume: sum_f (x_j x x_k)/6 = (sum_f n_f + x_i x sum_f (e_{s-1} - e_s))/6
        d3 vg = nsum + cross3(xi, esum);
        axpy3(g, fV / 6.0, vg);
    }

    // repulsion (mesh_properties.cpp:139-146, mesh_repulsion.h:56-78)
    if (nl.kind >= 0 && P.rep_lc1 > 0.0 && P.kappa_r != 0.0) {
        d3 gr = mk3(0, 0, 0);
        const double lc1 = P.rep_lc1;
        const int rr = P.rep_r;
        for_each_pair_fast<false, true>(mv, nl, i, xi, lc1 * lc1, [&gr, lc1, rr](int, d3 dv, double l2) {
            double sl;
            repulsion_pair<true>(l2, lc1, rr, sl);
            axpy3(gr, sl, dv);
        });
        axpy3(g, 2.0 * P.kappa_r, gr);
    }

    double* o = grad + 3 * (size_t)i;
    o[0] = g.x; o[1] = g.y; o[2] = g.z;
}

}  // namespace

int tm_launch_props(tm_mesh* m, const NlView& nl, const tm_energy_params& p, bool full, bool write_vprops) {
    if (tm_ensure_layout(m)) return 1;
    int nb = (m->V + TM_PROPS_THREADS - 1) / TM_PROPS_THREADS;
    TM_CHECK(nb <= TM_MAX_PARTIAL_BLOCKS, "mesh too large for the reduction workspace (%d vertices)", m->V);
    DevParams P = tm_to_dev(p);
    MeshView mv = m->view();
    if (write_vprops) {
        if (!m->vprops) TM_CUDA(cudaMalloc(&m->vprops, sizeof(double) * 6 * (size_t)m->V));
        k_props<true, true><<<nb, TM_PROPS_THREADS, 0, m->stream>>>(mv, nl, P, m->pos, m->vprops,
                                                                    m->partials, m->ticket, m->sums);
    } else if (full) {
        k_props<true, false><<<nb, TM_PROPS_THREADS, 0, m->stream>>>(mv, nl, P, m->pos, nullptr,
                                                                     m->partials, m->ticket, m->sums);
    } else {
        k_props<false, false><<<nb, TM_PROPS_THREADS, 0, m->stream>>>(mv, nl, P, m->pos, nullptr,
                                                                      m->partials, m->ticket, m->sums);
    }
    TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    m->geo_epoch = m->pos_epoch;        // sums[0..2] and rho now belong to the current positions
    return 0;
}

int tm_launch_grad(tm_mesh* m, const NlView& nl, const tm_energy_params& p) {
    if (tm_ensure_layout(m)) return 1;
    int nb = (m->V + TM_GRAD_THREADS - 1) / TM_GRAD_THREADS;
    DevParams P = tm_to_dev(p);
    k_grad<<<nb, TM_GRAD_THREADS, 0, m->stream>>>(m->view(), nl, P, m->sums, m->grad);
    TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    return 0;
}
