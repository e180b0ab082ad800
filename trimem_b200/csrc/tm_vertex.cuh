// Per-vertex property evaluation shared by the property sweep (tm_energy.cu) and the
// flip kernels (tm_flips.cu), so that both produce bit-identical per-vertex values.
//
// Reference: vertex_properties (src/mesh_properties.cpp:12-49), tether potentials
// (src/mesh_tether.h:29-121), surface repulsion (src/mesh_repulsion.h:28-54),
// trimem_energy (src/kernel.h:110-122).
#pragma once
#include "tm_common.cuh"
#include "tm_neigh.cuh"
#include "tm_math.cuh"

struct DevParams {
    double kappa_b, kappa_a, kappa_v, kappa_c, kappa_t, kappa_r;
    double ref_area, ref_volume, ref_curvature;
    double lc0, lc1, a0;
    int bond_r, bond_type;
    double rep_lc1;
    int rep_r;
    // derived on the host once per launch (the sweeps multiply instead of dividing)
    double inv_ref_area, inv_ref_volume, inv_ref_curvature, inv_a0, lc0sq, lc1sq;
};

inline DevParams tm_to_dev(const tm_energy_params& p) {
    DevParams P;
    P.kappa_b = p.kappa_b; P.kappa_a = p.kappa_a; P.kappa_v = p.kappa_v;
    P.kappa_c = p.kappa_c; P.kappa_t = p.kappa_t; P.kappa_r = p.kappa_r;
    P.ref_area = p.ref_area; P.ref_volume = p.ref_volume; P.ref_curvature = p.ref_curvature;
    P.lc0 = p.bond_lc0; P.lc1 = p.bond_lc1; P.a0 = p.bond_a0;
    P.bond_r = p.bond_r; P.bond_type = p.bond_type;
    P.rep_lc1 = p.rep_lc1; P.rep_r = p.rep_r;
    P.inv_ref_area = 1.0 / p.ref_area; P.inv_ref_volume = 1.0 / p.ref_volume;
    P.inv_ref_curvature = 1.0 / p.ref_curvature; P.inv_a0 = 1.0 / p.bond_a0;
    P.lc0sq = p.bond_lc0 * p.bond_lc0; P.lc1sq = p.bond_lc1 * p.bond_lc1;
    return P;
}

#ifdef __CUDACC__
__device__ __forceinline__ double powi(double x, int r) {
    if (r == 2) return x * x;
    int a = r < 0 ? -r : r;
    double acc = 1.0, b = x;
    while (a) { if (a & 1) acc *= b; b *= b; a >>= 1; }
    return r < 0 ? 1.0 / acc : acc;
}

// FlatBottomEdgePenalty::vertex_property (mesh_tether.h:40-54)
__device__ __forceinline__ double tether_edge(double l, const DevParams& P) {
    if (l > P.lc0) return powi((double)P.bond_r, P.bond_r + 1) * powi(l - P.lc0, P.bond_r);
    if (l < P.lc1) return tm_math::exp_nonpos(l / (l - P.lc1)) / powi(l, P.bond_r);
    return 0.0;
}
// slope factor of FlatBottomEdgePenalty::vertex_property_grad (mesh_tether.h:62-76)
__device__ __forceinline__ double tether_edge_fac(double l, const DevParams& P) {
    double fac = 0.0;
    if (l > P.lc0) fac = powi((double)P.bond_r, P.bond_r + 2) * powi(l - P.lc0, P.bond_r - 1);
    if (l < P.lc1) {
        double lr = powi(l, P.bond_r);
        double lmc = l - P.lc1;
        fac -= P.lc1 / (lmc * lmc) + P.bond_r / l;
        fac /= lr;
        fac *= tm_math::exp_nonpos(l / lmc);
    }
    return fac;
}

// trimem_energy (kernel.h:110-122)
__device__ __forceinline__ double energy_of(const DevParams& P, const double* s) {
    double e = 0.0, d;
    d = s[0] / P.ref_area - 1.0;      e += P.kappa_a * d * d;
    d = s[1] / P.ref_volume - 1.0;    e += P.kappa_v * d * d;
    d = s[2] / P.ref_curvature - 1.0; e += P.kappa_c * d * d;
    e += P.kappa_t * s[4];
    e += P.kappa_r * s[5];
    e += P.kappa_b * s[3];
    return e;
}

// surface-repulsion pair potential exp(l/(l-lc1)) / l^r (mesh_repulsion.h:44-53) from the
// squared length; `sl` receives the slope factor of mesh_repulsion.h:65-75,
//   -(lc1/(l-lc1)^2 + r/l) / l^(r+1) * exp(l/(l-lc1)),  when WITH_SLOPE.
// One rsqrt, one reciprocal and one exp per pair, no division.
template <bool WITH_SLOPE>
__device__ __forceinline__ double repulsion_pair(double l2, double lc1, int r, double& sl) {
    const double il = rsqrt_pos(l2);
    const double l = l2 * il;
    sl = 0.0;
    if (!(l < lc1)) return 0.0;            // the rounded length decides, as in the reference
    const double w = rcp_pos(l - lc1);     // 1/(l - lc1) < 0
    const double ex = tm_math::exp_nonpos(l * w);
    const double ilr = r == 2 ? il * il : powi(il, r);
    const double phi = ex * ilr;
    if (WITH_SLOPE) sl = -fma(lc1 * w, w, (double)r * il) * (phi * il);
    return phi;
}

// sqrt for non-negative arguments: x * rsqrt(x), 0 for 0 (branch-free select)
__device__ __forceinline__ double sqrt_nn(double x) {
    const double y = x * rsqrt_pos(x);
    return x > 0.0 ? y : 0.0;
}

// Six properties of vertex i whose clockwise one-ring is ring(0..d-1) (neighbour id |
// TM_RING_FACE | TM_RING_OFACE). With e_s = x_{j_s} - x_i, n_s = e_s x e_{s-1}:
//   A_f = |n_s|/2, V_f = x_i . n_s / 6, theta_s = atan2(-|e_s| (e_{s+1}.n_s), n_s.n_{s+1}).
// FULL adds tether and repulsion; LIVE selects the exclusion walk used by cell lists.
template <bool FULL, bool LIVE, class RingFn>
__device__ __forceinline__ void vertex_props(const MeshView& mv, const NlView& nl, const DevParams& P,
                                             int i, int d, RingFn ring, double out[6]) {
    const d3 xi = ldpos(mv.pos, i);
    double area = 0, vol = 0, curv = 0, teth = 0;
    if (d > 0) {
        // ring state of one slot: spoke e_s, face normal n_s = e_s x e_{s-1}, ring word
        struct Slot { d3 e, n; int r; };
        // one ring slot; `cur` is read, `nxt` is produced. The caller alternates two Slot objects
        // so the rotation cur <- nxt costs no register moves.
        auto step = [&](const Slot& cur, Slot& nxt, int s) {
            const int sn = s + 1 == d ? 0 : s + 1;
            nxt.r = ring(sn);
            nxt.e = ldpos(mv.pos, nxt.r & TM_RING_MASK) - xi;
            nxt.n = cross3(nxt.e, cur.e);
            if (cur.r & TM_RING_FACE) {           // `not he.is_boundary()`
                const double l = sqrt_nn(dot3(cur.e, cur.e));
                const double a_f = 0.5 * sqrt_nn(dot3(cur.n, cur.n));
                area += a_f;
                vol += dot3(xi, cur.n);           // = (x_j x x_k) . x_i
                if (cur.r & TM_RING_OFACE) {      // dihedral is 0 on boundary edges
                    const double th = tm_math::atan2_fast(-l * dot3(nxt.e, cur.n), dot3(cur.n, nxt.n));
                    curv = fma(l, th, curv);
                }
                if (FULL) {
                    if (P.bond_type == TM_BOND_EDGE) { if (l > P.lc0 || l < P.lc1) teth += tether_edge(l, P); }
                    else { const double q = a_f / P.a0 - 1.0; teth += q * q; }
                }
            }
        };
        Slot a, b;
        a.r = ring(0);
        a.e = ldpos(mv.pos, a.r & TM_RING_MASK) - xi;
        a.n = cross3(a.e, ldpos(mv.pos, ring(d - 1) & TM_RING_MASK) - xi);
        for (int s = 0; s < d; s += 2) {
            step(a, b, s);
            if (s + 1 < d) step(b, a, s + 1);
        }
    }
    area = area / 3.0;
    vol = vol / 18.0;                              // /6 per face, /3 multiplicity
    curv = curv / 4.0;                             // 0.5 l theta per half-edge, /2 multiplicity
    double rep = 0.0;
    if (FULL) {
        teth = teth / (P.bond_type == TM_BOND_EDGE ? 2.0 : 3.0);
        if (nl.kind >= 0 && P.rep_lc1 > 0.0) {
            const double lc1 = P.rep_lc1;
            const int rr = P.rep_r;
            for_each_pair_fast<LIVE, false>(mv, nl, i, xi, lc1 * lc1, [&rep, lc1, rr](int, d3, double l2) {
                double sl;
                rep += repulsion_pair<false>(l2, lc1, rr, sl);
            });
        }
    }
    out[0] = area; out[1] = vol; out[2] = curv;
    out[3] = 2.0 * curv * curv / area;
    out[4] = teth; out[5] = rep;
}
#endif
