// fp64 elementary functions of the sweeps, written for the B200 instruction mix.
//
// The reference calls libm: std::acos inside OpenMesh's calc_dihedral_angle /
// calc_sector_angle (mesh_util.h:132,148-161), std::exp and std::pow in the tether
// and repulsion potentials (mesh_tether.h:46-71, mesh_repulsion.h:51,69). CUDA's own
// exp()/atan2() materialise every polynomial coefficient with two UMOVs per use and
// carry slow-path branches (atan2: ~35 DFMA inside ~200 SASS instructions; exp: 14
// DFMA inside ~60), and the sweeps turned out issue-bound on exactly that overhead
// (profiles/r01c). These versions keep the coefficients in the constant bank, where a
// DFMA reads them as a direct operand, and are branch-free.
//
// Accuracy: <= ~2 ulp over the argument ranges the sweeps produce (tests/test_gpu_math.py
// compares against numpy over dense samples); the parity bar of the path is 1e-10 relative.
#pragma once
#include "tm_common.cuh"

#ifdef __CUDACC__
namespace tm_math {

// 1/k!, k = 13 .. 2
static __constant__ double EXP_C[12] = {
    1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
    2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,
    8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 5.0e-01};

// (-1)^n / (2n+1), n = 8 .. 1   (atan r = r + r^3 * P(r^2))
static __constant__ double ATAN_C[8] = {
    5.8823529411764705e-02, -6.6666666666666666e-02, 7.6923076923076927e-02, -9.0909090909090912e-02,
    1.1111111111111111e-01, -1.4285714285714285e-01, 2.0e-01, -3.3333333333333331e-01};

// exp(t) for t <= 0 (t = l/(l - lc1) of the flat-bottom potentials). Returns 0 below -708
// (the true value is < 3.4e-308 there).
__device__ __forceinline__ double exp_nonpos(double t) {
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: round-to-nearest integer
    double kd = fma(t, 1.4426950408889634, MAGIC);
    const int k = __double2loint(kd);                        // k <= 0
    kd -= MAGIC;
    double r = fma(kd, -6.93147180369123816490e-01, t);      // ln2 hi (32 trailing zero bits)
    r = fma(kd, -1.90821492927058770002e-10, r);             // ln2 lo
    double p = EXP_C[0];
#pragma unroll
    for (int i = 1; i < 12; i++) p = fma(p, r, EXP_C[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    // p * 2^k through the exponent field (p in [0.70, 1.42], k >= -1021 keeps it normal)
    const double s = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    return t < -708.0 ? 0.0 : s;
}

// atan2(y, x), all quadrants, y == x == 0 -> 0.
//   q = min(|x|,|y|)/max(|x|,|y|) in [0,1];  c = q rounded to a multiple of 1/4;
//   atan q = atan c + atan r,  r = (q - c)/(1 + q c) = (mn - c mx)/(mx + c mn),  |r| <= 1/8 (+1e-6)
//   -> odd series to r^17.
// c only selects the interval, so it is taken from a 20-bit estimate of q (the MUFU seed alone, no Newton
// steps), and r needs ONE full-precision reciprocal. The octant bookkeeping (swap of the arguments,
// x < 0) and atan c are folded into one table: with o = swap + 2 (x < 0) the result is
//   ATAN_T[o][k] + sigma atan r,  sigma = -1 for o = 1, 2  (atan is odd: the sign goes onto r),
// one indexed constant load instead of a chain of selects. max(|x|,|y|) is offset by 1e-300, which is
// invisible for normal arguments and makes x = y = 0 come out as atan2(0, +-0) without a guard.
// ~25 fp64 operations, no branch.
static __constant__ double ATAN_T[32] = {
    // o = 0: atan c                       (k = 0 .. 4; 3 pad entries per row)
    0.0, 0.24497866312686414, 0.46364760900080615, 0.6435011087932844, 0.7853981633974483, 0, 0, 0,
    // o = 1: pi/2 - atan c
    1.5707963267948966, 1.3258176636680326, 1.1071487177940904, 0.9272952180016122, 0.7853981633974483, 0, 0, 0,
    // o = 2: pi - atan c
    3.141592653589793, 2.896613990462929, 2.677945044588987, 2.498091544796509, 2.356194490192345, 0, 0, 0,
    // o = 3: pi/2 + atan c
    1.5707963267948966, 1.8157749899217608, 2.0344439357957027, 2.214297435588181, 2.356194490192345, 0, 0, 0};

__device__ __forceinline__ double atan2_fast(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const bool swap = ay > ax;
    const double mx = (swap ? ay : ax) + 1.0e-300, mn = swap ? ax : ay;
    double seed;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(seed) : "d"(mx));
    const double M = 1688849860263936.0;                     // 1.5 * 2^50: ulp = 1/4
    const double t = fma(mn, seed, M);
    const double c = t - M;
    const unsigned neg = (unsigned)__double2hiint(x) >> 31;
    const unsigned o = (swap ? 1u : 0u) | (neg << 1);
    const unsigned idx = (o << 3) | ((unsigned)__double2loint(t) & 7u);      // k = 0 .. 4
    double r = fma(-c, mx, mn) * rcp_pos(fma(c, mn, mx));
    // sigma = -1 for o = 1, 2
    r = __hiloint2double(__double2hiint(r) ^ (int)(((o ^ (o >> 1)) & 1u) << 31), __double2loint(r));
    const double s = r * r;
    double p = ATAN_C[0];
#pragma unroll
    for (int i = 1; i < 8; i++) p = fma(p, s, ATAN_C[i]);
    const double a = fma(r * s, p, r) + ATAN_T[idx];
    return copysign(a, y);
}

}  // namespace tm_math
#endif
