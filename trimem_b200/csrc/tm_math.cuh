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

// 2^(j/64), j = 0 .. 63: table of the pair sweep's exponential (exp_tab below)
static __constant__ double EXP2_64[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951,
};
// exp(t) for t <= 0 with the table in shared memory (`tab` = EXP2_64 copied there by the block):
//   t = (64 q + j) ln2/64 + r, |r| <= ln2/128  ->  exp(t) = 2^q * tab[j] * (1 + r + ... + r^5/120)
// (truncation 3.4e-17 relative). 10 fp64 operations and one shared-memory load, against 18 for the
// table-free exp_nonpos: the pair sweep is bound by instruction issue and evaluates this once per pair.
__device__ __forceinline__ double exp_tab(double t, const double* tab) {
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: round-to-nearest integer
    double kd = fma(t, 92.33248261689366, MAGIC);            // 64 / ln 2
    const int k = __double2loint(kd);                        // k <= 0
    kd -= MAGIC;
    double r = fma(kd, -0.010830424609594047, t);            // ln2/64 hi (26 trailing zero bits)
    r = fma(kd, -8.66550983900947e-11, r);                   // ln2/64 lo
    double p = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p *= tab[k & 63];
    // p * 2^(k >> 6) through the exponent field (p in [0.99, 2.0], k >= -64 * 1021 keeps it normal)
    return __hiloint2double(__double2hiint(p) + ((k >> 6) << 20), __double2loint(p));
}

// The same with the power of two clamped at 2^-1021 (one integer max) instead of a comparison of t with -708 by
// the caller: for t < ~-708 the result is ~1e-308 (normal) where exp() underflows to 0. Valid for
// -2^24 < t <= 0 (the integer k must fit 32 bits). Used by the pair sweep (pair_eval).
__device__ __forceinline__ double exp_tab_clamped(double t, const double* tab) {
    const double MAGIC = 6755399441055744.0;
    double kd = fma(t, 92.33248261689366, MAGIC);
    const int k = __double2loint(kd);
    kd -= MAGIC;
    double r = fma(kd, -0.010830424609594047, t);
    r = fma(kd, -8.66550983900947e-11, r);
    double p = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p *= tab[k & 63];
    return __hiloint2double(__double2hiint(p) + (max(k >> 6, -1021) << 20), __double2loint(p));
}

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
