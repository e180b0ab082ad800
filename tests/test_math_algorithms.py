"""The branch-free fp64 elementary functions of the sweeps (trimem_b200/csrc/tm_math.cuh), restated in
numpy operation for operation and compared with numpy's own over dense samples. (The CUDA code itself
is covered by the parity tests: every dihedral angle and every flat-bottom potential goes through it.)"""
import numpy as np

ATAN_C = np.array([5.8823529411764705e-02, -6.6666666666666666e-02, 7.6923076923076927e-02, -9.0909090909090912e-02,
                   1.1111111111111111e-01, -1.4285714285714285e-01, 2.0e-01, -3.3333333333333331e-01])
ATAN_T = np.array([
    [0.0, 0.24497866312686414, 0.46364760900080615, 0.6435011087932844, 0.7853981633974483],
    [1.5707963267948966, 1.3258176636680326, 1.1071487177940904, 0.9272952180016122, 0.7853981633974483],
    [3.141592653589793, 2.896613990462929, 2.677945044588987, 2.498091544796509, 2.356194490192345],
    [1.5707963267948966, 1.8157749899217608, 2.0344439357957027, 2.214297435588181, 2.356194490192345]])


def rcp_newton(x, seed):
    """rcp_pos (tm_common.cuh): one third-order step on a 2^-20 seed"""
    e = 1.0 - x * seed             # fma(-x, r, 1)
    return seed + seed * (e + e * e)


def atan2_fast(y, x, rng):
    ax, ay = np.abs(x), np.abs(y)
    swap = ay > ax
    mx, mn = np.where(swap, ay, ax) + 1.0e-300, np.where(swap, ax, ay)
    # MUFU.RCP64H: a ~20-bit reciprocal seed (modelled as the exact reciprocal with a 2^-20 relative error)
    seed = (1.0 / mx) * (1.0 + rng.uniform(-1, 1, size=mx.shape) * 2.0 ** -20)
    M = 1688849860263936.0
    t = mn * seed + M
    c = t - M
    k = np.clip(np.rint(4.0 * c).astype(int), 0, 4)
    o = swap.astype(int) + 2 * np.signbit(x).astype(int)
    den = mx + c * mn
    r = (mn - c * mx) * rcp_newton(den, (1.0 / den) * (1.0 + rng.uniform(-1, 1, size=mx.shape) * 2.0 ** -20))
    r = np.where((o == 1) | (o == 2), -r, r)
    s = r * r
    p = np.full_like(s, ATAN_C[0])
    for coef in ATAN_C[1:]:
        p = p * s + coef
    a = (r * s * p + r) + ATAN_T[o, k]
    return np.copysign(a, y)


def test_atan2_algorithm_accuracy():
    rng = np.random.default_rng(0)
    n = 400000
    ang = rng.uniform(-np.pi, np.pi, n)
    rad = 10.0 ** rng.uniform(-30, 30, n)
    x, y = rad * np.cos(ang), rad * np.sin(ang)
    # plus the neighbourhoods of the axes, the diagonals and the interval boundaries q = (2j+1)/8
    q = np.concatenate([np.linspace(0, 1, 20001)] + [(2 * j + 1) / 8 + np.linspace(-1e-5, 1e-5, 2001) for j in range(4)])
    xs = np.concatenate([x, np.ones_like(q), q, -np.ones_like(q), -q, np.ones_like(q), q])
    ys = np.concatenate([y, q, np.ones_like(q), q, np.ones_like(q), -q, -np.ones_like(q)])
    got = atan2_fast(ys, xs, rng)
    want = np.arctan2(ys, xs)
    err = np.abs(got - want)
    # relative to the result, except where the result is ~0 against pi-sized terms being cancelled
    tol = 6.0e-16 * np.maximum(np.abs(want), 2.0e-16) + 2.3e-16 * (np.abs(want) > 0.24)
    assert np.all(err <= tol), float(np.max(err / np.maximum(np.abs(want), 1e-300)))
    # the interval selector never drives the series argument beyond its design range
    assert atan2_fast(np.array([0.0]), np.array([0.0]), rng)[0] == 0.0
    assert atan2_fast(np.array([0.0]), np.array([-2.0]), rng)[0] == np.pi
    assert atan2_fast(np.array([-0.0]), np.array([-2.0]), rng)[0] == -np.pi


def test_exp_nonpos_algorithm_accuracy():
    C = [1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
         2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,
         8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 5.0e-01]
    t = -10.0 ** np.random.default_rng(1).uniform(-8, np.log10(700.0), 300000)
    kd = np.rint(t * 1.4426950408889634)
    r = (t - kd * 6.93147180369123816490e-01) - kd * 1.90821492927058770002e-10
    p = np.full_like(r, C[0])
    for c in C[1:]:
        p = p * r + c
    p = p * r + 1.0
    p = p * r + 1.0
    got = np.ldexp(p, kd.astype(int))
    want = np.exp(t)
    assert np.max(np.abs(got - want) / want) < 4.0e-16


def test_rsqrt_rcp_refinement_accuracy():
    """One cubic step on a seed of relative error <= 2^-20 (what MUFU delivers on B200, scripts/ubench/seed_accuracy.cu)."""
    rng = np.random.default_rng(2)
    x = 10.0 ** rng.uniform(-30, 30, 400000)
    err = (1.0 + rng.uniform(-1, 1, size=x.shape) * 2.0 ** -20)
    r = rcp_newton(x, (1.0 / x) * err)
    assert np.max(np.abs(r * x - 1.0)) < 5e-16
    y = (1.0 / np.sqrt(x)) * err
    e = 1.0 - (x * y) * y
    y = y + (y * e) * (0.375 * e + 0.5)
    assert np.max(np.abs(y * np.sqrt(x) - 1.0)) < 5e-16


def test_exp_table_algorithm_accuracy():
    """exp_tab (tm_math.cuh): 2^(j/64) table + degree-5 polynomial, the pair sweep's exponential."""
    tab = 2.0 ** (np.arange(64) / 64.0)
    t = -10.0 ** np.random.default_rng(3).uniform(-8, np.log10(700.0), 300000)
    kd = np.rint(t * 92.33248261689366)
    k = kd.astype(np.int64)
    r = (t - kd * 0.010830424609594047) - kd * 8.66550983900947e-11
    assert np.max(np.abs(r)) < 0.0055
    p = r * 8.3333333333333332e-03 + 4.1666666666666664e-02
    for c in (1.6666666666666666e-01, 0.5, 1.0, 1.0):
        p = p * r + c
    got = np.ldexp(p * tab[k & 63], (k >> 6).astype(int))
    want = np.exp(t)
    assert np.max(np.abs(got - want) / want) < 4.0e-16
