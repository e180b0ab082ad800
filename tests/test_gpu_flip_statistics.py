"""Sweep-level flip statistics against the REFERENCE (north_star: "flip sweeps are checked by
acceptance-rate and energy-histogram agreement"). The reference's flip / pflip are clock-seeded
(flips.cpp:19-20), so sequences cannot be compared; distributions can. tests/golden/flip_statistics.npz holds
samples drawn from the reference build (tests/golden/make_golden.py flip_statistics):

  small  ico3_flip (642 vertices, the reference's tests/test_flips.py parameters), flip_ratio 0.1: 240 independent
         chains of 30 sweeps from the initial mesh, for flip and for pflip (one OpenMP thread); per chain and sweep
         the accepted flips and the energy after the sweep.
  large  ico6_flip (40,962 vertices, config 2's physical variant, all kappa = 1): chains of 100 sweeps, flip (1 thread)
         and pflip (8 threads).

Here the same chains run on the device, in both modes (serial warp, conflict-free batches), and are compared with
two-sample tests whose false-alarm probability is ~1e-3 per assertion.

What is compared with what: the device's pflip attempts EVERY proposal (conflicting ones wait for a later batch),
like the reference's flip and like its pflip on one thread; the reference's pflip on T threads drops a proposal whose
patch is locked (flips.cpp:118-124) and proposes T*floor(n/T) of them, so its acceptance COUNT depends on T. On the
large mesh a conflict has probability ~2e-3 per proposal and the 8-thread samples are comparable; on the small one
only the 1-thread samples are."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

GOLD = cases.load_golden("flip_statistics")


def ks_2samp(a, b):
    """two-sample Kolmogorov-Smirnov statistic and its 0.1 % critical value"""
    a, b = np.sort(a), np.sort(b)
    allv = np.concatenate([a, b])
    d = np.max(np.abs(np.searchsorted(a, allv, side="right") / len(a) - np.searchsorted(b, allv, side="right") / len(b)))
    return d, 1.95 * np.sqrt((len(a) + len(b)) / (len(a) * len(b)))


def device_chains(m, fn, chains, sweeps, ratio, seed0):
    acc = np.zeros((chains, sweeps), np.int64)
    ener = np.zeros((chains, sweeps))
    for c in range(chains):
        mesh, es = cases.build(m, "ico3_flip")
        for s in range(sweeps):
            acc[c, s] = fn(mesh, es, ratio, seed=seed0 + 1000 * c + s)
            ener[c, s] = es.energy(mesh)
    return acc, ener


@pytest.mark.skipif(GOLD is None, reason="tests/golden/flip_statistics.npz missing")
@pytest.mark.parametrize("mode,chains", [("flip", 64), ("pflip", 160)])
def test_small_mesh_distributions(gpu_core, mode, chains):
    m = gpu_core
    _, sweeps, ratio = cases.FLIPSTAT_SMALL
    acc, ener = device_chains(m, getattr(m, mode), chains, sweeps, ratio, 7 if mode == "flip" else 500007)
    burn = sweeps // 3
    for ref in ("flip", "pflip"):                 # the reference's two functions are one Markov kernel on one thread
        racc, rener = GOLD[f"small_{ref}_acc"], GOLD[f"small_{ref}_energy"]
        # energy histogram after the last sweep, and half way
        for s in (sweeps // 2, sweeps - 1):
            d, crit = ks_2samp(ener[:, s], rener[:, s])
            assert d < crit, (mode, ref, s, d, crit)
        # acceptance: per-chain mean of the accepted flips per sweep after burn-in (chains are independent)
        a, r = acc[:, burn:].mean(axis=1), racc[:, burn:].mean(axis=1)
        z = (a.mean() - r.mean()) / np.sqrt(a.var(ddof=1) / len(a) + r.var(ddof=1) / len(r))
        assert abs(z) < 3.5, (mode, ref, a.mean(), r.mean(), z)
        d, crit = ks_2samp(a, r)
        assert d < crit, (mode, ref, d, crit)
        # the relaxation itself: mean accepted flips of the first sweeps
        for s in (0, 1, 2, 5):
            a, r = acc[:, s].astype(float), racc[:, s].astype(float)
            z = (a.mean() - r.mean()) / np.sqrt(a.var(ddof=1) / len(a) + r.var(ddof=1) / len(r))
            assert abs(z) < 3.5, (mode, ref, s, a.mean(), r.mean(), z)


@pytest.mark.skipif(GOLD is None, reason="tests/golden/flip_statistics.npz missing")
def test_large_mesh_trajectories(gpu_core):
    """40,962 vertices, 100 sweeps of pflip (and 6 of the serial flip): accepted flips per sweep and the energy
    along the chain against the band of the reference's chains. A sweep proposes 12,288 flips, so its acceptance
    count is binomial-like with a standard deviation of ~55 around a mean that both implementations must share."""
    m = gpu_core
    nsw, ratio, plan = cases.FLIPSTAT_LARGE
    racc = np.concatenate([GOLD[f"large_{fn}{thr}_acc"] for fn, thr, _ in plan]).astype(float)
    rener = np.concatenate([GOLD[f"large_{fn}{thr}_energy"] for fn, thr, _ in plan])
    mesh, es = cases.build(m, "ico6_flip")
    assert abs(es.energy(mesh) - float(GOLD["large_energy0"])) <= 1e-10 * abs(float(GOLD["large_energy0"]))
    acc, ener = [], []
    for s in range(nsw):
        acc.append(m.pflip(mesh, es, ratio, seed=900 + s))
        ener.append(es.energy(mesh))
    acc, ener = np.array(acc, float), np.array(ener)
    n_prop = int(mesh.n_edges() * ratio)
    mu, sd_chain = racc.mean(axis=0), racc.std(axis=0, ddof=1)
    # per sweep: within 5 binomial standard deviations (+ the reference's own chain-to-chain spread) of the reference mean
    sd = np.sqrt(n_prop * (mu / n_prop) * (1 - mu / n_prop) * (1 + 1 / len(racc)))
    assert np.all(np.abs(acc - mu) < 5 * sd + 3 * sd_chain), np.max(np.abs(acc - mu) / sd)
    # windows of 10 sweeps: means agree to a fraction of a percent
    for w in range(0, nsw, 10):
        a, r = acc[w:w + 10].mean(), racc[:, w:w + 10].mean()
        assert abs(a - r) < 5 * sd[w:w + 10].mean() / np.sqrt(10) + 3 * racc[:, w:w + 10].mean(axis=1).std(ddof=1), (w, a, r)
    # total acceptance rate of the chain
    assert abs(acc.sum() - racc.sum(axis=1).mean()) < 5 * np.sqrt((sd ** 2).sum()) + 3 * racc.sum(axis=1).std(ddof=1)
    # energy along the chain: inside the reference band (its chains differ by a few percent from one another)
    emu, esd = rener.mean(axis=0), rener.std(axis=0, ddof=1)
    for s in (0, 1, 4, 9, 19, 49, 99):
        assert abs(ener[s] - emu[s]) < 5 * esd[s] + 0.01 * abs(emu[s]), (s, ener[s], emu[s], esd[s])
    # the serial device sweep on the same mesh: its first sweeps
    mesh2, es2 = cases.build(m, "ico6_flip")
    for s in range(6):
        a = m.flip(mesh2, es2, ratio, seed=1900 + s)
        assert abs(a - mu[s]) < 5 * sd[s] + 3 * sd_chain[s], (s, a, mu[s])
        assert abs(es2.energy(mesh2) - emu[s]) < 5 * esd[s] + 0.01 * abs(emu[s])
