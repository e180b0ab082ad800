"""Device-resident velocity-Verlet trajectory (tm_vv_integrate, core.vv_integrate) against the
reference's own Python integrator: the golden vectors were produced by running
src/trimem/mc/hmc.py::_vv_integration on the reference build (tests/golden/make_golden.py
trajectory_cases). Tolerances: the force agrees to 1e-10 norm-wise per evaluation (BASELINE.json);
over the 6-10 steps of a case the end point's DISPLACEMENT and the momenta are held to 1e-9
norm-wise, the energies to 1e-10 relative."""
import numpy as np
import pytest

import cases
from trimem_b200 import cabi, workload

pytestmark = pytest.mark.gpu
TOL_TRAJ = 1e-9
TOL_E = 1e-10


def nrm(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("name", list(cases.VV_CASES))
def test_pybind_vv_integrate_vs_reference(gpu_core, name):
    gold = cases.load_golden(name)
    case, p0, mass, dt, n = cases.vv_inputs(name)
    mesh, es = cases.build(gpu_core, case)
    x0 = mesh.points().copy()
    x, p, st = gpu_core.vv_integrate(mesh, es, p0, mass, dt, n)
    assert np.array_equal(mesh.points(), x0)                 # the mesh itself is left alone
    assert nrm(x - x0, gold["x"] - x0) < TOL_TRAJ
    assert nrm(p, gold["p"]) < TOL_TRAJ
    assert np.all(np.abs(st - gold["stats"]) <= TOL_E * np.abs(gold["stats"]) * np.array([1, 1, 10, 10]))
    # the same trajectory driven from Python through energy()/gradient(), as the reference does it
    def force(xx):
        np.copyto(mesh.points(), xx)
        return -es.gradient(mesh)
    xr, pr, a = x0, p0, force(x0)
    for _ in range(n):
        xr = xr + (pr * dt + 0.5 * a * dt**2) / mass
        an = force(xr)
        pr = pr + 0.5 * (a + an) * dt
        a = an
    np.copyto(mesh.points(), x0)
    # equal gradients -> bit-identical update arithmetic
    assert np.array_equal(x, xr) and np.array_equal(p, pr)


def test_c_abi_async_trajectory_and_port(gpu_core, port):
    """Through the C ABI with the momenta resident on the device, against the oracle port."""
    name = "vv_ico4_stress_verlet2"
    case, p0, mass, dt, n = cases.vv_inputs(name)
    v, f = cases.mesh_of(case)
    pm = port.TriMesh(v, f)
    pe = port.EnergyManager(pm, port.workload_params(pm, **cases.CASES[case][2]))
    xr, pr, sr = port.vv_integration(pm, pe, p0, mass, dt, n)
    mesh, es = cases.build(gpu_core, case)
    he_to, he_next, he_face, v_he, _ = mesh.halfedge_arrays()
    cm = cabi.Mesh(he_to, he_next, he_face, v_he, xyz=v)
    rp = es.eparams.repulse_params
    nl = cabi.NList(cm, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
    par = workload.abi_params(es)
    cm.set_momenta(p0)
    assert np.array_equal(cm.momenta(), p0)
    # in two legs: 4 steps, then the remaining ones (the state stays on the device in between)
    cm.vv_integrate_async(nl, par, mass, dt, 4, with_energies=True)
    st_a, _, _ = cm.vv_stats(par)
    cm.vv_integrate_async(nl, par, mass, dt, n - 4, with_energies=True)
    st_b, props0, props1 = cm.vv_stats(par)
    x, p = cm.positions(), cm.momenta()
    assert nrm(x - v, xr - v) < TOL_TRAJ and nrm(p, pr) < TOL_TRAJ
    assert abs(st_a[0] - sr[0]) <= TOL_E * abs(sr[0]) and abs(st_a[1] - sr[1]) <= TOL_E * abs(sr[1])
    assert abs(st_b[2] - sr[2]) <= 10 * TOL_E * abs(sr[2]) and abs(st_b[3] - sr[3]) <= 10 * TOL_E * abs(sr[3])
    assert st_a[2] == st_b[0] and st_a[3] == st_b[1]        # leg 2 starts where leg 1 ended, same bits
    # the energy the trajectory reports at its end is the energy of its end point
    e, props = cm.energy(nl, par)
    assert e == st_b[2] and np.array_equal(props, props1)


@pytest.mark.parametrize("subdiv,n,dt", [(6, 10, 1.0e-6), (8, 5, 1.0e-7)])
def test_reversibility_and_energy_conservation(gpu_core, subdiv, n, dt):
    """Size-independent properties at BASELINE sizes: velocity Verlet is time reversible (integrate,
    flip the momenta, integrate again -> the start) and conserves H = U + K to O(dt^2)."""
    mesh, es, x0 = workload.make_case(gpu_core, subdiv, repulsion="stress", n_search="verlet-list")
    V = mesh.n_vertices()
    p0 = np.random.default_rng(5).normal(size=(V, 3))
    mass = 1.0          # (the stress potential is stiff: the reference itself diverges at dt = 5e-6 here)
    x1, p1, st = gpu_core.vv_integrate(mesh, es, p0, mass, dt, n)
    h0, h1 = st[0] + st[1], st[2] + st[3]
    assert abs(h1 - h0) < 2e-3 * abs(h0)
    assert np.abs(x1 - x0).max() > 0
    np.copyto(mesh.points(), x1)
    x2, p2, st2 = gpu_core.vv_integrate(mesh, es, -p1, mass, dt, n)
    assert np.linalg.norm(x2 - x0) < 1e-9 * np.linalg.norm(x1 - x0)
    assert nrm(-p2, p0) < 1e-9
    assert abs(st2[2] - st[0]) <= 1e-10 * abs(st[0])
    # zero steps: nothing moves, energies of the start point
    np.copyto(mesh.points(), x0)
    x3, p3, st3 = gpu_core.vv_integrate(mesh, es, p0, mass, dt, 0)
    assert np.array_equal(x3, x0) and np.array_equal(p3, p0) and st3[0] == st3[2] and st3[1] == st3[3]
    assert abs(st3[0] - es.energy(mesh)) <= 1e-12 * abs(st3[0])


def test_argument_errors(gpu_core):
    mesh, es = cases.build(gpu_core, "ico3_edge")
    p0 = np.zeros((mesh.n_vertices(), 3))
    with pytest.raises(RuntimeError):
        gpu_core.vv_integrate(mesh, es, p0[:-1], 1.0, 1e-4, 2)
    with pytest.raises(RuntimeError):
        gpu_core.vv_integrate(mesh, es, p0, 0.0, 1e-4, 2)
    with pytest.raises(RuntimeError):
        gpu_core.vv_integrate(mesh, es, p0, 1.0, 1e-4, -1)


@pytest.mark.parametrize("name", ["vv_ico4_stress_verlet2", "vv_ico4_stress_cell2"])
def test_repeated_trajectory_replays_identically(gpu_core, name):
    """The second request of a trajectory is captured into a CUDA graph and later ones replay it
    (tm_integrate.cu): direct launch, capture and replay must give the same bits, for both parities
    of the gradient ping-pong, and a parameter change must not reuse a stale graph."""
    gold = cases.load_golden(name)
    case, p0, mass, dt, n = cases.vv_inputs(name)
    mesh, es = cases.build(gpu_core, case)
    x0 = mesh.points().copy()
    runs = [gpu_core.vv_integrate(mesh, es, p0, mass, dt, n) for _ in range(5)]
    for x, p, st in runs[1:]:
        assert np.array_equal(x, runs[0][0]) and np.array_equal(p, runs[0][1]) and np.array_equal(st, runs[0][2])
    assert nrm(runs[0][0] - x0, gold["x"] - x0) < TOL_TRAJ
    # odd step count: the two gradient buffers change roles from call to call
    odd = [gpu_core.vv_integrate(mesh, es, p0, mass, dt, 3) for _ in range(5)]
    for x, p, st in odd[1:]:
        assert np.array_equal(x, odd[0][0]) and np.array_equal(p, odd[0][1]) and np.array_equal(st, odd[0][2])
    # a changed parameter is a different trajectory
    es.eparams.kappa_b = 2.0 * es.eparams.kappa_b
    xb, pb, stb = gpu_core.vv_integrate(mesh, es, p0, mass, dt, n)
    assert not np.array_equal(xb, runs[0][0])
    es.eparams.kappa_b = 0.5 * es.eparams.kappa_b
    xa, pa, sta = gpu_core.vv_integrate(mesh, es, p0, mass, dt, n)
    assert np.array_equal(xa, runs[0][0]) and np.array_equal(sta, runs[0][2])
    # continuing a trajectory on the device (C ABI, no re-upload) through replays
    he_to, he_next, he_face, v_he, _ = mesh.halfedge_arrays()
    rp = es.eparams.repulse_params
    kind = cabi.NLIST_CELL if rp.n_search == "cell-list" else cabi.NLIST_VERLET
    par = workload.abi_params(es)
    ends = []
    for graph_legs in (False, True):
        cm = cabi.Mesh(he_to, he_next, he_face, v_he, xyz=x0)
        nl = cabi.NList(cm, kind, rp.rlist, rp.exclusion_level)
        cm.set_momenta(p0)
        if graph_legs:
            for _ in range(n // 2):
                cm.vv_integrate_async(nl, par, mass, dt, 2)          # same request over and over: replays
        else:
            cm.vv_integrate_async(nl, par, mass, dt, 2 * (n // 2))   # one direct launch sequence
        ends.append((cm.positions(), cm.momenta()))
    assert np.array_equal(ends[0][0], ends[1][0]) and np.array_equal(ends[0][1], ends[1][1])


@pytest.mark.parametrize("name", list(cases.HMC_CASES))
def test_hmc_chain_vs_reference_hmc(gpu_core, name):
    """trimem_b200.hmc.MeshHMC (device-resident trajectory) against chains of the reference's own HMC
    class run on the reference build with the same numpy seed (tests/golden/make_golden.py
    hmc_chain_cases): the same states, the same accept/reject decisions, the same temperature schedule."""
    from trimem_b200 import hmc
    gold = cases.load_golden(name)
    case, seed, opts, nsteps = cases.HMC_CASES[name]
    mesh, es = cases.build(gpu_core, case)
    x0 = mesh.points().copy()
    np.random.seed(seed)
    chain = hmc.MeshHMC(mesh, es, options=dict(opts), verbose=False)
    prev = x0
    for k in range(nsteps):
        mine_before = chain.x.copy()
        chain.step()
        assert chain.acc == int(gold["acc"][k])
        want = gold["xs"][k]
        moved = np.linalg.norm(want - prev)
        if moved == 0.0:
            assert np.array_equal(chain.x, mine_before)           # rejected: the state stays
        else:
            assert np.linalg.norm((chain.x - prev) - (want - prev)) < 1e-8 * moved
        assert np.array_equal(mesh.points(), chain.x)             # MeshHMC.step updates the mesh
        prev = want
    assert abs(chain.T - float(gold["temperature"][0])) <= 1e-15 * chain.T
    assert chain.counter["move"] == nsteps


def test_rebuild_trigger_measured_on_device(gpu_core):
    """the trajectory measures, on the device, how far its end point is from the positions the verlet
    list was built from; the value equals the host-side computation and tm_nlist_displacement's."""
    m = gpu_core
    mesh, es = cases.build(m, "ico4_physical_cell2")
    es.eparams.repulse_params.n_search = "verlet-list"
    es.update_repulsion(mesh)
    x0 = mesh.points().copy()
    rp = es.eparams.repulse_params
    p0 = np.random.default_rng(5).normal(size=x0.shape)
    xn, pn, st = m.vv_integrate(mesh, es, p0, 1.0, 2.0e-5, 6)
    disp, margin = m.vv_displacement(mesh, es)
    want = np.sqrt(((xn - x0) ** 2).sum(axis=1)).max()
    assert abs(disp - want) <= 1e-12 * want and want > 0.0
    assert margin == rp.rlist - rp.lc1
    # a second trajectory from the same start, list rebuilt in between: same end point, same displacement
    es.update_repulsion(mesh)
    xn2, _, _ = m.vv_integrate(mesh, es, p0, 1.0, 2.0e-5, 6)
    assert np.array_equal(xn, xn2) and m.vv_displacement(mesh, es)[0] == disp
    # cell lists keep no build positions: the query says so instead of returning something stale
    es.eparams.repulse_params.n_search = "cell-list"
    es.update_repulsion(mesh)
    m.vv_integrate(mesh, es, p0, 1.0, 2.0e-5, 2)
    with pytest.raises(RuntimeError, match="did not measure"):
        m.vv_displacement(mesh, es)


def test_lazy_repulsion_chain_equals_rebuild_every_move(gpu_core):
    """A chain that rebuilds its verlet list only when the device-side trigger fires visits the same states, with the
    same Hamiltonians, as the chain that rebuilds after every move (the reference's cadence, mc/evaluators.py:181-182):
    a list inside its skin margin is complete. The trigger does fire once the state has moved far enough."""
    from trimem_b200 import hmc
    m = gpu_core
    opts = {"time_step": 2.0e-5, "num_integration_steps": 6, "initial_temperature": 50.0, "info_step": 0}
    runs = []
    for lazy in (False, True):
        mesh, es = cases.build(m, "ico4_physical_cell2")
        es.eparams.repulse_params.n_search = "verlet-list"
        es.update_repulsion(mesh)
        chain = hmc.MeshHMC(mesh, es, options=dict(opts), verbose=False, rng=np.random.RandomState(31))
        upd = hmc.LazyRepulsion(chain, fraction=0.5)
        states, hams = [], []
        for k in range(24):
            chain.step()
            if lazy:
                upd()
            else:
                es.update_repulsion(mesh)
            states.append(chain.x.copy())
            hams.append(np.array(chain.last[2]))
        runs.append((np.array(states), np.array(hams), upd))
    (xs0, h0, _), (xs1, h1, upd) = runs
    assert upd.skipped > 0 and upd.rebuilds > 0, (upd.skipped, upd.rebuilds)       # both branches exercised
    # same lists up to entries outside the cut-off: same energies; the row ORDER differs between a fresh and an older
    # list, so sums agree to rounding, not bitwise
    assert np.abs(h1 - h0).max() <= 1e-11 * np.abs(h0).max()
    assert np.linalg.norm(xs1 - xs0) <= 1e-9 * np.linalg.norm(xs0 - xs0[0] + 1e-300) + 1e-12


def philox4x32_10_numpy(c, k):
    """Philox4x32-10 (Salmon et al., SC'11) on numpy arrays: c (N, 4) uint32 counters, k (2,) uint32 key."""
    c = c.astype(np.uint64).copy()
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[:, 0], M1 * c[:, 2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & MASK, (k1 + np.uint64(0xBB67AE85)) & MASK
    return c.astype(np.uint32)


def test_device_momenta_are_the_documented_generator(gpu_core):
    """tm_mesh_random_momenta: Philox4x32-10 keyed by the seed, counter (vertex, half, sequence), 53-bit uniforms,
    Box-Muller -- restated in numpy and compared value for value; then the moments and a Kolmogorov-Smirnov test."""
    from scipy import stats
    from trimem_b200 import cabi
    import halfedge
    v, f = cases.mesh_of("ico6_stress_verlet2")
    V = len(v)
    he_to, he_next, he_face, v_he, _ = halfedge.build(V, f)
    cm = cabi.Mesh(he_to, he_next, he_face, v_he, xyz=v)
    # Philox known-answer (Random123 kat_vectors): counter = key = 0 and all ones
    kat0 = philox4x32_10_numpy(np.zeros((1, 4), np.uint32), np.zeros(2, np.uint32))[0]
    assert [int(x) for x in kat0] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    kat1 = philox4x32_10_numpy(np.full((1, 4), 0xffffffff, np.uint32), np.full(2, 0xffffffff, np.uint32))[0]
    assert [int(x) for x in kat1] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    seed, seq, scale = 0x1234567890ABCDEF, 7, 2.5
    cm.random_momenta(seed, seq, scale)
    p = cm.momenta()
    key = np.array([seed & 0xFFFFFFFF, seed >> 32], np.uint32)
    z = np.empty((V, 4))
    for half in (0, 1):
        c = np.stack([np.arange(V, dtype=np.uint32), np.full(V, half, np.uint32), np.full(V, seq, np.uint32),
                      np.zeros(V, np.uint32)], axis=1)
        w = philox4x32_10_numpy(c, key).astype(np.float64)
        u = lambda hi, lo: (np.floor(hi / 32.0) * 67108864.0 + np.floor(lo / 64.0) + 0.5) / 9007199254740992.0
        r = np.sqrt(-2.0 * np.log(u(w[:, 0], w[:, 1])))
        ang = 2.0 * np.pi * u(w[:, 2], w[:, 3])
        z[:, 2 * half], z[:, 2 * half + 1] = r * np.cos(ang), r * np.sin(ang)
    assert np.abs(p - scale * z[:, :3]).max() <= 1e-12 * scale * np.abs(z).max()
    # reproducible, and a different stream for another sequence number / seed
    cm.random_momenta(seed, seq, scale)
    assert np.array_equal(cm.momenta(), p)
    cm.random_momenta(seed, seq + 1, scale)
    q = cm.momenta()
    assert abs(np.corrcoef(p.ravel(), q.ravel())[0, 1]) < 5.0 / np.sqrt(p.size)
    n = p.size
    s = p.ravel() / scale
    assert abs(s.mean()) < 5.0 / np.sqrt(n) and abs(s.var() - 1.0) < 5.0 * np.sqrt(2.0 / n)
    assert abs((s ** 4).mean() - 3.0) < 5.0 * np.sqrt(96.0 / n)
    assert stats.kstest(s, "norm").pvalue > 1e-3


def test_hmc_with_device_momenta(gpu_core):
    """MeshHMC(device_momenta=seed): reproducible from the seed, K(p0) consistent with the drawn momenta, and the same
    acceptance behaviour as the chain whose momenta numpy draws on the host."""
    from trimem_b200 import hmc
    m = gpu_core
    opts = {"time_step": 1.0e-5, "num_integration_steps": 8, "initial_temperature": 50.0, "info_step": 0}
    def run(**kw):
        mesh, es = cases.build(m, "ico4_stress_verlet2")
        chain = hmc.MeshHMC(mesh, es, options=dict(opts), verbose=False, **kw)
        hams, acc = [], 0
        for _ in range(12):
            chain.step()
            hams.append(np.array(chain.last[2]))
            acc += chain.last[1]
        return chain.x.copy(), np.array(hams), acc
    x1, h1, a1 = run(device_momenta=11, rng=np.random.RandomState(3))
    x2, h2, a2 = run(device_momenta=11, rng=np.random.RandomState(3))
    assert np.array_equal(x1, x2) and np.array_equal(h1, h2)                       # same seed, same chain
    x3, h3, a3 = run(device_momenta=12, rng=np.random.RandomState(3))
    assert not np.array_equal(x1, x3)
    xh, hh, ah = run(rng=np.random.RandomState(3))
    # kinetic energy of the draw: 3V/2 * T on average (equipartition), both ways of drawing
    V = len(x1)
    for h in (h1, h3, hh):
        assert abs(h[:, 1].mean() / (1.5 * V * 50.0) - 1.0) < 0.05
    assert abs(a1 - ah) <= 4 and abs(a3 - ah) <= 4
