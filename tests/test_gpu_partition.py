"""One mesh partitioned across ranks (tm_part_*): every rank owns a slice of the
vertices, reads halo positions from the owners' arrays and adds the ranks' reduced sums in rank
order. Here the ranks are mesh handles of ONE process on ONE device (peer pointers attached
directly); tests/test_gpu_partition_ipc.py runs the same path with one process per rank over
cudaIpc handles. Parity: sums / energy 1e-10 relative against the reference's golden values,
gradient rows 1e-10 norm-wise, trajectories against the reference's integrator as in
test_gpu_trajectory.py; all ranks must agree on the global sums bit for bit."""
import numpy as np
import pytest

import cases
from trimem_b200 import cabi, workload

pytestmark = pytest.mark.gpu
TOL = 1e-10


def nrm(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


class Ranks:
    def __init__(self, core, case, nranks, n_search=None, x=None):
        self.mesh, self.es = cases.build(core, case)
        self.v = self.mesh.points().copy() if x is None else x
        arrs = self.mesh.halfedge_arrays()[:4]
        rp = self.es.eparams.repulse_params
        kind = cabi.NLIST_CELL if rp.n_search == "cell-list" else cabi.NLIST_VERLET
        self.par = workload.abi_params(self.es)
        self.cm = [cabi.Mesh(*arrs, xyz=self.v) for _ in range(nranks)]
        for r, m in enumerate(self.cm):
            m.part_create(r, nranks)
        for r, m in enumerate(self.cm):
            for q, o in enumerate(self.cm):
                if q != r:
                    m.part_attach_local(q, o)
        self.nl = [cabi.NList(m, kind, rp.rlist, rp.exclusion_level) for m in self.cm]
        for m, nl in zip(self.cm, self.nl):
            m.part_prepare(nl)          # all local allocation before any collective call
        self.n = nranks

    def each(self, fn):
        """fn(mesh, nlist) on every rank, one host thread per rank (as one process per rank would):
        collective calls wait for their peers on the device, so the ranks' launches must not be
        serialised behind each other on the host."""
        import threading
        out, err = [None] * self.n, [None] * self.n

        def run(r):
            try:
                out[r] = fn(self.cm[r], self.nl[r])
            except BaseException as exc:      # noqa: BLE001  (re-raised in the caller's thread)
                err[r] = exc
        th = [threading.Thread(target=run, args=(r,)) for r in range(self.n)]
        [t.start() for t in th]
        [t.join() for t in th]
        for e in err:
            if e is not None:
                raise e
        return out


@pytest.mark.parametrize("case,nranks", [("ico4_stress_verlet2", 2), ("ico4_stress_verlet2", 3), ("ico4_stress_cell2", 2),
                                         ("ico4_config1", 4), ("ico4_area_stress", 2), ("ico6_stress_verlet2", 8),
                                         ("ico8_stress_verlet2", 4)])
def test_energy_gradient_partitioned(gpu_core, case, nranks):
    gold = cases.load_golden(case)
    R = Ranks(gpu_core, case, nranks)
    info = [m.part_info() for m in R.cm]
    V = R.cm[0].V
    assert sum(i["count"] for i in info) == V and info[0]["first"] == 0
    assert all(i["first"] % 32 == 0 for i in info)
    assert all(0 < i["n_ghost1"] <= i["n_pull"] < V for i in info)
    # asynchronous on every rank, then fetch: nothing may block on a peer that has not been launched
    R.each(lambda m, nl: m.energy_async(nl, R.par))
    res = R.each(lambda m, nl: m.fetch_energy(R.par))
    for e, props in res:
        assert e == res[0][0] and np.array_equal(props, res[0][1])      # identical bits on every rank
    e, props = res[0]
    for q in range(6):
        if gold["props"][q] == 0.0:
            assert props[q] == 0.0
        else:
            assert abs(props[q] - gold["props"][q]) <= TOL * abs(gold["props"][q])
    assert abs(e - float(gold["energy"])) <= TOL * abs(float(gold["energy"]))
    R.each(lambda m, nl: m.gradient_async(nl, R.par))
    gs = R.each(lambda m, nl: m.fetch_gradient())
    # each rank reports the rows it owns and zeros elsewhere; the rows are disjoint
    owned = [np.any(g != 0.0, axis=1) for g in gs]
    assert np.all(np.sum(owned, axis=0) == 1)
    g = np.sum(gs, axis=0)
    if "grad" in gold:
        assert nrm(g, gold["grad"]) < TOL
    else:
        rows = gold["rows"]
        assert nrm(g[rows], gold["grad_rows"]) < TOL
        assert abs(np.linalg.norm(g) - float(gold["grad_norm"])) <= TOL * float(gold["grad_norm"])


@pytest.mark.parametrize("name,nranks", [("vv_ico4_stress_verlet2", 2), ("vv_ico4_stress_verlet2", 3),
                                         ("vv_ico4_stress_cell2", 2), ("vv_ico4_config1", 4)])
def test_trajectory_partitioned(gpu_core, name, nranks):
    gold = cases.load_golden(name)
    case, p0, mass, dt, n = cases.vv_inputs(name)
    R = Ranks(gpu_core, case, nranks)
    R.each(lambda m, nl: m.set_momenta(p0))
    R.each(lambda m, nl: m.vv_integrate_async(nl, R.par, mass, dt, n, with_energies=True))
    stats = R.each(lambda m, nl: m.vv_stats(R.par)[0])
    ps = R.each(lambda m, nl: m.momenta())
    xs = R.each(lambda m, nl: m.positions())            # collective: whole arrays become current
    for x in xs:
        assert np.array_equal(x, xs[0])
    assert nrm(xs[0] - R.v, gold["x"] - R.v) < 1e-9
    assert nrm(np.sum(ps, axis=0), gold["p"]) < 1e-9
    st = np.array(stats)
    assert np.all(st[:, 0] == st[0, 0]) and np.all(st[:, 2] == st[0, 2])
    k0, k1 = st[:, 1].sum(), st[:, 3].sum()             # kinetic energy: each rank's owned share
    want = gold["stats"]
    assert abs(st[0, 0] - want[0]) <= TOL * abs(want[0]) and abs(k0 - want[1]) <= TOL * abs(want[1])
    assert abs(st[0, 2] - want[2]) <= 10 * TOL * abs(want[2]) and abs(k1 - want[3]) <= 10 * TOL * abs(want[3])


def test_trajectory_partitioned_with_split_pair_sweep(gpu_core):
    """655,362 vertices on 2 ranks: every rank's pair sweep is cut across its side streams, its ring gradient follows range
    by range, and the repulsion sum is reduced and exchanged in a second step (tm_energy.cu: join_side_streams). Against
    the same trajectory of the whole mesh on one handle: the ranks add their sums in rank order, so equal to rounding."""
    case = "ico8_stress_verlet2"
    R = Ranks(gpu_core, case, 2)
    p0 = np.random.default_rng(4).standard_normal(R.v.shape)
    mass, dt, n = 1.0, 2.0e-8, 3
    R.each(lambda m, nl: m.set_momenta(p0))
    R.each(lambda m, nl: m.vv_integrate_async(nl, R.par, mass, dt, n, with_energies=True))
    stats = np.array(R.each(lambda m, nl: m.vv_stats(R.par)[0]))
    ps = np.sum(R.each(lambda m, nl: m.momenta()), axis=0)
    xs = R.each(lambda m, nl: m.positions())
    assert np.array_equal(xs[0], xs[1])
    arrs = R.mesh.halfedge_arrays()[:4]
    rp = R.es.eparams.repulse_params
    whole = cabi.Mesh(*arrs, xyz=R.v)
    nl = cabi.NList(whole, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
    x1, p1, st1 = whole.vv_integrate(nl, R.par, p0, mass, dt, n)
    assert nrm(xs[0] - R.v, x1 - R.v) < 1e-11 and nrm(ps, p1) < 1e-11
    assert np.all(stats[:, 0] == stats[0, 0]) and np.all(stats[:, 2] == stats[0, 2])      # identical bits on every rank
    assert abs(stats[0, 0] - st1[0]) <= 1e-13 * abs(st1[0]) and abs(stats[0, 2] - st1[2]) <= 1e-13 * abs(st1[2])
    assert abs(stats[:, 1].sum() - st1[1]) <= 1e-12 * abs(st1[1]) and abs(stats[:, 3].sum() - st1[3]) <= 1e-12 * abs(st1[3])


def test_rebuild_after_trajectory_and_allsum(gpu_core):
    """A neighbour list rebuilt after a partitioned trajectory sees every rank's current positions
    (tm_nlist_create pulls the far field first), and equals the list a single device builds."""
    name = "vv_ico4_stress_verlet2"
    case, p0, mass, dt, n = cases.vv_inputs(name)
    R = Ranks(gpu_core, case, 2)
    # the sums path itself: rank-ordered addition, same bits everywhere
    vec = [np.arange(6) * 0.1 + r for r in range(2)]
    res = R.each(lambda m, nl: m.part_allsum6(vec[m.part_info()["rank"]]))
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], vec[0] + vec[1])
    R.each(lambda m, nl: m.set_momenta(p0))
    R.each(lambda m, nl: m.vv_integrate_async(nl, R.par, mass, dt, n, with_energies=False))
    rp = R.es.eparams.repulse_params
    new = R.each(lambda m, nl: cabi.NList(m, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level))
    # single-device list at the same end point
    gold = cases.load_golden(name)
    x_end = R.cm[0].positions()
    arrs = R.mesh.halfedge_arrays()[:4]
    one = cabi.Mesh(*arrs, xyz=x_end)
    nl1 = cabi.NList(one, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
    d1 = nl1.distance_matrix(rp.rlist)
    # every rank builds the rows of the vertices it owns; together they are the single-device list
    parts = [new[r].distance_matrix(rp.rlist) for r in range(2)]
    key1 = d1[1].astype(np.int64) * one.V + d1[2]
    keyp = np.concatenate([p[1].astype(np.int64) * one.V + p[2] for p in parts])
    assert len(keyp) == len(key1) and np.array_equal(np.sort(keyp), np.sort(key1))
    order = np.argsort(keyp)
    assert np.array_equal(np.concatenate([p[0] for p in parts])[order], d1[0][np.argsort(key1)])
    assert sum(new[r].info()[0] for r in range(2)) == nl1.info()[0]
    assert nrm(x_end - R.v, gold["x"] - R.v) < 1e-9


def test_partition_errors(gpu_core):
    mesh, es = cases.build(gpu_core, "ico3_edge")
    arrs = mesh.halfedge_arrays()[:4]
    cm = cabi.Mesh(*arrs, xyz=mesh.points().copy())
    with pytest.raises(cabi.TMError):
        cm.part_create(2, 2)
    cm.part_create(0, 2)
    with pytest.raises(cabi.TMError):
        cm.part_create(0, 2)
    with pytest.raises(cabi.TMError, match="not every peer is attached"):
        cm.energy_async(None, workload.abi_params(es))


@pytest.mark.parametrize("gen,nranks", [("tube", 2), ("tube", 3), ("rect", 2), ("fan", 2)])
def test_partitioned_open_and_irregular_meshes(gpu_core, port, gen, nranks):
    """Boundary vertices (ring flags, open fans) and a valence-12 vertex (generic ring walk) across the cuts:
    the partitioned sums and gradient rows equal the unpartitioned ones and the oracle port's."""
    v, f = cases.open_mesh(gen)
    mesh = gpu_core.TriMesh(v, f)
    ep = cases.open_mesh_params(gpu_core, mesh)
    a, l = gpu_core.avg_tri_props(mesh)
    es = gpu_core.EnergyManager(mesh, ep)
    e1, g1 = es.energy(mesh), es.gradient(mesh)
    pm = port.TriMesh(v, f)
    pe = port.EnergyManager(pm, port.params_from(ep))
    assert abs(e1 - pe.energy(pm)) <= TOL * abs(e1)
    # Gradient near a boundary: the reference's dihedral_angle is 0 on boundary edges while its
    # dihedral_angle_grad is still evaluated there, on the fictitious sector of the boundary loop (SURVEY
    # App. C #11): the reference gradient is not the gradient of its energy on open meshes. Parity means
    # reproducing exactly that (k_grad: boundary_spoke_cot).
    assert nrm(g1, pe.gradient(pm)) < 1e-9
    gold = cases.load_golden("open_meshes")          # the reference build's own values
    assert abs(e1 - float(gold[f"{gen}_energy"])) <= TOL * abs(e1)
    assert nrm(g1, gold[f"{gen}_grad"]) < 1e-9
    for q in range(6):
        assert abs(cases.props6(es.properties(mesh))[q] - gold[f"{gen}_props"][q]) <= TOL * abs(gold[f"{gen}_props"][q]) + 1e-300
    par = workload.abi_params(es)
    arrs = mesh.halfedge_arrays()[:4]
    cm = [cabi.Mesh(*arrs, xyz=v) for _ in range(nranks)]
    for r, m in enumerate(cm):
        m.part_create(r, nranks)
    for r, m in enumerate(cm):
        for q, o in enumerate(cm):
            if q != r:
                m.part_attach_local(q, o)
    rp = ep.repulse_params
    nls = [cabi.NList(m, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level) for m in cm]
    for m, nl in zip(cm, nls):
        m.part_prepare(nl)
    import threading
    out = [None] * nranks

    def run(r):
        out[r] = (cm[r].energy(nls[r], par)[0], cm[r].gradient(nls[r], par))
    th = [threading.Thread(target=run, args=(r,)) for r in range(nranks)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert all(o[0] == out[0][0] for o in out)
    assert abs(out[0][0] - e1) <= 1e-13 * abs(e1)
    g = np.sum([o[1] for o in out], axis=0)
    assert nrm(g, g1) < 1e-13


def test_replicated_flip_sweep_on_partitioned_mesh(gpu_core):
    """Flip sweeps are not partitioned: every rank runs the same sweep over the whole (replicated) mesh after
    bringing all positions up to date. Same proposals -> the same flips on every rank and the flips of an
    unpartitioned sweep; the next partitioned evaluation rebuilds its halo lists for the new connectivity."""
    name = "flip_ico4_cell"
    gold = cases.load_golden(name)
    case, edges, u = cases.flip_inputs(name)
    R = Ranks(gpu_core, case, 2)
    # move the vertices on the device first, so that the far field of each rank is stale before the sweep
    p0 = 0.05 * np.random.default_rng(1).normal(size=R.v.shape)
    R.each(lambda m, nl: m.set_momenta(p0))
    R.each(lambda m, nl: m.vv_integrate_async(nl, R.par, 1.0, 1.0e-5, 3, with_energies=False))
    props0 = R.each(lambda m, nl: m.properties(nl, R.par))
    assert np.array_equal(props0[0], props0[1])
    out = R.each(lambda m, nl: m.flip_sweep(nl, R.par, cabi.FLIP_SERIAL, edges, u, props0[0]))
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    topo = R.each(lambda m, nl: m.topology())
    for a, b in zip(topo[0], topo[1]):
        assert np.array_equal(a, b)
    # the same sweep on one device from the same positions
    x = R.cm[0].positions()
    one = cabi.Mesh(*R.mesh.halfedge_arrays()[:4], xyz=x)
    rp = R.es.eparams.repulse_params
    nl1 = cabi.NList(one, cabi.NLIST_CELL, rp.rlist, rp.exclusion_level)
    nacc1, flips1, props1 = one.flip_sweep(nl1, R.par, cabi.FLIP_SERIAL, edges, u, one.properties(nl1, R.par))
    assert nacc1 == out[0][0] and np.array_equal(flips1, out[0][1])
    # evaluation after the sweep: halo lists follow the new connectivity (prepared locally first: with the
    # ranks sharing one device, a rank still rebuilding its tables must not meet a peer that already waits)
    for m, nl in zip(R.cm, R.nl):
        m.part_prepare(nl)
    e = R.each(lambda m, nl: m.energy(nl, R.par)[0])
    assert e[0] == e[1] and abs(e[0] - one.energy(nl1, R.par)[0]) <= 1e-12 * abs(e[0])


def test_open_mesh_area_bond(gpu_core):
    gold = cases.load_golden("open_meshes")
    v, f = cases.open_mesh("rect")
    mesh = gpu_core.TriMesh(v, f)
    es = gpu_core.EnergyManager(mesh, cases.open_mesh_params(gpu_core, mesh, bond="Area"))
    e, g = es.energy(mesh), es.gradient(mesh)
    assert abs(e - float(gold["rect_area_energy"])) <= TOL * abs(e)
    assert nrm(g, gold["rect_area_grad"]) < 1e-9


def test_partition_through_the_drop_in_classes(gpu_core):
    """core.TriMesh.part_* / EnergyManager.part_prepare: two ranks as two TriMesh / EnergyManager pairs of one
    process (one host thread per rank; the bindings release the interpreter lock around device work), evaluated
    through the same calls a trimem user makes, plus one MeshHMC move; against the unpartitioned results."""
    from trimem_b200 import hmc
    m, n = gpu_core, 2
    name = "vv_ico4_stress_verlet2"
    case, p0, mass, dt, nsteps = cases.vv_inputs(name)
    mesh1, es1 = cases.build(m, case)
    e1, g1 = es1.energy(mesh1), es1.gradient(mesh1)
    x1, pn1, st1 = m.vv_integrate(mesh1, es1, p0, mass, dt, nsteps)
    pairs = [cases.build(m, case) for _ in range(n)]
    for r, (mesh, es) in enumerate(pairs):
        mesh.part_create(r, n)
    for r, (mesh, es) in enumerate(pairs):
        for q, (other, _) in enumerate(pairs):
            if q != r:
                mesh.part_attach_local(q, other)
    for mesh, es in pairs:
        es.part_prepare(mesh)
    info = [mesh.part_info() for mesh, _ in pairs]
    assert sum(i["count"] for i in info) == mesh1.n_vertices() and info[1]["rank"] == 1

    def each(fn):
        import threading
        out, err = [None] * n, [None] * n

        def run(r):
            try:
                out[r] = fn(*pairs[r])
            except BaseException as exc:      # noqa: BLE001
                err[r] = exc
        th = [threading.Thread(target=run, args=(r,)) for r in range(n)]
        [t.start() for t in th]
        [t.join() for t in th]
        for e in err:
            if e is not None:
                raise e
        return out
    res = each(lambda mesh, es: (es.energy(mesh), es.gradient(mesh)))
    assert res[0][0] == res[1][0] and abs(res[0][0] - e1) <= 1e-13 * abs(e1)
    assert nrm(res[0][1] + res[1][1], g1) < 1e-13            # every rank delivers the rows it owns, zeros elsewhere
    traj = each(lambda mesh, es: m.vv_integrate(mesh, es, p0, mass, dt, nsteps))
    assert np.array_equal(traj[0][0], traj[1][0]) and nrm(traj[0][0] - mesh1.points(), x1 - mesh1.points()) < 1e-12
    assert nrm(traj[0][1] + traj[1][1], pn1) < 1e-12
    # potential energies are global sums (same on every rank); a rank's kinetic energy covers the rows it owns
    assert abs(traj[0][2][2] - st1[2]) <= 1e-12 * abs(st1[2]) and traj[0][2][2] == traj[1][2][2]
    assert abs(traj[0][2][3] + traj[1][2][3] - st1[3]) <= 1e-12 * abs(st1[3])
    # collective list rebuild through the manager
    each(lambda mesh, es: es.update_repulsion(mesh))
    res2 = each(lambda mesh, es: es.energy(mesh))
    assert res2[0] == res2[1] and abs(res2[0] - e1) <= 1e-13 * abs(e1)
