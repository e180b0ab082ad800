"""Neighbour lists on the GPU: pair sets must be bit-exact (reference tests/test_neighbours.py)."""
import numpy as np
import pytest
from scipy.sparse import coo_matrix
from scipy.spatial import KDTree

import cases
from trimem_b200 import meshgen, workload

pytestmark = pytest.mark.gpu


def get_nlist(m, mesh, ltype, rlist, excl):
    params = m.EnergyParams()
    params.repulse_params.n_search = ltype
    params.repulse_params.rlist = rlist
    params.repulse_params.exclusion_level = excl
    return m.make_nlist(mesh, params)


@pytest.mark.parametrize("ltype", ["cell-list", "verlet-list"])
@pytest.mark.parametrize("rlist", [0.2, 0.5, 1.2])
def test_distance_matrix(gpu_core, ltype, rlist):
    v, f = meshgen.rect(20, 20)
    mesh = gpu_core.TriMesh(v, f)
    x = mesh.points()
    nl = get_nlist(gpu_core, mesh, ltype, rlist, 0)
    d, i, j = nl.distance_matrix(mesh, 0.123)
    assert isinstance(d, list) and isinstance(i, list)
    M = coo_matrix((d, (i, j)), shape=(len(x), len(x)))
    if ltype == "cell-list":
        M = M + M.T
    tree = KDTree(x)
    C = tree.sparse_distance_matrix(tree, 0.123)
    assert np.allclose(C.todense(), M.todense())
    # bit-exact against the reference's own output
    gold = cases.load_golden("neighbours_rect20")
    tag = ltype.split("-")[0]
    o = np.lexsort((j, i))
    assert np.array_equal(np.asarray(i)[o], gold[f"{tag}_dm_{rlist}_i"])
    assert np.array_equal(np.asarray(j)[o], gold[f"{tag}_dm_{rlist}_j"])
    assert np.array_equal(np.asarray(d)[o], gold[f"{tag}_dm_{rlist}_d"])


@pytest.mark.parametrize("ltype", ["cell-list", "verlet-list"])
@pytest.mark.parametrize("excl,nn", [(0, 44), (1, 38), (2, 26)])
def test_exclusion(gpu_core, ltype, excl, nn):
    v, f = meshgen.rect(20, 20)
    mesh = gpu_core.TriMesh(v, f)
    nl = get_nlist(gpu_core, mesh, ltype, 0.2, excl)
    dist, jdx = nl.point_distances(mesh, 151, 0.2)
    assert len(jdx) == nn and len(dist) == nn
    gold = cases.load_golden("neighbours_rect20")
    assert np.array_equal(np.sort(jdx), gold[f"{ltype.split('-')[0]}_pd151_excl{excl}"])
    # distance components are x_pid - x_j
    x = mesh.points()
    for p, j in zip(dist, jdx):
        assert np.array_equal(np.array([p[0], p[1], p[2]]), x[151] - x[j])


@pytest.mark.parametrize("case", ["ico4_stress_cell0", "ico4_stress_cell1", "ico4_stress_cell2",
                                  "ico4_stress_verlet0", "ico4_stress_verlet1", "ico4_stress_verlet2",
                                  "ico6_stress_verlet2", "ico8_stress_verlet2"])
def test_pair_sets_exact(gpu_core, case):
    gold = cases.load_golden(case)
    mesh, es = cases.build(gpu_core, case)
    V = mesh.n_vertices()
    nl = gpu_core.make_nlist(mesh, es.eparams)
    d, i, j = nl.distance_matrix(mesh, es.eparams.repulse_params.rlist)
    i, j = np.asarray(i, np.int64), np.asarray(j, np.int64)
    key = np.unique(np.minimum(i, j) * V + np.maximum(i, j))
    assert len(key) == int(gold["n_pairs_undirected"])
    assert int(key.sum() % (2 ** 61 - 1)) == int(gold["pair_key_sum"][0])
    if "pair_keys" in gold:
        assert np.array_equal(key, gold["pair_keys"])
    if "verlet" in case:
        assert len(i) <= 2 * len(key)            # strict < at query vs <= at build


def test_stale_cells_live_exclusions_after_motion(gpu_core, port):
    """Cell lists keep build-time occupancy but use live positions/topology."""
    v, f = meshgen.jittered_icosphere(3, seed=9)
    mesh = gpu_core.TriMesh(v, f)
    ep = workload.energy_params(gpu_core, mesh, repulsion="stress", n_search="cell-list", exclusion=1)
    es = gpu_core.EnergyManager(mesh, ep)
    pm = port.TriMesh(v, f)
    pe = port.EnergyManager(pm, port.params_from(ep))
    rng = np.random.default_rng(2)
    dx = 0.05 * rng.normal(size=v.shape)
    np.copyto(mesh.points(), v + dx)
    pm.x[:] = v + dx
    got, want = cases.props6(es.properties(mesh)), pe.properties(pm)
    assert abs(got[5] - want[5]) / abs(want[5]) < 1e-10
    nl = gpu_core.make_nlist(mesh, ep)            # rebuilt at the new positions
    _, jd = nl.point_distances(mesh, 17, ep.repulse_params.lc1)
    pnl = port.NeighbourList(pm, "cell-list", ep.repulse_params.rlist, 1)
    _, jp = pnl.point_distances(pm, 17, ep.repulse_params.lc1)
    assert sorted(jd) == sorted(jp.tolist())
    # the same with exclusion level 2 (sorted two-ring rows in the sweep), energy and gradient
    ep2 = workload.energy_params(gpu_core, mesh, repulsion="stress", n_search="cell-list", exclusion=2)
    np.copyto(mesh.points(), v)
    pm.x[:] = v
    es2 = gpu_core.EnergyManager(mesh, ep2)
    pe2 = port.EnergyManager(pm, port.params_from(ep2))
    np.copyto(mesh.points(), v + dx)
    pm.x[:] = v + dx
    assert abs(es2.energy(mesh) - pe2.energy(pm)) <= 1e-10 * abs(pe2.energy(pm))
    g, gp = es2.gradient(mesh), pe2.gradient(pm)
    assert np.linalg.norm(g - gp) <= 1e-10 * np.linalg.norm(gp)


def test_errors(gpu_core):
    v, f = meshgen.rect(5, 5)
    mesh = gpu_core.TriMesh(v, f)
    with pytest.raises(RuntimeError, match="Unsupported exclusion level"):
        get_nlist(gpu_core, mesh, "verlet-list", 0.2, 3)
    with pytest.raises(RuntimeError, match="Unknown neighbour search algorithm."):
        get_nlist(gpu_core, mesh, "ball-tree", 0.2, 0)


def test_skin_criterion(gpu_core):
    """tm_nlist_displacement / EnergyManager.repulsion_outdated: the frozen verlet rows are complete for the
    cut-off lc1 exactly as long as no vertex moved more than (rlist - lc1) / 2 since the build."""
    mesh, es = cases.build(gpu_core, "ico4_stress_verlet2")
    rp = es.eparams.repulse_params
    skin = rp.rlist - rp.lc1
    x0 = mesh.points().copy()
    assert not es.repulsion_outdated(mesh)
    rng = np.random.default_rng(3)
    d = rng.normal(size=x0.shape)
    d /= np.linalg.norm(d, axis=1)[:, None]
    np.copyto(mesh.points(), x0 + 0.49 * skin * d)
    assert not es.repulsion_outdated(mesh)
    # while the criterion holds, the frozen rows give the repulsion sum of a freshly built list
    # (same pairs inside the cut-off; the row order, hence the rounding, may differ)
    r_frozen = es.properties(mesh).repulsion
    es2 = gpu_core.EnergyManager(mesh, es.eparams)
    assert abs(es2.properties(mesh).repulsion - r_frozen) <= 1e-12 * abs(r_frozen)
    x1 = x0 + 0.49 * skin * d
    x1[17] = x0[17] + 0.51 * skin * d[17]
    np.copyto(mesh.points(), x1)
    assert es.repulsion_outdated(mesh)
    es.update_repulsion(mesh)
    assert not es.repulsion_outdated(mesh)
    # cell lists are always reported as due
    mesh_c, es_c = cases.build(gpu_core, "ico4_stress_cell2")
    assert es_c.repulsion_outdated(mesh_c)
