"""Parity of the CUDA energy / gradient path against the reference's golden vectors, the
oracle port and (where present) the live reference build. Tolerances: 1e-10 relative on
the six reduced properties and the energy, 1e-10 norm-wise on the gradient (BASELINE.json)."""
import numpy as np
import pytest

import cases
from trimem_b200 import cabi, meshgen, workload

pytestmark = pytest.mark.gpu
TOL = 1e-10
ENERGY_CASES = [c for c in cases.CASES if not c.startswith(("ico3_flip", "ico4_flip", "ico6_flip", "rect_flip"))]


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def check_props(got, want, tol=TOL):
    for q in range(len(want)):
        if want[q] == 0.0:
            assert got[q] == 0.0, q          # a term that is exactly 0 in the reference stays exactly 0
        else:
            assert rel(got[q], want[q]) < tol, (q, got[q], want[q])


@pytest.mark.parametrize("case", ENERGY_CASES)
def test_pybind_path_vs_golden(gpu_core, case):
    gold = cases.load_golden(case)
    mesh, es = cases.build(gpu_core, case)
    check_props(cases.props6(es.initial_props), gold["initial_props"])
    check_props(cases.props6(es.properties(mesh)), gold["props"])
    assert rel(es.energy(mesh), float(gold["energy"])) < TOL
    g = es.gradient(mesh)
    # a fresh caller-owned array (backed by a recycled pinned block: base is a capsule, not the mesh)
    assert g.shape == (mesh.n_vertices(), 3) and g.flags.writeable and not np.shares_memory(g, mesh.points())
    assert rel(np.linalg.norm(g), float(gold["grad_norm"])) < TOL
    if "grad" in gold:
        assert np.linalg.norm(g - gold["grad"]) / np.linalg.norm(gold["grad"]) < TOL
    else:
        rows = gold["rows"]
        assert np.linalg.norm(g[rows] - gold["grad_rows"]) / np.linalg.norm(gold["grad_rows"]) < TOL
        assert np.linalg.norm(g.sum(axis=0) - gold["grad_sum"]) < TOL * float(gold["grad_norm"])
        # every row, through the vector sums of tiles of 64 consecutive rows of the reference's gradient: a tile's
        # sum cancels, so the yardstick is the tile's sum of absolute values (what rounding in a 64-term sum scales with)
        tiles = cases.load_golden(case + "_gtiles")
        mine = cases.tile_sums(g)
        assert mine.shape == tiles["tile_sums"].shape
        ratio = np.linalg.norm(mine - tiles["tile_sums"], axis=1) / np.linalg.norm(tiles["tile_abs"], axis=1)
        # Per tile an order looser than the norm-wise bar. Up to two tiles may stand out further: where an edge is flat
        # to ~1e-9 rad the REFERENCE's gradient row is off (its acos-based angle cannot resolve it) --
        # on the jittered 655,362-vertex sphere that is edge (113071, 420275), |theta| = 1.4e-9, row 113071 off by 7e-8
        # of its length, and central differences of the reference's own energy agree with this implementation's row
        # to 2e-8, their noise level, against 7e-8 for the reference's row.
        outliers = np.flatnonzero(ratio > 10 * TOL)
        assert len(outliers) <= 2 and ratio.max() < 1e-6, (ratio.max(), outliers[:8])


@pytest.mark.parametrize("case", ["ico4_config1", "ico4_stress_cell2", "ico4_stress_verlet2", "ico4_area_stress", "ico3_area"])
def test_c_abi_path_vs_golden_and_port(gpu_core, port, case):
    """Straight through the C ABI (ctypes), with half-edge arrays from the oracle's own builder."""
    import halfedge
    gold = cases.load_golden(case)
    v, f = cases.mesh_of(case)
    he_to, he_next, he_face, v_he, _ = halfedge.build(len(v), f)
    cm = cabi.Mesh(he_to, he_next, he_face, v_he, xyz=v)
    # parameters from the port's EnergyManager (independent of the product's host code)
    pm = port.TriMesh(v, f)
    pe = port.EnergyManager(pm, port.workload_params(pm, **cases.CASES[case][2]))
    cp = pe.cparams()
    p = cabi.EnergyParamsABI()
    for name, _ in cabi.EnergyParamsABI._fields_:
        if name != "_pad":
            setattr(p, name, getattr(cp, name))
    kind = cabi.NLIST_CELL if pe.params["n_search"] == "cell-list" else cabi.NLIST_VERLET
    nl = cabi.NList(cm, kind, pe.params["rlist"], pe.params["exclusion_level"])
    props = cm.properties(nl, p)
    check_props(props, gold["props"])
    check_props(props, pe.properties(pm))
    e, _ = cm.energy(nl, p)
    assert rel(e, float(gold["energy"])) < TOL
    g = cm.gradient(nl, p)
    assert np.linalg.norm(g - gold["grad"]) / np.linalg.norm(gold["grad"]) < TOL
    assert np.linalg.norm(g - pe.gradient(pm)) / np.linalg.norm(g) < TOL
    # per-vertex properties (vertex_properties, mesh_properties.cpp:12-49)
    vp = cm.vertex_properties(nl, p)
    scale = np.abs(gold["vprops"]).max(axis=0) + 1e-300
    assert np.max(np.abs(vp - gold["vprops"]) / scale) < TOL
    # grid geometry of the neighbour list equals the port's (bounding-box quirks included)
    n_entries, shape, box, cell = nl.info()
    pshape, pbox, pcell = pe.nlist.grid()
    assert np.array_equal(shape, pshape) and np.array_equal(box, pbox) and np.array_equal(cell, pcell)
    if kind == cabi.NLIST_VERLET:
        assert n_entries == pe.nlist.n_entries() == 2 * int(gold["n_pairs_undirected"])


def test_docs_known_answer(gpu_core):
    # reference doc/python-module.rst:60-98
    m = gpu_core
    v, f = meshgen.icosphere(3, project_each_round=False)
    v = v.astype(np.float32).astype(np.float64)
    mesh = m.TriMesh(v, f)
    ep = m.EnergyParams()
    ep.kappa_b, ep.kappa_a, ep.kappa_v, ep.kappa_c, ep.kappa_t, ep.kappa_r = 30.0, 1e6, 1e6, 1e6, 1e5, 1e3
    ep.area_frac = 0.5
    es = m.EnergyManager(mesh, ep)
    assert rel(es.energy(mesh), 1000755.3602085959) < 1e-12
    gold = cases.load_golden("analytic")
    g = es.gradient(mesh)
    assert np.linalg.norm(g - gold["docs_grad"]) / np.linalg.norm(gold["docs_grad"]) < TOL


def test_analytic_sphere_and_tube(gpu_core):
    # reference tests/test_energy.py (closed and OPEN meshes), 1e-2 vs analytic + exact reference values
    m = gpu_core
    gold = cases.load_golden("analytic")
    for tag, r in (("r1", 1.0), ("r05", 0.5)):
        v, f = meshgen.icosphere(5)
        mesh = m.TriMesh(v * r, f)
        p = cases.props6(m.EnergyManager(mesh, m.EnergyParams()).properties(mesh))
        for got, want in zip(p[:4], (4 * np.pi * r**2, 4 / 3 * np.pi * r**3, 4 * np.pi * r, 8 * np.pi)):
            assert rel(got, want) < 1e-2
        check_props(p, gold[f"sphere_{tag}"])
        v, f = meshgen.tube(1.0, r, 32)
        mesh = m.TriMesh(v, f)
        p = cases.props6(m.EnergyManager(mesh, m.EnergyParams()).properties(mesh))
        for got, want in zip(p[:4], (2 * np.pi * r, 2 / 3 * np.pi * r**2, np.pi, np.pi / r)):
            assert rel(got, want) < 1e-2
        # The tube's quad diagonals are exactly flat edges. The reference takes acos of a cosine that
        # rounds to 1 - 1ulp there and returns +1.5e-8 rad instead of 0 (always with a + sign since
        # its sign test is `>= 0`), a systematic 5e-8 relative excess in curvature (ours is pi to the
        # last digit, the analytic value). This is the reference's own noise floor on flat edges
        # ("dihedral angle"); area and volume still match to 1e-10.
        check_props(p[:2], gold[f"tube_{tag}"][:2])
        for q in (2, 3):
            assert rel(p[q], gold[f"tube_{tag}"][q]) < 1e-6


@pytest.mark.parametrize("bond", ["Edge", "Area"])
def test_gradient_vs_finite_differences(gpu_core, bond):
    # reference tests/test_gradient.py: forward FD with eps=1e-8 through core.gradient(...)
    m = gpu_core
    v, f = meshgen.icosphere(3)
    mesh = m.TriMesh(v, f)
    ep = workload.energy_params(m, mesh, kappa=cases.K1, fracs=(0.8, 0.8, 0.8), bond=bond, n_search="cell-list")
    es = m.EnergyManager(mesh, ep)
    ref_grad = np.empty((mesh.n_vertices(), 3))
    m.gradient(mesh, es, ref_grad, 1.0e-8)
    grad = es.gradient(mesh)
    assert np.linalg.norm(grad - ref_grad) / np.linalg.norm(ref_grad) < 1.0e-4
    # perturbations undone the way the reference undoes them (`+= eps; -= eps`, bindings.cpp:40-47):
    # back to within an ulp of the coordinate, not bitwise
    assert np.allclose(mesh.points(), v, rtol=0, atol=1e-15)


def test_bit_reproducible(gpu_core):
    mesh, es = cases.build(gpu_core, "ico6_stress_verlet2")
    e = [es.energy(mesh) for _ in range(3)]
    g = [es.gradient(mesh) for _ in range(3)]
    assert e[0] == e[1] == e[2]
    assert np.array_equal(g[0], g[1]) and np.array_equal(g[0], g[2])


def test_positions_are_reuploaded_every_call(gpu_core):
    mesh, es = cases.build(gpu_core, "ico4_config1")
    e0 = es.energy(mesh)
    x = mesh.points()
    x *= 1.01                                   # Python writes through the view behind our back
    e1 = es.energy(mesh)
    assert e1 != e0
    x /= 1.01
    assert rel(es.energy(mesh), e0) < 1e-9


def test_eparams_are_live_and_continuation(gpu_core):
    m = gpu_core
    mesh, es = cases.build(m, "ico3_edge")
    p0 = es.properties(mesh)
    e0 = es.energy(mesh)
    assert rel(es.energy(p0), e0) < 1e-14         # energy(vprops) overload
    es.eparams.kappa_b = 2.0 * es.eparams.kappa_b   # nested fields of the read-only attribute are live
    e1 = es.energy(mesh)
    assert rel(e1 - e0, p0.bending * 1.0) < 1e-9
    with pytest.raises(AttributeError):
        es.eparams = m.EnergyParams()
    # update_reference_properties: lambda += delta while lambda < 1 (energy.cpp:47-55)
    es.eparams.continuation_params.lam = 0.0
    es.eparams.continuation_params.delta = 0.25
    e_l0 = es.energy(mesh)
    es.update_reference_properties()
    assert es.eparams.continuation_params.lam == 0.25
    assert es.energy(mesh) != e_l0
    for _ in range(10):
        es.update_reference_properties()
    assert es.eparams.continuation_params.lam == 1.0


def test_update_repulsion_and_print_info(gpu_core, capsys):
    mesh, es = cases.build(gpu_core, "ico4_stress_verlet2")
    e0 = es.energy(mesh)
    x = mesh.points()
    x *= 1.2                                      # all pair distances grow: frozen rows -> fewer within lc1
    e_frozen = es.energy(mesh)
    es.update_repulsion(mesh)
    e_fresh = es.energy(mesh)
    assert e_frozen != e0 and np.isfinite(e_fresh)
    es.print_info(mesh)
    out = capsys.readouterr().out
    assert "----- EnergyManager info" in out and "total:" in out and "repulsion:" in out


def test_setup_reductions(gpu_core, port):
    # area / edges_length / avg_tri_props (src/util.cpp:9-41)
    v, f = meshgen.jittered_icosphere(4)
    mesh = gpu_core.TriMesh(v, f)
    pm = port.TriMesh(v, f)
    a, l = gpu_core.avg_tri_props(mesh)
    pa, pl = port.avg_tri_props(pm)
    assert rel(a, pa) < 1e-12 and rel(l, pl) < 1e-12
    assert rel(gpu_core.area(mesh), pa * len(f)) < 1e-12
    assert rel(gpu_core.edges_length(mesh), pl * mesh.n_edges()) < 1e-12


def test_live_reference_build(gpu_core, refmod):
    """Where the reference build travelled with the tree, compare live on a fresh random case."""
    if refmod is None:
        pytest.skip("oracle/_ref not present on this machine (goldens cover it)")
    v, f = meshgen.jittered_icosphere(5, seed=123)
    for ns, ex in (("cell-list", 1), ("verlet-list", 2)):
        ma, mb = refmod.TriMesh(v, f), gpu_core.TriMesh(v, f)
        kw = dict(repulsion="stress", n_search=ns, exclusion=ex)
        ea = refmod.EnergyManager(ma, workload.energy_params(refmod, ma, **kw))
        eb = gpu_core.EnergyManager(mb, workload.energy_params(gpu_core, mb, **kw))
        check_props(cases.props6(eb.properties(mb)), cases.props6(ea.properties(ma)))
        assert rel(eb.energy(mb), ea.energy(ma)) < TOL
        ga, gb = ea.gradient(ma), eb.gradient(mb)
        assert np.linalg.norm(ga - gb) / np.linalg.norm(ga) < TOL


def test_full_size_properties(gpu_core):
    """Size-independent properties at the headline size (655,362 vertices)."""
    mesh, es = cases.build(gpu_core, "ico8_norep")
    V = mesh.n_vertices()
    g = es.gradient(mesh)
    # translation invariance of every term => the gradient sums to zero
    assert np.linalg.norm(g.sum(axis=0)) < 1e-9 * np.linalg.norm(g) * np.sqrt(V)
    # rotation invariance => zero total torque
    tq = np.cross(mesh.points(), g).sum(axis=0)
    assert np.linalg.norm(tq) < 1e-9 * np.linalg.norm(g) * np.sqrt(V)
    # Euler characteristic from the device ring table: sum of valences = 2E = 6V - 12
    assert mesh.n_edges() == 3 * V - 6
    # directional derivative along a random direction vs central differences of the energy
    rng = np.random.default_rng(5)
    d = rng.normal(size=g.shape)
    d /= np.linalg.norm(d)
    x0 = mesh.points().copy()
    h = 1e-7
    np.copyto(mesh.points(), x0 + h * d); ep = es.energy(mesh)
    np.copyto(mesh.points(), x0 - h * d); em = es.energy(mesh)
    np.copyto(mesh.points(), x0)
    assert rel((ep - em) / (2 * h), float((g * d).sum())) < 1e-5


def test_minimiser_path_lbfgs(gpu_core):
    """The minimisation driver of the reference (mc/util.py:240-306) runs scipy's L-BFGS-B over
    EnergyEvaluators.fun / .grad with the neighbour list refreshed every iteration: the same calls here
    (positions written through points(), energy(), gradient(), update_repulsion from the callback)."""
    from scipy.optimize import minimize
    mesh, es = cases.build(gpu_core, "ico4_stress_verlet2")
    x0 = mesh.points().copy()

    def fun(x):
        np.copyto(mesh.points(), x.reshape(x0.shape))
        return es.energy(mesh)

    def grad(x):
        np.copyto(mesh.points(), x.reshape(x0.shape))
        return es.gradient(mesh).ravel()

    calls = []

    def callback(x):
        np.copyto(mesh.points(), x.reshape(x0.shape))
        es.update_repulsion(mesh)
        es.update_reference_properties()
        calls.append(es.energy(mesh))

    e0 = fun(x0.ravel())
    res = minimize(fun, x0.ravel(), jac=grad, callback=callback, method="L-BFGS-B", options={"maxiter": 15})
    assert len(calls) >= 5 and res.fun < 0.5 * e0
    # the gradient the minimiser used is the gradient of the energy it used (directional finite difference)
    x = res.x.copy()
    g = grad(x)
    d = np.random.default_rng(0).normal(size=x.shape)
    d /= np.linalg.norm(d)
    h = 1e-6
    fd = (fun(x + h * d) - fun(x - h * d)) / (2 * h)
    assert abs(fd - g.dot(d)) <= 1e-5 * max(abs(fd), np.linalg.norm(g) * 1e-3)
