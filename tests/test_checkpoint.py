"""Checkpoint hand-off (src/trimem/mc/output/checkpoint.py, mc/util.py:50-160).

CPU: the writer/reader pair (enumerated parts, XML index, array container, configuration round trip).
GPU: the restart flow against the reference -- state handed over = (x, f, continuation lambda/delta) after a flip
sweep; EnergyManager rebuilt on the initial mesh, mesh rebuilt from the checkpoint, list rebuilt on it; energy,
the six sums and the gradient against the reference's answers on ITS restarted objects
(tests/golden/restart_*.npz) at 1e-10, and against the uninterrupted objects of this implementation."""
import configparser

import numpy as np
import pytest

import cases
from trimem_b200 import checkpoint, workload


def make_config(prefix):
    cfg = configparser.ConfigParser()
    cfg.read_dict({"GENERAL": {"restart_prefix": str(prefix), "algorithm": "hmc"},
                   "ENERGY": {"kappa_b": "30.0", "continuation_lambda": "0.0", "continuation_delta": "0.0"},
                   "HMC": {"num_steps": "10"}})
    return cfg


def test_writer_reader_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    cfg = make_config(tmp_path / "run")
    pts = [rng.normal(size=(7, 3)) for _ in range(3)]
    cells = [rng.integers(0, 7, (4, 3)).astype(np.int32) for _ in range(3)]
    names = []
    for p, c in zip(pts, cells):
        w = checkpoint.CheckpointWriter(tmp_path / "run")
        w.write(p, c, cfg)
        names.append(w.fname.name)
    assert names == ["run.p0.cpt", "run.p1.cpt", "run.p2.cpt"]            # enumerated like output/_common.py
    for k in (0, 1, 2, -1):
        p, c, conf = checkpoint.CheckpointReader(tmp_path / "run", k).read()
        assert np.array_equal(p, pts[k]) and p.dtype == np.float64          # bit-exact positions
        assert np.array_equal(c, cells[k]) and c.dtype == np.int32
        assert conf["ENERGY"]["kappa_b"] == "30.0" and conf["GENERAL"]["algorithm"] == "hmc"
    # a damaged index is reported the way the reference reports it
    bad = (tmp_path / "run.p0.cpt").read_text().replace("<points", "<pints").replace("</points", "</pints")
    (tmp_path / "run.p0.cpt").write_text(bad)
    with pytest.raises(ValueError, match="Checkpoint file corrupt"):
        checkpoint.CheckpointReader(tmp_path / "run", 0).read()


def test_golden_restart_state_is_a_valid_mesh(core):
    """host side only: the reference's post-flip faces build a closed manifold with the same vertex count."""
    gold = cases.load_golden("restart_ico4_flip_verlet")
    mesh = core.TriMesh(gold["x"], gold["f"].astype(np.int32))
    assert mesh.n_vertices() == len(gold["x"]) and mesh.n_edges() == 3 * len(gold["x"]) - 6
    assert np.array_equal(mesh.fv_indices(), gold["f"])
    v0, f0 = cases.mesh_of("ico4_flip_verlet")
    assert (np.sort(gold["f"], axis=1) != np.sort(f0, axis=1)).any()      # the sweep did flip something


@pytest.mark.gpu
def test_restart_from_checkpoint_matches_reference(gpu_core, tmp_path):
    m = gpu_core
    case = "ico4_flip_verlet"
    gold = cases.load_golden("restart_" + case)
    fcase, edges, u = cases.flip_inputs(cases.RESTART_FLIPS)
    assert fcase == case

    # --- the run that is interrupted: flips (same injected randomness as the reference), continuation, motion
    mesh, es = cases.build(m, case)
    x_init, f_init = mesh.points().copy(), mesh.fv_indices().copy()
    m.flip_with(mesh, es, edges.tolist(), u.tolist())
    es.eparams.continuation_params.delta, es.eparams.continuation_params.lam = cases.RESTART_CONT
    es.eparams.area_frac, es.eparams.volume_frac, es.eparams.curvature_frac = cases.RESTART_FRACS
    es.update_reference_properties()
    es.update_reference_properties()
    wrapped = checkpoint.Mesh(core=m)
    wrapped.trimesh = mesh
    wrapped.x = cases.restart_displacement(mesh.points())
    es.update_repulsion(mesh)
    e_cont, g_cont = es.energy(mesh), es.gradient(mesh)
    assert np.array_equal(wrapped.f, gold["f"])                  # every flip of the sweep as in the reference
    cfg = make_config(tmp_path / "run")
    checkpoint.write_checkpoint_handle(cfg)(wrapped, es, {"move": 3})

    # --- restart (mc/util.py:153-160)
    cfg2 = make_config(tmp_path / "run")
    mesh2, cfg2 = checkpoint.read_checkpoint(cfg2, -1, core=m)
    assert np.array_equal(mesh2.x, gold["x"]) and np.array_equal(mesh2.f, gold["f"])
    assert cfg2["DEFAULT"]["init_step"] == '{"move": 3}'
    mesh0 = m.TriMesh(x_init, f_init)
    ep = workload.energy_params(m, mesh0, **cases.CASES[case][2])
    ep.area_frac, ep.volume_frac, ep.curvature_frac = cases.RESTART_FRACS
    cp = ep.continuation_params
    cp.lam, cp.delta = float(cfg2["ENERGY"]["continuation_lambda"]), float(cfg2["ENERGY"]["continuation_delta"])
    ep.continuation_params = cp
    assert cp.lam == float(gold["lam"][0]) and cp.delta == float(gold["delta"][0])
    es2 = m.EnergyManager(mesh0, ep)
    # the list of es2 was built on mesh0: using it on another TriMesh object is refused, not silently wrong
    with pytest.raises(RuntimeError, match="another mesh object"):
        es2.energy(mesh2.trimesh)
    es2.update_repulsion(mesh2.trimesh)
    e2, g2 = es2.energy(mesh2.trimesh), es2.gradient(mesh2.trimesh)
    p2 = cases.props6(es2.properties(mesh2.trimesh))

    tol = 1e-10
    assert abs(e2 - gold["energy"][0]) <= tol * abs(gold["energy"][0])
    assert np.all(np.abs(p2 - gold["props"]) <= tol * np.abs(gold["props"]))
    assert np.linalg.norm(g2 - gold["grad"]) <= tol * np.linalg.norm(gold["grad"])
    # and the restarted objects continue where the interrupted ones were (ring order differs: not bitwise)
    assert abs(e2 - e_cont) <= 1e-12 * abs(e_cont)
    assert np.linalg.norm(g2 - g_cont) <= 1e-12 * np.linalg.norm(g_cont)
