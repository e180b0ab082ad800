"""Host-side logic of the product and the C-ABI boundary. CPU only: no compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

import cases
from trimem_b200 import cabi, meshgen


def test_cabi_library_exports_every_declared_symbol():
    names = cabi.header_symbols()
    assert len(names) >= 35
    L = ctypes.CDLL(cabi.lib_path())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert cabi.lib().tm_abi_version() == 4


def test_cabi_param_struct_layout():
    # struct tm_energy_params: 12 doubles, 2 ints, 1 double, 2 ints
    assert ctypes.sizeof(cabi.EnergyParamsABI) == 12 * 8 + 8 + 8 + 8


def test_no_cpu_fallback_without_device(core):
    if core.device_count() > 0:
        pytest.skip("a CUDA device is present")
    v, f = meshgen.icosphere(1)
    mesh = core.TriMesh(v, f)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        core.EnergyManager(mesh, core.EnergyParams())
    h = ctypes.c_void_p()
    assert cabi.lib().tm_mesh_create(0, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in cabi.lib().tm_last_error()


def test_halfedge_numbering_matches_reference(core):
    gold = cases.load_golden("topology")
    for name, (v, f) in {"ico3": meshgen.icosphere(3), "rect20": meshgen.rect(20, 20),
                         "tube32": meshgen.tube(1.0, 1.0, 32), "ico4": meshgen.icosphere(4)}.items():
        mesh = core.TriMesh(v, f)
        for k, a in zip(("he_to", "he_next", "he_face", "v_he", "f_he"), mesh.halfedge_arrays()):
            assert np.array_equal(a, gold[f"{name}_{k}"]), (name, k)
        assert np.array_equal(mesh.fv_indices(), gold[f"{name}_fv"])
        assert mesh.fv_indices().dtype == np.int32
        assert mesh.n_vertices() == len(v) and mesh.n_edges() == len(gold[f"{name}_he_to"]) // 2


def test_flip_sequence_matches_reference(core):
    gold = cases.load_golden("topology")
    v, f = meshgen.icosphere(3)
    mesh = core.TriMesh(v, f)
    for e, ok in zip(gold["flipseq_edges"], gold["flipseq_ok"]):
        assert mesh.is_flip_ok(int(e)) == bool(ok)
        if ok:
            mesh.flip(int(e))
    for k, a in zip(("he_to", "he_next", "he_face", "v_he", "f_he"), mesh.halfedge_arrays()):
        assert np.array_equal(a, gold[f"flipseq_{k}"]), k
    assert np.array_equal(mesh.fv_indices(), gold["flipseq_fv"])


def test_points_is_a_writable_view_owned_by_the_mesh(core):
    v, f = meshgen.icosphere(2)
    mesh = core.TriMesh(v, f)
    x = mesh.points()
    assert x.shape == (len(v), 3) and x.dtype == np.float64
    assert x.flags.writeable and not x.flags.owndata and x.base is mesh
    np.copyto(x, v + 1.0)                         # trimem.mc.mesh.Mesh.x setter
    assert np.array_equal(mesh.points(), v + 1.0)


def test_handles(core):
    v, f = meshgen.icosphere(1)
    mesh = core.TriMesh(v, f)
    he_to, he_next, he_face, v_he, f_he = mesh.halfedge_arrays()
    h = mesh.halfedge_handle(5)
    assert isinstance(h, core.HalfedgeHandle) and h.idx() == 5
    assert mesh.opposite_halfedge_handle(h).idx() == 4
    assert mesh.next_halfedge_handle(h).idx() == he_next[5]
    core.HalfedgeHandle()                         # default constructible


def test_construction_edge_cases(core):
    assert core.TriMesh().n_vertices() == 0
    assert core.TriMesh(np.zeros((0, 3)), np.zeros((0, 3), dtype=np.int32)).n_vertices() == 0
    with pytest.raises(RuntimeError, match=r"Array 'points' must have shape \(n,3\)"):
        core.TriMesh(np.zeros((4, 2)), np.zeros((1, 3), dtype=np.int32))
    with pytest.raises(RuntimeError, match=r"Array 'cells' must have shape \(n,3\)"):
        core.TriMesh(np.zeros((4, 3)), np.zeros((1, 4), dtype=np.int32))
    # faces with out-of-range ids are skipped silently (src/mesh_py.cpp:54-63)
    v, f = meshgen.icosahedron()
    bad = np.vstack([f, [[0, 1, 99]]]).astype(np.int32)
    mesh = core.TriMesh(v, bad)
    assert mesh.fv_indices().shape == (20, 3)
    # int64 / float32 inputs are converted
    mesh = core.TriMesh(v.astype(np.float32), f.astype(np.int64))
    assert mesh.n_edges() == 30


def test_parameter_defaults_match_reference(core):
    ep = core.EnergyParams()
    assert (ep.kappa_b, ep.kappa_a, ep.kappa_v, ep.kappa_c, ep.kappa_t, ep.kappa_r) == (0, 0, 0, 0, 0, 0)
    assert (ep.area_frac, ep.volume_frac, ep.curvature_frac) == (1, 1, 1)
    bp = ep.bond_params
    assert (bp.lc0, bp.lc1, bp.a0, bp.r, bp.type) == (1, 0, 1, 2, core.BondType.Edge)
    rp = ep.repulse_params
    assert (rp.lc1, rp.r, rp.n_search, rp.rlist, rp.exclusion_level) == (0, 2, "cell-list", 0.1, 2)
    cp = ep.continuation_params
    assert (cp.delta, cp.lam) == (0, 1)
    assert core.Edge == core.BondType.Edge and core.Area == core.BondType.Area    # export_values
    vp = core.VertexProperties()
    for n in ("area", "volume", "curvature", "bending", "tethering", "repulsion"):
        assert getattr(vp, n) == 0.0
        setattr(vp, n, 1.5)


def test_argument_errors_precede_device_work(core):
    v, f = meshgen.icosphere(1)
    mesh = core.TriMesh(v, f)
    ep = core.EnergyParams()
    ep.repulse_params.exclusion_level = 3
    with pytest.raises(RuntimeError, match="Unsupported exclusion level"):
        core.make_nlist(mesh, ep)
    ep.repulse_params.n_search = "kd-tree"
    with pytest.raises(RuntimeError, match="Unknown neighbour search algorithm."):
        core.make_nlist(mesh, ep)
    ep = core.EnergyParams()
    ep.bond_params.type = getattr(core.BondType, "None")
    with pytest.raises(RuntimeError, match="Unknown bond potential"):
        core.EnergyManager(mesh, ep)


def test_product_never_touches_the_oracle():
    """The shipped package must not import, link or execute anything under oracle/."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "trimem_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                for line in text.splitlines():
                    if re.search(r"^\s*(import|from|#include)\b.*oracle", line) or "oracle/" in line or "oracle\"" in line:
                        raise AssertionError(f"{fn}: {line.strip()}")
