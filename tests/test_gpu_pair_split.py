"""The pair sweep cut into ranges of tiles on side streams, each range's ring gradient following it on that range's
stream, the repulsion sum added up by the sweep itself (launch_pairs / launch_grad_kernels / k_pairs in
tm_energy.cu; meshes of >= 131,072 vertices with a verlet list). Same kernels, same arithmetic per vertex, another
launch shape: gradients, momenta and positions must be bit-identical to the path where one pair sweep is joined in
front of the reduction (TM_PAIR_SPLIT=0) and to the synchronous calls; the repulsion sum is added in another (fixed)
order, so the energy is held to 1e-14 relative and to identical bits run to run. The expressions themselves
(mesh_repulsion.h:28-79, mesh_properties.cpp:139-146) are pinned against the reference at this size by the ico8
goldens of tests/test_gpu_energy.py, which run through the same path."""
import os
import subprocess
import sys

import numpy as np
import pytest

from trimem_b200 import cabi, workload

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUBDIV = 7                                    # 163,842 vertices = 1,281 tiles: above the split threshold


def make(core, subdiv):
    mesh, estore, x0 = workload.make_case(core, subdiv, repulsion="stress", n_search="verlet-list")
    he = mesh.halfedge_arrays()
    cm = cabi.Mesh(he[0], he[1], he[2], he[3], xyz=x0)
    rp = estore.eparams.repulse_params
    nl = cabi.NList(cm, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
    return cm, nl, workload.abi_params(estore), x0


def test_async_evaluations_equal_synchronous_ones(gpu_core):
    cm, nl, p, x0 = make(gpu_core, SUBDIV)
    rng = np.random.default_rng(7)
    amp = 1e-3 * np.linalg.norm(x0[0] - x0[1])
    for it in range(8):
        x = x0 + amp * rng.standard_normal(x0.shape)
        cm.set_positions(x)
        if it % 3 == 1:                       # an energy evaluation nobody follows with a gradient
            cm.energy_async(nl, p)
            cm.sync()
        if it % 4 == 3:
            cm.energy_async(nl, p)
            cm.energy_async(nl, p)            # (the second one finds everything cached)
        cm.energy_async(nl, p)
        cm.gradient_async(nl, p)              # the pair sweep's ranges are still running
        e1, props1 = cm.fetch_energy(p)
        g1 = cm.fetch_gradient()
        # everything cached and complete: the ring gradient alone, one launch, reads the stored pair force
        g2 = cm.gradient(nl, p)
        e2, props2 = cm.energy(nl, p)
        assert np.array_equal(g1, g2), it
        assert e1 == e2 and np.array_equal(props1, props2), it
        # gradient first: geometry-only property sweep, force-only pair sweep beside it
        cm.set_positions(x + 1e-9)
        cm.gradient_async(nl, p)
        g3 = cm.fetch_gradient()
        cm.energy_async(nl, p)
        cm.sync()
        g4 = cm.gradient(nl, p)
        assert np.array_equal(g3, g4), it


_RUN = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
import trimem_b200.core as core
from trimem_b200 import cabi, workload
mesh, estore, x0 = workload.make_case(core, {subdiv}, repulsion="stress", n_search="verlet-list")
he = mesh.halfedge_arrays()
cm = cabi.Mesh(he[0], he[1], he[2], he[3], xyz=x0)
rp = estore.eparams.repulse_params
nl = cabi.NList(cm, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
p = workload.abi_params(estore)
p0 = np.random.default_rng(3).standard_normal(x0.shape)
out = []
for rep in range(3):                          # direct run, graph capture, graph replay
    cm.set_positions(x0)
    x, pm, st = cm.vv_integrate(nl, p, p0, 1.0, 5e-8, 5)
    assert np.isfinite(x).all() and np.isfinite(pm).all() and np.isfinite(st).all()
    out += [x, pm, st]
cm.set_positions(x0)
e, props = cm.energy(nl, p)
g = cm.gradient(nl, p)
out += [np.array([e]), props, g]
np.savez(sys.argv[1], *out)
"""


def test_split_and_joined_builds_of_a_trajectory_agree(gpu_core, tmp_path):
    """Five fused leapfrog steps, three times over (direct, captured, replayed), then E and G: the default (2 ranges),
    4 ranges, and one joined sweep (TM_PAIR_SPLIT=0)."""
    script = tmp_path / "run.py"
    script.write_text(_RUN.format(root=ROOT, subdiv=SUBDIV))
    res = {}
    for tag, env in (("default", {}), ("four", {"TM_PAIR_SPLIT": "4"}), ("joined", {"TM_PAIR_SPLIT": "0"})):
        out = tmp_path / (tag + ".npz")
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, str(script), str(out)], check=True, env=e, timeout=900)
        res[tag] = np.load(out)
    a = res["default"]
    for rep in (1, 2):                         # run to run, graph or not: identical bits
        for j in range(3):
            assert np.array_equal(a["arr_%d" % j], a["arr_%d" % (3 * rep + j)]), (rep, j)
    for other in ("four", "joined"):
        b = res[other]
        for j in (0, 1, 3, 4, 11):             # positions, momenta, gradient: the same arithmetic per vertex
            assert np.array_equal(a["arr_%d" % j], b["arr_%d" % j]), (other, j)
        for j in (2, 5, 8, 9, 10):             # energies and sums: the repulsion sum is added in another fixed order
            assert np.allclose(a["arr_%d" % j], b["arr_%d" % j], rtol=1e-14, atol=0.0), (other, j)


def test_split_path_with_generic_vertices(gpu_core):
    """After flip sweeps the mesh has vertices of valence > 8 (no fixed-width ring row: k_props_generic / k_grad_generic,
    launched on the main stream behind the join): asynchronous evaluations still equal the synchronous ones."""
    mesh, estore, x0 = workload.make_case(gpu_core, SUBDIV, repulsion="physical", n_search="verlet-list",
                                          kappa=(1.0, 1.0, 1.0, 1.0, 1.0, 1.0), fracs=(1.0, 1.0, 1.0))
    for k in range(4):
        gpu_core.pflip(mesh, estore, 0.3, seed=20 + k)
        estore.update_repulsion(mesh)
    he = mesh.halfedge_arrays()
    valence = np.bincount(he[0], minlength=mesh.n_vertices())
    assert valence.max() > 8, "the sweeps were meant to produce generic vertices"
    cm = cabi.Mesh(he[0], he[1], he[2], he[3], xyz=x0)
    rp = estore.eparams.repulse_params
    nl = cabi.NList(cm, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
    p = workload.abi_params(estore)
    rng = np.random.default_rng(11)
    amp = 1e-3 * np.linalg.norm(x0[0] - x0[1])
    for it in range(3):
        cm.set_positions(x0 + amp * rng.standard_normal(x0.shape))
        cm.energy_async(nl, p)
        cm.gradient_async(nl, p)
        e1, props1 = cm.fetch_energy(p)
        g1 = cm.fetch_gradient()
        g2 = cm.gradient(nl, p)
        e2, props2 = cm.energy(nl, p)
        assert np.array_equal(g1, g2) and e1 == e2 and np.array_equal(props1, props2), it
    # and against the pybind path (synchronous calls on its own mesh object)
    np.copyto(mesh.points(), cm.positions())
    assert estore.energy(mesh) == e2
    assert np.array_equal(estore.gradient(mesh), g2)
