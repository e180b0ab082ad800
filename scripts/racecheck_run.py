"""Small run of the kernels that use shared memory or cross-thread ordering, for
`compute-sanitizer --tool racecheck` / `--tool memcheck`: verlet build, cell-list pair sweep, ring sweeps, both flip
modes, a short trajectory (icosphere subdiv 4, stress repulsion)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import workload
for ns in ("verlet-list", "cell-list"):
    mesh, es, x0 = workload.make_case(core, 4, repulsion="stress", n_search=ns)
    e, g = es.energy(mesh), es.gradient(mesh)
    es.update_repulsion(mesh)
    a = core.pflip(mesh, es, 0.1, seed=1)
    b = core.flip(mesh, es, 0.02, seed=2)
    p0 = np.random.default_rng(0).normal(size=x0.shape)
    xn, pn, st = core.vv_integrate(mesh, es, p0, 1.0, 1.0e-6, 3)
    print(ns, "E", e, "|g|", float(np.linalg.norm(g)), "pflip", a, "flip", b, "H", st[0] + st[1], st[2] + st[3])
if "split" in sys.argv:
    # the pair sweep cut across the side streams, each range's ring gradient behind it (>= 131,072 vertices)
    from trimem_b200 import cabi
    mesh, es, x0 = workload.make_case(core, 7, repulsion="stress", n_search="verlet-list")
    he = mesh.halfedge_arrays()
    cm = cabi.Mesh(he[0], he[1], he[2], he[3], xyz=x0)
    rp = es.eparams.repulse_params
    nl = cabi.NList(cm, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
    p = workload.abi_params(es)
    cm.energy_async(nl, p); cm.gradient_async(nl, p)
    e, _ = cm.fetch_energy(p)
    g = cm.fetch_gradient()
    p0 = np.random.default_rng(0).normal(size=x0.shape)
    xn, pn, st = cm.vv_integrate(nl, p, p0, 1.0, 5.0e-8, 3)
    print("split", "E", e, "|g|", float(np.linalg.norm(g)), "H", st[0] + st[1], st[2] + st[3])
