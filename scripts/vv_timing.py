"""Per-call timing of repeated identical trajectories (direct / capture / replay)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import cabi, workload
subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 6
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10
mesh, estore, x0 = workload.make_case(core, subdiv, repulsion="stress", n_search="verlet-list")
he = mesh.halfedge_arrays()
cm = cabi.Mesh(he[0], he[1], he[2], he[3], xyz=x0)
rp = estore.eparams.repulse_params
nl = cabi.NList(cm, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
p = workload.abi_params(estore)
p0 = np.random.default_rng(1).normal(size=x0.shape)
for k in range(10):
    cm.set_positions(x0); cm.set_momenta(p0); cm.sync()
    n0 = cabi.lib().tm_launch_count()
    t0 = time.perf_counter()
    cm.vv_integrate_async(nl, p, 1.0, 2e-8, L)
    t1 = time.perf_counter()
    cm.sync()
    t2 = time.perf_counter()
    print("call %d: enqueue %.3f ms, total %.3f ms, launches %d" % (k, 1e3 * (t1 - t0), 1e3 * (t2 - t0), cabi.lib().tm_launch_count() - n0))
