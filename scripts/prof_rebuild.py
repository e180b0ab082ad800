"""A few verlet-list rebuilds on the headline mesh (for ncu launch lists)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import cabi, workload
subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mesh, estore, x0 = workload.make_case(core, subdiv, repulsion="stress", n_search="verlet-list")
he = mesh.halfedge_arrays()
cm = cabi.Mesh(he[0], he[1], he[2], he[3], xyz=x0)
rp = estore.eparams.repulse_params
for k in range(4):
    cm.sync(); t0 = time.perf_counter()
    nl = cabi.NList(cm, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level); cm.sync()
    print("rebuild %d: %.3f ms" % (k, 1e3 * (time.perf_counter() - t0)))
    nl.close()
