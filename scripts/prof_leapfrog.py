"""A short device-resident trajectory on the headline workload (for ncu launch lists / captures)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import cabi, workload
subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mesh, estore, x0 = workload.make_case(core, subdiv, repulsion="stress", n_search="verlet-list")
he = mesh.halfedge_arrays()
cm = cabi.Mesh(he[0], he[1], he[2], he[3], xyz=x0)
rp = estore.eparams.repulse_params
nl = cabi.NList(cm, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
p = workload.abi_params(estore)
cm.set_momenta(np.random.default_rng(1).normal(size=x0.shape))
cm.vv_integrate_async(nl, p, 1.0, 2e-8, steps, True)
print(cm.vv_stats(p)[0])
