"""Launch the sweeps a few times on the headline workload (for ncu captures).
usage: prof_kernels.py [subdiv] [repulsion] [which ...]   which: 0 energy sweeps, 1 gradient's property sweep,
2 ring gradient, 3 pair sweep (default: 0 1 2)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import cabi, workload
subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rep = sys.argv[2] if len(sys.argv) > 2 else "stress"
which = [int(a) for a in sys.argv[3:]] or [0, 1, 2]
mesh, estore, x0 = workload.make_case(core, subdiv, repulsion=rep, n_search="verlet-list")
he = mesh.halfedge_arrays()
cm = cabi.Mesh(he[0], he[1], he[2], he[3], xyz=x0)
rp = estore.eparams.repulse_params
nl = cabi.NList(cm, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
p = workload.abi_params(estore)
for _ in range(3):
    for w in which:
        cm.time_kernel(nl, p, w, 1, 0)
cm.sync()
