"""Energy sweeps with a cell list on the headline mesh (for ncu captures of the cell-list pair sweep)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import cabi, workload
subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mesh, estore, x0 = workload.make_case(core, subdiv, repulsion="stress", n_search="cell-list")
he = mesh.halfedge_arrays()
cm = cabi.Mesh(he[0], he[1], he[2], he[3], xyz=x0)
rp = estore.eparams.repulse_params
nl = cabi.NList(cm, cabi.NLIST_CELL, rp.rlist, rp.exclusion_level)
p = workload.abi_params(estore)
for _ in range(3):
    cm.time_kernel(nl, p, 3, 1, 0)
cm.sync()
