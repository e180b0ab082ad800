"""Device-resident timing of the three sweeps (CUDA events, L2 flushed before every launch) on the
headline workload, through the C ABI only. TRIMEM_B200_LIB selects a kernel variant."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import cabi, workload

subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rep = sys.argv[2] if len(sys.argv) > 2 else "stress"
iters = 30
ns = sys.argv[3] if len(sys.argv) > 3 else "verlet-list"
mesh, estore, x0 = workload.make_case(core, subdiv, repulsion=rep, n_search=ns)
he = mesh.halfedge_arrays()
cm = cabi.Mesh(he[0], he[1], he[2], he[3], xyz=x0)
rp = estore.eparams.repulse_params
KIND = cabi.NLIST_CELL if ns == "cell-list" else cabi.NLIST_VERLET
nl = cabi.NList(cm, KIND, rp.rlist, rp.exclusion_level)
p = workload.abi_params(estore)
e, props = cm.energy(nl, p)
g = cm.gradient(nl, p)
out = {"list": ns, "lib": os.environ.get("TRIMEM_B200_LIB", "default"), "V": cm.V, "rep": rep, "E": e, "gnorm": float(np.linalg.norm(g)),
       "E_core": estore.energy(mesh), "gnorm_core": float(np.linalg.norm(estore.gradient(mesh)))}
for name, w in (("props_full", 0), ("props_geo", 1), ("grad", 2), ("pairs", 3), ("EG", 4), ("G_newpos", 5)):
    out[name + "_us"] = round(1e3 * cm.time_kernel(nl, p, w, iters, 512 << 20), 2)
    out[name + "_warmL2_us"] = round(1e3 * cm.time_kernel(nl, p, w, iters, 0), 2)
cm.sync()
import time
ts = []
for _ in range(3):
    cm.sync(); t0 = time.time(); n2 = cabi.NList(cm, KIND, rp.rlist, rp.exclusion_level); cm.sync(); ts.append(time.time() - t0); n2.close()
out["verlet_rebuild_ms"] = round(1e3 * min(ts), 2)
print(json.dumps(out))
