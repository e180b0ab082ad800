"""Debug driver: 2 ranks on one device, short trajectory, dump the sequencing blocks."""
import os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cases
import trimem_b200.core as core
from trimem_b200 import cabi, workload

nranks = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with_e = (sys.argv[3] == "1") if len(sys.argv) > 3 else False
case, p0, mass, dt, n = cases.vv_inputs("vv_ico4_stress_verlet2")
mesh, es = cases.build(core, case)
v = mesh.points().copy()
arrs = mesh.halfedge_arrays()[:4]
rp = es.eparams.repulse_params
par = workload.abi_params(es)
cm = [cabi.Mesh(*arrs, xyz=v) for _ in range(nranks)]
for r, m in enumerate(cm):
    m.part_create(r, nranks)
for r, m in enumerate(cm):
    for q, o in enumerate(cm):
        if q != r:
            m.part_attach_local(q, o)
nl = [cabi.NList(m, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level) for m in cm]
for m, l in zip(cm, nl):
    m.part_prepare(l)
    m.set_momenta(p0)
print([m.part_info() for m in cm], flush=True)

def each(fn):
    out = [None] * nranks
    def run(r):
        try:
            out[r] = fn(cm[r], nl[r])
        except Exception as exc:
            out[r] = exc
    th = [threading.Thread(target=run, args=(r,)) for r in range(nranks)]
    [t.start() for t in th]; [t.join() for t in th]
    return out

print("launch", each(lambda m, l: m.vv_integrate_async(l, par, mass, dt, nsteps, with_energies=with_e)), flush=True)
print("sync", each(lambda m, l: m.sync()), flush=True)
for m in cm:
    print(m.part_debug(), flush=True)
