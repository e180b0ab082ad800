"""Phases of the step positions in -> energy -> gradient out at 655,362 vertices: the reference's idiom
np.copyto(mesh.points(), x) against TriMesh.set_points(x) (TM_SET_POINTS_THREADS host threads)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import workload
subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mesh, estore, x0 = workload.make_case(core, subdiv, repulsion="stress", n_search="verlet-list")
xs = [x0 + 1e-6 * np.random.default_rng(s).normal(size=x0.shape) for s in range(2)]
for name, put in (("copyto", lambda x: np.copyto(mesh.points(), x)), ("set_points", lambda x: mesh.set_points(x))):
    ph = [0.0, 0.0, 0.0]
    for k in range(24):
        if k == 4: ph = [0.0, 0.0, 0.0]
        t0 = time.perf_counter(); put(xs[k % 2])
        t1 = time.perf_counter(); e = estore.energy(mesh)
        t2 = time.perf_counter(); g = estore.gradient(mesh)
        t3 = time.perf_counter()
        ph[0] += t1 - t0; ph[1] += t2 - t1; ph[2] += t3 - t2
    print(name, "threads", os.environ.get("TM_SET_POINTS_THREADS", "4"), "ms:", [round(1e3 * p / 20, 3) for p in ph], "sum", round(1e3 * sum(ph) / 20, 3), "E", e)
