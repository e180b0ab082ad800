"""One pflip sweep at 40,962 vertices (config 2 flip statistics setup) for ncu launch lists."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import workload
mesh, es, _ = workload.make_case(core, 6, repulsion="physical", n_search="cell-list",
                                 kappa=(1.0, 1.0, 1.0, 1.0, 1.0, 1.0), fracs=(1.0, 1.0, 1.0))
core.pflip(mesh, es, 0.1, seed=1)
n0 = core.launch_count()
t0 = time.perf_counter()
acc = core.pflip(mesh, es, 0.1, seed=2)
print("pflip ms", 1e3 * (time.perf_counter() - t0), "accepted", acc, "launches", core.launch_count() - n0)
for k in range(4):
    t0 = time.perf_counter()
    acc = core.pflip(mesh, es, 0.1, seed=3 + k)
    print("pflip ms", 1e3 * (time.perf_counter() - t0), "accepted", acc)
