"""pflip / flip sweeps at 40,962 vertices (config 2 flip statistics setup) for timings and ncu launch lists."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import workload
ns = sys.argv[1] if len(sys.argv) > 1 else "cell-list"
nsweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mesh, es, _ = workload.make_case(core, 6, repulsion="physical", n_search=ns,
                                 kappa=(1.0, 1.0, 1.0, 1.0, 1.0, 1.0), fracs=(1.0, 1.0, 1.0))
core.pflip(mesh, es, 0.1, seed=1)
for k in range(nsweeps):
    n0 = core.launch_count()
    t0 = time.perf_counter()
    acc = core.pflip(mesh, es, 0.1, seed=2 + k)
    print(ns, "pflip ms %.2f" % (1e3 * (time.perf_counter() - t0)), "accepted", acc, "launches", core.launch_count() - n0)
t0 = time.perf_counter()
acc = core.flip(mesh, es, 0.1, seed=99)
print(ns, "flip (serial) ms %.1f" % (1e3 * (time.perf_counter() - t0)), "accepted", acc)
