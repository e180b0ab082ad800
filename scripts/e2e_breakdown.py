"""Where the end-to-end (host-buffer) E+G step spends its time."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import cabi, workload
subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mesh, estore, x0 = workload.make_case(core, subdiv, repulsion="stress", n_search="verlet-list")
xs = [x0 + 1e-6 * np.random.default_rng(s).normal(size=x0.shape) for s in range(2)]
def T(f, n=10):
    f(); t0 = time.perf_counter()
    for _ in range(n): f()
    return 1e3 * (time.perf_counter() - t0) / n
k = [0]
def copy():
    np.copyto(mesh.points(), xs[k[0] % 2]); k[0] += 1
print("np.copyto(points, x)        %.3f ms" % T(copy))
print("estore.energy(mesh)         %.3f ms" % T(lambda: estore.energy(mesh)))
print("estore.gradient(mesh)       %.3f ms" % T(lambda: estore.gradient(mesh)))
print("estore.properties(mesh)     %.3f ms" % T(lambda: estore.properties(mesh)))
def step():
    copy(); estore.energy(mesh); estore.gradient(mesh)
print("full step                   %.3f ms" % T(step))
he = mesh.halfedge_arrays()
cm = cabi.Mesh(he[0], he[1], he[2], he[3], xyz=x0)
pageable = np.array(x0)
print("cabi set_positions pageable %.3f ms" % T(lambda: cm.set_positions(pageable)))
out = np.empty_like(x0)
import ctypes as C
print("cabi get_positions pageable %.3f ms" % T(lambda: cm.positions()))
print("np.empty+fill 15MB          %.3f ms" % T(lambda: np.empty_like(x0).fill(0.0)))
a = np.empty_like(x0)
print("memcpy 15MB warm            %.3f ms" % T(lambda: np.copyto(a, x0)))

# raw PCIe copies of the same size through torch (pinned host memory), both directions
try:
    import torch
    h = torch.empty(x0.size, dtype=torch.float64).pin_memory()
    d = torch.empty(x0.size, dtype=torch.float64, device="cuda")
    def h2d():
        d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    def d2h():
        h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    print("torch pinned H2D 15.7MB     %.3f ms" % T(h2d))
    print("torch pinned D2H 15.7MB     %.3f ms" % T(d2h))
except Exception as exc:
    print("torch copy probe failed:", exc)
for rep in range(3):
    g = None
    ts = []
    for _ in range(10):
        copy(); estore.energy(mesh)
        t0 = time.perf_counter(); g = estore.gradient(mesh); ts.append(1e3 * (time.perf_counter() - t0))
    print("gradient() in a step loop   " + " ".join("%.2f" % t for t in ts))
