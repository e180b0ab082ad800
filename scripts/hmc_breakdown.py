"""Host-side breakdown of one MeshHMC move at 40,962 vertices (config 5 chain): where the milliseconds go."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import workload, hmc
mesh, es, x0 = workload.make_case(core, 6, repulsion="physical", n_search="verlet-list")
L = 10
for rng in (np.random.RandomState(1), np.random.default_rng(1)):
    ch = hmc.MeshHMC(mesh, es, options={"time_step": 1e-6, "num_integration_steps": L, "info_step": 0}, verbose=False, rng=rng)
    for _ in range(5): ch.step(); es.update_repulsion(mesh)
    T = {}
    def tick(name, t0): T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    n = 30
    for _ in range(n):
        t0 = time.perf_counter(); p = ch.rng.normal(size=ch.x.shape) * np.sqrt(ch.m * ch.T); tick("momenta draw", t0)
        t0 = time.perf_counter(); ch._set_mesh(ch.x); tick("copy x into points()", t0)
        t0 = time.perf_counter(); xn, pn, st = core.vv_integrate(ch.trimesh, ch.estore, p, ch.m, ch.dt, ch.L); tick("core.vv_integrate (upload p, x; trajectory; download x, p, stats)", t0)
        t0 = time.perf_counter(); core.vv_displacement(ch.trimesh, ch.estore); ch.x = xn; ch._set_mesh(ch.x); tick("bookkeeping + copy back", t0)
        t0 = time.perf_counter(); es.update_repulsion(mesh); tick("update_repulsion", t0)
    t0 = time.perf_counter()
    for _ in range(n): ch.step(); es.update_repulsion(mesh)
    whole = (time.perf_counter() - t0) / n
    print(type(rng).__name__, {k: round(1e3 * v / n, 3) for k, v in T.items()}, "step()+update ms", round(1e3 * whole, 3))
