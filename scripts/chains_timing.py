"""Wall clock of repeated rounds of moves of R independent HMC + flip chains (bench.py: replica_run), to see the spread."""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trimem_b200.core as core
from trimem_b200 import workload, hmc
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if len(sys.argv) > 3:
    sys.setswitchinterval(float(sys.argv[3]))      # the interpreter's thread switch interval (default 5 ms)
chains = []
for gid in range(R):
    mesh, estore, x0 = workload.make_case(core, 6, repulsion="physical", n_search="verlet-list", seed=gid,
                                          kappa=(10.0 * (1 + gid % 8), 1.0e6, 1.0e6, 0.0, 1.0e5, 1.0e3),
                                          fracs=(1.0, 0.5 + 0.5 * ((gid // 8) % 8) / 7.0, 1.0))
    ch = hmc.MeshHMC(mesh, estore, options={"time_step": 1.0e-6, "num_integration_steps": 10, "info_step": 0},
                     verbose=False, rng=np.random.default_rng(1000 + gid))
    chains.append((mesh, estore, ch, gid))
tt = [[0.0, 0.0, 0.0] for _ in range(R)]
def run(k, n_moves, first):
    mesh, estore, ch, gid = chains[k]
    for mv in range(first, first + n_moves):
        t0 = time.perf_counter(); ch.step()
        t1 = time.perf_counter(); estore.update_repulsion(mesh)
        t2 = time.perf_counter()
        if (mv + 1) % 10 == 0:
            core.pflip(mesh, estore, 0.1, seed=977 * gid + mv)
        t3 = time.perf_counter()
        tt[k][0] += t1 - t0; tt[k][1] += t2 - t1; tt[k][2] += t3 - t2
def all_chains(n, first):
    th = [threading.Thread(target=run, args=(k, n, first)) for k in range(R)]
    [t.start() for t in th]; [t.join() for t in th]
all_chains(10, 0)
for rep in range(reps):
    for k in range(R): tt[k] = [0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    all_chains(20, 10 + 20 * rep)
    w = time.perf_counter() - t0
    s = np.array(tt).sum(axis=0) / R
    print("rep %d: %.0f moves/s  wall %.3f s; per chain: step %.3f s, update_repulsion %.3f s, pflip %.3f s" % (rep, R * 20 / w, w, s[0], s[1], s[2]), flush=True)
