#!/usr/bin/env python
"""Benchmark of the force-evaluation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one E+G evaluation = one EnergyManager.energy() + one .gradient() on the
headline workload (BASELINE.json configs[2]): jittered icosphere, subdiv 8 (655,362
vertices), tether + stress surface repulsion over a verlet list, exclusion level 2.

  value    device-resident E+G evals/s (positions already in HBM, CUDA events on the
           library's stream, L2 flushed before every timed step), whole job (all ranks)
  e2e      the same through trimem_b200.core with host buffers: per step
           np.copyto(mesh.points(), x) ; estore.energy(mesh) ; estore.gradient(mesh)
  roofline dominant kernel (gather gradient sweep): algorithmic bytes / event time
  cpu_baseline / --impl reference: the reference's own OpenMP C++ path (oracle/_ref)
           on the host cores, same mesh and parameters.
Multi-GPU (N>1): independent replicas, one per rank, no data-path collective (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLUSH_BYTES = 512 << 20          # > 126 MB L2

# Counters of the three sweeps from the committed `ncu --set full` capture of this build at 655,362 vertices
# (per launch; profiles/README.md says how they were taken). They explain the live timings of the roofline object.
NCU_SOURCE = "profiles/r02_ncu_full_sweeps.txt, profiles/r02_ncu_full_grad.txt (per launch, 655,362 vertices)"
# issue bound = 2 * (fp64 warp instructions) + (other warp instructions) cycles per scheduler: an fp64 warp instruction holds a
# scheduler's issue port for two cycles on B200 (64 fp64 lanes per SM), DESIGN.md section 4
NCU = {
    "k_pairs": {
        "ncu_us_cold_serialised": 63.104,
        "dram_bytes_per_launch": 81661952.0,
        "warp_instructions": 36611824.0,
        "fp64_warp_instructions": 16903951,
        "fp64_pipe_pct": 53.8,
        "issue_bound_cycles_per_scheduler": 90398,
        "active_cycles_per_scheduler": 106085,
        "issue_bound_fraction": 0.852,
        "registers": 80,
        "warps_active_pct": 34.057766
    },
    "k_props": {
        "ncu_us_cold_serialised": 37.216,
        "dram_bytes_per_launch": 44592639.99999999,
        "warp_instructions": 18982564.0,
        "fp64_warp_instructions": 9878346,
        "fp64_pipe_pct": 54.8,
        "issue_bound_cycles_per_scheduler": 48752,
        "active_cycles_per_scheduler": 60866,
        "issue_bound_fraction": 0.801,
        "registers": 64,
        "warps_active_pct": 44.341362
    },
    "k_grad": {
        "ncu_us_cold_serialised": 38.56,
        "dram_bytes_per_launch": 95247104.0,
        "warp_instructions": 12132445.0,
        "fp64_warp_instructions": 7113958,
        "fp64_pipe_pct": 37.4,
        "issue_bound_cycles_per_scheduler": 32511,
        "active_cycles_per_scheduler": 64324,
        "issue_bound_fraction": 0.505,
        "registers": 128,
        "warps_active_pct": 21.139982
    }
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


_POLLER = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
while True:
    try:
        print(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM),
              nv.nvmlDeviceGetCurrentClocksThrottleReasons(h), flush=True)
    except Exception as exc:
        print("ERR", repr(exc), flush=True)
        break
    time.sleep(0.005)
"""


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions (B200_PROFILING.md). NVML is
    polled every 5 ms from a separate PROCESS: the timed loops are ~10-100 ms long (too short for
    `nvidia-smi -lms`), and a sampling thread in this process would fight the e2e loop for the GIL."""
    REASONS = (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
               ("hw_thermal_slowdown", 0x40))

    def __init__(self, gpu):
        self.gpu, self.proc = gpu, None

    @staticmethod
    def _nvml_index(gpu):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if gpu < len(ids) and ids[gpu].strip().isdigit():
                return int(ids[gpu])
        return gpu

    def start(self):
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _POLLER, str(self._nvml_index(self.gpu))],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.proc.stdout.readline()            # first sample = the poller is up
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml poller unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=3)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = line.split()
            if len(f) != 3 or f[0] == "ERR":
                continue
            sm.append(float(f[0])); mx.append(float(f[1]))
            for name, bit in self.REASONS:
                if int(f[2]) & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_reference():
    """The reference's trimem.core compiled from its own sources (oracle/_ref), or None."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    try:
        import trimem.core as ref
        return ref
    except ImportError:
        return None


def time_reference(subdiv, steps, warmup, threads=None):
    """E+G evals/s of the reference's OpenMP path on the host cores (same mesh, same parameters)."""
    from trimem_b200 import workload
    threads = threads or os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(threads)
    ref = load_reference()
    if ref is not None:
        kind = "reference"
        t0 = time.time()
        mesh, estore, x = workload.make_case(ref, subdiv, repulsion="stress", n_search="verlet-list")
        setup = time.time() - t0
        def step():
            np.copyto(mesh.points(), x)
            e = estore.energy(mesh)
            g = estore.gradient(mesh)
            return e, g
    else:
        kind = "port"
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_port
        case = oracle_port.make_case(subdiv, repulsion="stress", n_search="verlet-list", threads=threads)
        setup = case.setup_seconds
        step = case.energy_and_gradient
    for _ in range(warmup):
        step()
    t0 = time.time()
    for _ in range(steps):
        e, g = step()
    dt = time.time() - t0
    return {"value": steps / dt, "ms_per_step": 1e3 * dt / steps, "kind": kind, "cores": threads,
            "setup_s": setup, "energy": float(e), "grad_norm": float(np.linalg.norm(g))}


def replica_run(args, core, cabi, workload, rank, world, local, barrier):
    """BASELINE config 5: independent HMC + flip chains, `--replicas` per GPU, each a 40,962-vertex membrane of the
    reduced-volume / kappa_b sweep; no communication. Every chain runs through the PUBLIC API (core.TriMesh,
    core.EnergyManager, trimem_b200.hmc.MeshHMC, core.pflip) on a host thread of its own: per move one MeshHMC.step()
    (L velocity-Verlet steps on the device + Metropolis, end point and momenta copied back), update_repulsion
    (mc/evaluators.py:181-182), and every 10th move one pflip(ratio 0.1) (mc/hmc.py:325-372 with the cadence of
    config 2). The bindings release the interpreter lock around device work, so the chains' kernels overlap on the
    GPU; timed: host wall clock from the common start to the last chain's finish."""
    from trimem_b200 import hmc
    R = args.replicas
    L = max(1, args.leapfrog_steps)
    chains = []
    for k in range(R):
        gid = rank * R + k                                        # position in the 8 x 8 sweep of config 5
        mesh, estore, x0 = workload.make_case(core, 6, repulsion="physical", n_search="verlet-list", seed=gid,
                                              kappa=(10.0 * (1 + gid % 8), 1.0e6, 1.0e6, 0.0, 1.0e5, 1.0e3),
                                              fracs=(1.0, 0.5 + 0.5 * ((gid // 8) % 8) / 7.0, 1.0))
        ch = hmc.MeshHMC(mesh, estore, options={"time_step": 1.0e-6, "num_integration_steps": L, "info_step": 0},
                         verbose=False, rng=np.random.default_rng(1000 + gid))   # (a Generator draws outside the interpreter lock)
        chains.append((mesh, estore, ch, gid))
    V = chains[0][0].n_vertices()
    stats = [None] * R

    lazy_upd = [hmc.LazyRepulsion(c[2]) for c in chains]

    def run(k, n_moves, first_move, lazy=False):
        mesh, estore, ch, gid = chains[k]
        flips = acc0 = 0
        for mv in range(first_move, first_move + n_moves):
            ch.step()
            if lazy:
                lazy_upd[k]()               # rebuild only when the device-side trigger fires
            else:
                estore.update_repulsion(mesh)
            if (mv + 1) % 10 == 0:
                flips += lazy_upd[k].flips_accepted(core.pflip(mesh, estore, 0.1, seed=977 * gid + mv))
        stats[k] = (ch.acc, ch.i, flips)

    def all_chains(n_moves, first_move, lazy=False):
        th = [threading.Thread(target=run, args=(k, n_moves, first_move, lazy)) for k in range(R)]
        [t.start() for t in th]
        [t.join() for t in th]
    all_chains(10, 0)                   # warm-up: trajectory graphs captured, one flip sweep, list pools filled
    n_moves = 20
    # Host wall clock of R Python threads sharing one interpreter and one GPU; every figure of this section is the
    # shortest of three runs, like timeit. (Until the pinned result blocks of the bindings were kept up to a byte budget
    # instead of four per size, single runs spread from 0.15 to 2 s: eight chains' arrays of one size overflowed the
    # pool, and every cudaFreeHost / cudaHostAlloc stalled all of them; profiles/r02be_chains_spread.txt.)
    REPS = 3
    wall = float("inf")
    for rep in range(REPS):
        launches0 = core.launch_count()
        barrier()
        t0 = time.perf_counter()
        all_chains(n_moves, 10 + 1000 * rep)
        w = time.perf_counter() - t0
        if w < wall:
            wall, launches, stats_timed = w, core.launch_count() - launches0, list(stats)
    # the same chains with the list rebuilt only when the device-side trigger fires
    for c in chains:
        c[1].update_repulsion(c[0]); c[2].displacement = 0.0
    wall_lazy = float("inf")
    for rep in range(REPS):
        barrier()
        t0 = time.perf_counter()
        all_chains(n_moves, 40 + 100 * rep, lazy=True)
        wall_lazy = min(wall_lazy, time.perf_counter() - t0)
    rebuilt = sum(u.rebuilds for u in lazy_upd) / max(1, sum(u.rebuilds + u.skipped for u in lazy_upd))
    # ... and with the momenta drawn on the device as well (counter-based generator; the host draw is most of a move)
    for c in chains:
        c[2].device_momenta = 5000 + c[3]
    all_chains(5, 60, lazy=True)
    wall_dev = float("inf")
    for rep in range(REPS):
        barrier()
        t0 = time.perf_counter()
        all_chains(n_moves, 70 + 100 * rep, lazy=True)
        wall_dev = min(wall_dev, time.perf_counter() - t0)
    for c in chains:
        c[2].device_momenta = None
    stats[:] = stats_timed
    # one chain alone, for the concurrency gain
    t0 = time.perf_counter()
    run(0, 10, 30)
    solo = (time.perf_counter() - t0) / 10
    e_last = chains[0][1].energy(chains[0][0])
    assert np.isfinite(e_last)
    return {"wall_s": wall, "moves": n_moves, "L": L, "R": R, "V": V, "solo_ms_per_move": 1e3 * solo,
            "hmc_acceptance": float(np.mean([s[0] / max(s[1], 1) for s in stats])),
            "flips_accepted_per_sweep": float(np.mean([s[2] for s in stats])) / max(1, n_moves // 10),
            "launches": int(launches), "wall_lazy_s": wall_lazy, "lazy_rebuild_fraction": rebuilt, "wall_device_momenta_s": wall_dev}


def hmc_step_run(args, core, workload, rank):
    """BASELINE config 2: the pieces of one full HMC step at 40,962 vertices through the public API
    (trimem_b200.core / trimem_b200.hmc): a 10-step trajectory move with cell-list stress repulsion,
    update_repulsion, and the batch-parallel flip sweep pflip(ratio 0.1) that runs every 10 moves
    (flip statistics on the physical variant with all kappa = 1, as tests/test_flips.py of the reference)."""
    from trimem_b200 import hmc
    out = {}
    mesh, estore, x0 = workload.make_case(core, 6, repulsion="stress", n_search="cell-list")
    V = mesh.n_vertices()
    np.random.seed(5)
    chain = hmc.MeshHMC(mesh, estore, options={"time_step": 1.0e-6, "num_integration_steps": 10, "info_step": 0,
                                               "initial_temperature": 1.0e3}, verbose=False)
    for _ in range(3):
        chain.step()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        chain.step()
    out["hmc_move_ms"] = 1e3 * (time.perf_counter() - t0) / n
    out["hmc_acceptance"] = chain.acc / max(chain.i, 1)
    # the same move with the momenta drawn on the device (opt-in; the host draw of 3 V normal variates is most of a move)
    chain.device_momenta = 4242
    for _ in range(3):
        chain.step()
    t0 = time.perf_counter()
    for _ in range(n):
        chain.step()
    out["hmc_move_ms_device_momenta"] = 1e3 * (time.perf_counter() - t0) / n
    chain.device_momenta = None
    estore.update_repulsion(mesh)
    t0 = time.perf_counter()
    for _ in range(5):
        estore.update_repulsion(mesh)
    out["update_repulsion_cell_list_ms"] = 1e3 * (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    e = estore.energy(mesh); g = estore.gradient(mesh)
    out["eg_cell_list_ms_host_buffers"] = 1e3 * (time.perf_counter() - t0)
    # flips
    fmesh, festore, _ = workload.make_case(core, 6, repulsion="physical", n_search="cell-list",
                                           kappa=(1.0, 1.0, 1.0, 1.0, 1.0, 1.0), fracs=(1.0, 1.0, 1.0))
    core.pflip(fmesh, festore, 0.1, seed=1)
    acc, ts = [], []
    for k in range(5):
        t0 = time.perf_counter()
        acc.append(core.pflip(fmesh, festore, 0.1, seed=10 + k))
        ts.append(1e3 * (time.perf_counter() - t0))
    out["pflip_ms"] = float(np.median(ts))
    out["pflip_ms_min"] = float(np.min(ts))
    out["pflip_proposals"] = int(fmesh.n_edges() * 0.1)
    out["pflip_acceptance"] = float(np.mean(acc)) / out["pflip_proposals"]
    t0 = time.perf_counter()
    a = core.flip(fmesh, festore, 0.1, seed=3)
    out["flip_serial_ms"] = 1e3 * (time.perf_counter() - t0)
    out["flip_serial_acceptance"] = a / out["pflip_proposals"]
    out["what"] = (f"icosphere subdiv=6 ({V} vertices), public API with host buffers: MeshHMC.step() = 10 velocity-Verlet steps "
                   "on the device + Metropolis (stress repulsion, cell-list, excl 2); update_repulsion; "
                   "pflip / flip over 10 % of the edges on the physical variant with kappa = 1")
    out["mvertex_steps_per_s_hmc_move"] = V * 10 / (out["hmc_move_ms"] * 1e-3) / 1e6
    return out


def partitioned_run(args, core, cabi, workload, rank, world, local, barrier, dist, torch):
    """BASELINE config 4: ONE mesh (subdiv --part-subdiv) whose vertices are partitioned across the ranks
    (csrc/tm_part.cu): E+G evals/s and leapfrog steps/s of the whole mesh, strong scaling. At one rank the
    same mesh is evaluated unpartitioned (the base of the scaling curve)."""
    from trimem_b200 import meshgen, partition
    t0 = time.time()
    v, f = meshgen.jittered_icosphere(args.part_subdiv)
    mesh = core.TriMesh(v, f)
    ep = workload.energy_params(core, mesh, repulsion="stress", n_search="verlet-list")
    estore = core.EnergyManager(mesh, ep)            # whole mesh on this device: reference sums + a parity value
    e_single = estore.energy(mesh)
    par = workload.abi_params(estore)
    rp = ep.repulse_params
    arrs = mesh.halfedge_arrays()[:4]
    V = mesh.n_vertices()
    cm = cabi.Mesh(*arrs, xyz=v, device=local)
    make_nl = lambda m: cabi.NList(m, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
    if world > 1:
        def allgather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        nl = partition.setup(cm, rank, world, allgather, make_nl)
        info = cm.part_info()
    else:
        nl = make_nl(cm)
        info = {"rank": 0, "nranks": 1, "first": 0, "count": V, "n_ghost1": 0, "n_pull": 0}
    setup_s = time.time() - t0
    steps = max(5, min(args.steps, 20))
    for _ in range(3):
        cm.energy_async(nl, par); cm.gradient_async(nl, par)
    cm.sync()
    barrier()
    eg_ms = 0.0
    for _ in range(steps):
        cm.flush_l2(FLUSH_BYTES)
        cm.timer_start()
        cm.energy_async(nl, par); cm.gradient_async(nl, par)
        eg_ms += cm.timer_stop()
    barrier()
    e_part, _ = cm.fetch_energy(par)
    # parity with the REFERENCE on this mesh (tests/golden/ico10_stress_verlet2.npz: the reference's six sums, energy
    # and 4,096 sampled gradient rows, tests/golden/make_golden.py big): every rank checks the sampled rows it owns
    gold_path = os.path.join(ROOT, "tests", "golden", f"ico{args.part_subdiv}_stress_verlet2.npz")
    gold_cmp = None
    if os.path.exists(gold_path):
        gold = np.load(gold_path)
        g_own = cm.fetch_gradient()
        lo, hi = info["first"], info["first"] + info["count"]
        d2h = cm.vertex_order()                       # device slot -> host vertex
        owned = np.zeros(V, bool); owned[d2h[lo:hi]] = True
        rows = gold["rows"][owned[gold["rows"]]]
        want = gold["grad_rows"][owned[gold["rows"]]]
        num = float(np.sum((g_own[rows] - want) ** 2)); den = float(np.sum(want ** 2))
        gold_cmp = {"energy_rel_diff": abs(e_part - float(gold["energy"])) / abs(float(gold["energy"])),
                    "grad_rows_sq_err": num, "grad_rows_sq_norm": den, "rows_checked": int(len(rows))}
    # leapfrog
    L = max(1, args.leapfrog_steps)
    lf_dt = 2.0e-8 * (655362.0 / V)
    p_host = np.random.default_rng(7).normal(size=v.shape)

    def lf_reset():
        cm.set_positions(v)
        cm.set_momenta(p_host)
        cm.sync()
    for _ in range(2 if world > 1 else 5):      # unpartitioned: both record-array parities direct, captured, replayed
        lf_reset()
        barrier()
        cm.vv_integrate_async(nl, par, 1.0, lf_dt, L)
        cm.sync()
    barrier()
    n_traj = 3
    lf_ms = 0.0
    for _ in range(n_traj):
        lf_reset()
        barrier()
        cm.flush_l2(FLUSH_BYTES)
        cm.timer_start()
        cm.vv_integrate_async(nl, par, 1.0, lf_dt, L)
        lf_ms += cm.timer_stop()
    lf_reset()
    barrier()
    cm.vv_integrate_async(nl, par, 1.0, lf_dt, L, True)
    st = cm.vv_stats(par)[0]
    kin = torch.tensor([st[1], st[3]], dtype=torch.float64, device="cuda")     # kinetic energy: owned shares
    if world > 1:
        dist.all_reduce(kin, op=dist.ReduceOp.SUM)
    h0, h1 = st[0] + float(kin[0]), st[2] + float(kin[1])
    assert np.isfinite([h0, h1]).all(), "partitioned leapfrog trajectory diverged"
    # list rebuild on current positions (collective: pulls the far field first); second build timed (warm pool)
    make_nl(cm).close(); cm.sync()
    barrier()
    t1 = time.time()
    nl2 = make_nl(cm); cm.sync()
    rebuild_ms = 1e3 * (time.time() - t1)
    t = torch.tensor([eg_ms, lf_ms, rebuild_ms], dtype=torch.float64, device="cuda")
    halo = torch.tensor([info["n_ghost1"], info["n_pull"], info["count"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(halo, op=dist.ReduceOp.MAX)
    eg_ms, lf_ms, rebuild_ms = float(t[0]), float(t[1]), float(t[2])
    nl2.close(); nl.close(); cm.close()
    if gold_cmp is not None:
        gt = torch.tensor([gold_cmp["grad_rows_sq_err"], gold_cmp["grad_rows_sq_norm"], gold_cmp["rows_checked"]],
                          dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(gt, op=dist.ReduceOp.SUM)
        gold_cmp = {"energy_rel_diff_vs_reference": gold_cmp["energy_rel_diff"],
                    "grad_rows_rel_diff_vs_reference": float(np.sqrt(float(gt[0]) / float(gt[1]))),
                    "rows_checked": int(gt[2]), "golden": os.path.relpath(gold_path, ROOT)}
        assert gold_cmp["energy_rel_diff_vs_reference"] < 1e-10 and gold_cmp["grad_rows_rel_diff_vs_reference"] < 1e-10, gold_cmp
    # the same partition through the drop-in classes: core.TriMesh.part_* / EnergyManager.part_prepare, then one
    # MeshHMC move (collective; every rank seeds its chain identically)
    from trimem_b200 import hmc
    api = None
    try:
        if world > 1:
            partition.setup_core(mesh, estore, rank, world, allgather)
        chain = hmc.MeshHMC(mesh, estore, options={"time_step": lf_dt, "num_integration_steps": L, "info_step": 0},
                            verbose=False, rng=np.random.RandomState(4242))
        chain.step()
        barrier()
        t1 = time.perf_counter()
        chain.step()
        barrier()
        api = {"hmc_move_ms": 1e3 * (time.perf_counter() - t1), "L": L, "accepted": bool(chain.last[1]),
               "what": "MeshHMC.step() through trimem_b200.core on the partitioned mesh, host buffers in and out"}
    except Exception as exc:            # noqa: BLE001  (reported, not fatal for the device-resident numbers)
        api = {"error": repr(exc)[:300]}
    return {"workload": f"ONE icosphere subdiv={args.part_subdiv} ({V} vertices), tether + stress repulsion, verlet-list excl 2, "
                        f"vertices partitioned over {world} rank(s)" + ("" if world > 1 else " (unpartitioned base of the scaling curve)"),
            "scaling": "strong", "n_vertices": V, "ranks": world,
            "ms_per_eg": eg_ms / steps, "evals_per_s": steps / (eg_ms / 1e3),
            "mvertex_evals_per_s": steps / (eg_ms / 1e3) * V / 1e6,
            "leapfrog_ms_per_step": lf_ms / n_traj / L, "mvertex_steps_per_s": n_traj * L / (lf_ms / 1e3) * V / 1e6,
            "verlet_rebuild_ms": rebuild_ms,
            "hamiltonian": {"H0": h0, "H1": h1, "time_step": lf_dt},
            "max_owned": int(halo[2]), "max_ghost_ring1": int(halo[0]), "max_pulled_per_update": int(halo[1]),
            "energy_rel_diff_vs_whole_mesh_on_one_device": abs(e_part - e_single) / abs(e_single),
            "reference_golden": gold_cmp, "public_api": api,
            "exchange": "halo positions pulled over NVLink from the owners' arrays (cudaIpc), six sums added in rank order; "
                        "device-side sequencing, no NCCL / host on the data path" if world > 1 else "none",
            "setup_s": setup_s}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--subdiv", type=int, default=8, help="icosphere subdivisions (8 = 655,362 vertices, the headline)")
    ap.add_argument("--cpu-steps", type=int, default=20, help="E+G evaluations of the bounded cpu_baseline sample (~10 s of host work with the setup)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--part-subdiv", type=int, default=10,
                    help="icosphere subdivisions of the ONE mesh partitioned across the ranks (BASELINE config 4: 10 = "
                         "10,485,762 vertices); also evaluated unpartitioned at --gpus 1 as the strong-scaling base")
    ap.add_argument("--no-partition", action="store_true", help="skip the partitioned-mesh measurement")
    ap.add_argument("--replicas", type=int, default=8,
                    help="independent 40,962-vertex membranes per GPU in the replica measurement (BASELINE config 5; 0 = skip)")
    ap.add_argument("--no-hmc-step", action="store_true", help="skip the config-2 HMC-step measurement")
    ap.add_argument("--leapfrog-steps", type=int, default=10, help="velocity-Verlet steps per timed trajectory")
    args = ap.parse_args()

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    from trimem_b200 import workload
    V = workload.N_VERTICES.get(args.subdiv, None)
    wl = f"icosphere subdiv={args.subdiv} ({V} vertices) E+G: energy()+gradient(), tether + stress repulsion, verlet-list excl 2"

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, args.steps)
        # bounded: one E+G takes ~1.4 s on 8 cores at 655k vertices; cap the run at ~3 minutes
        probe_budget = 150.0
        r = time_reference(args.subdiv, 1, 1)
        steps = max(1, min(steps, int(probe_budget / max(r["ms_per_step"] / 1e3, 1e-3))))
        if steps > 1:
            r = time_reference(args.subdiv, steps, max(0, args.warmup))
        line = {"impl": "reference", "metric": "energy+gradient evals/s", "value": r["value"], "unit": "evals/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": max(0, args.warmup), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl, "n_vertices": V, "timing": "host wall clock, working set (~70 MB) exceeds CPU caches"},
                "cpu_baseline": {"value": r["value"], "unit": "evals/s", "cores": r["cores"], "kind": r["kind"],
                                 "sample": f"{steps} E+G evaluations of the same {V}-vertex workload, OMP_NUM_THREADS={r['cores']}"},
                "e2e": {"value": r["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "mvertex_evals_per_s": r["value"] * V / 1e6}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------------ our arm
    bound_cpus = None
    orig_affinity = os.sched_getaffinity(0)
    if True:
        # one process per GPU: run on (and first-touch the pinned buffers from) the cores next to this rank's
        # GPU; the end-to-end path is host-memory- and PCIe-bound, so locality matters (a single rank takes all the
        # cores next to its GPU and gives the process its original cores back before the CPU baseline leg)
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(ClockSampler._nvml_index(local))
            words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
            allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
            if allowed:
                # spread the ranks that share a socket over its cores
                per = max(2, len(allowed) // max(1, min(world, 8)))
                mine = allowed[(local * per) % len(allowed):][:per] or allowed
                os.sched_setaffinity(0, mine)
                bound_cpus = len(mine)
        except Exception:
            bound_cpus = None
    import torch
    import torch.distributed as dist
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's own banner ("NCCL version ...") off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE"):
            os.environ["NCCL_DEBUG_FILE"] = os.environ.get("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    import trimem_b200.core as core
    from trimem_b200 import cabi
    if core.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    core.set_device(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the public-API objects (e2e path) ...
    t0 = time.time()
    mesh, estore, x0 = workload.make_case(core, args.subdiv, repulsion="stress", n_search="verlet-list")
    setup_s = time.time() - t0
    V = mesh.n_vertices()
    # ... and a C-ABI mirror of the same mesh for the device-resident loop
    he_to, he_next, he_face, v_he, _ = mesh.halfedge_arrays()
    cm = cabi.Mesh(he_to, he_next, he_face, v_he, xyz=x0, device=local)
    rp = estore.eparams.repulse_params
    t0 = time.time()
    nl = cabi.NList(cm, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
    cm.sync()
    rebuild_first_ms = 1e3 * (time.time() - t0)
    n_entries = nl.info()[0]
    P = n_entries / V
    p = workload.abi_params(estore)

    def eg_device():
        cm.energy_async(nl, p)
        cm.gradient_async(nl, p)

    for _ in range(max(args.warmup, 3)):
        eg_device()
    cm.sync()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = cabi.lib().tm_launch_count()
    barrier()
    total_ms = 0.0
    for _ in range(args.steps):
        cm.flush_l2(FLUSH_BYTES)          # evict L2 (outside the timed span)
        cm.timer_start()
        eg_device()
        total_ms += cm.timer_stop()
    barrier()
    launches = cabi.lib().tm_launch_count() - launches0

    # sweep sets timed alone, L2 flushed before every launch: energy() sweeps (ring properties +
    # pair sweep + reduction), the gradient's own property sweep, the ring gradient kernel
    # (dominant single kernel), the pair sweep
    k_ms = [cm.time_kernel(nl, p, w, max(10, min(args.steps, 50)), FLUSH_BYTES) for w in (0, 1, 2, 3)]
    # verlet-list rebuild (update_repulsion), device resident
    reb = []
    for _ in range(3):
        cm.sync(); t0 = time.time()
        nl2 = cabi.NList(cm, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level); cm.sync()
        reb.append(1e3 * (time.time() - t0)); nl2.close()

    # ---- e2e through the public API with host buffers (H2D of positions, D2H of the gradient)
    xs = [x0 + 1e-6 * np.random.default_rng(s).normal(size=x0.shape) for s in range(2)]
    phase = [0.0, 0.0, 0.0]
    def eg_api(k):
        t_a = time.perf_counter()
        np.copyto(mesh.points(), xs[k % 2])
        t_b = time.perf_counter()
        e = estore.energy(mesh)
        t_c = time.perf_counter()
        g = estore.gradient(mesh)
        t_d = time.perf_counter()
        phase[0] += t_b - t_a; phase[1] += t_c - t_b; phase[2] += t_d - t_c
        return e, g
    for k in range(4):
        e_last, g_last = eg_api(k)          # same ownership pattern as the timed loop (two result arrays circulate)
    e2e_steps = max(5, min(args.steps, 30))
    barrier()
    phase[:] = [0.0, 0.0, 0.0]
    t0 = time.time()
    for k in range(e2e_steps):
        e_last, g_last = eg_api(k)
    barrier()
    e2e_s = time.time() - t0
    e2e_phase_ms = [1e3 * t / e2e_steps for t in phase]

    # ... the same step with the opt-in TriMesh.set_points(x) (one threaded copy + one upload, no re-upload inside
    # energy()/gradient()): what a caller gains by leaving the reference's idiom
    phase_sp = [0.0, 0.0, 0.0]
    def eg_set_points(k):
        t_a = time.perf_counter()
        mesh.set_points(xs[k % 2])
        t_b = time.perf_counter()
        e = estore.energy(mesh)
        t_c = time.perf_counter()
        g = estore.gradient(mesh)
        t_d = time.perf_counter()
        phase_sp[0] += t_b - t_a; phase_sp[1] += t_c - t_b; phase_sp[2] += t_d - t_c
        return e, g
    for k in range(4):
        e_sp, g_sp = eg_set_points(k)
    barrier()
    phase_sp[:] = [0.0, 0.0, 0.0]
    t0 = time.time()
    for k in range(e2e_steps):
        e_sp, g_sp = eg_set_points(k)
    barrier()
    e2e_sp_s = time.time() - t0
    e2e_sp_phase_ms = [1e3 * t / e2e_steps for t in phase_sp]
    sp_equal = bool(e_sp == e_last and np.array_equal(g_sp, g_last))     # (reported, not asserted: a bench line either way)
    # ... and what the host side of this box gives: one core's copy rate into the mesh's pinned memory, and the PCIe
    # rates of the position upload / gradient download (pinned, torch copies): the floor of the drop-in step is
    # copy + upload + E+G + download
    V3 = x0.size * 8
    t0 = time.perf_counter()
    for k in range(6):
        np.copyto(mesh.points(), xs[k % 2])
    host_copy_gbs = 6 * V3 / (time.perf_counter() - t0) / 1e9

    # ---- device-resident velocity-Verlet trajectory (tm_vv_integrate_async; hmc.py:14-26): L steps per
    # timed trajectory = L+1 gradient evaluations + L fused update kernels, state never leaves HBM
    L = max(1, args.leapfrog_steps)
    lf_dt = 2.0e-8                      # stable for the stiff stress potential at this resolution
    p_host = np.random.default_rng(100 + rank).normal(size=x0.shape)

    def lf_reset():                     # every trajectory starts from (x0, p0): no drift across trajectories
        cm.set_positions(x0)
        cm.set_momenta(p_host)
        cm.sync()
    for _ in range(6):                  # both record-array parities: direct launch, graph capture, first replay
        lf_reset()
        cm.vv_integrate_async(nl, p, 1.0, lf_dt, L)
    cm.sync()
    n_traj = max(3, min(10, args.steps // 5))
    barrier()
    lf_ms = 0.0
    for _ in range(n_traj):
        lf_reset()
        cm.flush_l2(FLUSH_BYTES)
        cm.timer_start()
        cm.vv_integrate_async(nl, p, 1.0, lf_dt, L)
        lf_ms += cm.timer_stop()
    barrier()
    clocks = sampler.stop()       # covers the device-resident loops, the per-kernel timings and the e2e loop
    lf_reset()
    cm.vv_integrate_async(nl, p, 1.0, lf_dt, L, True)
    st = cm.vv_stats(p)[0]
    lf_stats = {"H0": st[0] + st[1], "H1": st[2] + st[3], "time_step": lf_dt}
    assert np.isfinite(st).all(), "leapfrog trajectory diverged"
    cm.set_positions(x0)

    # ---- one HMC move (L leapfrog steps + both Hamiltonian evaluations + Metropolis) through the public API with
    # host buffers: what a trimem user's MeshHMC.step() costs end to end at the headline size
    from trimem_b200 import hmc as hmc_mod
    np.copyto(mesh.points(), x0)
    chain = hmc_mod.MeshHMC(mesh, estore, options={"time_step": lf_dt, "num_integration_steps": L, "info_step": 0},
                            verbose=False, rng=np.random.RandomState(7 + rank))
    chain.step(); chain.step()
    barrier()
    t0 = time.perf_counter()
    n_moves_api = 5
    for _ in range(n_moves_api):
        chain.step()
    barrier()
    hmc_api_ms = 1e3 * (time.perf_counter() - t0) / n_moves_api
    # ... and with the momenta drawn on the device (opt-in): the host's draw of 3 V normal variates is most of the move above
    chain.device_momenta = 99 + rank
    chain.step(); chain.step()
    t0 = time.perf_counter()
    for _ in range(n_moves_api):
        chain.step()
    hmc_api_dev_ms = 1e3 * (time.perf_counter() - t0) / n_moves_api
    chain.device_momenta = None
    np.copyto(mesh.points(), x0)

    repl = replica_run(args, core, cabi, workload, rank, world, local, barrier) if args.replicas > 0 else None
    hmc2 = hmc_step_run(args, core, workload, rank) if (rank == 0 and not args.no_hmc_step) else None
    part = None
    if not args.no_partition:
        part = partitioned_run(args, core, cabi, workload, rank, world, local, barrier, dist if world > 1 else None, torch)

    # PCIe rates of this box (torch copies of a pinned buffer of the positions' size); measured here, after the
    # host-thread-heavy sections
    hp = torch.empty(x0.size, dtype=torch.float64).pin_memory()
    dp = torch.zeros(x0.size, dtype=torch.float64, device="cuda")
    pcie = {}
    # (download first: the DMA leaves no dirty lines of the pinned buffer in the CPU caches, which would slow the upload)
    for name, src, dst in (("d2h", dp, hp), ("h2d", hp, dp)):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(6):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        pcie[name] = 6 * V3 / (time.perf_counter() - t0) / 1e9
    del hp, dp

    # parity spot check of the timed configuration against itself through both boundaries
    e_dev, _ = cm.energy(nl, p)

    # max over ranks
    t = torch.tensor([total_ms, e2e_s, lf_ms, repl["wall_s"] if repl else 0.0, hmc_api_ms, e2e_sp_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, lf_ms, repl_s, hmc_api_ms, e2e_sp_s = (float(v) for v in t)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, orig_affinity)
        r = time_reference(args.subdiv, args.cpu_steps, 1)
        cpu = {"value": r["value"], "unit": "evals/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"{args.cpu_steps} E+G evaluations (after 1 warm-up) of the same {V}-vertex workload on the host, "
                         f"OMP_NUM_THREADS={r['cores']}; reference E={r['energy']:.12e} |g|={r['grad_norm']:.12e}",
               "ms_per_step": r["ms_per_step"]}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        ms_step = total_ms / args.steps
        value = world * args.steps / (total_ms / 1e3)
        # ALGORITHMIC bytes per launch, per vertex:
        #   property sweep   x 24 + ring row 28 + rho 8                                  =  60
        #   pair sweep       x 24 + list 4P+4 + pair force 24                            =  52 + 4P
        #   gradient sweep   x 24 + ring row 28 + rho 8 + g 24 (+ pair force 24 read)    = 108
        # (the slot quantity tau the property sweep hands to the gradient sweep, 8 B per ring slot, is an
        #  implementation choice, not algorithmic traffic: it shows up in `traffic`, not in `achieved`)
        alg = {"k_props": 60.0 * V, "k_pairs": (52.0 + 4.0 * P) * V, "k_grad": (workload.bytes_grad_sweep(0.0, False) + 24.0) * V}
        kms = {"k_props": k_ms[1], "k_pairs": k_ms[3], "k_grad": k_ms[2]}
        kernels = {k: {"ms_per_launch": kms[k], "algorithmic_bytes_per_launch": alg[k],
                       "achieved_GBps": alg[k] / (kms[k] * 1e-3) / 1e9, "frac": alg[k] / (kms[k] * 1e-3) / 1e9 / peak,
                       **NCU.get(k, {})} for k in alg}
        dom = max(kms, key=kms.get)                      # the dominant kernel of the step = the longest one
        alg_eg = workload.bytes_eg(P, True) * V
        lf_ms_step = lf_ms / n_traj / L
        # 1 gradient() (its property sweep + its gradient sweep, the list read in both) + the x, p update
        alg_lf = (workload.bytes_grad_props(P, True) + workload.bytes_grad_sweep(P, True) + 96.0) * V
        part_compact = None if not part else {k: part[k] for k in ("n_vertices", "ranks", "ms_per_eg", "leapfrog_ms_per_step",
                                                                   "verlet_rebuild_ms") if k in part}
        line = {
            "metric": "energy+gradient evals/s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl, "n_vertices": V, "verlet_entries": int(n_entries), "entries_per_vertex": P,
                       "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
                       "host_cores_bound_per_rank": bound_cpus,
                       "l2": f"flushed before every timed step ({FLUSH_BYTES >> 20} MiB memset, outside the timed span)",
                       # strong scaling of ONE partitioned mesh (BASELINE config 4), in short; details under "partitioned"
                       "partitioned": part_compact},
            "clocks": clocks,
            "e2e": {"value": world * e2e_steps / e2e_sp_s, "unit": "evals/s",
                    "h2d_bytes_per_step": V * 24, "d2h_bytes_per_step": V * 24 + 2 * 48,
                    "ms_per_step": 1e3 * e2e_sp_s / e2e_steps, "steps": e2e_steps,
                    "what": "per step, through trimem_b200.core with host buffers: mesh.set_points(x) (x a host array; copy into the "
                            "mesh's pinned memory + ONE upload), estore.energy(mesh), estore.gradient(mesh) (gradient downloaded into a "
                            "host array); set_points is this repo's opt-in extension (INTEGRATION.md section C), the reference's own "
                            "idiom through the same module is under reference_idiom",
                    "bit_identical_to_reference_idiom": sp_equal,
                    "phase_ms": {"mesh.set_points(x)": e2e_sp_phase_ms[0], "estore.energy(mesh)": e2e_sp_phase_ms[1],
                                 "estore.gradient(mesh)": e2e_sp_phase_ms[2]},
                    "reference_idiom": {"what": "np.copyto(mesh.points(), x); estore.energy(mesh); estore.gradient(mesh) -- the drop-in path: "
                                                "every evaluation re-uploads the positions and compares them bitwise (round 1's e2e figure)",
                                        "value": world * e2e_steps / e2e_s, "unit": "evals/s", "ms_per_step": 1e3 * e2e_s / e2e_steps,
                                        "h2d_bytes_per_step": 2 * V * 24, "d2h_bytes_per_step": V * 24 + 2 * 48,
                                        "phase_ms": {"np.copyto(mesh.points(), x)": e2e_phase_ms[0], "estore.energy(mesh)": e2e_phase_ms[1],
                                                     "estore.gradient(mesh)": e2e_phase_ms[2]}},
                    "host": {"copy_into_pinned_GBps_one_core": host_copy_gbs, "h2d_GBps": pcie["h2d"], "d2h_GBps": pcie["d2h"]},
                    "hmc_move": {"what": f"MeshHMC.step() through trimem_b200.core / trimem_b200.hmc: {L} leapfrog steps on the device, "
                                         "both Hamiltonian evaluations, Metropolis; momenta and positions in, end point and momenta out",
                                 "ms_per_move": hmc_api_ms, "ms_per_move_device_momenta_this_rank": hmc_api_dev_ms, "h2d_bytes_per_move": 2 * V * 24, "d2h_bytes_per_move": 2 * V * 24 + 32,
                                 "mvertex_steps_per_s": world * L * V / (hmc_api_ms * 1e-3) / 1e6}},
            "leapfrog": {"what": f"device-resident velocity-Verlet trajectory of {L} steps (tm_vv_integrate_async): "
                                 f"{L + 1} gradient evaluations + {L} fused update kernels, L2 flushed before each trajectory",
                         "ms_per_step": lf_ms_step, "steps_per_s": world * n_traj * L / (lf_ms / 1e3),
                         "mvertex_steps_per_s": world * n_traj * L / (lf_ms / 1e3) * V / 1e6,
                         "trajectories": n_traj, "hamiltonian": lf_stats},
            "replicas": None if not repl else {
                "what": f"{repl['R']} independent HMC + flip chains per GPU (BASELINE config 5: {repl['V']}-vertex membranes of the kappa_b x "
                        f"volume_frac sweep, physical repulsion, verlet-list) through trimem_b200.hmc.MeshHMC / core.pflip / "
                        f"update_repulsion, one host thread per chain; per move {repl['L']} leapfrog steps, every 10th move a "
                        "pflip(0.1); host wall clock of the shortest of three runs of 20 moves per chain, max over ranks",
                "chains_total": world * repl["R"], "moves_per_chain": repl["moves"],
                "moves_per_s": world * repl["R"] * repl["moves"] / repl_s,
                "mvertex_steps_per_s": world * repl["R"] * repl["moves"] * repl["L"] / repl_s * repl["V"] / 1e6,
                "ms_per_move_per_chain_concurrent": 1e3 * repl_s / repl["moves"],
                "ms_per_move_one_chain_alone": repl["solo_ms_per_move"],
                "hmc_acceptance": repl["hmc_acceptance"], "flips_accepted_per_sweep": repl["flips_accepted_per_sweep"],
                "gpu_launches": repl["launches"],
                "lazy_repulsion": {"what": "the same chains with hmc.LazyRepulsion: the verlet list is rebuilt only when the "
                                           "displacement measured on the device behind each trajectory exceeds a quarter of "
                                           "rlist - lc1, or flips were accepted (this rank's wall clock)",
                                   "moves_per_s": world * repl["R"] * repl["moves"] / repl["wall_lazy_s"],
                                   "rebuilds_per_move": repl["lazy_rebuild_fraction"],
                                   "moves_per_s_with_device_momenta": world * repl["R"] * repl["moves"] / repl["wall_device_momenta_s"],
                                   "device_momenta": "MeshHMC(device_momenta=seed): momenta of a move drawn on the device "
                                                     "(Philox4x32-10 + Box-Muller, tm_mesh_random_momenta) instead of by numpy"}},
            "hmc_step_config2": hmc2,
            "partitioned": part,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom + " (longest kernel of the step; the sweeps are bound by fp64 instruction issue and by the "
                                                     "latency of a warp's dependent loads, not by HBM: DESIGN.md section 4)",
                         "achieved": kernels[dom]["achieved_GBps"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac"],
                         "traffic": NCU.get(dom, {}).get("dram_bytes_per_launch") if args.subdiv == 8 else None,
                         "traffic_source": NCU_SOURCE, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg[dom], "ms_per_launch": kms[dom],
                         "kernels": kernels,
                         # whole step: energy() + gradient() (one property sweep, one pair sweep, one gradient sweep: the
                         # gradient reuses what the energy evaluated at the same positions) ...
                         "eg_algorithmic_bytes": alg_eg, "eg_frac_of_peak": alg_eg / (ms_step * 1e-3) / 1e9 / peak,
                         # ... and the production figure: one leapfrog step (gradient only, nothing cached)
                         "leapfrog_algorithmic_bytes": alg_lf, "leapfrog_frac_of_peak": alg_lf / (lf_ms_step * 1e-3) / 1e9 / peak,
                         "kernel_ms": {"energy_sweeps": k_ms[0], "props_for_gradient": k_ms[1], "ring_grad": k_ms[2],
                                       "pair_sweep": k_ms[3]}},
            "cpu_baseline": cpu,
            "mvertex_evals_per_s": value * V / 1e6,
            "verlet_rebuild_ms": float(np.median(reb)),
            "setup_s": setup_s,
            "check": {"energy_device": e_dev, "energy_api_last": float(e_last), "grad_norm_api_last": float(np.linalg.norm(g_last))},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
