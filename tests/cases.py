"""Parity cases shared by tests/golden/make_golden.py (which records the reference's
answers) and the tests (which compare the oracle port and the CUDA path with them)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from trimem_b200 import meshgen, workload  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
K1 = (1.0, 1.0, 1.0, 1.0, 1.0, 0.0)
K1R = (1.0, 1.0, 1.0, 1.0, 1.0, 1.0)

# name -> (mesh generator, args, workload.energy_params kwargs)
CASES = {
    # reference tests/test_gradient.py configuration, both bond types
    "ico3_edge": ("icosphere", (3,), dict(kappa=K1, fracs=(0.8, 0.8, 0.8), bond="Edge", n_search="cell-list")),
    "ico3_area": ("icosphere", (3,), dict(kappa=K1, fracs=(0.8, 0.8, 0.8), bond="Area", n_search="cell-list")),
    # BASELINE config 1: subdiv 4, no repulsion
    "ico4_config1": ("jittered_icosphere", (4,), dict(repulsion="none", n_search="cell-list")),
    # stress repulsion, every list type x exclusion level
    "ico4_stress_cell0": ("jittered_icosphere", (4,), dict(repulsion="stress", n_search="cell-list", exclusion=0)),
    "ico4_stress_cell1": ("jittered_icosphere", (4,), dict(repulsion="stress", n_search="cell-list", exclusion=1)),
    "ico4_stress_cell2": ("jittered_icosphere", (4,), dict(repulsion="stress", n_search="cell-list", exclusion=2)),
    "ico4_stress_verlet0": ("jittered_icosphere", (4,), dict(repulsion="stress", n_search="verlet-list", exclusion=0)),
    "ico4_stress_verlet1": ("jittered_icosphere", (4,), dict(repulsion="stress", n_search="verlet-list", exclusion=1)),
    "ico4_stress_verlet2": ("jittered_icosphere", (4,), dict(repulsion="stress", n_search="verlet-list", exclusion=2)),
    "ico4_physical_cell2": ("jittered_icosphere", (4,), dict(repulsion="physical", n_search="cell-list", exclusion=2)),
    "ico4_area_stress": ("jittered_icosphere", (4,), dict(repulsion="stress", n_search="verlet-list", bond="Area")),
    # BASELINE config 2 (40,962 vertices) and config 3 (655,362 vertices, headline)
    "ico6_stress_cell2": ("jittered_icosphere", (6,), dict(repulsion="stress", n_search="cell-list", exclusion=2)),
    "ico6_stress_verlet2": ("jittered_icosphere", (6,), dict(repulsion="stress", n_search="verlet-list", exclusion=2)),
    "ico8_stress_verlet2": ("jittered_icosphere", (8,), dict(repulsion="stress", n_search="verlet-list", exclusion=2)),
    "ico8_norep": ("jittered_icosphere", (8,), dict(repulsion="none", n_search="verlet-list", exclusion=2)),
}
# BASELINE config 4 (10,485,762 vertices): golden made on request only (make_golden.py big; ~5 minutes of the
# reference on 8 threads), never built by the CPU tests (BIG_CASES are kept out of CASES)
BIG_CASES = {
    "ico10_stress_verlet2": ("jittered_icosphere", (10,), dict(repulsion="stress", n_search="verlet-list", exclusion=2)),
}
N_SAMPLED_ROWS_BIG = 4096
FULL_GRADIENT_MAX_V = 3000      # cases up to this size store the whole gradient
N_SAMPLED_ROWS = 512


def mesh_of(case):
    gen, args, _ = CASES[case]
    if gen.startswith("open:"):
        return open_mesh(gen[5:])
    return getattr(meshgen, gen)(*args)


def build(mod, case):
    """(mesh, estore) of `case` on module `mod` (reference build or trimem_b200.core)."""
    v, f = mesh_of(case)
    mesh = mod.TriMesh(v, f)
    ep = workload.energy_params(mod, mesh, **CASES[case][2])
    return mesh, mod.EnergyManager(mesh, ep)


GRAD_TILE = 64


def tile_sums(a, tile=None):
    """(ceil(V / tile), 3): vector sums of consecutive row tiles of a (V, 3) array."""
    tile = tile or GRAD_TILE
    V = len(a)
    pad = (-V) % tile
    b = np.vstack([a, np.zeros((pad, a.shape[1]))]) if pad else a
    return b.reshape(-1, tile, a.shape[1]).sum(axis=1)


def sample_rows(n_vertices, n=None):
    rng = np.random.default_rng(7)
    return np.sort(rng.choice(n_vertices, size=min(n or N_SAMPLED_ROWS, n_vertices), replace=False))


def props6(p):
    return np.array([p.area, p.volume, p.curvature, p.bending, p.tethering, p.repulsion])


def load_golden(name):
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    if not os.path.exists(path):
        return None
    return dict(np.load(path, allow_pickle=False))


# flip sweeps with injected randomness: name -> (case, n_proposals, seed)
FLIP_CASES = {
    "flip_ico3_k1": ("ico3_flip", 960, 1),
    "flip_ico4_cell": ("ico4_flip_cell", 768, 2),
    "flip_ico4_verlet": ("ico4_flip_verlet", 768, 3),
    # an open mesh: proposals on boundary edges, patches that contain boundary vertices
    "flip_rect_open": ("rect_flip", 600, 4),
}
CASES.update({
    # reference tests/test_flips.py configuration
    "ico3_flip": ("icosphere", (3,), dict(kappa=K1, fracs=(1.0, 1.0, 1.0), n_search="cell-list")),
    "ico4_flip_cell": ("jittered_icosphere", (4,), dict(kappa=K1R, fracs=(1.0, 1.0, 1.0), repulsion="stress",
                                                        n_search="cell-list", exclusion=2)),
    "ico4_flip_verlet": ("jittered_icosphere", (4,), dict(kappa=K1R, fracs=(1.0, 1.0, 1.0), repulsion="stress",
                                                          n_search="verlet-list", exclusion=2)),
    "rect_flip": ("open:rect", (), dict(kappa=K1R, fracs=(1.0, 1.0, 1.0), repulsion="physical", n_search="cell-list",
                                        exclusion=2)),
})


# sweep-level flip statistics against the reference (tests/golden/make_golden.py flip_statistics):
# small mesh: (independent chains, sweeps per chain, flip_ratio); large mesh: (sweeps, flip_ratio,
# [(function, OpenMP threads, chains)])
FLIPSTAT_SMALL = (240, 30, 0.1)
FLIPSTAT_LARGE = (100, 0.1, [("flip", 1, 2), ("pflip", 8, 3)])
CASES["ico6_flip"] = ("jittered_icosphere", (6,), dict(kappa=K1R, fracs=(1.0, 1.0, 1.0), repulsion="physical",
                                                       n_search="cell-list", exclusion=2))

# checkpoint hand-off (tests/golden/make_golden.py restart_case): flips, continuation state, displacement
RESTART_FLIPS = "flip_ico4_verlet"
RESTART_CONT = (0.25, 0.0)              # continuation delta, initial lambda (two updates -> 0.5)
RESTART_FRACS = (0.9, 0.8, 0.95)


def restart_displacement(x):
    return x + 0.004 * np.sin(3.0 * x[:, [1, 2, 0]]) * np.linalg.norm(x, axis=1, keepdims=True)


def flip_inputs(name):
    case, n, seed = FLIP_CASES[name]
    v, f = mesh_of(case)
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    n_edges = len(np.unique(e[:, 0].astype(np.int64) * len(v) + e[:, 1]))      # = 3V - 6 on a closed sphere
    rng = np.random.default_rng(seed)
    return case, rng.integers(0, n_edges, n).astype(np.int32), rng.random(n)


# velocity-Verlet trajectories (reference: _vv_integration, src/trimem/mc/hmc.py:14-26):
# name -> (case, momentum seed, momentum scale, mass, time step, steps)
VV_CASES = {
    "vv_ico4_config1": ("ico4_config1", 11, 1.0, 1.0, 2.5e-5, 10),
    "vv_ico4_stress_verlet2": ("ico4_stress_verlet2", 12, 1.0, 1.0, 2.5e-5, 10),
    "vv_ico4_stress_cell2": ("ico4_stress_cell2", 13, 2.0, 2.0, 2.0e-5, 6),
    "vv_ico3_area": ("ico3_area", 14, 1.0, 1.0, 1.0e-3, 8),
    # an open mesh: the gradient's boundary terms, the update as a kernel of its own
    "vv_rect_open": ("rect_flip", 15, 0.2, 1.0, 4.0e-4, 8),
}


def vv_inputs(name):
    case, seed, scale, mass, dt, n = VV_CASES[name]
    v, _ = mesh_of(case)
    p0 = scale * np.random.default_rng(seed).normal(size=v.shape)
    return case, p0, mass, dt, n


# HMC chains (reference: HMC.step, src/trimem/mc/hmc.py:133-190): name -> (case, numpy seed, options, steps)
HMC_CASES = {
    "hmc_ico3_edge": ("ico3_edge", 21, {"time_step": 2.0e-3, "num_integration_steps": 5, "info_step": 0}, 6),
    "hmc_ico4_stress_verlet2": ("ico4_stress_verlet2", 22, {"time_step": 1.0e-5, "num_integration_steps": 8,
                                                             "initial_temperature": 50.0, "info_step": 0}, 5),
    "hmc_ico3_cooling": ("ico3_area", 23, {"time_step": 2.0e-3, "num_integration_steps": 4, "info_step": 0,
                                           "cooling_factor": 0.5, "cooling_start_step": 1, "mass": 2.0}, 5),
}


# open / irregular meshes (boundary rings, a valence-12 vertex) with a parametrisation that exercises every
# term, for the gradient near boundaries
def fan_mesh(n_spokes=12):
    ang = np.linspace(0, 2 * np.pi, n_spokes, endpoint=False)
    ring1 = np.column_stack([np.cos(ang), np.sin(ang), 0.15 * np.cos(3 * ang)])
    ring2 = 2.0 * np.column_stack([np.cos(ang + np.pi / n_spokes), np.sin(ang + np.pi / n_spokes), 0.1 * np.sin(2 * ang)])
    v = np.vstack([[0.0, 0.0, 0.3], ring1, ring2])
    f = []
    for k in range(n_spokes):
        a, b = 1 + k, 1 + (k + 1) % n_spokes
        f.append([0, a, b])
        c = 1 + n_spokes + k
        f.append([a, c, b])
        f.append([b, c, 1 + n_spokes + (k + 1) % n_spokes])
    return v, np.array(f, np.int32)


OPEN_MESHES = ("tube", "rect", "fan")


def open_mesh(name):
    if name == "tube":
        v, f = meshgen.tube(2.0, 0.7, 24)
        # (off the exactly flat quad diagonals, where the reference's acos cannot resolve the angle)
        return v + 0.01 * np.random.default_rng(6).normal(size=v.shape), f
    if name == "rect":
        v, f = meshgen.rect(12, 9)
        return v + 0.02 * np.random.default_rng(4).normal(size=v.shape), f
    return fan_mesh()


def open_mesh_params(mod, mesh, bond="Edge"):
    ep = mod.EnergyParams()
    ep.kappa_b, ep.kappa_a, ep.kappa_v, ep.kappa_c, ep.kappa_t, ep.kappa_r = 3.0, 2.0, 1.5, 1.0, 0.5, 0.7
    a, l = mod.avg_tri_props(mesh)
    bp = ep.bond_params
    bp.lc0, bp.lc1, bp.a0 = 1.05 * l, 0.95 * l, a
    if bond != "Edge":
        bp.type = getattr(mod.BondType, bond)
    ep.bond_params = bp
    rp = ep.repulse_params
    rp.lc1, rp.rlist, rp.exclusion_level, rp.n_search = 2.2 * l, 2.5 * l, 1, "verlet-list"
    ep.repulse_params = rp
    ep.area_frac = ep.volume_frac = ep.curvature_frac = 0.9
    return ep
