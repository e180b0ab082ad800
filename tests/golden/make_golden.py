#!/usr/bin/env python
"""Generate the golden fixtures from the REFERENCE (oracle/_ref = the unmodified reference
sources compiled against oracle/omshim; `make -C oracle ref`). Runs only in the build
container (needs /root/reference to have been compiled); the .npz files it writes are
committed so that the GPU box, where the reference does not exist, can check against them.

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py
"""
import os
import sys

os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import trimem.core as ref          # noqa: E402
import refprobe                    # noqa: E402
import cases                       # noqa: E402
from trimem_b200 import meshgen, workload    # noqa: E402


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print("wrote", name, {k: np.shape(v) for k, v in arrays.items()})


def energy_cases():
    for name in cases.CASES:
        if name.endswith(("_flip", "_flip_cell", "_flip_verlet")):
            continue
        mesh, es = cases.build(ref, name)
        V = mesh.n_vertices()
        props = cases.props6(es.properties(mesh))
        energy = es.energy(mesh)
        grad = es.gradient(mesh)
        out = dict(n_vertices=V, props=props, energy=energy, grad_norm=np.linalg.norm(grad),
                   grad_sum=grad.sum(axis=0), initial_props=cases.props6(es.initial_props))
        if V <= cases.FULL_GRADIENT_MAX_V:
            out["grad"] = grad
            out["vprops"] = refprobe.vertex_properties(mesh, es)
        else:
            rows = cases.sample_rows(V)
            out["rows"] = rows
            out["grad_rows"] = grad[rows]
        # neighbour-list content of the case
        nl = ref.make_nlist(mesh, es.eparams)
        rl = es.eparams.repulse_params.rlist
        d, i, j = nl.distance_matrix(mesh, rl)
        i, j = np.asarray(i, np.int64), np.asarray(j, np.int64)
        lo, hi = np.minimum(i, j), np.maximum(i, j)
        key = np.unique(lo * V + hi)
        out["n_pairs_undirected"] = len(key)
        out["pair_key_sum"] = np.array([int(key.sum() % (2 ** 61 - 1))], dtype=np.int64)
        if V <= cases.FULL_GRADIENT_MAX_V:
            out["pair_keys"] = key
        save(name, **out)


def gradient_tile_sums():
    """For the cases too large to store the whole gradient: the vector sum of every tile of cases.GRAD_TILE
    consecutive rows (host numbering) of the reference's gradient, so that EVERY row of the device's gradient is
    checked against the reference through exactly one tile, not only the sampled rows. Stored beside the case's
    other fixtures as <case>_gtiles.npz."""
    for name in cases.CASES:
        if name.endswith(("_flip", "_flip_cell", "_flip_verlet")):
            continue
        mesh, es = cases.build(ref, name)
        V = mesh.n_vertices()
        if V <= cases.FULL_GRADIENT_MAX_V:
            continue
        grad = es.gradient(mesh)
        save(name + "_gtiles", tile_sums=cases.tile_sums(grad), tile_abs=cases.tile_sums(np.abs(grad)))


def big_case():
    """BASELINE config 4: the 10,485,762-vertex mesh on the reference (six sums, energy, 4,096 sampled gradient rows,
    gradient sum and norm, pair count). Run with OMP_NUM_THREADS=8; ~5 minutes and ~8 GB."""
    import time
    for name, (gen, args, kw) in cases.BIG_CASES.items():
        t0 = time.time()
        v, f = getattr(meshgen, gen)(*args)
        mesh = ref.TriMesh(v, f)
        es = ref.EnergyManager(mesh, workload.energy_params(ref, mesh, **kw))
        V = mesh.n_vertices()
        print(name, "built", V, "vertices in", time.time() - t0, "s", flush=True)
        props = cases.props6(es.properties(mesh))
        energy = es.energy(mesh)
        grad = es.gradient(mesh)
        rows = cases.sample_rows(V, cases.N_SAMPLED_ROWS_BIG)
        save(name, n_vertices=V, props=props, energy=energy, grad_norm=np.linalg.norm(grad), grad_sum=grad.sum(axis=0),
             rows=rows, grad_rows=grad[rows], initial_props=cases.props6(es.initial_props))
        print(name, "done in", time.time() - t0, "s")


def neighbour_cases():
    # reference tests/test_neighbours.py: rect(20,20)
    v, f = meshgen.rect(20, 20)
    mesh = ref.TriMesh(v, f)
    out = {}
    for lt in ("cell-list", "verlet-list"):
        tag = lt.split("-")[0]
        for rl in (0.2, 0.5, 1.2):
            p = ref.EnergyParams()
            p.repulse_params.n_search, p.repulse_params.rlist, p.repulse_params.exclusion_level = lt, rl, 0
            d, i, j = ref.make_nlist(mesh, p).distance_matrix(mesh, 0.123)
            order = np.lexsort((j, i))
            out[f"{tag}_dm_{rl}_d"] = np.asarray(d)[order]
            out[f"{tag}_dm_{rl}_i"] = np.asarray(i, np.int32)[order]
            out[f"{tag}_dm_{rl}_j"] = np.asarray(j, np.int32)[order]
        for ex in (0, 1, 2):
            p = ref.EnergyParams()
            p.repulse_params.n_search, p.repulse_params.rlist, p.repulse_params.exclusion_level = lt, 0.2, ex
            _, jd = ref.make_nlist(mesh, p).point_distances(mesh, 151, 0.2)
            out[f"{tag}_pd151_excl{ex}"] = np.sort(np.asarray(jd, np.int32))
    save("neighbours_rect20", **out)


def topology_cases():
    out = {}
    for name, (v, f) in {"ico3": meshgen.icosphere(3), "rect20": meshgen.rect(20, 20),
                         "tube32": meshgen.tube(1.0, 1.0, 32), "ico4": meshgen.icosphere(4)}.items():
        mesh = ref.TriMesh(v, f)
        for k, a in zip(("he_to", "he_next", "he_face", "v_he", "f_he"), refprobe.halfedge_arrays(mesh)):
            out[f"{name}_{k}"] = a
        out[f"{name}_fv"] = mesh.fv_indices()
    # after a fixed sequence of flips
    v, f = meshgen.icosphere(3)
    mesh = ref.TriMesh(v, f)
    rng = np.random.default_rng(0)
    edges = rng.integers(0, mesh.n_edges(), 2000).astype(np.int32)
    ok = np.zeros(len(edges), np.int32)
    for k, e in enumerate(edges):
        ok[k] = refprobe.is_flip_ok(mesh, int(e))
        if ok[k]:
            refprobe.flip(mesh, int(e))
    out["flipseq_edges"], out["flipseq_ok"] = edges, ok
    for k, a in zip(("he_to", "he_next", "he_face", "v_he", "f_he"), refprobe.halfedge_arrays(mesh)):
        out[f"flipseq_{k}"] = a
    out["flipseq_fv"] = mesh.fv_indices()
    save("topology", **out)


def analytic_cases():
    # reference tests/test_energy.py: sphere r=1,0.5 (icosphere(5)), tube n=32 r=1,0.5
    out = {}
    for tag, r in (("r1", 1.0), ("r05", 0.5)):
        v, f = meshgen.icosphere(5)
        mesh = ref.TriMesh(v * r, f)
        out[f"sphere_{tag}"] = cases.props6(ref.EnergyManager(mesh, ref.EnergyParams()).properties(mesh))
        v, f = meshgen.tube(1.0, r, 32)
        mesh = ref.TriMesh(v, f)
        out[f"tube_{tag}"] = cases.props6(ref.EnergyManager(mesh, ref.EnergyParams()).properties(mesh))
    # docs example (doc/python-module.rst:60-98)
    v, f = meshgen.icosphere(3, project_each_round=False)
    v = v.astype(np.float32).astype(np.float64)
    mesh = ref.TriMesh(v, f)
    ep = ref.EnergyParams()
    ep.kappa_b, ep.kappa_a, ep.kappa_v, ep.kappa_c, ep.kappa_t, ep.kappa_r = 30.0, 1e6, 1e6, 1e6, 1e5, 1e3
    ep.area_frac = 0.5
    es = ref.EnergyManager(mesh, ep)
    out["docs_energy"] = es.energy(mesh)
    out["docs_grad"] = es.gradient(mesh)
    out["docs_props"] = cases.props6(es.properties(mesh))
    save("analytic", **out)


def primitive_cases():
    # reference tests/test_properties.py: two triangles hinged at +-45 degrees
    out = {}
    for tag, alpha in (("p45", np.pi / 4), ("m45", -np.pi / 4)):
        x, z = np.cos(alpha) * 0.5, np.sin(alpha) * 0.5
        pts = np.array([[0.0, -0.5, 0.0], [0.0, 0.5, 0.0], [-x, 0.0, z], [x, 0.0, z]])
        cells = np.array([[0, 1, 2], [1, 0, 3]])
        mesh = ref.TriMesh(pts, cells)
        e = mesh.halfedge_handle(0)
        out[f"{tag}_scalars"] = np.array([ref.edge_length(mesh, e), ref.face_area(mesh, e),
                                          ref.face_volume(mesh, e), ref.dihedral_angle(mesh, e)])
        out[f"{tag}_d_length"] = ref.edge_length_grad(mesh, e)
        out[f"{tag}_d_area"] = ref.face_area_grad(mesh, e)
        out[f"{tag}_d_volume"] = ref.face_volume_grad(mesh, e)
        out[f"{tag}_d_angle"] = ref.dihedral_angle_grad(mesh, e)
    # every half-edge of a jittered icosphere(2): scalars + all gradients via the same calls
    v, f = meshgen.jittered_icosphere(2, sigma_rel=0.1, seed=3)
    mesh = ref.TriMesh(v, f)
    nh = 2 * mesh.n_edges()
    geo = np.zeros((nh, 40))
    for h in range(nh):
        e = mesh.halfedge_handle(h)
        geo[h, 0:4] = [ref.edge_length(mesh, e), ref.face_area(mesh, e), ref.face_volume(mesh, e), ref.dihedral_angle(mesh, e)]
        geo[h, 4:10] = ref.edge_length_grad(mesh, e).ravel()
        geo[h, 10:19] = ref.face_area_grad(mesh, e).ravel()
        geo[h, 19:28] = ref.face_volume_grad(mesh, e).ravel()
        geo[h, 28:40] = ref.dihedral_angle_grad(mesh, e).ravel()
    out["ico2_geometry"] = geo
    save("primitives", **out)


def flip_cases():
    for name in cases.FLIP_CASES:
        case, edges, u = cases.flip_inputs(name)
        mesh, es = cases.build(ref, case)
        # per-proposal edge_vertex_properties before the sweep (on the initial mesh)
        evp = np.array([refprobe.edge_vertex_properties(mesh, es, int(e)) for e in edges[:64]])
        flags, ener, props = refprobe.flip_sweep(mesh, es, edges, u)
        out = dict(edges=edges, u=u, flags=flags, energies=ener, props_after=props, evp_first64=evp,
                   fv_after=mesh.fv_indices())
        for k, a in zip(("he_to", "he_next", "he_face", "v_he", "f_he"), refprobe.halfedge_arrays(mesh)):
            out[f"after_{k}"] = a
        save(name, **out)


def reference_hmc_module():
    """The reference's own Python integrator (src/trimem/mc/hmc.py), imported from where it lies
    with the oracle build of trimem.core behind its `from .. import core`."""
    import importlib.util
    import types
    pkg = types.ModuleType("trimem.mc")
    pkg.__path__ = ["/root/reference/src/trimem/mc"]
    sys.modules["trimem.mc"] = pkg
    spec = importlib.util.spec_from_file_location("trimem.mc.hmc", "/root/reference/src/trimem/mc/hmc.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["trimem.mc.hmc"] = mod
    spec.loader.exec_module(mod)
    return mod


def trajectory_cases():
    hmc = reference_hmc_module()
    for name in cases.VV_CASES:
        case, p0, mass, dt, n = cases.vv_inputs(name)
        mesh, es = cases.build(ref, case)
        x0 = mesh.points().copy()

        def force(x):                      # mc/evaluators.py:104-107,139-156: mesh.x = x ; -estore.gradient
            np.copyto(mesh.points(), x)
            return -es.gradient(mesh)

        def nlog_prob(x):
            np.copyto(mesh.points(), x)
            return es.energy(mesh)

        u0 = nlog_prob(x0)
        xn, pn = hmc._vv_integration(x0, p0, force, mass, dt, n)
        u1 = nlog_prob(xn)
        k0 = 0.5 * p0.ravel().dot(p0.ravel()) / mass
        k1 = 0.5 * pn.ravel().dot(pn.ravel()) / mass
        save(name, x=xn, p=pn, stats=np.array([u0, k0, u1, k1]), moved=np.array([np.abs(xn - x0).max()]))


def hmc_chain_cases():
    """Chains of the reference's own HMC class (src/trimem/mc/hmc.py:48-190) on the reference build."""
    hmc = reference_hmc_module()
    import contextlib, io
    for name, (case, seed, opts, nsteps) in cases.HMC_CASES.items():
        mesh, es = cases.build(ref, case)
        x0 = mesh.points().copy()

        def nlog_prob(x):
            np.copyto(mesh.points(), x)
            return es.energy(mesh)

        def grad(x):
            np.copyto(mesh.points(), x)
            return es.gradient(mesh)

        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            chain = hmc.HMC(x0, nlog_prob, grad, counter=hmc.get_step_counters(), options=dict(opts))
            xs, accs = [], []
            for _ in range(nsteps):
                chain.step()
                xs.append(chain.x.copy())
                accs.append(chain.acc)
        save(name, xs=np.array(xs), acc=np.array(accs), temperature=np.array([chain.T]))


def restart_case():
    """Checkpoint hand-off (src/trimem/mc/util.py:153-160): after a flip sweep, two continuation updates and a
    displacement, the state (x, f, lam, delta) is all that is handed over; the restart builds the EnergyManager on
    the INITIAL mesh with the checkpoint's continuation parameters, the mesh from (x, f), rebuilds the list on it
    and evaluates. Golden = the reference's answers on the restarted objects."""
    case, edges, u = cases.flip_inputs(cases.RESTART_FLIPS)
    mesh, es = cases.build(ref, case)
    x_init, f_init = mesh.points().copy(), mesh.fv_indices().copy()
    refprobe.flip_sweep(mesh, es, edges, u)
    # (eparams is handed out by reference: nested assignments are live, tests/test_gpu_energy.py)
    es.eparams.continuation_params.delta, es.eparams.continuation_params.lam = cases.RESTART_CONT
    es.eparams.area_frac, es.eparams.volume_frac, es.eparams.curvature_frac = cases.RESTART_FRACS
    es.update_reference_properties()
    es.update_reference_properties()
    x = cases.restart_displacement(mesh.points())
    np.copyto(mesh.points(), x)
    es.update_repulsion(mesh)
    e_cont, g_cont = es.energy(mesh), es.gradient(mesh)
    lam, delta, f = es.eparams.continuation_params.lam, es.eparams.continuation_params.delta, mesh.fv_indices().copy()
    # restart
    mesh0 = ref.TriMesh(x_init, f_init)
    ep2 = workload.energy_params(ref, mesh0, **cases.CASES[case][2])
    ep2.area_frac, ep2.volume_frac, ep2.curvature_frac = cases.RESTART_FRACS
    cp2 = ep2.continuation_params
    cp2.lam, cp2.delta = lam, delta
    ep2.continuation_params = cp2
    es2 = ref.EnergyManager(mesh0, ep2)
    mesh2 = ref.TriMesh(x, f)
    es2.update_repulsion(mesh2)
    e2, g2, p2 = es2.energy(mesh2), es2.gradient(mesh2), cases.props6(es2.properties(mesh2))
    print("restart: continued", e_cont, "restarted", e2, "grad diff", np.abs(g2 - g_cont).max() / np.abs(g_cont).max())
    save("restart_" + case, x=x, f=f, lam=np.array([lam]), delta=np.array([delta]), energy=np.array([e2]), grad=g2,
         props=p2, energy_continued=np.array([e_cont]))


def open_mesh_cases():
    out = {}
    for name in cases.OPEN_MESHES:
        v, f = cases.open_mesh(name)
        mesh = ref.TriMesh(v, f)
        es = ref.EnergyManager(mesh, cases.open_mesh_params(ref, mesh))
        out[f"{name}_props"] = cases.props6(es.properties(mesh))
        out[f"{name}_energy"] = es.energy(mesh)
        out[f"{name}_grad"] = es.gradient(mesh)
    # the harmonic area bond on an open mesh
    v, f = cases.open_mesh("rect")
    mesh = ref.TriMesh(v, f)
    es = ref.EnergyManager(mesh, cases.open_mesh_params(ref, mesh, bond="Area"))
    out["rect_area_energy"] = es.energy(mesh)
    out["rect_area_grad"] = es.gradient(mesh)
    # every half-edge primitive (scalars and gradients, bindings.cpp:249-407) of the small open meshes,
    # boundary half-edges included: the reference evaluates them on the boundary loop's sectors
    for name in ("rect", "fan"):
        v, f = cases.open_mesh(name)
        mesh = ref.TriMesh(v, f)
        nh = 2 * mesh.n_edges()
        geo = np.zeros((nh, 40))
        bnd = np.zeros(nh, bool)
        arrs = refprobe.halfedge_arrays(mesh)
        for h in range(nh):
            e = mesh.halfedge_handle(h)
            bnd[h] = arrs[2][h] < 0 or arrs[2][h ^ 1] < 0
            geo[h, 0:4] = [ref.edge_length(mesh, e), ref.face_area(mesh, e), ref.face_volume(mesh, e), ref.dihedral_angle(mesh, e)]
            geo[h, 4:10] = ref.edge_length_grad(mesh, e).ravel()
            geo[h, 10:19] = ref.face_area_grad(mesh, e).ravel()
            geo[h, 19:28] = ref.face_volume_grad(mesh, e).ravel()
            geo[h, 28:40] = ref.dihedral_angle_grad(mesh, e).ravel()
        out[f"{name}_geometry"] = geo
        out[f"{name}_on_boundary_edge"] = bnd
    save("open_meshes", **out)


def thread_spread():
    """The reference's own 1-vs-N-thread spread on the headline case (its noise floor)."""
    import subprocess
    code = ("import sys,os;sys.path.insert(0,%r);sys.path.insert(0,%r);sys.path.insert(0,%r);"
            "import trimem.core as ref,cases,numpy as np;m,e=cases.build(ref,'ico8_stress_verlet2');"
            "print(repr(e.energy(m)),repr(float(np.linalg.norm(e.gradient(m)))))"
            % (os.path.join(ROOT, "oracle", "_ref"), ROOT, os.path.join(ROOT, "tests")))
    res = {}
    for thr in (1, 8):
        env = dict(os.environ, OMP_NUM_THREADS=str(thr))
        o = subprocess.check_output([sys.executable, "-c", code], env=env, text=True).split()
        res[thr] = (float(o[0]), float(o[1]))
    save("thread_spread", energy=np.array([res[1][0], res[8][0]]), grad_norm=np.array([res[1][1], res[8][1]]))


def flip_statistics():
    """Sweep-level statistics of the reference's flip / pflip (flips.cpp:22-165; clock-seeded, so distributions
    rather than sequences): acceptance counts and energies of independent chains of sweeps at fixed vertex
    positions. Two sizes:
      * ico3_flip (642 vertices, reference tests/test_flips.py parameters): FLIPSTAT_SMALL chains of sweeps,
        every chain from the initial mesh; per chain the accepted flips of each sweep and the energy after each.
      * ico6_flip (40,962 vertices, config 2's physical variant, all kappa = 1): a few chains of 100 sweeps.
    pflip runs with ONE OpenMP thread here for the small mesh: its conflict handling drops proposals whose patch
    is locked by another thread (flips.cpp:118-124), so its acceptance COUNT depends on the thread count; with one
    thread it is the same Markov kernel as flip. The large mesh also records pflip at 8 threads (a conflict there
    has probability ~0.2 % per proposal)."""
    import subprocess
    chains, sweeps, ratio = cases.FLIPSTAT_SMALL
    out = {}
    for fn in ("flip", "pflip"):
        acc = np.zeros((chains, sweeps), np.int32)
        ener = np.zeros((chains, sweeps))
        for c in range(chains):
            mesh, es = cases.build(ref, "ico3_flip")
            for s in range(sweeps):
                acc[c, s] = getattr(ref, fn)(mesh, es, ratio)
                ener[c, s] = es.energy(mesh)
        out[f"small_{fn}_acc"], out[f"small_{fn}_energy"] = acc, ener
        print(fn, "small: mean acc/sweep", acc[:, sweeps // 3:].mean(), "final energy", ener[:, -1].mean(), ener[:, -1].std())
    mesh, es = cases.build(ref, "ico3_flip")
    out["small_energy0"] = es.energy(mesh)
    out["small_n_edges"] = mesh.n_edges()
    # large mesh: separate processes (thread count per run)
    code = ("import sys,os;sys.path.insert(0,%r);sys.path.insert(0,%r);sys.path.insert(0,%r);"
            "import trimem.core as ref,cases,numpy as np;m,e=cases.build(ref,'ico6_flip');fn=getattr(ref,sys.argv[1]);"
            "n=int(sys.argv[2]);r=float(sys.argv[3]);print(repr(e.energy(m)));\n"
            "for s in range(n): a=fn(m,e,r);print(a,repr(e.energy(m)),flush=True)"
            % (os.path.join(ROOT, "oracle", "_ref"), ROOT, os.path.join(ROOT, "tests")))
    nsw, lratio, plan = cases.FLIPSTAT_LARGE
    for fn, threads, nchains in plan:
        accs, eners = [], []
        procs = [subprocess.Popen([sys.executable, "-c", code, fn, str(nsw), str(lratio)], text=True, stdout=subprocess.PIPE,
                                  env=dict(os.environ, OMP_NUM_THREADS=str(threads))) for _ in range(nchains)]
        for pr in procs:
            lines = pr.communicate()[0].split("\n")
            out["large_energy0"] = float(lines[0])
            rows = [l.split() for l in lines[1:] if l.strip()]
            accs.append([int(r[0]) for r in rows]); eners.append([float(r[1]) for r in rows])
        out[f"large_{fn}{threads}_acc"] = np.array(accs, np.int32)
        out[f"large_{fn}{threads}_energy"] = np.array(eners)
        print(fn, threads, "large: acc/sweep", np.mean(accs), "energy after", nsw, "sweeps", [e[-1] for e in eners])
    save("flip_statistics", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["topology", "analytic", "primitives", "neighbours", "flips", "energy", "spread", "trajectory", "hmc", "open", "restart"]
    if "topology" in which: topology_cases()
    if "analytic" in which: analytic_cases()
    if "primitives" in which: primitive_cases()
    if "neighbours" in which: neighbour_cases()
    if "flips" in which: flip_cases()
    if "energy" in which: energy_cases()
    if "spread" in which: thread_spread()
    if "trajectory" in which: trajectory_cases()
    if "hmc" in which: hmc_chain_cases()
    if "open" in which: open_mesh_cases()
    if "restart" in which: restart_case()
    if "gtiles" in which: gradient_tile_sums()
    if "flipstats" in which: flip_statistics()
    if "big" in which: big_case()
