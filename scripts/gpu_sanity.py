"""Ad-hoc GPU sanity run: product (trimem_b200.core) vs oracle/_ref on a few cases."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref")); sys.path.insert(0, ROOT)
import trimem.core as ref
import trimem_b200.core as m
from trimem_b200 import meshgen

def params(mod, mesh, bond="Edge", rep=None, nsearch="cell-list", excl=2, kappa=None, fracs=(1.0, 0.9, 1.0)):
    a, l = mod.avg_tri_props(mesh)
    bp = mod.BondParams(); bp.type = getattr(mod.BondType, bond); bp.r = 2
    bp.lc0 = 1.15 * l; bp.lc1 = 0.85 * l; bp.a0 = a
    ep = mod.EnergyParams()
    kb, ka, kv, kc, kt, kr = kappa or (30.0, 1e6, 1e6, 0.0, 1e5, 0.0)
    ep.kappa_b, ep.kappa_a, ep.kappa_v, ep.kappa_c, ep.kappa_t, ep.kappa_r = kb, ka, kv, kc, kt, kr
    ep.area_frac, ep.volume_frac, ep.curvature_frac = fracs
    ep.bond_params = bp
    rp = mod.SurfaceRepulsionParams(); rp.n_search = nsearch; rp.exclusion_level = excl; rp.r = 2
    if rep == "stress": rp.lc1 = 3.0 * l; rp.rlist = 3.2 * l
    else: rp.lc1 = 0.0; rp.rlist = 4.0 / 3.0 * l
    ep.repulse_params = rp
    return ep

def rel(a, b): return abs(a - b) / max(abs(b), 1e-300)

def compare(name, v, f, **kw):
    ma, mb = ref.TriMesh(v, f), m.TriMesh(v, f)
    pa, pb = params(ref, ma, **kw), params(m, mb, **kw)
    t0 = time.time(); ea = ref.EnergyManager(ma, pa); t1 = time.time(); eb = m.EnergyManager(mb, pb); t2 = time.time()
    ra, rb = ea.properties(ma), eb.properties(mb)
    names = ["area", "volume", "curvature", "bending", "tethering", "repulsion"]
    d = {n: rel(getattr(rb, n), getattr(ra, n)) if getattr(ra, n) != 0 else abs(getattr(rb, n)) for n in names}
    Ea, Eb = ea.energy(ma), eb.energy(mb)
    ga, gb = ea.gradient(ma), eb.gradient(mb)
    print(f"{name:38s} V={len(v):7d} props {max(d.values()):.2e} ({max(d, key=d.get)}) E {rel(Eb, Ea):.2e} "
          f"|dg|/|g| {np.linalg.norm(ga - gb) / np.linalg.norm(ga):.2e}  ctor ref {t1 - t0:.2f}s gpu {t2 - t1:.2f}s", flush=True)
    return ma, mb, ea, eb

print("devices", m.device_count())
v, f = meshgen.icosphere(3)
k1 = (1.0, 1.0, 1.0, 1.0, 1.0, 0.0)
compare("ico3 kappa=1 Edge", v, f, kappa=k1, fracs=(0.8, 0.8, 0.8))
compare("ico3 kappa=1 Area", v, f, kappa=k1, fracs=(0.8, 0.8, 0.8), bond="Area")
v, f = meshgen.jittered_icosphere(4)
compare("ico4 jitter config1", v, f)
for ns in ("cell-list", "verlet-list"):
    for ex in (0, 1, 2):
        compare(f"ico4 stress {ns} excl{ex}", v, f, rep="stress", nsearch=ns, excl=ex, kappa=(30.0, 1e6, 1e6, 0.0, 1e5, 1e3))
v, f = meshgen.tube(1.0, 1.0, 32)
ma, mb = ref.TriMesh(v, f), m.TriMesh(v, f)
ea, eb = ref.EnergyManager(ma, ref.EnergyParams()), m.EnergyManager(mb, m.EnergyParams())
ra, rb = ea.properties(ma), eb.properties(mb)
print("tube props", [(getattr(ra, n), getattr(rb, n)) for n in ("area", "volume", "curvature", "bending")])
v, f = meshgen.jittered_icosphere(6)
compare("ico6 stress verlet excl2", v, f, rep="stress", nsearch="verlet-list", kappa=(30.0, 1e6, 1e6, 0.0, 1e5, 1e3))
compare("ico6 stress cell excl2", v, f, rep="stress", nsearch="cell-list", kappa=(30.0, 1e6, 1e6, 0.0, 1e5, 1e3))
# neighbour lists on rect
from scipy.sparse import coo_matrix
from scipy.spatial import KDTree
v, f = meshgen.rect(20, 20)
mb = m.TriMesh(v, f); ma = ref.TriMesh(v, f)
for lt in ("cell-list", "verlet-list"):
    for rl in (0.2, 0.5, 1.2):
        p = m.EnergyParams(); p.repulse_params.n_search = lt; p.repulse_params.rlist = rl; p.repulse_params.exclusion_level = 0
        nl = m.make_nlist(mb, p)
        d, i, j = nl.distance_matrix(mb, 0.123)
        M = coo_matrix((d, (i, j)), shape=(400, 400))
        if lt == "cell-list": M = M + M.T
        tree = KDTree(v); C = tree.sparse_distance_matrix(tree, 0.123)
        print(lt, rl, "dm ok", np.allclose(C.todense(), M.todense()), len(d))
    for ex, nn in ((0, 44), (1, 38), (2, 26)):
        p = m.EnergyParams(); p.repulse_params.n_search = lt; p.repulse_params.rlist = 0.2; p.repulse_params.exclusion_level = ex
        nl = m.make_nlist(mb, p); _, jd = nl.point_distances(mb, 151, 0.2)
        pr = ref.EnergyParams(); pr.repulse_params.n_search = lt; pr.repulse_params.rlist = 0.2; pr.repulse_params.exclusion_level = ex
        _, jr = ref.make_nlist(ma, pr).point_distances(ma, 151, 0.2)
        print(lt, "excl", ex, len(jd), nn, sorted(jd) == sorted(jr))
# big one
v, f = meshgen.jittered_icosphere(8)
t0 = time.time(); mb = m.TriMesh(v, f); print("TriMesh 655k", time.time() - t0)
pb = params(m, mb, rep="stress", nsearch="verlet-list", kappa=(30.0, 1e6, 1e6, 0.0, 1e5, 1e3))
t0 = time.time(); eb = m.EnergyManager(mb, pb); print("EnergyManager 655k", time.time() - t0)
for _ in range(3):
    t0 = time.time(); E = eb.energy(mb); t1 = time.time(); g = eb.gradient(mb); t2 = time.time()
    print("655k E", repr(E), "|g|", repr(np.linalg.norm(g)), "energy %.2f ms grad %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
print("expected (SURVEY App. D, 8 thr ref): E 3.693272603980e+11 |g| 3.697022759067e+11")
