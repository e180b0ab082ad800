"""TEST INFRASTRUCTURE ONLY. ctypes front-end of oracle/trimem_oracle.c.

Importable only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs (when the reference build oracle/_ref is unavailable). It
mirrors just enough of the trimem.core operator interface (TriMesh, EnergyManager,
make_nlist, flip sweep with injected randomness) for the parity tests to read like
the reference's own tests.
"""
import ctypes as C
import os
import subprocess
import time

import numpy as np

try:
    from . import halfedge
except ImportError:  # imported as a top-level module with oracle/ on sys.path
    import halfedge

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "_build", "libtrimem_oracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _Mesh(C.Structure):
    _fields_ = [("V", C.c_int), ("HE", C.c_int), ("he_to", _ip), ("he_next", _ip), ("he_face", _ip),
                ("v_he", _ip), ("x", _dp)]


class _MMesh(C.Structure):
    _fields_ = [("V", C.c_int), ("HE", C.c_int), ("F", C.c_int), ("he_to", _ip), ("he_next", _ip),
                ("he_face", _ip), ("v_he", _ip), ("f_he", _ip), ("x", _dp)]


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("kappa_b", "kappa_a", "kappa_v", "kappa_c", "kappa_t", "kappa_r",
                 "ref_area", "ref_volume", "ref_curvature", "bond_lc0", "bond_lc1", "bond_a0")] + [
        ("bond_r", C.c_int), ("bond_type", C.c_int), ("rep_lc1", C.c_double), ("rep_r", C.c_int), ("pad", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "trimem_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
        L = C.CDLL(_SO)
        L.to_nlist_build.restype = C.c_void_p
        L.to_nlist_build.argtypes = [C.POINTER(_Mesh), C.c_int, C.c_double, C.c_int, C.c_double]
        L.to_nlist_free.argtypes = [C.c_void_p]
        L.to_nlist_entries.restype = C.c_long
        L.to_nlist_entries.argtypes = [C.c_void_p]
        L.to_nlist_grid.argtypes = [C.c_void_p, _ip, _dp, _dp]
        L.to_nlist_distance_matrix.restype = C.c_long
        L.to_nlist_distance_matrix.argtypes = [C.c_void_p, C.POINTER(_Mesh), C.c_double, C.POINTER(_dp),
                                               C.POINTER(_ip), C.POINTER(_ip)]
        L.to_nlist_point_distances.argtypes = [C.c_void_p, C.POINTER(_Mesh), C.c_int, C.c_double, C.c_int, _dp, _ip]
        L.to_free.argtypes = [C.c_void_p]
        L.to_vertex_properties.argtypes = [C.POINTER(_Mesh), C.POINTER(Params), C.c_void_p, _dp]
        L.to_reduce.argtypes = [C.c_int, _dp, _dp]
        L.to_energy.restype = C.c_double
        L.to_energy.argtypes = [C.POINTER(Params), _dp]
        L.to_gradient.argtypes = [C.POINTER(_Mesh), C.POINTER(Params), C.c_void_p, _dp, _dp, _dp]
        L.to_area_edges.argtypes = [C.POINTER(_Mesh), C.c_int, _ip, _dp, _dp]
        L.to_halfedge_geometry.argtypes = [C.POINTER(_Mesh), C.c_int, _dp]
        L.to_is_flip_ok.argtypes = [C.POINTER(_MMesh), C.c_int]
        L.to_flip.argtypes = [C.POINTER(_MMesh), C.c_int]
        L.to_edge_vertex_properties.argtypes = [C.POINTER(_MMesh), C.POINTER(Params), C.c_void_p, C.c_int, _dp]
        L.to_flip_sweep.argtypes = [C.POINTER(_MMesh), C.POINTER(Params), C.c_void_p, C.c_int, _ip, _dp, _dp, _ip, _dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class TriMesh:
    """Half-edge mesh with OpenMesh numbering (see halfedge.py)."""

    def __init__(self, points, faces, exact=None):
        self.x = np.ascontiguousarray(points, dtype=np.float64).copy()
        faces = np.asarray(faces)
        if exact is None:
            exact = len(faces) <= 20000
        builder = halfedge.build if exact else halfedge.build_closed
        self.he_to, self.he_next, self.he_face, self.v_he, self.f_he = builder(len(self.x), faces)

    def points(self):
        return self.x

    def n_vertices(self):
        return len(self.x)

    def n_edges(self):
        return len(self.he_to) // 2

    def fv_indices(self):
        h0 = self.f_he
        h1 = self.he_next[h0]
        h2 = self.he_next[h1]
        return np.stack([self.he_to[h0], self.he_to[h1], self.he_to[h2]], axis=1).astype(np.int32)

    def _c(self):
        return _Mesh(len(self.x), len(self.he_to), _i(self.he_to), _i(self.he_next), _i(self.he_face),
                     _i(self.v_he), _d(self.x))

    def _cm(self):
        return _MMesh(len(self.x), len(self.he_to), len(self.f_he), _i(self.he_to), _i(self.he_next),
                      _i(self.he_face), _i(self.v_he), _i(self.f_he), _d(self.x))

    def is_flip_ok(self, e):
        return bool(lib().to_is_flip_ok(C.byref(self._cm()), int(e)))

    def flip(self, e):
        lib().to_flip(C.byref(self._cm()), int(e))

    def halfedge_geometry(self, he):
        out = np.empty(40)
        lib().to_halfedge_geometry(C.byref(self._c()), int(he), _d(out))
        return out


def avg_tri_props(mesh):
    a, l = C.c_double(), C.c_double()
    lib().to_area_edges(C.byref(mesh._c()), len(mesh.f_he), _i(mesh.f_he), C.byref(a), C.byref(l))
    return a.value / len(mesh.f_he), l.value / mesh.n_edges()


class NeighbourList:
    def __init__(self, mesh, n_search, rlist, exclusion_level, box_eps=1.0e-6):
        kinds = {"cell-list": 0, "verlet-list": 1}
        if n_search not in kinds:
            raise RuntimeError("Unknown neighbour search algorithm.")
        if exclusion_level not in (0, 1, 2):
            raise RuntimeError("Unsupported exclusion level")
        self.h = lib().to_nlist_build(C.byref(mesh._c()), kinds[n_search], rlist, exclusion_level, box_eps)

    def __del__(self):
        try:
            lib().to_nlist_free(self.h)
        except Exception:
            pass

    def n_entries(self):
        return lib().to_nlist_entries(self.h)

    def grid(self):
        shape = np.zeros(3, np.int32)
        box = np.zeros(6)
        cell = np.zeros(3)
        lib().to_nlist_grid(self.h, _i(shape), _d(box), _d(cell))
        return shape, box, cell

    def distance_matrix(self, mesh, rdist):
        d, i, j = _dp(), _ip(), _ip()
        n = lib().to_nlist_distance_matrix(self.h, C.byref(mesh._c()), rdist, C.byref(d), C.byref(i), C.byref(j))
        if n == 0:
            return np.zeros(0), np.zeros(0, np.int32), np.zeros(0, np.int32)
        out = (np.ctypeslib.as_array(d, (n,)).copy(), np.ctypeslib.as_array(i, (n,)).copy(),
               np.ctypeslib.as_array(j, (n,)).copy())
        for p in (d, i, j):
            lib().to_free(C.cast(p, C.c_void_p))
        return out

    def point_distances(self, mesh, pid, rdist):
        cap = 1 << 16
        d = np.empty((cap, 3))
        j = np.empty(cap, np.int32)
        n = lib().to_nlist_point_distances(self.h, C.byref(mesh._c()), pid, rdist, cap, _d(d), _i(j))
        return d[:n].copy(), j[:n].copy()


class EnergyManager:
    """EnergyManager (reference src/energy.cpp) on the C restatement. `params` is a dict with the
    EnergyParams fields (flat): kappa_*, *_frac, bond_{type,lc0,lc1,a0,r}, rep_{lc1,r}, n_search,
    rlist, exclusion_level, lam."""

    def __init__(self, mesh, params):
        self.params = dict(params)
        self.nlist = NeighbourList(mesh, params.get("n_search", "cell-list"), params.get("rlist", 0.1),
                                   params.get("exclusion_level", 2))
        self.initial_props = np.zeros(6)
        self.initial_props = self.properties(mesh)

    def update_repulsion(self, mesh):
        p = self.params
        self.nlist = NeighbourList(mesh, p.get("n_search", "cell-list"), p.get("rlist", 0.1), p.get("exclusion_level", 2))

    def cparams(self):
        p = self.params
        lam = p.get("lam", 1.0)
        c = Params()
        for k in ("kappa_b", "kappa_a", "kappa_v", "kappa_c", "kappa_t", "kappa_r"):
            setattr(c, k, p.get(k, 0.0))
        c.ref_area = (1.0 - lam * (1.0 - p.get("area_frac", 1.0))) * self.initial_props[0]
        c.ref_volume = (1.0 - lam * (1.0 - p.get("volume_frac", 1.0))) * self.initial_props[1]
        c.ref_curvature = (1.0 - lam * (1.0 - p.get("curvature_frac", 1.0))) * self.initial_props[2]
        c.bond_lc0, c.bond_lc1, c.bond_a0 = p.get("bond_lc0", 1.0), p.get("bond_lc1", 0.0), p.get("bond_a0", 1.0)
        c.bond_r = p.get("bond_r", 2)
        c.bond_type = {"Edge": 0, "Area": 1}[p.get("bond_type", "Edge")]
        c.rep_lc1, c.rep_r = p.get("rep_lc1", 0.0), p.get("rep_r", 2)
        return c

    def vertex_properties(self, mesh):
        vp = np.empty((mesh.n_vertices(), 6))
        c = self.cparams()
        lib().to_vertex_properties(C.byref(mesh._c()), C.byref(c), self.nlist.h, _d(vp))
        return vp

    def properties(self, mesh):
        vp = self.vertex_properties(mesh)
        out = np.empty(6)
        lib().to_reduce(len(vp), _d(vp), _d(out))
        return out

    def energy(self, mesh_or_props):
        props = mesh_or_props if isinstance(mesh_or_props, np.ndarray) else self.properties(mesh_or_props)
        c = self.cparams()
        props = np.ascontiguousarray(props, dtype=np.float64)
        return lib().to_energy(C.byref(c), _d(props))

    def gradient(self, mesh):
        vp = self.vertex_properties(mesh)
        sums = np.empty(6)
        lib().to_reduce(len(vp), _d(vp), _d(sums))
        g = np.empty((mesh.n_vertices(), 3))
        c = self.cparams()
        lib().to_gradient(C.byref(mesh._c()), C.byref(c), self.nlist.h, _d(vp), _d(sums), _d(g))
        return g

    def edge_vertex_properties(self, mesh, e):
        out = np.empty(6)
        c = self.cparams()
        lib().to_edge_vertex_properties(C.byref(mesh._cm()), C.byref(c), self.nlist.h, int(e), _d(out))
        return out

    def flip_sweep(self, mesh, edges, u):
        """flip_serial with injected randomness; mutates mesh; returns (flags, energies, props)."""
        edges = np.ascontiguousarray(edges, dtype=np.int32)
        u = np.ascontiguousarray(u, dtype=np.float64)
        props = self.properties(mesh)
        flags = np.zeros(len(edges), np.int32)
        ener = np.zeros(len(edges))
        c = self.cparams()
        lib().to_flip_sweep(C.byref(mesh._cm()), C.byref(c), self.nlist.h, len(edges), _i(edges), _d(u),
                            _d(props), _i(flags), _d(ener))
        return flags, ener, props


def vv_integration(mesh, estore, p0, mass, dt, n_steps):
    """Restatement of _vv_integration (src/trimem/mc/hmc.py:14-26) with force = -gradient evaluated
    the way mc/evaluators.py:104-107,139-156 does (positions written into the mesh, then
    EnergyManager.gradient). Returns (x, p, [U(x0), K(p0), U(xN), K(pN)]); the mesh is restored."""
    x0 = mesh.points().copy()

    def force(x):
        np.copyto(mesh.points(), x)
        return -estore.gradient(mesh)

    u0 = estore.energy(mesh)
    x, p = x0, np.asarray(p0, dtype=np.float64)
    a = force(x)
    for _ in range(n_steps):
        x = x + (p * dt + 0.5 * a * dt**2) / mass
        an = force(x)
        p = p + 0.5 * (a + an) * dt
        a = an
    np.copyto(mesh.points(), x)
    u1 = estore.energy(mesh)
    np.copyto(mesh.points(), x0)
    stats = np.array([u0, 0.5 * p0.ravel().dot(p0.ravel()) / mass, u1, 0.5 * p.ravel().dot(p.ravel()) / mass])
    return x, p, stats


def params_from(core_params):
    """Flat dict from a trimem.core-style EnergyParams object (reference or product module)."""
    ep = core_params
    bt = str(ep.bond_params.type).split(".")[-1]
    return dict(kappa_b=ep.kappa_b, kappa_a=ep.kappa_a, kappa_v=ep.kappa_v, kappa_c=ep.kappa_c,
                kappa_t=ep.kappa_t, kappa_r=ep.kappa_r, area_frac=ep.area_frac, volume_frac=ep.volume_frac,
                curvature_frac=ep.curvature_frac, bond_type=bt, bond_lc0=ep.bond_params.lc0,
                bond_lc1=ep.bond_params.lc1, bond_a0=ep.bond_params.a0, bond_r=ep.bond_params.r,
                rep_lc1=ep.repulse_params.lc1, rep_r=ep.repulse_params.r, n_search=ep.repulse_params.n_search,
                rlist=ep.repulse_params.rlist, exclusion_level=ep.repulse_params.exclusion_level,
                lam=ep.continuation_params.lam)


def workload_params(mesh, repulsion="none", n_search="verlet-list", exclusion=2, bond="Edge",
                    kappa=None, fracs=(1.0, 0.9, 1.0)):
    """Same parametrisation as trimem_b200.workload.energy_params, as a flat dict."""
    a, l = avg_tri_props(mesh)
    kr = 0.0 if repulsion == "none" else 1.0e3
    kb, ka, kv, kc, kt, kr = kappa or (30.0, 1.0e6, 1.0e6, 0.0, 1.0e5, kr)
    lc1, rlist = {"stress": (3.0 * l, 3.2 * l), "physical": (1.0 * l, 4.0 / 3.0 * l)}.get(repulsion, (0.0, 4.0 / 3.0 * l))
    return dict(kappa_b=kb, kappa_a=ka, kappa_v=kv, kappa_c=kc, kappa_t=kt, kappa_r=kr,
                area_frac=fracs[0], volume_frac=fracs[1], curvature_frac=fracs[2], bond_type=bond,
                bond_lc0=1.15 * l, bond_lc1=0.85 * l, bond_a0=a, bond_r=2, rep_lc1=lc1, rep_r=2,
                n_search=n_search, rlist=rlist, exclusion_level=exclusion, lam=1.0)


class _Case:
    pass


def make_case(subdiv, repulsion="stress", n_search="verlet-list", threads=None, seed=42):
    """Benchmark case on the port (bench.py cpu_baseline when oracle/_ref is unavailable)."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from trimem_b200 import meshgen
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    t0 = time.time()
    v, f = meshgen.jittered_icosphere(subdiv, seed=seed)
    mesh = TriMesh(v, f)
    es = EnergyManager(mesh, workload_params(mesh, repulsion=repulsion, n_search=n_search))
    case = _Case()
    case.mesh, case.estore, case.setup_seconds = mesh, es, time.time() - t0

    def eg():
        return es.energy(mesh), es.gradient(mesh)
    case.energy_and_gradient = eg
    return case
