"""ctypes binding of the C ABI (include/trimem_b200.h).

This is what a non-C++ caller of libtrimem_b200.so looks like; the parity tests
and bench.py use it to call the kernels through the same boundary the pybind11
host module uses.
"""
import ctypes as C
import os
import re

import numpy as np

from . import lib_path

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class EnergyParamsABI(C.Structure):
    """struct tm_energy_params"""
    _fields_ = [(n, C.c_double) for n in
                ("kappa_b", "kappa_a", "kappa_v", "kappa_c", "kappa_t", "kappa_r",
                 "ref_area", "ref_volume", "ref_curvature",
                 "bond_lc0", "bond_lc1", "bond_a0")] + [
        ("bond_r", C.c_int), ("bond_type", C.c_int),
        ("rep_lc1", C.c_double), ("rep_r", C.c_int), ("_pad", C.c_int)]


BOND_EDGE, BOND_AREA = 0, 1
NLIST_CELL, NLIST_VERLET = 0, 1
FLIP_SERIAL, FLIP_BATCH = 0, 1

_lib = None


def header_symbols():
    """All function names include/trimem_b200.h declares."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "include", "trimem_b200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(tm_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Load libtrimem_b200.so (raises if it is not built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise OSError(f"{path} is missing: build it with `python -m trimem_b200.build`")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.tm_last_error.restype = C.c_char_p
    L.tm_host_alloc.restype = vp
    L.tm_host_alloc.argtypes = [C.c_ulonglong]
    L.tm_host_free.argtypes = [vp]
    L.tm_free.argtypes = [vp]
    L.tm_launch_count.restype = C.c_ulonglong
    L.tm_mesh_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.tm_mesh_destroy.argtypes = [vp]
    L.tm_mesh_set_topology.argtypes = [vp, C.c_int, C.c_int, _ip, _ip, _ip, _ip]
    L.tm_mesh_get_topology.argtypes = [vp, _ip, _ip, _ip, _ip]
    L.tm_mesh_get_rings.argtypes = [vp, _ip, _ip, _ip]
    L.tm_mesh_set_positions.argtypes = [vp, _dp]
    L.tm_mesh_set_positions_rows.argtypes = [vp, _dp, C.c_int, C.c_int, C.c_int]
    L.tm_mesh_get_positions.argtypes = [vp, _dp]
    L.tm_mesh_reorder.argtypes = [vp]
    L.tm_mesh_get_vertex_order.argtypes = [vp, _ip]
    L.tm_mesh_n_vertices.argtypes = [vp]
    L.tm_mesh_n_halfedges.argtypes = [vp]
    L.tm_nlist_create.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_double, C.POINTER(vp)]
    L.tm_nlist_destroy.argtypes = [vp]
    L.tm_nlist_distance_matrix.argtypes = [vp, vp, C.c_double, C.POINTER(C.c_longlong),
                                           C.POINTER(_dp), C.POINTER(_ip), C.POINTER(_ip)]
    L.tm_nlist_point_distances.argtypes = [vp, vp, C.c_int, C.c_double, _ip, C.POINTER(_dp), C.POINTER(_ip)]
    L.tm_nlist_info.argtypes = [vp, C.POINTER(C.c_longlong), _ip, _dp, _dp]
    L.tm_nlist_displacement.argtypes = [vp, vp, C.c_double, _dp, _ip]
    pp = C.POINTER(EnergyParamsABI)
    L.tm_properties.argtypes = [vp, vp, pp, _dp]
    L.tm_vertex_properties.argtypes = [vp, vp, pp, _dp]
    L.tm_energy.argtypes = [vp, vp, pp, _dp, _dp]
    L.tm_gradient.argtypes = [vp, vp, pp, _dp, _dp]
    L.tm_energy_async.argtypes = [vp, vp, pp]
    L.tm_gradient_async.argtypes = [vp, vp, pp]
    L.tm_fetch_energy.argtypes = [vp, pp, _dp, _dp]
    L.tm_fetch_gradient.argtypes = [vp, _dp]
    L.tm_sync.argtypes = [vp]
    L.tm_timer_start.argtypes = [vp]
    L.tm_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.tm_flush_l2.argtypes = [vp, C.c_ulonglong]
    L.tm_time_kernel.argtypes = [vp, vp, pp, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(C.c_float)]
    L.tm_area_edges.argtypes = [vp, _dp, _dp]
    L.tm_halfedge_geometry.argtypes = [vp, C.c_int, _dp]
    L.tm_flip_sweep.argtypes = [vp, vp, pp, C.c_int, C.c_int, _ip, _dp, _ip, _ip, _ip, _dp]
    L.tm_edge_vertex_properties.argtypes = [vp, vp, pp, C.c_int, _dp]
    L.tm_mesh_is_flip_ok.argtypes = [vp, C.c_int, _ip]
    L.tm_mesh_flip.argtypes = [vp, C.c_int]
    L.tm_mesh_set_momenta.argtypes = [vp, _dp]
    L.tm_mesh_get_momenta.argtypes = [vp, _dp]
    L.tm_vv_integrate_async.argtypes = [vp, vp, pp, C.c_double, C.c_double, C.c_int, C.c_int]
    L.tm_vv_stats.argtypes = [vp, pp, _dp, _dp, _dp]
    L.tm_vv_displacement.argtypes = [vp, vp, C.c_double, _dp, _dp]
    L.tm_mesh_random_momenta.argtypes = [vp, C.c_ulonglong, C.c_ulonglong, C.c_double]
    L.tm_vv_integrate.argtypes = [vp, vp, pp, _dp, C.c_double, C.c_double, C.c_int, _dp, _dp, _dp]
    L.tm_part_create.argtypes = [vp, C.c_int, C.c_int]
    L.tm_part_info.argtypes = [vp, _ip, _ip, _ip, _ip, _ip, _ip]
    L.tm_part_export.argtypes = [vp, C.c_char_p]
    L.tm_part_attach_ipc.argtypes = [vp, C.c_int, C.c_char_p]
    L.tm_part_attach_local.argtypes = [vp, C.c_int, vp]
    L.tm_part_prepare.argtypes = [vp, vp]
    L.tm_part_sync_positions.argtypes = [vp]
    L.tm_part_allsum6.argtypes = [vp, _dp, _dp]
    L.tm_part_check.argtypes = [vp]
    L.tm_part_debug.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    _lib = L
    return L


class TMError(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise TMError(lib().tm_last_error().decode())


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class Mesh:
    """Device mesh handle fed with flattened half-edge arrays."""

    def __init__(self, he_to, he_next, he_face, v_he, xyz=None, device=0):
        self.h = C.c_void_p()
        _chk(lib().tm_mesh_create(device, C.byref(self.h)))
        self.V = len(v_he)
        self.HE = len(he_to)
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (he_to, he_next, he_face, v_he)]
        _chk(lib().tm_mesh_set_topology(self.h, self.V, self.HE, *[_i(a) for a in arrs]))
        if xyz is not None:
            self.set_positions(xyz)

    def close(self):
        if self.h:
            lib().tm_mesh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_positions(self, xyz):
        x = np.ascontiguousarray(xyz, dtype=np.float64)
        assert x.shape == (self.V, 3)
        _chk(lib().tm_mesh_set_positions(self.h, _d(x)))
        _chk(lib().tm_sync(self.h))

    def positions(self):
        x = np.empty((self.V, 3))
        _chk(lib().tm_mesh_get_positions(self.h, _d(x)))
        return x

    def reorder(self):
        _chk(lib().tm_mesh_reorder(self.h))

    def vertex_order(self):
        o = np.empty(self.V, np.int32)
        _chk(lib().tm_mesh_get_vertex_order(self.h, _i(o)))
        return o

    def topology(self):
        a = [np.empty(self.HE, np.int32) for _ in range(3)] + [np.empty(self.V, np.int32)]
        _chk(lib().tm_mesh_get_topology(self.h, *[_i(x) for x in a]))
        return a

    def rings(self):
        off = np.empty(self.V + 1, np.int32)
        rv = np.empty(self.HE, np.int32)
        fl = np.empty(self.HE, np.int32)
        _chk(lib().tm_mesh_get_rings(self.h, _i(off), _i(rv), _i(fl)))
        return off, rv[:off[-1]], fl[:off[-1]]

    def properties(self, nl, p):
        out = np.empty(6)
        _chk(lib().tm_properties(self.h, nl.h if nl else None, C.byref(p), _d(out)))
        return out

    def vertex_properties(self, nl, p):
        out = np.empty((self.V, 6))
        _chk(lib().tm_vertex_properties(self.h, nl.h if nl else None, C.byref(p), _d(out)))
        return out

    def energy(self, nl, p):
        e = C.c_double()
        props = np.empty(6)
        _chk(lib().tm_energy(self.h, nl.h if nl else None, C.byref(p), C.byref(e), _d(props)))
        return e.value, props

    def gradient(self, nl, p):
        g = np.empty((self.V, 3))
        _chk(lib().tm_gradient(self.h, nl.h if nl else None, C.byref(p), _d(g), None))
        return g

    def energy_async(self, nl, p):
        _chk(lib().tm_energy_async(self.h, nl.h if nl else None, C.byref(p)))

    def gradient_async(self, nl, p):
        _chk(lib().tm_gradient_async(self.h, nl.h if nl else None, C.byref(p)))

    def sync(self):
        _chk(lib().tm_sync(self.h))

    def timer_start(self):
        _chk(lib().tm_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        _chk(lib().tm_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def flush_l2(self, nbytes):
        _chk(lib().tm_flush_l2(self.h, nbytes))

    def time_kernel(self, nl, p, which, iters, flush_bytes=0):
        ms = C.c_float()
        _chk(lib().tm_time_kernel(self.h, nl.h if nl else None, C.byref(p), which, iters, flush_bytes, C.byref(ms)))
        return ms.value

    def area_edges(self):
        a, l = C.c_double(), C.c_double()
        _chk(lib().tm_area_edges(self.h, C.byref(a), C.byref(l)))
        return a.value, l.value

    def halfedge_geometry(self, he):
        out = np.empty(40)
        _chk(lib().tm_halfedge_geometry(self.h, he, _d(out)))
        return out

    def flip_sweep(self, nl, p, mode, edges, u, props):
        edges = np.ascontiguousarray(edges, dtype=np.int32)
        u = np.ascontiguousarray(u, dtype=np.float64)
        n = len(edges)
        nacc, nfl = C.c_int(), C.c_int()
        flipped = np.empty(max(2 * n, 1), np.int32)
        props = np.array(props, dtype=np.float64)
        _chk(lib().tm_flip_sweep(self.h, nl.h if nl else None, C.byref(p), mode, n, _i(edges), _d(u),
                                 C.byref(nacc), C.byref(nfl), _i(flipped), _d(props)))
        return nacc.value, flipped[:nfl.value].copy(), props

    def edge_vertex_properties(self, nl, p, edge):
        out = np.empty(6)
        _chk(lib().tm_edge_vertex_properties(self.h, nl.h if nl else None, C.byref(p), edge, _d(out)))
        return out

    def is_flip_ok(self, edge):
        ok = C.c_int()
        _chk(lib().tm_mesh_is_flip_ok(self.h, edge, C.byref(ok)))
        return bool(ok.value)

    def flip(self, edge):
        _chk(lib().tm_mesh_flip(self.h, edge))

    # ---- velocity-Verlet trajectory (device resident) ----
    def set_momenta(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        assert p.shape == (self.V, 3)
        _chk(lib().tm_mesh_set_momenta(self.h, _d(p)))

    def random_momenta(self, seed, sequence, scale=1.0):
        _chk(lib().tm_mesh_random_momenta(self.h, seed, sequence, scale))

    def momenta(self):
        p = np.zeros((self.V, 3))
        _chk(lib().tm_mesh_get_momenta(self.h, _d(p)))
        return p

    def vv_integrate_async(self, nl, p, mass, dt, n_steps, with_energies=False):
        _chk(lib().tm_vv_integrate_async(self.h, nl.h if nl else None, C.byref(p), mass, dt, n_steps,
                                         1 if with_energies else 0))

    def vv_stats(self, p):
        st, p0, p1 = np.empty(4), np.empty(6), np.empty(6)
        _chk(lib().tm_vv_stats(self.h, C.byref(p), _d(st), _d(p0), _d(p1)))
        return st, p0, p1

    def vv_integrate(self, nl, p, p0, mass, dt, n_steps, stats=True):
        """-> (x_new, p_new, [U0, K0, U1, K1] or None): n_steps of hmc.py's _vv_integration from the
        mesh's current positions."""
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        assert p0.shape == (self.V, 3)
        x, pn = np.zeros((self.V, 3)), np.zeros((self.V, 3))
        st = np.empty(4) if stats else None
        _chk(lib().tm_vv_integrate(self.h, nl.h if nl else None, C.byref(p), _d(p0), mass, dt, n_steps,
                                   _d(x), _d(pn), _d(st) if stats else None))
        return x, pn, st

    # ---- one mesh partitioned across devices ----
    def part_create(self, rank, nranks):
        _chk(lib().tm_part_create(self.h, rank, nranks))

    def part_info(self):
        v = [C.c_int() for _ in range(6)]
        _chk(lib().tm_part_info(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("rank", "nranks", "first", "count", "n_ghost1", "n_pull"), [x.value for x in v]))

    def part_export(self):
        buf = C.create_string_buffer(128)
        _chk(lib().tm_part_export(self.h, buf))
        return buf.raw

    def part_attach_ipc(self, peer, handle):
        assert len(handle) == 128
        _chk(lib().tm_part_attach_ipc(self.h, peer, handle))

    def part_attach_local(self, peer, other):
        _chk(lib().tm_part_attach_local(self.h, peer, other.h))

    def part_prepare(self, nl):
        _chk(lib().tm_part_prepare(self.h, nl.h if nl else None))

    def part_sync_positions(self):
        _chk(lib().tm_part_sync_positions(self.h))

    def part_allsum6(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty(6)
        _chk(lib().tm_part_allsum6(self.h, _d(v), _d(out)))
        return out

    def part_debug(self):
        out = (C.c_ulonglong * 21)()
        _chk(lib().tm_part_debug(self.h, out))
        v = list(out)
        return dict(pos_seq=v[0], sums_seq=v[1], error=v[2], host_pos_seq=v[3], host_sums_seq=v[4], ack=v[5:])

    def part_check(self):
        _chk(lib().tm_part_check(self.h))

    def fetch_energy(self, p):
        e = C.c_double()
        props = np.empty(6)
        _chk(lib().tm_fetch_energy(self.h, C.byref(p), C.byref(e), _d(props)))
        return e.value, props

    def fetch_gradient(self):
        g = np.zeros((self.V, 3))
        _chk(lib().tm_fetch_gradient(self.h, _d(g)))
        return g


class NList:
    def __init__(self, mesh, kind, rlist, excl, box_eps=1.0e-6):
        self.h = C.c_void_p()
        self.mesh = mesh
        _chk(lib().tm_nlist_create(mesh.h, kind, rlist, excl, box_eps, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().tm_nlist_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        n = C.c_longlong()
        shape = np.zeros(3, np.int32)
        box = np.zeros(6)
        cell = np.zeros(3)
        _chk(lib().tm_nlist_info(self.h, C.byref(n), _i(shape), _d(box), _d(cell)))
        return n.value, shape, box, cell

    def displacement(self, dmax):
        """(largest displacement since the build, rows may have become incomplete for cut-off dmax)"""
        d, need = C.c_double(), C.c_int()
        _chk(lib().tm_nlist_displacement(self.h, self.mesh.h, dmax, C.byref(d), C.byref(need)))
        return d.value, bool(need.value)

    def distance_matrix(self, rdist):
        n = C.c_longlong()
        d, i, j = _dp(), _ip(), _ip()
        _chk(lib().tm_nlist_distance_matrix(self.h, self.mesh.h, rdist, C.byref(n), C.byref(d), C.byref(i), C.byref(j)))
        k = n.value
        out = (np.ctypeslib.as_array(d, (max(k, 1),))[:k].copy(),
               np.ctypeslib.as_array(i, (max(k, 1),))[:k].copy(),
               np.ctypeslib.as_array(j, (max(k, 1),))[:k].copy())
        for p in (d, i, j):
            lib().tm_free(C.cast(p, C.c_void_p))
        return out

    def point_distances(self, pid, rdist):
        n = C.c_int()
        d, j = _dp(), _ip()
        _chk(lib().tm_nlist_point_distances(self.h, self.mesh.h, pid, rdist, C.byref(n), C.byref(d), C.byref(j)))
        k = n.value
        out = (np.ctypeslib.as_array(d, (max(3 * k, 1),))[:3 * k].reshape(k, 3).copy(),
               np.ctypeslib.as_array(j, (max(k, 1),))[:k].copy())
        lib().tm_free(C.cast(d, C.c_void_p))
        lib().tm_free(C.cast(j, C.c_void_p))
        return out
