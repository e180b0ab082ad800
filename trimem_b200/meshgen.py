"""Synthetic triangle meshes used by the tests and by bench.py (numpy only).

The reference's tests take their meshes from ``trimesh`` (absent here) and from
two small generators (reference tests/util.py:5-68, a tube and a rectangle).
These are our own generators producing the same vertex/face *ordering* as the
ones the reference's golden values were recorded on:

* ``icosahedron``   12-vertex / 20-face table.
* ``icosphere(n)``  n rounds of 1->4 midpoint subdivision; original vertices
  first, midpoints appended in sorted-unique-edge order; vertices re-projected
  to the sphere after every round (``project_each_round=False`` projects once
  at the end, which is what the docs example of the reference used).
* ``tube`` / ``rect``  open meshes used for the boundary cases.
"""
import numpy as np


def icosahedron():
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array(
        [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
         [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1)[:, None]
    f = np.array(
        [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
         [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
         [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
         [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int32)
    return v, f


def subdivide(v, f):
    """One round of 1->4 midpoint subdivision."""
    nv = len(v)
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    key = e[:, 0].astype(np.int64) * nv + e[:, 1]
    ukey, inv = np.unique(key, return_inverse=True)
    ue = np.stack([ukey // nv, ukey % nv], axis=1)
    mid = 0.5 * (v[ue[:, 0]] + v[ue[:, 1]])
    nf = len(f)
    ab = nv + inv[:nf]
    bc = nv + inv[nf:2 * nf]
    ca = nv + inv[2 * nf:]
    a, b, c = f[:, 0], f[:, 1], f[:, 2]
    nfaces = np.column_stack([a, ab, ca, ab, b, bc, ca, bc, c, ab, bc, ca]).reshape(-1, 3)
    return np.vstack([v, mid]), nfaces.astype(np.int32)


def icosphere(n, project_each_round=True):
    v, f = icosahedron()
    for _ in range(n):
        v, f = subdivide(v, f)
        if project_each_round:
            v /= np.linalg.norm(v, axis=1)[:, None]
    if not project_each_round:
        v /= np.linalg.norm(v, axis=1)[:, None]
    return np.ascontiguousarray(v), np.ascontiguousarray(f)


def tube(length=1.0, radius=1.0, segments=30):
    """Open tube around the y axis (same ordering as reference tests/util.py:5)."""
    circ = 2 * np.pi * radius
    n = int(length // (circ / segments)) + 1
    ang = np.linspace(0, 2 * np.pi, segments + 1)[:-1]
    y = np.linspace(-length / 2, length / 2, n)
    x, z = np.cos(ang) * radius, np.sin(ang) * radius
    verts = np.column_stack([np.tile(x, n), np.repeat(y, segments), np.tile(z, n)])
    i = np.arange(segments)
    ip = (i + 1) % segments
    ring = np.vstack([np.column_stack([i, i + segments, ip]),
                      np.column_stack([ip, i + segments, ip + segments])])
    faces = np.vstack([ring + k * segments for k in range(n - 1)])
    return np.ascontiguousarray(verts, dtype=np.float64), faces.astype(np.int32)


def rect(nx, ny):
    """Flat open rectangle in z=0 (same ordering as reference tests/util.py:46)."""
    xx, yy = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny))
    verts = np.column_stack([xx.ravel(), yy.ravel(), np.zeros(nx * ny)])
    i = np.arange(nx - 1)
    row = np.vstack([np.column_stack([i, i + 1, i + 1 + nx]),
                     np.column_stack([i + nx, i, i + 1 + nx])])
    faces = np.vstack([row + k * nx for k in range(ny - 1)])
    return np.ascontiguousarray(verts, dtype=np.float64), faces.astype(np.int32)


def mean_tri_props(v, f):
    """(mean face area, mean edge length) as reference src/util.cpp:37-41."""
    p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    area = 0.5 * np.linalg.norm(np.cross(p1 - p0, p2 - p0), axis=1)
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    # unique undirected edges in lexicographic order (an int64 key sorts like np.unique(e, axis=0)
    # and is ~10x faster at 10M vertices)
    nv = len(v)
    ukey = np.unique(e[:, 0].astype(np.int64) * nv + e[:, 1])
    ln = np.linalg.norm(v[ukey // nv] - v[ukey % nv], axis=1)
    return area.mean(), ln.mean()


def jittered_icosphere(n, sigma_rel=0.02, seed=42):
    """The bench/parity mesh of SURVEY.md section 8(d): unit icosphere plus
    N(0, (sigma_rel * mean_edge)^2) noise on every coordinate."""
    v, f = icosphere(n)
    _, l = mean_tri_props(v, f)
    rng = np.random.default_rng(seed)
    v = v + rng.normal(0.0, sigma_rel * l, size=v.shape)
    return np.ascontiguousarray(v), f
