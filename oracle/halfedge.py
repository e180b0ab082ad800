"""TEST INFRASTRUCTURE ONLY (oracle). Half-edge arrays with OpenMesh's numbering.

Restates what ``from_points_cells`` + OpenMesh ``add_face`` produce (reference
src/mesh_py.cpp:12-70; OpenMesh ArrayKernel semantics):
edge k <-> half-edges 2k/2k+1 numbered by first appearance while faces are added
in input order, opposite(h) = h ^ 1, halfedge(face) = v[2] -> v[0], clockwise
circulation h -> next(opposite(h)).

``build`` is an exact, sequential restatement of add_face (any manifold input,
open or closed; pure Python, use it for small meshes). ``build_closed`` is a
vectorised numpy construction for closed manifold meshes: identical edge /
half-edge / face numbering, but every vertex starts its circulation at its
lowest-numbered outgoing half-edge (OpenMesh's choice depends on the insertion
history) -- that only permutes the order of floating-point summation.
"""
import numpy as np


def build(n_vertices, faces):
    to, nxt, prv, face = [], [], [], []
    v_he = [-1] * n_vertices
    f_he = []

    def is_boundary_v(v):
        h = v_he[v]
        return not (h >= 0 and face[h] >= 0)

    def find(a, b):
        h0 = v_he[a]
        if h0 < 0:
            return -1
        h = h0
        while True:
            if to[h] == b:
                return h
            h = nxt[h ^ 1]
            if h == h0:
                return -1

    def set_next(h, n):
        nxt[h] = n
        prv[n] = h

    for f in faces:
        v = [int(f[0]), int(f[1]), int(f[2])]
        if min(v) < 0 or max(v) >= n_vertices:
            continue
        eh, new = [0, 0, 0], [False] * 3
        ok = True
        for i in range(3):
            if not is_boundary_v(v[i]):
                ok = False
                break
            eh[i] = find(v[i], v[(i + 1) % 3])
            new[i] = eh[i] < 0
            if not new[i] and face[eh[i]] >= 0:
                ok = False
                break
        if not ok:
            continue
        relink = []
        bad = False
        for i in range(3):
            ii = (i + 1) % 3
            if new[i] or new[ii]:
                continue
            ip, inx = eh[i], eh[ii]
            if nxt[ip] == inx:
                continue
            bp = inx ^ 1
            while True:
                bp = nxt[bp] ^ 1
                if face[bp] < 0:
                    break
            bn = nxt[bp]
            if bp == ip:
                bad = True
                break
            relink += [(bp, nxt[ip]), (prv[inx], bn), (ip, inx)]
        if bad:
            continue
        for i in range(3):
            if new[i]:
                h = len(to)
                to += [v[(i + 1) % 3], v[i]]
                nxt += [-1, -1]
                prv += [-1, -1]
                face += [-1, -1]
                eh[i] = h
        fh = len(f_he)
        f_he.append(eh[2])
        adjust = [False] * 3
        for i in range(3):
            ii = (i + 1) % 3
            vh = v[ii]
            ip, inx = eh[i], eh[ii]
            idc = (1 if new[i] else 0) | (2 if new[ii] else 0)
            if idc:
                op, on = inx ^ 1, ip ^ 1
                if idc == 1:
                    relink.append((prv[inx], on))
                    v_he[vh] = on
                elif idc == 2:
                    bn = nxt[ip]
                    relink.append((op, bn))
                    v_he[vh] = bn
                elif v_he[vh] < 0:
                    v_he[vh] = on
                    relink.append((op, on))
                else:
                    bn = v_he[vh]
                    relink += [(prv[bn], on), (op, bn)]
                relink.append((ip, inx))
            else:
                adjust[ii] = v_he[vh] == inx
            face[eh[i]] = fh
        for a, b in relink:
            set_next(a, b)
        for i in range(3):
            if adjust[i]:
                h0 = v_he[v[i]]
                h = h0
                while True:
                    if face[h] < 0:
                        v_he[v[i]] = h
                        break
                    h = nxt[h ^ 1]
                    if h == h0:
                        break
    i32 = lambda a: np.asarray(a, dtype=np.int32)
    return i32(to), i32(nxt), i32(face), i32(v_he), i32(f_he)


def build_closed(n_vertices, faces):
    f = np.asarray(faces, dtype=np.int64)
    nf = len(f)
    # directed edges in add_face creation order: (v0,v1), (v1,v2), (v2,v0) per face
    a = f[:, [0, 1, 2]].ravel()
    b = f[:, [1, 2, 0]].ravel()
    key = np.minimum(a, b) * n_vertices + np.maximum(a, b)
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")           # edges by first appearance
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    eid = rank[inv]                                    # edge id of every directed edge
    is_first = np.zeros(3 * nf, bool)
    is_first[first] = True
    he = 2 * eid + (~is_first).astype(np.int64)        # 2k for the creating direction
    nhe = 2 * len(order)
    if np.bincount(he, minlength=nhe).max() != 1 or len(he) != nhe:
        raise ValueError("build_closed needs a closed, consistently oriented manifold mesh")
    to = np.empty(nhe, np.int64)
    to[he] = b
    he3 = he.reshape(nf, 3)
    nxt = np.empty(nhe, np.int64)
    nxt[he3] = he3[:, [1, 2, 0]]
    face = np.empty(nhe, np.int64)
    face[he3] = np.arange(nf)[:, None]
    f_he = he3[:, 2]
    frm = to[np.arange(nhe) ^ 1]
    v_he = np.full(n_vertices, nhe, np.int64)
    np.minimum.at(v_he, frm, np.arange(nhe))
    v_he[v_he == nhe] = -1
    i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
    return i32(to), i32(nxt), i32(face), i32(v_he), i32(f_he)
