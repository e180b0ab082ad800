"""read_mesh (src/mesh.cpp:7-16 -> OpenMesh::IO::read_mesh): STL binary / ASCII with OpenMesh's
corner welding (bitwise-equal float32 coordinates, vertices numbered by first appearance, degenerate
facets dropped), OFF, OBJ, PLY (ascii and both binary byte orders). Host-only logic: runs in the CPU suite; the docs' known answer through an
STL file (doc/python-module.rst:60-98) is checked on the GPU."""
import os
import struct

import numpy as np
import pytest

import cases
from trimem_b200 import meshgen


def write_stl_binary(path, v, f):
    v32 = v.astype(np.float32)
    with open(path, "wb") as fh:
        fh.write(b"trimem_b200 test".ljust(80, b" "))
        fh.write(struct.pack("<I", len(f)))
        for a, b, c in f:
            fh.write(struct.pack("<3f", 0.0, 0.0, 0.0))
            fh.write(v32[[a, b, c]].tobytes())
            fh.write(b"\x00\x00")


def write_stl_ascii(path, v, f):
    v32 = v.astype(np.float32)
    with open(path, "w") as fh:
        fh.write("solid test\n")
        for a, b, c in f:
            fh.write(" facet normal 0 0 0\n  outer loop\n")
            for k in (a, b, c):
                fh.write("   vertex %.9e %.9e %.9e\n" % tuple(float(t) for t in v32[k]))
            fh.write("  endloop\n endfacet\n")
        fh.write("endsolid test\n")


def first_appearance(v, f):
    """Expected welded numbering: vertices in order of first appearance over the facets."""
    order, seen = [], {}
    faces = []
    for tri in f:
        ids = []
        for k in tri:
            if k not in seen:
                seen[k] = len(order)
                order.append(k)
            ids.append(seen[k])
        faces.append(ids)
    return v.astype(np.float32).astype(np.float64)[order], np.array(faces, np.int32)


@pytest.mark.parametrize("writer", [write_stl_binary, write_stl_ascii])
def test_read_stl_welds_like_openmesh(core, tmp_path, writer):
    v, f = meshgen.icosphere(2)
    path = str(tmp_path / "m.stl")
    writer(path, v, f)
    mesh = core.read_mesh(path)
    want_v, want_f = first_appearance(v, f)
    assert mesh.n_vertices() == len(v) and mesh.n_edges() == 3 * len(v) - 6
    assert np.array_equal(mesh.points(), want_v)
    # same construction as TriMesh(points, cells): identical half-edge numbering
    ref = core.TriMesh(want_v, want_f)
    for a, b in zip(mesh.halfedge_arrays(), ref.halfedge_arrays()):
        assert np.array_equal(a, b)
    assert np.array_equal(mesh.fv_indices(), ref.fv_indices())


def test_read_stl_drops_degenerate_facets(core, tmp_path):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], dtype=np.float64)
    f = np.array([[0, 1, 2], [1, 3, 2], [1, 1, 2]])          # the last facet has a repeated corner
    path = str(tmp_path / "d.stl")
    write_stl_binary(path, v, f)
    mesh = core.read_mesh(path)
    assert mesh.n_vertices() == 4 and mesh.fv_indices().shape == (2, 3)


def test_read_off_and_obj(core, tmp_path):
    v, f = meshgen.icosphere(1)
    off = str(tmp_path / "m.off")
    with open(off, "w") as fh:
        fh.write("OFF\n# comment\n%d %d 0\n" % (len(v), len(f)))
        for p in v:
            fh.write("%.17g %.17g %.17g\n" % tuple(p))
        for t in f:
            fh.write("3 %d %d %d\n" % tuple(t))
    obj = str(tmp_path / "m.obj")
    with open(obj, "w") as fh:
        for p in v:
            fh.write("v %.17g %.17g %.17g\n" % tuple(p))
        fh.write("vn 0 0 1\n")
        for t in f:
            fh.write("f %d//1 %d//1 %d//1\n" % tuple(t + 1))
    ref = core.TriMesh(v, f)
    for path in (off, obj):
        mesh = core.read_mesh(path)
        assert np.array_equal(mesh.points(), v)
        for a, b in zip(mesh.halfedge_arrays(), ref.halfedge_arrays()):
            assert np.array_equal(a, b)
    # a quad is fan-triangulated from its first corner
    quad = str(tmp_path / "q.off")
    with open(quad, "w") as fh:
        fh.write("OFF 4 1 0\n0 0 0\n1 0 0\n1 1 0\n0 1 0\n4 0 1 2 3\n")
    assert np.array_equal(core.read_mesh(quad).fv_indices(), core.TriMesh(
        np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], float), np.array([[0, 1, 2], [0, 2, 3]])).fv_indices())


def write_ply(path, v, f, fmt="ascii", ftype="double", extra=False, index_name="vertex_indices"):
    """PLY with optional extra vertex properties (a uchar colour triple between the coordinates and a trailing
    float) and an extra face property, so that the reader has to skip by declared type."""
    end = {"ascii": "", "binary_little_endian": "<", "binary_big_endian": ">"}[fmt]
    fc = {"float": "f", "double": "d"}[ftype]
    head = ["ply", "format %s 1.0" % fmt, "comment made by the test", "element vertex %d" % len(v),
            "property %s x" % ftype]
    if extra:
        head += ["property uchar red"]
    head += ["property %s y" % ftype, "property %s z" % ftype]
    if extra:
        head += ["property float quality"]
    head += ["element face %d" % len(f)]
    if extra:
        head += ["property short flag"]
    head += ["property list uchar int %s" % index_name, "end_header"]
    with open(path, "wb") as fh:
        fh.write(("\n".join(head) + "\n").encode())
        if fmt == "ascii":
            rep = (lambda t: "%.9g" % np.float32(t)) if ftype == "float" else (lambda t: "%.17g" % t)
            for p in v:
                row = [rep(p[0])] + (["7"] if extra else []) + [rep(p[1]), rep(p[2])] + (["0.5"] if extra else [])
                fh.write((" ".join(row) + "\n").encode())
            for t in f:
                fh.write((("-3 " if extra else "") + "3 %d %d %d\n" % tuple(t)).encode())
        else:
            for p in v:
                fh.write(struct.pack(end + fc, p[0]))
                if extra:
                    fh.write(struct.pack("B", 7))
                fh.write(struct.pack(end + 2 * fc, p[1], p[2]))
                if extra:
                    fh.write(struct.pack(end + "f", 0.5))
            for t in f:
                if extra:
                    fh.write(struct.pack(end + "h", -3))
                fh.write(struct.pack(end + "B3i", 3, *[int(k) for k in t]))


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
@pytest.mark.parametrize("ftype,extra", [("double", False), ("float", True)])
def test_read_ply(core, tmp_path, fmt, ftype, extra):
    v, f = meshgen.icosphere(2)
    path = str(tmp_path / "m.ply")
    write_ply(path, v, f, fmt, ftype, extra, "vertex_index" if extra else "vertex_indices")
    mesh = core.read_mesh(path)
    want = v if ftype == "double" else v.astype(np.float32).astype(np.float64)     # float32 widened, like STL
    assert np.array_equal(mesh.points(), want)
    ref = core.TriMesh(want, f)
    for a, b in zip(mesh.halfedge_arrays(), ref.halfedge_arrays()):
        assert np.array_equal(a, b)
    assert np.array_equal(mesh.fv_indices(), ref.fv_indices())


def test_read_ply_errors(core, tmp_path):
    v, f = meshgen.icosphere(1)
    good = tmp_path / "g.ply"
    write_ply(str(good), v, f, "binary_little_endian")
    data = good.read_bytes()
    (tmp_path / "trunc.ply").write_bytes(data[:-7])
    (tmp_path / "nohead.ply").write_bytes(data.replace(b"end_header\n", b""))
    bad_index = tmp_path / "idx.ply"
    write_ply(str(bad_index), v, np.array([[0, 1, len(v)]]), "ascii")
    for name in ("trunc.ply", "nohead.ply", "idx.ply"):
        with pytest.raises(RuntimeError, match="read error on file"):
            core.read_mesh(str(tmp_path / name))


def test_read_errors(core, tmp_path):
    with pytest.raises(RuntimeError, match="read error on file"):
        core.read_mesh(str(tmp_path / "missing.stl"))
    bad = tmp_path / "bad.xyz"
    bad.write_text("nothing")
    with pytest.raises(RuntimeError, match="read error on file"):
        core.read_mesh(str(bad))
    trunc = tmp_path / "trunc.stl"
    trunc.write_bytes(b" " * 80 + struct.pack("<I", 10) + b"\x00" * 20)
    with pytest.raises(RuntimeError, match="read error on file"):
        core.read_mesh(str(trunc))


@pytest.mark.gpu
def test_docs_example_through_stl(gpu_core, tmp_path):
    """doc/python-module.rst:60-98: icosphere(3) written to a binary STL (float32), read back, energy
    1000755.3602085959 and the printed gradient rows."""
    gold = cases.load_golden("analytic")
    v, f = meshgen.icosphere(3, project_each_round=False)
    path = str(tmp_path / "sphere.stl")
    write_stl_binary(path, v, f)
    mesh = gpu_core.read_mesh(path)
    ep = gpu_core.EnergyParams()
    ep.kappa_b, ep.kappa_a, ep.kappa_v, ep.kappa_c, ep.kappa_t, ep.kappa_r = 30.0, 1e6, 1e6, 1e6, 1e5, 1e3
    ep.area_frac = 0.5
    es = gpu_core.EnergyManager(mesh, ep)
    e = es.energy(mesh)
    assert abs(e - 1000755.3602085959) <= 1e-10 * 1000755.3602085959
    assert abs(e - float(gold["docs_energy"])) <= 1e-10 * e
    # the welded numbering permutes the rows relative to the generator's: compare through the positions
    g = es.gradient(mesh)
    key = lambda a: [tuple(r) for r in a.astype(np.float32)]
    order = {k: i for i, k in enumerate(key(v))}
    perm = np.array([order[k] for k in key(mesh.points())])
    assert np.linalg.norm(g - gold["docs_grad"][perm]) <= 1e-10 * np.linalg.norm(gold["docs_grad"])
