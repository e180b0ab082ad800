"""Single half-edge primitives on the device (reference tests/test_properties.py)."""
import numpy as np
import pytest

import cases
from trimem_b200 import meshgen

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,alpha", [("p45", np.pi / 4), ("m45", -np.pi / 4)])
def test_hinge(gpu_core, tag, alpha):
    m = gpu_core
    gold = cases.load_golden("primitives")
    x, z = np.cos(alpha) * 0.5, np.sin(alpha) * 0.5
    pts = np.array([[0.0, -0.5, 0.0], [0.0, 0.5, 0.0], [-x, 0.0, z], [x, 0.0, z]])
    mesh = m.TriMesh(pts, np.array([[0, 1, 2], [1, 0, 3]]))
    e = mesh.halfedge_handle(0)
    s = gold[f"{tag}_scalars"]
    assert abs(m.edge_length(mesh, e) - s[0]) / s[0] < 1e-10
    assert abs(m.face_area(mesh, e) - s[1]) / s[1] < 1e-10
    assert abs(m.face_volume(mesh, e) - s[2]) < 1e-10
    assert abs(m.dihedral_angle(mesh, e) - s[3]) < 1e-10
    for fn, key, shape in ((m.edge_length_grad, "d_length", (2, 3)), (m.face_area_grad, "d_area", (3, 3)),
                           (m.face_volume_grad, "d_volume", (3, 3)), (m.dihedral_angle_grad, "d_angle", (4, 3))):
        got, want = fn(mesh, e), gold[f"{tag}_{key}"]
        assert got.shape == shape
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-10


def test_every_halfedge_of_a_rough_sphere(gpu_core):
    from trimem_b200 import cabi
    gold = cases.load_golden("primitives")["ico2_geometry"]
    v, f = meshgen.jittered_icosphere(2, sigma_rel=0.1, seed=3)
    mesh = gpu_core.TriMesh(v, f)
    he_to, he_next, he_face, v_he, _ = mesh.halfedge_arrays()
    cm = cabi.Mesh(he_to, he_next, he_face, v_he, xyz=v)
    geo = np.array([cm.halfedge_geometry(h) for h in range(len(he_to))])
    scale = np.abs(gold).max(axis=0) + 1e-300
    assert np.max(np.abs(geo - gold) / scale) < 1e-10


@pytest.mark.parametrize("name", ["rect", "fan"])
def test_every_halfedge_of_open_meshes(gpu_core, name):
    """Every half-edge of open meshes, boundary half-edges included, against the reference build: the
    reference evaluates those on the boundary loop's sectors as OpenMesh defines them (prev / next along
    the loop, sector angles signed against the face across the edge)."""
    from trimem_b200 import cabi
    gold = cases.load_golden("open_meshes")
    want, on_b = gold[f"{name}_geometry"], gold[f"{name}_on_boundary_edge"]
    v, f = cases.open_mesh(name)
    mesh = gpu_core.TriMesh(v, f)
    he_to, he_next, he_face, v_he, _ = mesh.halfedge_arrays()
    cm = cabi.Mesh(he_to, he_next, he_face, v_he, xyz=v)
    geo = np.array([cm.halfedge_geometry(h) for h in range(len(he_to))])
    scale = np.abs(want).max(axis=0) + 1e-300
    err = np.abs(geo - want) / scale
    assert np.max(err[~on_b]) < 1e-10
    worst = err[on_b].max(axis=0)
    assert np.max(err[on_b]) < 1e-10, np.nonzero(worst > 1e-10)[0].tolist()
