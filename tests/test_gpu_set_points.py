"""TriMesh.set_points(x) (extension): positions copied into the mesh memory and uploaded in one call, after which
energy()/gradient() skip their own re-upload until points() hands out a writable view again or a trajectory moves the
device positions. It must give exactly what the reference's idiom np.copyto(mesh.points(), x) (src/trimem/mc/mesh.py:39-42)
gives, and never hide a later write through points()."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["ico4_stress_verlet2", "ico4_stress_cell2", "ico3_area"])
def test_set_points_equals_copyto(gpu_core, case):
    mesh_a, es_a = cases.build(gpu_core, case)
    mesh_b, es_b = cases.build(gpu_core, case)
    x0 = mesh_a.points().copy()
    rng = np.random.default_rng(5)
    amp = 1e-3 * np.linalg.norm(x0[0] - x0[1])
    for it in range(4):
        x = x0 + amp * rng.standard_normal(x0.shape)
        np.copyto(mesh_a.points(), x)
        mesh_b.set_points(x)
        assert es_a.energy(mesh_a) == es_b.energy(mesh_b)
        assert np.array_equal(es_a.gradient(mesh_a), es_b.gradient(mesh_b))
        assert np.array_equal(mesh_b.points(), x)                  # the mesh memory holds the new positions too
    # a write through points() after set_points is honoured
    x = x0 + amp * rng.standard_normal(x0.shape)
    mesh_b.set_points(x0)
    e0 = es_b.energy(mesh_b)
    np.copyto(mesh_b.points(), x)
    np.copyto(mesh_a.points(), x)
    e1 = es_b.energy(mesh_b)
    assert e1 == es_a.energy(mesh_a) and e1 != e0
    assert np.array_equal(es_a.gradient(mesh_a), es_b.gradient(mesh_b))
    # a trajectory moves the device positions: the next evaluation is at the mesh's own positions again
    mesh_b.set_points(x0)
    p0 = np.random.default_rng(1).standard_normal(x0.shape)
    gpu_core.vv_integrate(mesh_b, es_b, p0, 1.0, 1.0e-6, 3)
    np.copyto(mesh_a.points(), x0)
    assert es_b.energy(mesh_b) == es_a.energy(mesh_a)
    assert np.array_equal(es_a.gradient(mesh_a), es_b.gradient(mesh_b))
    with pytest.raises(RuntimeError):
        mesh_b.set_points(x0[:-1])
