"""N>1 path with one PROCESS per rank (tests/part_worker.py under torch.distributed.run, gloo).
CPU: the host-side protocol of trimem_b200.partition with world_size 2 and 3. GPU: the real
partitioned evaluation over cudaIpc handles (both ranks share the device when the box has one
GPU), compared with the reference's golden vectors."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from trimem_b200 import partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch(mode, world, out, timeout):
    port = 29500 + (os.getpid() * 7 + world) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "part_worker.py"), mode, out]
    env = dict(os.environ, OMP_NUM_THREADS="1", TM_PART_TIMEOUT_MS="60000")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-4000:]
    return dict(np.load(out))


def test_slot_bounds():
    for V, n in [(2562, 2), (2562, 3), (655362, 8), (10485762, 8), (40, 4), (31, 2)]:
        b = partition.slot_bounds(V, n)
        assert b[0] == 0 and b[-1] == V and all(x <= y for x, y in zip(b, b[1:]))
        assert all(x % 32 == 0 for x in b[:-1])
        sizes = np.diff(b)
        assert sizes.max() - sizes.min() <= 64 or V < 64 * n


@pytest.mark.parametrize("world", [2, 3])
def test_protocol_gloo_cpu(tmp_path, world):
    res = launch("cpu", world, str(tmp_path / "out.npz"), 300)
    assert res["ok"][0] == 1 and list(res["bounds"]) == partition.slot_bounds(655362, world)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2])
def test_partitioned_over_ipc(gpu_core, tmp_path, world):
    res = launch("gpu", world, str(tmp_path / "out.npz"), 600)
    name = "vv_ico4_stress_verlet2"
    case = cases.VV_CASES[name][0]
    gold, gvv = cases.load_golden(case), cases.load_golden(name)
    v, _ = cases.mesh_of(case)
    assert abs(res["energy"] - float(gold["energy"])) <= 1e-10 * abs(float(gold["energy"]))
    assert np.all(np.abs(res["props"] - gold["props"]) <= 1e-10 * np.abs(gold["props"]))
    assert np.linalg.norm(res["grad"] - gold["grad"]) <= 1e-10 * np.linalg.norm(gold["grad"])
    assert np.linalg.norm((res["x"] - v) - (gvv["x"] - v)) <= 1e-9 * np.linalg.norm(gvv["x"] - v)
    assert np.linalg.norm(res["p"] - gvv["p"]) <= 1e-9 * np.linalg.norm(gvv["p"])
    assert np.all(np.abs(res["stats"] - gvv["stats"]) <= 1e-9 * np.abs(gvv["stats"]))
