"""Worker of the multi-process partition tests (one process per rank, torch.distributed/gloo for
the rendezvous and the handle exchange only).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ... tests/part_worker.py MODE OUT.npz

MODE gpu : the real thing on cuda:(rank % device_count): partitioned energy, gradient, trajectory,
           list rebuild; rank 0 writes the results for the parent test to compare with the goldens.
MODE cpu : no device: the same plumbing (trimem_b200.partition.setup) against a recording stand-in
           for the mesh handle, so the protocol (who exports, who attaches whom, barriers) is covered
           by the CPU suite.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch.distributed as dist  # noqa: E402

from trimem_b200 import partition  # noqa: E402


def allgather(obj):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


class FakeMesh:
    """Records the partition calls; the handle encodes the exporting rank."""
    def __init__(self, rank):
        self.rank, self.calls = rank, []

    def part_create(self, rank, nranks):
        self.calls.append(("create", rank, nranks))

    def part_export(self):
        self.calls.append(("export",))
        return bytes([self.rank]) * 128

    def part_attach_ipc(self, peer, handle):
        assert len(handle) == 128 and handle[0] == peer
        self.calls.append(("attach", peer))

    def part_prepare(self, nl):
        self.calls.append(("prepare", nl))


def main_cpu(out):
    rank, world = dist.get_rank(), dist.get_world_size()
    fm = FakeMesh(rank)
    nl = partition.setup(fm, rank, world, allgather, nlist_factory=lambda m: "nl")
    assert nl == "nl"
    want = [("create", rank, world), ("export",)] + [("attach", p) for p in range(world) if p != rank] + [("prepare", "nl")]
    assert fm.calls == want, fm.calls
    b = partition.slot_bounds(655362, world)
    rows = np.zeros((10, 3))
    rows[rank::world] = rank + 1.0
    full = partition.combine_rows(allgather, rows)
    assert np.all(full[:, 0] == (np.arange(10) % world) + 1.0)
    if rank == 0:
        np.savez(out, bounds=np.array(b), ok=np.array([1]))


def main_gpu(out):
    import cases
    import trimem_b200.core as core
    from trimem_b200 import cabi, workload
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = rank % core.device_count()
    core.set_device(dev)
    name = "vv_ico4_stress_verlet2"
    case, p0, mass, dt, n = cases.vv_inputs(name)
    mesh, es = cases.build(core, case)
    v = mesh.points().copy()
    arrs = mesh.halfedge_arrays()[:4]
    rp = es.eparams.repulse_params
    par = workload.abi_params(es)
    cm = cabi.Mesh(*arrs, xyz=v, device=dev)
    nl = partition.setup(cm, rank, world, allgather,
                         lambda m: cabi.NList(m, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level))
    e, props = cm.energy(nl, par)
    g = partition.combine_rows(allgather, cm.gradient(nl, par))
    cm.set_momenta(p0)
    cm.vv_integrate_async(nl, par, mass, dt, n, with_energies=True)
    st = cm.vv_stats(par)[0]
    p = partition.combine_rows(allgather, cm.momenta())
    x = cm.positions()
    kin = np.sum(allgather(np.array([st[1], st[3]])), axis=0)
    nl2 = cabi.NList(cm, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level)
    n_entries = sum(allgather(nl2.info()[0]))          # every rank holds the rows of its own vertices
    es_all = allgather((e, tuple(props), st[0], st[2], x.tobytes()))
    assert all(t == es_all[0] for t in es_all), "ranks disagree on global results"
    if rank == 0:
        one = cabi.Mesh(*arrs, xyz=x, device=dev)      # the same list on one device
        assert cabi.NList(one, cabi.NLIST_VERLET, rp.rlist, rp.exclusion_level).info()[0] == n_entries
        np.savez(out, energy=e, props=props, grad=g, x=x, p=p, stats=np.array([st[0], kin[0], st[2], kin[1]]),
                 n_entries=n_entries, info=np.array(list(cm.part_info().values())))


if __name__ == "__main__":
    mode, out = sys.argv[1], sys.argv[2]
    dist.init_process_group("gloo")
    try:
        (main_gpu if mode == "gpu" else main_cpu)(out)
        dist.barrier()
    finally:
        dist.destroy_process_group()
