"""Host-side plumbing of a mesh partitioned across ranks (include/trimem_b200.h, tm_part_*).

One process per GPU: every rank creates the same mesh, calls ``setup`` with a function that
all-gathers bytes (``torch.distributed.all_gather_object`` in bench.py and the tests), and from
then on makes the same sequence of evaluation calls as its peers. The data path never comes
back here: halo positions and the six reduced sums move between the devices inside the kernels'
own stream (csrc/tm_part.cu)."""
from . import cabi


def slot_bounds(n_vertices, nranks):
    """Device-slot range [bounds[r], bounds[r+1]) rank r owns: equal numbers of 32-vertex warps
    (the mirror of tm_part_create; a warp never straddles a cut)."""
    nw = (n_vertices + 31) // 32
    b = [min(n_vertices, 32 * (nw * r // nranks)) for r in range(nranks + 1)]
    b[nranks] = n_vertices
    return b


def setup(mesh, rank, nranks, allgather, nlist_factory=None):
    """Partition ``mesh`` (a cabi.Mesh holding the same topology and positions on every rank).

    allgather(obj) -> list of every rank's obj, in rank order.
    nlist_factory(mesh) -> cabi.NList or None; built after the peers are attached.
    Returns the neighbour list."""
    if not 0 <= rank < nranks:
        raise ValueError(f"rank {rank} out of range for {nranks} ranks")
    mesh.part_create(rank, nranks)
    handles = allgather(mesh.part_export())
    if len(handles) != nranks:
        raise RuntimeError(f"all-gather returned {len(handles)} handles for {nranks} ranks")
    for peer, h in enumerate(handles):
        if peer != rank:
            mesh.part_attach_ipc(peer, h)
    nl = nlist_factory(mesh) if nlist_factory else None
    mesh.part_prepare(nl)
    # nobody starts a collective call before every rank has finished its local preparation
    allgather(b"ready")
    return nl


def setup_core(mesh, estore, rank, nranks, allgather):
    """The same for the drop-in classes: ``mesh`` a ``core.TriMesh`` holding the same topology and positions on
    every rank, ``estore`` its ``core.EnergyManager`` (built before the mesh is partitioned: reference sums and
    neighbour list of the whole mesh). Afterwards ``estore.energy / gradient / update_repulsion``, ``core.vv_integrate``
    and ``hmc.MeshHMC.step`` are collective: every rank makes the same calls (seed the chains identically)."""
    if not 0 <= rank < nranks:
        raise ValueError(f"rank {rank} out of range for {nranks} ranks")
    mesh.part_create(rank, nranks)
    handles = allgather(mesh.part_export())
    for peer, h in enumerate(handles):
        if peer != rank:
            mesh.part_attach_ipc(peer, h)
    estore.part_prepare(mesh)
    allgather(b"ready")


def combine_rows(allgather, rows):
    """Full (V,3) array from the ranks' owned rows (each rank reports zeros elsewhere)."""
    parts = allgather(rows)
    out = parts[0].copy()
    for p in parts[1:]:
        out += p
    return out
