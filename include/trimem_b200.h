/* trimem_b200.h — C ABI of libtrimem_b200.so
 *
 * B200 (sm_100a) implementation of trimem's force-evaluation hot path. This is
 * the drop-in boundary: plain pointers and sizes, no C++/torch types. The
 * pybind11 module trimem_b200.core (host C++, mirrors the reference's
 * trimem.core) and the benchmark both sit on top of exactly these entry points.
 *
 * Each entry cites the reference interface it replaces (paths relative to the
 * reference tree, bio-phys/trimem @ 1a2a05c).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; tm_last_error()
 *     returns a thread-local message for the last failing call.
 *   - all reals are fp64, all indices int32, positions are (V,3) row-major,
 *     exactly the memory TriMesh.points() exposes (src/mesh_py.cpp:96-106).
 *   - vertex / half-edge / edge numbering at this boundary is the host's
 *     (OpenMesh ArrayKernel numbering: edge k <-> half-edges 2k, 2k+1).
 *   - calls are synchronous at the boundary unless suffixed _async.
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     fails with a message; nothing is silently evaluated on the host.
 */
#ifndef TRIMEM_B200_H
#define TRIMEM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tm_mesh  tm_mesh;   /* device mirror of one TriMesh              */
typedef struct tm_nlist tm_nlist;  /* device neighbour list (cell / verlet)      */

/* ---- library ------------------------------------------------------------- */
const char* tm_last_error(void);
/* 1: energy / gradient / neighbour lists / flips; 2: + trajectories (tm_vv_*), partitioned meshes
 * (tm_part_*), tm_nlist_displacement; 3: + tm_vv_displacement (rebuild trigger measured behind the
 * trajectory), tm_mesh_random_momenta, tm_time_kernel composites 4/5; 4: + tm_mesh_set_positions_rows (upload in
 * pieces). Additions only: a caller of version 1 keeps working.                                                    */
int  tm_abi_version(void);
/* number of visible CUDA devices (0 when there is no driver / device)         */
int  tm_device_count(void);
/* pinned host memory for TriMesh.points() storage; falls back to malloc when
 * there is no device (host-only unit tests).                                   */
void* tm_host_alloc(unsigned long long bytes);
void  tm_host_free(void* p);
/* frees arrays returned by the tm_nlist_* query functions                      */
void  tm_free(void* p);
/* number of kernels this library has launched so far (bench: gpu_launches)     */
unsigned long long tm_launch_count(void);

/* ---- mesh mirror  (replaces OpenMesh TriMesh_ArrayKernelT storage,
 *      src/mesh.h:15-31; flattened to SoA + CSR one-ring on the device) ------ */
int tm_mesh_create(int device, tm_mesh** out);
int tm_mesh_destroy(tm_mesh* m);
/* Upload the half-edge connectivity (after construction and after host-side
 * topology changes). he_next must be defined for boundary half-edges too
 * (boundary loops), he_face = -1 for boundary half-edges, v_he = outgoing
 * half-edge at which the clockwise one-ring circulation starts (-1: isolated). */
int tm_mesh_set_topology(tm_mesh* m, int n_vertices, int n_halfedges,
                         const int* he_to, const int* he_next,
                         const int* he_face, const int* v_he);
/* Download the device connectivity (device-side flips mutate it).              */
int tm_mesh_get_topology(tm_mesh* m, int* he_to, int* he_next, int* he_face, int* v_he);
/* CSR one-ring the kernels use: ring_off[V+1]; ring_v[n_halfedges] holds the
 * neighbour ids in clockwise circulation order (tests / diagnostics).          */
int tm_mesh_get_rings(tm_mesh* m, int* ring_off, int* ring_v, int* ring_flags);
/* positions: the writable view of src/mesh_py.cpp:96-106 is host memory that
 * Python mutates behind our back, so every evaluation re-uploads.              */
int tm_mesh_set_positions(tm_mesh* m, const double* xyz);
int tm_mesh_get_positions(tm_mesh* m, double* xyz);
/* Extension (ABI 4): the same upload in pieces, so that a caller who is still filling its (pinned) array can hand
 * over the rows that are ready: rows [first, first + count) of the (V,3) array `xyz` (the pointer is the array's
 * base). Pieces in any number, `last` != 0 on the final one -- only then do the new positions count as uploaded;
 * no compute entry in between. (TriMesh.set_points of the host layer: copy and upload overlapped.)              */
int tm_mesh_set_positions_rows(tm_mesh* m, const double* xyz, int first, int count, int last);
/* Extension (no reference counterpart): the device stores vertices sorted by the Morton code
 * of a coarse grid cell so that warps work on compact surface patches; the order is built from
 * the first positions a mesh receives. All ids at this boundary stay in the caller's numbering.
 * tm_mesh_reorder recomputes the order from the current positions (neighbour lists built before
 * the call are refused afterwards: rebuild them); tm_mesh_get_vertex_order returns, per device
 * slot, the caller's vertex id stored there (V ints; diagnostics).                              */
int tm_mesh_reorder(tm_mesh* m);
int tm_mesh_get_vertex_order(tm_mesh* m, int* slot_to_vertex);
int tm_mesh_n_vertices(const tm_mesh* m);
int tm_mesh_n_halfedges(const tm_mesh* m);

/* ---- parameters  (POD mirror of src/params.h:19-71) ------------------------ */
enum { TM_BOND_EDGE = 0, TM_BOND_AREA = 1, TM_BOND_NONE = 2 };
enum { TM_NLIST_CELL = 0, TM_NLIST_VERLET = 1 };
typedef struct tm_energy_params {
    double kappa_b, kappa_a, kappa_v, kappa_c, kappa_t, kappa_r;
    /* reference values interpolate_reference_properties() yields
     * (src/energy.cpp:30-45); the host computes them, the device only uses them */
    double ref_area, ref_volume, ref_curvature;
    /* BondParams (src/params.h:19-26) */
    double bond_lc0, bond_lc1, bond_a0;
    int    bond_r, bond_type;
    /* SurfaceRepulsionParams.{lc1,r} (src/params.h:34-41) */
    double rep_lc1;
    int    rep_r, _pad;
} tm_energy_params;

/* ---- neighbour lists  (src/nlists/nlist.cpp:10-53, cell_list.h, verlet_list.h) */
/* make_nlist: kind x exclusion_level in {0,1,2}; built from the mesh's current
 * device positions and connectivity.                                            */
int tm_nlist_create(tm_mesh* m, int kind, double rlist, int exclusion_level,
                    double box_eps, tm_nlist** out);
int tm_nlist_destroy(tm_nlist* nl);
/* NeighbourList::distance_matrix (cell_list.h:221-312 / verlet_list.h:37-72):
 * COO triplets; cell-list: each pair once with i<j and dist <= rdist;
 * verlet-list: both directions and dist < rdist. Output is sorted by (i,j);
 * arrays are library-allocated, release with tm_free.                           */
int tm_nlist_distance_matrix(tm_nlist* nl, tm_mesh* m, double rdist,
                             long long* n, double** dist, int** idx, int** jdx);
/* NeighbourList::point_distances (cell_list.h:131-219 / verlet_list.h:74-101):
 * d_ij = x_pid - x_j (n x 3) and j for all listed j with |d_ij| < rdist.        */
int tm_nlist_point_distances(tm_nlist* nl, tm_mesh* m, int pid, double rdist,
                             int* n, double** dxyz, int** jdx);
/* Skin criterion of a verlet list (extension; the reference rebuilds on a fixed
 * cadence, mc/evaluators.py:181-182): the frozen rows stay complete for the cut-off
 * dmax (the repulsion's lc1) while no vertex has moved further than
 * (rlist - dmax) / 2 since the build. displacement = the largest such move on the
 * mesh's current device positions; needs_rebuild = 2 * displacement > rlist - dmax. */
int tm_nlist_displacement(tm_nlist* nl, tm_mesh* m, double dmax, double* displacement, int* needs_rebuild);
/* diagnostics: number of directed verlet entries (0 for a cell list), cells     */
int tm_nlist_info(tm_nlist* nl, long long* n_entries, int shape[3], double box[6], double cell[3]);

/* ---- energy / gradient  (src/energy.cpp:63-124, kernel.h, mesh_properties.cpp) */
/* EnergyManager::properties: the six reduced sums
 * (area, volume, curvature, bending, tethering, repulsion).                     */
int tm_properties(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, double out[6]);
/* per-vertex vertex_properties (mesh_properties.cpp:12-49): out is (V,6)        */
int tm_vertex_properties(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, double* out);
/* EnergyManager::energy(mesh) = trimem_energy(properties) (kernel.h:110-122)    */
int tm_energy(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, double* energy, double props[6]);
/* EnergyManager::gradient (energy.cpp:93-124): grad is (V,3) host memory; props
 * (optional) receives the reduced sums the gradient was evaluated with.         */
int tm_gradient(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, double* grad, double props[6]);

/* device-resident variants: positions already on the device (tm_mesh_set_positions
 * or a previous device update); results stay on the device until fetched.       */
int tm_energy_async(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p);
int tm_gradient_async(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p);
int tm_fetch_energy(tm_mesh* m, const tm_energy_params* p, double* energy, double props[6]);
int tm_fetch_gradient(tm_mesh* m, double* grad);
int tm_sync(tm_mesh* m);
/* CUDA-event timing on the mesh's stream (bench.py)                             */
int tm_timer_start(tm_mesh* m);
int tm_timer_stop(tm_mesh* m, float* ms);
/* write `bytes` of scratch on the mesh's stream to evict L2 between timed runs  */
int tm_flush_l2(tm_mesh* m, unsigned long long bytes);

/* average device time (CUDA events on the mesh's stream) of `iters` launches of ONE sweep set:
 * which = 0 the energy() sweeps (ring properties + pair sweep + reduction), 1 the gradient's own
 * property sweep (+ reduction), 2 the ring gradient sweep (pair force already evaluated), 3 the
 * pair sweep alone, 4 energy + gradient (all sweeps), 5 one gradient evaluation at new positions
 * (geometry sweep, pair force, ring gradient: what a leapfrog step costs); flush_bytes > 0 evicts
 * L2 before every launch (outside the timed span).                                              */
int tm_time_kernel(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, int which, int iters,
                   unsigned long long flush_bytes, float* ms_avg);

/* ---- setup reductions  (src/util.cpp:9-41) --------------------------------- */
/* area(mesh) = sum of face areas, edges_length(mesh) = sum over edges (each edge
 * measured on its half-edge 2k, as util.cpp:29-31)                              */
int tm_area_edges(tm_mesh* m, double* area, double* edges_length);

/* ---- single half-edge primitives  (src/mesh_util.h:15-185; exposed to Python
 *      by src/bindings.cpp:249-407 and pinned by tests/test_properties.py) ------
 * out[0..3]   edge_length, face_area, face_volume, dihedral_angle
 * out[4..9]   edge_length_grad     rows: from-vertex, to-vertex
 * out[10..18] face_area_grad       rows: from, to, apex
 * out[19..27] face_volume_grad     rows: from, to, apex
 * out[28..39] dihedral_angle_grad  rows: from, to, apex, opposite apex             */
int tm_halfedge_geometry(tm_mesh* m, int halfedge, double out[40]);

/* ---- edge flips  (src/flips.cpp:22-165, flip_utils.cpp:19-114) --------------- */
enum { TM_FLIP_SERIAL = 0, TM_FLIP_BATCH = 1 };
/* One sweep over n_proposals proposed edges with their uniform variates u[k]
 * (the host draws them; std::mt19937 in the reference, flips.cpp:20,105-106).
 * TM_FLIP_SERIAL evaluates proposals strictly in order (flip_serial);
 * TM_FLIP_BATCH processes them in conflict-free rounds (flip_parallel_batches),
 * Metropolis on the device in proposal order inside a round.
 * flipped[0..*n_flips) receives, in application order, the edge id of EVERY flip applied to
 * the connectivity: each trial flip and, for a rejected trial, the flip back (the reference
 * also flips back, flips.cpp:72,157, which does not restore half-edge orientation). Replaying
 * that list on the host half-edge mirror keeps it identical to the device. Capacity must be
 * 2 * n_proposals. props_io: reduced properties before (in) / after (out) the sweep.          */
int tm_flip_sweep(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, int mode,
                  int n_proposals, const int* edges, const double* u,
                  int* n_accepted, int* n_flips, int* flipped, double props_io[6]);
/* edge_vertex_properties (flip_utils.cpp:19-41) of one edge on the device mesh  */
int tm_edge_vertex_properties(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p,
                              int edge, double out[6]);
/* TriConnectivity::is_flip_ok / flip on the device connectivity (tests)         */
int tm_mesh_is_flip_ok(tm_mesh* m, int edge, int* ok);
int tm_mesh_flip(tm_mesh* m, int edge);

/* ---- velocity-Verlet trajectory, device resident   ---------------
 * Replaces _vv_integration (src/trimem/mc/hmc.py:14-26) and the two Hamiltonian
 * evaluations of HMC._step (hmc.py:129-155): positions, momenta and gradients stay
 * in HBM for the n_steps steps; the neighbour list is fixed during a trajectory, as
 * in the reference (update_repulsion runs in the per-move callback,
 * mc/evaluators.py:181-182). Momenta are (V,3) in the caller's vertex numbering.   */
int tm_mesh_set_momenta(tm_mesh* m, const double* p);
int tm_mesh_get_momenta(tm_mesh* m, double* p);
/* Extension: momenta <- scale * N(0, 1) drawn ON THE DEVICE (the reference draws them with numpy on the host,
 * src/trimem/mc/hmc.py:135-137; with the trajectory on the device that draw is two thirds of a move). Philox4x32-10
 * keyed by `seed`, counter = (vertex id, half, sequence), two 53-bit uniforms -> Box-Muller: the values depend on
 * (seed, sequence, vertex) only, so a chain is reproducible from its seed and move numbers. Asynchronous.
 * tm_vv_integrate(..., p0 = NULL, ...) then integrates from the momenta on the device.                         */
int tm_mesh_random_momenta(tm_mesh* m, unsigned long long seed, unsigned long long sequence, double scale);
/* x, p <- n_steps velocity-Verlet steps from the mesh's current device positions and
 * momenta, force = -gradient of the functional `params` describes; asynchronous.
 * with_energies != 0 also evaluates U(x0), U(xN) (the full sweeps replace the
 * geometry-only ones at both ends; nothing is evaluated twice) and the kinetic
 * energies 0.5 p.p / mass.                                                          */
int tm_vv_integrate_async(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, double mass, double dt,
                          int n_steps, int with_energies);
/* waits; stats = { U(x0), K(p0), U(xN), K(pN) }, props0/props1 (optional) the reduced
 * properties at both ends. On a partitioned mesh K is the owned vertices' share.     */
int tm_vv_stats(tm_mesh* m, const tm_energy_params* p, double stats[4], double props0[6], double props1[6]);
/* the synchronous whole: upload p0, integrate, download x (optional), p (optional),
 * stats (optional; NULL skips the energy sweeps)                                     */
int tm_vv_integrate(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, const double* p0, double mass,
                    double dt, int n_steps, double* x_out, double* p_out, double stats[4]);

/* On-device rebuild trigger (the cadence the reference leaves to its per-move callback,
 * src/trimem/mc/evaluators.py:181-182). Every tm_vv_integrate over a verlet list also measures, on the device and
 * behind the trajectory (one more kernel, no extra synchronisation: the value travels with the statistics), how far
 * the END POINT of the trajectory is from the positions the list was built from. displacement = the largest such
 * vertex displacement; margin = rlist - rep_lc1: as long as no vertex has moved more than margin / 2 since the build,
 * the frozen rows cannot miss a pair inside the cut-off and a rebuild would change no energy or gradient.
 * Returns an error if the last trajectory did not measure it (cell list, no list, asynchronous entry).          */
int tm_vv_displacement(tm_mesh* m, tm_nlist* nl, double rep_lc1, double* displacement, double* margin);

/* ---- one mesh partitioned across devices  (BASELINE.json config 4) ----
 * No reference counterpart (its sweeps are OpenMP loops in one address space,
 * src/kernel.h:243-262). Every rank holds the whole mesh, owns a contiguous slice of
 * the (Morton-ordered) vertices, evaluates / integrates / reports only those, reads
 * halo positions straight from the owners' memory over NVLink and adds the ranks'
 * reduced sums in rank order (csrc/tm_part.cu). Every rank must make the same
 * sequence of evaluation calls. After tm_part_create:
 *   tm_energy / tm_properties      global sums, identical bits on every rank
 *   tm_gradient / tm_fetch_gradient, tm_mesh_get_momenta
 *                                  rows of the owned vertices, zeros elsewhere
 *   tm_mesh_get_positions, tm_nlist_create
 *                                  first bring every rank's whole array up to date
 *   flips, per-vertex queries      operate on the whole replicated mesh            */
int tm_part_create(tm_mesh* m, int rank, int nranks);
int tm_part_info(tm_mesh* m, int* rank, int* nranks, int* first_slot, int* n_owned, int* n_ghost_ring1, int* n_pull);
/* peers in other processes: 128 opaque bytes (two cudaIpcMemHandle_t) to pass around */
int tm_part_export(tm_mesh* m, unsigned char handle[128]);
int tm_part_attach_ipc(tm_mesh* m, int peer_rank, const unsigned char handle[128]);
/* peers in this process (other device, or the same one in tests)                    */
int tm_part_attach_local(tm_mesh* m, int peer_rank, tm_mesh* peer);
/* local: halo lists and every on-demand buffer, so that no allocation happens
 * between the collective calls that follow                                          */
int tm_part_prepare(tm_mesh* m, tm_nlist* nl);
/* collective: the whole position array of every rank becomes current                */
int tm_part_sync_positions(tm_mesh* m);
/* collective: out[q] = sum over ranks of in[q], added in rank order (the path the
 * property sums take; tests)                                                        */
int tm_part_allsum6(tm_mesh* m, const double in[6], double out[6]);
/* diagnostics: the rank's sequencing block {pos_seq, sums_seq, error, host-side pos_seq,
 * host-side sums_seq, ack[0..16)}                                                    */
int tm_part_debug(tm_mesh* m, unsigned long long out[21]);
/* syncs the mesh's stream and reports an exchange time-out, if any                  */
int tm_part_check(tm_mesh* m);

#ifdef __cplusplus
}
#endif
#endif /* TRIMEM_B200_H */
