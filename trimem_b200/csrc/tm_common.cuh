// Shared declarations of libtrimem_b200: handle structs, device views, helpers.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/trimem_b200.h"

// ------------------------------------------------------------------ errors --
void tm_set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_tm_launches;      // all threads (tm_launch_count)
extern thread_local unsigned long long tl_tm_launches;     // this thread (per-call deltas)

#define TM_CUDA(call)                                                           \
    do {                                                                        \
        cudaError_t err__ = (call);                                             \
        if (err__ != cudaSuccess) {                                             \
            tm_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,   \
                         cudaGetErrorString(err__));                            \
            return 1;                                                           \
        }                                                                       \
    } while (0)

#define TM_CHECK(cond, ...)                                                     \
    do {                                                                        \
        if (!(cond)) { tm_set_error(__VA_ARGS__); return 1; }                   \
    } while (0)

#define TM_LAUNCHED() (++tl_tm_launches, g_tm_launches.fetch_add(1, std::memory_order_relaxed))

// --------------------------------------------------------------- constants --
constexpr int TM_RING_FACE  = 1 << 30;          // slot's outgoing half-edge has a face
constexpr int TM_RING_OFACE = (int)0x80000000u;         // its opposite half-edge has a face
constexpr int TM_RING_MASK  = (1 << 30) - 1;
// fixed-width one-ring rows (R8) of the vertices with a CLOSED fan of valence 3..8 (every spoke has
// a face on both sides): plain neighbour ids, unused slots hold TM_RING_PAD. Every other vertex
// (valence > 8 or < 3, or on the boundary of an open mesh) has TM_RING_GENERIC in slot 0 and is
// evaluated from the CSR row instead
constexpr int TM_RING_PAD     = TM_RING_MASK;
constexpr int TM_RING_GENERIC = TM_RING_MASK - 1;
// launch shapes of the two sweeps (overridable with -D for timing experiments, scripts/build_variant.py)
#ifndef TM_PROPS_NT
#define TM_PROPS_NT 128
#endif
#ifndef TM_GRAD_NT
#define TM_GRAD_NT 128
#endif
// resident blocks per SM the register allocation aims at (measured on B200 at 655,362 vertices, profiles/r02*
// kernel timings): property sweep 8 blocks of 128 = 64 registers (44 bytes of spills) 42.2 us, 6 blocks / 80
// registers 44.6, 5 / 96 47.0, 4 / 120 50.9. Ring gradient, after the pair force and the sums moved to the head of
// the kernel: 4 blocks / 128 registers (no spills) 42.5 us, 5 / 96 45.8, 6 / 80 (470 bytes of spills) 44.1, 8 / 64
// 48.7: the sweep is bound by the round trips of a warp, not by issue, and the larger allocation lets the scheduler
// keep more of a warp's loads in flight. 64- and 256-thread blocks at the same number of resident warps: within 1 us.
#ifndef TM_PROPS_MINB
#define TM_PROPS_MINB 8
#endif
#ifndef TM_GRAD_MINB
#define TM_GRAD_MINB 4
#endif
constexpr int TM_PROPS_THREADS = TM_PROPS_NT;
constexpr int TM_GRAD_THREADS  = TM_GRAD_NT;
constexpr int TM_PAIR_STREAMS = 4;             // side streams a pair sweep can be cut across
constexpr int TM_REDUCE_BLOCKS = 128;           // second-stage blocks of the property reduction (upper bound)

// ------------------------------------------------------------ device views --
// One 32-byte record per vertex = one DRAM/L2 sector: position plus rho = c/a (curvature /
// area ratio of the last property sweep), so the gather sweeps fetch everything they need
// about a ring neighbour with a single 256-bit load.
struct __align__(32) P4 { double x, y, z, w; };
// One 32-byte record per vertex as well: the clockwise one-ring (neighbour id | face flags) of a
// vertex of valence <= 8. A thread fetches its whole ring with one coalesced 256-bit load instead
// of the dependent ring_off -> ring_v chain, and can then issue all position gathers at once.
struct __align__(32) R8 { int v[8]; };

// All device arrays use the DEVICE vertex numbering: vertices sorted by the Morton code of a
// coarse grid cell (tm_layout.cu), so that the 32 vertices of a warp form a compact surface
// patch and their ring / neighbour-list gathers hit the same few cache lines. d2h / h2d map
// to and from the host (OpenMesh) numbering at the C-ABI boundary; half-edge, edge and face
// ids are not renumbered.
struct MeshView {
    int V, HE;
    const int* he_to;
    const int* he_next;
    const int* he_face;
    const int* v_he;
    const int* ring_off;   // V+1
    const int* ring_v;     // HE entries, flags in the two top bits
    const R8* ring8;       // V fixed-width rows (see TM_RING_PAD / TM_RING_GENERIC)
    const P4* pos;         // V packed (x, y, z, rho) records, device vertex numbering
    const int* ex2_off;    // sorted two-ring rows (exclusion level 2), valid after tm_ensure_ring2
    const int* ex2_v;
    // per ring slot, written by the property sweep and read by the gradient sweep (slot-major, [8][aux_vs]):
    // aux[s][i] = tau = theta_s / l_s of spoke s of vertex i
    double* aux;
    int aux_vs;            // V rounded up to a multiple of 32
};

// Vertices one sweep launch covers: device slots [first, first + count), or the explicit list
// vlist[0..count) (ghost sweeps of a partitioned mesh, tm_part.cu). `first` is a multiple of 32,
// so a warp's partial of a partitioned sweep is the same sum as in the single-device sweep.
struct SweepRange {
    int first, count;
    const int* vlist;
};

struct NlView {
    int kind;              // -1: none, TM_NLIST_CELL, TM_NLIST_VERLET
    int excl;
    // cell grid (build-time occupancy)
    double box_min[3];
    double cell[3];
    int shape[3];
    int strides[3];
    const int* cell_start; // ncell+1
    const int* cell_items; // V, ascending vertex id inside a cell
    const P4* cell_rec;    // V, (x, y, z, id bits) of cell_items[k] at the positions of the current sweep: the
                           //    candidates of a cell-list scan as contiguous records (refreshed per sweep)
    // verlet rows, fixed width, sliced by warp: entry k of vertex i is
    // row_j[((i >> 5) * row_cap + k) * 32 + (i & 31)], k < row_cnt[i]. The 32 threads of a warp
    // read 32 consecutive ints per step and a warp's whole slice is one contiguous block
    // (row_cap * 128 bytes), so the sweep stays inside a page or two.
    const int* row_cnt;    // V
    const int* row_j;      // ceil(V/32) * row_cap * 32
    int row_cap;
    unsigned long long serial;   // identity of the list (pair-force cache)
};

// ----------------------------------------------------------------- handles --
struct tm_mesh {
    int device = 0;
    int sm_count = 148;          // multiprocessors of the device (persistent grids)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // the pair sweep depends on positions only, like the ring property sweep: it runs beside it on
    // a second stream (fork after the positions are final, join before its first consumer)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool pairs_inflight = false;
    // the pair sweep cut into ranges of tiles, one side stream each; the ring gradient of a range follows that range's
    // pair sweep on its stream (launch_pairs / launch_grad_kernels in tm_energy.cu)
    cudaStream_t sp_stream[TM_PAIR_STREAMS] = {};
    cudaEvent_t sp_ev[TM_PAIR_STREAMS] = {};    // end of range k's last kernel
    cudaEvent_t ev_sums = nullptr;              // the reduced sums of the property sweep are in place (main stream)
    int split = 0;                              // ranges of the pair sweep in flight (0: one sweep on stream2)
    int rep_cols = 0;                           // > 0: columns of warp partials whose repulsion column awaits its reduction
    unsigned long long vv_disp_list = 0;   // serial of the list the last trajectory measured its end point's displacement against
    cudaEvent_t ev_pre_upload = nullptr;   // recorded on `stream` just before a position upload is enqueued
    int V = 0, HE = 0;
    int *he_to = nullptr, *he_next = nullptr, *he_face = nullptr, *v_he = nullptr;
    int *ring_off = nullptr, *ring_v = nullptr, *ring_cnt = nullptr;
    R8* ring8 = nullptr;
    double* aux = nullptr;       // 8 * aux_vs slot quantities (tau) handed from the property to the gradient sweep
    int* gen_list = nullptr;     // vertices without a fixed-width row (TM_RING_GENERIC), ascending (tm_rings.cu)
    int n_generic = 0;
    bool rings_valid = false;
    bool has_boundary = false;   // the uploaded connectivity has boundary half-edges (open mesh)
    // host-numbered connectivity as uploaded (he_to_h, v_he_h) and the vertex permutation
    int *he_to_h = nullptr, *v_he_h = nullptr;
    int *d2h = nullptr, *h2d = nullptr;
    bool perm_valid = false;       // d2h/h2d computed for the current vertex set
    bool topo_dev_valid = false;   // he_to / v_he (device numbering) reflect he_to_h / v_he_h
    bool raw_stale = false;        // device connectivity mutated by flips; he_to_h / v_he_h are old
    bool stage_pending = false;    // xio3 holds uploaded positions not yet imported into pos
    bool have_positions = false;
    unsigned long long layout_epoch = 0;   // bumped whenever the permutation changes
    unsigned long long pos_epoch = 1;      // bumped whenever positions or connectivity change
    unsigned long long geo_epoch = 0;      // pos_epoch at which sums[0..2] and rho were evaluated
    bool sums_full = false;                // the last property sweep also reduced tether and repulsion (sums[4..5])
    P4* pos = nullptr;           // V   (x, y, z, rho), device numbering
    P4* pos_alt = nullptr;       // V   second record array: a trajectory step writes x_{n+1} here while the gradient
                                 //     sweep still gathers x_n from `pos`, then the two swap (tm_integrate.cu)
    double* xio3 = nullptr;      // V*3 staging in HOST numbering: uploaded positions
    double* xout3 = nullptr;     // V*3 staging in HOST numbering: results on their way out
    double* vprops = nullptr;    // V*6 (allocated on demand)
    double* grad = nullptr;      // V*3 device numbering
    // velocity-Verlet trajectory state (tm_integrate.cu; allocated on demand)
    double* grad_prev = nullptr; // V*3 gradient at the previous trajectory point (ping-pong with grad)
    double* mom = nullptr;       // V*3 momenta, device numbering
    unsigned long long grad_epoch = 0;     // pos_epoch the content of `grad` was evaluated at
    double* h_stats = nullptr;   // pinned, 16: energies / reduced properties a trajectory reports
    double* dot_part = nullptr;  // block partials of the kinetic-energy reduction
    double* d_stats = nullptr;   // 16: what h_stats receives, collected on the device during a trajectory
    // vertex-partitioned evaluation across devices (tm_part.cu); nullptr = whole mesh on this device
    struct tm_part* part = nullptr;
    struct tm_vvcache* vvcache = nullptr;   // captured trajectories (tm_integrate.cu)
    double* frep = nullptr;      // V*3 pair force of the last pair sweep (allocated on demand)
    unsigned long long cellrec_epoch = 0, cellrec_list = 0;   // positions / cell list the list's cell_rec holds
    unsigned long long frep_epoch = 0;     // pos_epoch, potential and list it was evaluated for
    double frep_lc1 = 0; int frep_r = 0; unsigned long long frep_list = 0;
    // sorted, de-duplicated two-ring rows (verlet exclusion level 2), rebuilt when the topology changed
    int *ex2_off = nullptr, *ex2_v = nullptr;
    unsigned long long topo_epoch = 1, ex2_epoch = 0;
    int verlet_cap_hint = 0;     // row width that fitted the last verlet build of this mesh
    double* partials = nullptr;  // 6 * (warps of the property sweep) + 6 * TM_REDUCE_BLOCKS
    size_t partials_n = 0;
    double* sums = nullptr;      // 8
    unsigned* ticket = nullptr;
    double* h_sums = nullptr;    // pinned, 8
    int* scan_tmp = nullptr;     // block sums for scans
    size_t scan_tmp_n = 0;
    void* flush = nullptr;
    size_t flush_bytes = 0;
    // flip workspace (allocated on demand)
    void* flipws = nullptr;
    MeshView view() const {
        MeshView v;
        v.V = V; v.HE = HE;
        v.he_to = he_to; v.he_next = he_next; v.he_face = he_face; v.v_he = v_he;
        v.ring_off = ring_off; v.ring_v = ring_v; v.ring8 = ring8; v.pos = pos;
        v.ex2_off = ex2_off; v.ex2_v = ex2_v;
        v.aux = aux; v.aux_vs = (V + 31) & ~31;
        return v;
    }
};

struct tm_nlist {
    int device = 0;
    int kind = -1, excl = 0;
    double rlist = 0, box_eps = 0;
    int V = 0;
    double box_min[3] = {0, 0, 0}, box_max[3] = {0, 0, 0}, cell[3] = {0, 0, 0};
    int shape[3] = {1, 1, 1}, strides[3] = {1, 1, 1};
    long long ncell = 0;
    int *cell_start = nullptr, *cell_items = nullptr;
    P4* cell_rec = nullptr;
    int *row_cnt = nullptr, *row_j = nullptr;
    int row_cap = 0;
    int rows_first = 0, rows_count = 0;   // vertices whose verlet rows were built (all, or a partition's slice)
    unsigned long long serial = 0;
    int* cell_of = nullptr;      // build-time cell of every vertex (cell lists)
    double* x_build = nullptr;   // V*3 positions the verlet rows were built from (device numbering): skin criterion
    long long n_entries = 0;
    unsigned long long layout_epoch = 0;   // vertex numbering the list was built in
    NlView view() const {
        NlView v;
        v.kind = kind; v.excl = excl;
        for (int k = 0; k < 3; k++) {
            v.box_min[k] = box_min[k]; v.cell[k] = cell[k];
            v.shape[k] = shape[k]; v.strides[k] = strides[k];
        }
        v.cell_start = cell_start; v.cell_items = cell_items; v.cell_rec = cell_rec;
        v.row_cnt = row_cnt; v.row_j = row_j; v.row_cap = row_cap; v.serial = serial;
        return v;
    }
};

// first entry of vertex i's row; consecutive entries are 32 ints apart
#ifdef __CUDACC__
__host__ __device__
#endif
inline size_t tm_row_base(int i, int row_cap) { return ((size_t)(i >> 5) * row_cap) * 32 + (i & 31); }

inline int tm_sm_count(const tm_mesh* m) { return m->sm_count; }

inline NlView tm_no_nlist() {
    NlView v;
    memset(&v, 0, sizeof(v));
    v.kind = -1;
    return v;
}

// ------------------------------------------------- internal host interfaces --
int tm_ensure_layout(tm_mesh* m);                                // tm_layout.cu: permutation, device-numbered
                                                                 // connectivity, staged positions, rings
int tm_refresh_raw_topology(tm_mesh* m);                         // tm_layout.cu
int tm_export_positions(tm_mesh* m);                             // tm_layout.cu: pos -> xio3 (host numbering)
int tm_export_gradient(tm_mesh* m);                              // tm_layout.cu: grad -> xio3 (host numbering)
int tm_bbox(tm_mesh* m, const double* x, int stride, int V, double out[6]);   // tm_layout.cu: min xyz, max xyz
int tm_sort_segments(tm_mesh* m, long long nseg, const int* start, int* items);   // tm_layout.cu
int tm_check_nlist(const tm_mesh* m, const tm_nlist* nl);        // tm_layout.cu
int tm_ensure_rings(tm_mesh* m);                                 // tm_rings.cu
int tm_ensure_ring2(tm_mesh* m);                                 // tm_rings.cu
// size-keyed cache of device blocks (neighbour lists are rebuilt every few steps: no cudaMalloc /
// cudaFree, and hence no device-wide sync, in steady state)                      tm_api.cu
void* tm_pool_alloc(int device, size_t bytes);
void tm_pool_free(int device, void* p);
int tm_exclusive_scan(tm_mesh* m, const int* in, int* out, int n, long long* total);  // tm_rings.cu
int tm_props_workspace(tm_mesh* m);                              // tm_energy.cu: partials + slot arrays
int tm_launch_props(tm_mesh* m, const NlView& nl, const tm_energy_params& p,
                    bool full, bool write_vprops);               // tm_energy.cu
// Velocity-Verlet update fused into the gradient sweep's epilogue (tm_integrate.cu): mode 1 = momentum
// update of the step that ends at the swept positions, mode 2 = that plus the position update of the next
// step, written to pos_next (the sweep itself still reads the old records).
struct VVFuse {
    int mode;                    // 0 none
    double* mom;
    const double* g_prev;
    P4* pos_next;
    double dt, dt2, mass;
};
int tm_launch_grad(tm_mesh* m, const NlView& nl, const tm_energy_params& p, const VVFuse* fuse = nullptr);  // tm_energy.cu
int tm_launch_pairs(tm_mesh* m, const NlView& nl, const tm_energy_params& p); // tm_energy.cu (force only)
// gradient sweep from the CURRENT device positions and caches on `st`, exported (host numbering) into
// m->xout3 and copied to `grad_host`; no layout / staging work (speculative path of tm_gradient)
int tm_launch_grad_speculative(tm_mesh* m, const NlView& nl, const tm_energy_params& p, cudaStream_t st, double* grad_host);
bool tm_pair_force_valid(const tm_mesh* m, const NlView& nl, const tm_energy_params& p);
int tm_nlist_measure_displacement(tm_nlist* nl, tm_mesh* m, double* out);   // tm_nlist.cu (asynchronous)
int tm_ensure_cellrec(tm_mesh* m, const NlView& nl);   // tm_energy.cu: cell_rec of a cell list at the current positions
int tm_random_momenta(tm_mesh* m, unsigned long long seed, unsigned long long seq, double scale);   // tm_integrate.cu
bool tm_pair_split_active(const tm_mesh* m, const NlView& nl);   // tm_energy.cu: pair sweeps of this mesh / list are cut across the side streams
int tm_join_pairs(tm_mesh* m);   // tm_energy.cu: the mesh's stream waits for a pair sweep still running beside it
int tm_flips_release(tm_mesh* m);                               // tm_flips.cu
double tm_energy_from_props(const tm_energy_params& p, const double s[6]);    // tm_api.cu (kernel.h:110-122)
int tm_vv_run(tm_mesh* m, const NlView& nl, const tm_energy_params& p, double mass, double dt, int n_steps,
              bool with_energies);                               // tm_integrate.cu
void tm_vv_release(tm_mesh* m);                                  // tm_integrate.cu
int tm_import_momenta(tm_mesh* m, const double* p_host);         // tm_integrate.cu
int tm_export_vec3_owned(tm_mesh* m, const double* dev, double* host);   // tm_integrate.cu (async copy)
// vertices this device sweeps (whole mesh, or the owned slice of a partitioned mesh)   tm_part.cu
SweepRange tm_owned_range(const tm_mesh* m);
// partitioned mesh: ring-1 ghost property sweep list (rho of non-owned ring neighbours), exchange
// of the locally reduced sums; no-ops returning 0 on an unpartitioned mesh                tm_part.cu
SweepRange tm_ghost_range(const tm_mesh* m);
double* tm_part_local_sums(tm_mesh* m);          // where k_reduce_props writes (m->sums if unpartitioned)
int tm_part_exchange_sums(tm_mesh* m, int qlo, int qhi);   // publish local sums, add all ranks' sums [qlo, qhi) in rank order -> m->sums
int tm_part_nranks(const tm_mesh* m);            // 1 if unpartitioned
int tm_part_positions_updated(tm_mesh* m);       // after a device-side update of the owned positions: halo pull
int tm_part_before_update(tm_mesh* m);           // peers must have pulled the previous positions
int tm_part_release(tm_mesh* m);
// CUDA loads kernels lazily, and the first launch of a kernel may synchronise the context. A rank
// whose stream waits (on the device) for a peer living in the SAME context would then hold up the
// very launch it waits for, so a partitioned mesh loads every kernel of the path up front.
int tm_preload_energy();      // tm_energy.cu
int tm_preload_integrate();   // tm_integrate.cu
int tm_preload_layout();      // tm_layout.cu
#define TM_PRELOAD(k) do { cudaFuncAttributes a__; TM_CUDA(cudaFuncGetAttributes(&a__, k)); } while (0)
int tm_part_refresh(tm_mesh* m, const NlView& nl);   // halo lists for the current connectivity / list (local)
int tm_part_make_current(tm_mesh* m);            // full position exchange if the far field is stale
int tm_part_quiesce(tm_mesh* m);                 // partitioned: drain the stream before enqueuing a copy
int tm_part_poll_error(tm_mesh* m);
void tm_part_positions_uploaded(tm_mesh* m);              // syncs; reports an exchange time-out

// ------------------------------------------------------------ device maths --
#ifdef __CUDACC__
struct d3 { double x, y, z; };
__device__ __forceinline__ d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator*(double s, d3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ double dot3(d3 a, d3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ d3 cross3(d3 a, d3 b) {
    return mk3(fma(a.y, b.z, -(a.z * b.y)), fma(a.z, b.x, -(a.x * b.z)), fma(a.x, b.y, -(a.y * b.x)));
}
__device__ __forceinline__ void axpy3(d3& acc, double s, d3 a) {
    acc.x = fma(s, a.x, acc.x); acc.y = fma(s, a.y, acc.y); acc.z = fma(s, a.z, acc.z);
}
// position of vertex i. Coherent loads of the first 24 bytes only: the property sweep writes
// rho (.w) of other vertices while it runs, so .w must not be touched (nor the record read
// through the non-coherent path) there.
__device__ __forceinline__ d3 ldpos(const P4* pos, int i) {
    const double2 xy = *reinterpret_cast<const double2*>(pos + i);
    const double z = pos[i].z;
    return mk3(xy.x, xy.y, z);
}
// whole record through the read-only path (one LDG.E.256.CONSTANT): only in kernels that do
// not write pos
__device__ __forceinline__ P4 ldrec(const P4* __restrict__ pos, int i) {
    P4 r;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(pos + i));
    return r;
}
// 1/x and 1/sqrt(x) for normal, positive x, branch-free (the compiler's own sequences carry a slow-path call
// for denormals/inf that these arguments never take). The MUFU seeds are good to 2^-20 on B200
// (scripts/ubench/seed_accuracy.cu: 9.4e-7), so ONE third-order step reaches the rounding level
// (measured max relative error 2.2e-16 for rcp, 2.7e-16 for rsqrt, the same as two Newton steps) with
// 3 / 5 dependent fp64 operations instead of 4 / 7.
#ifndef TM_CUBIC
#define TM_CUBIC 1
#endif
__device__ __forceinline__ double rcp_pos(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#if TM_CUBIC
    const double e = fma(-x, r, 1.0);
    return fma(r, fma(e, e, e), r);          // r (1 + e + e^2)
#else
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#endif
}
__device__ __forceinline__ double rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if TM_CUBIC
    const double e = fma(-x * y, y, 1.0);    // 1 - x y^2
    return fma(y * e, fma(0.375, e, 0.5), y);   // y (1 + e/2 + 3 e^2 / 8)
#else
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);     // 0.5 (1 - x y^2)
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    return fma(y, e, y);
#endif
}
// x^r for the small integer exponents of the potentials (std::pow(l, r_) in the reference)
__device__ __forceinline__ double ipow(double x, int r) {
    if (r == 2) return x * x;
    if (r < 0) return 1.0 / ipow(x, -r);
    double acc = 1.0, b = x;
    while (r) { if (r & 1) acc *= b; b *= b; r >>= 1; }
    return acc;
}
// Euclidean distance exactly as the reference evaluates it on x86-64 without FMA
// (cell_list.h:11-26, VectorT::norm): ((dx*dx + dy*dy) + dz*dz), correctly rounded sqrt.
__device__ __forceinline__ double dist_nofma(d3 d) {
    double s = __dmul_rn(d.x, d.x);
    s = __dadd_rn(s, __dmul_rn(d.y, d.y));
    s = __dadd_rn(s, __dmul_rn(d.z, d.z));
    return __dsqrt_rn(s);
}
__device__ __forceinline__ d3 sub_nofma(d3 a, d3 b) {
    return mk3(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z));
}
#endif
