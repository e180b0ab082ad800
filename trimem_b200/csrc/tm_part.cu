// Vertex-partitioned evaluation of ONE mesh across devices (BASELINE.json config 4).
//
// The reference has no distributed path: its sweeps are OpenMP loops over all vertices of one
// mesh in one address space (kernel.h:243-262). The partition keeps that contract -- same
// sums, same gradient rows -- and splits the vertex loop.
//
// Scheme. Every rank (one device, normally one process) holds the whole connectivity and a
// full-size position array; the device vertex order is the Morton order of tm_layout.cu, the
// same on every rank because it is built deterministically from the same first positions. Rank
// r owns the contiguous slot range [bounds[r], bounds[r+1]) (multiples of 32: a warp never
// straddles a cut, so a warp's partial sums are the sums of the single-device sweep). Owned
// vertices are the ones a rank evaluates, integrates and reports. What it needs from others:
//
//   * positions of every non-owned vertex within two rings of an owned one (ring positions for
//     the owned sweeps; the rings of ring-1 ghosts to recompute their rho = c/a locally, which
//     replaces a second exchange between the two gradient sweeps) and of every non-owned entry
//     of an owned verlet row: the PULL LIST, built on the device from the one-ring table and the
//     list rows (cell lists search live cells, so they pull everything);
//   * the other ranks' six locally reduced sums, every property sweep.
//
// Exchange = loads over NVLink from the peers' own arrays (peer-mapped: cudaIpc handles between
// processes, direct pointers inside one process), sequenced by counters in a small peer-mapped
// block per rank; no host round trip, no NCCL on the data path:
//
//   positions   after its update kernel a rank publishes pos_seq = s (fence + release store);
//               a one-warp kernel waits until every peer shows pos_seq >= s, the pull kernel
//               copies the listed records from the owners' arrays, a last kernel stores
//               ack[me] = s into every peer's block. A rank overwrites its positions only
//               after all peers acknowledged the previous sequence number (WAR).
//   sums        k_reduce_props writes the local sums into slot (q & 1) of the rank's block;
//               one kernel publishes sums_seq = q, waits for all peers, and adds the slots in
//               RANK ORDER into m->sums: every rank obtains bit-identical global sums, run to
//               run. Two slots suffice: a rank reaches sweep q+2 only after it has read every
//               peer's q+1, which each peer published after reading q.
//
// Every wait carries a timeout (TM_PART_TIMEOUT_MS, default 20 s) that raises an error flag the
// host reads at the next synchronising call, instead of hanging the device.
#include <algorithm>
#include <cstdint>

#include "tm_common.cuh"

constexpr int TM_MAX_RANKS = 16;

struct PartComm {
    unsigned long long pos_seq;              // position sets this rank has published
    unsigned long long sums_seq;             // local-sum sets this rank has published
    unsigned long long ack[TM_MAX_RANKS];    // ack[r]: last pos_seq of THIS rank that rank r has pulled
    double sums[2][8];
    int error;
    int pad;
};

struct PeerTable {
    const P4* pos[TM_MAX_RANKS];
    PartComm* comm[TM_MAX_RANKS];
    int bounds[TM_MAX_RANKS + 1];
    int rank, nranks;
};

struct tm_part {
    int rank = 0, nranks = 1;
    PeerTable peers;
    bool attached[TM_MAX_RANKS] = {false};
    void* ipc_pos[TM_MAX_RANKS] = {nullptr};      // mappings to close
    void* ipc_comm[TM_MAX_RANKS] = {nullptr};
    PartComm* comm = nullptr;
    int* g1 = nullptr; int n_g1 = 0;              // non-owned ring neighbours of owned vertices
    int* pull = nullptr; int n_pull = 0;          // non-owned vertices whose positions the owned sweeps read
    bool pull_all = false;                        // cell lists: everything non-owned
    int* mark = nullptr;                          // 3V + 4 workspace (two mark arrays, scan output)
    unsigned long long lists_topo = 0, lists_nl = 0; bool lists_valid = false;
    unsigned long long pos_seq = 0, sums_seq = 0; // host mirrors of the published counters
    bool stale_far = false;                       // non-halo positions are older than the owners'
    unsigned long long timeout_ns = 20000000000ull;
    int* h_err = nullptr;                         // pinned
};

namespace {

constexpr int NT = 256;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ bool spin_until(const unsigned long long* p, unsigned long long target, unsigned long long timeout_ns) {
    const unsigned long long t0 = now_ns();
    while (ld_acquire_sys(p) < target) {
        if (now_ns() - t0 > timeout_ns) return false;
        __nanosleep(64);
    }
    return true;
}
__device__ __forceinline__ int owner_of(const PeerTable& T, int v) {
    int r = 0;
    while (r + 1 < T.nranks && v >= T.bounds[r + 1]) ++r;
    return r;
}

// ---- sequencing kernels (one warp each) ---------------------------------------------------
__global__ void k_publish_pos(PartComm* own, unsigned long long seq) {
    if (threadIdx.x == 0) { __threadfence_system(); st_release_sys(&own->pos_seq, seq); }
}
__global__ void k_wait_pos(PeerTable T, PartComm* own, unsigned long long seq, unsigned long long timeout_ns) {
    const int r = threadIdx.x;
    if (r < T.nranks && r != T.rank && !spin_until(&T.comm[r]->pos_seq, seq, timeout_ns)) own->error = 1;
}
__global__ void k_ack(PeerTable T, unsigned long long seq) {
    const int r = threadIdx.x;
    if (r < T.nranks && r != T.rank) { __threadfence_system(); st_release_sys(&T.comm[r]->ack[T.rank], seq); }
}
__global__ void k_wait_acks(PeerTable T, PartComm* own, unsigned long long seq, unsigned long long timeout_ns) {
    const int r = threadIdx.x;
    if (r < T.nranks && r != T.rank && !spin_until(&own->ack[r], seq, timeout_ns)) own->error = 2;
}

// positions of the listed (or of all non-owned) vertices from their owners' arrays
__global__ void __launch_bounds__(NT)
k_pull(PeerTable T, const int* __restrict__ list, int n, P4* pos) {
    const int t = blockIdx.x * NT + threadIdx.x;
    if (t >= n) return;
    const int lo = T.bounds[T.rank], hi = T.bounds[T.rank + 1];
    const int v = list ? list[t] : (t < lo ? t : t + (hi - lo));
    const P4* src = T.pos[owner_of(T, v)] + v;
    const double2 xy = __ldcv(reinterpret_cast<const double2*>(src));
    const double z = __ldcv(&src->z);
    *reinterpret_cast<double2*>(pos + v) = xy;     // rho (.w) of a ghost is this rank's own business
    pos[v].z = z;
}

// publish the local sums of sweep `seq`, wait for every peer's, add in rank order
__global__ void k_exchange_sums(PeerTable T, PartComm* own, unsigned long long seq, int qlo, int qhi, double* __restrict__ out,
                                unsigned long long timeout_ns) {
    if (threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(&own->sums_seq, seq);
    }
    __syncwarp();
    const int r = threadIdx.x;
    if (r < T.nranks && r != T.rank && !spin_until(&T.comm[r]->sums_seq, seq, timeout_ns)) own->error = 3;
    __syncwarp();
    __threadfence_system();
    const int q = threadIdx.x;
    if (q >= qlo && q < qhi) {
        double s = 0.0;
        for (int k = 0; k < T.nranks; k++) s += __ldcv(&T.comm[k]->sums[seq & 1][q]);
        out[q] = s;
    }
}

// ---- halo lists ----------------------------------------------------------------------------
__global__ void k_mark_rings(int lo, int hi, const int* __restrict__ sel, int V, const int* __restrict__ ring_off,
                             const int* __restrict__ ring_v, int* __restrict__ mark) {
    // sel == nullptr: rings of the owned range; else rings of the vertices with sel[v] != 0
    const int v = blockIdx.x * NT + threadIdx.x;
    if (v >= V) return;
    const bool own = v >= lo && v < hi;
    if (sel ? (sel[v] == 0) : !own) return;
    for (int k = ring_off[v]; k < ring_off[v + 1]; k++) {
        const int j = ring_v[k] & TM_RING_MASK;
        if (j < lo || j >= hi) mark[j] = 1;
    }
}
__global__ void k_mark_rows(int lo, int hi, NlView nl, int* __restrict__ mark) {
    const int i = lo + blockIdx.x * NT + threadIdx.x;
    if (i >= hi) return;
    const int e = nl.row_cnt[i];
    const int* col = nl.row_j + tm_row_base(i, nl.row_cap);
    for (int k = 0; k < e; k++) {
        const int j = col[k * 32];
        if (j < lo || j >= hi) mark[j] = 1;
    }
}
__global__ void k_compact(int V, const int* __restrict__ mark, const int* __restrict__ off, int* __restrict__ out) {
    const int v = blockIdx.x * NT + threadIdx.x;
    if (v < V && mark[v]) out[off[v]] = v;
}

int compact(tm_mesh* m, const int* mark, int* scratch, int** list, int* n) {
    long long total = 0;
    if (tm_exclusive_scan(m, mark, scratch, m->V, &total)) return 1;
    if (*list) { tm_pool_free(m->device, *list); *list = nullptr; }
    *n = (int)total;
    *list = reinterpret_cast<int*>(tm_pool_alloc(m->device, sizeof(int) * (size_t)(total > 0 ? total : 1)));
    TM_CHECK(*list != nullptr, "out of device memory");
    k_compact<<<(m->V + NT - 1) / NT, NT, 0, m->stream>>>(m->V, mark, scratch, *list); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    return 0;
}

struct Six { double v[6]; };
__global__ void k_store6(double* dst, Six s) {
    if (threadIdx.x < 6) dst[threadIdx.x] = s.v[threadIdx.x];
}

// Copies are enqueued only on a drained stream: a copy waiting behind a kernel that waits for a
// peer would occupy the copy engine's in-order queue, and with several ranks on ONE device (tests)
// the peer's own copy could be stuck behind it.
int check_error(tm_mesh* m) {
    tm_part* P = m->part;
    TM_CUDA(cudaStreamSynchronize(m->stream));
    TM_CUDA(cudaMemcpyAsync(P->h_err, &P->comm->error, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    const int e = *P->h_err;
    TM_CHECK(e == 0, "partition exchange timed out on rank %d waiting for %s of a peer (is every rank making the same calls?)",
             P->rank, e == 1 ? "positions" : e == 2 ? "pull acknowledgements" : "reduced sums");
    return 0;
}

bool all_attached(const tm_part* P) {
    for (int r = 0; r < P->nranks; r++) if (!P->attached[r]) return false;
    return true;
}

}  // namespace

// --------------------------------------------------------------- internal interface --
SweepRange tm_owned_range(const tm_mesh* m) {
    SweepRange R;
    R.vlist = nullptr;
    if (m->part) { R.first = m->part->peers.bounds[m->part->rank]; R.count = m->part->peers.bounds[m->part->rank + 1] - R.first; }
    else { R.first = 0; R.count = m->V; }
    return R;
}

SweepRange tm_ghost_range(const tm_mesh* m) {
    SweepRange R;
    R.first = 0; R.count = 0; R.vlist = nullptr;
    if (m->part && m->part->lists_valid) { R.count = m->part->n_g1; R.vlist = m->part->g1; }
    return R;
}

double* tm_part_local_sums(tm_mesh* m) {
    if (!m->part) return m->sums;
    return m->part->comm->sums[(m->part->sums_seq + 1) & 1];
}

int tm_part_nranks(const tm_mesh* m) { return m->part ? m->part->nranks : 1; }

int tm_part_exchange_sums(tm_mesh* m, int qlo, int qhi) {
    tm_part* P = m->part;
    if (!P) return 0;
    TM_CHECK(all_attached(P), "partitioned mesh: not every peer is attached (tm_part_attach_*)");
    P->sums_seq++;
    k_exchange_sums<<<1, 32, 0, m->stream>>>(P->peers, P->comm, P->sums_seq, qlo, qhi, m->sums, P->timeout_ns); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    return 0;
}

int tm_part_before_update(tm_mesh* m) {
    tm_part* P = m->part;
    if (!P) return 0;
    k_wait_acks<<<1, 32, 0, m->stream>>>(P->peers, P->comm, P->pos_seq, P->timeout_ns); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    return 0;
}

static int exchange_positions(tm_mesh* m, bool everything) {
    tm_part* P = m->part;
    TM_CHECK(all_attached(P), "partitioned mesh: not every peer is attached (tm_part_attach_*)");
    TM_CHECK(everything || P->lists_valid, "internal: halo lists not built");
    P->pos_seq++;
    k_publish_pos<<<1, 32, 0, m->stream>>>(P->comm, P->pos_seq); TM_LAUNCHED();
    k_wait_pos<<<1, 32, 0, m->stream>>>(P->peers, P->comm, P->pos_seq, P->timeout_ns); TM_LAUNCHED();
    const SweepRange R = tm_owned_range(m);
    const bool all = everything || P->pull_all;
    const int n = all ? m->V - R.count : P->n_pull;
    if (n > 0) { k_pull<<<(n + NT - 1) / NT, NT, 0, m->stream>>>(P->peers, all ? nullptr : P->pull, n, m->pos); TM_LAUNCHED(); }
    k_ack<<<1, 32, 0, m->stream>>>(P->peers, P->pos_seq); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    P->stale_far = !all;
    return 0;
}

int tm_part_positions_updated(tm_mesh* m) {
    if (!m->part) return 0;
    return exchange_positions(m, false);
}

// whole position array current again (no-op unless a device-side update left the far field stale);
// collective when it does something: every rank is in the same state by construction
int tm_part_make_current(tm_mesh* m) {
    if (!m->part || !m->part->stale_far) return 0;
    if (exchange_positions(m, true)) return 1;
    m->pos_epoch++;
    return 0;
}

int tm_part_poll_error(tm_mesh* m) { return m->part ? check_error(m) : 0; }
// the caller uploaded a whole position array (the same one on every rank): nothing is stale
void tm_part_positions_uploaded(tm_mesh* m) { if (m->part) m->part->stale_far = false; }
int tm_part_quiesce(tm_mesh* m) {
    if (m->part) TM_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

// ghost lists for the current connectivity and neighbour list (local work, no peer involved)
int tm_part_refresh(tm_mesh* m, const NlView& nl) {
    tm_part* P = m->part;
    if (!P) return 0;
    const unsigned long long nls = nl.kind >= 0 ? nl.serial : 0;
    if (P->lists_valid && P->lists_topo == m->topo_epoch && P->lists_nl == nls) return 0;
    if (tm_ensure_rings(m)) return 1;
    const int V = m->V, vb = (V + NT - 1) / NT;
    const SweepRange R = tm_owned_range(m);
    const int lo = R.first, hi = R.first + R.count;
    // two mark arrays and the scan's output, which has V + 1 entries (offsets and the total)
    if (!P->mark) TM_CUDA(cudaMalloc(&P->mark, sizeof(int) * (3 * (size_t)V + 4)));
    int *m1 = P->mark, *m2 = P->mark + V, *scratch = P->mark + 2 * (size_t)V;
    TM_CUDA(cudaMemsetAsync(m1, 0, sizeof(int) * 2 * (size_t)V, m->stream));
    k_mark_rings<<<vb, NT, 0, m->stream>>>(lo, hi, nullptr, V, m->ring_off, m->ring_v, m1); TM_LAUNCHED();
    if (compact(m, m1, scratch, &P->g1, &P->n_g1)) return 1;
    // second ring: rings of the ring-1 ghosts; plus the ring-1 ghosts themselves
    TM_CUDA(cudaMemcpyAsync(m2, m1, sizeof(int) * (size_t)V, cudaMemcpyDeviceToDevice, m->stream));
    k_mark_rings<<<vb, NT, 0, m->stream>>>(lo, hi, m1, V, m->ring_off, m->ring_v, m2); TM_LAUNCHED();
    P->pull_all = nl.kind == TM_NLIST_CELL;
    if (nl.kind == TM_NLIST_VERLET && R.count > 0) {
        k_mark_rows<<<(R.count + NT - 1) / NT, NT, 0, m->stream>>>(lo, hi, nl, m2); TM_LAUNCHED();
    }
    if (compact(m, m2, scratch, &P->pull, &P->n_pull)) return 1;
    P->lists_topo = m->topo_epoch; P->lists_nl = nls; P->lists_valid = true;
    return 0;
}

int tm_part_release(tm_mesh* m) {
    tm_part* P = m->part;
    if (!P) return 0;
    for (int r = 0; r < TM_MAX_RANKS; r++) {
        if (P->ipc_pos[r]) cudaIpcCloseMemHandle(P->ipc_pos[r]);
        if (P->ipc_comm[r]) cudaIpcCloseMemHandle(P->ipc_comm[r]);
    }
    if (P->g1) tm_pool_free(m->device, P->g1);
    if (P->pull) tm_pool_free(m->device, P->pull);
    cudaFree(P->mark);
    cudaFree(P->comm);
    if (P->h_err) cudaFreeHost(P->h_err);
    delete P;
    m->part = nullptr;
    return 0;
}

// ------------------------------------------------------------------------- C ABI --
extern "C" {

int tm_part_create(tm_mesh* m, int rank, int nranks) {
    TM_CHECK(m != nullptr && m->V > 0, "mesh has no topology (call tm_mesh_set_topology first)");
    TM_CHECK(nranks >= 1 && nranks <= TM_MAX_RANKS, "number of ranks must be in 1..%d", TM_MAX_RANKS);
    TM_CHECK(rank >= 0 && rank < nranks, "rank %d out of range", rank);
    TM_CHECK(m->part == nullptr, "mesh is already partitioned");
    TM_CUDA(cudaSetDevice(m->device));
    tm_part* P = new tm_part();
    P->rank = rank; P->nranks = nranks;
    memset(&P->peers, 0, sizeof(P->peers));
    P->peers.rank = rank; P->peers.nranks = nranks;
    // equal slices of the Morton-ordered slots, cut at warp boundaries
    const long long nw = ((long long)m->V + 31) / 32;
    for (int r = 0; r <= nranks; r++) P->peers.bounds[r] = (int)std::min<long long>(m->V, 32 * (nw * r / nranks));
    P->peers.bounds[nranks] = m->V;
    TM_CUDA(cudaMalloc(&P->comm, sizeof(PartComm)));
    TM_CUDA(cudaMemset(P->comm, 0, sizeof(PartComm)));
    TM_CUDA(cudaHostAlloc(&P->h_err, sizeof(int), cudaHostAllocDefault));
    P->peers.pos[rank] = m->pos; P->peers.comm[rank] = P->comm; P->attached[rank] = true;
    if (const char* t = getenv("TM_PART_TIMEOUT_MS")) {
        const long long ms = atoll(t);
        if (ms > 0) P->timeout_ns = (unsigned long long)ms * 1000000ull;
    }
    m->part = P;
    // cached sweep results belong to whole-mesh sweeps
    m->geo_epoch = 0; m->frep_epoch = 0; m->grad_epoch = 0;
    return 0;
}

int tm_part_info(tm_mesh* m, int* rank, int* nranks, int* first, int* count, int* n_ghost1, int* n_pull) {
    TM_CHECK(m != nullptr && m->part != nullptr, "mesh is not partitioned");
    const tm_part* P = m->part;
    if (rank) *rank = P->rank;
    if (nranks) *nranks = P->nranks;
    if (first) *first = P->peers.bounds[P->rank];
    if (count) *count = P->peers.bounds[P->rank + 1] - P->peers.bounds[P->rank];
    if (n_ghost1) *n_ghost1 = P->lists_valid ? P->n_g1 : -1;
    if (n_pull) *n_pull = P->lists_valid ? (P->pull_all ? m->V - (P->peers.bounds[P->rank + 1] - P->peers.bounds[P->rank]) : P->n_pull) : -1;
    return 0;
}

// 128 bytes: cudaIpcMemHandle_t of the position records, then of the sequencing block
int tm_part_export(tm_mesh* m, unsigned char handle[128]) {
    TM_CHECK(m != nullptr && m->part != nullptr && handle != nullptr, "mesh is not partitioned");
    TM_CUDA(cudaSetDevice(m->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
    cudaIpcMemHandle_t h;
    TM_CUDA(cudaIpcGetMemHandle(&h, m->pos));
    memcpy(handle, &h, 64);
    TM_CUDA(cudaIpcGetMemHandle(&h, m->part->comm));
    memcpy(handle + 64, &h, 64);
    return 0;
}

int tm_part_attach_ipc(tm_mesh* m, int peer, const unsigned char handle[128]) {
    TM_CHECK(m != nullptr && m->part != nullptr && handle != nullptr, "mesh is not partitioned");
    tm_part* P = m->part;
    TM_CHECK(peer >= 0 && peer < P->nranks && peer != P->rank, "bad peer rank %d", peer);
    TM_CHECK(!P->attached[peer], "peer %d is already attached", peer);
    TM_CUDA(cudaSetDevice(m->device));
    cudaIpcMemHandle_t h;
    void *pp = nullptr, *pc = nullptr;
    memcpy(&h, handle, 64);
    TM_CUDA(cudaIpcOpenMemHandle(&pp, h, cudaIpcMemLazyEnablePeerAccess));
    memcpy(&h, handle + 64, 64);
    TM_CUDA(cudaIpcOpenMemHandle(&pc, h, cudaIpcMemLazyEnablePeerAccess));
    P->ipc_pos[peer] = pp; P->ipc_comm[peer] = pc;
    P->peers.pos[peer] = reinterpret_cast<const P4*>(pp);
    P->peers.comm[peer] = reinterpret_cast<PartComm*>(pc);
    P->attached[peer] = true;
    return 0;
}

int tm_part_attach_local(tm_mesh* m, int peer, tm_mesh* other) {
    TM_CHECK(m != nullptr && m->part != nullptr, "mesh is not partitioned");
    TM_CHECK(other != nullptr && other->part != nullptr, "peer mesh is not partitioned");
    tm_part* P = m->part;
    TM_CHECK(peer >= 0 && peer < P->nranks && peer != P->rank, "bad peer rank %d", peer);
    TM_CHECK(other->part->rank == peer && other->part->nranks == P->nranks && other->V == m->V,
             "peer mesh is rank %d of %d with %d vertices", other->part->rank, other->part->nranks, other->V);
    TM_CUDA(cudaSetDevice(m->device));
    if (other->device != m->device) {
        int can = 0;
        TM_CUDA(cudaDeviceCanAccessPeer(&can, m->device, other->device));
        TM_CHECK(can, "device %d cannot map memory of device %d", m->device, other->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) TM_CUDA(e);
        cudaGetLastError();
    }
    P->peers.pos[peer] = other->pos;
    P->peers.comm[peer] = other->part->comm;
    P->attached[peer] = true;
    return 0;
}

// local preparation: vertex order, one-rings, halo lists, every buffer the sweeps and the
// integrator allocate on demand. No peer is involved, nothing waits.
int tm_part_prepare(tm_mesh* m, tm_nlist* nl) {
    TM_CHECK(m != nullptr && m->part != nullptr, "mesh is not partitioned");
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_ensure_layout(m) || tm_check_nlist(m, nl)) return 1;
    if (tm_part_refresh(m, nl ? nl->view() : tm_no_nlist())) return 1;
    if (nl && nl->kind == TM_NLIST_CELL && nl->excl == 2 && tm_ensure_ring2(m)) return 1;   // the cell sweep's exclusion rows
    if (tm_preload_energy() || tm_preload_integrate() || tm_preload_layout()) return 1;
    TM_PRELOAD(k_publish_pos); TM_PRELOAD(k_wait_pos); TM_PRELOAD(k_ack); TM_PRELOAD(k_wait_acks);
    TM_PRELOAD(k_pull); TM_PRELOAD(k_exchange_sums); TM_PRELOAD(k_store6);
    const size_t n3 = sizeof(double) * 3 * (size_t)m->V;
    if (!m->frep) TM_CUDA(cudaMalloc(&m->frep, n3));
    if (!m->mom) { TM_CUDA(cudaMalloc(&m->mom, n3)); TM_CUDA(cudaMemsetAsync(m->mom, 0, n3, m->stream)); }
    if (!m->grad_prev) { TM_CUDA(cudaMalloc(&m->grad_prev, n3)); TM_CUDA(cudaMemsetAsync(m->grad_prev, 0, n3, m->stream)); }
    if (!m->h_stats) TM_CUDA(cudaHostAlloc(&m->h_stats, 16 * sizeof(double), cudaHostAllocDefault));
    if (!m->dot_part) TM_CUDA(cudaMalloc(&m->dot_part, 1024 * sizeof(double)));
    if (!m->d_stats) TM_CUDA(cudaMalloc(&m->d_stats, 16 * sizeof(double)));
    if (tm_props_workspace(m)) return 1;
    TM_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

// collective: every rank's whole position array becomes current (before a neighbour-list
// rebuild, a flip sweep, a download)
int tm_part_sync_positions(tm_mesh* m) {
    TM_CHECK(m != nullptr && m->part != nullptr, "mesh is not partitioned");
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_ensure_layout(m)) return 1;
    // (no acknowledgement wait: peers only ever read this rank's OWNED records, which stay as they are)
    if (exchange_positions(m, true)) return 1;
    m->pos_epoch++;
    return check_error(m);
}

// collective (tests / diagnostics): one exchange round of an arbitrary 6-vector through the
// same path the property sweeps use: out[q] = sum over ranks, in rank order
int tm_part_allsum6(tm_mesh* m, const double in[6], double out[6]) {
    TM_CHECK(m != nullptr && m->part != nullptr && in && out, "mesh is not partitioned");
    TM_CUDA(cudaSetDevice(m->device));
    Six v;
    for (int q = 0; q < 6; q++) v.v[q] = in[q];
    k_store6<<<1, 32, 0, m->stream>>>(tm_part_local_sums(m), v); TM_LAUNCHED();   // by value: no copy-engine traffic
    if (tm_part_exchange_sums(m, 0, 6)) return 1;
    m->geo_epoch = 0;                 // m->sums no longer holds property sums
    if (check_error(m)) return 1;     // waits for the exchange
    TM_CUDA(cudaMemcpy(out, m->sums, 6 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

// diagnostics: {pos_seq, sums_seq, error, host pos_seq, host sums_seq, ack[0..nranks)} of this rank's block
int tm_part_debug(tm_mesh* m, unsigned long long out[5 + 16]) {
    TM_CHECK(m != nullptr && m->part != nullptr && out, "mesh is not partitioned");
    TM_CUDA(cudaSetDevice(m->device));
    PartComm c;
    TM_CUDA(cudaMemcpy(&c, m->part->comm, sizeof(c), cudaMemcpyDeviceToHost));
    out[0] = c.pos_seq; out[1] = c.sums_seq; out[2] = (unsigned long long)c.error;
    out[3] = m->part->pos_seq; out[4] = m->part->sums_seq;
    for (int r = 0; r < 16; r++) out[5 + r] = c.ack[r];
    return 0;
}

int tm_part_check(tm_mesh* m) {
    if (!m || !m->part) return 0;
    TM_CUDA(cudaSetDevice(m->device));
    return check_error(m);
}

}  // extern "C"
