// C ABI of libtrimem_b200 (include/trimem_b200.h): handles, transfers, and the
// host-side sequencing of the sweeps. No arithmetic of the hot path lives here.
#include <cstdarg>
#include <cmath>

#include "tm_common.cuh"

std::atomic<unsigned long long> g_tm_launches{0};
thread_local unsigned long long tl_tm_launches = 0;
static thread_local char g_err[1024] = "";

void tm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- device block cache -------------------------------------------------------------
#include <map>
#include <mutex>
namespace {
struct Pool {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks[16];    // per device, keyed by size
    std::map<void*, size_t> live[16];
    size_t cached_bytes[16] = {0};
};
Pool g_pool;
constexpr size_t POOL_MAX_CACHED = (size_t)8 << 30;
}  // namespace

void* tm_pool_alloc(int device, size_t bytes) {
    if (bytes == 0) bytes = 16;
    // size classes of 1/8 octave (at most 12.5 % over): the cell arrays of a neighbour list change their size a little
    // from rebuild to rebuild with the membrane's bounding box, and an exact-size cache would miss -- a cudaMalloc per
    // rebuild and chain, and a cache that only grows
    {
        size_t step = 256;
        while (step * 16 <= bytes) step <<= 1;
        bytes = (bytes + step - 1) / step * step;
    }
    const int d = device & 15;
    {
        std::lock_guard<std::mutex> lk(g_pool.mu);
        auto it = g_pool.free_blocks[d].find(bytes);
        if (it != g_pool.free_blocks[d].end()) {
            void* p = it->second;
            g_pool.free_blocks[d].erase(it);
            g_pool.cached_bytes[d] -= bytes;
            g_pool.live[d][p] = bytes;
            return p;
        }
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        // release the cache and retry once
        std::lock_guard<std::mutex> lk(g_pool.mu);
        for (auto& kv : g_pool.free_blocks[d]) cudaFree(kv.second);
        g_pool.free_blocks[d].clear();
        g_pool.cached_bytes[d] = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    std::lock_guard<std::mutex> lk(g_pool.mu);
    g_pool.live[d][p] = bytes;
    return p;
}

void tm_pool_free(int device, void* p) {
    if (!p) return;
    const int d = device & 15;
    std::lock_guard<std::mutex> lk(g_pool.mu);
    auto it = g_pool.live[d].find(p);
    if (it == g_pool.live[d].end()) { cudaFree(p); return; }
    const size_t bytes = it->second;
    g_pool.live[d].erase(it);
    if (g_pool.cached_bytes[d] + bytes > POOL_MAX_CACHED) { cudaFree(p); return; }
    g_pool.free_blocks[d].emplace(bytes, p);
    g_pool.cached_bytes[d] += bytes;
}

// trimem_energy (kernel.h:110-122) from the six reduced properties
double tm_energy_from_props(const tm_energy_params& p, const double s[6]) {
    double e = 0.0, d;
    d = s[0] / p.ref_area - 1.0;      e += p.kappa_a * d * d;
    d = s[1] / p.ref_volume - 1.0;    e += p.kappa_v * d * d;
    d = s[2] / p.ref_curvature - 1.0; e += p.kappa_c * d * d;
    e += p.kappa_t * s[4];
    e += p.kappa_r * s[5];
    e += p.kappa_b * s[3];
    return e;
}

namespace {

template <class T>
int dev_alloc(T** p, size_t n) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (n == 0) n = 1;
    TM_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
    return 0;
}

double energy_from_props(const tm_energy_params& p, const double s[6]) { return tm_energy_from_props(p, s); }

// every compute entry: device selected, vertex order / device connectivity / staged positions
// / one-ring table brought up to date (tm_layout.cu)
int check_mesh(tm_mesh* m, bool need_layout = true, bool join = true) {
    TM_CHECK(m != nullptr, "null mesh handle");
    TM_CHECK(m->V > 0 && m->he_to_h != nullptr, "mesh has no topology (call tm_mesh_set_topology first)");
    TM_CUDA(cudaSetDevice(m->device));
    if (join && tm_join_pairs(m)) return 1;
    if (need_layout && tm_ensure_layout(m)) return 1;
    return 0;
}

int check_params(const tm_energy_params* p) {
    TM_CHECK(p != nullptr, "null parameter block");
    TM_CHECK(p->bond_type == TM_BOND_EDGE || p->bond_type == TM_BOND_AREA, "Unknown bond potential");
    return 0;
}

NlView view_of(tm_nlist* nl) { return nl ? nl->view() : tm_no_nlist(); }

int fetch_sums(tm_mesh* m, double out[6]) {
    if (tm_join_pairs(m) || tm_part_quiesce(m)) return 1;      // (a split pair sweep stores sums[5] itself)
    TM_CUDA(cudaMemcpyAsync(m->h_sums, m->sums, 6 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    for (int q = 0; q < 6; q++) out[q] = m->h_sums[q];
    return tm_part_poll_error(m);
}

}  // namespace

extern "C" {

const char* tm_last_error(void) { return g_err; }
int tm_abi_version(void) { return 4; }

int tm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void* tm_host_alloc(unsigned long long bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 8;
    if (tm_device_count() > 0 && cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) {
        // tag pinned blocks so tm_host_free knows how to release them
        return p;
    }
    cudaGetLastError();
    return malloc(bytes);
}

void tm_host_free(void* p) {
    if (!p) return;
    cudaPointerAttributes attr;
    if (tm_device_count() > 0 && cudaPointerGetAttributes(&attr, p) == cudaSuccess &&
        attr.type == cudaMemoryTypeHost) {
        cudaFreeHost(p);
        return;
    }
    cudaGetLastError();
    free(p);
}

void tm_free(void* p) { free(p); }
unsigned long long tm_launch_count(void) { return g_tm_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------- mesh --
int tm_mesh_create(int device, tm_mesh** out) {
    TM_CHECK(out != nullptr, "null output pointer");
    int n = tm_device_count();
    TM_CHECK(n > 0, "no CUDA device available: libtrimem_b200 has no CPU fallback");
    TM_CHECK(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
    TM_CUDA(cudaSetDevice(device));
    tm_mesh* m = new tm_mesh();
    m->device = device;
    TM_CUDA(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device));
    {
        // the main stream carries the dependent chain of an evaluation (property sweep -> reduction -> ...), the side
        // streams the pair sweep and what follows it, which have slack: blocks of the chain are placed first
        int lo = 0, hi = 0;
        TM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        { const char* e = getenv("TM_STREAM_PRIO"); if (e && e[0] == '0') hi = lo; }
        TM_CUDA(cudaStreamCreateWithPriority(&m->stream, cudaStreamNonBlocking, hi));
        TM_CUDA(cudaStreamCreateWithPriority(&m->stream2, cudaStreamNonBlocking, lo));
        for (int k = 0; k < TM_PAIR_STREAMS; k++) {
            TM_CUDA(cudaStreamCreateWithPriority(&m->sp_stream[k], cudaStreamNonBlocking, lo));
            TM_CUDA(cudaEventCreateWithFlags(&m->sp_ev[k], cudaEventDisableTiming));
        }
        TM_CUDA(cudaEventCreateWithFlags(&m->ev_sums, cudaEventDisableTiming));
    }
    TM_CUDA(cudaEventCreate(&m->ev0));
    TM_CUDA(cudaEventCreate(&m->ev1));
    TM_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
    TM_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    TM_CUDA(cudaEventCreateWithFlags(&m->ev_pre_upload, cudaEventDisableTiming));
    TM_CUDA(cudaMalloc(&m->sums, sizeof(double) * 8));
    TM_CUDA(cudaMalloc(&m->ticket, sizeof(unsigned) * 4));
    TM_CUDA(cudaMemset(m->ticket, 0, sizeof(unsigned) * 4));
    TM_CUDA(cudaMemset(m->sums, 0, sizeof(double) * 8));
    TM_CUDA(cudaHostAlloc(&m->h_sums, sizeof(double) * 8, cudaHostAllocDefault));
    *out = m;
    return 0;
}


int tm_mesh_destroy(tm_mesh* m) {
    if (!m) return 0;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream2);
    for (int k = 0; k < TM_PAIR_STREAMS; k++) cudaStreamSynchronize(m->sp_stream[k]);
    cudaStreamSynchronize(m->stream);
    tm_flips_release(m);
    cudaFree(m->he_to); cudaFree(m->he_next); cudaFree(m->he_face); cudaFree(m->v_he);
    cudaFree(m->ring_off); cudaFree(m->ring_v); cudaFree(m->ring_cnt); cudaFree(m->ring8);
    cudaFree(m->aux); tm_pool_free(m->device, m->gen_list);
    cudaFree(m->he_to_h); cudaFree(m->v_he_h); cudaFree(m->d2h); cudaFree(m->h2d);
    cudaFree(m->pos); cudaFree(m->pos_alt); cudaFree(m->xio3); cudaFree(m->xout3); cudaFree(m->vprops); cudaFree(m->grad); cudaFree(m->frep); cudaFree(m->ex2_off); tm_pool_free(m->device, m->ex2_v);
    cudaFree(m->partials); cudaFree(m->sums); cudaFree(m->ticket); cudaFree(m->scan_tmp);
    cudaFree(m->flush);
    tm_part_release(m);
    tm_vv_release(m);
    cudaFree(m->mom); cudaFree(m->grad_prev); cudaFree(m->dot_part); cudaFree(m->d_stats);
    if (m->h_stats) cudaFreeHost(m->h_stats);
    cudaFreeHost(m->h_sums);
    cudaEventDestroy(m->ev0); cudaEventDestroy(m->ev1);
    cudaEventDestroy(m->ev_fork); cudaEventDestroy(m->ev_join); cudaEventDestroy(m->ev_pre_upload);
    cudaStreamDestroy(m->stream2);
    for (int k = 0; k < TM_PAIR_STREAMS; k++) { cudaStreamDestroy(m->sp_stream[k]); cudaEventDestroy(m->sp_ev[k]); }
    cudaEventDestroy(m->ev_sums);
    cudaStreamDestroy(m->stream);
    delete m;
    return 0;
}

int tm_mesh_set_topology(tm_mesh* m, int V, int HE, const int* he_to, const int* he_next,
                         const int* he_face, const int* v_he) {
    TM_CHECK(m != nullptr, "null mesh handle");
    TM_CHECK(V > 0 && HE >= 0 && (HE % 2) == 0, "bad mesh sizes: %d vertices, %d half-edges", V, HE);
    TM_CHECK(V < TM_RING_FACE, "too many vertices (%d)", V);
    TM_CUDA(cudaSetDevice(m->device));
    TM_CHECK(!(m->part && V != m->V), "a partitioned mesh cannot change its vertex count");
    if (V != m->V) {
        if (dev_alloc(&m->v_he, V) || dev_alloc(&m->v_he_h, V) || dev_alloc(&m->d2h, V) || dev_alloc(&m->h2d, V) ||
            dev_alloc(&m->ring_off, (size_t)V + 1) || dev_alloc(&m->ring_cnt, (size_t)V + 1) || dev_alloc(&m->ring8, V) ||
            dev_alloc(&m->pos, V) || dev_alloc(&m->pos_alt, V) || dev_alloc(&m->xio3, 3 * (size_t)V) || dev_alloc(&m->xout3, 3 * (size_t)V) ||
            dev_alloc(&m->grad, 3 * (size_t)V))
            return 1;
        if (m->vprops) { cudaFree(m->vprops); m->vprops = nullptr; }
        if (m->aux) { cudaFree(m->aux); m->aux = nullptr; }
        if (m->mom) { cudaFree(m->mom); m->mom = nullptr; }
        if (m->grad_prev) { cudaFree(m->grad_prev); m->grad_prev = nullptr; }
        m->grad_epoch = 0;
        if (m->frep) { cudaFree(m->frep); m->frep = nullptr; }
        if (m->ex2_off) { cudaFree(m->ex2_off); m->ex2_off = nullptr; }
        m->frep_epoch = 0;
        TM_CUDA(cudaMemsetAsync(m->pos, 0, sizeof(P4) * (size_t)V, m->stream));
        TM_CUDA(cudaMemsetAsync(m->pos_alt, 0, sizeof(P4) * (size_t)V, m->stream));
        m->perm_valid = false;          // a new vertex set: the order is rebuilt from its first positions
        m->have_positions = false;
        m->stage_pending = false;
    }
    if (HE != m->HE) {
        if (dev_alloc(&m->he_to, HE) || dev_alloc(&m->he_to_h, HE) || dev_alloc(&m->he_next, HE) ||
            dev_alloc(&m->he_face, HE) || dev_alloc(&m->ring_v, HE))
            return 1;
    }
    m->V = V; m->HE = HE;
    m->has_boundary = false;
    for (int h = 0; h < HE; h++) if (he_face[h] < 0) { m->has_boundary = true; break; }
    // he_next / he_face do not hold vertex ids; he_to / v_he are kept as uploaded (host numbering)
    // and converted to the device numbering by tm_ensure_layout
    TM_CUDA(cudaMemcpyAsync(m->he_to_h, he_to, sizeof(int) * HE, cudaMemcpyHostToDevice, m->stream));
    TM_CUDA(cudaMemcpyAsync(m->he_next, he_next, sizeof(int) * HE, cudaMemcpyHostToDevice, m->stream));
    TM_CUDA(cudaMemcpyAsync(m->he_face, he_face, sizeof(int) * HE, cudaMemcpyHostToDevice, m->stream));
    TM_CUDA(cudaMemcpyAsync(m->v_he_h, v_he, sizeof(int) * V, cudaMemcpyHostToDevice, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    m->topo_dev_valid = false;
    m->raw_stale = false;
    m->rings_valid = false;
    m->pos_epoch++;
    m->topo_epoch++;
    return 0;
}

int tm_mesh_get_topology(tm_mesh* m, int* he_to, int* he_next, int* he_face, int* v_he) {
    if (check_mesh(m, false)) return 1;
    if (tm_refresh_raw_topology(m)) return 1;      // flips mutate the device-numbered arrays
    if (he_to) TM_CUDA(cudaMemcpyAsync(he_to, m->he_to_h, sizeof(int) * m->HE, cudaMemcpyDeviceToHost, m->stream));
    if (he_next) TM_CUDA(cudaMemcpyAsync(he_next, m->he_next, sizeof(int) * m->HE, cudaMemcpyDeviceToHost, m->stream));
    if (he_face) TM_CUDA(cudaMemcpyAsync(he_face, m->he_face, sizeof(int) * m->HE, cudaMemcpyDeviceToHost, m->stream));
    if (v_he) TM_CUDA(cudaMemcpyAsync(v_he, m->v_he_h, sizeof(int) * m->V, cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

int tm_mesh_get_rings(tm_mesh* m, int* ring_off, int* ring_v, int* ring_flags) {
    if (check_mesh(m)) return 1;
    const int V = m->V;
    std::vector<int> off(V + 1), rv(m->HE > 0 ? m->HE : 1), d2h(V), h2d(V);
    TM_CUDA(cudaMemcpyAsync(off.data(), m->ring_off, sizeof(int) * (V + 1), cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaMemcpyAsync(rv.data(), m->ring_v, sizeof(int) * m->HE, cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaMemcpyAsync(d2h.data(), m->d2h, sizeof(int) * V, cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaMemcpyAsync(h2d.data(), m->h2d, sizeof(int) * V, cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    // rows in host vertex order, neighbour ids in host numbering, circulation order kept
    int o = 0;
    for (int hv = 0; hv < V; hv++) {
        if (ring_off) ring_off[hv] = o;
        const int d = h2d[hv];
        for (int k = off[d]; k < off[d + 1]; k++, o++) {
            if (ring_v) ring_v[o] = d2h[rv[k] & TM_RING_MASK];
            if (ring_flags) ring_flags[o] = ((rv[k] & TM_RING_FACE) ? 1 : 0) | ((rv[k] & TM_RING_OFACE) ? 2 : 0);
        }
    }
    if (ring_off) ring_off[V] = o;
    return 0;
}

int tm_mesh_set_positions(tm_mesh* m, const double* xyz) {
    if (check_mesh(m, false)) return 1;
    TM_CHECK(xyz != nullptr, "null positions");
    // staged in host numbering; scattered into the packed device records by the next compute
    // entry on the same stream (tm_ensure_layout)
    TM_CUDA(cudaEventRecord(m->ev_pre_upload, m->stream));
    TM_CUDA(cudaMemcpyAsync(m->xio3, xyz, sizeof(double) * 3 * (size_t)m->V, cudaMemcpyHostToDevice, m->stream));
    m->stage_pending = true;
    m->have_positions = true;
    tm_part_positions_uploaded(m);
    return 0;
}

int tm_mesh_set_positions_rows(tm_mesh* m, const double* xyz, int first, int count, int last) {
    if (first == 0) {
        if (check_mesh(m, false)) return 1;
        TM_CUDA(cudaEventRecord(m->ev_pre_upload, m->stream));
    } else {
        TM_CHECK(m != nullptr && m->V > 0, "null mesh handle");
        TM_CUDA(cudaSetDevice(m->device));
    }
    TM_CHECK(xyz != nullptr, "null positions");
    TM_CHECK(first >= 0 && count >= 0 && first + count <= m->V, "rows [%d, %d) outside the mesh's %d vertices", first, first + count, m->V);
    if (count > 0)
        TM_CUDA(cudaMemcpyAsync(m->xio3 + 3 * (size_t)first, xyz + 3 * (size_t)first, sizeof(double) * 3 * (size_t)count,
                                cudaMemcpyHostToDevice, m->stream));
    if (last) {
        m->stage_pending = true;
        m->have_positions = true;
        tm_part_positions_uploaded(m);
    }
    return 0;
}

int tm_mesh_get_positions(tm_mesh* m, double* xyz) {
    if (check_mesh(m)) return 1;
    TM_CHECK(xyz != nullptr, "null positions");
    if (tm_part_make_current(m)) return 1;
    if (tm_export_positions(m) || tm_part_quiesce(m)) return 1;
    TM_CUDA(cudaMemcpyAsync(xyz, m->xout3, sizeof(double) * 3 * (size_t)m->V, cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

int tm_mesh_n_vertices(const tm_mesh* m) { return m ? m->V : 0; }
int tm_mesh_n_halfedges(const tm_mesh* m) { return m ? m->HE : 0; }

// ------------------------------------------------------- energy / gradient --
int tm_energy_async(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p) {
    if (check_mesh(m) || check_params(p) || tm_check_nlist(m, nl)) return 1;
    return tm_launch_props(m, view_of(nl), *p, true, false);
}

int tm_gradient_async(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p) {
    // (a pair sweep tm_energy_async left running on the side streams stays there when nothing is about to touch the
    // positions: the ring gradient of each range follows it on the range's stream, tm_launch_grad)
    const bool settled = m && !m->stage_pending && m->perm_valid && m->topo_dev_valid && m->rings_valid;
    if (check_mesh(m, true, !settled) || check_params(p) || tm_check_nlist(m, nl)) return 1;
    // Sweep 1: area, volume, curvature sums and rho = c/a per vertex (all the gradient needs from
    // the properties; no tether, no repulsion). They depend on positions and connectivity only,
    // so a property sweep already evaluated for the current positions (energy() immediately
    // before gradient(), the usual E+G pair) is reused. Sweep 2: the gather sweep. Both on the
    // mesh's stream, no host round trip in between.
    if (m->geo_epoch != m->pos_epoch && tm_launch_props(m, view_of(nl), *p, false, false)) return 1;
    return tm_launch_grad(m, view_of(nl), *p);
}

int tm_fetch_energy(tm_mesh* m, const tm_energy_params* p, double* energy, double props[6]) {
    if (check_mesh(m, false) || check_params(p)) return 1;
    TM_CHECK(m->sums_full, "the last property sweep was the gradient's geometry-only one (no tether / repulsion sums): "
             "call tm_energy_async before tm_fetch_energy");
    double s[6];
    if (fetch_sums(m, s)) return 1;
    if (props) for (int q = 0; q < 6; q++) props[q] = s[q];
    if (energy) *energy = energy_from_props(*p, s);
    return 0;
}

int tm_fetch_gradient(tm_mesh* m, double* grad) {
    if (check_mesh(m)) return 1;
    TM_CHECK(grad != nullptr, "null gradient buffer");
    // device order -> host vertex numbering; a partitioned mesh delivers the rows it owns, zeros elsewhere
    if (tm_export_vec3_owned(m, m->grad, grad)) return 1;
    TM_CUDA(cudaStreamSynchronize(m->stream));
    return tm_part_poll_error(m);
}

int tm_sync(tm_mesh* m) {
    if (check_mesh(m, false)) return 1;
    TM_CUDA(cudaStreamSynchronize(m->stream));
    return tm_part_poll_error(m);
}

int tm_properties(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, double out[6]) {
    if (tm_energy_async(m, nl, p)) return 1;
    return fetch_sums(m, out);
}

int tm_vertex_properties(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, double* out) {
    if (check_mesh(m) || check_params(p) || tm_check_nlist(m, nl)) return 1;
    if (tm_launch_props(m, view_of(nl), *p, true, true)) return 1;
    std::vector<double> tmp(6 * (size_t)m->V);
    std::vector<int> d2h(m->V);
    TM_CUDA(cudaMemcpyAsync(tmp.data(), m->vprops, sizeof(double) * 6 * (size_t)m->V, cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaMemcpyAsync(d2h.data(), m->d2h, sizeof(int) * m->V, cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    for (int d = 0; d < m->V; d++)
        for (int q = 0; q < 6; q++) out[6 * (size_t)d2h[d] + q] = tmp[6 * (size_t)d + q];
    return 0;
}

int tm_energy(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, double* energy, double props[6]) {
    if (tm_energy_async(m, nl, p)) return 1;
    return tm_fetch_energy(m, p, energy, props);
}

int tm_gradient(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, double* grad, double props[6]) {
    if (check_mesh(m, false) || check_params(p)) return 1;
    TM_CHECK(grad != nullptr, "null gradient buffer");
    // Speculation across the re-upload. The caller's positions cross the boundary on every call
    // (TriMesh.points() is writable host memory), but energy() followed by gradient() at ONE position
    // is the usual pair: when the property sums, rho and the pair force of the current device positions
    // are cached, the gradient sweep, its export and its download run on the second stream from the
    // positions already on the device WHILE the upload and its bitwise comparison run on the first
    // (PCIe is full duplex). Unchanged -> the result is the answer; changed -> it is discarded and the
    // normal path runs. Saves the upload's 0.3 ms of the 0.7 ms call at 655,362 vertices.
    const NlView nv = view_of(nl);
    const bool spec = m->stage_pending && !m->part && m->perm_valid && m->topo_dev_valid && m->rings_valid &&
                      m->geo_epoch == m->pos_epoch && (!props || m->sums_full) && tm_check_nlist(m, nl) == 0 &&
                      (!(p->rep_lc1 > 0.0 && p->kappa_r != 0.0) || tm_pair_force_valid(m, nv, *p));
    const unsigned long long epoch0 = m->pos_epoch;
    if (spec) {
        TM_CUDA(cudaStreamWaitEvent(m->stream2, m->ev_pre_upload, 0));
        if (tm_launch_grad_speculative(m, nv, *p, m->stream2, grad)) return 1;
    }
    if (tm_ensure_layout(m)) return 1;             // imports the upload; bumps pos_epoch if a coordinate changed
    if (spec) {
        TM_CUDA(cudaStreamSynchronize(m->stream2));
        if (m->pos_epoch == epoch0) {
            m->grad_epoch = m->pos_epoch;
            if (props) return fetch_sums(m, props);
            return 0;
        }
    }
    // `props` = all six reduced sums at the gradient's positions: the gradient's own property sweep is
    // geometry-only (tether = repulsion = 0), so a request for them takes the full sweep instead
    if (props && !(m->geo_epoch == m->pos_epoch && m->sums_full) &&
        (tm_check_nlist(m, nl) || tm_launch_props(m, nv, *p, true, false))) return 1;
    if (tm_gradient_async(m, nl, p)) return 1;
    if (props) {
        if (tm_part_quiesce(m)) return 1;
        TM_CUDA(cudaMemcpyAsync(m->h_sums, m->sums, 6 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    }
    if (tm_fetch_gradient(m, grad)) return 1;
    if (props) for (int q = 0; q < 6; q++) props[q] = m->h_sums[q];
    return 0;
}

// ---------------------------------------------------- velocity-Verlet trajectory --
int tm_mesh_set_momenta(tm_mesh* m, const double* p) {
    if (check_mesh(m)) return 1;
    TM_CHECK(p != nullptr, "null momenta");
    if (tm_import_momenta(m, p)) return 1;
    TM_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

int tm_mesh_random_momenta(tm_mesh* m, unsigned long long seed, unsigned long long sequence, double scale) {
    if (check_mesh(m)) return 1;
    TM_CHECK(!m->part, "device-drawn momenta are not available on a partitioned mesh");
    return tm_random_momenta(m, seed, sequence, scale);
}

int tm_mesh_get_momenta(tm_mesh* m, double* p) {
    if (check_mesh(m)) return 1;
    TM_CHECK(p != nullptr, "null momenta");
    TM_CHECK(m->mom != nullptr, "mesh has no momenta (call tm_mesh_set_momenta first)");
    if (tm_export_vec3_owned(m, m->mom, p)) return 1;
    TM_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

int tm_vv_integrate_async(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, double mass, double dt,
                          int n_steps, int with_energies) {
    if (check_mesh(m) || check_params(p) || tm_check_nlist(m, nl)) return 1;
    TM_CHECK(n_steps >= 0, "negative number of integration steps");
    TM_CHECK(mass > 0.0, "mass must be positive");
    TM_CHECK(m->mom != nullptr, "mesh has no momenta (call tm_mesh_set_momenta first)");
    return tm_vv_run(m, view_of(nl), *p, mass, dt, n_steps, with_energies != 0);
}

int tm_vv_stats(tm_mesh* m, const tm_energy_params* p, double stats[4], double props0[6], double props1[6]) {
    if (check_mesh(m, false) || check_params(p)) return 1;
    TM_CHECK(m->h_stats != nullptr && m->d_stats != nullptr, "no trajectory has been run on this mesh");
    TM_CUDA(cudaStreamSynchronize(m->stream));
    TM_CUDA(cudaMemcpyAsync(m->h_stats, m->d_stats, 16 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    if (tm_part_poll_error(m)) return 1;
    if (stats) {
        stats[0] = tm_energy_from_props(*p, m->h_stats);
        stats[1] = m->h_stats[12];
        stats[2] = tm_energy_from_props(*p, m->h_stats + 6);
        stats[3] = m->h_stats[13];
    }
    if (props0) for (int q = 0; q < 6; q++) props0[q] = m->h_stats[q];
    if (props1) for (int q = 0; q < 6; q++) props1[q] = m->h_stats[6 + q];
    return 0;
}

int tm_vv_integrate(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, const double* p0, double mass, double dt,
                    int n_steps, double* x_out, double* p_out, double stats[4]) {
    if (check_mesh(m) || check_params(p) || tm_check_nlist(m, nl)) return 1;
    TM_CHECK(n_steps >= 0, "negative number of integration steps");
    TM_CHECK(mass > 0.0, "mass must be positive");
    // p0 == NULL: the momenta already on the device (tm_mesh_set_momenta / tm_mesh_random_momenta)
    TM_CHECK(p0 != nullptr || m->mom != nullptr, "null momenta and none on the device");
    if (p0 && tm_import_momenta(m, p0)) return 1;
    if (tm_vv_run(m, view_of(nl), *p, mass, dt, n_steps, stats != nullptr)) return 1;
    // the rebuild trigger: displacement of the end point from the list's build positions (d_stats[14], squared)
    m->vv_disp_list = 0;
    if (stats && nl && !m->part && tm_nlist_measure_displacement(nl, m, m->d_stats + 14) == 0) m->vv_disp_list = nl->serial;
    if (p_out && tm_export_vec3_owned(m, m->mom, p_out)) return 1;
    if (x_out) {
        if (tm_part_make_current(m)) return 1;
        // xout3 is reused: the momentum copy above is ordered before this export on the stream
        if (tm_export_positions(m) || tm_part_quiesce(m)) return 1;
        TM_CUDA(cudaMemcpyAsync(x_out, m->xout3, sizeof(double) * 3 * (size_t)m->V, cudaMemcpyDeviceToHost, m->stream));
    }
    if (stats) return tm_vv_stats(m, p, stats, nullptr, nullptr);
    TM_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

int tm_vv_displacement(tm_mesh* m, tm_nlist* nl, double rep_lc1, double* displacement, double* margin) {
    TM_CHECK(m != nullptr && nl != nullptr, "null argument");
    TM_CHECK(m->h_stats != nullptr && m->vv_disp_list != 0 && m->vv_disp_list == nl->serial,
             "the last trajectory on this mesh did not measure the displacement against this list "
             "(needs tm_vv_integrate with statistics over a verlet list)");
    if (displacement) *displacement = sqrt(m->h_stats[14]);
    if (margin) *margin = nl->rlist - rep_lc1;
    return 0;
}

int tm_timer_start(tm_mesh* m) {
    if (check_mesh(m, false)) return 1;
    TM_CUDA(cudaEventRecord(m->ev0, m->stream));
    return 0;
}

int tm_timer_stop(tm_mesh* m, float* ms) {
    if (check_mesh(m, false)) return 1;
    TM_CUDA(cudaEventRecord(m->ev1, m->stream));
    TM_CUDA(cudaEventSynchronize(m->ev1));
    TM_CUDA(cudaEventElapsedTime(ms, m->ev0, m->ev1));
    return 0;
}

int tm_time_kernel(tm_mesh* m, tm_nlist* nl, const tm_energy_params* p, int which, int iters,
                   unsigned long long flush_bytes, float* ms_avg) {
    if (check_mesh(m) || check_params(p) || tm_check_nlist(m, nl)) return 1;
    TM_CHECK(iters > 0 && ms_avg, "bad arguments");
    TM_CHECK(which >= 0 && which <= 5, "which must be 0 (energy sweeps: ring properties + pairs + reduction), "
             "1 (the gradient's property sweep + reduction), 2 (ring gradient sweep), 3 (pair sweep alone), "
             "4 (energy + gradient) or 5 (a gradient evaluation at new positions: geometry sweep, pair force, ring gradient)");
    // the gradient sweep consumes rho, the reduced sums and the pair force of earlier sweeps
    if (tm_launch_props(m, view_of(nl), *p, true, false)) return 1;
    double total = 0.0;
    for (int it = 0; it < iters; it++) {
        if (flush_bytes && tm_flush_l2(m, flush_bytes)) return 1;
        TM_CUDA(cudaEventRecord(m->ev0, m->stream));
        if (which >= 4) m->frep_epoch = m->pos_epoch - 1;      // the composites re-evaluate the pair force
        int rc = which == 0 ? tm_launch_props(m, view_of(nl), *p, true, false)
               : which == 1 ? tm_launch_props(m, view_of(nl), *p, false, false)
               : which == 2 ? tm_launch_grad(m, view_of(nl), *p)
               : which == 3 ? tm_launch_pairs(m, view_of(nl), *p)
               : which == 4 ? (tm_launch_props(m, view_of(nl), *p, true, false) || tm_launch_grad(m, view_of(nl), *p))
                            : (tm_launch_props(m, view_of(nl), *p, false, false) || tm_launch_grad(m, view_of(nl), *p));
        if (rc || tm_join_pairs(m)) return 1;
        TM_CUDA(cudaEventRecord(m->ev1, m->stream));
        TM_CUDA(cudaEventSynchronize(m->ev1));
        float ms = 0.f;
        TM_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
        total += ms;
    }
    *ms_avg = (float)(total / iters);
    return 0;
}

int tm_flush_l2(tm_mesh* m, unsigned long long bytes) {
    if (check_mesh(m, false)) return 1;
    if (m->flush_bytes < bytes) {
        if (m->flush) cudaFree(m->flush);
        m->flush = nullptr;
        TM_CUDA(cudaMalloc(&m->flush, bytes));
        m->flush_bytes = bytes;
    }
    TM_CUDA(cudaMemsetAsync(m->flush, 1, bytes, m->stream));
    return 0;
}

}  // extern "C"
