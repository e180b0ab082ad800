// One-ring CSR construction from the device half-edge arrays, and the integer
// exclusive scan shared by the ring and neighbour-list builders.
//
// The reference walks OpenMesh circulators inside every sweep
// (mesh_properties.cpp:19,73; cell_list.h:181-199). Here the circulation
// h -> next(opposite(h)) (clockwise, starting at the vertex's stored half-edge,
// SURVEY.md App. B) is flattened once per topology change into
//   ring_off[V+1], ring_v[HE]  (neighbour id | face flags)
// so that the sweeps read contiguous rows.
#include "tm_common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                       // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void k_ring_count(int V, const int* __restrict__ v_he,
                             const int* __restrict__ he_next, int* __restrict__ cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    int h0 = v_he[i], n = 0;
    if (h0 >= 0) {
        int h = h0;
        do { ++n; h = he_next[h ^ 1]; } while (h != h0 && n < (1 << 20));
    }
    cnt[i] = n;
}

__global__ void k_ring_fill(int V, const int* __restrict__ v_he, const int* __restrict__ he_next,
                            const int* __restrict__ he_to, const int* __restrict__ he_face,
                            const int* __restrict__ off, int* __restrict__ ring_v, R8* __restrict__ ring8,
                            int* __restrict__ is_generic) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    R8 row;
#pragma unroll
    for (int k = 0; k < 8; k++) row.v[k] = TM_RING_PAD;
    int h0 = v_he[i];
    int o = off[i], end = off[i + 1], h = h0, n = 0;
    bool closed = true;
    if (h0 >= 0) {
        do {
            const int to = he_to[h];
            int v = to;
            if (he_face[h] >= 0) v |= TM_RING_FACE; else closed = false;
            if (he_face[h ^ 1] >= 0) v |= TM_RING_OFACE; else closed = false;
            ring_v[o++] = v;
#pragma unroll
            for (int k = 0; k < 8; k++) if (k == n) row.v[k] = to;
            ++n;
            h = he_next[h ^ 1];
        } while (h != h0 && o < end);
    }
    // the fixed-width row serves the closed-fan sweeps only (tm_energy.cu)
    const bool generic = n > 8 || n < 3 || !closed;
    if (generic) row.v[0] = TM_RING_GENERIC;
    ring8[i] = row;
    is_generic[i] = generic ? 1 : 0;
}

// ascending list of the vertices without a fixed-width row (off = exclusive scan of the flags)
__global__ void k_generic_list(int V, const R8* __restrict__ ring8, const int* __restrict__ off, int* __restrict__ list) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    if (ring8[i].v[0] == TM_RING_GENERIC) list[off[i]] = i;
}

// --- sorted two-ring rows ------------------------------------------------------------------
// all vertices within two rings of i (i itself excluded), ascending and de-duplicated: the
// exclusion set of level 2 (cell_list.h:181-199), looked up by binary search in the verlet build
constexpr int RING2_CAP = 160;
__device__ int collect_ring2(int i, const int* __restrict__ ring_off, const int* __restrict__ ring_v, int* buf) {
    int n = 0;
    for (int s = ring_off[i]; s < ring_off[i + 1]; s++) {
        const int v = ring_v[s] & TM_RING_MASK;
        if (n < RING2_CAP) buf[n] = v;
        ++n;
        for (int t = ring_off[v]; t < ring_off[v + 1]; t++) {
            const int w = ring_v[t] & TM_RING_MASK;
            if (w == i) continue;
            if (n < RING2_CAP) buf[n] = w;
            ++n;
        }
    }
    if (n > RING2_CAP) return -1;
    for (int a = 1; a < n; a++) {
        int v = buf[a], q = a - 1;
        while (q >= 0 && buf[q] > v) { buf[q + 1] = buf[q]; --q; }
        buf[q + 1] = v;
    }
    int m = 0;
    for (int a = 0; a < n; a++)
        if (m == 0 || buf[m - 1] != buf[a]) buf[m++] = buf[a];
    return m;
}

__global__ void k_ring2_count(int V, const int* __restrict__ ring_off, const int* __restrict__ ring_v,
                              int* __restrict__ cnt, int* __restrict__ err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    int buf[RING2_CAP];
    int n = collect_ring2(i, ring_off, ring_v, buf);
    if (n < 0) { *err = 1; n = 0; }
    cnt[i] = n;
}

__global__ void k_ring2_fill(int V, const int* __restrict__ ring_off, const int* __restrict__ ring_v,
                             const int* __restrict__ off, int* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    int buf[RING2_CAP];
    const int n = collect_ring2(i, ring_off, ring_v, buf);
    int* o = out + off[i];
    for (int a = 0; a < n; a++) o[a] = buf[a];
}

// --- three-phase exclusive scan over int (sizes here are <= 64M elements) ---
__device__ __forceinline__ int warp_incl_scan(int v) {
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns block total in *total
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
    __shared__ int wsum[SCAN_THREADS / 32];
    __shared__ int tot;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = warp_incl_scan(v);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
        int is = warp_incl_scan(s);
        if (lane < SCAN_THREADS / 32) wsum[lane] = is - s;
        if (lane == SCAN_THREADS / 32 - 1) tot = is;
    }
    __syncthreads();
    int r = inc - v + wsum[w];
    *total = tot;
    __syncthreads();
    return r;
}

__global__ void k_scan_tile_sums(const int* __restrict__ in, int n, int* __restrict__ tile_sums) {
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) s += (base + k < n) ? in[base + k] : 0;
    int tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of tile sums in place; writes grand total (as two ints) behind
__global__ void k_scan_tile_offsets(int* __restrict__ tile_sums, int ntiles, long long* __restrict__ total) {
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += SCAN_THREADS) {
        int idx = base + threadIdx.x;
        int v = idx < ntiles ? tile_sums[idx] : 0;
        int tot;
        int ex = block_excl_scan(v, &tot);
        long long c = carry;
        if (idx < ntiles) tile_sums[idx] = (int)(c + ex);
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void k_scan_apply(const int* __restrict__ in, int* __restrict__ out, int n,
                             const int* __restrict__ tile_off, const long long* __restrict__ total) {
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : 0; s += v[k]; }
    int tot;
    int ex = block_excl_scan(s, &tot) + tile_off[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (int)(*total);   // out has n+1 entries
}

}  // namespace

// out must hold n+1 ints; out[n] = total. in and out may alias.
int tm_exclusive_scan(tm_mesh* m, const int* in, int* out, int n, long long* total) {
    int ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (ntiles < 1) ntiles = 1;
    size_t need = (size_t)ntiles + 4;
    if (m->scan_tmp_n < need) {
        if (m->scan_tmp) cudaFree(m->scan_tmp);
        TM_CUDA(cudaMalloc(&m->scan_tmp, need * sizeof(int) + 16));
        m->scan_tmp_n = need;
    }
    // the grand total lives (8-byte aligned) behind the tile sums
    long long* d_total = (long long*)(m->scan_tmp + ((ntiles + 1) & ~1));
    k_scan_tile_sums<<<ntiles, SCAN_THREADS, 0, m->stream>>>(in, n, m->scan_tmp); TM_LAUNCHED();
    k_scan_tile_offsets<<<1, SCAN_THREADS, 0, m->stream>>>(m->scan_tmp, ntiles, d_total); TM_LAUNCHED();
    k_scan_apply<<<ntiles, SCAN_THREADS, 0, m->stream>>>(in, out, n, m->scan_tmp, d_total); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    if (total) {
        TM_CUDA(cudaMemcpyAsync(total, d_total, sizeof(long long), cudaMemcpyDeviceToHost, m->stream));
        TM_CUDA(cudaStreamSynchronize(m->stream));
    }
    return 0;
}

int tm_ensure_rings(tm_mesh* m) {
    if (m->rings_valid) return 0;
    TM_CHECK(m->V > 0 && m->he_to && m->topo_dev_valid, "mesh has no device connectivity (internal: tm_ensure_layout first)");
    int nb = (m->V + 255) / 256;
    k_ring_count<<<nb, 256, 0, m->stream>>>(m->V, m->v_he, m->he_next, m->ring_cnt); TM_LAUNCHED();
    long long total = 0;
    if (tm_exclusive_scan(m, m->ring_cnt, m->ring_off, m->V, &total)) return 1;
    TM_CHECK(total <= m->HE, "one-ring walk visited %lld half-edges but the mesh has %d "
             "(non-manifold vertex or corrupt connectivity)", total, m->HE);
    k_ring_fill<<<nb, 256, 0, m->stream>>>(m->V, m->v_he, m->he_next, m->he_to, m->he_face,
                                          m->ring_off, m->ring_v, m->ring8, m->ring_cnt); TM_LAUNCHED();
    // the vertices the closed-fan sweeps skip (valence > 8 or < 3, boundary vertices of open meshes), in
    // ascending order: the sweeps' generic kernels run over this list and only when it is not empty
    long long ngen = 0;
    if (tm_exclusive_scan(m, m->ring_cnt, m->ring_cnt, m->V, &ngen)) return 1;     // syncs the stream
    if (m->gen_list) { tm_pool_free(m->device, m->gen_list); m->gen_list = nullptr; }
    m->n_generic = (int)ngen;
    if (ngen > 0) {
        m->gen_list = (int*)tm_pool_alloc(m->device, sizeof(int) * (size_t)ngen);
        TM_CHECK(m->gen_list != nullptr, "out of device memory for the generic-vertex list");
        k_generic_list<<<nb, 256, 0, m->stream>>>(m->V, m->ring8, m->ring_cnt, m->gen_list); TM_LAUNCHED();
    }
    TM_CUDA(cudaGetLastError());
    m->rings_valid = true;
    return 0;
}


int tm_ensure_ring2(tm_mesh* m) {
    if (tm_ensure_rings(m)) return 1;
    if (m->ex2_off && m->ex2_epoch == m->topo_epoch) return 0;
    const int V = m->V, nb = (V + 127) / 128;
    if (!m->ex2_off) TM_CUDA(cudaMalloc(&m->ex2_off, sizeof(int) * ((size_t)V + 2)));
    int* err = m->ex2_off + V + 1;
    TM_CUDA(cudaMemsetAsync(err, 0, sizeof(int), m->stream));
    k_ring2_count<<<nb, 128, 0, m->stream>>>(V, m->ring_off, m->ring_v, m->ex2_off, err); TM_LAUNCHED();
    int herr = 0;
    TM_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    long long total = 0;
    if (tm_exclusive_scan(m, m->ex2_off, m->ex2_off, V, &total)) return 1;   // syncs the stream
    TM_CHECK(herr == 0, "a vertex has more than %d vertices within two rings", RING2_CAP);
    // from the block cache: rebuilt after every flip sweep, and cudaFree would synchronise the whole device
    // (the stream is idle here: the scan above waited for it)
    if (m->ex2_v) { tm_pool_free(m->device, m->ex2_v); m->ex2_v = nullptr; }
    m->ex2_v = (int*)tm_pool_alloc(m->device, sizeof(int) * (size_t)(total > 0 ? total : 1));
    TM_CHECK(m->ex2_v != nullptr, "out of device memory for the two-ring rows");
    k_ring2_fill<<<nb, 128, 0, m->stream>>>(V, m->ring_off, m->ring_v, m->ex2_off, m->ex2_v); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    m->ex2_epoch = m->topo_epoch;
    return 0;
}
