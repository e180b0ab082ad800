// Device data layout: vertex renumbering, packed position records, boundary
// conversions between the host (OpenMesh) numbering and the device numbering.
//
// The reference keeps positions and connectivity inside OpenMesh's array kernel
// (src/mesh.h:15-31) in whatever order the input file lists the vertices, and every
// sweep chases handles through it (mesh_properties.cpp:19,73). On the device the
// sweeps are gathers over ring neighbours and neighbour-list entries, so their cost
// is set by how many distinct 32-byte sectors a warp touches. Vertices are therefore
// renumbered once per mesh by the Morton code of a coarse grid cell (a counting sort:
// histogram -> scan -> fill -> ascending ids inside a cell, deterministic), so the 32
// vertices of a warp are a compact patch of the surface. Nothing of this is visible
// at the C ABI: ids are converted on the way in and out.
#include <cfloat>
#include <algorithm>
#include <atomic>
#include <cmath>

#include "tm_common.cuh"

namespace {

constexpr int NT = 256;

// ---- bounding box of `V` points stored with a stride of `stride` doubles -----------
__global__ void k_bbox_partial(const double* __restrict__ x, int stride, int V, double* __restrict__ part) {
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int i = blockIdx.x * NT + threadIdx.x; i < V; i += gridDim.x * NT) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double v = x[(size_t)stride * i + k];
            mn[k] = fmin(mn[k], v);
            mx[k] = fmax(mx[k], v);
        }
    }
    __shared__ double smn[3][NT / 32], smx[3][NT / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if (lane == 0) { smn[k][w] = mn[k]; smx[k][w] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int k = threadIdx.x;
        double a = smn[k][0], b = smx[k][0];
        for (int q = 1; q < NT / 32; q++) { a = fmin(a, smn[k][q]); b = fmax(b, smx[k][q]); }
        part[blockIdx.x * 6 + k] = a;
        part[blockIdx.x * 6 + 3 + k] = b;
    }
}

__global__ void k_bbox_final(const double* __restrict__ part, int nb, double* __restrict__ out) {
    // 6 warps, one per output; lanes stride over the block partials (min/max: order-free)
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= 6) return;
    double a = k < 3 ? DBL_MAX : -DBL_MAX;
    for (int b = lane; b < nb; b += 32) a = k < 3 ? fmin(a, part[b * 6 + k]) : fmax(a, part[b * 6 + k]);
    for (int o = 16; o > 0; o >>= 1) {
        double t = __shfl_xor_sync(0xffffffffu, a, o);
        a = k < 3 ? fmin(a, t) : fmax(a, t);
    }
    if (lane == 0) out[k] = a;
}

// ascending ids inside every segment (cells / verlet rows): deterministic output
__global__ void k_sort_segments(long long nseg, const int* __restrict__ start, int* __restrict__ items) {
    long long c = (long long)blockIdx.x * NT + threadIdx.x;
    if (c >= nseg) return;
    int b = start[c], e = start[c + 1];
    for (int a = b + 1; a < e; a++) {
        int v = items[a], q = a - 1;
        while (q >= b && items[q] > v) { items[q + 1] = items[q]; --q; }
        items[q + 1] = v;
    }
}

// ---- Morton cell keys ----------------------------------------------------------------
__device__ __forceinline__ unsigned spread3(unsigned v) {   // 8 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

struct SortGrid {
    double lo[3], inv[3];   // cell = (x - lo) * inv, clamped to [0, G)
    int G;
};

__global__ void k_key_count(SortGrid g, const double* __restrict__ x3, int V, int* __restrict__ key,
                            int* __restrict__ count) {
    int i = blockIdx.x * NT + threadIdx.x;
    if (i >= V) return;
    unsigned c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int q = (int)((x3[3 * (size_t)i + k] - g.lo[k]) * g.inv[k]);
        c[k] = (unsigned)min(max(q, 0), g.G - 1);
    }
    int kk = (int)(spread3(c[0]) | (spread3(c[1]) << 1) | (spread3(c[2]) << 2));
    key[i] = kk;
    atomicAdd(count + kk, 1);
}

__global__ void k_key_fill(int V, const int* __restrict__ key, const int* __restrict__ start,
                           int* __restrict__ cursor, int* __restrict__ items) {
    int i = blockIdx.x * NT + threadIdx.x;
    if (i >= V) return;
    int c = key[i];
    items[start[c] + atomicAdd(cursor + c, 1)] = i;
}

__global__ void k_invert(int V, const int* __restrict__ d2h, int* __restrict__ h2d) {
    int d = blockIdx.x * NT + threadIdx.x;
    if (d < V) h2d[d2h[d]] = d;
}

__global__ void k_iota(int V, int* __restrict__ a, int* __restrict__ b) {
    int d = blockIdx.x * NT + threadIdx.x;
    if (d < V) { a[d] = d; b[d] = d; }
}

// ---- numbering conversions -----------------------------------------------------------
__global__ void k_map_ids(int n, const int* __restrict__ in, const int* __restrict__ map, int* __restrict__ out) {
    int k = blockIdx.x * NT + threadIdx.x;
    if (k < n) { int v = in[k]; out[k] = v >= 0 ? map[v] : v; }
}
// out[k] = in[map[k]]  (per-vertex arrays whose VALUES are not vertex ids)
__global__ void k_gather_int(int n, const int* __restrict__ in, const int* __restrict__ map, int* __restrict__ out) {
    int k = blockIdx.x * NT + threadIdx.x;
    if (k < n) out[k] = in[map[k]];
}

// changed (optional): set to 1 when any coordinate differs bitwise from the record it replaces.
__global__ void k_import_positions(int V, const double* __restrict__ x3, const int* __restrict__ d2h,
                                   P4* __restrict__ pos, int* __restrict__ changed) {
    int d = blockIdx.x * NT + threadIdx.x;
    bool diff = false;
    if (d < V) {
        const double* s = x3 + 3 * (size_t)d2h[d];
        const double sx = s[0], sy = s[1], sz = s[2];
        if (changed) {
            const d3 old = ldpos(pos, d);
            diff = __double_as_longlong(old.x) != __double_as_longlong(sx) ||
                   __double_as_longlong(old.y) != __double_as_longlong(sy) ||
                   __double_as_longlong(old.z) != __double_as_longlong(sz);
        }
        // rho (.w) is left alone: it belongs to the last property sweep
        if (!changed || diff) {
            *reinterpret_cast<double2*>(pos + d) = make_double2(sx, sy);
            pos[d].z = sz;
        }
    }
    if (changed && __any_sync(0xffffffffu, diff) && (threadIdx.x & 31) == 0) *changed = 1;
}

__global__ void k_export_positions(int V, const P4* __restrict__ pos, const int* __restrict__ d2h,
                                   double* __restrict__ x3) {
    int d = blockIdx.x * NT + threadIdx.x;
    if (d >= V) return;
    const d3 p = ldpos(pos, d);
    double* o = x3 + 3 * (size_t)d2h[d];
    o[0] = p.x; o[1] = p.y; o[2] = p.z;
}

__global__ void k_export_vec3(int V, const double* __restrict__ in, const int* __restrict__ d2h,
                              double* __restrict__ out) {
    int d = blockIdx.x * NT + threadIdx.x;
    if (d >= V) return;
    const double* s = in + 3 * (size_t)d;
    double* o = out + 3 * (size_t)d2h[d];
    o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
}

template <class T>
struct Tmp {
    T* p = nullptr;
    ~Tmp() { if (p) cudaFree(p); }
    int alloc(size_t n) { TM_CUDA(cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T))); return 0; }
};

// Permutation from the staged (host-numbered) positions.
int build_permutation(tm_mesh* m) {
    const int V = m->V, vb = (V + NT - 1) / NT;
    // grid resolution: ~ a handful of vertices per occupied cell for a surface-like point set
    int G = 1;
    while (G < 256 && (double)(2 * G) * (2 * G) * 16.0 <= (double)V) G *= 2;
    if (G == 1) {
        k_iota<<<vb, NT, 0, m->stream>>>(V, m->d2h, m->h2d); TM_LAUNCHED();
        TM_CUDA(cudaGetLastError());
        return 0;
    }
    double box[6];
    if (tm_bbox(m, m->xio3, 3, V, box)) return 1;
    SortGrid g;
    g.G = G;
    for (int k = 0; k < 3; k++) {
        TM_CHECK(std::isfinite(box[k]) && std::isfinite(box[3 + k]), "vertex positions are not finite");
        double dim = box[3 + k] - box[k];
        g.lo[k] = box[k];
        g.inv[k] = dim > 0.0 ? G / dim : 0.0;
    }
    const size_t ncell = (size_t)G * G * G;
    Tmp<int> key, count, start;
    if (key.alloc(V) || count.alloc(ncell + 1) || start.alloc(ncell + 1)) return 1;
    TM_CUDA(cudaMemsetAsync(count.p, 0, sizeof(int) * (ncell + 1), m->stream));
    k_key_count<<<vb, NT, 0, m->stream>>>(g, m->xio3, V, key.p, count.p); TM_LAUNCHED();
    if (tm_exclusive_scan(m, count.p, start.p, (int)ncell, nullptr)) return 1;
    TM_CUDA(cudaMemsetAsync(count.p, 0, sizeof(int) * (ncell + 1), m->stream));
    k_key_fill<<<vb, NT, 0, m->stream>>>(V, key.p, start.p, count.p, m->d2h); TM_LAUNCHED();
    if (tm_sort_segments(m, (long long)ncell, start.p, m->d2h)) return 1;
    k_invert<<<vb, NT, 0, m->stream>>>(V, m->d2h, m->h2d); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    TM_CUDA(cudaStreamSynchronize(m->stream));   // temporaries are released on return
    return 0;
}

}  // namespace

int tm_bbox(tm_mesh* m, const double* x, int stride, int V, double out[6]) {
    int nb = std::min((V + NT - 1) / NT, 1024);
    Tmp<double> part, box;
    if (part.alloc((size_t)nb * 6) || box.alloc(6)) return 1;
    k_bbox_partial<<<nb, NT, 0, m->stream>>>(x, stride, V, part.p); TM_LAUNCHED();
    k_bbox_final<<<1, 192, 0, m->stream>>>(part.p, nb, box.p); TM_LAUNCHED();
    TM_CUDA(cudaMemcpyAsync(out, box.p, sizeof(double) * 6, cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

int tm_sort_segments(tm_mesh* m, long long nseg, const int* start, int* items) {
    if (nseg <= 0) return 0;
    k_sort_segments<<<(unsigned)((nseg + NT - 1) / NT), NT, 0, m->stream>>>(nseg, start, items); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    return 0;
}

// Host-numbered copies of the connectivity from the (flip-mutated) device-numbered arrays.
int tm_refresh_raw_topology(tm_mesh* m) {
    if (!m->raw_stale) return 0;
    TM_CHECK(m->perm_valid && m->topo_dev_valid, "internal: stale raw topology without a device layout");
    k_map_ids<<<(m->HE + NT - 1) / NT, NT, 0, m->stream>>>(m->HE, m->he_to, m->d2h, m->he_to_h); TM_LAUNCHED();
    k_gather_int<<<(m->V + NT - 1) / NT, NT, 0, m->stream>>>(m->V, m->v_he, m->h2d, m->v_he_h); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    m->raw_stale = false;
    return 0;
}

int tm_ensure_layout(tm_mesh* m) {
    TM_CHECK(m->V > 0 && m->he_to_h, "mesh has no topology (call tm_mesh_set_topology first)");
    TM_CHECK(m->have_positions, "mesh has no positions (call tm_mesh_set_positions first)");
    const int V = m->V, vb = (V + NT - 1) / NT;
    if (!m->perm_valid) {
        TM_CHECK(m->stage_pending, "internal: no staged positions to build the vertex order from");
        if (build_permutation(m)) return 1;
        m->perm_valid = true;
        m->topo_dev_valid = false;
        // unique over all mesh handles of the process: a neighbour list built on another handle (rows in that
        // handle's vertex order) can then never pass tm_check_nlist by coincidence
        static std::atomic<unsigned long long> g_layout{0};
        m->layout_epoch = ++g_layout;
    }
    if (!m->topo_dev_valid) {
        k_map_ids<<<(m->HE + NT - 1) / NT, NT, 0, m->stream>>>(m->HE, m->he_to_h, m->h2d, m->he_to); TM_LAUNCHED();
        k_gather_int<<<vb, NT, 0, m->stream>>>(V, m->v_he_h, m->d2h, m->v_he); TM_LAUNCHED();
        TM_CUDA(cudaGetLastError());
        m->topo_dev_valid = true;
        m->raw_stale = false;
        m->rings_valid = false;
        m->pos_epoch++;
        m->topo_epoch++;
    }
    if (m->stage_pending) {
        // Positions cross the C ABI on every evaluation (the caller may have written through
        // TriMesh.points()). When results cached for the current positions exist (the sums / rho
        // of a property sweep, the pair force), compare while importing: an unchanged upload --
        // energy() then gradient() at one position, the usual pair -- keeps them valid.
        const bool have_cache = m->geo_epoch == m->pos_epoch || m->frep_epoch == m->pos_epoch;
        if (have_cache) {
            int* flag = reinterpret_cast<int*>(m->ticket) + 2;
            TM_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), m->stream));
            k_import_positions<<<vb, NT, 0, m->stream>>>(V, m->xio3, m->d2h, m->pos, flag); TM_LAUNCHED();
            int changed = 1;
            TM_CUDA(cudaMemcpyAsync(&changed, flag, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
            TM_CUDA(cudaStreamSynchronize(m->stream));
            if (changed) m->pos_epoch++;
        } else {
            k_import_positions<<<vb, NT, 0, m->stream>>>(V, m->xio3, m->d2h, m->pos, nullptr); TM_LAUNCHED();
            m->pos_epoch++;
        }
        TM_CUDA(cudaGetLastError());
        m->stage_pending = false;
    }
    return tm_ensure_rings(m);
}

int tm_export_positions(tm_mesh* m) {
    k_export_positions<<<(m->V + NT - 1) / NT, NT, 0, m->stream>>>(m->V, m->pos, m->d2h, m->xout3); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    return 0;
}

int tm_export_gradient(tm_mesh* m) {
    k_export_vec3<<<(m->V + NT - 1) / NT, NT, 0, m->stream>>>(m->V, m->grad, m->d2h, m->xout3); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    return 0;
}

int tm_check_nlist(const tm_mesh* m, const tm_nlist* nl) {
    if (!nl) return 0;
    TM_CHECK(nl->V == m->V, "neighbour list was built for %d vertices, mesh has %d", nl->V, m->V);
    TM_CHECK(nl->layout_epoch == m->layout_epoch,
             "neighbour list is stale: it was built on another mesh object or before the mesh's vertex order "
             "changed (call update_repulsion(mesh) / rebuild it)");
    if (nl->kind == TM_NLIST_VERLET) {
        const SweepRange R = tm_owned_range(m);
        TM_CHECK(nl->rows_first <= R.first && nl->rows_first + nl->rows_count >= R.first + R.count,
                 "verlet list holds the rows of vertices [%d, %d) but the mesh now evaluates [%d, %d): rebuild it",
                 nl->rows_first, nl->rows_first + nl->rows_count, R.first, R.first + R.count);
    }
    return 0;
}

int tm_preload_layout() {
    TM_PRELOAD(k_import_positions); TM_PRELOAD(k_export_positions); TM_PRELOAD(k_export_vec3);
    return 0;
}

extern "C" {

// Recompute the vertex order from the current positions (extension; e.g. after long runs in
// which vertices have diffused across the surface). Neighbour lists must be rebuilt afterwards.
int tm_mesh_reorder(tm_mesh* m) {
    TM_CHECK(m != nullptr, "null mesh handle");
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_ensure_layout(m)) return 1;
    if (tm_refresh_raw_topology(m) || tm_export_positions(m)) return 1;
    TM_CUDA(cudaMemcpyAsync(m->xio3, m->xout3, sizeof(double) * 3 * (size_t)m->V, cudaMemcpyDeviceToDevice, m->stream));
    m->stage_pending = true;
    m->perm_valid = false;
    return tm_ensure_layout(m);
}

// d2h[d] = host vertex id stored at device slot d (diagnostics / tests)
int tm_mesh_get_vertex_order(tm_mesh* m, int* d2h) {
    TM_CHECK(m != nullptr && d2h != nullptr, "null argument");
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_ensure_layout(m)) return 1;
    TM_CUDA(cudaMemcpyAsync(d2h, m->d2h, sizeof(int) * m->V, cudaMemcpyDeviceToHost, m->stream));
    TM_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

}  // extern "C"
