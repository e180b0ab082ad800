// Shared declarations of libtrimem_b200: handle structs, device views, helpers.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/trimem_b200.h"

// ------------------------------------------------------------------ errors --
void tm_set_error(const char* fmt, ...);
extern unsigned long long g_tm_launches;

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

#define TM_LAUNCHED() (++g_tm_launches)

// --------------------------------------------------------------- constants --
constexpr int TM_RING_FACE  = 1 << 30;          // slot's outgoing half-edge has a face
constexpr int TM_RING_OFACE = (int)0x80000000u;         // its opposite half-edge has a face
constexpr int TM_RING_MASK  = (1 << 30) - 1;
constexpr int TM_PROPS_THREADS = 128;
constexpr int TM_GRAD_THREADS  = 128;
constexpr int TM_MAX_PARTIAL_BLOCKS = 1 << 17;  // 16.7M vertices at 128 threads

// ------------------------------------------------------------ device views --
struct MeshView {
    int V, HE;
    const int* he_to;
    const int* he_next;
    const int* he_face;
    const int* v_he;
    const int* ring_off;   // V+1
    const int* ring_v;     // HE entries, flags in the two top bits
    const double* x3;      // V*3 positions, row-major (host layout)
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
    // verlet rows
    const int* row_off;    // V+1
    const int* row_j;      // directed entries, ascending inside a row
};

// ----------------------------------------------------------------- handles --
struct tm_mesh {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int V = 0, HE = 0;
    int *he_to = nullptr, *he_next = nullptr, *he_face = nullptr, *v_he = nullptr;
    int *ring_off = nullptr, *ring_v = nullptr, *ring_cnt = nullptr;
    bool rings_valid = false;
    double* x3 = nullptr;        // V*3
    double* rho = nullptr;       // V   curvature/area ratio of the last props sweep
    double* vprops = nullptr;    // V*6 (allocated on demand)
    double* grad = nullptr;      // V*3
    double* partials = nullptr;  // 6 * TM_MAX_PARTIAL_BLOCKS
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
        v.ring_off = ring_off; v.ring_v = ring_v; v.x3 = x3;
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
    int *row_off = nullptr, *row_j = nullptr;
    int* cell_of = nullptr;      // build-time cell of every vertex (cell lists)
    long long n_entries = 0;
    NlView view() const {
        NlView v;
        v.kind = kind; v.excl = excl;
        for (int k = 0; k < 3; k++) {
            v.box_min[k] = box_min[k]; v.cell[k] = cell[k];
            v.shape[k] = shape[k]; v.strides[k] = strides[k];
        }
        v.cell_start = cell_start; v.cell_items = cell_items;
        v.row_off = row_off; v.row_j = row_j;
        return v;
    }
};

inline NlView tm_no_nlist() {
    NlView v;
    memset(&v, 0, sizeof(v));
    v.kind = -1;
    return v;
}

// ------------------------------------------------- internal host interfaces --
int tm_ensure_rings(tm_mesh* m);                                 // tm_rings.cu
int tm_exclusive_scan(tm_mesh* m, const int* in, int* out, int n, long long* total);  // tm_rings.cu
int tm_launch_props(tm_mesh* m, const NlView& nl, const tm_energy_params& p,
                    bool full, bool write_vprops);               // tm_energy.cu
int tm_launch_grad(tm_mesh* m, const NlView& nl, const tm_energy_params& p);  // tm_energy.cu
int tm_flips_release(tm_mesh* m);                               // tm_flips.cu

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
__device__ __forceinline__ d3 ldpos(const double* __restrict__ x3, int i) {
    const double* p = x3 + 3 * (size_t)i;
    return mk3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
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
