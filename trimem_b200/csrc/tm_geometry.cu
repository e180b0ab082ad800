// Setup-time reductions and single half-edge primitives.
//
// Replaces, in the reference:
//   area / edges_length                         src/util.cpp:9-35
//   edge_length, face_area, face_volume, dihedral_angle and their gradients
//                                               src/mesh_util.h:15-185
// (the primitives are exposed to Python one half-edge at a time by
// src/bindings.cpp:249-407 and pinned by tests/test_properties.py).
#include "tm_common.cuh"

namespace {

constexpr int NT = 256;

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sm[NT / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = lane < NT / 32 ? sm[lane] : 0.0;
        for (int o = NT / 64; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;
}

// per-block partials of (sum of face areas, sum of edge lengths); grid-stride so the
// partial count is fixed and the final pass is a fixed-order sum
__global__ void k_area_edges_partial(MeshView mv, double* __restrict__ part) {
    double a = 0.0, l = 0.0;
    for (int h = blockIdx.x * NT + threadIdx.x; h < mv.HE; h += gridDim.x * NT) {
        int to = mv.he_to[h], from = mv.he_to[h ^ 1];
        d3 x0 = ldpos(mv.pos, from), x1 = ldpos(mv.pos, to);
        if ((h & 1) == 0) { d3 e = x1 - x0; l += sqrt(dot3(e, e)); }
        if (mv.he_face[h] >= 0) {
            d3 x2 = ldpos(mv.pos, mv.he_to[mv.he_next[h]]);
            d3 n = cross3(x2 - x1, x0 - x1);
            a += 0.5 * sqrt(dot3(n, n)) / 3.0;      // every face is visited by its 3 half-edges
        }
    }
    a = block_sum(a);
    l = block_sum(l);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = l; }
}

__global__ void k_area_edges_final(const double* __restrict__ part, int nb, double* __restrict__ out) {
    double a = 0.0, l = 0.0;
    for (int b = threadIdx.x; b < nb; b += NT) { a += part[2 * b]; l += part[2 * b + 1]; }
    a = block_sum(a);
    l = block_sum(l);
    if (threadIdx.x == 0) { out[0] = a; out[1] = l; }
}

__device__ __forceinline__ void st3(double* o, d3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }

// half-edge h: i -> j, apex k = to(next h); opposite half-edge j -> i with apex m
__global__ void k_halfedge_geometry(MeshView mv, int h, double* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int o = h ^ 1;
    const int i = mv.he_to[o], j = mv.he_to[h];
    const int k = mv.he_to[mv.he_next[h]], m = mv.he_to[mv.he_next[o]];
    const d3 xi = ldpos(mv.pos, i), xj = ldpos(mv.pos, j), xk = ldpos(mv.pos, k), xm = ldpos(mv.pos, m);
    const bool boundary = mv.he_face[h] < 0 || mv.he_face[o] < 0;

    // edge (mesh_util.h:30-56)
    const d3 e = xj - xi;
    const double l = sqrt(dot3(e, e));
    const d3 eh = (1.0 / l) * e;
    out[0] = l;
    st3(out + 4, -1.0 * eh);
    st3(out + 7, eh);

    // face (i,j,k): sector normal of h (mesh_util.h:58-92)
    const d3 n0 = cross3(xk - xj, xi - xj);
    const double nn0 = dot3(n0, n0), ln0 = sqrt(nn0);
    const d3 u0 = (1.0 / ln0) * n0;
    out[1] = 0.5 * ln0;
    st3(out + 10, 0.5 * cross3(u0, xk - xj));
    // edge_vector(prev h): k -> i in a triangle; on a boundary half-edge the loop's previous half-edge u -> i
    d3 e_prev = xi - xk;
    if (mv.he_face[h] < 0) {
        int q = h;
        while (mv.he_next[q] != h) q = mv.he_next[q];
        e_prev = xi - ldpos(mv.pos, mv.he_to[q ^ 1]);
    }
    st3(out + 13, 0.5 * cross3(u0, e_prev));
    st3(out + 16, 0.5 * cross3(u0, xj - xi));

    // volume (mesh_util.h:94-128)
    out[2] = dot3(cross3(xj, xk), xi) / 6.0;
    st3(out + 19, (1.0 / 6.0) * cross3(xj, xk));
    st3(out + 22, (1.0 / 6.0) * cross3(xk, xi));
    st3(out + 25, (1.0 / 6.0) * cross3(xi, xj));

    // dihedral angle and its gradient (mesh_util.h:130-185): cot of a sector angle is
    // dot / |cross|, and |cross| is the face's un-normalised normal length
    const d3 n1 = cross3(xm - xi, xj - xi);
    const double nn1 = dot3(n1, n1), ln1 = sqrt(nn1);
    const d3 u1 = (1.0 / ln1) * n1;
    double theta = 0.0;
    if (!boundary && ln0 * ln1 != 0.0)
        theta = atan2(dot3(cross3(n0, n1), e) / l, dot3(n0, n1));
    out[3] = theta;
    double cot01 = dot3(xj - xi, xk - xi) / ln0;   // a01: angle at i in face (i,j,k)
    double cot03 = dot3(xk - xj, xi - xj) / ln0;   // a03: angle at j in face (i,j,k)
    double cot02 = dot3(xm - xi, xj - xi) / ln1;   // a02: angle at i in face (j,i,m)
    double cot04 = dot3(xi - xj, xm - xj) / ln1;   // a04: angle at j in face (j,i,m)
    // Boundary edges. The reference evaluates the four sector angles through OpenMesh, whose sectors of a
    // boundary half-edge b run along the boundary loop: calc_sector_angle(b) is the angle between
    // next(b) and opposite(b), signed against the normal of the face across the edge, and prev(b) is the
    // loop's previous half-edge (mesh_util.h:135-185). With b the boundary side of the
    // edge, c = its boundary predecessor u -> from(b), f the face normal across b, g the one across c:
    auto signed_cot = [](d3 v0, d3 v1, d3 fn) {
        const d3 cr = cross3(v0, v1);
        const double sn = sqrt(dot3(cr, cr));
        const double c = sn > 0.0 ? dot3(v0, v1) / sn : 0.0;
        return dot3(cr, fn) >= 0.0 ? c : -c;
    };
    auto loop_prev = [&](int b) { int q = b; while (mv.he_next[q] != b) q = mv.he_next[q]; return q; };
    if (mv.he_face[o] < 0 && mv.he_face[h] >= 0) {
        // b = o (j -> i): sector of o = (i -> m, i -> j) -> a02; sector of c = (u -> j): (j -> i, j -> u) -> a04
        cot02 = signed_cot(xm - xi, xj - xi, n0);
        const int c = loop_prev(o), u = mv.he_to[c ^ 1];
        const d3 xu = ldpos(mv.pos, u);
        const int cf = c ^ 1;                                           // j -> u, has the face (j, u, q)
        const d3 xq = ldpos(mv.pos, mv.he_to[mv.he_next[cf]]);
        cot04 = signed_cot(xi - xj, xu - xj, cross3(xq - xu, xj - xu));
    } else if (mv.he_face[h] < 0 && mv.he_face[o] >= 0) {
        // b = h (i -> j): sector of h = (j -> k, j -> i) -> a03; sector of c = (u -> i): (i -> j, i -> u) -> a01
        cot03 = signed_cot(xk - xj, xi - xj, n1);
        const int c = loop_prev(h), u = mv.he_to[c ^ 1];
        const d3 xu = ldpos(mv.pos, u);
        const int cf = c ^ 1;                                           // i -> u, has the face (i, u, q)
        const d3 xq = ldpos(mv.pos, mv.he_to[mv.he_next[cf]]);
        cot01 = signed_cot(xj - xi, xu - xi, cross3(xq - xu, xi - xu));
    }
    const double il = 1.0 / l;
    st3(out + 28, il * (cot03 * u0 + cot04 * u1));
    st3(out + 31, il * (cot01 * u0 + cot02 * u1));
    st3(out + 34, (-(cot01 + cot03) * il) * u0);
    st3(out + 37, (-(cot02 + cot04) * il) * u1);
}

}  // namespace

extern "C" {

int tm_area_edges(tm_mesh* m, double* area, double* edges_length) {
    TM_CHECK(m != nullptr && m->V > 0 && m->he_to_h, "mesh has no topology (call tm_mesh_set_topology first)");
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_ensure_layout(m)) return 1;
    const int nb = 296;
    double *part = nullptr, *out = nullptr;
    TM_CUDA(cudaMalloc(&part, sizeof(double) * 2 * nb));
    TM_CUDA(cudaMalloc(&out, sizeof(double) * 2));
    k_area_edges_partial<<<nb, NT, 0, m->stream>>>(m->view(), part); TM_LAUNCHED();
    k_area_edges_final<<<1, NT, 0, m->stream>>>(part, nb, out); TM_LAUNCHED();
    double h[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(h, out, sizeof(h), cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    cudaFree(part); cudaFree(out);
    TM_CHECK(e == cudaSuccess, "tm_area_edges failed: %s", cudaGetErrorString(e));
    if (area) *area = h[0];
    if (edges_length) *edges_length = h[1];
    return 0;
}

int tm_halfedge_geometry(tm_mesh* m, int halfedge, double out[40]) {
    TM_CHECK(m != nullptr && m->V > 0 && m->he_to_h, "mesh has no topology (call tm_mesh_set_topology first)");
    TM_CHECK(halfedge >= 0 && halfedge < m->HE, "half-edge index %d out of range", halfedge);
    TM_CUDA(cudaSetDevice(m->device));
    if (tm_ensure_layout(m)) return 1;
    double* d = nullptr;
    TM_CUDA(cudaMalloc(&d, sizeof(double) * 40));
    k_halfedge_geometry<<<1, 32, 0, m->stream>>>(m->view(), halfedge, d); TM_LAUNCHED();
    cudaError_t e = cudaMemcpyAsync(out, d, sizeof(double) * 40, cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    cudaFree(d);
    TM_CHECK(e == cudaSuccess, "tm_halfedge_geometry failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
