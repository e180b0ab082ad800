// Device-resident velocity-Verlet trajectory of hybrid Monte Carlo.
//
// Replaces, in the reference's Python driver, _vv_integration (src/trimem/mc/hmc.py:14-26)
// and the two Hamiltonian evaluations around it (hmc.py:133-155):
//
//     a = force(x)                              force = -EnergyManager.gradient
//     for i in range(N):
//         x  = x + (p*dt + 0.5*a*dt**2) / m
//         an = force(x)
//         p  = p + 0.5*(a + an)*dt
//         a  = an
//
// There every force() call crosses the pybind boundary (positions up, gradient down, ~10
// numpy temporaries of V*3 doubles per step). Here x, p and both gradients stay in HBM for the
// N steps; per step the device runs the gradient sweeps (tm_energy.cu) and ONE update kernel
// that applies the momentum update of step i and the position update of step i+1 together.
// The two gradients ping-pong between m->grad and m->grad_prev, so a = -g_old and an = -g_new
// are read where they were written. The update is evaluated in numpy's operation order with
// explicitly rounded multiplies/adds (no FMA contraction): given equal gradients it is
// bit-identical to the reference expression.
//
// Energies. HMC needs U(x0) and U(xN) (nlog_prob, hmc.py:129-131). U(x0) comes from the
// full property sweep that also feeds the first gradient; at xN the last gradient evaluation
// runs the full sweeps (tether + pair energies) instead of the geometry-only ones, so the
// trajectory costs N+1 sweep sets in total, none repeated. The kinetic terms 0.5 p.p/m are
// reduced on the device in a fixed order.
#include <utility>

#include "tm_common.cuh"

namespace {

constexpr int NT = 256;

template <bool KICK, bool DRIFT>
__global__ void __launch_bounds__(NT)
k_vv_update(SweepRange R, P4* pos, double* __restrict__ mom, const double* __restrict__ g_old,
            const double* __restrict__ g_new, double dt, double dt2, double mass) {
    const int t = blockIdx.x * NT + threadIdx.x;
    if (t >= R.count) return;
    const size_t i = (size_t)(R.first + t);
    double p[3], an[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        p[k] = mom[3 * i + k];
        an[k] = -g_new[3 * i + k];
    }
    if (KICK) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double a = -g_old[3 * i + k];
            // p + 0.5*(a + an)*dt
            p[k] = __dadd_rn(p[k], __dmul_rn(__dmul_rn(0.5, __dadd_rn(a, an[k])), dt));
            mom[3 * i + k] = p[k];
        }
    }
    if (DRIFT) {
        const d3 x = ldpos(pos, (int)i);
        double xn[3] = {x.x, x.y, x.z};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            // x + (p*dt + 0.5*a*dt**2)/m
            const double num = __dadd_rn(__dmul_rn(p[k], dt), __dmul_rn(__dmul_rn(0.5, an[k]), dt2));
            xn[k] = __dadd_rn(xn[k], __ddiv_rn(num, mass));
        }
        *reinterpret_cast<double2*>(pos + i) = make_double2(xn[0], xn[1]);
        pos[i].z = xn[2];
    }
}

// p.p over the swept vertices: fixed grid, fixed strides, fixed trees -> bit-reproducible
constexpr int DOT_BLOCKS = 296;
__global__ void __launch_bounds__(NT)
k_dot_partial(SweepRange R, const double* __restrict__ v, double* __restrict__ part) {
    double s = 0.0;
    for (int t = blockIdx.x * NT + threadIdx.x; t < R.count; t += DOT_BLOCKS * NT) {
        const double* q = v + 3 * (size_t)(R.first + t);
        s += fma(q[0], q[0], fma(q[1], q[1], q[2] * q[2]));
    }
    __shared__ double sm[NT / 32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < NT / 32; k++) tot += sm[k];
        part[blockIdx.x] = tot;
    }
}

__global__ void k_dot_final(const double* __restrict__ part, double scale, double* __restrict__ out) {
    // one warp: lanes stride over the block partials, then a fixed tree
    double s = 0.0;
    for (int b = threadIdx.x; b < DOT_BLOCKS; b += 32) s += part[b];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *out = scale * s;
}

// The energies a trajectory reports are collected in a device block and fetched once, after the
// trajectory: a device-to-host copy enqueued between the sweeps would sit in the copy engine's
// in-order queue, where it can hold back another rank's copy when several ranks share a device.
__global__ void k_store(double* __restrict__ dst, const double* __restrict__ src, int n) {
    if (threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}

// host-numbered (V,3) array <-> device-numbered
__global__ void k_import_vec3(int V, const double* __restrict__ in, const int* __restrict__ d2h,
                              double* __restrict__ out) {
    int d = blockIdx.x * NT + threadIdx.x;
    if (d >= V) return;
    const double* s = in + 3 * (size_t)d2h[d];
    double* o = out + 3 * (size_t)d;
    o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
}

// rows of the swept range only (a partitioned mesh exports what it owns)
__global__ void k_export_vec3_range(SweepRange R, const double* __restrict__ in, const int* __restrict__ d2h,
                                    double* __restrict__ out) {
    const int t = blockIdx.x * NT + threadIdx.x;
    if (t >= R.count) return;
    const int d = R.first + t;
    const double* s = in + 3 * (size_t)d;
    double* o = out + 3 * (size_t)d2h[d];
    o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
}

int ensure_state(tm_mesh* m) {
    const size_t n = 3 * (size_t)m->V;
    if (!m->mom) {
        TM_CUDA(cudaMalloc(&m->mom, n * sizeof(double)));
        TM_CUDA(cudaMemsetAsync(m->mom, 0, n * sizeof(double), m->stream));
    }
    if (!m->grad_prev) {
        TM_CUDA(cudaMalloc(&m->grad_prev, n * sizeof(double)));
        TM_CUDA(cudaMemsetAsync(m->grad_prev, 0, n * sizeof(double), m->stream));
    }
    if (!m->h_stats) TM_CUDA(cudaHostAlloc(&m->h_stats, 16 * sizeof(double), cudaHostAllocDefault));
    if (!m->dot_part) TM_CUDA(cudaMalloc(&m->dot_part, DOT_BLOCKS * sizeof(double)));
    if (!m->d_stats) TM_CUDA(cudaMalloc(&m->d_stats, 16 * sizeof(double)));
    return 0;
}

int kinetic(tm_mesh* m, double mass, double* out_dev) {
    double* part = m->dot_part;
    k_dot_partial<<<DOT_BLOCKS, NT, 0, m->stream>>>(tm_owned_range(m), m->mom, part); TM_LAUNCHED();
    k_dot_final<<<1, 32, 0, m->stream>>>(part, 0.5 / mass, out_dev); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

static int vv_run_direct(tm_mesh* m, const NlView& nl, const tm_energy_params& p, double mass, double dt, int n_steps,
                         bool with_energies) {
    if (ensure_state(m) || tm_part_refresh(m, nl)) return 1;
    const double dt2 = dt * dt;
    if (with_energies) {
        // U(x0): full sweeps; their sums, rho and pair force feed the first gradient
        if (tm_launch_props(m, nl, p, true, false) || tm_join_pairs(m)) return 1;      // (a split pair sweep stores sums[5])
        k_store<<<1, 32, 0, m->stream>>>(m->d_stats, m->sums, 6); TM_LAUNCHED();
        if (kinetic(m, mass, m->d_stats + 12)) return 1;
    }
    // a = force(x0). The ring gradient is always re-evaluated (it depends on every parameter);
    // the property sweep and the pair force are reused when they belong to x0.
    if (m->geo_epoch != m->pos_epoch && tm_launch_props(m, nl, p, false, false)) return 1;
    if (tm_launch_grad(m, nl, p)) return 1;
    const SweepRange R = tm_owned_range(m);
    const int nb = (R.count + NT - 1) / NT;
    if (n_steps > 0) {
        if (tm_part_before_update(m)) return 1;
        if (nb > 0) k_vv_update<false, true><<<nb, NT, 0, m->stream>>>(R, m->pos, m->mom, m->grad, m->grad, dt, dt2, mass);
        TM_LAUNCHED();
    }
    // Steps. On an unpartitioned mesh the update rides in the gradient sweep's epilogue: the sweep of step n
    // gathers x_n from `pos` and writes x_{n+1} into `pos_alt`, then the two arrays swap roles (a partitioned
    // mesh keeps ONE array, which its peers have mapped, and runs the separate update kernel).
    const bool fused = m->part == nullptr && !m->has_boundary;   // (open meshes: the boundary terms follow the sweep)
    for (int n = 1; n <= n_steps; n++) {
        m->pos_epoch++;
        if (tm_part_positions_updated(m)) return 1;
        std::swap(m->grad, m->grad_prev);          // a = -grad_prev, an = -grad (about to be evaluated)
        const bool last = n == n_steps;
        if (tm_launch_props(m, nl, p, last && with_energies, false)) return 1;
        if (fused) {
            VVFuse U;
            U.mode = last ? 1 : 2;
            U.mom = m->mom; U.g_prev = m->grad_prev; U.pos_next = m->pos_alt;
            U.dt = dt; U.dt2 = dt2; U.mass = mass;
            if (tm_launch_grad(m, nl, p, &U)) return 1;
            if (!last) std::swap(m->pos, m->pos_alt);
            continue;
        }
        if (tm_launch_grad(m, nl, p)) return 1;
        if (nb <= 0) {
            if (!last && tm_part_before_update(m)) return 1;
        } else if (last) {
            k_vv_update<true, false><<<nb, NT, 0, m->stream>>>(R, m->pos, m->mom, m->grad_prev, m->grad, dt, dt2, mass);
        } else {
            if (tm_part_before_update(m)) return 1;
            k_vv_update<true, true><<<nb, NT, 0, m->stream>>>(R, m->pos, m->mom, m->grad_prev, m->grad, dt, dt2, mass);
        }
        TM_LAUNCHED();
    }
    TM_CUDA(cudaGetLastError());
    if (with_energies) {
        if (n_steps == 0 && tm_launch_props(m, nl, p, true, false)) return 1;
        if (tm_join_pairs(m)) return 1;
        k_store<<<1, 32, 0, m->stream>>>(m->d_stats + 6, m->sums, 6); TM_LAUNCHED();
        if (kinetic(m, mass, m->d_stats + 13)) return 1;
        TM_CUDA(cudaGetLastError());
    }
    return 0;
}

// ---- trajectory as a CUDA graph ------------------------------------------------------------
// A trajectory is ~6 launches per step; on a 40,962-vertex membrane each kernel runs for 4-8 us, so
// a chain of them (and even more so eight replicas' chains issued by one host thread) is bound by
// launch cost, not by the device. The second time the same trajectory is requested (same step
// count, step size, mass, parameters, neighbour list, connectivity, cache state and gradient-buffer
// parity) the launch sequence is captured into a graph -- including the pair sweep's side stream --
// and from then on replayed with one cudaGraphLaunch; the host-side bookkeeping the launch functions
// do (epochs, buffer ping-pong) is replayed beside it. Partitioned meshes launch directly: their
// exchange kernels carry sequence numbers as arguments, and so do meshes whose pair sweep is cut across the side
// streams (>= 131,072 vertices with a verlet list: tm_vv_run). TM_VV_GRAPH=0 disables.
struct VVKey {
    int n_steps, with_e, geo_valid, frep_valid, V;
    double mass, dt;
    tm_energy_params p;
    unsigned long long nl_serial, topo_epoch, layout_epoch;
    const double* grad;
    const P4* pos;
};
struct VVGraph {
    VVKey key;
    cudaGraphExec_t exec = nullptr;
    int uses = 0;
    unsigned long long launches = 0, stamp = 0;
    bool frep_touched = false;
};
struct tm_vvcache {
    VVGraph g[4];
    unsigned long long clock = 0;
    bool disabled = false;
};

bool tm_pair_force_valid(const tm_mesh* m, const NlView& nl, const tm_energy_params& p);   // tm_energy.cu

static bool same_key(const VVKey& a, const VVKey& b) { return memcmp(&a, &b, sizeof(VVKey)) == 0; }

void tm_vv_release(tm_mesh* m) {
    if (!m->vvcache) return;
    for (auto& g : m->vvcache->g) if (g.exec) cudaGraphExecDestroy(g.exec);
    delete m->vvcache;
    m->vvcache = nullptr;
}

int tm_vv_run(tm_mesh* m, const NlView& nl, const tm_energy_params& p, double mass, double dt, int n_steps,
              bool with_energies) {
    static const bool enabled = [] { const char* e = getenv("TM_VV_GRAPH"); return !(e && e[0] == '0'); }();
    // Large meshes launch directly too: their kernels run for tens of microseconds, so the host stays ahead of the
    // device anyway, and a replayed graph loses what cutting the pair sweep across prioritised side streams gains
    // (10 steps at 655,362 vertices: 1.40 ms direct, 1.48 ms replayed; profiles/r02as).
    if (!enabled || m->part || n_steps < 1 || tm_pair_split_active(m, nl))
        return vv_run_direct(m, nl, p, mass, dt, n_steps, with_energies);
    if (!m->vvcache) m->vvcache = new tm_vvcache();
    tm_vvcache* C = m->vvcache;
    if (C->disabled) return vv_run_direct(m, nl, p, mass, dt, n_steps, with_energies);

    VVKey key;
    memset(&key, 0, sizeof(key));
    key.n_steps = n_steps; key.with_e = with_energies; key.V = m->V;
    key.geo_valid = m->geo_epoch == m->pos_epoch;
    key.frep_valid = tm_pair_force_valid(m, nl, p);
    key.mass = mass; key.dt = dt;
    memcpy(&key.p, &p, sizeof(p));
    key.p._pad = 0;
    key.nl_serial = nl.kind >= 0 ? nl.serial : 0;
    key.topo_epoch = m->topo_epoch; key.layout_epoch = m->layout_epoch;
    key.grad = m->grad;
    key.pos = m->pos;

    VVGraph* slot = nullptr;
    for (auto& g : C->g) if (g.uses > 0 && same_key(g.key, key)) slot = &g;
    if (!slot) {                                   // first request: run directly (allocates what is needed), remember
        slot = &C->g[0];
        for (auto& g : C->g) if (g.stamp < slot->stamp) slot = &g;
        if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
        memcpy(&slot->key, &key, sizeof(VVKey));          // (bytewise: same_key compares the padding too)
        slot->uses = 1; slot->stamp = ++C->clock;
        return vv_run_direct(m, nl, p, mass, dt, n_steps, with_energies);
    }
    slot->stamp = ++C->clock;
    slot->uses++;
    if (!slot->exec) {
        // second request: capture. The launch functions update host-side state while they enqueue;
        // snapshot it so that a failed capture can be undone and rerun directly.
        const tm_mesh snap = *m;
        const unsigned long long l0 = tl_tm_launches;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            const int rc = vv_run_direct(m, nl, p, mass, dt, n_steps, with_energies);
            const cudaError_t e = cudaStreamEndCapture(m->stream, &graph);
            ok = rc == 0 && e == cudaSuccess && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&slot->exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {
            cudaGetLastError();
            tm_vvcache* keep = m->vvcache;
            *m = snap;
            m->vvcache = keep;
            g_tm_launches.fetch_sub(tl_tm_launches - l0, std::memory_order_relaxed);   // nothing was launched
            tl_tm_launches = l0;
            C->disabled = true;
            slot->exec = nullptr;
            return vv_run_direct(m, nl, p, mass, dt, n_steps, with_energies);
        }
        slot->launches = tl_tm_launches - l0;
        slot->frep_touched = m->frep_epoch == m->pos_epoch;
        TM_CUDA(cudaGraphLaunch(slot->exec, m->stream));
        return 0;
    }
    // replay: one launch, plus the bookkeeping of vv_run_direct
    TM_CUDA(cudaGraphLaunch(slot->exec, m->stream));
    g_tm_launches.fetch_add(slot->launches, std::memory_order_relaxed);
    tl_tm_launches += slot->launches;
    m->pos_epoch += (unsigned long long)n_steps;
    m->geo_epoch = m->pos_epoch;
    m->grad_epoch = m->pos_epoch;
    if (slot->frep_touched) {
        m->frep_epoch = m->pos_epoch; m->frep_lc1 = p.rep_lc1; m->frep_r = p.rep_r; m->frep_list = nl.serial;
    }
    if (n_steps & 1) std::swap(m->grad, m->grad_prev);
    if (!m->has_boundary && ((n_steps - 1) & 1)) std::swap(m->pos, m->pos_alt);   // fused steps ping-pong the record arrays
    return 0;
}

// ---- momenta drawn on the device (extension) ---------------------------------------------------
// The reference draws an HMC move's momenta with numpy on the host (src/trimem/mc/hmc.py:135-137); at 40,962
// vertices that draw is two thirds of a move once the trajectory runs on the device (scripts/hmc_breakdown.py).
// Counter-based generator, so a chain is reproducible from (seed, move number) whatever the launch shape:
// Philox4x32-10 (Salmon et al., SC'11; constants as published) keyed by the seed, counter = (vertex id in HOST
// numbering, half, sequence lo, sequence hi); the four words of a block make two 53-bit uniforms in (0, 1) and
// Box-Muller turns them into two standard normals; a vertex takes three of its two blocks' four.
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                              unsigned (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u53(unsigned hi, unsigned lo) {          // (0, 1), 53 bits
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}
__global__ void k_random_momenta(int V, const int* __restrict__ d2h, unsigned long long seed, unsigned long long seq,
                                 double scale, double* __restrict__ mom) {
    const int d = blockIdx.x * NT + threadIdx.x;
    if (d >= V) return;
    const unsigned v = (unsigned)d2h[d];               // host numbering: the draw does not depend on the device order
    double z[4];
#pragma unroll
    for (int half = 0; half < 2; half++) {
        unsigned w[4];
        philox4x32_10(v, (unsigned)half, (unsigned)seq, (unsigned)(seq >> 32), (unsigned)seed, (unsigned)(seed >> 32), w);
        const double r = sqrt(-2.0 * log(u53(w[0], w[1])));
        double sn, cs;
        sincospi(2.0 * u53(w[2], w[3]), &sn, &cs);
        z[2 * half] = r * cs; z[2 * half + 1] = r * sn;
    }
    mom[3 * (size_t)d] = scale * z[0]; mom[3 * (size_t)d + 1] = scale * z[1]; mom[3 * (size_t)d + 2] = scale * z[2];
}

int tm_random_momenta(tm_mesh* m, unsigned long long seed, unsigned long long seq, double scale) {
    if (ensure_state(m)) return 1;
    k_random_momenta<<<(m->V + NT - 1) / NT, NT, 0, m->stream>>>(m->V, m->d2h, seed, seq, scale, m->mom); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    return 0;
}

int tm_import_momenta(tm_mesh* m, const double* p_host) {
    if (ensure_state(m)) return 1;
    // through the host-numbered staging block, on the mesh's stream (positions staged there have
    // been imported by the caller's tm_ensure_layout)
    TM_CHECK(!m->stage_pending, "internal: staged positions not imported");
    TM_CUDA(cudaMemcpyAsync(m->xio3, p_host, sizeof(double) * 3 * (size_t)m->V, cudaMemcpyHostToDevice, m->stream));
    k_import_vec3<<<(m->V + NT - 1) / NT, NT, 0, m->stream>>>(m->V, m->xio3, m->d2h, m->mom); TM_LAUNCHED();
    TM_CUDA(cudaGetLastError());
    return 0;
}

int tm_export_vec3_owned(tm_mesh* m, const double* dev, double* host) {
    const SweepRange R = tm_owned_range(m);
    if (R.count < m->V) TM_CUDA(cudaMemsetAsync(m->xout3, 0, sizeof(double) * 3 * (size_t)m->V, m->stream));
    if (R.count > 0) { k_export_vec3_range<<<(R.count + NT - 1) / NT, NT, 0, m->stream>>>(R, dev, m->d2h, m->xout3); TM_LAUNCHED(); }
    TM_CUDA(cudaGetLastError());
    if (tm_part_quiesce(m)) return 1;
    TM_CUDA(cudaMemcpyAsync(host, m->xout3, sizeof(double) * 3 * (size_t)m->V, cudaMemcpyDeviceToHost, m->stream));
    return 0;
}

int tm_preload_integrate() {
    TM_PRELOAD((k_vv_update<false, true>)); TM_PRELOAD((k_vv_update<true, false>)); TM_PRELOAD((k_vv_update<true, true>));
    TM_PRELOAD(k_dot_partial); TM_PRELOAD(k_dot_final); TM_PRELOAD(k_store);
    TM_PRELOAD(k_import_vec3); TM_PRELOAD(k_export_vec3_range);
    return 0;
}
