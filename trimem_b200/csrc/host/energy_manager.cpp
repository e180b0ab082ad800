#include "energy_manager.h"

#include <chrono>
#include <cmath>
#include <random>
#include <sstream>
#include <stdexcept>

namespace trimem_b200 {

namespace {
[[noreturn]] void fail() { throw std::runtime_error(tm_last_error()); }

int nlist_kind(const EnergyParams& p) {
    // make_nlist dispatch and its error messages (src/nlists/nlist.cpp:10-53)
    const auto& s = p.repulse_params.n_search;
    int kind;
    if (s == "cell-list") kind = TM_NLIST_CELL;
    else if (s == "verlet-list") kind = TM_NLIST_VERLET;
    else throw std::runtime_error("Unknown neighbour search algorithm.");
    int e = p.repulse_params.exclusion_level;
    if (e < 0 || e > 2) throw std::runtime_error("Unsupported exclusion level");
    return kind;
}

VertexProperties from6(const double s[6]) {
    VertexProperties p;
    p.area = s[0]; p.volume = s[1]; p.curvature = s[2];
    p.bending = s[3]; p.tethering = s[4]; p.repulsion = s[5];
    return p;
}
}  // namespace

// ------------------------------------------------------------- NeighbourList --
NeighbourList::NeighbourList(HalfedgeMesh& mesh, const EnergyParams& params) {
    int kind = nlist_kind(params);
    tm_mesh* dm = mesh.device();
    // box_eps default of the reference constructors (cell_list.h:43, verlet_list.h:18)
    if (tm_nlist_create(dm, kind, params.repulse_params.rlist, params.repulse_params.exclusion_level,
                        1.0e-6, &nl_))
        fail();
}

NeighbourList::~NeighbourList() { tm_nlist_destroy(nl_); }

std::tuple<std::vector<double>, std::vector<int>, std::vector<int>>
NeighbourList::distance_matrix(HalfedgeMesh& mesh, double rdist) const {
    long long n = 0;
    double* d = nullptr;
    int *i = nullptr, *j = nullptr;
    if (tm_nlist_distance_matrix(nl_, mesh.device(), rdist, &n, &d, &i, &j)) fail();
    std::tuple<std::vector<double>, std::vector<int>, std::vector<int>> out(
        std::vector<double>(d, d + n), std::vector<int>(i, i + n), std::vector<int>(j, j + n));
    tm_free(d); tm_free(i); tm_free(j);
    return out;
}

std::tuple<std::vector<std::array<double, 3>>, std::vector<int>>
NeighbourList::point_distances(HalfedgeMesh& mesh, int pid, double rdist) const {
    int n = 0;
    double* d = nullptr;
    int* j = nullptr;
    if (tm_nlist_point_distances(nl_, mesh.device(), pid, rdist, &n, &d, &j)) fail();
    std::vector<std::array<double, 3>> dv(n);
    for (int k = 0; k < n; k++) dv[k] = {d[3 * k], d[3 * k + 1], d[3 * k + 2]};
    std::vector<int> jv(j, j + n);
    tm_free(d); tm_free(j);
    return {std::move(dv), std::move(jv)};
}

std::unique_ptr<NeighbourList> make_nlist(HalfedgeMesh& mesh, const EnergyParams& params) {
    return std::make_unique<NeighbourList>(mesh, params);
}

// -------------------------------------------------------------- EnergyManager --
EnergyManager::EnergyManager(HalfedgeMesh& mesh, const EnergyParams& energy_params)
    : params(energy_params) {
    // same order as src/energy.cpp:13-28: bonds, neighbour list, repulsion, initial properties
    if (params.bond_params.type != BondType::Edge && params.bond_params.type != BondType::Area)
        throw std::runtime_error("Unknown bond potential");
    nlist = make_nlist(mesh, params);
    initial_props = properties(mesh);
}

VertexProperties EnergyManager::interpolate_reference_properties() const {
    // src/energy.cpp:30-45
    const double lam = params.continuation_params.lambda;
    VertexProperties ref;
    ref.area = (1.0 - lam * (1.0 - params.area_frac)) * initial_props.area;
    ref.volume = (1.0 - lam * (1.0 - params.volume_frac)) * initial_props.volume;
    ref.curvature = (1.0 - lam * (1.0 - params.curvature_frac)) * initial_props.curvature;
    return ref;
}

void EnergyManager::update_reference_properties() {
    auto& c = params.continuation_params;
    if (c.lambda < 1.0) c.lambda += c.delta;
}

void EnergyManager::update_repulsion(HalfedgeMesh& mesh) {
    nlist.reset();
    nlist = make_nlist(mesh, params);
}

bool EnergyManager::repulsion_outdated(HalfedgeMesh& mesh) {
    if (params.repulse_params.n_search != "verlet-list") return true;
    int need = 1;
    if (tm_nlist_displacement(nlist_handle(), mesh.device(), params.repulse_params.lc1, nullptr, &need)) fail();
    return need != 0;
}

tm_energy_params EnergyManager::abi_params() const {
    if (params.bond_params.type != BondType::Edge && params.bond_params.type != BondType::Area)
        throw std::runtime_error("Unknown bond potential");
    VertexProperties ref = interpolate_reference_properties();
    tm_energy_params p{};
    p.kappa_b = params.kappa_b; p.kappa_a = params.kappa_a; p.kappa_v = params.kappa_v;
    p.kappa_c = params.kappa_c; p.kappa_t = params.kappa_t; p.kappa_r = params.kappa_r;
    p.ref_area = ref.area; p.ref_volume = ref.volume; p.ref_curvature = ref.curvature;
    p.bond_lc0 = params.bond_params.lc0; p.bond_lc1 = params.bond_params.lc1;
    p.bond_a0 = params.bond_params.a0; p.bond_r = params.bond_params.r;
    p.bond_type = params.bond_params.type == BondType::Edge ? TM_BOND_EDGE : TM_BOND_AREA;
    p.rep_lc1 = params.repulse_params.lc1; p.rep_r = params.repulse_params.r;
    return p;
}

VertexProperties EnergyManager::properties(HalfedgeMesh& mesh) {
    tm_energy_params p = abi_params();
    double s[6];
    if (tm_properties(mesh.device(), nlist_handle(), &p, s)) fail();
    return from6(s);
}

double EnergyManager::energy(HalfedgeMesh& mesh) {
    tm_energy_params p = abi_params();
    double e = 0;
    if (tm_energy(mesh.device(), nlist_handle(), &p, &e, nullptr)) fail();
    return e;
}

double EnergyManager::energy(const VertexProperties& s) const {
    // trimem_energy (src/kernel.h:110-122) on given reduced properties: six flops, evaluated
    // where the properties live (they were reduced on the device)
    VertexProperties ref = interpolate_reference_properties();
    double e = 0.0, d;
    d = s.area / ref.area - 1.0;           e += params.kappa_a * d * d;
    d = s.volume / ref.volume - 1.0;       e += params.kappa_v * d * d;
    d = s.curvature / ref.curvature - 1.0; e += params.kappa_c * d * d;
    e += params.kappa_t * s.tethering;
    e += params.kappa_r * s.repulsion;
    e += params.kappa_b * s.bending;
    return e;
}

void EnergyManager::gradient(HalfedgeMesh& mesh, double* out) {
    tm_energy_params p = abi_params();
    if (tm_gradient(mesh.device(), nlist_handle(), &p, out, nullptr)) fail();
}

std::string EnergyManager::info(HalfedgeMesh& mesh) {
    // layout of EnergyManager::print_info (src/energy.cpp:126-151)
    VertexProperties s = properties(mesh);
    VertexProperties ref = interpolate_reference_properties();
    auto sq = [](double x) { return x * x; };
    std::ostringstream out;
    out << "----- EnergyManager info\n";
    out << "reference properties:\n";
    out << "  area:      " << ref.area << "\n";
    out << "  volume:    " << ref.volume << "\n";
    out << "  curvature: " << ref.curvature << "\n";
    out << "current properties:\n";
    out << "  area:      " << s.area << "\n";
    out << "  volume:    " << s.volume << "\n";
    out << "  curvature: " << s.curvature << "\n";
    out << "energies:\n";
    out << "  area:      " << params.kappa_a * sq(s.area / ref.area - 1.0) << "\n";
    out << "  volume:    " << params.kappa_v * sq(s.volume / ref.volume - 1.0) << "\n";
    out << "  area diff: " << params.kappa_c * sq(s.curvature / ref.curvature - 1.0) << "\n";
    out << "  bending:   " << params.kappa_b * s.bending << "\n";
    out << "  tether:    " << params.kappa_t * s.tethering << "\n";
    out << "  repulsion: " << params.kappa_r * s.repulsion << "\n";
    out << "  total:     " << energy(s) << "\n";
    out << "\n";
    return out.str();
}

// ----------------------------------------------------------------------- flips --
namespace {
int flip_sweep_with(HalfedgeMesh& mesh, EnergyManager& estore, const std::vector<int>& edges, const std::vector<double>& u,
                    int mode) {
    const int nflips = (int)edges.size();
    if (nflips <= 0) return 0;
    tm_energy_params p = estore.abi_params();
    tm_mesh* dm = mesh.device();
    double props[6];
    if (tm_properties(dm, estore.nlist_handle(), &p, props)) fail();
    std::vector<int> flipped(2 * (size_t)nflips);
    int nacc = 0, napplied = 0;
    if (tm_flip_sweep(dm, estore.nlist_handle(), &p, mode, nflips, edges.data(), u.data(), &nacc,
                      &napplied, flipped.data(), props)) {
        // the device may hold a half-finished sweep: the next device() call uploads this mesh's (pre-sweep)
        // connectivity again, so host and device agree and the failed sweep is rolled back as a whole
        mesh.invalidate_device_topology();
        fail();
    }
    // replay every applied flip (trial flips and flip-backs) on the host mirror, in order
    for (int k = 0; k < napplied; k++) mesh.flip(flipped[k]);
    mesh.mark_device_in_sync();
    return nacc;
}

int flip_sweep(HalfedgeMesh& mesh, EnergyManager& estore, double flip_ratio, long long seed, int mode) {
    const int nedges = mesh.n_edges();
    const int nflips = (int)(nedges * flip_ratio);
    if (nflips <= 0 || nedges <= 0) return 0;
    unsigned long long s = seed >= 0 ? (unsigned long long)seed
        : (unsigned long long)std::chrono::high_resolution_clock::now().time_since_epoch().count();
    std::mt19937 gen((std::mt19937::result_type)s);
    std::uniform_int_distribution<int> propose(0, nedges - 1);
    std::uniform_real_distribution<double> accept(0.0, 1.0);
    std::vector<int> edges(nflips);
    std::vector<double> u(nflips);
    for (int i = 0; i < nflips; i++) { edges[i] = propose(gen); u[i] = accept(gen); }
    return flip_sweep_with(mesh, estore, edges, u, mode);
}
}  // namespace

int flip_with(HalfedgeMesh& mesh, EnergyManager& estore, const std::vector<int>& edges, const std::vector<double>& u,
              bool parallel) {
    if (edges.size() != u.size()) throw std::invalid_argument("flip_with: one acceptance variate per proposed edge");
    for (int e : edges)
        if (e < 0 || e >= mesh.n_edges()) throw std::out_of_range("flip_with: edge id out of range");
    return flip_sweep_with(mesh, estore, edges, u, parallel ? TM_FLIP_BATCH : TM_FLIP_SERIAL);
}

int flip_serial(HalfedgeMesh& mesh, EnergyManager& estore, double flip_ratio, long long seed) {
    // the reference constructs a runtime_error for flip_ratio > 1 but never throws it
    // (src/flips.cpp:24-25): proposals are simply int(n_edges * ratio)
    return flip_sweep(mesh, estore, flip_ratio, seed, TM_FLIP_SERIAL);
}

int flip_parallel_batches(HalfedgeMesh& mesh, EnergyManager& estore, double flip_ratio, long long seed) {
    if (flip_ratio > 1.0) throw std::range_error("flip_ratio must be <= 1.0");   // src/flips.cpp:84-85
    return flip_sweep(mesh, estore, flip_ratio, seed, TM_FLIP_BATCH);
}

// ------------------------------------------------------------------ trajectory --
void vv_integrate(HalfedgeMesh& mesh, EnergyManager& estore, const double* p0, double mass, double dt,
                  int n_steps, double* x_out, double* p_out, double* stats) {
    tm_energy_params p = estore.abi_params();
    tm_mesh* dm = mesh.device();
    mesh.points_exposed();          // the trajectory moves the device positions away from the mesh's own
    if (tm_vv_integrate(dm, estore.nlist_handle(), &p, p0, mass, dt, n_steps, x_out, p_out, stats))
        fail();
}

void vv_integrate_random(HalfedgeMesh& mesh, EnergyManager& estore, unsigned long long seed, unsigned long long sequence,
                         double scale, double mass, double dt, int n_steps, double* x_out, double* stats) {
    tm_energy_params p = estore.abi_params();
    tm_mesh* dm = mesh.device();
    mesh.points_exposed();
    if (tm_mesh_random_momenta(dm, seed, sequence, scale)) fail();
    if (tm_vv_integrate(dm, estore.nlist_handle(), &p, nullptr, mass, dt, n_steps, x_out, nullptr, stats)) fail();
}

std::tuple<double, double> vv_displacement(HalfedgeMesh& mesh, EnergyManager& estore) {
    double disp = 0, margin = 0;
    if (tm_vv_displacement(mesh.device(), estore.nlist_handle(), estore.params.repulse_params.lc1, &disp, &margin)) fail();
    return {disp, margin};
}

// --------------------------------------------------------------- setup reductions --
double area(HalfedgeMesh& mesh) {
    double a = 0, l = 0;
    if (tm_area_edges(mesh.device(), &a, &l)) fail();
    return a;
}

double edges_length(HalfedgeMesh& mesh) {
    double a = 0, l = 0;
    if (tm_area_edges(mesh.device(), &a, &l)) fail();
    return l;
}

std::tuple<double, double> mean_tri_props(HalfedgeMesh& mesh) {
    double a = 0, l = 0;
    if (tm_area_edges(mesh.device(), &a, &l)) fail();
    return {a / mesh.n_faces(), l / mesh.n_edges()};
}

}  // namespace trimem_b200
