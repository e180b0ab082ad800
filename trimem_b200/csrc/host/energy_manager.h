// Host-side operator interface of the hot path: parameter PODs, EnergyManager,
// NeighbourList and the flip-sweep drivers. Same names, argument meaning and
// error behaviour as the reference's C++ core (src/params.h, src/energy.h,
// src/nlists/nlist.h, src/flips.h); every evaluation is forwarded to the C ABI
// (include/trimem_b200.h) and runs on the device.
#pragma once
#include <array>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../../include/trimem_b200.h"
#include "halfedge_mesh.h"

namespace trimem_b200 {

// ---- parameters: defaults as in src/params.h:19-71 ---------------------------
enum class BondType : int { Edge, Area, None, Default = Edge };

struct BondParams {
    double lc0 = 1;
    double lc1 = 0;
    double a0 = 1;
    int r = 2;
    BondType type = BondType::Edge;
};

struct ContinuationParams {
    double delta = 0;
    double lambda = 1;
};

struct SurfaceRepulsionParams {
    double lc1 = 0;
    int r = 2;
    std::string n_search = "cell-list";
    double rlist = 0.1;
    int exclusion_level = 2;
};

struct EnergyParams {
    double kappa_b = 0, kappa_a = 0, kappa_v = 0, kappa_c = 0, kappa_t = 0, kappa_r = 0;
    double area_frac = 1, volume_frac = 1, curvature_frac = 1;
    BondParams bond_params;
    ContinuationParams continuation_params;
    SurfaceRepulsionParams repulse_params;
};

// reduced (or per-vertex) properties, src/mesh_properties.h:15-48
struct VertexProperties {
    double area = 0, volume = 0, curvature = 0, bending = 0, tethering = 0, repulsion = 0;
};

// ---- neighbour list handle (src/nlists/nlist.h:17-41) ---------------------------
class NeighbourList {
public:
    NeighbourList(HalfedgeMesh& mesh, const EnergyParams& params);   // = make_nlist
    ~NeighbourList();
    NeighbourList(const NeighbourList&) = delete;
    NeighbourList& operator=(const NeighbourList&) = delete;

    std::tuple<std::vector<double>, std::vector<int>, std::vector<int>>
    distance_matrix(HalfedgeMesh& mesh, double rdist) const;
    // (n x 3 distance components x_pid - x_j, neighbour ids)
    std::tuple<std::vector<std::array<double, 3>>, std::vector<int>>
    point_distances(HalfedgeMesh& mesh, int pid, double rdist) const;

    tm_nlist* handle() const { return nl_; }

private:
    tm_nlist* nl_ = nullptr;
};

std::unique_ptr<NeighbourList> make_nlist(HalfedgeMesh& mesh, const EnergyParams& params);

// ---- EnergyManager (src/energy.h:20-56) ---------------------------------------------
class EnergyManager {
public:
    EnergyManager(HalfedgeMesh& mesh, const EnergyParams& params);

    void update_reference_properties();
    VertexProperties interpolate_reference_properties() const;
    void update_repulsion(HalfedgeMesh& mesh);
    // extension: has any vertex moved far enough since the last update_repulsion that the verlet rows may
    // miss a pair inside the repulsion cut-off? (always true for cell lists, which are cheap to keep current)
    bool repulsion_outdated(HalfedgeMesh& mesh);

    VertexProperties properties(HalfedgeMesh& mesh);
    double energy(HalfedgeMesh& mesh);
    double energy(const VertexProperties& props) const;
    // gradient into caller memory (n_vertices x 3)
    void gradient(HalfedgeMesh& mesh, double* out);
    std::string info(HalfedgeMesh& mesh);

    // C-ABI parameter block for the current state (reference values interpolated)
    tm_energy_params abi_params() const;
    tm_nlist* nlist_handle() const { return nlist ? nlist->handle() : nullptr; }

    EnergyParams params;
    VertexProperties initial_props;
    std::unique_ptr<NeighbourList> nlist;
};

// ---- flips (src/flips.h) ----------------------------------------------------------------
// `seed` < 0 seeds from the clock like the reference (flips.cpp:20,105-106).
int flip_serial(HalfedgeMesh& mesh, EnergyManager& estore, double flip_ratio, long long seed = -1);
int flip_parallel_batches(HalfedgeMesh& mesh, EnergyManager& estore, double flip_ratio, long long seed = -1);
// extension: the same sweeps with the caller's proposals and acceptance variates (one per proposal) instead of the
// internal generator's: reproducible sweeps for tests and restarts
int flip_with(HalfedgeMesh& mesh, EnergyManager& estore, const std::vector<int>& edges, const std::vector<double>& u,
              bool parallel);

// ---- velocity-Verlet trajectory (src/trimem/mc/hmc.py:14-26, device resident) ------------
// n_steps of _vv_integration from the mesh's current points() and the momenta p0 (n_vertices x 3);
// x_out / p_out (n_vertices x 3) receive the end point, stats (optional) = {U(x0), K(p0), U(xN), K(pN)}.
// The mesh's points() are left untouched, like the reference's evaluator-driven integration leaves
// the decision to the Metropolis test.
void vv_integrate(HalfedgeMesh& mesh, EnergyManager& estore, const double* p0, double mass, double dt,
                  int n_steps, double* x_out, double* p_out, double* stats);

// extension: the same with the momenta drawn on the device, scale * N(0, 1) from (seed, sequence) (tm_mesh_random_momenta);
// the end point's momenta stay on the device (the Metropolis test needs their kinetic energy only: stats[3])
void vv_integrate_random(HalfedgeMesh& mesh, EnergyManager& estore, unsigned long long seed, unsigned long long sequence,
                         double scale, double mass, double dt, int n_steps, double* x_out, double* stats);
// extension (rebuild trigger): (largest vertex displacement of the last trajectory's end point from the
// positions the verlet list was built from, margin = rlist - lc1), measured on the device behind the trajectory
std::tuple<double, double> vv_displacement(HalfedgeMesh& mesh, EnergyManager& estore);

// ---- setup-time reductions (src/util.h) ------------------------------------------------
double area(HalfedgeMesh& mesh);
double edges_length(HalfedgeMesh& mesh);
std::tuple<double, double> mean_tri_props(HalfedgeMesh& mesh);

}  // namespace trimem_b200
