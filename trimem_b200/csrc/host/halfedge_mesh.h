// Host half-edge mirror of the device mesh.
//
// The reference's TriMesh is OpenMesh::TriMesh_ArrayKernelT<fp64 traits>
// (src/mesh.h:15-31). Python-visible handles, fv_indices() and the edge ids used
// by the flip sweeps all follow OpenMesh's ArrayKernel numbering, so this mirror
// reproduces that numbering exactly:
//   * edge k <-> half-edges 2k (towards b) and 2k+1 (towards a) for the first
//     face that introduced edge (a,b); opposite(h) = h ^ 1;
//   * faces keep their input order, halfedge(face) = v[2] -> v[0];
//   * vertex circulation is clockwise, h -> next(opposite(h)), starting at the
//     vertex's stored half-edge, which is a boundary half-edge if one exists;
//   * flip(e) keeps edge and half-edge ids (TriConnectivity::flip).
// Positions live in (pinned, when a device is present) host memory that Python
// aliases through points(); the flattened arrays are what tm_mesh_set_topology
// uploads.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct tm_mesh;

namespace trimem_b200 {

// CUDA device new meshes attach to (one process per GPU: set it from LOCAL_RANK)
void set_default_device(int device);
int default_device();

class HalfedgeMesh {
public:
    HalfedgeMesh() = default;
    ~HalfedgeMesh();
    HalfedgeMesh(const HalfedgeMesh&) = delete;
    HalfedgeMesh& operator=(const HalfedgeMesh&) = delete;

    // construction (reference: from_points_cells, src/mesh_py.cpp:12-70)
    void set_points(const double* xyz, int n);
    // returns face id or -1 when OpenMesh would refuse the face (complex vertex/edge)
    int add_face(int a, int b, int c);

    int n_vertices() const { return nv_; }
    int n_halfedges() const { return (int)he_to_.size(); }
    int n_edges() const { return (int)he_to_.size() / 2; }
    int n_faces() const { return (int)f_he_.size(); }

    double* points() { resident_ = false; return xyz_; }      // (writable: the compute entries upload and compare again)
    const double* points() const { return xyz_; }

    // connectivity queries
    int to_vertex(int h) const { return he_to_[h]; }
    int from_vertex(int h) const { return he_to_[h ^ 1]; }
    int next(int h) const { return he_next_[h]; }
    int prev(int h) const { return he_prev_[h]; }
    int face(int h) const { return he_face_[h]; }
    int opposite(int h) const { return h ^ 1; }
    int vertex_halfedge(int v) const { return v_he_[v]; }
    int face_halfedge(int f) const { return f_he_[f]; }
    bool is_boundary_halfedge(int h) const { return he_face_[h] < 0; }
    bool is_boundary_edge(int e) const { return he_face_[2 * e] < 0 || he_face_[2 * e + 1] < 0; }
    bool is_boundary_vertex(int v) const { int h = v_he_[v]; return !(h >= 0 && he_face_[h] >= 0); }
    int find_halfedge(int a, int b) const;

    // TriConnectivity::is_flip_ok / flip
    bool is_flip_ok(int e) const;
    void flip(int e);

    const std::vector<int>& he_to() const { return he_to_; }
    const std::vector<int>& he_next() const { return he_next_; }
    const std::vector<int>& he_face() const { return he_face_; }
    const std::vector<int>& v_he() const { return v_he_; }
    const std::vector<int>& f_he() const { return f_he_; }

    // device mirror: created lazily, connectivity re-uploaded when it changed on the host,
    // positions re-uploaded on every call (Python writes through points() behind our back).
    tm_mesh* device(bool upload_positions = true);
    // Opt-in extension (TriMesh.set_points): copy `xyz` (n_vertices x 3) into the mesh's memory AND upload it, once.
    // From then on the compute entries skip their own re-upload -- until a writable view is handed out again
    // (the non-const points()) or the device positions move on their own (a trajectory: points_exposed()).
    void update_points(const double* xyz);
    void points_exposed() { resident_ = false; }
    bool points_resident() const { return resident_; }
    uint64_t topology_version() const { return topo_version_; }
    // the device connectivity was mutated by a flip sweep and the host replayed the same
    // flips: mark both sides as in sync without a re-upload
    void mark_device_in_sync() { dev_topo_version_ = topo_version_; }
    void invalidate_device_topology() { dev_topo_version_ = 0; }

private:
    void set_next(int h, int n) { he_next_[h] = n; he_prev_[n] = h; }
    void adjust_outgoing_halfedge(int v);

    int nv_ = 0;
    double* xyz_ = nullptr;
    std::vector<int> he_to_, he_next_, he_prev_, he_face_, v_he_, f_he_;
    uint64_t topo_version_ = 1;
    tm_mesh* dev_ = nullptr;
    uint64_t dev_topo_version_ = 0;
    bool resident_ = false;       // the device holds exactly xyz_ and nobody can have written to xyz_ since
};

// OpenMesh::IO::read_mesh for the formats trimem's drivers use (STL, OFF, OBJ); mesh_io.cpp
bool read_mesh_file(const std::string& fname, HalfedgeMesh& mesh);

}  // namespace trimem_b200
