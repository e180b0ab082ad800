// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for OpenMesh handle types.
#pragma once
namespace OpenMesh {
struct BaseHandle {
    int idx_;
    explicit BaseHandle(int i = -1) : idx_(i) {}
    int idx() const { return idx_; }
    bool is_valid() const { return idx_ >= 0; }
    bool operator==(const BaseHandle& o) const { return idx_ == o.idx_; }
    bool operator!=(const BaseHandle& o) const { return idx_ != o.idx_; }
};
struct VertexHandle : BaseHandle { explicit VertexHandle(int i = -1) : BaseHandle(i) {} };
struct HalfedgeHandle : BaseHandle { explicit HalfedgeHandle(int i = -1) : BaseHandle(i) {} };
struct EdgeHandle : BaseHandle { explicit EdgeHandle(int i = -1) : BaseHandle(i) {} };
struct FaceHandle : BaseHandle { explicit FaceHandle(int i = -1) : BaseHandle(i) {} };
}  // namespace OpenMesh
