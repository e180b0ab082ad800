// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Minimal stand-in for OpenMesh's VectorT, exposing the subset of the API that
// the reference sources under /root/reference/src use. The
// OpenMesh submodule (pinned db8c0608830ae7324b2b2e1d91e4db565b2d3eae,
// /root/reference/.SUBMODULES.json:10) is absent from the reference tree.
#pragma once
#include <array>
#include <cmath>
#include <ostream>
#include <type_traits>

namespace OpenMesh {

template <typename S, int N>
struct VectorT {
    typedef S value_type;
    std::array<S, N> v_{};

    VectorT() {}
    explicit VectorT(const S& s) { v_.fill(s); }
    template <typename... T,
              typename = std::enable_if_t<sizeof...(T) == N && (N > 1)>>
    VectorT(T... t) : v_{{static_cast<S>(t)...}} {}

    S& operator[](size_t i) { return v_[i]; }
    const S& operator[](size_t i) const { return v_[i]; }
    S* data() { return v_.data(); }
    const S* data() const { return v_.data(); }
    static constexpr size_t size() { return N; }

    VectorT& operator+=(const VectorT& o) { for (int i = 0; i < N; i++) v_[i] += o.v_[i]; return *this; }
    VectorT& operator-=(const VectorT& o) { for (int i = 0; i < N; i++) v_[i] -= o.v_[i]; return *this; }
    template <typename T, typename = std::enable_if_t<std::is_arithmetic<T>::value>>
    VectorT& operator*=(const T& s) { for (int i = 0; i < N; i++) v_[i] *= s; return *this; }
    template <typename T, typename = std::enable_if_t<std::is_arithmetic<T>::value>>
    VectorT& operator/=(const T& s) { for (int i = 0; i < N; i++) v_[i] /= s; return *this; }

    VectorT operator+(const VectorT& o) const { VectorT r(*this); r += o; return r; }
    VectorT operator-(const VectorT& o) const { VectorT r(*this); r -= o; return r; }
    VectorT operator-() const { VectorT r; for (int i = 0; i < N; i++) r.v_[i] = -v_[i]; return r; }
    template <typename T, typename = std::enable_if_t<std::is_arithmetic<T>::value>>
    VectorT operator*(const T& s) const { VectorT r(*this); r *= s; return r; }
    template <typename T, typename = std::enable_if_t<std::is_arithmetic<T>::value>>
    VectorT operator/(const T& s) const { VectorT r(*this); r /= s; return r; }

    S dot(const VectorT& o) const { S r = 0; for (int i = 0; i < N; i++) r += v_[i] * o.v_[i]; return r; }
    S operator|(const VectorT& o) const { return dot(o); }
    S sqrnorm() const { return dot(*this); }
    S norm() const { return std::sqrt(sqrnorm()); }
    VectorT cross(const VectorT& o) const {
        static_assert(N == 3, "cross product needs 3 components");
        return VectorT(v_[1] * o.v_[2] - v_[2] * o.v_[1],
                       v_[2] * o.v_[0] - v_[0] * o.v_[2],
                       v_[0] * o.v_[1] - v_[1] * o.v_[0]);
    }
};

template <typename T, typename S, int N,
          typename = std::enable_if_t<std::is_arithmetic<T>::value>>
VectorT<S, N> operator*(const T& s, const VectorT<S, N>& v) { return v * s; }
template <typename S, int N> S norm(const VectorT<S, N>& v) { return v.norm(); }
template <typename S, int N> S dot(const VectorT<S, N>& a, const VectorT<S, N>& b) { return a.dot(b); }
template <typename S> VectorT<S, 3> cross(const VectorT<S, 3>& a, const VectorT<S, 3>& b) { return a.cross(b); }
template <typename S, int N>
std::ostream& operator<<(std::ostream& os, const VectorT<S, N>& v) {
    for (int i = 0; i < N; i++) { os << v[i]; if (i < N - 1) os << " "; }
    return os;
}

typedef VectorT<double, 3> Vec3d;
typedef VectorT<double, 2> Vec2d;
typedef VectorT<float, 4> Vec4f;
typedef VectorT<float, 3> Vec3f;

}  // namespace OpenMesh
