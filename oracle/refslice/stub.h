// oracle/refslice/stub.h — the minimal environment the SLICED reference sources compile in.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp).  The reference cannot be built here (it needs
// GCC >= 14 for deducing-this and <print>, plus Boost/Eigen/nanobind and external catalog data,
// DESIGN.md section 2), but the function BODIES that carry the hot path's arithmetic use only
// `Numeric`, `std::` math and three tiny value types.  oracle/slice_ref.py cuts those bodies out of
// /root/reference at build time (by line range, each range guarded by anchors on its first and last
// line) into oracle/_ref/refslice_gen.cpp; this header supplies what they name and nothing that
// computes: the constant-data base class of matpack (`cdata_t`: an array, element-wise compound
// assignment, the tuple protocol) and a row-major view.  The constants come from the reference's
// own util headers, which DO compile with g++ 13 and are included where they lie (-I).
//
// Nothing here restates reference arithmetic: every formula under test lives in the sliced text.
#pragma once

#include <omp.h>

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <numbers>
#include <numeric>
#include <span>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

// the reference's own headers, compiled in place (oracle/Makefile passes -I$(REF)/src/core/util)
#include <arts_constants.h>
#include <arts_constexpr_math.h>
#include <arts_conversions.h>
#include <configtypes.h>
#include <nonstd.h>

using Complex = std::complex<Numeric>;

inline bool arts_omp_in_parallel() { return omp_in_parallel(); }

namespace matpack {
// what matpack_mdspan_cdata_t.h provides to rtepack's value types: storage + element-wise compound
// assignment (:144-167) + structured bindings (:253-266, :348-356) + [i] / [i, j] access
template <class T, Size... dims>
struct cdata_t {
  static constexpr Size ndata = (dims * ...);
  static constexpr Size N     = sizeof...(dims);
  static constexpr std::array<Size, N> shape_{dims...};
  using value_type            = T;
  using refslice_cdata_tag    = void;

  std::array<T, ndata> data;

  constexpr T& operator[](Size i) requires(N == 1) { return data[i]; }
  constexpr const T& operator[](Size i) const requires(N == 1) { return data[i]; }
  constexpr T& operator[](Size i, Size j) requires(N == 2) { return data[i * shape_[1] + j]; }
  constexpr const T& operator[](Size i, Size j) const requires(N == 2) { return data[i * shape_[1] + j]; }

  constexpr cdata_t& operator+=(const cdata_t& x) {
    for (Size i = 0; i < ndata; i++) data[i] += x.data[i];
    return *this;
  }
  constexpr cdata_t& operator-=(const cdata_t& x) {
    for (Size i = 0; i < ndata; i++) data[i] -= x.data[i];
    return *this;
  }
  constexpr cdata_t& operator*=(const T& x) {
    for (Size i = 0; i < ndata; i++) data[i] *= x;
    return *this;
  }
  constexpr cdata_t& operator/=(const T& x) {
    for (Size i = 0; i < ndata; i++) data[i] /= x;
    return *this;
  }
  constexpr cdata_t& operator/=(const cdata_t& x) {
    for (Size i = 0; i < ndata; i++) data[i] /= x.data[i];
    return *this;
  }

  template <Index i>
  constexpr T& get() & { return std::get<i>(data); }
  template <Index i>
  constexpr const T& get() const& { return std::get<i>(data); }
  template <Index i>
  constexpr T&& get() && { return std::get<i>(std::move(data)); }
};

template <class T>
concept any_cdata = requires { typename std::remove_cvref_t<T>::refslice_cdata_tag; };

// scaling by a value of the element type: matpack_mdspan_cdata_t.h:301-319 (these templates, not an implicit conversion
// of the scalar to the class, are what `propmat / Numeric` resolves to: element-wise)
template <any_cdata T>
constexpr T operator*(T x, const std::convertible_to<typename T::value_type> auto& y) {
  x *= static_cast<typename T::value_type>(y);
  return x;
}
template <any_cdata T>
constexpr T operator*(const std::convertible_to<typename T::value_type> auto& y, T x) {
  x *= static_cast<typename T::value_type>(y);
  return x;
}
template <any_cdata T>
constexpr T operator/(T x, const std::convertible_to<typename T::value_type> auto& y) {
  x /= static_cast<typename T::value_type>(y);
  return x;
}

// row-major dense view: what the sliced loops of rtepack_rtestep.cc index with [i], [i, j] and npages()/nrows()/ncols()
template <class T, Size N>
struct view_t {
  T* p{};
  std::array<Size, N> ext{};

  constexpr view_t() = default;
  constexpr view_t(T* p_, std::array<Size, N> e) : p(p_), ext(e) {}
  template <class U>
    requires(std::is_same_v<const U, T>)
  constexpr view_t(const view_t<U, N>& o) : p(o.p), ext(o.ext) {}

  [[nodiscard]] constexpr Size stride0() const {
    Size s = 1;
    for (Size k = 1; k < N; k++) s *= ext[k];
    return s;
  }
  constexpr decltype(auto) operator[](Size i) const {
    if constexpr (N == 1) {
      return (p[i]);
    } else {
      std::array<Size, N - 1> e{};
      for (Size k = 1; k < N; k++) e[k - 1] = ext[k];
      return view_t<T, N - 1>{p + i * stride0(), e};
    }
  }
  constexpr T& operator[](Size i, Size j) const requires(N == 2) { return p[i * ext[1] + j]; }
  constexpr T& operator[](Size i, Size j, Size k) const requires(N == 3) { return p[(i * ext[1] + j) * ext[2] + k]; }
  [[nodiscard]] constexpr Size size() const {
    Size s = 1;
    for (Size k = 0; k < N; k++) s *= ext[k];
    return s;
  }
  [[nodiscard]] constexpr T* begin() const requires(N == 1) { return p; }
  [[nodiscard]] constexpr T* end() const requires(N == 1) { return p + ext[0]; }
  [[nodiscard]] constexpr Size ncols() const { return ext[N - 1]; }
  [[nodiscard]] constexpr Size nrows() const requires(N >= 2) { return ext[N - 2]; }
  [[nodiscard]] constexpr Size npages() const requires(N >= 3) { return ext[N - 3]; }
};
}  // namespace matpack

namespace std {
template <matpack::any_cdata T>
struct tuple_size<T> : std::integral_constant<std::size_t, std::remove_cvref_t<T>::ndata> {};
template <std::size_t I, matpack::any_cdata T>
struct tuple_element<I, T> {
  using type = typename std::remove_cvref_t<T>::value_type;
};
}  // namespace std

using Vector4  = matpack::cdata_t<Numeric, 4>;
using Vector7  = matpack::cdata_t<Numeric, 7>;
using Matrix44 = matpack::cdata_t<Numeric, 4, 4>;
using ComplexMatrix44 = matpack::cdata_t<Complex, 4, 4>;
using ConstVectorView = matpack::view_t<const Numeric, 1>;

// rtepack_common.h forward-declares these; rtepack_stokes_vector.h names the two enums only in functions that are not sliced
namespace rtepack {
struct propmat;
struct muelmat;
struct stokvec;
struct specmat;
}  // namespace rtepack
