// atmpath.cu — atm_pathFromPath for a 1-D atmosphere: the AtmField flattened once, a path is np linear
// interpolations on flat arrays.  Replaces forward_atm_path (reference src/core/path/atm_path.cpp:19-28) ->
// Atm::Field::at (src/core/atm/atm_field.cpp:928-947) -> Atm::Data::at (:890-924) -> interp::get(GeodeticField3)
// (src/core/functional/functional_atm_field_interp.cpp:53-65, altitude lag of order 1 :6-10) for fields whose latitude and
// longitude grids have one point.  Host code: a path is O(np) numbers and ab200_path_upload consumes host arrays.
#include <cmath>
#include <string>

#include "common.cuh"

namespace {
struct Stencil {
  int64_t i0;      // first grid index of the two-point stencil
  double w0, w1;   // Lagrange weights (w1 = 1 - w0: the last weight closes the sum, lagrange_interp.h set_weights)
  bool zero;       // InterpolationExtrapolation::Zero applies
};

// Atm::find_limit + select (atm_field.cpp:536-566) on the altitude axis, then the order-1 lag (lagrange_interp.h find_pos /
// update_pos / set_weights for an ascending grid): the stencil starts at the last index whose right neighbour is not
// below x, clamped to [0, n - 2].
int stencil(const ab200_atm_profile& f, double alt, Stencil& s) {
  s = Stencil{0, 1.0, 0.0, false};
  const int64_t n = f.nalt;
  const double lo = f.alt[0], hi = f.alt[n - 1];
  int type = AB200_EXTRAP_LINEAR;
  if (alt < lo) {
    type = n == 1 ? AB200_EXTRAP_NEAREST : f.alt_low;  // adjust_interpolation_extrapolation :57-83
    if (type == AB200_EXTRAP_NEAREST) alt = lo;
  } else if (hi < alt) {
    type = n == 1 ? AB200_EXTRAP_NEAREST : f.alt_upp;
    if (type == AB200_EXTRAP_NEAREST) alt = hi;
  }
  if (type == AB200_EXTRAP_NONE)
    return ab200::set_error(AB200_ERR_INVALID, "Limit breached.  Position (" + std::to_string(alt) +
                                                   ", 0, 0) is out-of-bounds when no extrapolation is wanted");
  if (type == AB200_EXTRAP_ZERO) {
    s.zero = true;
    return AB200_OK;
  }
  if (n == 1) return AB200_OK;
  int64_t xp = 0;
  const int64_t xe = n - 2;
  // same walk as update_pos from the fractional-index guess, written as a bisection: the result is the unique index with
  // alt[xp] <= x (or xp == 0) and not alt[xp + 1] < x (or xp == n - 2)
  int64_t a = 0, b = xe;
  while (a < b) {
    const int64_t m = (a + b) / 2;
    if (f.alt[m + 1] < alt) a = m + 1;
    else b = m;
  }
  xp = a;
  s.i0 = xp;
  const double x0 = f.alt[xp], x1 = f.alt[xp + 1];
  s.w0 = (alt - x1) / (x0 - x1);
  s.w1 = 1.0 - s.w0;
  return AB200_OK;
}
inline double at(const Stencil& s, const double* v, int64_t stride, int64_t n) {
  if (s.zero) return 0.0;
  if (n == 1) return v[0];
  return s.w0 * v[s.i0 * stride] + s.w1 * v[(s.i0 + 1) * stride];
}
}  // namespace

int ab200_atm_path_from_profile(const ab200_atm_profile* f, int32_t n_species, int32_t n_isot, int32_t np, const double* alt,
                                const uint8_t* in_atm, double* T, double* P, double* vmr, double* isorat, double* Q, double* dQdT,
                                double* mag, double* wind) {
  using ab200::set_error;
  if (!f || !alt || !T || !P || (n_species > 0 && !vmr) || (n_isot > 0 && !isorat))
    return set_error(AB200_ERR_INVALID, "ab200_atm_path_from_profile: null argument");
  if (f->nalt < 1 || !f->alt || !f->T || !f->P || (n_species > 0 && !f->vmr) || (n_isot > 0 && !f->isorat))
    return set_error(AB200_ERR_INVALID, "ab200_atm_path_from_profile: the profile needs alt, T, P, vmr and isorat");
  for (int64_t i = 1; i < f->nalt; i++)
    if (!(f->alt[i - 1] < f->alt[i])) return set_error(AB200_ERR_INVALID, "ab200_atm_path_from_profile: the altitude grid must be ascending");
  for (int ip = 0; ip < np; ip++) {
    const double a = (in_atm && !in_atm[ip]) ? f->top_of_atmosphere : alt[ip];  // atm_path.cpp:24-26
    if (a > f->top_of_atmosphere)
      return set_error(AB200_ERR_INVALID, "Cannot get values above the top of the atmosphere, which is at: " +
                                              std::to_string(f->top_of_atmosphere) + " m.\nYour max input altitude is: " +
                                              std::to_string(a) + " m.");
    Stencil s;
    AB_TRY(stencil(*f, a, s));
    T[ip] = at(s, f->T, 1, f->nalt);
    P[ip] = at(s, f->P, 1, f->nalt);
    if (std::isnan(P[ip])) return set_error(AB200_ERR_INVALID, "Pressure is NaN");  // Point::check_and_fix :601-640
    if (std::isnan(T[ip])) return set_error(AB200_ERR_INVALID, "Temperature is NaN");
    for (int sp = 0; sp < n_species; sp++) {
      const double v = at(s, f->vmr + sp, n_species, f->nalt);
      if (std::isnan(v) || v < 0.0) return set_error(AB200_ERR_INVALID, "VMR for species " + std::to_string(sp) + " is " + std::to_string(v));
      vmr[static_cast<size_t>(ip) * n_species + sp] = v;
    }
    for (int i = 0; i < n_isot; i++) {
      if (f->isorat[i] < 0.0) return set_error(AB200_ERR_INVALID, "Isotopologue ratio for isotopologue " + std::to_string(i) + " is " + std::to_string(f->isorat[i]));
      isorat[static_cast<size_t>(ip) * n_isot + i] = f->isorat[i];
    }
    for (int c = 0; c < 3; c++) {
      if (mag) mag[3 * ip + c] = f->mag ? at(s, f->mag + c, 3, f->nalt) : 0.0;
      if (wind) wind[3 * ip + c] = f->wind ? at(s, f->wind + c, 3, f->nalt) : 0.0;
    }
  }
  if (f->partfun && Q && n_isot > 0 && np > 0) AB_TRY(ab200_partfun_eval(f->partfun, n_isot, np, T, Q, dQdT));
  return AB200_OK;
}
