// oracle/oracle.cpp — CPU restatement of the reference's clear-sky spectral hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (arts_b200/, include/) may
// link, import or execute this file; it is used by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs
// as the checker and the CPU baseline.
//
// What it restates (all paths relative to /root/reference):
//   stage 1  src/core/lbl/lbl_lineshape_voigt_lte.cpp  (engine A, the one behind
//            spectral_propmatAddLines, src/m_lbl.cc:242-300)
//            src/core/lbl/lbl_lineshape_model.cpp, lbl_temperature_model.h,
//            lbl_data.h, lbl_zeeman.{h,cpp}, lbl_lineshape.cpp:168-208
//   stage 2  src/core/rtepack/rtepack_transmission.cc, rtepack_source.cc,
//            rtepack_rtestep.cc, rtepack_multitype.h, rtepack_propagation_matrix.h
//   scalars  src/core/physics/physics_funcs.{h,cc}, src/core/util/arts_constants.h,
//            src/core/operators/spectral_radiance_transform_operator.cc
// Each function cites the lines it follows.  Faddeeva::w is NOT restated: the
// reference's own 3rdparty/Faddeeva/Faddeeva.cc is compiled where it lies and
// linked in (oracle/Makefile), so the dominant arithmetic of the path is the
// reference's object code.
//
// Parity pins (tests/test_oracle_*.py): Faddeeva 57-point KAT
// (Faddeeva.cc:4041-4195), the catalog-free fixture
// tests/core/linsrc/test_linsrc_convergence.py, exp(-K) against scipy expm
// (src/tests/test_rtepack.cc:12-33).  Full-pipeline goldens of the reference need
// external catalogs (arts-cat-data) and are unpinned here: see DESIGN.md.
//
// The flattened inputs are the C-ABI structs of include/arts_b200.h, so tests
// feed byte-identical inputs to this oracle and to the CUDA library.

#include <omp.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numbers>
#include <numeric>
#include <string>
#include <vector>

#include "../include/arts_b200.h"
#include "../arts_b200/csrc/predef_tables.h"  // published line lists of PWR98 / MPM89 (data only; pinned through the sliced reference)

namespace Faddeeva {
// reference 3rdparty/Faddeeva/Faddeeva.hh:36
extern std::complex<double> w(std::complex<double> z, double relerr);
// reference 3rdparty/Faddeeva/Faddeeva.hh:56 (used by tran::linsrc_linprop, rtepack_transmission.cc:453)
extern double Dawson(double x);
// reference 3rdparty/Faddeeva/Faddeeva.hh:55 (element-wise in rtepack::dawson(specmat), rtepack_spectral_matrix.cc:6-26)
extern std::complex<double> Dawson(std::complex<double> z, double relerr);
// reference 3rdparty/Faddeeva/Faddeeva.hh:46 (the speed-dependent line shape of PWR20xx::compute_h2o, PWR20xx.cc:145)
extern std::complex<double> erfcx(std::complex<double> z, double relerr);
}  // namespace Faddeeva

namespace {
using Numeric = double;
using Complex = std::complex<double>;
using Index   = std::int64_t;

thread_local std::string g_err;

// ---------------------------------------------------------------------------
// Constants: src/core/util/arts_constants.h:57-254 (same expressions, so the
// doubles are bit-identical)
// ---------------------------------------------------------------------------
namespace Constant {
constexpr Numeric pi          = std::numbers::pi;
constexpr Numeric inv_pi      = std::numbers::inv_pi;
constexpr Numeric two_pi      = 2 * pi;
constexpr Numeric inv_two_pi  = inv_pi / 2;
constexpr Numeric inv_sqrt_pi = std::numbers::inv_sqrtpi;
constexpr Numeric c           = 299792458;
constexpr Numeric h           = 6.62607015e-34;
constexpr Numeric h_bar       = h * inv_two_pi;
constexpr Numeric e           = 1.602176634e-19;
constexpr Numeric k           = 1.380649e-23;
constexpr Numeric NA          = 6.02214076e23;
constexpr Numeric alpha       = 7.2973525693e-3;
constexpr Numeric R_inf       = 10973731.568160;
constexpr Numeric m_e         = 2 * h * R_inf / (c * (alpha * alpha));
constexpr Numeric bohr_magneton = e * h_bar / (2 * m_e);
constexpr Numeric R           = k * NA;
constexpr Numeric doppler_broadening_const_squared = 2'000 * R / (c * c);
}  // namespace Constant

constexpr Numeric pow2(Numeric x) { return x * x; }
constexpr Numeric pow3(Numeric x) { return x * x * x; }
constexpr Numeric pow4(Numeric x) { return pow2(pow2(x)); }
inline Numeric deg2rad(Numeric x) { return x * (Constant::pi / 180); }

// ---------------------------------------------------------------------------
// physics: src/core/physics/physics_funcs.cc:153-158,192-197,254-263 and
// physics_funcs.h:54-72
// ---------------------------------------------------------------------------
Numeric planck(Numeric f, Numeric t) {
  constexpr Numeric a = 2 * Constant::h / pow2(Constant::c);
  constexpr Numeric b = Constant::h / Constant::k;
  return a * pow3(f) / std::expm1((b * f) / t);
}

Numeric dplanck_dt(Numeric f, Numeric t) {
  constexpr Numeric a        = 2 * Constant::h / pow2(Constant::c);
  constexpr Numeric b        = Constant::h / Constant::k;
  const Numeric inv_exp_t_m1 = 1.0 / std::expm1(b * f / t);
  return a * b * pow4(f) * inv_exp_t_m1 * (1 + inv_exp_t_m1) / pow2(t);
}

Numeric invplanck(Numeric i, Numeric f) {
  constexpr Numeric a = Constant::h / Constant::k;
  constexpr Numeric b = 2 * Constant::h / (Constant::c * Constant::c);
  return (a * f) / std::log1p((b * f * f * f) / i);
}

constexpr Numeric number_density(Numeric p, Numeric t) { return p / (Constant::k * t); }
constexpr Numeric dnumber_density_dt(Numeric p, Numeric t) { return -p / (Constant::k * pow2(t)); }

// ---------------------------------------------------------------------------
// Temperature models: src/core/lbl/lbl_temperature_model.h:62-314
// ---------------------------------------------------------------------------
// nonstd::pow, src/core/util/nonstd.h:29-33: the reference's temperature models do NOT call std::pow but
// exp(v log x) ("factor 4 faster"); the two differ in the last bits (pinned by tests/test_refslice_pins.py)
inline Numeric ns_pow(Numeric x, Numeric v) {
  return std::signbit(x) ? -std::exp(v * std::log(x < 0 ? -x : x)) : std::exp(v * std::log(x));
}

Numeric tm_value(int type, const double* x, Numeric T0, Numeric T) {
  switch (type) {
    case AB200_TM_T0: return x[0];
    case AB200_TM_T1: return x[0] * ns_pow(T0 / T, x[1]);
    case AB200_TM_T2: return x[0] * ns_pow(T0 / T, x[1]) * (1 + x[2] * std::log(T / T0));
    case AB200_TM_T3: return x[0] + x[1] * (T - T0);
    case AB200_TM_T4: return (x[0] + x[1] * (T0 / T - 1)) * ns_pow(T0 / T, x[2]);
    case AB200_TM_T5: return x[0] * ns_pow(T0 / T, 0.25 + 1.5 * x[1]);
    case AB200_TM_AER:
      if (T < 250.0) return x[0] + (T - 200.0) * (x[1] - x[0]) / (250.0 - 200.0);
      if (T > 296.0) return x[2] + (T - 296.0) * (x[3] - x[2]) / (340.0 - 296.0);
      return x[1] + (T - 250.0) * (x[2] - x[1]) / (296.0 - 250.0);
    case AB200_TM_DPL: return x[0] * ns_pow(T0 / T, x[1]) + x[2] * ns_pow(T0 / T, x[3]);
    case AB200_TM_POLY: {
      Numeric poly_fac = 1.0, poly_sum = 0.0;
      for (int i = 0; i < 4; i++) {
        poly_sum += x[i] * poly_fac;
        poly_fac *= T;
      }
      return poly_sum;
    }
  }
  return std::numeric_limits<Numeric>::quiet_NaN();
}

// d/dT of the above: lbl_temperature_model.h:81-283 (d*_dT members)
Numeric tm_dT(int type, const double* x, Numeric T0, Numeric T) {
  switch (type) {
    case AB200_TM_T0: return 0;
    case AB200_TM_T1: return -x[0] * x[1] * ns_pow(T0 / T, x[1]) / T;
    case AB200_TM_T2:
      return -x[0] * x[1] * ns_pow(T0 / T, x[1]) * (x[2] * std::log(T / T0) + 1.) / T +
             x[0] * x[2] * ns_pow(T0 / T, x[1]) / T;
    case AB200_TM_T3: return x[1];
    case AB200_TM_T4:
      return -x[2] * ns_pow(T0 / T, x[2]) * (x[0] + x[1] * (T0 / T - 1.)) / T -
             T0 * x[1] * ns_pow(T0 / T, x[2]) / (T * T);
    case AB200_TM_T5: return -x[0] * ns_pow(T0 / T, 1.5 * x[1] + 0.25) * (1.5 * x[1] + 0.25) / T;
    case AB200_TM_AER:
      if (T < 250.0) return (x[1] - x[0]) / (250.0 - 200.0);
      if (T > 296.0) return (x[3] - x[2]) / (340.0 - 296.0);
      return (x[2] - x[1]) / (296.0 - 250.0);
    case AB200_TM_DPL:
      return -x[0] * x[1] * ns_pow(T0 / T, x[1]) / T + -x[2] * x[3] * ns_pow(T0 / T, x[3]) / T;
    case AB200_TM_POLY: {
      Numeric poly_fac = 1.0, poly_sum = 0.0;
      for (int i = 1; i < 4; ++i) {
        poly_sum += static_cast<Numeric>(i) * x[i] * poly_fac;
        poly_fac *= T;
      }
      return poly_sum;
    }
  }
  return std::numeric_limits<Numeric>::quiet_NaN();
}

// ---------------------------------------------------------------------------
// Atmosphere point view
// ---------------------------------------------------------------------------
struct AtmPt {
  Numeric T, P;
  const double* vmr;     // [n_species]
  const double* isorat;  // [n_isot]
  const double* Q;
  const double* dQdT;
  Numeric mag[3];
  Numeric los[2];
  Numeric wind[3];
  Numeric vmr_of(int s) const { return s < 0 ? 0.0 : vmr[s]; }
};

AtmPt atm_at(const ab200_catalog_desc& d, const ab200_atm_path& a, int ip) {
  AtmPt p{};
  p.T      = a.T[ip];
  p.P      = a.P[ip];
  p.vmr    = a.vmr + static_cast<Index>(ip) * d.n_species;
  p.isorat = a.isorat + static_cast<Index>(ip) * d.n_isot;
  p.Q      = a.Q + static_cast<Index>(ip) * d.n_isot;
  p.dQdT   = a.dQdT ? a.dQdT + static_cast<Index>(ip) * d.n_isot : nullptr;
  for (int i = 0; i < 3; i++) p.mag[i] = a.mag ? a.mag[3 * ip + i] : 0.0;
  for (int i = 0; i < 2; i++) p.los[i] = a.los ? a.los[2 * ip + i] : 0.0;
  for (int i = 0; i < 3; i++) p.wind[i] = a.wind ? a.wind[3 * ip + i] : 0.0;
  return p;
}

// wind_shift, src/m_frequency_grid.cc:4-84: the frequency scaling factor and freq_wind_shift_jac
// (= d fac / d(u, v, w) / fac), with path::mirror, src/core/path/path_point.cpp:33-39.
// Returns false for a non-positive factor (the reference throws).
bool wind_factor(const AtmPt& atm, Numeric& fac, Numeric* jac = nullptr) {
  constexpr Numeric c = Constant::c;
  const Numeric u = atm.wind[0], v = atm.wind[1], w = atm.wind[2];
  Numeric za = 180 - atm.los[0], aa = atm.los[1] + 180;
  if (aa > 180) aa -= 360;
  const Numeric u2v2 = u * u + v * v;
  const Numeric w2   = w * w;
  const Numeric f2   = u2v2 + w2;
  const Numeric f    = std::sqrt(f2);
  const Numeric za_f = f == w ? 0.0 : std::acos(w / f);
  const Numeric aa_f = std::atan2(u, v);
  const Numeric za_p = deg2rad(za);
  const Numeric aa_p = deg2rad(aa);
  const Numeric scl  = f2 * std::sqrt(f2 - w2);
  const Numeric czaf = std::cos(za_f);
  const Numeric szaf = std::sin(za_f);
  const Numeric czap = std::cos(za_p);
  const Numeric szap = std::sin(za_p);
  const Numeric caa  = std::cos(aa_f - aa_p);
  const Numeric saa  = std::sin(aa_p - aa_f);
  const Numeric dp   = czaf * czap + szaf * szap * caa;
  fac                = 1.0 - (f * dp) / c;
  if (fac <= 0) return false;
  if (std::isnan(fac)) {  // "Zero shift if nan" :40-44
    fac = 1.0;
    if (jac) jac[0] = jac[1] = jac[2] = 0.0;
    return true;
  }
  if (jac) {
    {  // :56-63
      const Numeric df_du    = (f == 0) ? 1.0 : u / f;
      const Numeric dczaf_du = (f2 == 0) ? 0.0 : (-w * df_du / f2);
      const Numeric dszaf_du = (f2 == w2) ? 0.0 : (w2 * df_du / scl);
      const Numeric dcaa_du  = (u2v2 == 0) ? 0.0 : (v * saa / u2v2);
      const Numeric ddp_du   = czap * dczaf_du + szap * caa * dszaf_du + szap * szaf * dcaa_du;
      jac[0]                 = -(dp * df_du + f * ddp_du) / c;
    }
    {  // :65-72
      const Numeric df_dv    = (f == 0) ? 1.0 : v / f;
      const Numeric dczaf_dv = (f2 == 0) ? 0.0 : (-w * df_dv / f2);
      const Numeric dszaf_dv = (f2 == w2) ? 0.0 : (w2 * df_dv / scl);
      const Numeric dcaa_dv  = (u2v2 == 0) ? 0.0 : (-u * saa / u2v2);
      const Numeric ddp_dv   = czap * dczaf_dv + szap * caa * dszaf_dv + szap * szaf * dcaa_dv;
      jac[1]                 = -(dp * df_dv + f * ddp_dv) / c;
    }
    {  // :74-80
      const Numeric df_dw    = (f == 0) ? 1.0 : w / f;
      const Numeric dczaf_dw = (f2 == 0) ? 0.0 : (-w * df_dw / f2 + 1.0 / f);
      const Numeric dszaf_dw = (scl == 0) ? 0.0 : ((w2 * df_dw - f * w) / scl);
      const Numeric ddp_dw   = czap * dczaf_dw + szap * caa * dszaf_dw;
      jac[2]                 = -(dp * df_dw + f * ddp_dw) / c;
    }
    for (int i = 0; i < 3; i++) jac[i] /= fac;  // :82
  }
  return true;
}

// d/dX0 .. d/dX3 of the temperature models: lbl_temperature_model.h:65-275 (the d*_dX* members; coefficients a model
// does not have give 0, the EMPTY overloads :37-60)
Numeric tm_dX(int type, int k, const double* x, Numeric T0, Numeric T) {
  using std::log;
  const auto pow = [](Numeric x, Numeric v) { return ns_pow(x, v); };  // nonstd::pow, as the reference
  switch (type) {
    case AB200_TM_T0: return k == 0 ? 1 : 0;
    case AB200_TM_T1:
      if (k == 0) return pow(T0 / T, x[1]);
      if (k == 1) return x[0] * pow(T0 / T, x[1]) * log(T0 / T);
      return 0;
    case AB200_TM_T2:
      if (k == 0) return pow(T0 / T, x[1]) * (1 + x[2] * log(T / T0));
      if (k == 1) return x[0] * pow(T0 / T, x[1]) * (x[2] * log(T / T0) + 1.) * log(T0 / T);
      if (k == 2) return x[0] * pow(T0 / T, x[1]) * log(T / T0);
      return 0;
    case AB200_TM_T3: return k == 0 ? 1 : k == 1 ? T - T0 : 0;
    case AB200_TM_T4:
      if (k == 0) return pow(T0 / T, x[2]);
      if (k == 1) return pow(T0 / T, x[2]) * (T0 / T - 1.);
      if (k == 2) return pow(T0 / T, x[2]) * (x[0] + x[1] * (T0 / T - 1)) * log(T0 / T);
      return 0;
    case AB200_TM_T5:
      if (k == 0) return pow(T0 / T, 1.5 * x[1] + 0.25);
      if (k == 1) return 1.5 * x[0] * pow(T0 / T, 1.5 * x[1] + 0.25) * log(T0 / T);
      return 0;
    case AB200_TM_AER:
      if (k == 0) return T < 250.0 ? 1 - (T - 200.0) / (250.0 - 200.0) : 0;
      if (k == 1) return T < 250.0 ? (T - 200.0) / (250.0 - 200.0) : T > 296.0 ? 0 : 1 - (T - 250.0) / (296.0 - 250.0);
      if (k == 2) return T < 250.0 ? 0 : T > 296.0 ? 1 - (T - 296.0) / (340.0 - 296.0) : (T - 250.0) / (296.0 - 250.0);
      return T > 296.0 ? (T - 296.0) / (340.0 - 296.0) : 0;
    case AB200_TM_DPL:
      if (k == 0) return pow(T0 / T, x[1]);
      if (k == 1) return x[0] * pow(T0 / T, x[1]) * log(T0 / T);
      if (k == 2) return pow(T0 / T, x[3]);
      return x[2] * pow(T0 / T, x[3]) * log(T0 / T);
    case AB200_TM_POLY: return k == 0 ? 1.0 : k == 1 ? T : k == 2 ? T * T : T * T * T;
  }
  return std::numeric_limits<Numeric>::quiet_NaN();
}

// ---------------------------------------------------------------------------
// Line-shape model: src/core/lbl/lbl_lineshape_model.cpp:14-35 (pressure
// scaling), :70-113 (mixing + dVMR), :127-148 (dT).
// ---------------------------------------------------------------------------
struct LineView {
  const ab200_catalog_desc& d;
  Index l;
  Numeric a() const { return d.a[l]; }
  Numeric f0() const { return d.f0[l]; }
  Numeric e0() const { return d.e0[l]; }
  Numeric gu() const { return d.gu[l]; }
  Numeric T0() const { return d.T0[l]; }

  static Numeric pscale(int var, Numeric P) {
    return (var == AB200_VAR_G || var == AB200_VAR_DV) ? P * P : P;
  }

  // species_model::VAR(T0,T,P): lbl_lineshape_model.cpp:14-35
  Numeric single(Index ils, int var, Numeric T, Numeric P, bool dT) const {
    const int type = d.ls_type[ils * AB200_NVAR + var];
    if (type == AB200_TM_ABSENT) return 0.0;
    const double* x = d.ls_X + (ils * AB200_NVAR + var) * 4;
    return pscale(var, P) * (dT ? tm_dT(type, x, T0(), T) : tm_value(type, x, T0(), T));
  }

  // model::VAR(atm) and model::dVAR_dT(atm): lbl_lineshape_model.cpp:70-90,127-148
  Numeric mix(int var, const AtmPt& atm, bool dT = false) const {
    Numeric vmr = 0.0, res = 0.0, bth = std::numeric_limits<Numeric>::quiet_NaN();
    for (Index i = d.ls_offset[l]; i < d.ls_offset[l + 1]; i++) {
      const Numeric this_res = single(i, var, atm.T, atm.P, dT);
      if (d.ls_species[i] != AB200_SPECIES_BATH) {
        const Numeric this_vmr  = atm.vmr_of(d.ls_species[i]);
        vmr                    += this_vmr;
        res                    += this_vmr * this_res;
      } else {
        bth = this_res;
      }
    }
    if (not std::isnan(bth)) return res + (1.0 - vmr) * bth;
    return res / vmr;
  }

  // model::dVAR_dVMR(atm, species): lbl_lineshape_model.cpp:92-113
  Numeric dmix_dVMR(int var, const AtmPt& atm, int species) const {
    Index ptr = -1, bth = -1;
    for (Index i = d.ls_offset[l]; i < d.ls_offset[l + 1]; i++) {
      if (d.ls_species[i] == species) ptr = i;
      if (d.ls_species[i] == AB200_SPECIES_BATH) bth = i;
    }
    if (ptr < 0) return 0.0;
    const Numeric x = single(ptr, var, atm.T, atm.P, false);
    if (species == AB200_SPECIES_BATH) return -x;
    if (bth >= 0) return x - single(bth, var, atm.T, atm.P, false);
    Numeric t = 0.0;
    for (Index i = d.ls_offset[l]; i < d.ls_offset[l + 1]; i++) t += atm.vmr_of(d.ls_species[i]);
    return (t - x) / t * t;  // sic, lbl_lineshape_model.cpp:112
  }

  // model::dVAR_dX(atm, species, coeff): lbl_lineshape_model.cpp:150-246 with species_model::dVAR_dXk :38-63
  Numeric dmix_dX(int var, const AtmPt& atm, int species, int coeff) const {
    Index ptr = -1, bth = -1;
    for (Index i = d.ls_offset[l]; i < d.ls_offset[l + 1]; i++) {
      if (d.ls_species[i] == species) ptr = i;
      if (d.ls_species[i] == AB200_SPECIES_BATH) bth = i;
    }
    if (ptr < 0) return 0.0;
    const int type  = d.ls_type[ptr * AB200_NVAR + var];
    const Numeric x = type == AB200_TM_ABSENT
                          ? 0.0
                          : pscale(var, atm.P) * tm_dX(type, coeff, d.ls_X + (ptr * AB200_NVAR + var) * 4, T0(), atm.T);
    if (species == AB200_SPECIES_BATH) {
      Numeric vmr = 0.0;
      for (Index i = d.ls_offset[l]; i < d.ls_offset[l + 1]; i++)
        vmr += d.ls_species[i] == AB200_SPECIES_BATH ? 0.0 : atm.vmr_of(d.ls_species[i]);
      return (1 - vmr) * x;
    }
    if (bth >= 0) return atm.vmr_of(species) * x;
    Numeric vmr = 0.0;
    for (Index i = d.ls_offset[l]; i < d.ls_offset[l + 1]; i++) vmr += atm.vmr_of(d.ls_species[i]);
    return x * atm.vmr_of(species) / vmr;
  }

  // line::s(T,Q): lbl_data.h:66-68
  Numeric s(Numeric T, Numeric Q) const {
    return a() * gu() * std::exp(-e0() / (Constant::k * T)) / (pow3(f0()) * Q);
  }
  // line::ds_dT: lbl_data.h:138-142
  Numeric ds_dT(Numeric T, Numeric Q, Numeric dQ_dT) const {
    return a() * gu() * (e0() * Q - Constant::k * pow2(T) * dQ_dT) * std::exp(-e0() / (Constant::k * T)) /
           (pow3(f0()) * Constant::k * pow2(T) * pow2(Q));
  }
};

// ---------------------------------------------------------------------------
// Zeeman: src/core/lbl/lbl_zeeman.h:18-160,342-352 and lbl_zeeman.cpp:261-309,
// :321-455.  wigner3j(Jl,1,Ju,ml,dm,-mu) is the reference's 3rdparty/wigner
// call (lbl_zeeman.cpp:276); here it is the Racah formula in long double, all
// arguments doubled integers.
// ---------------------------------------------------------------------------
long double lfact(int n) { return std::lgamma(static_cast<long double>(n) + 1.0L); }

Numeric wigner3j_2(int tj1, int tj2, int tj3, int tm1, int tm2, int tm3) {
  if (tm1 + tm2 + tm3 != 0) return 0;
  if (std::abs(tm1) > tj1 || std::abs(tm2) > tj2 || std::abs(tm3) > tj3) return 0;
  if (tj3 > tj1 + tj2 || tj3 < std::abs(tj1 - tj2)) return 0;
  if ((tj1 + tj2 + tj3) % 2) return 0;
  if ((tj1 + tm1) % 2 || (tj2 + tm2) % 2 || (tj3 + tm3) % 2) return 0;
  auto h = [](int x) { return x / 2; };
  const long double ldelta = lfact(h(tj1 + tj2 - tj3)) + lfact(h(tj1 - tj2 + tj3)) + lfact(h(-tj1 + tj2 + tj3)) -
                             lfact(h(tj1 + tj2 + tj3) + 1);
  const long double lpre = 0.5L * (ldelta + lfact(h(tj1 + tm1)) + lfact(h(tj1 - tm1)) + lfact(h(tj2 + tm2)) +
                                   lfact(h(tj2 - tm2)) + lfact(h(tj3 + tm3)) + lfact(h(tj3 - tm3)));
  const int kmin = std::max({0, h(tj2 - tj3 - tm1), h(tj1 - tj3 + tm2)});
  const int kmax = std::min({h(tj1 + tj2 - tj3), h(tj1 - tm1), h(tj2 + tm2)});
  long double sum = 0;
  for (int k = kmin; k <= kmax; k++) {
    const long double lden = lfact(k) + lfact(h(tj1 + tj2 - tj3) - k) + lfact(h(tj1 - tm1) - k) +
                             lfact(h(tj2 + tm2) - k) + lfact(h(tj3 - tj2 + tm1) + k) +
                             lfact(h(tj3 - tj1 - tm2) + k);
    sum += ((k % 2) ? -1.0L : 1.0L) * std::exp(lpre - lden);
  }
  const int ph = h(tj1 - tj2 - tm3);
  return static_cast<Numeric>(((ph % 2) ? -1.0L : 1.0L) * sum);
}

enum Pol { POL_NO = 0, POL_PI = 1, POL_SM = 2, POL_SP = 3 };

int zeeman_dM(Pol p) { return p == POL_SM ? -1 : (p == POL_SP ? 1 : 0); }                    // lbl_zeeman.h:18-26
Numeric polarization_factor(Pol p) { return p == POL_PI ? 1.5 : (p == POL_NO ? 1.0 : .75); }  // :154-162

struct ZeemanView {
  bool on;
  Numeric gu, gl;
  int tJu, tJl;

  // model::size, lbl_zeeman.cpp:298-309 with zeeman::size lbl_zeeman.h:92-96
  Index size(Pol p) const {
    if (on) return p == POL_NO ? 0 : tJl + 1;
    return p == POL_NO ? 1 : 0;
  }
  // 2*Ml, 2*Mu: lbl_zeeman.h:112-136
  int tMl(Pol, Index n) const { return -tJl + 2 * static_cast<int>(n); }
  int tMu(Pol p, Index n) const { return tMl(p, n) + 2 * zeeman_dM(p); }

  // model::Strength, lbl_zeeman.cpp:261-277
  Numeric Strength(Pol p, Index n) const {
    if (p == POL_NO) return 1.0;
    const int ml = tMl(p, n), mu = tMu(p, n);
    if (std::abs(ml) > tJl or std::abs(mu) > tJu) return 0.0;
    const Numeric C = polarization_factor(p);
    return C * pow2(wigner3j_2(tJl, 2, tJu, ml, 2 * zeeman_dM(p), -mu));
  }
  // model::Splitting, lbl_zeeman.h:342-352
  Numeric Splitting(Pol p, Index n) const {
    constexpr Numeric C = Constant::bohr_magneton / Constant::h;
    if (p == POL_NO) return 0.0;
    return C * (0.5 * tMu(p, n) * gu - 0.5 * tMl(p, n) * gl);
  }
};

// zeeman::norm_view, lbl_zeeman.cpp:321-331,413-455
void norm_view(Pol p, const Numeric mag[3], const Numeric los[2], Numeric npm[7]) {
  const Numeric u = mag[0], v = mag[1], w = mag[2];
  const Numeric sa = std::sin(deg2rad(los[1])), ca = std::cos(deg2rad(los[1]));
  const Numeric sz = std::sin(deg2rad(los[0])), cz = std::cos(deg2rad(los[0]));
  const Numeric H    = std::hypot(u, v, w);
  const Numeric uct  = sz * sa * u + sz * ca * v + cz * w;
  const Numeric duct = u * sa * cz + v * ca * cz - w * sz;
  const Numeric theta = H == 0 ? 0 : std::acos(uct / H);
  const Numeric eta   = -std::atan2(ca * u - sa * v, -duct);
  const Numeric CT    = std::cos(theta);
  const Numeric ST2   = pow2(std::sin(theta));
  const Numeric Q     = ST2 * std::cos(2 * eta);
  const Numeric U     = ST2 * std::sin(2 * eta);
  const Numeric pi_[7] = {ST2, -Q, U, 0, 0, U, Q};
  const Numeric sm_[7] = {2 - ST2, Q, -U, 2 * CT, -2 * CT, -U, -Q};
  const Numeric sp_[7] = {2 - ST2, Q, -U, -2 * CT, 2 * CT, -U, -Q};
  const Numeric no_[7] = {1, 0, 0, 0, 0, 0, 0};
  const Numeric* src = p == POL_PI ? pi_ : p == POL_SM ? sm_ : p == POL_SP ? sp_ : no_;
  for (int i = 0; i < 7; i++) npm[i] = src[i];
}

// zeeman::dnorm_view_du / dv / dw, lbl_zeeman.cpp:457-536, with magnetic_angles::dtheta_d* and deta_d* (:361-411);
// comp = 0, 1, 2 for u, v, w
void dnorm_view(Pol p, int comp, const Numeric mag[3], const Numeric los[2], Numeric dnpm[7]) {
  const Numeric u = mag[0], v = mag[1], w = mag[2];
  const Numeric sa = std::sin(deg2rad(los[1])), ca = std::cos(deg2rad(los[1]));
  const Numeric sz = std::sin(deg2rad(los[0])), cz = std::cos(deg2rad(los[0]));
  const Numeric H    = std::hypot(u, v, w);
  const Numeric uct  = sz * sa * u + sz * ca * v + cz * w;
  const Numeric duct = u * sa * cz + v * ca * cz - w * sz;
  const Numeric theta = H == 0 ? 0 : std::acos(uct / H);
  const Numeric eta   = -std::atan2(ca * u - sa * v, -duct);
  Numeric dtheta, deta;
  {
    const Numeric rat = pow2(uct / H);
    const Numeric nom = comp == 0   ? u * uct - sa * sz * pow2(H)
                        : comp == 1 ? v * uct - ca * sz * pow2(H)
                                    : w * uct - cz * pow2(H);
    dtheta = (H == 0.0 or rat == 1.0) ? 0 : nom / (std::sqrt(1.0 - rat) * (H * H * H));
    const Numeric den = pow2(ca * u - sa * v) + pow2(duct);
    deta = H == 0 ? 0
                  : (comp == 0   ? (cz * v - ca * sz * w)
                     : comp == 1 ? (sa * sz * w - cz * u)
                                 : sz * (ca * u - sa * v)) /
                        den;
  }
  const Numeric CT   = std::cos(theta);
  const Numeric ST   = std::sin(theta);
  const Numeric CE   = std::cos(2 * eta);
  const Numeric SE   = std::sin(2 * eta);
  const Numeric ST2  = pow2(ST);
  const Numeric dST2 = 2 * dtheta * ST * CT;
  const Numeric dQ   = 2 * dtheta * ST * CE * CT - 2 * deta * SE * ST2;
  const Numeric dU   = 2 * deta * ST2 * CE + 2 * dtheta * SE * ST * CT;
  const Numeric dCT  = -dtheta * ST;
  const Numeric pi_[7] = {dST2, -dQ, dU, 0, 0, dU, dQ};
  const Numeric sm_[7] = {-dST2, dQ, -dU, 2 * dCT, -2 * dCT, -dU, -dQ};
  const Numeric sp_[7] = {-dST2, dQ, -dU, -2 * dCT, 2 * dCT, -dU, -dQ};
  const Numeric no_[7] = {0, 0, 0, 0, 0, 0, 0};
  const Numeric* src = p == POL_PI ? pi_ : p == POL_SM ? sm_ : p == POL_SP ? sp_ : no_;
  for (int i = 0; i < 7; i++) dnpm[i] = src[i];
}

// ---------------------------------------------------------------------------
// single_shape: src/core/lbl/lbl_lineshape_voigt_lte.h:20-56 and
// lbl_lineshape_voigt_lte.cpp:22-36,145-204,239-268
// ---------------------------------------------------------------------------
struct single_shape {
  Numeric f0{}, inv_gd{}, z_imag{};
  Complex s{};
  // VP_LTE_MIRROR (lbl_lineshape_voigt_lte_mirrored.{h,cpp}): the same shape plus its mirror image at -f0,
  // F(f) = w(z(f)) + w(zm(f)), zm = inv_gd (f + f0) + i z_imag (.h:44-46, .cpp:220)
  bool mirror{false};
  Complex z(Numeric f) const { return Complex{inv_gd * (f - f0), z_imag}; }
  Complex zm(Numeric f) const { return Complex{inv_gd * (f + f0), z_imag}; }
  static Complex F(Complex z_) { return Faddeeva::w(z_, 0); }
  Complex operator()(Numeric f) const { return mirror ? s * (F(z(f)) + F(zm(f))) : s * F(z(f)); }
  // forward finite difference, lbl_lineshape_voigt_lte.cpp:250-268
  static Complex dF(Complex z_, Complex F_) {
    const Complex dz{std::max(1e-4 * std::abs(z_.real()), 1e-4), std::max(1e-4 * std::abs(z_.imag()), 1e-4)};
    const Complex F_2 = Faddeeva::w(z_ + dz, 0);
    return (F_2 - F_) / dz;
  }
  // single_shape::dH, lbl_lineshape_voigt_lte.cpp:305-307; mirrored: s dz_dH (dFp + dFm), ..._mirrored.cpp
  Complex dH(Complex dz_dH, Numeric f) const {
    const Complex z_ = z(f);
    if (mirror) {
      const Complex zm_ = zm(f);
      return s * dz_dH * (dF(z_, F(z_)) + dF(zm_, F(zm_)));
    }
    return s * dz_dH * dF(z_, F(z_));
  }
  // single_shape::df, lbl_lineshape_voigt_lte.cpp:275 (dF(f) :245-248); mirrored: ..._mirrored.cpp:244-248, :262
  Complex df(Numeric f) const {
    const Complex z_ = z(f);
    if (mirror) {
      const Complex zm_ = zm(f);
      return s * inv_gd * (dF(z_, F(z_)) + dF(zm_, F(zm_)));
    }
    return s * inv_gd * dF(z_, F(z_));
  }
  // single_shape::dT / dVMR, lbl_lineshape_voigt_lte.cpp:310-323; mirrored: lbl_lineshape_voigt_lte_mirrored.cpp:305-325
  // (z_ = zp - zm, F_ = Fp + Fm, dF_ = dFp + dFm - literal, including the frequency-independent z_)
  Complex dX(Complex ds, Complex dz, Numeric dz_fac, Numeric f) const {
    if (mirror) {
      const Complex zp_ = z(f), zm_ = zm(f);
      const Complex Fp_ = F(zp_), Fm_ = F(zm_);
      const Complex dFp_ = dF(zp_, Fp_), dFm_ = dF(zm_, Fm_);
      const Complex z_ = zp_ - zm_;
      return ds * (Fp_ + Fm_) + s * (dz + dz_fac * z_) * (dFp_ + dFm_);
    }
    const Complex z_ = z(f);
    const Complex F_ = F(z_);
    const Complex dF_ = dF(z_, F_);
    return ds * F_ + s * (dz + dz_fac * z_) * dF_;
  }
};

struct line_pos {
  Index line;
  Index iz;
};

// line_strength_calc, lbl_lineshape_voigt_lte.cpp:22-36
Complex line_strength_calc(Numeric inv_gd, int isot, int spec, const LineView& ln, const AtmPt& atm) {
  const auto s    = ln.s(atm.T, atm.Q[isot]);
  const Numeric G = ln.mix(AB200_VAR_G, atm);
  const Numeric Y = ln.mix(AB200_VAR_Y, atm);
  const Complex lm{1 + G, -Y};
  const Numeric r = atm.isorat[isot];
  const Numeric x = atm.vmr_of(spec);
  return Constant::inv_sqrt_pi * inv_gd * r * x * lm * s;
}

// dline_strength_calc_dY / dG / df0, lbl_lineshape_voigt_lte.cpp:38-84 (line::ds_df0_s_ratio = -3 / f0, lbl_data.h:118)
Complex dline_strength_calc_dY(Numeric dY, Numeric inv_gd, int isot, int spec, const LineView& ln, const AtmPt& atm) {
  const auto s    = ln.s(atm.T, atm.Q[isot]);
  const Numeric r = atm.isorat[isot];
  const Numeric x = atm.vmr_of(spec);
  return Constant::inv_sqrt_pi * inv_gd * r * x * Complex(0, -dY) * s;
}
Complex dline_strength_calc_dG(Numeric dG, Numeric inv_gd, int isot, int spec, const LineView& ln, const AtmPt& atm) {
  const auto s    = ln.s(atm.T, atm.Q[isot]);
  const Numeric r = atm.isorat[isot];
  const Numeric x = atm.vmr_of(spec);
  return Constant::inv_sqrt_pi * inv_gd * r * x * dG * s;
}
Complex dline_strength_calc_df0(Numeric f0, Numeric inv_gd, int isot, int spec, const LineView& ln, const AtmPt& atm) {
  const auto s    = ln.s(atm.T, atm.Q[isot]);
  const auto ds   = (-3 / ln.f0()) * s;
  const Numeric G = ln.mix(AB200_VAR_G, atm);
  const Numeric Y = ln.mix(AB200_VAR_Y, atm);
  const Complex lm{1 + G, -Y};
  const Numeric r = atm.isorat[isot];
  const Numeric x = atm.vmr_of(spec);
  return Constant::inv_sqrt_pi * inv_gd * r * x * (f0 * ds - s) * lm / f0;
}

// dline_strength_calc_dVMR, lbl_lineshape_voigt_lte.cpp:86-114
Complex dline_strength_calc_dVMR(Numeric inv_gd, Numeric f0, int isot, int spec, int target_spec,
                                 const LineView& ln, const AtmPt& atm) {
  const auto s      = ln.s(atm.T, atm.Q[isot]);
  const Numeric G   = ln.mix(AB200_VAR_G, atm);
  const Numeric Y   = ln.mix(AB200_VAR_Y, atm);
  const Numeric dG  = ln.dmix_dVMR(AB200_VAR_G, atm, target_spec);
  const Numeric dY  = ln.dmix_dVMR(AB200_VAR_Y, atm, target_spec);
  const Numeric dD0 = ln.dmix_dVMR(AB200_VAR_D0, atm, target_spec);
  const Numeric dDV = ln.dmix_dVMR(AB200_VAR_DV, atm, target_spec);
  const Numeric df0 = dD0 + dDV;
  const Complex lm{1 + G, -Y};
  const Complex dlm = {dG, -dY};
  const Numeric r   = atm.isorat[isot];
  const Numeric x   = atm.vmr_of(spec);
  if (target_spec == spec) {
    return -Constant::inv_sqrt_pi * inv_gd * r * s * (x * (df0 / f0) * lm - (x * dlm + lm));
  }
  return -Constant::inv_sqrt_pi * inv_gd * r * s * x * ((df0 / f0) * lm - dlm);
}

// dline_strength_calc_dT, lbl_lineshape_voigt_lte.cpp:116-143
Complex dline_strength_calc_dT(Numeric inv_gd, Numeric f0, int isot, int spec, const LineView& ln,
                               const AtmPt& atm) {
  const Numeric T   = atm.T;
  const auto s      = ln.s(T, atm.Q[isot]);
  const auto ds     = ln.ds_dT(T, atm.Q[isot], atm.dQdT ? atm.dQdT[isot] : 0.0);
  const Numeric G   = ln.mix(AB200_VAR_G, atm);
  const Numeric Y   = ln.mix(AB200_VAR_Y, atm);
  const Numeric dG  = ln.mix(AB200_VAR_G, atm, true);
  const Numeric dY  = ln.mix(AB200_VAR_Y, atm, true);
  const Numeric dD0 = ln.mix(AB200_VAR_D0, atm, true);
  const Numeric dDV = ln.mix(AB200_VAR_DV, atm, true);
  const Numeric df0 = dD0 + dDV;
  const Complex lm{1 + G, -Y};
  const Complex dlm = {dG, -dY};
  const Numeric r   = atm.isorat[isot];
  const Numeric x   = atm.vmr_of(spec);
  return Constant::inv_sqrt_pi * inv_gd * r * x *
         (2 * T * (dlm * s + lm * ds) * f0 - 2 * T * df0 * lm * s - f0 * lm * s) / (2 * T * f0);
}

// line_center_calc, :145-147
Numeric line_center_calc(const LineView& ln, const AtmPt& atm) {
  return ln.f0() + ln.mix(AB200_VAR_D0, atm) + ln.mix(AB200_VAR_DV, atm);
}

// band_shape_helper + lines_push_back + zeeman_push_back + single_shape_builder,
// lbl_lineshape_voigt_lte.cpp:165-204,338-429.  ByLine uses band_data::active_lines
// (lbl_data.cpp:61-68) on the catalog f0 (lines of a band are sorted by f0).
void band_shape_helper(std::vector<single_shape>& lines, std::vector<line_pos>& pos,
                       const ab200_catalog_desc& d, int ib, const AtmPt& atm, Numeric fmin, Numeric fmax,
                       Pol pol) {
  lines.resize(0);
  pos.resize(0);
  const int isot = d.band_isot[ib];
  const int spec = d.isot_species[isot];
  Index lo = d.band_offset[ib], hi = d.band_offset[ib + 1];
  if (d.band_cutoff_type[ib] == AB200_CUTOFF_BYLINE) {
    const Numeric c = d.band_cutoff_value[ib];
    const double* b = d.f0 + lo;
    const double* e = d.f0 + hi;
    const double* low = std::lower_bound(b, e, fmin - c);
    const double* upp = std::upper_bound(low, e, fmax + c);
    hi = lo + (upp - b);
    lo = lo + (low - b);
  }
  const Numeric H = std::hypot(atm.mag[0], atm.mag[1], atm.mag[2]);
  for (Index il = lo; il < hi; il++) {
    const LineView ln{d, il};
    const ZeemanView z{d.z_on[il] != 0, d.z_gu[il], d.z_gl[il], d.two_Ju[il], d.two_Jl[il]};
    if (not((z.on and pol != POL_NO) or (not z.on and pol == POL_NO))) continue;
    // single_shape_builder ctor :177-186
    const Numeric f0             = line_center_calc(ln, atm);
    const Numeric scaled_gd_part = std::sqrt(Constant::doppler_broadening_const_squared * atm.T / d.isot_mass[isot]);
    const Numeric G0             = ln.mix(AB200_VAR_G0, atm);
    if (pol == POL_NO) {
      single_shape s;  // operator single_shape() :198-204
      s.f0     = f0;
      s.inv_gd = 1.0 / (scaled_gd_part * f0);
      s.z_imag = G0 * s.inv_gd;
      s.s      = line_strength_calc(s.inv_gd, isot, spec, ln, atm);
      s.mirror = d.band_lineshape[ib] == AB200_LINESHAPE_VP_LTE_MIRROR;
      lines.push_back(s);
      pos.push_back({il, std::numeric_limits<Index>::max()});
    } else {
      const Index nz = z.size(pol);
      for (Index iz = 0; iz < nz; iz++) {
        single_shape s;  // as_zeeman :188-196 (inv_gd uses the unsplit centre)
        s.f0     = f0 + H * z.Splitting(pol, iz);
        s.inv_gd = 1.0 / (scaled_gd_part * f0);
        s.z_imag = G0 * s.inv_gd;
        s.s      = z.Strength(pol, iz) * line_strength_calc(s.inv_gd, isot, spec, ln, atm);
        s.mirror = d.band_lineshape[ib] == AB200_LINESHAPE_VP_LTE_MIRROR;
        if (s.s == 0.0) continue;  // pop_back :354-357
        lines.push_back(s);
        pos.push_back({il, iz});
      }
    }
  }
  // stdr::sort(zip(lines,pos)) by f0 :426-428
  std::vector<Index> order(lines.size());
  std::iota(order.begin(), order.end(), Index{0});
  std::sort(order.begin(), order.end(), [&](Index a, Index b) { return lines[a].f0 < lines[b].f0; });
  std::vector<single_shape> l2(lines.size());
  std::vector<line_pos> p2(pos.size());
  for (size_t i = 0; i < order.size(); i++) {
    l2[i] = lines[order[i]];
    p2[i] = pos[order[i]];
  }
  lines.swap(l2);
  pos.swap(p2);
}

// find_offset_and_count_of_frequency_range, lbl_lineshape_voigt_lte.h:123-133
std::pair<Index, Index> freq_range(const std::vector<single_shape>& lines, Numeric f, Numeric cutoff) {
  if (cutoff < std::numeric_limits<Numeric>::infinity()) {
    auto low = std::lower_bound(lines.begin(), lines.end(), f - cutoff,
                                [](const single_shape& l, Numeric v) { return l.f0 < v; });
    auto upp = std::upper_bound(lines.begin(), lines.end(), f + cutoff,
                                [](Numeric v, const single_shape& l) { return v < l.f0; });
    return {low - lines.begin(), upp - low};
  }
  return {0, static_cast<Index>(lines.size())};
}

// zeeman::scale, lbl_zeeman.h:432-440
inline void add_scaled(double* pm, const Numeric npm[7], Complex F) {
  pm[0] += npm[0] * F.real();
  pm[1] += npm[1] * F.real();
  pm[2] += npm[2] * F.real();
  pm[3] += npm[3] * F.real();
  pm[4] += npm[4] * F.imag();
  pm[5] += npm[5] * F.imag();
  pm[6] += npm[6] * F.imag();
}

// voigt::lte::calculate for one (band, pol) over one frequency range,
// lbl_lineshape_voigt_lte.cpp:1652-1725 with ComputeData ctor :936-956,
// core_calc :959-981, dt_core_calc :984-1033, dVMR_core_calc :1153-1189 and
// compute_derivative :1463-1561.
void calculate_band(double* pm, double* dpm, Index nf_total, const double* f_grid, Index f_lo, Index f_n,
                    const ab200_catalog_desc& d, int ib, const AtmPt& atm, Pol pol, int nq,
                    const ab200_target* targets, bool no_negative_absorption, std::vector<single_shape>& lines,
                    std::vector<line_pos>& pos) {
  Numeric npm[7];
  norm_view(pol, atm.mag, atm.los, npm);
  if (std::all_of(npm, npm + 7, [](Numeric n) { return n == 0; })) return;
  if (f_n == 0) return;
  const double* fg = f_grid + f_lo;
  const int isot   = d.band_isot[ib];
  const int spec   = d.isot_species[isot];

  // fmin/fmax of band_data::active_lines (:1672-1680): the reference takes them from the
  // frequency range of the call.  Its chunked stand-alone branch (m_lbl.cc:273-295) therefore
  // selects a slightly different line set per OpenMP chunk when a pressure shift moves a line
  // across the cutoff window of a chunk-edge frequency (thread-count dependent, DESIGN.md
  // quirk 7).  The parity target is the unchunked call (m_lbl.cc:256-271, the branch every
  // path-level caller takes, m_propmat.cc:42): the bounds are those of the whole grid of the
  // call even when this oracle splits the grid over threads.
  band_shape_helper(lines, pos, d, ib, atm, f_grid[0], f_grid[nf_total - 1], pol);
  if (lines.empty()) return;
  const bool has_cut   = d.band_cutoff_type[ib] != AB200_CUTOFF_NONE;
  const Numeric cutoff = has_cut ? d.band_cutoff_value[ib] : std::numeric_limits<Numeric>::infinity();
  const size_t nl      = lines.size();

  // ComputeData ctor: scl :944-953
  std::vector<Numeric> scl(f_n);
  {
    const Numeric N = number_density(atm.P, atm.T), T = atm.T;
    for (Index i = 0; i < f_n; i++) {
      constexpr Numeric c = Constant::c * Constant::c / (8 * Constant::pi);
      const Numeric r     = (Constant::h * fg[i]) / (Constant::k * T);
      scl[i]              = -N * fg[i] * std::expm1(-r) * c;
    }
  }

  // core_calc :959-981 (+ band_shape::operator() :431-436, :591-608)
  std::vector<Complex> shape(f_n), cut(nl);
  if (has_cut) {
    for (size_t i = 0; i < nl; i++) cut[i] = lines[i](lines[i].f0 + cutoff);
    for (Index i = 0; i < f_n; i++) {
      const auto [start, count] = freq_range(lines, fg[i], cutoff);
      Complex out{};
      for (Index j = start; j < start + count; j++) out += lines[j](fg[i]) - cut[j];
      shape[i] = out;
    }
  } else {
    for (Index i = 0; i < f_n; i++) {
      Complex out{};
      for (size_t j = 0; j < nl; j++) out += lines[j](fg[i]);
      shape[i] = out;
    }
  }

  // :1688-1692
  for (Index i = 0; i < f_n; i++) {
    const auto F = scl[i] * shape[i];
    if (no_negative_absorption and F.real() < 0) continue;
    add_scaled(pm + (f_lo + i) * 7, npm, F);
  }

  // Jacobian targets :1694-1708
  std::vector<Complex> ds(nl), dz(nl), dcut(nl), dshape(f_n);
  std::vector<Numeric> dz_fac(nl), dscl(f_n);
  for (int iq = 0; iq < nq; iq++) {
    double* dp = dpm + (static_cast<Index>(iq) * nf_total + f_lo) * 7;
    const Numeric T = atm.T;
    bool is_T = targets[iq].kind == AB200_TARGET_T;
    const bool is_wind = targets[iq].kind >= AB200_TARGET_WIND_U and targets[iq].kind <= AB200_TARGET_WIND_W;
    if (is_wind) {
      // df_core_calc :1036-1062 (band_shape::df :438-443, :610-627; single_shape::df :275), compute_derivative
      // :1514-1523; the three wind components share it, spectral_propmat_jacWindFix tells them apart
      const Numeric N = number_density(atm.P, T);
      for (Index i = 0; i < f_n; i++) {
        constexpr Numeric c = Constant::c * Constant::c / (8 * Constant::pi);
        const Numeric r     = (Constant::h * fg[i]) / (Constant::k * T);
        dscl[i]             = N * (r * std::exp(-r) - std::expm1(-r)) * c;
      }
      if (has_cut) {
        for (size_t i = 0; i < nl; i++) dcut[i] = lines[i].df(lines[i].f0 + cutoff);
        for (Index i = 0; i < f_n; i++) {
          const auto [start, count] = freq_range(lines, fg[i], cutoff);
          Complex out{};
          for (Index j = start; j < start + count; j++) out += lines[j].df(fg[i]) - dcut[j];
          dshape[i] = out;
        }
      } else {
        for (Index i = 0; i < f_n; i++) {
          Complex out{};
          for (size_t j = 0; j < nl; j++) out += lines[j].df(fg[i]);
          dshape[i] = out;
        }
      }
      for (Index i = 0; i < f_n; i++) add_scaled(dp + i * 7, npm, dscl[i] * shape[i] + scl[i] * dshape[i]);
      continue;
    }
    if (targets[iq].kind == AB200_TARGET_ISORAT) {
      // compute_derivative(SpeciesIsotope) :1526-1544 (a zero ratio is rejected by orc_propmat_levels, :1539)
      if (targets[iq].species != isot) continue;
      const Numeric isorat = atm.isorat[isot];
      for (Index i = 0; i < f_n; i++) add_scaled(dp + i * 7, npm, scl[i] * shape[i] / isorat);
      continue;
    }
    if (targets[iq].kind >= AB200_TARGET_LINE_F0 and targets[iq].kind <= AB200_TARGET_LINE_LS) {
      // compute_derivative(line_key) :1562-1637 without cutoff: only the band that holds the line (lbl_lineshape.cpp
      // hands line targets to their own band), only its (Zeeman sub-)lines (set_filter :1192-1201)
      const ab200_target& key = targets[iq];
      if (key.line < d.band_offset[ib] or key.line >= d.band_offset[ib + 1]) continue;
      if (has_cut) continue;  // rejected by orc_propmat_levels before it gets here
      std::fill(dshape.begin(), dshape.end(), Complex{});
      for (size_t i = 0; i < nl; i++) {
        if (pos[i].line != key.line) continue;
        const LineView ln{d, pos[i].line};
        const ZeemanView z{d.z_on[ln.l] != 0, d.z_gu[ln.l], d.z_gl[ln.l], d.two_Ju[ln.l], d.two_Jl[ln.l]};
        const single_shape& lshp = lines[i];
        const Numeric inv_gd = lshp.inv_gd, f0 = lshp.f0;
        Complex dsi{}, dzi{};
        Numeric dzf  = 0;
        bool only_ds = false, only_dz = false;
        switch (key.kind) {
          case AB200_TARGET_LINE_F0:  // df0_core_calc :1204-1238, single_shape::df0 :277-283
            dzf = -1.0 / f0;
            dsi = z.Strength(pol, pos[i].iz) * dline_strength_calc_df0(f0, inv_gd, isot, spec, ln, atm);
            dzi = -inv_gd;
            break;
          case AB200_TARGET_LINE_E0:  // de0_core_calc :1241-1265, ds_de0_s_ratio lbl_data.h:101-103, de0 :329-331
            dsi     = (-1 / (Constant::k * atm.T)) * lshp.s;
            only_ds = true;
            break;
          case AB200_TARGET_LINE_A:  // da_core_calc :1268-1290, da :325-327
            dsi     = (1.0 / ln.a()) * lshp.s;
            only_ds = true;
            break;
          default: {
            const Numeric dv = ln.dmix_dX(key.ls_var, atm, key.species, key.coeff);
            switch (key.ls_var) {
              case AB200_VAR_G0:  // dG0_core_calc :1293-1317, dG0 :301-303
                dzi     = Complex(0, inv_gd * dv);
                only_dz = true;
                break;
              case AB200_VAR_D0:  // dD0_core_calc :1320-1349, dD0 :293-299
              case AB200_VAR_DV:  // dDV_core_calc :1419-1450, dDV :285-291
                dzf = -dv / f0;
                dsi = key.ls_var == AB200_VAR_D0 ? -dv * lshp.s / f0 : lshp.s * dzf;
                dzi = -dv * inv_gd;
                break;
              case AB200_VAR_Y:  // dY_core_calc :1352-1383, dY :337-339
                dsi     = z.Strength(pol, pos[i].iz) * dline_strength_calc_dY(dv, inv_gd, isot, spec, ln, atm);
                only_ds = true;
                break;
              case AB200_VAR_G:  // dG_core_calc :1386-1416, dG :333-335
                dsi     = z.Strength(pol, pos[i].iz) * dline_strength_calc_dG(dv, inv_gd, isot, spec, ln, atm);
                only_ds = true;
                break;
              default: break;
            }
          }
        }
        for (Index j = 0; j < f_n; j++) {
          const Complex z_ = lshp.z(fg[j]);
          const Complex F_ = single_shape::F(z_);
          if (only_ds) {
            dshape[j] += dsi * F_;
          } else if (only_dz) {
            dshape[j] += lshp.s * dzi * single_shape::dF(z_, F_);
          } else {
            dshape[j] += dsi * F_ + lshp.s * (dzi + dzf * z_) * single_shape::dF(z_, F_);
          }
        }
      }
      for (Index i = 0; i < f_n; i++) add_scaled(dp + i * 7, npm, scl[i] * dshape[i]);
      continue;
    }
    if (targets[iq].kind >= AB200_TARGET_MAG_U and targets[iq].kind <= AB200_TARGET_MAG_W) {
      // compute_derivative :1484-1513 with dmag_{u,v,w}_core_calc :1066-1162 (band_shape::dH :445-455, :629-655)
      if (pol == POL_NO) continue;
      const int comp      = targets[iq].kind - AB200_TARGET_MAG_U;
      const Numeric H     = std::hypot(atm.mag[0], atm.mag[1], atm.mag[2]);
      const Numeric dH_dm = atm.mag[comp] / H;
      for (size_t i = 0; i < nl; i++) {
        const LineView ln{d, pos[i].line};
        const ZeemanView z{d.z_on[ln.l] != 0, d.z_gu[ln.l], d.z_gl[ln.l], d.two_Ju[ln.l], d.two_Jl[ln.l]};
        dz[i] = -lines[i].inv_gd * dH_dm * z.Splitting(pol, pos[i].iz);
      }
      if (has_cut) {
        for (size_t i = 0; i < nl; i++) dcut[i] = lines[i].dH(dz[i], lines[i].f0 + cutoff);
        for (Index i = 0; i < f_n; i++) {
          const auto [start, count] = freq_range(lines, fg[i], cutoff);
          Complex out{};
          for (Index j = start; j < start + count; j++) out += lines[j].dH(dz[j], fg[i]) - dcut[j];
          dshape[i] = out;
        }
      } else {
        for (Index i = 0; i < f_n; i++) {
          Complex out{};
          for (size_t j = 0; j < nl; j++) out += lines[j].dH(dz[j], fg[i]);
          dshape[i] = out;
        }
      }
      Numeric dnpm[7];
      dnorm_view(pol, comp, atm.mag, atm.los, dnpm);
      for (Index i = 0; i < f_n; i++) {  // zeeman::scale(a, da, F, dF), lbl_zeeman.h:442-453
        const Complex F = scl[i] * shape[i], dF = scl[i] * dshape[i];
        double* o = dp + i * 7;
        for (int c = 0; c < 4; c++) o[c] += dnpm[c] * F.real() + npm[c] * dF.real();
        for (int c = 4; c < 7; c++) o[c] += dnpm[c] * F.imag() + npm[c] * dF.imag();
      }
      continue;
    }
    if (is_T) {
      // dt_core_calc :984-1033
      const Numeric N = number_density(atm.P, T), dN = dnumber_density_dt(atm.P, T);
      for (Index i = 0; i < f_n; i++) {
        constexpr Numeric c = Constant::c * Constant::c / (8 * Constant::pi);
        const Numeric r     = (Constant::h * fg[i]) / (Constant::k * T);
        dscl[i]             = -fg[i] * (N * r * std::exp(-r) / T + dN * std::expm1(-r)) * c;
      }
      for (size_t i = 0; i < nl; i++) {
        const LineView ln{d, pos[i].line};
        const ZeemanView z{d.z_on[ln.l] != 0, d.z_gu[ln.l], d.z_gl[ln.l], d.two_Ju[ln.l], d.two_Jl[ln.l]};
        const Numeric inv_gd = lines[i].inv_gd;
        const Numeric f0     = lines[i].f0;
        const Numeric dD0 = ln.mix(AB200_VAR_D0, atm, true), dDV = ln.mix(AB200_VAR_DV, atm, true);
        dz_fac[i] = (-2 * T * dD0 - 2 * T * dDV - f0) / (2 * T * f0);
        ds[i]     = z.Strength(pol, pos[i].iz) * dline_strength_calc_dT(inv_gd, f0, isot, spec, ln, atm);
        dz[i]     = inv_gd * Complex{-(dD0 + dDV), ln.mix(AB200_VAR_G0, atm, true)};
      }
    } else {
      // dVMR_core_calc :1153-1189
      const int tspec = targets[iq].species;
      for (size_t i = 0; i < nl; i++) {
        const LineView ln{d, pos[i].line};
        const ZeemanView z{d.z_on[ln.l] != 0, d.z_gu[ln.l], d.z_gl[ln.l], d.two_Ju[ln.l], d.two_Jl[ln.l]};
        const Numeric inv_gd = lines[i].inv_gd;
        const Numeric f0     = lines[i].f0;
        const Numeric dD0 = ln.dmix_dVMR(AB200_VAR_D0, atm, tspec), dDV = ln.dmix_dVMR(AB200_VAR_DV, atm, tspec);
        dz_fac[i] = -(dD0 + dDV) / f0;
        ds[i]     = z.Strength(pol, pos[i].iz) * dline_strength_calc_dVMR(inv_gd, f0, isot, spec, tspec, ln, atm);
        dz[i]     = inv_gd * Complex{-(dD0 + dDV), ln.dmix_dVMR(AB200_VAR_G0, atm, tspec)};
      }
    }
    // band_shape::dT / dVMR with and without cutoff :475-507, :655-720
    if (has_cut) {
      for (size_t i = 0; i < nl; i++) dcut[i] = lines[i].dX(ds[i], dz[i], dz_fac[i], lines[i].f0 + cutoff);
      for (Index i = 0; i < f_n; i++) {
        const auto [start, count] = freq_range(lines, fg[i], cutoff);
        Complex out{};
        for (Index j = start; j < start + count; j++) out += lines[j].dX(ds[j], dz[j], dz_fac[j], fg[i]) - dcut[j];
        dshape[i] = out;
      }
    } else {
      for (Index i = 0; i < f_n; i++) {
        Complex out{};
        for (size_t j = 0; j < nl; j++) out += lines[j].dX(ds[j], dz[j], dz_fac[j], fg[i]);
        dshape[i] = out;
      }
    }
    // compute_derivative :1474-1481 (T) and :1553-1560 (VMR)
    for (Index i = 0; i < f_n; i++) {
      const Complex dF = is_T ? dscl[i] * shape[i] + scl[i] * dshape[i] : scl[i] * dshape[i];
      add_scaled(dp + i * 7, npm, dF);
    }
  }
}

// lbl::calculate, src/core/lbl/lbl_lineshape.cpp:75-209 (VP_LTE bands only):
// pol=no over all selected bands, then pi, sm, sp.
void lbl_calculate(double* pm, double* dpm, Index nf, const double* f_grid, Index f_lo, Index f_n,
                   const ab200_catalog_desc& d, const AtmPt& atm, int select_species, int nq,
                   const ab200_target* targets, bool no_negative_absorption) {
  std::vector<single_shape> lines;
  std::vector<line_pos> pos;
  for (Pol pol : {POL_NO, POL_PI, POL_SM, POL_SP}) {
    for (int ib = 0; ib < d.n_bands; ib++) {
      const int spec = d.isot_species[d.band_isot[ib]];
      if (select_species == spec or select_species == AB200_SPECIES_BATH) {
        calculate_band(pm, dpm, nf, f_grid, f_lo, f_n, d, ib, atm, pol, nq, targets, no_negative_absorption,
                       lines, pos);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// rtepack value types
// ---------------------------------------------------------------------------
struct propmat {
  Numeric v[7];
  Numeric A() const { return v[0]; }
  Numeric B() const { return v[1]; }
  Numeric C() const { return v[2]; }
  Numeric D() const { return v[3]; }
  Numeric U() const { return v[4]; }
  Numeric V() const { return v[5]; }
  Numeric W() const { return v[6]; }
  // rtepack_propagation_matrix.h:41-49
  bool is_rotational() const { return A() == 0.0 and B() == 0.0 and C() == 0.0 and D() == 0.0; }
  bool is_polarized() const { return B() != 0 or C() != 0 or D() != 0 or U() != 0 or V() != 0 or W() != 0; }
};
inline propmat load_pm(const double* p) {
  propmat k;
  std::memcpy(k.v, p, sizeof(k.v));
  return k;
}

struct stokvec {
  Numeric v[4]{0, 0, 0, 0};
};
inline stokvec operator+(stokvec a, const stokvec& b) {
  for (int i = 0; i < 4; i++) a.v[i] += b.v[i];
  return a;
}
inline stokvec operator-(stokvec a, const stokvec& b) {
  for (int i = 0; i < 4; i++) a.v[i] -= b.v[i];
  return a;
}
inline stokvec operator-(stokvec a) {
  for (int i = 0; i < 4; i++) a.v[i] = -a.v[i];
  return a;
}
// rtepack_stokes_vector.h:120 (std::midpoint per element)
inline stokvec avg(const stokvec& a, const stokvec& b) {
  stokvec o;
  for (int i = 0; i < 4; i++) o.v[i] = std::midpoint(a.v[i], b.v[i]);
  return o;
}

struct muelmat {
  Numeric m[16];
  muelmat() { id(1.0); }
  muelmat(Numeric d) { id(d); }  // diagonal ctor, rtepack_mueller_matrix.h
  void id(Numeric d) {
    for (auto& x : m) x = 0;
    m[0] = m[5] = m[10] = m[15] = d;
  }
  static muelmat zero() { return muelmat(0.0); }
};
inline muelmat make_mm(std::initializer_list<Numeric> l) {
  muelmat o;
  std::copy(l.begin(), l.end(), o.m);
  return o;
}
inline muelmat operator*(const muelmat& a, const muelmat& b) {  // rtepack_mueller_matrix.h:101-150
  muelmat o = muelmat::zero();
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      o.m[4 * i + j] = a.m[4 * i + 0] * b.m[0 + j] + a.m[4 * i + 1] * b.m[4 + j] + a.m[4 * i + 2] * b.m[8 + j] +
                       a.m[4 * i + 3] * b.m[12 + j];
  return o;
}
// muelmat * Numeric, rtepack_mueller_matrix.h:190-193 -> `a *= b`.  muelmat declares operator*=(const muelmat&)
// (:101), which HIDES the element-wise operator*= of its cdata_t base, so the scalar is converted to b*I by the
// diagonal constructor (:13-14) and the full 4x4 product runs.  Same value for finite input; the sign of a zero
// entry (x*b + y*0 + z*0 + w*0 is +0 where x*b is -0) and NaN/Inf propagation differ from an element-wise scaling.
// Found by the bitwise pin against the sliced reference code (tests/test_refslice_pins.py).
inline muelmat operator*(muelmat a, Numeric s) { return a * muelmat(s); }
inline muelmat operator*(Numeric s, muelmat a) { return a * s; }
inline muelmat operator+(muelmat a, const muelmat& b) {
  for (int i = 0; i < 16; i++) a.m[i] += b.m[i];
  return a;
}
// rtepack_multitype.h:58-68
inline stokvec operator*(const muelmat& a, const stokvec& b) {
  const Numeric* m = a.m;
  const Numeric s1 = b.v[0], s2 = b.v[1], s3 = b.v[2], s4 = b.v[3];
  stokvec o;
  o.v[0] = m[0] * s1 + m[1] * s2 + m[2] * s3 + m[3] * s4;
  o.v[1] = m[4] * s1 + m[5] * s2 + m[6] * s3 + m[7] * s4;
  o.v[2] = m[9] * s2 + m[10] * s3 + m[11] * s4 + m[8] * s1;
  o.v[3] = m[12] * s1 + m[13] * s2 + m[14] * s3 + m[15] * s4;
  return o;
}


// ---------------------------------------------------------------------------
// rtepack::specmat — src/core/rtepack/rtepack_spectral_matrix.h:12-242: a 4x4 complex matrix, row major.  Only what the
// polarised branch of tran::linsrc_linprop (rtepack_transmission.cc:467-474) touches.  Scaling by a Complex goes
// through the diagonal constructor and the FULL matrix product (`a *= b` finds specmat::operator*=(const specmat&),
// :56-107, like muelmat above), so `inv(A) = adj(A) / det(A)` (:242, :185-188) is adj(A) x ((1/det) I).
// ---------------------------------------------------------------------------
struct specmat {
  Complex m[16];
  specmat(Complex tau = 1.0) {
    for (auto& x : m) x = 0;
    m[0] = m[5] = m[10] = m[15] = tau;
  }
};
inline specmat operator*(const specmat& a, const specmat& b) {  // :56-107
  specmat o(0.0);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      o.m[4 * i + j] = a.m[4 * i + 0] * b.m[0 + j] + a.m[4 * i + 1] * b.m[4 + j] + a.m[4 * i + 2] * b.m[8 + j] +
                       a.m[4 * i + 3] * b.m[12 + j];
  return o;
}
inline specmat operator/(const specmat& a, const Complex& b) { return a * specmat(1.0 / b); }  // :185-188
inline specmat operator-(specmat a, const specmat& b) {
  for (int i = 0; i < 16; i++) a.m[i] -= b.m[i];
  return a;
}
inline Complex det(const specmat& A) {  // :203-210
  const Complex *q = A.m, a = q[0], b = q[1], c = q[2], d = q[3], e = q[4], f = q[5], g = q[6], h = q[7], i = q[8], j = q[9],
                k = q[10], l = q[11], m = q[12], n = q[13], o = q[14], p = q[15];
  return a * (f * (k * p - l * o) + g * (l * n - j * p) + h * (j * o - k * n)) +
         b * (e * (l * o - k * p) + g * (i * p - l * m) + h * (k * m - i * o)) +
         c * (e * (j * p - l * n) + f * (l * m - i * p) + h * (i * n - j * m)) +
         d * (e * (k * n - j * o) + f * (i * o - k * m) + g * (j * m - i * n));
}
inline specmat adj(const specmat& A) {  // :223-240
  const Complex *q = A.m, a = q[0], b = q[1], c = q[2], d = q[3], e = q[4], f = q[5], g = q[6], h = q[7], i = q[8], j = q[9],
                k = q[10], l = q[11], m = q[12], n = q[13], o = q[14], p = q[15];
  specmat r(0.0);
  r.m[0]  = f * (k * p - l * o) + g * (l * n - j * p) + h * (j * o - k * n);
  r.m[1]  = b * (l * o - k * p) + c * (j * p - l * n) + d * (k * n - j * o);
  r.m[2]  = b * (g * p - h * o) + c * (h * n - f * p) + d * (f * o - g * n);
  r.m[3]  = b * (h * k - g * l) + c * (f * l - h * j) + d * (g * j - f * k);
  r.m[4]  = e * (l * o - k * p) + g * (i * p - l * m) + h * (k * m - i * o);
  r.m[5]  = a * (k * p - l * o) + c * (l * m - i * p) + d * (i * o - k * m);
  r.m[6]  = a * (h * o - g * p) + c * (e * p - h * m) + d * (g * m - e * o);
  r.m[7]  = a * (g * l - h * k) + c * (h * i - e * l) + d * (e * k - g * i);
  r.m[8]  = e * (j * p - l * n) + f * (l * m - i * p) + h * (i * n - j * m);
  r.m[9]  = a * (l * n - j * p) + b * (i * p - l * m) + d * (j * m - i * n);
  r.m[10] = a * (f * p - h * n) + b * (h * m - e * p) + d * (e * n - f * m);
  r.m[11] = a * (h * j - f * l) + b * (e * l - h * i) + d * (f * i - e * j);
  r.m[12] = e * (k * n - j * o) + f * (i * o - k * m) + g * (j * m - i * n);
  r.m[13] = a * (j * o - k * n) + b * (k * m - i * o) + c * (i * n - j * m);
  r.m[14] = a * (g * n - f * o) + b * (e * o - g * m) + c * (f * m - e * n);
  r.m[15] = a * (f * k - g * j) + b * (g * i - e * k) + c * (e * j - f * i);
  return r;
}
inline specmat inv(const specmat& A) { return adj(A) / det(A); }  // :242
// specmat * propmat, rtepack_multitype.h:145-168 (the order of the four terms is the reference's)
inline specmat operator*(const specmat& s, const propmat& k) {
  const Numeric a = k.v[0], b = k.v[1], c = k.v[2], d = k.v[3], u = k.v[4], v = k.v[5], w = k.v[6];
  specmat o(0.0);
  for (int i = 0; i < 4; i++) {
    const Complex m1 = s.m[4 * i], m2 = s.m[4 * i + 1], m3 = s.m[4 * i + 2], m4 = s.m[4 * i + 3];
    o.m[4 * i + 0] = a * m1 + b * m2 + c * m3 + d * m4;
    o.m[4 * i + 1] = a * m2 + b * m1 - m3 * u - m4 * v;
    o.m[4 * i + 2] = a * m3 + c * m1 + m2 * u - m4 * w;
    o.m[4 * i + 3] = a * m4 + d * m1 + m2 * v + m3 * w;
  }
  return o;
}
// muelmat * specmat, rtepack_multitype.h:200-250
inline specmat operator*(const muelmat& a, const specmat& b) {
  specmat o(0.0);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      o.m[4 * i + j] = a.m[4 * i + 0] * b.m[0 + j] + a.m[4 * i + 1] * b.m[4 + j] + a.m[4 * i + 2] * b.m[8 + j] +
                       a.m[4 * i + 3] * b.m[12 + j];
  return o;
}
// element-wise Dawson function, rtepack_spectral_matrix.cc:6-26
inline specmat dawson(const specmat& A) {
  specmat o(0.0);
  for (int i = 0; i < 16; i++) o.m[i] = Faddeeva::Dawson(A.m[i], 0);
  return o;
}
inline muelmat real(const specmat& A) {  // rtepack_multitype.cc:116-133
  muelmat o;
  for (int i = 0; i < 16; i++) o.m[i] = std::real(A.m[i]);
  return o;
}
// propmat / Numeric: element-wise (matpack_mdspan_cdata_t.h:316-319); propmat - propmat
inline propmat operator/(propmat a, Numeric b) {
  for (auto& x : a.v) x /= b;
  return a;
}
inline propmat operator-(propmat a, const propmat& b) {
  for (int i = 0; i < 7; i++) a.v[i] -= b.v[i];
  return a;
}
inline propmat operator+(propmat a, const propmat& b) {
  for (int i = 0; i < 7; i++) a.v[i] += b.v[i];
  return a;
}
inline propmat operator*(propmat a, Numeric b) {
  for (auto& x : a.v) x *= b;
  return a;
}

// specmat sqrt(const propmat&), rtepack_transmission.cc:872-1002: the principal square root of the propagation matrix
// through its Cayley-Hamilton coefficients d0..d3 (complex: a - x and a +- i y may lie left of the branch cut).
inline specmat sqrt_pm(const propmat& pm) {
  const Numeric a      = pm.A();
  const Complex sqrt_a = std::sqrt(Complex(a));
  if (not pm.is_polarized()) return specmat(sqrt_a);
  const Numeric b = pm.B(), c = pm.C(), d = pm.D(), u = pm.U(), v = pm.V(), w = pm.W();
  const Numeric b2 = b * b, c2 = c * c, d2 = d * d, u2 = u * u, v2 = v * v, w2 = w * w;
  Complex d0c{}, d1c{}, d2c{}, d3c{};
  constexpr Numeric eps = std::numeric_limits<Numeric>::epsilon();
  if (pm.is_rotational()) {
    const Numeric rho = std::hypot(u, v, w);
    if (rho <= eps) return specmat(0.0);
    const Numeric r = std::sqrt(2.0 * rho);
    d0c             = 0.0;
    d1c             = 1.0 / r;
    d2c             = -1.0 / (rho * r);
    d3c             = 0.0;
  } else {
    const Numeric B      = u2 + v2 + w2 - b2 - c2 - d2;
    const Numeric C      = -pow2(d * u - c * v + b * w);
    const Numeric S      = std::sqrt(B * B - 4 * C);
    const Numeric x2     = std::max(0.0, 0.5 * (S - B));
    const Numeric abs_y2 = std::max(0.0, 0.5 * (S + B));
    const Numeric x      = std::sqrt(x2);
    const Complex y      = Complex(0, std::sqrt(abs_y2));
    const Complex sx     = std::sqrt(Complex(a + x));
    const Complex dx     = std::sqrt(Complex(a - x));
    const Complex sy     = std::sqrt(a + y);
    const Complex dy     = std::sqrt(a - y);
    const Complex Sx = sx + dx, Dx = sx - dx, Sy = sy + dy, Dy = sy - dy;
    if (x2 + abs_y2 <= eps) {
      d0c = sqrt_a;
      if (a <= eps) {  // sic (:934): the series coefficients are set for a SMALL a
        d1c = 0.5 / sqrt_a;
        d2c = 0.125 / (a * sqrt_a);
        d3c = 0.0625 / (a * a * sqrt_a);
      }
    } else {
      const Numeric inv_sum_sq = 1.0 / (x2 + abs_y2);
      d0c                      = (abs_y2 * Sx + x2 * Sy) * (0.5 * inv_sum_sq);
      d2c                      = (Sx - Sy) * (0.5 * inv_sum_sq);
      const Complex term1 = (x <= eps and a <= eps) ? Complex(0.0) : (x <= eps) ? 0.5 / sqrt_a : 0.5 * Dx / x;
      const Complex term2 = (abs_y2 <= eps and a <= eps) ? Complex(0.0) : (abs_y2 <= eps) ? 0.5 / sqrt_a : 0.5 * Dy / y;
      d1c                 = (abs_y2 * term1 + x2 * term2) * inv_sum_sq;
      d3c                 = (term1 - term2) * inv_sum_sq;
    }
  }
  const Numeric k2_00 = b2 + c2 + d2, k2_11 = b2 - u2 - v2, k2_22 = c2 - u2 - w2, k2_33 = d2 - v2 - w2;
  const Numeric k2_01 = -(c * u + d * v), k2_02 = b * u - d * w, k2_03 = b * v + c * w;
  const Numeric k2_12 = b * c - v * w, k2_13 = b * d + u * w, k2_23 = c * d - u * v;
  const Numeric k3_01 = b * k2_00 - u * k2_02 - v * k2_03;
  const Numeric k3_02 = c * k2_00 + u * k2_01 - w * k2_03;
  const Numeric k3_03 = d * k2_00 + v * k2_01 + w * k2_02;
  const Numeric k3_12 = -c * k2_01 + u * k2_11 - w * k2_13;
  const Numeric k3_13 = -d * k2_01 + v * k2_11 + w * k2_12;
  const Numeric k3_23 = -d * k2_02 + v * k2_12 + w * k2_22;
  specmat K;  // the default constructor is the identity (:14-16); every element is assigned below
  K.m[0]  = d0c + d2c * k2_00;
  K.m[5]  = d0c + d2c * k2_11;
  K.m[10] = d0c + d2c * k2_22;
  K.m[15] = d0c + d2c * k2_33;
  K.m[1]  = d1c * b + d2c * k2_01 + d3c * k3_01;
  K.m[4]  = d1c * b - d2c * k2_01 + d3c * k3_01;
  K.m[2]  = d1c * c + d2c * k2_02 + d3c * k3_02;
  K.m[8]  = d1c * c - d2c * k2_02 + d3c * k3_02;
  K.m[3]  = d1c * d + d2c * k2_03 + d3c * k3_03;
  K.m[12] = d1c * d - d2c * k2_03 + d3c * k3_03;
  K.m[6]  = d1c * u + d2c * k2_12 + d3c * k3_12;
  K.m[9]  = -d1c * u + d2c * k2_12 - d3c * k3_12;
  K.m[7]  = d1c * v + d2c * k2_13 + d3c * k3_13;
  K.m[13] = -d1c * v + d2c * k2_13 - d3c * k3_13;
  K.m[11] = d1c * w + d2c * k2_23 + d3c * k3_23;
  K.m[14] = -d1c * w + d2c * k2_23 - d3c * k3_23;
  return K;
}

// ---------------------------------------------------------------------------
// rtepack::tran — src/core/rtepack/rtepack_transmission.cc:20-150 (ctor and
// operator()), :207-275 (linsrc), :277-447 (linsrc_deriv), :558-674 (deriv).
// `exact` selects the mathematically exact eigen pair x^2=(S-B)/2, y^2=(S+B)/2
// instead of the reference's literal lines :67-70 (x2 = sqrt(t1), x = sqrt(x2)),
// see DESIGN.md "reference quirks".  Default (exact=false) is the reference.
// ---------------------------------------------------------------------------
constexpr Numeric too_small = 1e-4;

struct tran {
  Numeric a, exp_a;
  Numeric b, c, d, u, v, w;
  Numeric b2, c2, d2, u2, v2, w2;
  Numeric B, C, S;
  Numeric x2, y2, x, y, cy, sy, cx, sx;
  Numeric ix, iy, inv_x2y2;
  Numeric C0, C1, C2, C3;
  bool polarized, x_zero, y_zero, both_zero, either_zero;

  tran(const propmat& k1, const propmat& k2, const Numeric r, bool exact)
      : a{-0.5 * r * (k1.A() + k2.A())}, exp_a{std::exp(a)}, polarized(k1.is_polarized() or k2.is_polarized()) {
    if (not polarized) return;
    b = -0.5 * r * (k1.B() + k2.B());
    c = -0.5 * r * (k1.C() + k2.C());
    d = -0.5 * r * (k1.D() + k2.D());
    u = -0.5 * r * (k1.U() + k2.U());
    v = -0.5 * r * (k1.V() + k2.V());
    w = -0.5 * r * (k1.W() + k2.W());
    b2 = b * b;
    c2 = c * c;
    d2 = d * d;
    u2 = u * u;
    v2 = v * v;
    w2 = w * w;
    B  = u2 + v2 + w2 - b2 - c2 - d2;
    C  = -pow2(d * u - c * v + b * w);
    const Numeric disc = B * B - 4 * C;
    S                  = std::sqrt(std::max<Numeric>(0.0, disc));
    const Numeric t1   = 0.5 * (S - B);
    const Numeric t2   = 0.5 * (S + B);
    if (exact) {
      x2 = std::max<Numeric>(0.0, t1);
      y2 = std::max<Numeric>(0.0, t2);
    } else {
      x2 = std::sqrt(std::max<Numeric>(0.0, t1));  // :67
      y2 = std::sqrt(std::max<Numeric>(0.0, t2));  // :68
    }
    x  = std::sqrt(x2);
    y  = std::sqrt(y2);
    cy = std::cos(y);
    sy = std::sin(y);
    cx = std::cosh(x);
    sx = std::sinh(x);
    x_zero      = x < too_small;
    y_zero      = y < too_small;
    both_zero   = y_zero and x_zero;
    either_zero = y_zero or x_zero;
    ix          = x_zero ? 0.0 : 1.0 / x;
    iy          = y_zero ? 0.0 : 1.0 / y;
    inv_x2y2    = both_zero ? 1.0 : 1.0 / (x2 + y2);
    C0          = either_zero ? 1.0 : (cy * x2 + cx * y2) * inv_x2y2;
    C1          = either_zero ? 1.0 : (sy * x2 * iy + sx * y2 * ix) * inv_x2y2;
    C2          = both_zero ? 0.5 : (cx - cy) * inv_x2y2;
    C3          = both_zero ? 1.0 / 6.0 : ((x_zero ? 1.0 : sx * ix) - (y_zero ? 1.0 : sy * iy)) * inv_x2y2;
    polarized   = std::isfinite(C0) and std::isfinite(C1) and std::isfinite(C2) and std::isfinite(C3);
  }

  muelmat operator()() const {  // :118-150
    if (not polarized) return muelmat(exp_a);
    const Numeric C2b = C2 * (c * u + d * v);
    const Numeric C2c = C2 * (b * u - d * w);
    const Numeric C2d = C2 * (b * v + c * w);
    const Numeric C2u = C2 * (b * c - v * w);
    const Numeric C2v = C2 * (b * d + u * w);
    const Numeric C2w = C2 * (c * d - u * v);
    const Numeric C3b = C3 * (b * (B - w2) + w * (c * v - d * u));
    const Numeric C3c = C3 * (c * (v2 - B) - v * (d * u + b * w));
    const Numeric C3d = C3 * (d * (u2 - B) - u * (c * v - b * w));
    const Numeric C3u = C3 * (d * (c * v - b * w) - u * (B + d2));
    const Numeric C3v = C3 * (c * (d * u + b * w) - v * (B + c2));
    const Numeric C3w = C3 * (b * (c * v - d * u) - w * (B + b2));
    const Numeric M00 = C0 + C2 * (b2 + c2 + d2);
    const Numeric M11 = C0 + C2 * (b2 - u2 - v2);
    const Numeric M22 = C0 + C2 * (c2 - u2 - w2);
    const Numeric M33 = C0 + C2 * (d2 - v2 - w2);
    return exp_a * make_mm({M00, C1 * b - C2b - C3b, C1 * c + C2c + C3c, C1 * d + C2d + C3d,
                            C1 * b + C2b - C3b, M11, C1 * u + C2u + C3u, C1 * v + C2v + C3v,
                            C1 * c - C2c + C3c, -C1 * u + C2u - C3u, M22, C1 * w + C2w + C3w,
                            C1 * d - C2d + C3d, -C1 * v + C2v - C3v, -C1 * w + C2w - C3w, M33});
  }

  static Numeric func_F(Numeric z) { return std::abs(z) < 1e-8 ? 1.0 + z * 0.5 + z * z / 6.0 : std::expm1(z) / z; }
  static Numeric func_Fp(Numeric z) {
    if (std::abs(z) < too_small) return 0.5 + z / 3.0 + z * z / 8.0;
    const Numeric ez = std::exp(z);
    return (ez * (z - 1.0) + 1.0) / (z * z);
  }
  static Numeric func_Fpp(Numeric z) {
    if (std::abs(z) < too_small) return 1.0 / 3.0 + z / 4.0 + z * z / 10.0;
    const Numeric ez = std::exp(z);
    return (ez * (z * z - 2.0 * z + 2.0) - 2.0) / (z * z * z);
  }
  static Numeric func_F3p(Numeric z) {
    if (std::abs(z) < too_small) return 0.25 + z * 0.2;
    const Numeric ez = std::exp(z);
    const Numeric z2 = z * z;
    const Numeric z3 = z2 * z;
    return (ez * (z3 - 3.0 * z2 + 6.0 * z - 6.0) + 6.0) / (z3 * z);
  }
  static Numeric func_F4p(Numeric z) {
    if (std::abs(z) < too_small) return 0.2 + z / 6.0;
    const Numeric ez = std::exp(z);
    const Numeric z2 = z * z;
    const Numeric z3 = z2 * z;
    const Numeric z4 = z2 * z2;
    const Numeric z5 = z4 * z;
    return (ez * (z4 - 4.0 * z3 + 12.0 * z2 - 24.0 * z + 24.0) - 24.0) / z5;
  }

  muelmat S_mat() const { return make_mm({0, b, c, d, b, 0, u, v, c, -u, 0, w, d, -v, -w, 0}); }

  muelmat linsrc() const {  // :207-275
    if (not polarized) return muelmat(func_F(a));
    Numeric l0, l1, l2, l3;
    if (both_zero) {
      l0 = func_F(a);
      l1 = func_Fp(a);
      if (std::abs(a) < too_small) {
        l2 = 1.0 / 6.0 + a / 12.0;
        l3 = 1.0 / 24.0 + a / 60.0;
      } else {
        const Numeric ez = std::exp(a);
        const Numeric a2 = a * a;
        const Numeric a3 = a2 * a;
        l2               = 0.5 * (ez * (a2 - 2.0 * a + 2.0) - 2.0) / a3;
        l3               = (ez * (a3 - 3.0 * a2 + 6.0 * a - 6.0) + 6.0) / (6.0 * a2 * a2);
      }
    } else {
      Numeric Pp, Pm_div_x;
      if (x_zero) {
        Pp       = func_F(a);
        Pm_div_x = func_Fp(a);
      } else {
        const Numeric f1 = func_F(a + x);
        const Numeric f2 = func_F(a - x);
        Pp               = 0.5 * (f1 + f2);
        Pm_div_x         = 0.5 * (f1 - f2) / x;
      }
      Numeric Qp, q_im;
      if (y_zero) {
        Qp   = func_F(a);
        q_im = func_Fp(a);
      } else {
        const Numeric denom       = a * a + y * y;
        const Numeric ea_cy_m1    = exp_a * cy - 1.0;
        const Numeric ea_sy       = exp_a * sy;
        Qp                        = (a * ea_cy_m1 + y * ea_sy) / denom;
        const Numeric sin_y_div_y = (std::abs(y) < 1e-6) ? 1.0 - y * y / 6.0 : std::sin(y) / y;
        q_im                      = (a * exp_a * sin_y_div_y - ea_cy_m1) / denom;
      }
      l2 = (Pp - Qp) * inv_x2y2;
      l0 = Pp - l2 * x2;
      l3 = (Pm_div_x - q_im) * inv_x2y2;
      l1 = Pm_div_x - l3 * x2;
    }
    const muelmat Sm  = S_mat();
    const muelmat S2m = Sm * Sm;
    const muelmat S3m = Sm * S2m;
    return muelmat(l0) + Sm * l1 + S2m * l2 + S3m * l3;
  }

  muelmat deriv(const muelmat& t, const propmat& k1, const propmat& k2, const propmat& dk, const Numeric r,
                const Numeric dr) const {  // :558-674
    const Numeric da = -0.5 * (r * dk.A() + dr * (k1.A() + k2.A()));
    if (not polarized) return muelmat(da * exp_a);
    const Numeric db  = -0.5 * (r * dk.B() + dr * (k1.B() + k2.B()));
    const Numeric dc  = -0.5 * (r * dk.C() + dr * (k1.C() + k2.C()));
    const Numeric dd  = -0.5 * (r * dk.D() + dr * (k1.D() + k2.D()));
    const Numeric du  = -0.5 * (r * dk.U() + dr * (k1.U() + k2.U()));
    const Numeric dv  = -0.5 * (r * dk.V() + dr * (k1.V() + k2.V()));
    const Numeric dw  = -0.5 * (r * dk.W() + dr * (k1.W() + k2.W()));
    const Numeric db2 = 2 * db * b;
    const Numeric dc2 = 2 * dc * c;
    const Numeric dd2 = 2 * dd * d;
    const Numeric du2 = 2 * du * u;
    const Numeric dv2 = 2 * dv * v;
    const Numeric dw2 = 2 * dw * w;
    const Numeric dB  = du2 + dv2 + dw2 - db2 - dc2 - dd2;
    const Numeric dC  = -2 * (d * u - c * v + b * w) * (dd * u + d * du - dc * v - c * dv + db * w + b * dw);
    const Numeric dS  = (B * dB - 2 * dC) / S;
    const Numeric dx2 = 0.25 * (dS - dB) / x2;
    const Numeric dy2 = 0.25 * (dS + dB) / y2;
    const Numeric dx  = 0.5 * dx2 / x;
    const Numeric dy  = 0.5 * dy2 / y;
    const Numeric dcy = -sy * dy;
    const Numeric dsy = cy * dy;
    const Numeric dcx = sx * dx;
    const Numeric dsx = cx * dx;
    const Numeric dix = -dx * ix * ix;
    const Numeric diy = -dy * iy * iy;
    const Numeric dx2dy2 = dx2 + dy2;
    const Numeric dC0 =
        either_zero ? 0.0 : (dcy * x2 + cy * dx2 + dcx * y2 + cx * dy2 - C0 * dx2dy2) * inv_x2y2;
    const Numeric dC1 = either_zero ? 0.0
                                    : (dsy * x2 * iy + sy * dx2 * iy + sy * x2 * diy + dsx * y2 * ix +
                                       sx * dy2 * ix + sx * y2 * dix - C1 * dx2dy2) *
                                          inv_x2y2;
    const Numeric dC2 =
        both_zero ? 0.0 : ((x_zero ? 0.0 : (dcx - C2 * dx2)) - (y_zero ? 0.0 : (dcy + C2 * dy2))) * inv_x2y2;
    const Numeric dC3 = both_zero ? 0.0
                                  : ((x_zero ? 0.0 : (dsx * ix + sx * dix - C3 * dx2)) -
                                     (y_zero ? 0.0 : (dsy * iy + sy * diy + C3 * dy2))) *
                                        inv_x2y2;
    const Numeric dC2b = dC2 * (c * u + d * v) + C2 * (dc * u + c * du + dd * v + d * dv);
    const Numeric dC2c = dC2 * (b * u - d * w) + C2 * (db * u + b * du - dd * w - d * dw);
    const Numeric dC2d = dC2 * (b * v + c * w) + C2 * (db * v + b * dv + dc * w + c * dw);
    const Numeric dC2u = dC2 * (b * c - v * w) + C2 * (db * c + b * dc - dv * w - v * dw);
    const Numeric dC2v = dC2 * (b * d + u * w) + C2 * (db * d + b * dd + du * w + u * dw);
    const Numeric dC2w = dC2 * (c * d - u * v) + C2 * (dc * d + c * dd - du * v - u * dv);
    const Numeric dC3b = dC3 * (b * (B - w2) + w * (c * v - d * u)) +
                         C3 * (db * (B - w2) + b * (dB - dw2) + dw * (c * v - d * u) +
                               w * (dc * v + c * dv - dd * u - d * du));
    const Numeric dC3c = dC3 * (c * (v2 - B) - v * (d * u + b * w)) +
                         C3 * (dc * (v2 - B) + c * (dv2 - dB) - dv * (d * u + b * w) -
                               v * (dd * u + d * du + db * w + b * dw));
    const Numeric dC3d = dC3 * (d * (u2 - B) - u * (c * v - b * w)) +
                         C3 * (dd * (u2 - B) + d * (du2 - dB) - du * (c * v - b * w) -
                               u * (dc * v + c * dv - db * w - b * dw));
    const Numeric dC3u = dC3 * (d * (c * v - b * w) - u * (B + d2)) +
                         C3 * (dd * (c * v - b * w) + d * (dc * v + c * dv - db * w - b * dw) - du * (B + d2) -
                               u * (dB + dd2));
    const Numeric dC3v = dC3 * (c * (d * u + b * w) - v * (B + c2)) +
                         C3 * (dc * (d * u + b * w) + c * (dd * u + d * du + db * w + b * dw) - dv * (B + c2) -
                               v * (dB + dc2));
    const Numeric dC3w = dC3 * (b * (c * v - d * u) - w * (B + b2)) +
                         C3 * (db * (c * v - d * u) + b * (dc * v + c * dv - dd * u - d * du) - dw * (B + b2) -
                               w * (dB + db2));
    const Numeric dM00 = dC0 + dC2 * (b2 + c2 + d2) + C2 * (db2 + dc2 + dd2);
    const Numeric dM11 = dC0 + dC2 * (b2 - u2 - v2) + C2 * (db2 - du2 - dv2);
    const Numeric dM22 = dC0 + dC2 * (c2 - u2 - w2) + C2 * (dc2 - du2 - dw2);
    const Numeric dM33 = dC0 + dC2 * (d2 - v2 - w2) + C2 * (dd2 - dv2 - dw2);
    return da * t +
           exp_a * make_mm({dM00, dC1 * b + C1 * db - dC2b - dC3b, dC1 * c + C1 * dc + dC2c + dC3c,
                            dC1 * d + C1 * dd + dC2d + dC3d, dC1 * b + C1 * db + dC2b - dC3b, dM11,
                            dC1 * u + C1 * du + dC2u + dC3u, dC1 * v + C1 * dv + dC2v + dC3v,
                            dC1 * c + C1 * dc - dC2c + dC3c, -dC1 * u - C1 * du + dC2u - dC3u, dM22,
                            dC1 * w + C1 * dw + dC2w + dC3w, dC1 * d + C1 * dd - dC2d + dC3d,
                            -dC1 * v - C1 * dv + dC2v - dC3v, -dC1 * w - C1 * dw + dC2w - dC3w, dM33});
  }

  muelmat linsrc_deriv(const propmat& dk, const Numeric r, const Numeric dr) const {  // :277-447
    const Numeric inv_r = (std::abs(r) > 1e-20) ? 1.0 / r : 0.0;
    const Numeric dr_r  = dr * inv_r;
    const Numeric da    = dr_r * a - 0.5 * r * dk.A();
    if (not polarized) return muelmat(func_Fp(a) * da);
    const Numeric db  = dr_r * b - 0.5 * r * dk.B();
    const Numeric dc  = dr_r * c - 0.5 * r * dk.C();
    const Numeric dd  = dr_r * d - 0.5 * r * dk.D();
    const Numeric du  = dr_r * u - 0.5 * r * dk.U();
    const Numeric dv  = dr_r * v - 0.5 * r * dk.V();
    const Numeric dw  = dr_r * w - 0.5 * r * dk.W();
    const Numeric db2 = 2.0 * db * b;
    const Numeric dc2 = 2.0 * dc * c;
    const Numeric dd2 = 2.0 * dd * d;
    const Numeric du2 = 2.0 * du * u;
    const Numeric dv2 = 2.0 * dv * v;
    const Numeric dw2 = 2.0 * dw * w;
    const Numeric dB  = du2 + dv2 + dw2 - db2 - dc2 - dd2;
    const Numeric dC  = -2.0 * (d * u - c * v + b * w) * (dd * u + d * du - dc * v - c * dv + db * w + b * dw);
    const Numeric dS_val = (S > 1e-9) ? (B * dB - 2.0 * dC) / S : 0.0;
    const Numeric dx2    = (x2 > 1e-9) ? 0.25 * (dS_val - dB) / x2 : 0.0;
    const Numeric dy2    = (y2 > 1e-9) ? 0.25 * (dS_val + dB) / y2 : 0.0;
    const Numeric dx     = (x > 1e-9) ? 0.5 * dx2 / x : 0.0;
    const Numeric dy     = (y > 1e-9) ? 0.5 * dy2 / y : 0.0;
    Numeric l1, l2, l3;
    Numeric dl0, dl1, dl2, dl3;
    if (both_zero) {
      const Numeric fpa  = func_Fp(a);
      const Numeric fppa = func_Fpp(a);
      dl0                = fpa * da;
      l1                 = fpa;
      dl1                = fppa * da;
      if (std::abs(a) < too_small) {
        l2  = 1.0 / 6.0 + a / 12.0;
        dl2 = da / 12.0;
        l3  = 1.0 / 24.0 + a / 60.0;
        dl3 = da / 60.0;
      } else {
        const Numeric f3pa = func_F3p(a);
        const Numeric f4pa = func_F4p(a);
        l2                 = 0.5 * fppa;
        dl2                = 0.5 * f3pa * da;
        l3                 = f3pa / 6.0;
        dl3                = f4pa / 6.0 * da;
      }
    } else {
      Numeric Pp = 0.0, Pm_div_x = 0.0, Qp = 0.0, q_im = 0.0;
      Numeric dPp = 0.0, dPm_div_x = 0.0, dQp = 0.0, dq_im = 0.0;
      if (x_zero) {
        Pp        = func_F(a);
        dPp       = func_Fp(a) * da + 0.5 * func_Fpp(a) * dx2;
        Pm_div_x  = func_Fp(a);
        dPm_div_x = func_Fpp(a) * da + (func_F3p(a) / 6.0) * dx2;
      } else {
        const Numeric f_apx   = func_F(a + x);
        const Numeric f_amx   = func_F(a - x);
        const Numeric fp_apx  = func_Fp(a + x);
        const Numeric fp_amx  = func_Fp(a - x);
        Pp                    = 0.5 * (f_apx + f_amx);
        const Numeric sum_fp  = fp_apx + fp_amx;
        const Numeric diff_fp = fp_apx - fp_amx;
        dPp                   = 0.5 * sum_fp * da + 0.5 * diff_fp * dx;
        Pm_div_x              = 0.5 * (f_apx - f_amx) / x;
        const Numeric dPm_da  = 0.5 * diff_fp / x;
        Numeric dPm_dx;
        if (x < 1e-3) {
          dPm_dx = func_F3p(a) * x / 3.0;
        } else {
          const Numeric diff_f = f_apx - f_amx;
          dPm_dx               = (x * sum_fp - diff_f) / (2.0 * x * x);
        }
        dPm_div_x = dPm_da * da + dPm_dx * dx;
      }
      if (y_zero) {
        Qp    = func_F(a);
        dQp   = func_Fp(a) * da - 0.5 * func_Fpp(a) * dy2;
        q_im  = func_Fp(a);
        dq_im = func_Fpp(a) * da - (func_F3p(a) / 6.0) * dy2;
      } else {
        const Numeric denom       = a * a + y * y;
        const Numeric ea_cy_m1    = exp_a * cy - 1.0;
        const Numeric ea_sy       = exp_a * sy;
        Qp                        = (a * ea_cy_m1 + y * ea_sy) / denom;
        const Numeric sin_y_div_y = (std::abs(y) < 1e-6) ? 1.0 - y * y / 6.0 : std::sin(y) / y;
        q_im                      = (a * exp_a * sin_y_div_y - ea_cy_m1) / denom;
        const Numeric ImF         = (a * exp_a * sy - y * ea_cy_m1) / denom;
        const Numeric A_val       = exp_a * ((a - 1.0) * cy - y * sy) + 1.0;
        const Numeric B_val       = exp_a * ((a - 1.0) * sy + y * cy);
        const Numeric C_val       = a * a - y * y;
        const Numeric D_val       = 2.0 * a * y;
        const Numeric denom2      = C_val * C_val + D_val * D_val;
        const Numeric Fp_re       = (A_val * C_val + B_val * D_val) / denom2;
        const Numeric Fp_im       = (B_val * C_val - A_val * D_val) / denom2;
        dQp                       = Fp_re * da - Fp_im * dy;
        const Numeric dImF        = Fp_im * da + Fp_re * dy;
        dq_im                     = (y * dImF - ImF * dy) / (y * y);
      }
      const Numeric inv  = inv_x2y2;
      const Numeric dinv = -inv * inv * (dx2 + dy2);
      l2                 = (Pp - Qp) * inv;
      dl2                = (dPp - dQp) * inv + (Pp - Qp) * dinv;
      dl0                = dPp - dl2 * x2 - l2 * dx2;
      l3                 = (Pm_div_x - q_im) * inv;
      dl3                = (dPm_div_x - dq_im) * inv + (Pm_div_x - q_im) * dinv;
      l1                 = Pm_div_x - l3 * x2;
      dl1                = dPm_div_x - dl3 * x2 - l3 * dx2;
    }
    const muelmat Sm   = S_mat();
    const muelmat dSm  = make_mm({0, db, dc, dd, db, 0, du, dv, dc, -du, 0, dw, dd, -dv, -dw, 0});
    const muelmat S2m  = Sm * Sm;
    const muelmat dS2m = dSm * Sm + Sm * dSm;
    const muelmat S3m  = Sm * S2m;
    const muelmat dS3m = dSm * S2m + Sm * dS2m;
    return muelmat(dl0) + Sm * dl1 + dSm * l1 + S2m * dl2 + dS2m * l2 + S3m * dl3 + dS3m * l3;
  }

  // tran::linsrc_linprop, :449-475: the unpolarised Dawson form and the polarised one through the complex matrix
  // square root of the absorption gradient, its inverse and the ELEMENT-WISE Dawson function of :467-474.
  muelmat linsrc_linprop(const muelmat& t, const propmat& k1, const propmat& k2, const Numeric r) const {
    const propmat alpha2 = (k2 - k1) / (2.0 * r);
    if (alpha2.A() < 1e-8) return linsrc();  // "Ignore when the gradient is negative"
    if (not polarized) {
      const Numeric alpha = std::sqrt(alpha2.A());
      const Numeric u0    = k1.A() / (2.0 * alpha);
      const Numeric u1    = k2.A() / (2.0 * alpha);
      return muelmat((Faddeeva::Dawson(u1) - t.m[0] * Faddeeva::Dawson(u0)) / (r * alpha));
    }
    const specmat alpha     = sqrt_pm(alpha2);
    const specmat alpha_inv = inv(alpha);
    const specmat u0        = alpha_inv * (k1 / 2.0);
    const specmat u1        = alpha_inv * (k2 / 2.0);
    muelmat o               = real(alpha_inv * (dawson(u1) - t * dawson(u0)));
    for (auto& x : o.m) x /= r;  // muelmat / Numeric is element-wise (rtepack_mueller_matrix.h:197-200)
    return o;
  }

  // tran::linsrc_linprop_deriv, :477-556: closed form for unpolarised layers, a forward perturbation of 1e-6 times
  // (dk, dr) for polarised ones ("These derivaties don't work so we use perturbations...", :543).  `exact` is the
  // oracle's own switch and is handed on to the perturbed tran.
  muelmat linsrc_linprop_deriv(const muelmat& lambda, const muelmat& t, const propmat& k1, const propmat& k2, const propmat& dk_in,
                               const muelmat& dt, const Numeric r, const Numeric dr, bool k1_deriv, bool exact) const {
    const Numeric alpha2A = ((k2 - k1) / (2.0 * r)).A();
    if (alpha2A < 1e-8) return linsrc_deriv(dk_in, r, dr);
    if (polarized) {
      constexpr Numeric eps = 1e-6;
      const propmat k1p = k1_deriv ? k1 + dk_in * eps : k1;
      const propmat k2p = k1_deriv ? k2 : k2 + dk_in * eps;
      const Numeric rp  = r + dr * eps;
      const tran tran_p{k1p, k2p, rp, exact};
      const muelmat tp      = tran_p();
      const muelmat lambdap = tran_p.linsrc_linprop(tp, k1p, k2p, rp);
      muelmat o;
      for (int i = 0; i < 16; i++) o.m[i] = (lambdap.m[i] - lambda.m[i]) / eps;
      return o;
    }
    const Numeric k1a = k1.A(), k2a = k2.A(), dk = dk_in.A();
    const Numeric delta_k = k2a - k1a;
    const Numeric denom   = 2.0 * r;
    const Numeric alpha   = std::sqrt(std::max(0.0, delta_k / denom));
    const Numeric u0 = k1a / (2.0 * alpha), u1 = k2a / (2.0 * alpha);
    const Numeric D0 = Faddeeva::Dawson(u0), D1 = Faddeeva::Dawson(u1);
    const Numeric dD0 = 1.0 - 2.0 * u0 * D0, dD1 = 1.0 - 2.0 * u1 * D1;
    const Numeric t00 = t.m[0], dt00 = dt.m[0];
    Numeric d_alpha, d_u0, d_u1;
    if (k1_deriv) {
      d_alpha = -0.5 * dk / (denom * alpha);
      d_u0    = (dk * 2.0 * alpha - k1a * 2.0 * d_alpha) / (4.0 * alpha * alpha);
      d_u1    = -k2a * d_alpha / (2.0 * alpha * alpha);
    } else {
      d_alpha = 0.5 * dk / (denom * alpha);
      d_u0    = -k1a * d_alpha / (2.0 * alpha * alpha);
      d_u1    = (dk * 2.0 * alpha - k2a * 2.0 * d_alpha) / (4.0 * alpha * alpha);
    }
    const Numeric d_num       = dD1 * d_u1 - dt00 * D0 - t00 * dD0 * d_u0 - t00 * D0 * d_u0;  // sic, :531-532
    const Numeric denom_val   = r * alpha;
    const Numeric d_denom_val = dr * alpha + r * d_alpha;
    return muelmat((d_num * denom_val - (D1 - t00 * D0) * d_denom_val) / (denom_val * denom_val));
  }
};

void store(double* p, const muelmat& m) { std::memcpy(p, m.m, 16 * sizeof(double)); }
muelmat load_mm(const double* p) {
  muelmat m;
  std::memcpy(m.m, p, 16 * sizeof(double));
  return m;
}
stokvec load_sv(const double* p) {
  stokvec s;
  std::memcpy(s.v, p, 4 * sizeof(double));
  return s;
}
void store(double* p, const stokvec& s) { std::memcpy(p, s.v, 4 * sizeof(double)); }

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
}  // namespace

// ===========================================================================
// exported oracle entry points (same argument meaning as include/arts_b200.h)
// ===========================================================================
// freq_grid_pathFromPath (src/m_ppvar.cc:47-77) for a whole path: with winds, [np][nf] shifted grids
// fac[ip] * freq_grid replace the caller's grid
struct PathGrid {
  std::vector<double> buf;
  const double* f;
  int64_t stride;
};
int path_grid(const ab200_catalog_desc& d, const ab200_atm_path& atm, int64_t nf, const double* f, int64_t stride,
              PathGrid& out) {
  if (!atm.wind) {
    out.f      = f;
    out.stride = stride;
    return 0;
  }
  out.buf.resize(static_cast<size_t>(atm.np) * nf);
  for (int ip = 0; ip < atm.np; ip++) {
    Numeric fac;
    if (not wind_factor(atm_at(d, atm, ip), fac)) return fail(AB200_ERR_INVALID, "Negative frequency scaling factor");
    for (int64_t i = 0; i < nf; i++) out.buf[static_cast<size_t>(ip) * nf + i] = fac * f[ip * stride + i];
  }
  out.f      = out.buf.data();
  out.stride = nf;
  return 0;
}

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }
int orc_num_threads(void) { return omp_get_max_threads(); }
// bench.py's CPU arm sets the team size explicitly: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers
int orc_set_num_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
}

// Faddeeva::w of the reference, vectorised (lbl_lineshape_voigt_lte.cpp:239)
int orc_faddeeva_w(int64_t n, const double* zr, const double* zi, double* wr, double* wi) {
#pragma omp parallel for
  for (int64_t i = 0; i < n; i++) {
    const Complex w = Faddeeva::w(Complex{zr[i], zi[i]}, 0);
    wr[i]           = w.real();
    wi[i]           = w.imag();
  }
  return 0;
}

int orc_wigner3j(int tj1, int tj2, int tj3, int tm1, int tm2, int tm3, double* out) {
  *out = wigner3j_2(tj1, tj2, tj3, tm1, tm2, tm3);
  return 0;
}

// zeeman sub-line table for one line and polarisation (tests compare the library's expansion)
int orc_zeeman_components(int on, double gu, double gl, int tJu, int tJl, int pol, int64_t cap, double* strength,
                          double* splitting) {
  const ZeemanView z{on != 0, gu, gl, tJu, tJl};
  const Index n = z.size(static_cast<Pol>(pol));
  for (Index i = 0; i < n and i < cap; i++) {
    strength[i]  = z.Strength(static_cast<Pol>(pol), i);
    splitting[i] = z.Splitting(static_cast<Pol>(pol), i);
  }
  return static_cast<int>(n);
}

int orc_norm_view(int pol, const double* mag, const double* los, double* npm) {
  norm_view(static_cast<Pol>(pol), mag, los, npm);
  return 0;
}

int orc_dnorm_view(int pol, int comp, const double* mag, const double* los, double* dnpm) {
  dnorm_view(static_cast<Pol>(pol), comp, mag, los, dnpm);
  return 0;
}

// spectral_propmatAddLines / spectral_propmat_pathFromPath with the lines-only
// agenda.  Parallelised like the reference: over levels (m_propmat.cc:42) when
// there are at least as many levels as threads, else over contiguous frequency
// chunks per level (m_lbl.cc:273-295 with omp_offset_count,
// matpack_mdspan_algorithm.cc:4-18).  The result does not depend on the choice.
// f, f_level_stride: the levels' grids after freq_grid_pathFromPath (path_grid above); atm still carries the winds
static int propmat_levels_on_path_grids(const ab200_catalog_desc* d, int64_t nf, const double* f, int64_t f_level_stride,
                                        const ab200_atm_path* atm, int32_t select_species, int32_t no_negative_absorption,
                                        int32_t nq, const ab200_target* targets, double* K, double* dK) {
  for (int ib = 0; ib < d->n_bands; ib++)
    if (d->band_lineshape[ib] != AB200_LINESHAPE_VP_LTE && d->band_lineshape[ib] != AB200_LINESHAPE_VP_LTE_MIRROR)
      return fail(AB200_ERR_UNSUPPORTED, "only VP_LTE and VP_LTE_MIRROR bands");
  if (nq > 0 && !dK) return fail(AB200_ERR_INVALID, "dK is null with nq > 0");
  for (int q = 0; q < nq; q++) {
    if (targets[q].kind == AB200_TARGET_P) return fail(AB200_ERR_UNSUPPORTED, "Not implemented, pressure derivative");  // :1482
    if (targets[q].kind == AB200_TARGET_ISORAT) {
      if (targets[q].species < 0 or targets[q].species >= d->n_isot) return fail(AB200_ERR_INVALID, "isotopologue out of range");
      for (int ip = 0; ip < atm->np; ip++)
        if (atm->isorat[static_cast<Index>(ip) * d->n_isot + targets[q].species] == 0)
          return fail(AB200_ERR_INVALID, "Does not support 0 for isotopologue ratios");  // :1539
    }
  }
  for (int q = 0; q < nq; q++) {
    if (targets[q].kind < AB200_TARGET_LINE_F0 or targets[q].kind > AB200_TARGET_LINE_LS) continue;
    if (targets[q].line < 0 or targets[q].line >= d->n_lines) return fail(AB200_ERR_INVALID, "line target out of range");
    const int64_t b = std::upper_bound(d->band_offset, d->band_offset + d->n_bands + 1, targets[q].line) - d->band_offset - 1;
    // with a cutoff the reference indexes the window's sub-span with whole-band indices (:723-739): nothing to restate
    if (d->band_cutoff_type[b] != AB200_CUTOFF_NONE or d->band_lineshape[b] != AB200_LINESHAPE_VP_LTE)
      return fail(AB200_ERR_UNSUPPORTED, "line targets need a VP_LTE band without cutoff");
  }
  const int np       = atm->np;
  const int nthreads = omp_get_max_threads();
  if (np >= nthreads || nf < nthreads) {
#pragma omp parallel for schedule(dynamic)
    for (int ip = 0; ip < np; ip++) {
      const AtmPt a = atm_at(*d, *atm, ip);
      lbl_calculate(K + static_cast<Index>(ip) * nf * 7, nq ? dK + static_cast<Index>(ip) * nq * nf * 7 : nullptr,
                    nf, f + ip * f_level_stride, 0, nf, *d, a, select_species, nq, targets,
                    no_negative_absorption != 0);
    }
  } else {
    for (int ip = 0; ip < np; ip++) {
      const AtmPt a = atm_at(*d, *atm, ip);
      // omp_offset_count: dn = nf / n, last chunk takes the remainder
      const Index n  = nthreads;
      const Index dn = nf / n;
#pragma omp parallel for
      for (Index i = 0; i < n; i++) {
        const Index lo = i * dn;
        const Index cn = (i == n - 1) ? nf - lo : dn;
        lbl_calculate(K + static_cast<Index>(ip) * nf * 7,
                      nq ? dK + static_cast<Index>(ip) * nq * nf * 7 : nullptr, nf, f + ip * f_level_stride, lo,
                      cn, *d, a, select_species, nq, targets, no_negative_absorption != 0);
      }
    }
  }
  // spectral_propmat_jacWindFix, src/m_frequency_grid.cc:106-182, on the level's (shifted) grid with the level's
  // freq_wind_shift_jac (freq_grid_pathFromPath runs wind_shift at every point, calm ones included)
  for (int q = 0; q < nq; q++) {
    const int kind = targets[q].kind;
    if (kind < AB200_TARGET_WIND_U or kind > AB200_TARGET_WIND_W) continue;
    for (int ip = 0; ip < np; ip++) {
      Numeric fac, jac[3];
      if (not wind_factor(atm_at(*d, *atm, ip), fac, jac)) return fail(AB200_ERR_INVALID, "Negative frequency scaling factor");
      const Numeric df_dx = jac[kind - AB200_TARGET_WIND_U];
      double* row         = dK + (static_cast<Index>(ip) * nq + q) * nf * 7;
      const double* fl    = f + ip * f_level_stride;
      for (int64_t i = 0; i < nf; i++)
        for (int c = 0; c < 7; c++) row[i * 7 + c] = row[i * 7 + c] * fl[i] * df_dx;
    }
  }
  return 0;
}

int orc_propmat_levels(const ab200_catalog_desc* d, int64_t nf, const double* f_in, int64_t f_level_stride_in,
                       const ab200_atm_path* atm, int32_t select_species, int32_t no_negative_absorption,
                       int32_t nq, const ab200_target* targets, double* K, double* dK) {
  if (!d || !atm || !f_in || !K) return fail(AB200_ERR_INVALID, "null argument");
  PathGrid pg;
  if (int rc = path_grid(*d, *atm, nf, f_in, f_level_stride_in, pg)) return rc;
  return propmat_levels_on_path_grids(d, nf, pg.f, pg.stride, atm, select_species, no_negative_absorption, nq, targets, K,
                                      dK);
}

// wind_shift alone (tests): fac and freq_wind_shift_jac of one path point
int orc_wind_shift(const double* wind, const double* los, double* fac, double* jac) {
  AtmPt a{};
  for (int i = 0; i < 3; i++) a.wind[i] = wind[i];
  a.los[0] = los[0];
  a.los[1] = los[1];
  Numeric fc;
  if (not wind_factor(a, fc, jac)) return fail(AB200_ERR_INVALID, "Negative frequency scaling factor");
  *fac = fc;
  return 0;
}

// TransmittanceMatrix::init, rtepack_transmission.cc:1254-1328 with constant
// :1114-1149 and linsrc :1151-1193.
int orc_tramat(int32_t np, int64_t nf, int32_t nq, const double* K, const double* dK, const double* r,
               const double* dr, int32_t rte_option, uint32_t flags, double* T, double* L, double* P, double* dT,
               double* dL) {
  if (rte_option != AB200_RTE_CONSTANT && rte_option != AB200_RTE_LINSRC && rte_option != AB200_RTE_LINPROP)
    return fail(AB200_ERR_INVALID, "unknown rte_option");
  const bool exact   = flags & AB200_FLAG_TRAN_EXACT;
  const bool linprop = rte_option == AB200_RTE_LINPROP;
  const bool linsrc  = rte_option == AB200_RTE_LINSRC || linprop;  // L, dL exist for both (:1300-1306)
  const muelmat id;
  const muelmat zero = muelmat::zero();
  // :1300-1314 identity / zero init
#pragma omp parallel for
  for (Index iv = 0; iv < nf; iv++) {
    for (int i = 0; i < np; i++) {
      store(T + (iv * np + i) * 16, id);
      if (linsrc && L) store(L + (iv * np + i) * 16, id);
      for (int t = 0; t < 2; t++)
        for (int j = 0; j < nq; j++) {
          store(dT + (((static_cast<Index>(t) * nf + iv) * np + i) * nq + j) * 16, zero);
          if (linsrc && dL) store(dL + (((static_cast<Index>(t) * nf + iv) * np + i) * nq + j) * 16, zero);
        }
    }
  }
  auto dK_at = [&](int i, int j, Index iv) { return load_pm(dK + ((static_cast<Index>(i) * nq + j) * nf + iv) * 7); };
  auto dXi   = [&](double* base, int t, Index iv, int i, int j) {
    return base + (((static_cast<Index>(t) * nf + iv) * np + i) * nq + j) * 16;
  };
#pragma omp parallel for collapse(2)
  for (int i = 1; i < np; i++) {
    for (Index iv = 0; iv < nf; ++iv) {
      const propmat k1 = load_pm(K + (static_cast<Index>(i - 1) * nf + iv) * 7);
      const propmat k2 = load_pm(K + (static_cast<Index>(i) * nf + iv) * 7);
      const tran ts{k1, k2, r[i - 1], exact};
      const muelmat Tm = ts();
      store(T + (iv * np + i) * 16, Tm);
      muelmat Lm;
      if (linprop) Lm = ts.linsrc_linprop(Tm, k1, k2, r[i - 1]);
      else if (linsrc) Lm = ts.linsrc();
      if (linsrc) store(L + (iv * np + i) * 16, Lm);
      for (int j = 0; j < nq; j++) {
        const Numeric dr0 = dr[(0 * (np - 1) + (i - 1)) * nq + j];
        const Numeric dr1 = dr[(1 * (np - 1) + (i - 1)) * nq + j];
        const muelmat dT0m = ts.deriv(Tm, k1, k2, dK_at(i - 1, j, iv), r[i - 1], dr0);
        const muelmat dT1m = ts.deriv(Tm, k1, k2, dK_at(i, j, iv), r[i - 1], dr1);
        store(dXi(dT, 0, iv, i - 1, j), dT0m);
        store(dXi(dT, 1, iv, i, j), dT1m);
        if (linprop) {  // TransmittanceMatrix::linprop :1225-1247; note dr1 in BOTH calls (:1238)
          store(dXi(dL, 0, iv, i - 1, j),
                ts.linsrc_linprop_deriv(Lm, Tm, k1, k2, dK_at(i - 1, j, iv), dT0m, r[i - 1], dr1, true, exact));
          store(dXi(dL, 1, iv, i, j), ts.linsrc_linprop_deriv(Lm, Tm, k1, k2, dK_at(i, j, iv), dT1m, r[i - 1], dr1, false, exact));
        } else if (linsrc) {
          store(dXi(dL, 0, iv, i - 1, j), ts.linsrc_deriv(dK_at(i - 1, j, iv), r[i - 1], dr0));
          store(dXi(dL, 1, iv, i, j), ts.linsrc_deriv(dK_at(i, j, iv), r[i - 1], dr1));
        }
      }
    }
  }
  // :1322-1327 cumulative transmission
#pragma omp parallel for
  for (Index i = 0; i < nf; i++) {
    muelmat acc;
    store(P + (i * np + 0) * 16, acc);
    for (int j = 1; j < np; j++) {
      acc = acc * load_mm(T + (i * np + j) * 16);
      store(P + (i * np + j) * 16, acc);
    }
  }
  return 0;
}

// SourceVector::init (spectral), rtepack_source.cc:52-105, in LTE (nlte = 0, dnlte = 0):
// J = B(f,T) (+ K^-1 * 0), dJ[k] = [it==k ? dB/dT : 0, 0,0,0] - K^-1 (dK * 0 - 0).
int orc_srcvec(int32_t np, int64_t nf, int32_t nq, const double* K, const double* f, int64_t f_level_stride,
               const double* T_level, int32_t it, double* J, double* dJ) {
#pragma omp parallel for collapse(2)
  for (int i = 0; i < np; i++) {
    for (Index j = 0; j < nf; j++) {
      const propmat k = load_pm(K + (static_cast<Index>(i) * nf + j) * 7);
      stokvec Jv;
      const bool rot = k.is_rotational();
      if (not rot) Jv.v[0] = planck(f[i * f_level_stride + j], T_level[i]);
      store(J + (j * np + i) * 4, Jv);
      for (int q = 0; q < nq; q++) {
        stokvec d;
        if (not rot and it == q) d.v[0] = dplanck_dt(f[i * f_level_stride + j], T_level[i]);
        store(dJ + ((j * np + i) * nq + q) * 4, d);
      }
    }
  }
  return 0;
}

// rte_emission, rtepack_rtestep.cc:265-404 (+ the background copy and zero
// init of m_spectral_radiance.cc:36-40)
int orc_rte_emission(int32_t rte_option, int32_t np, int64_t nf, int32_t nq, const double* T, const double* L,
                     const double* P, const double* dT, const double* dL, const double* J, const double* dJ,
                     const double* I_bkg, double* I, double* dI) {
  if (rte_option != AB200_RTE_CONSTANT && rte_option != AB200_RTE_LINSRC && rte_option != AB200_RTE_LINPROP)
    return fail(AB200_ERR_INVALID, "unknown rte_option");
  if (rte_option == AB200_RTE_LINPROP) rte_option = AB200_RTE_LINSRC;  // rte_emission: linsrc and linprop share linevo (:392-401)
  auto dXi = [&](const double* base, int t, Index iv, int i, int j) {
    return load_mm(base + (((static_cast<Index>(t) * nf + iv) * np + i) * nq + j) * 16);
  };
#pragma omp parallel for
  for (Index iv = 0; iv < nf; iv++) {
    stokvec Iv = load_sv(I_bkg + iv * 4);
    std::vector<stokvec> dIv(static_cast<size_t>(np) * nq);
    auto Jv  = [&](int i) { return load_sv(J + (iv * np + i) * 4); };
    auto dJv = [&](int i, int q) { return load_sv(dJ + ((iv * np + i) * nq + q) * 4); };
    for (int i = np - 2; i >= 0; i--) {
      const muelmat Tm = load_mm(T + (iv * np + i + 1) * 16);
      if (rte_option == AB200_RTE_CONSTANT) {  // :287-309
        const stokvec Jm = avg(Jv(i), Jv(i + 1));
        Iv               = Iv - Jm;
        if (nq) {
          const muelmat Pm = load_mm(P + (iv * np + i) * 16);
          for (int iq = 0; iq < nq; iq++) {
            const stokvec dJ0 = dJv(i, iq), dJ1 = dJv(i + 1, iq);
            dIv[i * nq + iq]       = dIv[i * nq + iq] + Pm * (dXi(dT, 0, iv, i, iq) * Iv + avg(dJ0, -(Tm * dJ0)));
            dIv[(i + 1) * nq + iq] = dIv[(i + 1) * nq + iq] + Pm * (dXi(dT, 1, iv, i + 1, iq) * Iv + avg(dJ1, -(Tm * dJ1)));
          }
        }
        Iv = Tm * Iv + Jm;
      } else {  // linevo :341-370
        const muelmat Lm    = load_mm(L + (iv * np + i + 1) * 16);
        const stokvec J0    = Jv(i + 1);
        const stokvec J1    = Jv(i);
        const stokvec ImJ0  = Iv - J0;
        const stokvec J0mJ1 = J0 - J1;
        if (nq) {
          const muelmat Pm = load_mm(P + (iv * np + i) * 16);
          for (int iq = 0; iq < nq; iq++) {
            const stokvec dJ0 = dJv(i, iq), dJ1 = dJv(i + 1, iq);
            dIv[i * nq + iq] = dIv[i * nq + iq] + Pm * (dJ1 - Lm * dJ0 + dXi(dT, 0, iv, i, iq) * ImJ0 +
                                                        dXi(dL, 0, iv, i, iq) * J0mJ1);
            dIv[(i + 1) * nq + iq] =
                dIv[(i + 1) * nq + iq] + Pm * (dXi(dT, 1, iv, i + 1, iq) * ImJ0 + dXi(dL, 1, iv, i + 1, iq) * J0mJ1 +
                                               Lm * dJ1 - Tm * dJ0);
          }
        }
        Iv = Tm * ImJ0 + Lm * J0mJ1 + J1;
      }
    }
    store(I + iv * 4, Iv);
    for (int i = 0; i < np; i++)
      for (int q = 0; q < nq; q++) store(dI + ((iv * np + i) * nq + q) * 4, dIv[i * nq + q]);
  }
  return 0;
}

// rte_transmission, rtepack_rtestep.cc:456-503 (spectral_radCumulativeTransmission, m_spectral_radiance.cc:49-74).
// Forward part literal: I = P[iv][N-1] * I0 (:469-470).  Jacobian: the reference's loop (:476-491) reads the layer
// transmittance as Ts[i + 1][iv] - frequency and level swapped, out of bounds unless nf == np - and adds the
// dTs[1, iv, i] term to level i - 1 instead of i, so it has no defined output to restate.  What is restated here is
// its evident intent, the exact product rule of I = T_1 ... T_{N-1} I0 in the reference's own conventions
// (dT[0][i] = dT_{i+1}/dx_i, dT[1][i] = dT_i/dx_i; suffix product P, prefix Pi):
//   dI[i] += Pi[i] dT[0][i] (T_{i+2} ... T_{N-1}) I0  +  Pi[i-1] dT[1][i] (T_{i+1} ... T_{N-1}) I0
// pinned by tests/test_oracle_pins.py against perturbed forward runs and against rte_emission with J = 0.
int orc_rte_transmission(int32_t np, int64_t nf, int32_t nq, const double* T, const double* P, const double* dT,
                         const double* I_bkg, double* I, double* dI) {
  if (np == 0) return 0;
  auto dXi = [&](int t, Index iv, int i, int j) {
    return load_mm(dT + (((static_cast<Index>(t) * nf + iv) * np + i) * nq + j) * 16);
  };
#pragma omp parallel for
  for (Index iv = 0; iv < nf; iv++) {
    const stokvec src = load_sv(I_bkg + iv * 4);
    store(I + iv * 4, load_mm(P + (iv * np + (np - 1)) * 16) * src);
    if (nq == 0) continue;
    std::vector<stokvec> dIv(static_cast<size_t>(np) * nq);
    muelmat Sfx = 1.0;  // T_{i+2} ... T_{N-1}
    for (int i = np - 2; i >= 0; i--) {
      const muelmat R = load_mm(T + (iv * np + i + 1) * 16) * Sfx;  // T_{i+1} ... T_{N-1}
      const muelmat Pm = load_mm(P + (iv * np + i) * 16);
      for (int iq = 0; iq < nq; iq++) {
        dIv[i * nq + iq]       = dIv[i * nq + iq] + Pm * (dXi(0, iv, i, iq) * (Sfx * src));
        dIv[(i + 1) * nq + iq] = dIv[(i + 1) * nq + iq] + Pm * (dXi(1, iv, i + 1, iq) * (Sfx * src));
      }
      Sfx = R;
    }
    for (int i = 0; i < np; i++)
      for (int q = 0; q < nq; q++) store(dI + ((iv * np + i) * nq + q) * 4, dIv[i * nq + q]);
  }
  return 0;
}

// The canonical sequence of spectral_radClearskyEmission
// (workspace_meta_methods.cpp:166-181) from spectral_propmat_pathFromPath on,
// un-fused exactly like the reference (T, L, P, dT, dL, J, dJ materialised).
int orc_clearsky_emission(const ab200_catalog_desc* d, int64_t nf, const double* f, int64_t f_level_stride,
                          const ab200_atm_path* atm, int32_t select_species, int32_t no_negative_absorption,
                          int32_t nq, const ab200_target* targets, const double* r, int32_t hse_derivative,
                          int32_t rte_option, const double* I_bkg, uint32_t flags, double* I, double* dI,
                          double* K_out) {
  const int np = atm->np;
  std::vector<double> K(static_cast<size_t>(np) * nf * 7, 0.0), dK(static_cast<size_t>(np) * nq * nf * 7, 0.0);
  PathGrid pg;
  if (int rcg = path_grid(*d, *atm, nf, f, f_level_stride, pg)) return rcg;
  f              = pg.f;  // the grids below are already shifted
  f_level_stride = pg.stride;
  int rc = propmat_levels_on_path_grids(d, nf, f, f_level_stride, atm, select_species, no_negative_absorption, nq,
                                        targets, K.data(), dK.data());
  if (rc) return rc;
  // m_tramat.cc:14-24
  int it = -1;
  for (int q = 0; q < nq; q++)
    if (targets[q].kind == AB200_TARGET_T) it = q;
  std::vector<double> dr(static_cast<size_t>(2) * (np - 1) * nq, 0.0);
  if (hse_derivative and it >= 0) {
    for (int ip = 0; ip < np - 1; ip++) {
      dr[(0 * (np - 1) + ip) * nq + it] = r[ip] / (2.0 * atm->T[ip]);
      dr[(1 * (np - 1) + ip) * nq + it] = r[ip] / (2.0 * atm->T[ip + 1]);
    }
  }
  const size_t nm = static_cast<size_t>(nf) * np * 16;
  std::vector<double> T(nm), L(nm), P(nm), dT(2 * nm * nq), dL(2 * nm * nq);
  rc = orc_tramat(np, nf, nq, K.data(), dK.data(), r, dr.data(), rte_option, flags, T.data(), L.data(), P.data(),
                  dT.data(), dL.data());
  if (rc) return rc;
  std::vector<double> J(static_cast<size_t>(nf) * np * 4), dJ(static_cast<size_t>(nf) * np * nq * 4);
  rc = orc_srcvec(np, nf, nq, K.data(), f, f_level_stride, atm->T, it, J.data(), dJ.data());
  if (rc) return rc;
  rc = orc_rte_emission(rte_option, np, nf, nq, T.data(), L.data(), P.data(), dT.data(), dL.data(), J.data(),
                        dJ.data(), I_bkg, I, dI);
  if (rc) return rc;
  if (K_out) std::memcpy(K_out, K.data(), K.size() * sizeof(double));
  return 0;
}

// spectral_planck_op, spectral_radiance_transform_operator.cc:46-87 (forward part)
int orc_planck_tb(int64_t nf, const double* f, double* I) {
  for (int64_t j = 0; j < nf; j++) {
    double* v        = I + 4 * j;
    const Numeric fj = f[j];
    const Numeric n0 = invplanck(v[0], fj);
    const Numeric n1 = invplanck(0.5 * (v[0] + v[1]), fj) - invplanck(0.5 * (v[0] - v[1]), fj);
    const Numeric n2 = invplanck(0.5 * (v[0] + v[2]), fj) - invplanck(0.5 * (v[0] - v[2]), fj);
    const Numeric n3 = invplanck(0.5 * (v[0] + v[3]), fj) - invplanck(0.5 * (v[0] - v[3]), fj);
    v[0] = n0;
    v[1] = n1;
    v[2] = n2;
    v[3] = n3;
  }
  return 0;
}

int orc_planck(int64_t n, const double* f, double T, double* B) {
  for (int64_t i = 0; i < n; i++) B[i] = planck(f[i], T);
  return 0;
}

// ---------------------------------------------------------------------------
// observer epilogue: the host glue around the path
// ---------------------------------------------------------------------------
// physics_funcs.cc:76-83
static Numeric dinvplanckdI(Numeric i, Numeric f) {
  constexpr Numeric a = Constant::h / Constant::k;
  constexpr Numeric b = 2 * Constant::h / (Constant::c * Constant::c);
  const Numeric d     = b * f * f * f / i;
  const Numeric binv  = a * f / std::log1p(d);
  return binv * binv / (a * f * i * (1 / d + 1));
}
// physics_funcs.cc:172-176
static Numeric invrayjean(Numeric i, Numeric f) {
  constexpr Numeric a = Constant::c * Constant::c / (2 * Constant::k);
  return (a * i) / (f * f);
}

// from_temp, m_background.cc:55-63 (spectral_radSurfaceBlackbody :113-141, spectral_radUniformCosmicBackground :65-72);
// dB [nf] = dplanck_dt(f, T) of the surface-temperature Jacobian (:131-138)
int orc_background(int64_t nf, const double* f, double T, double* I_bkg, double* dB) {
  for (int64_t j = 0; j < nf; j++) {
    I_bkg[4 * j] = planck(f[j], T);
    I_bkg[4 * j + 1] = I_bkg[4 * j + 2] = I_bkg[4 * j + 3] = 0.0;
    if (dB) dB[j] = dplanck_dt(f[j], T);
  }
  return 0;
}

// Steps 2-5 of the observer epilogue on host arrays.
//   P  [nf][np][16] cumulative transmittance of the path (TransmittanceMatrix::P)
//   I  [nf][4] in: spectral_rad, out: transformed;  dI [nf][np][nq][4] spectral_rad_jac_path
//   Jx [nx][nf][4] out;  y [n_channels], Jy [n_channels][nx] out (this path's contribution)
int orc_observer(int32_t np, int64_t nf, int32_t nq, const double* f, const ab200_observer* o, const double* P,
                 double* I, const double* dI, double* Jx, double* y, double* Jy) {
  const Index nx = o->nx;
  std::vector<double> scratch;
  if (!Jx) {
    scratch.assign(static_cast<size_t>(nx) * nf * 4, 0.0);
    Jx = scratch.data();
  }
  auto jx = [&](Index i, Index j) { return Jx + (i * nf + j) * 4; };
  // spectral_radSurfaceBlackbody's spectral_rad_jac, m_background.cc:126-140
  for (Index k = 0; k < nx * nf * 4; k++) Jx[k] = 0.0;
  if (o->bkg_kind == AB200_BKG_PLANCK)
    for (int b = 0; b < o->n_bkg; b++)
      for (Index j = 0; j < nf; j++) jx(o->bkg_x[b], j)[0] += o->bkg_w[b] * dplanck_dt(f[j], o->bkg_T);
  // spectral_rad_jacFromBackground, m_rad.cc:26-60
  if (nq > 0 || o->n_bkg > 0)
    for (Index i = 0; i < nx; i++)
      for (Index j = 0; j < nf; j++) {
        const stokvec r = load_mm(P + (j * np + (np - 1)) * 16) * load_sv(jx(i, j));
        std::memcpy(jx(i, j), r.v, sizeof(r.v));
      }
  // spectral_rad_jacAddPathPropagation, m_rad.cc:62-127 (targets outermost, then path points, then weights)
  for (int t = 0; t < nq; t++)
    for (int ip = 0; ip < np; ip++)
      for (int64_t e = o->map_offset[ip * nq + t]; e < o->map_offset[ip * nq + t + 1]; e++) {
        const Numeric w = o->map_w[e];
        if (w == 0.0) continue;
        for (Index j = 0; j < nf; j++) {
          const double* a = dI + ((j * np + ip) * nq + t) * 4;
          double* b       = jx(o->map_x[e], j);
          for (int c = 0; c < 4; c++) b[c] = std::fma(w, a[c], b[c]);
        }
      }
  // spectral_rad_transform_operator, spectral_radiance_transform_operator.cc:8-122
  const Numeric n2 = o->n_real * o->n_real;
  for (Index j = 0; j < nf; j++) {
    double* v = I + 4 * j;
    Numeric dv[4] = {1, 1, 1, 1};
    switch (o->unit) {
      case AB200_UNIT_UNIT:
        if (o->n_real != 1.0) {
          for (int c = 0; c < 4; c++) { v[c] *= n2; dv[c] = n2; }
        }
        break;
      case AB200_UNIT_RJBT: {
        const Numeric df = invrayjean(1.0, f[j]);
        for (int c = 0; c < 4; c++) { v[c] *= df; dv[c] = df; }
      } break;
      case AB200_UNIT_PLANCKBT: {
        const Numeric fj = f[j];
        dv[0] = dinvplanckdI(v[0], fj);
        for (int c = 1; c < 4; c++) dv[c] = dinvplanckdI(0.5 * (v[0] + v[c]), fj) - dinvplanckdI(0.5 * (v[0] - v[c]), fj);
        Numeric n[4];
        n[0] = invplanck(v[0], fj);
        for (int c = 1; c < 4; c++) n[c] = invplanck(0.5 * (v[0] + v[c]), fj) - invplanck(0.5 * (v[0] - v[c]), fj);
        for (int c = 0; c < 4; c++) v[c] = n[c];
      } break;
      case AB200_UNIT_W_M2_M_SR: {
        const Numeric df = (f[j] * (f[j] / Constant::c));
        for (int c = 0; c < 4; c++) { v[c] *= df * n2; dv[c] = df * n2; }
      } break;
      case AB200_UNIT_W_M2_M1_SR:
        for (int c = 0; c < 4; c++) { v[c] *= n2 * Constant::c; dv[c] = n2 * Constant::c; }
        break;
      default: return fail(AB200_ERR_INVALID, "unknown spectral radiance unit");
    }
    const bool scale = !(o->unit == AB200_UNIT_UNIT && o->n_real == 1.0);
    if (scale)
      for (Index i = 0; i < nx; i++)
        for (int c = 0; c < 4; c++) jx(i, j)[c] *= dv[c];
  }
  // SensorObsel::sumup, obsel.cpp:246-279 (this path's poslos row; entries sorted by icol)
  for (int ch = 0; ch < o->n_channels; ch++) {
    Numeric sum = 0.0;
    for (int64_t e = o->w_offset[ch]; e < o->w_offset[ch + 1]; e++) {
      const double* a = I + 4 * o->w_freq[e];
      const double* w = o->w_stokes + 4 * e;
      sum += a[0] * w[0] + a[1] * w[1] + a[2] * w[2] + a[3] * w[3];  // dot = transform_reduce, matpack_mdspan_helpers_reduce.h:218-222
    }
    if (y) y[ch] = sum;
    if (Jy)
      for (Index i = 0; i < nx; i++) {
        Numeric s = 0.0;
        for (int64_t e = o->w_offset[ch]; e < o->w_offset[ch + 1]; e++) {
          const double* a = jx(i, o->w_freq[e]);
          const double* w = o->w_stokes + 4 * e;
          s += a[0] * w[0] + a[1] * w[1] + a[2] * w[2] + a[3] * w[3];
        }
        Jy[static_cast<Index>(ch) * nx + i] = s;
      }
  }
  return 0;
}

// scalar entry points of the unit conversions (pins against their analytic inverses in tests/test_oracle_pins.py)
double orc_invplanck(double i, double f) { return invplanck(i, f); }
double orc_dinvplanckdI(double i, double f) { return dinvplanckdI(i, f); }
double orc_invrayjean(double i, double f) { return invrayjean(i, f); }
double orc_dplanck_dt(double f, double t) { return dplanck_dt(f, t); }

// ---------------------------------------------------------------------------
// collision-induced absorption
// ---------------------------------------------------------------------------
namespace cia {
// lagrange_interp::update_pos for the identity transform on an ascending grid (lagrange_interp.h:160-248): the stencil of
// P = order + 1 points starts at clamp(xp, xf, xe) - Of, xp = the last index whose right neighbour is not below x
Index start_index(const double* xi, Index n, Index order, Numeric x) {
  const Index P = order + 1;
  if (n <= P) return 0;
  const Index Of = order / 2;
  const Index xf = Of, xe = n - P / 2 - 1;
  Index xp = xf;
  while (xp < xe and xi[xp + 1] < x) ++xp;
  while (xp > xf and xi[xp] > x) --xp;
  return std::clamp(xp, xf, xe) - xf;
}
// set_weights, lagrange_interp.h:300-440 (non-cyclic branch): the last weight is one minus the others
void weights(double* w, const double* xi, Index i0, Index order, Numeric x) {
  for (Index j = 0; j < order; j++) {
    const Numeric xj = xi[i0 + j];
    Numeric numer = 1.0, denom = 1.0;
    for (Index k = 0; k < order; k++) {
      const Index m = i0 + k + (k >= j);
      numer *= x - xi[m];
      denom *= xj - xi[m];
    }
    w[j] = numer / denom;
  }
  w[order] = 1.0;
  for (Index j = 0; j < order; j++) w[order] -= w[j];
}
// check_limit, lagrange_interp.h:572-650 (ascending, not cyclic): false = outside the extrapolation range
bool in_limits(const double* xi, Index n, Index order, Numeric limit, Numeric xmin, Numeric xmax) {
  if (order == 0 or limit <= 0.0) return true;
  const Numeric hi = xi[n - 1] + limit * (xi[n - 1] - xi[n - 2]);
  const Numeric lo = xi[0] - limit * (xi[1] - xi[0]);
  return not(hi < xmax or lo > xmin);
}
// cia_interpolation, src/core/absorption/cia.cc:76-190.  Returns false for the exception path (temperature outside the
// extrapolation range): with `robust` the reference then fills the result with NaN, otherwise it throws.
bool interpolate(double* result, const double* f, Index nf, Numeric T, const ab200_cia_dataset& ds, Numeric T_extrapolfac) {
  for (Index i = 0; i < nf; i++) result[i] = 0;
  Index i_fstart = 0, i_fstop = nf - 1;
  for (; i_fstart < nf; ++i_fstart)
    if (f[i_fstart] >= ds.f_grid[0]) break;
  if (i_fstart == nf) return true;
  for (; i_fstop >= 0; --i_fstop)
    if (f[i_fstop] <= ds.f_grid[ds.nf - 1]) break;
  if (i_fstop == -1) return true;
  if (i_fstop - i_fstart + 1 < 1) return true;
  constexpr Index f_order = 3;
  const Index T_order = std::min<Index>(3, ds.nT - 1);
  // make_lags: check_limit first (frequencies are inside the grid by construction; the temperature may not be)
  if (not in_limits(ds.T_grid, ds.nT, T_order, T_extrapolfac, T, T)) return false;
  double wT[4] = {1, 0, 0, 0};
  const Index iT = start_index(ds.T_grid, ds.nT, T_order, T);
  if (T_order > 0) weights(wT, ds.T_grid, iT, T_order, T);
  for (Index i = i_fstart; i <= i_fstop; i++) {
    double wf[4];
    const Index i0 = start_index(ds.f_grid, ds.nf, f_order, f[i]);
    weights(wf, ds.f_grid, i0, f_order, f[i]);
    Numeric out = 0;  // interp, lagrange_interp.h:920-940: (field * wf) * wT, frequency index outermost
    for (Index a = 0; a <= f_order; a++) {
      if (T_order == 0) {
        out += ds.data[(i0 + a) * ds.nT + 0] * wf[a];
      } else {
        for (Index b = 0; b <= T_order; b++) out += ds.data[(i0 + a) * ds.nT + iT + b] * wf[a] * wT[b];
      }
    }
    result[i] = out < 0 ? 0 : out;  // :181-182
  }
  return true;
}
}  // namespace cia

// spectral_propmatAddCIA, src/m_cia.cc:27-178, for every level of a path; K [np][nf][7], dK [np][nq][nf][7] are +=.
int orc_cia_levels(const ab200_cia_record* records, int32_t n_records, int64_t nf, const double* f_in, int64_t f_level_stride,
                   const ab200_atm_path* atm, int32_t n_species, int32_t select_species, int32_t nq,
                   const ab200_target* targets, double dt, double T_extrapolfac, int32_t ignore_errors, double* K, double* dK) {
  const int np = atm->np;
  int it = -1;
  for (int q = 0; q < nq; q++)
    if (targets[q].kind == AB200_TARGET_T and it < 0) it = q;
  if (it >= 0 and not std::isnormal(dt)) return fail(AB200_ERR_INVALID, "dt must be >0 and not NaN or Inf");
  std::vector<double> xsec(nf), dxsec(nf), tmp(nf);
  for (int ip = 0; ip < np; ip++) {
    const double* f = f_in + ip * f_level_stride;
    const Numeric T = atm->T[ip], P = atm->P[ip];
    if (T <= 0) return fail(AB200_ERR_INVALID, "Non-positive temperature");
    if (P <= 0) return fail(AB200_ERR_INVALID, "Non-positive pressure");
    const double* vmr = atm->vmr + static_cast<Index>(ip) * n_species;
    for (int r = 0; r < n_records; r++) {
      const ab200_cia_record& rec = records[r];
      if (select_species != AB200_SPECIES_BATH and select_species != rec.species1) continue;
      const Numeric nd_sec = number_density(P, T) * vmr[rec.species2];
      // CIARecord::Extract, cia.cc:214-226
      auto extract = [&](std::vector<double>& res, Numeric temp) -> int {
        std::fill(res.begin(), res.end(), 0.0);
        for (int k = 0; k < rec.n_datasets; k++) {
          const bool ok = cia::interpolate(tmp.data(), f, nf, temp, rec.datasets[k], T_extrapolfac);
          if (not ok) {
            if (not ignore_errors) return fail(AB200_ERR_INVALID, "Problem with CIA species: temperature outside the extrapolation range of the data");
            std::fill(tmp.begin(), tmp.end(), std::numeric_limits<double>::quiet_NaN());
          }
          for (Index i = 0; i < nf; i++) res[i] += tmp[i];
        }
        return 0;
      };
      if (int rc = extract(xsec, T)) return rc;
      if (it >= 0)
        if (int rc = extract(dxsec, T + dt)) return rc;
      const Numeric nd = number_density(P, T), dnd_dt = dnumber_density_dt(P, T);
      const Numeric dnd_dt_sec = dnumber_density_dt(P, T) * vmr[rec.species2];
      for (Index iv = 0; iv < nf; iv++) {
        K[(static_cast<Index>(ip) * nf + iv) * 7] += nd_sec * xsec[iv] * nd * vmr[rec.species1];
        auto dk = [&](int q) -> double& { return dK[((static_cast<Index>(ip) * nq + q) * nf + iv) * 7]; };
        if (it >= 0)
          dk(it) += ((nd_sec * (dxsec[iv] - xsec[iv]) / dt + xsec[iv] * dnd_dt_sec) * nd + xsec[iv] * nd_sec * dnd_dt) *
                    vmr[rec.species1];
        // jac_targets.find(species): the first target of that species (:166-175); both lines add the same expression
        for (int sp : {rec.species1, rec.species2})
          for (int q = 0; q < nq; q++)
            if (targets[q].kind == AB200_TARGET_VMR and targets[q].species == sp) {
              dk(q) += nd_sec * xsec[iv] * nd;
              break;
            }
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------
// absorption lookup tables
// ---------------------------------------------------------------------------
namespace lut {
constexpr int MAXP = 16;  // stencil points supported (orders up to 15; the defaults are 7)
struct Lag {
  Index i0 = 0, order = 0;
  double w[MAXP];
};
// lagrange_interp: check_limit (lagrange_interp.h:572-650), update_pos fixed point (:160-248, both grid orders, nearest
// neighbour for order 0) and set_weights (:300-440).  Returns false with a message for the reference's exceptions.
bool make_lag(Lag& l, const double* xi, Index n, Index order, Numeric x, Numeric limit, const char* info, std::string& err) {
  const bool ascending = n <= 1 or xi[0] < xi[1];
  if (order >= n) {
    err = std::string("Error in check_limit for ") + info + ":\nToo few grid points for the given polynomial order";
    return false;
  }
  if (order > 0 and limit > 0.0) {
    const Numeric hi = ascending ? xi[n - 1] + limit * (xi[n - 1] - xi[n - 2]) : xi[0] + limit * (xi[0] - xi[1]);
    const Numeric lo = ascending ? xi[0] - limit * (xi[1] - xi[0]) : xi[n - 1] - limit * (xi[n - 2] - xi[n - 1]);
    if (hi < x or lo > x) {
      err = std::string("Error in check_limit for ") + info + ":\nExtrapolation limit yields limits to the extrapolation of the grid that are outside the input grid.";
      return false;
    }
  }
  const Index P = order + 1;
  l.order = order;
  if (n <= P) {
    l.i0 = 0;
  } else {
    const Index Of = order / 2, xf = Of, xe = n - P / 2 - 1;
    Index xp = xf;
    if (ascending) {
      while (xp < xe and xi[xp + 1] < x) ++xp;
      while (xp > xf and xi[xp] > x) --xp;
    } else {
      while (xp < xe and xi[xp + 1] > x) ++xp;
      while (xp > xf and xi[xp] < x) --xp;
    }
    if (order == 0) {
      const Index xn = xp + 1;
      xp = (xn == n or std::abs(x - xi[xn]) > std::abs(x - xi[xp])) ? xp : xn;
    }
    l.i0 = std::clamp(xp, xf, xe) - xf;
  }
  for (Index j = 0; j < order; j++) {
    const Numeric xj = xi[l.i0 + j];
    Numeric numer = 1.0, denom = 1.0;
    for (Index k = 0; k < order; k++) {
      const Index m = l.i0 + k + (k >= j);
      numer *= x - xi[m];
      denom *= xj - xi[m];
    }
    l.w[j] = numer / denom;
  }
  l.w[order] = 1.0;
  for (Index j = 0; j < order; j++) l.w[order] -= l.w[j];
  return true;
}
Numeric interp1(const double* field, const Lag& l) {
  Numeric out = 0;
  for (Index i = 0; i <= l.order; i++) out += field[l.i0 + i] * l.w[i];
  return out;
}
// table::absorption, lookup_map.cpp:190-238
bool absorption(std::vector<double>& absorb, const ab200_lookup_table& t, const double* f, Index nf, Numeric T, Numeric P,
                const double* vmr, int h2o, int po, int to, int wo, int fo, Numeric extpol, std::string& err) {
  Lag plag, tlag, wlag;
  tlag.w[0] = wlag.w[0] = 1.0;
  std::vector<Lag> flag(nf);
  for (Index i = 0; i < nf; i++)
    if (not make_lag(flag[i], t.f_grid, t.nf, fo, f[i], extpol, "Frequency", err)) return false;
  if (not make_lag(plag, t.log_p_grid, t.np, po, std::log(P), extpol, "Log-Pressure", err)) return false;
  if (t.do_w) {  // water_lagrange :161-173
    const Numeric x = vmr[h2o] / interp1(t.water_atmref, plag);
    if (not make_lag(wlag, t.w_pert, t.nw, wo, x, extpol, "Water VMR", err)) return false;
  }
  if (t.do_t) {  // temperature_lagrange :175-188
    const Numeric x = T - interp1(t.t_atmref, plag);
    if (not make_lag(tlag, t.t_pert, t.nt, to, x, extpol, "Temperature", err)) return false;
  }
  const Numeric nd = vmr[t.species] * number_density(P, T);  // AtmPoint::number_density(species)
  for (Index i = 0; i < nf; i++) {
    // reinterp / interp, lagrange_interp.h:920-940: temperature outermost, then water, pressure, frequency;
    // every term is ((((field * wt) * ww) * wp) * wf), dimensions the table does not have are left out (:213-231)
    Numeric out = 0;
    for (Index a = 0; a <= tlag.order; a++)
      for (Index b = 0; b <= wlag.order; b++)
        for (Index c = 0; c <= plag.order; c++)
          for (Index d = 0; d <= flag[i].order; d++) {
            Numeric v = t.xsec[(((tlag.i0 + a) * t.nw + wlag.i0 + b) * t.np + plag.i0 + c) * t.nf + flag[i].i0 + d];
            if (t.do_t) v *= tlag.w[a];
            if (t.do_w) v *= wlag.w[b];
            v *= plag.w[c];
            v *= flag[i].w[d];
            out += v;
          }
    absorb[i] += out * nd;
  }
  return true;
}
}  // namespace lut

// _spectral_propmatAddLookup, src/m_lookup.cc:20-141, for every level of a path
int orc_lookup_levels(const ab200_lookup_table* tables, int32_t n_tables, int64_t nf, const double* f_in, int64_t f_level_stride,
                      const ab200_atm_path* atm, int32_t n_species, int32_t h2o_species, int32_t select_species, int32_t nq,
                      const ab200_target* targets, const double* target_d, int32_t no_negative_absorption, int32_t po, int32_t to,
                      int32_t wo, int32_t fo, double extpolfac, double* K, double* dK) {
  const int np = atm->np;
  if (std::max({po, to, wo, fo}) >= lut::MAXP) return fail(AB200_ERR_UNSUPPORTED, "interpolation order above 15");
  std::string err;
  auto total = [&](std::vector<double>& out, const double* f, Numeric T, Numeric P, const double* vmr) -> int {
    std::fill(out.begin(), out.end(), 0.0);
    bool found = false;
    for (int k = 0; k < n_tables; k++) {
      if (select_species != AB200_SPECIES_BATH and tables[k].species != select_species) continue;
      found = true;
      if (tables[k].nt * tables[k].nw * tables[k].np * tables[k].nf == 0) continue;  // xsec.empty(), :201
      if (not lut::absorption(out, tables[k], f, nf, T, P, vmr, h2o_species, po, to, wo, fo, extpolfac, err))
        return fail(AB200_ERR_INVALID, err);
    }
    if (select_species != AB200_SPECIES_BATH and not found) return fail(AB200_ERR_INVALID, "no lookup table for the selected species");  // .at()
    return 0;
  };
  std::vector<double> ab(nf), dab(nf), vm(n_species);
  for (int ip = 0; ip < np; ip++) {
    const double* f = f_in + ip * f_level_stride;
    const double* vmr = atm->vmr + static_cast<Index>(ip) * n_species;
    if (int rc = total(ab, f, atm->T[ip], atm->P[ip], vmr)) return rc;
    for (Index i = 0; i < nf; i++)
      if (no_negative_absorption == 0 or ab[i] > 0.0) K[(static_cast<Index>(ip) * nf + i) * 7] += ab[i];
    for (int q = 0; q < nq; q++) {
      const Numeric d = target_d[q];
      if (not std::isnormal(d)) return fail(AB200_ERR_INVALID, "The target is not good, it lacks a perturbation value.");
      if (targets[q].kind != AB200_TARGET_T and targets[q].kind != AB200_TARGET_VMR)
        return fail(AB200_ERR_UNSUPPORTED, "only temperature and VMR targets");  // the library's scope, m_lookup.cc:88-108 not restated
      std::copy(vmr, vmr + n_species, vm.begin());
      Numeric T = atm->T[ip];
      if (targets[q].kind == AB200_TARGET_T) T += d; else vm[targets[q].species] += d;
      if (int rc = total(dab, f, T, atm->P[ip], vm.data())) return rc;
      const Numeric d_inv = 1.0 / d;
      for (Index i = 0; i < nf; i++)
        if (no_negative_absorption == 0 or dab[i] > 0.0)
          dK[((static_cast<Index>(ip) * nq + q) * nf + i) * 7] = (dab[i] - ab[i]) * d_inv;  // '=' (sic), :130-135
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------
// predefined continua: the four "StandardType" models, src/core/predefined/standard.cc
// ---------------------------------------------------------------------------
namespace predef {
struct Pt {
  Numeric T, P, o2, n2, h2o, lwc = 0.0;
};

// PWR20xx::compute_h2o, src/core/predefined/PWR20xx.cc:21-166: 16 / 20 H2O lines with pressure shifts, the speed-dependent
// shape 2 (1 - sqrt(pi) xrt erfcx(xrt)) / (w2 - i d2) within 10 half-widths of a line that has a w2, Van Vleck-Weisskopf
// wings with Clough's 750 GHz local-line definition elsewhere, and the foreign + self continuum.  tab: one row of 19 numbers
// per line (arts_b200/csrc/predef_tables.h), sc: tref_lines, tref_cont, c_f, xc_f, c_s, xc_s.
Numeric pwr20xx_h2o(const double* tab, int nl, const double* sc, Numeric f_hz, const Pt& a) {
  using std::pow;
  const Numeric t = a.T, p_pa = a.P, h2o_vmr = a.h2o;
  const Numeric tref_lines = sc[0], tref_cont = sc[1], c_f = sc[2], xc_f = sc[3], c_s = sc[4], xc_s = sc[5];
  const Numeric p_hpa = p_pa * 1e-2, pvap_hpa = h2o_vmr * p_hpa, pdry_hpa = p_hpa - pvap_hpa;
  const Numeric pvap_bar = pvap_hpa * 1e-3, pdry_bar = pdry_hpa * 1e-3;
  const Numeric theta_cont = tref_cont / t, theta_line = tref_lines / t;
  const Numeric log_theta_line = std::log(theta_line);
  constexpr Numeric line_cutoff = 750.0;
  const Numeric f = f_hz * 1e-9;
  constexpr Numeric conv_cont = 1e-3;
  const Numeric cont = ((c_f * pdry_hpa * pow(theta_cont, xc_f) + c_s * pvap_hpa * pow(theta_cont, xc_s)) * pvap_hpa * pow2(f) * conv_cont);
  Numeric line_sum = 0.0;
  for (int i = 0; i < nl; i++) {
    const double* c = tab + 19 * i;
    const Numeric frequency_ghz = c[0], strength_296 = c[1], B = c[2], w0_air = c[3], xw_air = c[4], w0_self = c[5], xw_self = c[6],
                  d_air = c[7], d_self = c[9], a_air = c[11], a_self = c[12], w2_air = c[13], w2_self = c[15], d2_air = c[17],
                  d2_self = c[18];
    // exponents that are not given fall back to the width's (:63-76)
    const Numeric xd_air = c[8] <= 0 ? xw_air : c[8], xd_self = c[10] <= 0 ? xw_self : c[10];
    const Numeric x2_air = c[14] <= 0 ? xw_air : c[14], x2_self = c[16] <= 0 ? xw_self : c[16];
    const Numeric w0 = w0_air * pdry_bar * pow(theta_line, xw_air) + w0_self * pvap_bar * pow(theta_line, xw_self);
    const Numeric w2 = w2_air * pdry_bar * pow(theta_line, x2_air) + w2_self * pvap_bar * pow(theta_line, x2_self);
    const Numeric d2 = d2_air * pdry_bar + d2_self * pvap_bar;
    const Numeric shift_f = d_air * pdry_bar * (1.0 - a_air * log_theta_line) * pow(theta_line, xd_air);
    const Numeric shift_s = d_self * pvap_bar * (1.0 - a_self * log_theta_line) * pow(theta_line, xd_self);
    const Numeric shift = shift_f + shift_s;
    const Numeric strength = strength_296 * pow(theta_line, 2.5) * std::exp(B * (1.0 - theta_line));
    const Numeric base = w0 / (pow2(line_cutoff) + pow2(w0));
    const Numeric df_1 = f - frequency_ghz - shift, df_2 = f + frequency_ghz + shift;
    Numeric resonant = 0.0;
    if ((w2 > 0) && (std::abs(df_1) < (10.0 * w0))) {
      const Complex denom = Complex(w2, -d2);
      const Complex xc    = Complex(w0 - 1.5 * w2, df_1 + 1.5 * d2) / denom;
      const Complex xrt   = std::sqrt(xc);
      constexpr Numeric magic_number = 1.77245385090551603;
      const Complex pxw = magic_number * xrt * Faddeeva::erfcx(xrt, 0);
      const Complex sd  = 2.0 * (1.0 - pxw) / denom;
      resonant += sd.real() - base;
    } else if (std::abs(df_1) < line_cutoff) {
      resonant += w0 / (pow2(df_1) + pow2(w0)) - base;
    }
    if (std::abs(df_2) < line_cutoff) resonant += w0 / (pow2(df_2) + pow2(w0)) - base;
    line_sum += strength * resonant * pow2(f / frequency_ghz);
  }
  constexpr Numeric conv = 1e-13;
  line_sum = conv * Constant::inv_pi * line_sum * p_pa * h2o_vmr / (Constant::k * t);
  return line_sum + cont;
}
// PWR20xx::compute_o2, PWR20xx.cc:494-573: 49 O2 lines with first- and second-order mixing (y, g, delta_nu) and the dry
// continuum; a non-positive total adds nothing (:567-569).  tab: one row of 10 numbers per line.
Numeric pwr20xx_o2(const double* tab, int nl, Numeric f_hz, const Pt& a) {
  const Numeric t = a.T, p_pa = a.P, o2_vmr = a.o2, h2o_vmr = a.h2o;
  constexpr Numeric cont_width_300 = 0.56, x = 0.754, t_ref = 300.0;
  const Numeric theta = t_ref / t, theta_minus_1 = theta - 1.0;
  const Numeric b = std::pow(theta, x);
  const Numeric pvap_pa = h2o_vmr * p_pa, pdry_pa = p_pa - pvap_pa;
  const Numeric pvap_bar = pvap_pa * 1e-5, pdry_bar = pdry_pa * 1e-5;
  const Numeric den = pdry_bar * b + 1.2 * pvap_bar * theta;
  const Numeric df_cont = cont_width_300 * den, pe2 = pow2(den);
  const Numeric f_ghz = f_hz * 1e-9, f2_ghz = pow2(f_ghz);
  const Numeric cont = 1.584e-17 * f2_ghz * df_cont / (theta * (f2_ghz + pow2(df_cont)));
  Numeric lines_sum = 0.0;  // std::valarray::sum(): the first element, then += in order
  for (int i = 0; i < nl; i++) {
    const double* c = tab + 10 * i;  // frequency_ghz, strength_300, be, width_300, y0, y1, g0, g1, dnu0, dnu1
    const Numeric y = den * (c[4] + c[5] * theta_minus_1), delta_nu = pe2 * (c[8] + c[9] * theta_minus_1);
    const Numeric g = 1.0 + pe2 * (c[6] + c[7] * theta_minus_1), width = c[3] * den;
    const Numeric strength = c[1] * std::exp(-c[2] * theta_minus_1);
    const Numeric df_1 = f_ghz - c[0] - delta_nu, df_2 = f_ghz + c[0] + delta_nu;
    const Numeric den_1 = pow2(df_1) + pow2(width), den_2 = pow2(df_2) + pow2(width);
    const Numeric sfac_1 = (width * g + df_1 * y) / den_1, sfac_2 = (width * g - df_2 * y) / den_2;
    const Numeric line = strength * (sfac_1 + sfac_2) * pow2((f_ghz / c[0]));
    lines_sum = i == 0 ? line : lines_sum + line;
  }
  const Numeric sum = lines_sum + cont;
  constexpr Numeric conv = 1e-13;
  const Numeric absorption = 1.004 * conv * o2_vmr * Constant::inv_pi / (Constant::k * t_ref) * sum * pdry_pa * pow3(theta);
  return absorption > 0 ? absorption : 0.0;
}
Numeric model(int m, Numeric f, const Pt& a) {
  using std::pow;
  switch (m) {
    case AB200_PREDEF_H2O_PWR2021: return pwr20xx_h2o(ab200_pwr2021_h2o, AB200_PWR2021_H2O_LINES, ab200_pwr2021_h2o_scalars, f, a);
    case AB200_PREDEF_H2O_PWR2022: return pwr20xx_h2o(ab200_pwr2022_h2o, AB200_PWR2022_H2O_LINES, ab200_pwr2022_h2o_scalars, f, a);
    case AB200_PREDEF_O2_PWR2021: return pwr20xx_o2(ab200_pwr2021_o2, AB200_PWR2021_O2_LINES, f, a);
    case AB200_PREDEF_O2_PWR2022: return pwr20xx_o2(ab200_pwr2022_o2, AB200_PWR2022_O2_LINES, f, a);
    case AB200_PREDEF_O2_MPM2020: {  // MPM2020::compute + sum_lines, src/core/predefined/MPM2020.cc:18-36, :38-149
      constexpr Numeric conv = 0.1820 * 1e-7 / (2.0946 * std::numbers::log10e), x = 0.754;
      const Numeric p = a.P * 1e-5, theta = 300. / a.T, dt = theta - 1, tadapt = std::pow(theta, x);
      const Numeric ta1 = tadapt * p, ta2 = pow2(tadapt * p), tp = pow3(theta) * p;
      const Numeric f_ghz = f * 1e-9;
      Numeric acc = 0;
      for (int i = 0; i < AB200_MPM2020_O2_LINES; i++) {
        const double* l = ab200_mpm2020_o2 + 10 * i;  // f0, c, a2, ga, y0, y1, g0, g1, dv0, dv1
        const Numeric y = (l[4] + l[5] * dt) * ta1, g = (l[6] + l[7] * dt) * ta2, dv = (l[8] + l[9] * dt) * ta2;
        const Numeric ga = l[3] * ta1;
        const Numeric c  = (l[1] / l[0]) * tp * std::exp(-l[2] * dt);
        acc += c * ((ga * (1 + g) + y * (f_ghz - l[0] - dv)) / (pow2(ga) + pow2(f_ghz - l[0] - dv)) +
                    (ga * (1 + g) - y * (f_ghz + l[0] + dv)) / (pow2(ga) + pow2(f_ghz + l[0] + dv)));
      }
      return acc > 0 ? conv * a.o2 * pow2(f_ghz) * acc : 0.0;
    }
    case AB200_PREDEF_O2_TRE05: {  // TRE05::oxygen, src/core/predefined/TRE05.cc:115-296 (line shape :37-70)
      constexpr Numeric dB_km_to_1_m = (1.00000e-3 / (10.0 * std::numbers::log10e));
      constexpr Numeric VMRISO = 0.2085, S0 = 6.140e-5, G0 = 0.560e-3, X0 = 0.800, Hz_to_GHz = 1.000000e-9, Pa_to_hPa = 1.000000e-2;
      const Numeric t = a.T, p_pa = a.P, oxygen_vmr = a.o2, water_vmr = a.h2o;
      if (oxygen_vmr == 0.) return 0.0;
      const Numeric theta = (300.0 / t);
      const Numeric pwv = Pa_to_hPa * p_pa * water_vmr, pda = (Pa_to_hPa * p_pa) - pwv;
      const Numeric strength_cont = S0 * pda * pow(theta, 2.);
      const Numeric gam_cont      = G0 * (pwv + pda) * pow(theta, X0);
      const Numeric ff            = f * Hz_to_GHz;
      const Numeric Nppc          = strength_cont * ff * gam_cont / (pow(ff, 2.) + pow(gam_cont, 2.));
      Numeric Nppl = 0.0;
      for (int i = 0; i < AB200_TRE05_O2_LINES; i++) {
        const double* l = ab200_tre05_o2 + 7 * i;
        const Numeric strength = 1.000e-6 * pda * l[1] / l[0] * pow(theta, 3.) * std::exp(l[2] * (1.0 - theta));
        const Numeric gam      = (l[3] * 0.001 * ((pda * pow(theta, (0.8 - l[4]))) + (1.10 * pwv * theta)));
        const Numeric delta    = ((l[5] + l[6] * theta) * (pda + pwv) * pow(theta, 0.8) * 0.001);
        const Numeric f_minus  = (gam - delta * (l[0] - ff)) / ((l[0] - ff) * (l[0] - ff) + gam * gam);
        const Numeric f_plus   = (gam - delta * (l[0] + ff)) / ((l[0] + ff) * (l[0] + ff) + gam * gam);
        Nppl += strength * (ff * (f_minus + f_plus));
      }
      if (Nppl < 0.000) Nppl = 0.0000;
      return oxygen_vmr * dB_km_to_1_m * 0.1820 * ff * (Nppl + Nppc) / VMRISO;
    }
    case AB200_PREDEF_N2_SELFCONT_PWR2021: {  // PWR20xx::compute_n2, PWR20xx.cc:792-833
      const Numeric theta = 300.0 / a.T;
      const Numeric pdry_pa = a.P * (1.0 - a.h2o), pdry_hpa = pdry_pa * 1e-2;
      constexpr Numeric assumed_n2_vmr = 0.781, continuum_coefficient = 9.95e-14;
      const Numeric cont = (a.n2 / assumed_n2_vmr) * continuum_coefficient * pow2(pdry_hpa) * pow(theta, 3.22);
      const Numeric f_ghz = f * 1e-9;
      const Numeric frequency_dependence = 0.5 + 0.5 / (1.0 + pow2(f_ghz / 450.0));
      return cont * frequency_dependence * pow2(f_ghz) / 1000.0;
    }
    case AB200_PREDEF_LIQUIDCLOUD_ELL07: {  // ELL07::compute, src/core/predefined/ELL07.cc:39-188 (user errors: ell07_refused below)
      using Constant::pi;
      using Constant::two_pi;
      constexpr Numeric dB_km_to_1_m = (1e-3 / (10.0 * std::numbers::log10e));  // Constant::log10_euler, arts_constants.h
      const Numeric lwc = a.lwc;
      if (lwc < 1e-10) return 0.0;
      constexpr Numeric m = 1.00e3;
      // table 2 of Ellison (2007): Debye amplitudes a_i exp(-b_i t), relaxation times c_i exp(d_i / (t + tc)), two resonances p_i
      constexpr Numeric a1 = 79.23882, a2 = 3.815866, a3 = 1.634967, tc = 133.1383, b1 = 0.004300598, b2 = 0.01117295, b3 = 0.006841548;
      constexpr Numeric c1 = 1.382264e-13, c2 = 3.510354e-16, c3 = 6.30035e-15, d1 = 652.7648, d2 = 1249.533, d3 = 405.5169;
      constexpr Numeric p0 = 0.8379692, p1 = -0.006118594, p2 = -0.000012936798, p3 = 4235901000000.0, p4 = -14260880000.0,
                        p5 = 273815700.0, p6 = -1246943.0, p7 = 9.618642e-14, p8 = 1.795786e-16, p9 = -9.310017E-18, p10 = 1.655473e-19,
                        p11 = 0.6165532, p12 = 0.007238532, p13 = -0.00009523366, p14 = 15983170000000.0, p15 = -74413570000.0,
                        p16 = 497448000.0, p17 = 2.882476e-14, p18 = -3.142118e-16, p19 = 3.528051e-18;
      const Numeric t_cels    = a.T - 273.15;
      const Numeric epsilon_s = 87.9144 - 0.404399 * t_cels - 9.58726e-4 * pow2(t_cels) - 1.32802e-6 * pow3(t_cels);
      const Numeric delta1 = a1 * std::exp(-b1 * t_cels), delta2 = a2 * std::exp(-b2 * t_cels), delta3 = a3 * std::exp(-b3 * t_cels);
      const Numeric tau1 = c1 * std::exp(d1 / (t_cels + tc)), tau2 = c2 * std::exp(d2 / (t_cels + tc)), tau3 = c3 * std::exp(d3 / (t_cels + tc));
      const Numeric delta4 = p0 + p1 * t_cels + p2 * pow2(t_cels);
      const Numeric f0     = p3 + p4 * t_cels + p5 * pow2(t_cels) + p6 * pow3(t_cels);
      const Numeric tau4   = p7 + p8 * t_cels + p9 * pow2(t_cels) + p10 * pow3(t_cels);
      const Numeric delta5 = p11 + p12 * t_cels + p13 * pow2(t_cels);
      const Numeric f1     = p14 + p15 * t_cels + p16 * pow2(t_cels);
      const Numeric tau5   = p17 + p18 * t_cels + p19 * pow2(t_cels);
      auto debye_re = [&](Numeric tau, Numeric delta) { return pow2(tau) * delta / (1. + pow2(two_pi * f * tau)); };
      auto debye_im = [&](Numeric tau, Numeric delta) { return tau * delta / (1. + pow2(two_pi * f * tau)); };
      auto lor      = [&](Numeric tau, Numeric fc) { return 1. + pow2(two_pi * tau * fc); };
      const Numeric Reepsilon =
          epsilon_s - pow2((two_pi * f)) * (debye_re(tau1, delta1) + debye_re(tau2, delta2) + debye_re(tau3, delta3)) -
          pow2(two_pi * tau4) * delta4 / 2. * (f * (f0 + f) / lor(tau4, f0 + f) - f * (f0 - f) / lor(tau4, f0 - f)) -
          pow2(two_pi * tau5) * delta5 / 2. * (f * (f1 + f) / lor(tau5, f1 + f) - f * (f1 - f) / lor(tau5, f1 - f));
      const Numeric Imepsilon = two_pi * f * (debye_im(tau1, delta1) + debye_im(tau2, delta2) + debye_im(tau3, delta3)) +
                                pi * f * tau4 * delta4 * (1. / lor(tau4, f0 + f) + 1. / lor(tau4, f0 - f)) +
                                pi * f * tau5 * delta5 * (1. / lor(tau5, f1 + f) + 1. / lor(tau5, f1 - f));
      const Numeric ImNw = 1.500 / m * (3.000 * Imepsilon / (pow2((Reepsilon + 2.000)) + pow2(Imepsilon)));
      return lwc * 1.000e6 * dB_km_to_1_m * 0.1820 * (f * 1e-9) * ImNw;
    }
    case AB200_PREDEF_O2_SELFCONT_STANDARD: {  // Standard::oxygen :51-84
      constexpr Numeric C = (1.108e-14 / pow2(3.0e2));
      const Numeric G0 = 5600.000, G0A = 1.000, G0B = 1.100, XG0d = 0.800, XG0w = 1.000;
      const Numeric TH    = 3.0e2 / a.T;
      const Numeric ph2o  = a.P * a.h2o;
      const Numeric pdry  = a.P - ph2o;
      const Numeric gamma = G0 * (G0A * pdry * pow(TH, XG0d) + G0B * ph2o * pow(TH, XG0w));
      return a.o2 * C * a.P * pow2(TH) * (gamma * pow2(f) / (pow2(f) + pow2(gamma)));
    }
    case AB200_PREDEF_N2_SELFCONT_STANDARD: {  // Standard::nitrogen :118-138
      constexpr Numeric C = 1.05e-38, xf = 2.00, xt = 3.55, xp = 2.00;
      return a.n2 * C * pow(300.00 / a.T, xt) * pow(f, xf) * pow(a.P, xp) * pow(a.n2, xp - 1);
    }
    case AB200_PREDEF_H2O_FOREIGNCONT_STANDARD: {  // Standard::water_foreign :166-184
      constexpr Numeric C = 5.43e-35, x = 0.0;
      const Numeric pdry  = a.P * (1.000e0 - a.h2o);
      const Numeric dummy = C * pow(300. / a.T, x + 3) * a.P * pdry;
      return a.h2o * dummy * pow2(f);
    }
    case AB200_PREDEF_H2O_SELFCONT_STANDARD: {  // Standard::water_self :212-226
      constexpr Numeric C = 1.796e-33, x = 4.5;
      const Numeric dummy = C * pow(300. / a.T, x + 3) * pow2(a.P) * a.h2o;
      return a.h2o * dummy * pow2(f);
    }
    case AB200_PREDEF_H2O_PWR98: {  // PWR98::water, src/core/predefined/PWR98.cc:40-242
      const Numeric t = a.T, p_pa = a.P, vmr = a.h2o;
      const Numeric pvap_dummy = 1e-2 * p_pa;
      const Numeric pvap       = 1e-2 * p_pa * vmr;
      const Numeric pda        = (1e-2 * p_pa) - pvap;
      const Numeric den_dummy  = 3.335e16 * (2.1667 * p_pa / t);
      const Numeric ti         = (300.0 / t);
      const Numeric ti2        = pow(ti, 2.5);
      const Numeric con        = pvap_dummy * pow3(ti) * 1.000e-9 * ((0.543 * pda) + (17.96 * pvap * pow(ti, 4.5)));
      const Numeric ff         = f * 1e-9;
      Numeric sum              = 0.000;
      for (int l = 0; l < AB200_PWR98_H2O_LINES; l++) {
        const Numeric* c       = ab200_pwr98_h2o + 7 * l;  // fl, s1, b2, w3, x, ws, xs
        const Numeric width    = (c[3] * pda * pow(ti, c[4])) + (c[5] * pvap * pow(ti, c[6]));
        const Numeric wsq      = width * width;
        const Numeric strength = c[1] * ti2 * std::exp(c[2] * (1.0 - ti));
        const Numeric df0 = ff - c[0], df1 = ff + c[0];
        const Numeric base = width / (wsq + 562500.000);  // Clough's local line definition: 750 GHz cutoff
        Numeric res        = 0.000;
        if (std::fabs(df0) < 750.0) res += width / (df0 * df0 + wsq) - base;
        if (std::fabs(df1) < 750.0) res += width / (df1 * df1 + wsq) - base;
        sum += strength * res * pow2(ff / c[0]);
      }
      const Numeric absl = 0.3183e-4 * den_dummy * sum;
      return vmr * 1.000e-3 * (absl + (con * ff * ff));
    }
    case AB200_PREDEF_O2_PWR98: {  // PWR98::oxygen, PWR98.cc:297-434
      const Numeric t = a.T, p_pa = a.P, vmr = a.o2, h2o = a.h2o;
      constexpr Numeric WB300 = 0.56, X = 0.80;
      if (vmr == 0.) return 0.0;
      const Numeric TH = 3.0000e2 / t, TH1 = (TH - 1.000e0), B = pow(TH, X);
      const Numeric PRESWV = 1e-2 * (p_pa * h2o);
      const Numeric PRESDA = 1e-2 * (p_pa * (1.000e0 - h2o));
      const Numeric DEN    = 0.001 * (PRESDA * B + 1.1 * PRESWV * TH);
      const Numeric DENS   = 0.001 * (PRESDA + 1.1 * PRESWV) * TH;
      const Numeric DFNR   = WB300 * DEN;
      const Numeric CCONT  = 1.23e-10 * pow2(TH) * p_pa;
      const Numeric ff     = 1e-9 * f;
      const Numeric CONT   = CCONT * (ff * ff * DFNR / (ff * ff + DFNR * DFNR));
      Numeric SUM          = 0.000e0;
      for (int l = 0; l < AB200_PWR98_O2_LINES; ++l) {
        const Numeric* c = ab200_pwr98_o2 + 6 * l;  // F, S300, Y300, W300, BE, V
        const Numeric DF  = c[3] * ((std::fabs((c[0] - 118.75)) < 0.10) ? DENS : DEN);
        const Numeric Y   = 0.001 * 0.01 * p_pa * B * (c[2] + c[5] * TH1);
        const Numeric STR = c[1] * std::exp(-c[4] * TH1);
        const Numeric SF1 = (DF + (ff - c[0]) * Y) / ((ff - c[0]) * (ff - c[0]) + DF * DF);
        const Numeric SF2 = (DF - (ff + c[0]) * Y) / ((ff + c[0]) * (ff + c[0]) + DF * DF);
        SUM += STR * (SF1 + SF2) * (ff / c[0]) * (ff / c[0]);
      }
      return vmr * (CONT + (2.414322e7 * SUM * p_pa * pow3(TH) / Constant::pi));
    }
    case AB200_PREDEF_H2O_MPM89: {  // MPM89::water, MPM89.cc:95-180, line shape :36-65
      constexpr Numeric dB_km_to_1_m = (1e-3 / (10.0 * std::numbers::log10e));  // Constant::log10_euler
      const Numeric t = a.T, p_pa = a.P, vmr = a.h2o;
      const Numeric pwv_dummy = 1e-3 * p_pa;
      const Numeric theta     = (300.0 / t);
      const Numeric pwv       = pwv_dummy * vmr;
      const Numeric pda       = pwv_dummy - pwv;
      const Numeric Nppc      = pwv_dummy * pow3(theta) * 1.000e-5 * ((0.113 * pda) + (3.57 * pwv * pow(theta, 7.5)));
      const Numeric ff        = f * 1e-9;
      struct Row { Numeric v[7]; };
      const Row* rows = reinterpret_cast<const Row*>(ab200_mpm89_h2o);
      const Numeric Nppl = std::transform_reduce(rows, rows + AB200_MPM89_H2O_LINES, 0.0, std::plus{}, [&](const Row& r) {
        const Numeric* l       = r.v;
        const Numeric strength = pwv_dummy * l[1] * pow(theta, 3.5) * std::exp(l[2] * (1.000 - theta));
        const Numeric gam      = l[3] * 0.001 * (l[5] * pwv * pow(theta, l[6]) + pda * pow(theta, l[4]));
        const Numeric f_minus  = 1.000 / ((ff - l[0]) * (ff - l[0]) + gam * gam);
        const Numeric f_plus   = 1.000 / ((ff + l[0]) * (ff + l[0]) + gam * gam);
        return strength * (std::fabs(ff / l[0]) * gam * (f_minus + f_plus));
      });
      return vmr * dB_km_to_1_m * 0.1820 * ff * (Nppl + (Nppc * ff));
    }
    case AB200_PREDEF_O2_MPM89: {  // MPM89::oxygen, MPM89.cc:270-411, line shape :207-235
      constexpr Numeric dB_km_to_1_m = (1e-3 / (10.0 * std::numbers::log10e));
      const Numeric t = a.T, p_pa = a.P, vmr = a.o2, h2o = a.h2o;
      const Numeric S0 = 6.140e-4, G0 = 5.60e-3, X0 = 0.800, VMRISO = 0.2085;
      if (vmr == 0.) return 0.0;
      const Numeric theta = (300.0 / t);
      const Numeric pwv   = 1e-3 * p_pa * h2o;
      const Numeric pda   = (1e-3 * p_pa) - pwv;
      const Numeric strength_cont = S0 * pda * pow2(theta);
      const Numeric gam_cont      = G0 * (pwv + pda) * pow(theta, X0);
      const Numeric ff            = f * 1e-9;
      const Numeric Nppc          = strength_cont * ff * gam_cont / (pow2(ff) + pow2(gam_cont));
      struct Row { Numeric v[7]; };
      const Row* rows = reinterpret_cast<const Row*>(ab200_mpm89_o2);
      const Numeric Nppl = std::transform_reduce(rows, rows + AB200_MPM89_O2_LINES, 0.0, std::plus{}, [&](const Row& r) {
        const Numeric* l       = r.v;
        const Numeric strength = l[1] * 1.000e-6 * pda * pow3(theta) * std::exp(l[2] * (1.000 - theta)) / l[0];
        const Numeric gam      = (l[3] * 1.000e-3 * ((pda * pow(theta, (0.80 - l[4]))) + (1.10 * pwv * theta)));
        const Numeric delta    = ((l[5] + l[6] * theta) * 1.000e-3 * pda * pow(theta, 0.8));
        const Numeric f_minus  = (gam - delta * (l[0] - ff)) / ((l[0] - ff) * (l[0] - ff) + gam * gam);
        const Numeric f_plus   = (gam - delta * (l[0] + ff)) / ((l[0] + ff) * (l[0] + ff) + gam * gam);
        return strength * (ff * (f_minus + f_plus));
      });
      return vmr * dB_km_to_1_m * 0.1820 * ff * (((Nppl < 0.000) ? 0.0 : Nppl) + Nppc) / VMRISO;
    }
    default: {  // MPM93::nitrogen, MPM93.cc:33-73
      constexpr Numeric xT = 3.500, xf = 1.500, gxf = 9.000 * xf, S = 2.296e-31;
      static const Numeric G = 1.930e-5 * pow(10.000, -gxf);
      constexpr Numeric fac  = 4.0 * Constant::pi / Constant::c;
      const Numeric th       = 300.0 / a.T;
      const Numeric strength = S * pow((a.P * (1.0000 - a.h2o)), 2.0) * pow(th, xT);
      return a.n2 * fac * strength * pow(f, 2.0) / (1.000 + G * pow(f, xf)) * a.n2;
    }
  }
}
int species_of(int m, const ab200_predef_species& s) {  // isot.spec of the model tag
  switch (m) {
    case AB200_PREDEF_O2_SELFCONT_STANDARD: case AB200_PREDEF_O2_PWR98: case AB200_PREDEF_O2_MPM89: case AB200_PREDEF_O2_PWR2021:
    case AB200_PREDEF_O2_PWR2022: case AB200_PREDEF_O2_TRE05: case AB200_PREDEF_O2_MPM2020: return s.o2;
    case AB200_PREDEF_N2_SELFCONT_STANDARD: case AB200_PREDEF_N2_SELFCONT_MPM93: case AB200_PREDEF_N2_SELFCONT_PWR2021: return s.n2;
    case AB200_PREDEF_LIQUIDCLOUD_ELL07: return s.liquidcloud;
    default: return s.h2o;
  }
}
// ELL07.cc:52-54, :99-117: where there is liquid water, only up to 5e-3 kg/m3, inside 210-373 K and up to 25 THz
bool ell07_refused(int m, const Pt& a, const double* f, Index nf, Numeric df = 0.0) {
  if (m != AB200_PREDEF_LIQUIDCLOUD_ELL07 or a.lwc < 1e-10) return false;
  if (a.lwc > 5.00e-3 or a.T < 210 or a.T > 373) return true;
  return std::any_of(f, f + nf, [df](Numeric x) { return x + df > 25e12; });
}
// the full O2 models refuse a non-zero O2 mixing ratio below 1e-25 (PWR98.cc:363-370, MPM89.cc:345-352)
bool o2_vmr_refused(int m, const Pt& a) {
  return (m == AB200_PREDEF_O2_PWR98 or m == AB200_PREDEF_O2_MPM89 or m == AB200_PREDEF_O2_TRE05) and a.o2 != 0. and a.o2 < 1.000e-25;
}
// MT_CKD 4.x water continua, src/core/predefined/MT_CKD400.cc: RADFN_FUN :37-77, XINT_FUN :84-92, compute_foreign_h2o :99-172,
// compute_self_h2o :174-256 (MT_CKD430.cc: the same two functions).  Sequential like the reference: a cursor walks the regular
// wavenumber table along the ascending frequency grid and keeps the four scaled coefficients around it.  out [nf] +=.
Numeric mtckd_radfn(const Numeric XVI, const Numeric XKT) {
  if (XKT > 0.0) {
    const Numeric XVIOKT = XVI / XKT;
    if (XVIOKT <= 0.01) return 0.5 * XVIOKT * XVI;
    if (XVIOKT <= 10) {
      const Numeric EXPVKT = std::expm1(-XVIOKT);
      return -XVI * EXPVKT / (2 + EXPVKT);
    }
    return XVI;
  }
  return XVI;
}
Numeric mtckd_xint(const Numeric P, const std::array<Numeric, 4>& A) {
  const Numeric C = (3 - 2 * P) * P * P, B = 0.5 * P * (1 - P), B1 = B * (1 - P), B2 = B * P;
  return -A[0] * B1 + A[1] * (1 - C + B2) + A[2] * (C + B1) - A[3] * B2;
}
void mtckd(bool self, const ab200_mtckd_water& w, Index n, const double* f_grid, Numeric df, const Pt& a, Numeric* out) {
  auto kaycm = [](Numeric x) { return x / (100 * Constant::c); };  // Conversion::freq2kaycm
  const Numeric P = a.P, T = a.T, vmrh2o = a.h2o;
  if (n == 0) return;
  const Numeric* wn = w.wavenumbers;
  const Index data_size = w.n;
  const Numeric last_wavenumber = wn[data_size - 1];
  if (kaycm(f_grid[0] + df) > last_wavenumber) return;
  constexpr Numeric RADCN2 = 1.4387752;
  const Numeric dvc = wn[1] - wn[0], recdvc = 1 / dvc;
  const Numeric P0 = (1e-3 * w.ref_press) * 1e5;  // Conversion::bar2pa
  const Numeric T0 = w.ref_temp, xkt = T / RADCN2, rho_rat = (P / P0) * (T0 / T);
  const Numeric num_den_cm2 = 1e-6 * vmrh2o * P / (Constant::k * T);
  const Numeric r = T0 / T;
  auto scl = [&](Index i) {
    return self ? w.self_absco_ref[i] * vmrh2o * rho_rat * std::pow(r, w.self_texp[i]) * mtckd_radfn(wn[i], xkt)
                : w.for_absco_ref[i] * (1.0 - vmrh2o) * rho_rat * mtckd_radfn(wn[i], xkt);
  };
  Index cur = std::distance(wn, std::lower_bound(wn, wn + data_size, kaycm(f_grid[0] + df) - 2 * dvc));
  std::array<Numeric, 4> k{0, 0, 0, 0};
  for (Index i = -1; i < 3 and cur + i < data_size; i++) k[i + 1] = (i < 0 and cur == 0) ? scl(cur + i + 2) : scl(cur + i);
  for (Index s = 0; s < n; ++s) {
    const Numeric fs = f_grid[s] + df;
    if (fs < 0) continue;
    const Numeric x = kaycm(fs);
    if (x > last_wavenumber) return;
    while (x > wn[cur + 1]) {
      std::shift_left(k.begin(), k.end(), 1);
      k.back() = data_size > cur + 3 ? scl(cur + 3) : 0;
      cur++;
    }
    const Numeric o = 1e2 * num_den_cm2 * mtckd_xint(recdvc * (x - wn[cur]), k);
    out[s] += o >= 0 ? o : 0;
  }
}
bool is_mtckd(int m) { return m >= AB200_PREDEF_H2O_FOREIGNCONT_CKDMT400 and m <= AB200_PREDEF_H2O_SELFCONT_CKDMT430; }
}  // namespace predef

// spectral_propmatAddPredefined (m_predefined_absorption_models.cc:156-191) + PredefinedModel::compute
// (predefined_absorption_models.cc:219-317) for every level
int orc_predef_levels_data(const int32_t* models, int32_t n_models, const ab200_predef_species* sp, int64_t nf, const double* f_in,
                           int64_t f_level_stride, const ab200_atm_path* atm, int32_t n_species, int32_t select_species, int32_t nq,
                           const ab200_target* targets, const double* target_d, double* K, double* dK, const ab200_mtckd_water* ckdmt400,
                           const ab200_mtckd_water* ckdmt430);
int orc_predef_levels(const int32_t* models, int32_t n_models, const ab200_predef_species* sp, int64_t nf, const double* f_in,
                      int64_t f_level_stride, const ab200_atm_path* atm, int32_t n_species, int32_t select_species, int32_t nq,
                      const ab200_target* targets, const double* target_d, double* K, double* dK) {
  return orc_predef_levels_data(models, n_models, sp, nf, f_in, f_level_stride, atm, n_species, select_species, nq, targets, target_d, K, dK,
                                nullptr, nullptr);
}
int orc_predef_levels_data(const int32_t* models, int32_t n_models, const ab200_predef_species* sp, int64_t nf, const double* f_in,
                           int64_t f_level_stride, const ab200_atm_path* atm, int32_t n_species, int32_t select_species, int32_t nq,
                           const ab200_target* targets, const double* target_d, double* K, double* dK, const ab200_mtckd_water* ckdmt400,
                           const ab200_mtckd_water* ckdmt430) {
  const int np = atm->np;
  auto v = [&](const double* vmr, int idx) { return idx >= 0 ? vmr[idx] : 0.0; };
  int it = -1;
  for (int q = 0; q < nq; q++)
    if (targets[q].kind == AB200_TARGET_T and it < 0) it = q;
  for (int ip = 0; ip < np; ip++) {
    const double* f   = f_in + ip * f_level_stride;
    const double* vmr = atm->vmr + static_cast<Index>(ip) * n_species;
    const predef::Pt a{atm->T[ip], atm->P[ip], v(vmr, sp->o2), v(vmr, sp->n2), v(vmr, sp->h2o), v(vmr, sp->liquidcloud)};
    for (int k = 0; k < n_models; k++) {
      const int m = models[k];
      if (m < 0 or m > AB200_PREDEF_H2O_SELFCONT_CKDMT430) return fail(AB200_ERR_UNSUPPORTED, "predefined model outside the path");
      if (predef::is_mtckd(m)) {
        const ab200_mtckd_water* w = m >= AB200_PREDEF_H2O_FOREIGNCONT_CKDMT430 ? ckdmt430 : ckdmt400;
        if (not w or w->n == 0) return fail(AB200_ERR_INVALID, "No data");  // check(data), MT_CKD400.cc:94-98
        if (select_species != AB200_SPECIES_BATH and predef::species_of(m, *sp) != select_species) continue;
        const bool self = m == AB200_PREDEF_H2O_SELFCONT_CKDMT400 or m == AB200_PREDEF_H2O_SELFCONT_CKDMT430;
        std::vector<Numeric> pm(nf, 0.0), pq(nf);
        predef::mtckd(self, *w, nf, f, 0.0, a, pm.data());
        auto row = [&](int q, const predef::Pt& b, Numeric df) {  // (model' - model) / d, predefined_absorption_models.cc:256-314
          std::fill(pq.begin(), pq.end(), 0.0);
          predef::mtckd(self, *w, nf, f, df, b, pq.data());
          for (Index i = 0; i < nf; i++) dK[((static_cast<Index>(ip) * nq + q) * nf + i) * 7] += (pq[i] - pm[i]) / target_d[q];
        };
        for (Index i = 0; i < nf; i++) K[(static_cast<Index>(ip) * nf + i) * 7] += pm[i];
        if (it >= 0) {
          predef::Pt b = a;
          b.T += target_d[it];
          row(it, b, 0.0);
        }
        for (int kind : {AB200_TARGET_WIND_U, AB200_TARGET_WIND_V, AB200_TARGET_WIND_W})
          for (int q = 0; q < nq; q++)
            if (targets[q].kind == kind) {
              row(q, a, target_d[q]);
              break;
            }
        for (int idx : {sp->co2, sp->o2, sp->n2, sp->h2o, sp->liquidcloud}) {
          if (idx < 0) continue;
          for (int q = 0; q < nq; q++)
            if (targets[q].kind == AB200_TARGET_VMR and targets[q].species == idx) {
              predef::Pt b = a;
              if (idx == sp->h2o) b.h2o += target_d[q];
              row(q, b, 0.0);
              break;
            }
        }
        continue;
      }
      {  // every point the reference evaluates the model at raises its own range error
        bool bad = predef::ell07_refused(m, a, f, nf);
        if (it >= 0) {
          predef::Pt b = a;
          b.T += target_d[it];
          bad = bad or predef::ell07_refused(m, b, f, nf);
        }
        for (int q = 0; q < nq; q++) {
          if (targets[q].kind >= AB200_TARGET_WIND_U and targets[q].kind <= AB200_TARGET_WIND_W) bad = bad or predef::ell07_refused(m, a, f, nf, target_d[q]);
          if (targets[q].kind == AB200_TARGET_VMR and targets[q].species == sp->liquidcloud and sp->liquidcloud >= 0) {
            predef::Pt b = a;
            b.lwc += target_d[q];
            bad = bad or predef::ell07_refused(m, b, f, nf);
          }
        }
        if (bad) return fail(AB200_ERR_INVALID, "Liquid cloud absorption model ELL07 outside its range of validity");
      }
      if (predef::o2_vmr_refused(m, a))
        return fail(AB200_ERR_INVALID, "O2 full absorption model has detected a O2 volume mixing ratio which is below the threshold of 1e-25");
      if (select_species != AB200_SPECIES_BATH and predef::species_of(m, *sp) != select_species) continue;
      for (Index i = 0; i < nf; i++) {
        const Numeric pm = predef::model(m, f[i], a);
        K[(static_cast<Index>(ip) * nf + i) * 7] += pm;
        auto dk = [&](int q) -> double& { return dK[((static_cast<Index>(ip) * nq + q) * nf + i) * 7]; };
        if (it >= 0) {  // :256-268
          predef::Pt b = a;
          b.T += target_d[it];
          dk(it) += (predef::model(m, f[i], b) - pm) / target_d[it];
        }
        // freq_jac :280-296: every wind target present gets (model(f + d) - model(f)) / d, the frequency derivative that
        // spectral_propmat_jacWindFix later turns into a wind row (the first target of each component, jac_targets.find)
        for (int kind : {AB200_TARGET_WIND_U, AB200_TARGET_WIND_V, AB200_TARGET_WIND_W})
          for (int q = 0; q < nq; q++)
            if (targets[q].kind == kind) {
              dk(q) += (predef::model(m, f[i] + target_d[q], a) - pm) / target_d[q];
              break;
            }
        // vmrs_jac :237-241: the first target of CO2, O2, N2, H2O, liquidcloud, in that order
        for (int idx : {sp->co2, sp->o2, sp->n2, sp->h2o, sp->liquidcloud}) {
          if (idx < 0) continue;
          for (int q = 0; q < nq; q++)
            if (targets[q].kind == AB200_TARGET_VMR and targets[q].species == idx) {
              predef::Pt b = a;
              if (idx == sp->o2) b.o2 += target_d[q];
              if (idx == sp->n2) b.n2 += target_d[q];
              if (idx == sp->h2o) b.h2o += target_d[q];
              if (idx == sp->liquidcloud) b.lwc += target_d[q];
              dk(q) += (predef::model(m, f[i], b) - pm) / target_d[q];
              break;
            }
        }
      }
    }
  }
  return 0;
}

// rtepack::tran for single inputs (tests: exp(-K r) against scipy expm, src/tests/test_rtepack.cc:12-33)
// Faddeeva::Dawson(complex) of the reference's object, n points
int orc_dawson(int64_t n, const double* zr, const double* zi, double* dr, double* di) {
  for (int64_t i = 0; i < n; i++) {
    const Complex d = Faddeeva::Dawson(Complex(zr[i], zi[i]), 0);
    dr[i] = d.real();
    di[i] = d.imag();
  }
  return 0;
}

// specmat sqrt(const propmat&): out [16][2] (re, im)
int orc_sqrt_propmat(const double* k, double* out) {
  const specmat s = sqrt_pm(load_pm(k));
  for (int i = 0; i < 16; i++) out[2 * i] = s.m[i].real(), out[2 * i + 1] = s.m[i].imag();
  return 0;
}

int orc_tran(const double* k1, const double* k2, double r, uint32_t flags, double* T, double* L) {
  const tran ts{load_pm(k1), load_pm(k2), r, (flags & AB200_FLAG_TRAN_EXACT) != 0};
  store(T, ts());
  if (L) store(L, ts.linsrc());
  return 0;
}

// ---------------------------------------------------------------------------
// Unit-level views of stage 1 for tests/test_refslice_pins.py: the same numbers the pipeline above uses,
// exposed so that they can be compared BITWISE with the reference's own function bodies compiled from
// /root/reference (oracle/slice_ref.py -> oracle/_ref/librefslice.so).
// ---------------------------------------------------------------------------
int orc_tmodel(int type, const double* x, double T0, double T, double* val, double* dT) {
  *val = tm_value(type, x, T0, T);
  *dT  = tm_dT(type, x, T0, T);
  return 0;
}

// One catalog line at one path level: mix[15] = G0, D0, DV, G, Y, then their d/dT, then their d/dVMR(target_species);
// zee[2] = Splitting, Strength of component (pol, iz) (0, 1 for pol = POL_NO); shape[5] = f0, inv_gd, z_imag, Re s, Im s
// built as band_shape_helper does (mode 0 / 1: single_shape_builder, inv_gd from the unsplit centre) or as the
// single_shape constructor does (mode 2: inv_gd from the split centre, lbl_lineshape_voigt_lte.cpp:226-237);
// ds[4] = dline_strength_calc_dT, dline_strength_calc_dVMR(target_species) at that shape's inv_gd and f0.
int orc_line_level(const ab200_catalog_desc* d, const ab200_atm_path* atm_path, int32_t ip, int64_t il, int32_t pol, int32_t iz,
                   int32_t target_species, int32_t mode, double* mix, double* zee, double* shape, double* ds) {
  const AtmPt atm = atm_at(*d, *atm_path, ip);
  const LineView ln{*d, il};
  int ib = 0;
  while (ib + 1 < d->n_bands and d->band_offset[ib + 1] <= il) ib++;
  const int isot = d->band_isot[ib];
  const int spec = d->isot_species[isot];
  static const int vars[5] = {AB200_VAR_G0, AB200_VAR_D0, AB200_VAR_DV, AB200_VAR_G, AB200_VAR_Y};
  for (int k = 0; k < 5; k++) {
    mix[k]      = ln.mix(vars[k], atm);
    mix[5 + k]  = ln.mix(vars[k], atm, true);
    mix[10 + k] = ln.dmix_dVMR(vars[k], atm, target_species);
  }
  const ZeemanView z{d->z_on[il] != 0, d->z_gu[il], d->z_gl[il], d->two_Ju[il], d->two_Jl[il]};
  const Pol p = static_cast<Pol>(pol);
  zee[0]      = z.Splitting(p, iz);
  zee[1]      = z.Strength(p, iz);
  const Numeric H              = std::hypot(atm.mag[0], atm.mag[1], atm.mag[2]);
  const Numeric f0             = line_center_calc(ln, atm);
  const Numeric scaled_gd_part = std::sqrt(Constant::doppler_broadening_const_squared * atm.T / d->isot_mass[isot]);
  const Numeric G0             = ln.mix(AB200_VAR_G0, atm);
  single_shape s;
  if (mode == 0) {
    s.f0     = f0;
    s.inv_gd = 1.0 / (scaled_gd_part * f0);
    s.z_imag = G0 * s.inv_gd;
    s.s      = line_strength_calc(s.inv_gd, isot, spec, ln, atm);
  } else if (mode == 1) {
    s.f0     = f0 + H * zee[0];
    s.inv_gd = 1.0 / (scaled_gd_part * f0);
    s.z_imag = G0 * s.inv_gd;
    s.s      = zee[1] * line_strength_calc(s.inv_gd, isot, spec, ln, atm);
  } else {
    s.f0     = f0 + H * zee[0];
    s.inv_gd = 1.0 / (std::sqrt(Constant::doppler_broadening_const_squared * atm.T / d->isot_mass[isot]) * s.f0);
    s.z_imag = G0 * s.inv_gd;
    s.s      = zee[1] * line_strength_calc(s.inv_gd, isot, spec, ln, atm);
  }
  shape[0] = s.f0, shape[1] = s.inv_gd, shape[2] = s.z_imag, shape[3] = s.s.real(), shape[4] = s.s.imag();
  const Complex dT = dline_strength_calc_dT(s.inv_gd, s.f0, isot, spec, ln, atm);
  const Complex dV = dline_strength_calc_dVMR(s.inv_gd, s.f0, isot, spec, target_species, ln, atm);
  ds[0] = dT.real(), ds[1] = dT.imag(), ds[2] = dV.real(), ds[3] = dV.imag();
  return 0;
}

// The line-shape model of one catalog line at one path level, one variable: out[7] = VAR(atm), dVAR_dT(atm),
// dVAR_dVMR(atm, target_species), dVAR_dX0..3(atm, target_species) - the restatement of lbl_lineshape_model.cpp:70-246 above,
// for the bitwise comparison with the reference's own text (oracle/refslice/template_mix.cpp.in).
int orc_line_mix(const ab200_catalog_desc* d, const ab200_atm_path* atm_path, int32_t ip, int64_t il, int32_t var,
                 int32_t target_species, double* out) {
  const AtmPt atm = atm_at(*d, *atm_path, ip);
  const LineView ln{*d, il};
  out[0] = ln.mix(var, atm);
  out[1] = ln.mix(var, atm, true);
  out[2] = ln.dmix_dVMR(var, atm, target_species);
  for (int c = 0; c < 4; c++) out[3 + c] = ln.dmix_dX(var, atm, target_species, c);
  return 0;
}

// single_shape at n frequencies: out[i] = {Re, Im of s F(f); Re, Im of dF(f); Re, Im of dX(ds, dz, dz_fac, f)}
int orc_shape_eval(const double* shape, const double* dsdz, int64_t n, const double* f, double* out) {
  single_shape s;
  s.f0 = shape[0], s.inv_gd = shape[1], s.z_imag = shape[2], s.s = Complex{shape[3], shape[4]};
  const Complex ds{dsdz[0], dsdz[1]}, dz{dsdz[2], dsdz[3]};
  for (int64_t i = 0; i < n; i++) {
    const Complex z_ = s.z(f[i]);
    const Complex v = s(f[i]), dd = single_shape::dF(z_, single_shape::F(z_)), t = s.dX(ds, dz, dsdz[4], f[i]);
    out[6 * i + 0] = v.real(), out[6 * i + 1] = v.imag(), out[6 * i + 2] = dd.real(), out[6 * i + 3] = dd.imag();
    out[6 * i + 4] = t.real(), out[6 * i + 5] = t.imag();
  }
  return 0;
}

// ---------------------------------------------------------------------------
// atm_pathFromPath for a 1-D AtmField: forward_atm_path (src/core/path/atm_path.cpp:19-28) ->
// Atm::Field::at (src/core/atm/atm_field.cpp:928-947) -> Data::at (:890-924) with find_limit / select (:536-566) ->
// lagrange_interp::interp with the order-1 altitude lag (functional_atm_field_interp.cpp:6-10,53-65).  The lag is the
// same code the CIA restatement uses (cia::start_index / cia::weights above).
// ---------------------------------------------------------------------------
int orc_atm_path_from_profile(const ab200_atm_profile* f, int32_t n_species, int32_t n_isot, int32_t np, const double* alt,
                              const uint8_t* in_atm, double* T, double* P, double* vmr, double* isorat, double* mag, double* wind) {
  const Index n = f->nalt;
  for (int ip = 0; ip < np; ip++) {
    Numeric a = (in_atm and not in_atm[ip]) ? f->top_of_atmosphere : alt[ip];
    if (a > f->top_of_atmosphere) return fail(AB200_ERR_INVALID, "Cannot get values above the top of the atmosphere");
    // select(), :536-550
    int type = AB200_EXTRAP_LINEAR;
    const int lowt = n == 1 ? AB200_EXTRAP_NEAREST : f->alt_low, uppt = n == 1 ? AB200_EXTRAP_NEAREST : f->alt_upp;
    if (a < f->alt[0]) {
      type = lowt;
      if (type == AB200_EXTRAP_NEAREST) a = f->alt[0];
    } else if (f->alt[n - 1] < a) {
      type = uppt;
      if (type == AB200_EXTRAP_NEAREST) a = f->alt[n - 1];
    }
    if (type == AB200_EXTRAP_NONE) return fail(AB200_ERR_INVALID, "Limit breached");
    double w[2] = {1.0, 0.0};
    Index i0 = 0;
    const Index order = n == 1 ? 0 : 1;
    if (type != AB200_EXTRAP_ZERO and order == 1) {
      i0 = cia::start_index(f->alt, n, 1, a);
      cia::weights(w, f->alt, i0, 1, a);
    }
    auto at = [&](const double* v, Index stride) -> Numeric {
      if (type == AB200_EXTRAP_ZERO) return 0.0;
      if (order == 0) return v[0];
      Numeric out = 0.0;  // lagrange_interp::interp: sum over the stencil in index order
      for (Index j = 0; j < 2; j++) out += w[j] * v[(i0 + j) * stride];
      return out;
    };
    T[ip] = at(f->T, 1);
    P[ip] = at(f->P, 1);
    if (std::isnan(P[ip]) or std::isnan(T[ip])) return fail(AB200_ERR_INVALID, "Pressure or temperature is NaN");
    for (int s = 0; s < n_species; s++) {
      const Numeric v = at(f->vmr + s, n_species);
      if (std::isnan(v) or v < 0.0) return fail(AB200_ERR_INVALID, "bad VMR");
      vmr[static_cast<Index>(ip) * n_species + s] = v;
    }
    for (int i = 0; i < n_isot; i++) isorat[static_cast<Index>(ip) * n_isot + i] = f->isorat[i];
    for (int c = 0; c < 3; c++) {
      if (mag) mag[3 * ip + c] = f->mag ? at(f->mag + c, 3) : 0.0;
      if (wind) wind[3 * ip + c] = f->wind ? at(f->wind + c, 3) : 0.0;
    }
  }
  return 0;
}

}  // extern "C"
