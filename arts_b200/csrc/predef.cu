// predef.cu — predefined continuum models added into the resident propagation matrix.
//
//   predef_kernel   spectral_propmatAddPredefined (src/m_predefined_absorption_models.cc:156-191) with
//                   Absorption::PredefinedModel::compute (src/core/absorption/predefined_absorption_models.cc:219-317) for
//                   every (frequency, level): the four "StandardType" continua of src/core/predefined/standard.cc
//                   (O2 :51-84, N2 :118-138, H2O foreign :166-184, H2O self :212-226), the temperature row and the
//                   CO2 / O2 / N2 / H2O / liquidcloud VMR rows by the reference's perturbation (model(x + d) - model(x)) / d.
//
//                   The full microwave models — PWR98::water / oxygen (src/core/predefined/PWR98.cc:40-242, :297-434),
//                   MPM89::water / oxygen (MPM89.cc:95-180, :270-411) — are resonant line lists (15 to 44 lines) plus a
//                   continuum.  Their per-line strength, width and mixing terms depend on the level only (two `pow` and one
//                   `exp` per line): the CTA evaluates them ONCE per (level, perturbed state) into shared memory and every
//                   frequency then pays two divisions per line.  MPM93::nitrogen (MPM93.cc:33-73) is a closed form.
//                   Rosenkranz's 2021 / 2022 revisions (src/core/predefined/PWR20xx.cc) go the same way: compute_h2o (:21-166,
//                   16 / 20 lines, pressure shifts, the speed-dependent shape through the complex erfcx = w(i z) within ten
//                   half-widths of a line that has a quadratic width), compute_o2 (:494-573, 49 lines with first- and
//                   second-order mixing), compute_n2 (:792-833, closed form).
//
// One thread per (frequency, level); HBM bound on K (16 B per element, + 16 B per affected Jacobian row) for the closed-form
// continua, arithmetic bound (2 x lines divisions per frequency and state) for the line lists.
#include <cmath>

#include "predef.hpp"
#include "faddeeva.cuh"  // w(z): the complex erfcx of the speed-dependent PWR2021 / PWR2022 water lines

#define AB200_TABLE_Q static __device__ const
#include "predef_tables.h"

namespace ab200 {

// one MT_CKD 4.x water table on the device: wavenumbers, self, foreign, self exponent, each [n]
struct MtckdDev {
  int32_t n;
  double ref_temp, ref_press;
  const double *v, *self, *fore, *texp;
};

struct PredefParams {
  int32_t n_models;
  int32_t models[16];
  ab200_predef_species sp;
  int64_t nf;
  const double* f;
  int64_t f_stride;
  const double* ffac;
  const double *T, *P, *vmr;
  int32_t n_species, select_species;
  double* K;
  double* dK;
  int64_t k_pitch;
  int32_t nq, it;
  int32_t tg_kind[AB200_MAX_TARGETS], tg_species[AB200_MAX_TARGETS];
  double tg_d[AB200_MAX_TARGETS];
  int* flags;  // device error flags (bit 5: an O2 mixing ratio below the full models' threshold)
  MtckdDev ckd[2];     // MT_CKD 4.0 / 4.3 water tables on the device (n == 0: not loaded)
  const double* wjac;  // [np][3] freq_wind_shift_jac per level: wind rows are d/df times f times this (spectral_propmat_jacWindFix,
                       // what the line kernels leave in dK); null: the rows stay d/df (spectral_propmatAddPredefined alone)
};

// ELL07.cc:99-117
#define PREDEF_ELL07_RANGE_MSG                                                                                                       \
  "Liquid cloud absorption model ELL07: liquid water content above 5e-3 kg/m3, temperature outside 210-373 K or frequencies above " \
  "25 THz (only valid inside these ranges)"

struct PredefPoint {
  double T, P, o2, n2, h2o, lwc;
};

// ONE body (no inlining), like predef_line_model below: base and perturbed evaluations must round alike
__device__ __noinline__ double predef_model(int m, double f, const PredefPoint& a) {
  switch (m) {
    case AB200_PREDEF_O2_SELFCONT_STANDARD: {  // Standard::oxygen
      constexpr double C = (1.108e-14 / (3.0e2 * 3.0e2));
      const double G0 = 5600.000, G0A = 1.000, G0B = 1.100, XG0d = 0.800, XG0w = 1.000;
      const double TH    = 3.0e2 / a.T;
      const double ph2o  = a.P * a.h2o;
      const double pdry  = a.P - ph2o;
      const double gamma = G0 * (G0A * pdry * pow(TH, XG0d) + G0B * ph2o * pow(TH, XG0w));
      return a.o2 * C * a.P * (TH * TH) * (gamma * (f * f) / ((f * f) + (gamma * gamma)));
    }
    case AB200_PREDEF_N2_SELFCONT_STANDARD: {  // Standard::nitrogen
      constexpr double C = 1.05e-38, xf = 2.00, xt = 3.55, xp = 2.00;
      return a.n2 * C * pow(300.00 / a.T, xt) * pow(f, xf) * pow(a.P, xp) * pow(a.n2, xp - 1);
    }
    case AB200_PREDEF_H2O_FOREIGNCONT_STANDARD: {  // Standard::water_foreign
      constexpr double C = 5.43e-35, x = 0.0;
      const double pdry  = a.P * (1.000e0 - a.h2o);
      const double dummy = C * pow(300. / a.T, x + 3) * a.P * pdry;
      return a.h2o * dummy * (f * f);
    }
    case AB200_PREDEF_N2_SELFCONT_PWR2021: {  // PWR20xx::compute_n2
      const double theta = 300.0 / a.T;
      const double pdry_hpa = (a.P * (1.0 - a.h2o)) * 1e-2;
      const double cont = (a.n2 / 0.781) * 9.95e-14 * (pdry_hpa * pdry_hpa) * pow(theta, 3.22);
      const double f_ghz = f * 1e-9, q = f_ghz / 450.0;
      return cont * (0.5 + 0.5 / (1.0 + q * q)) * (f_ghz * f_ghz) / 1000.0;
    }
    case AB200_PREDEF_N2_SELFCONT_MPM93: {  // MPM93::nitrogen
      constexpr double xT = 3.500, xf = 1.500, S = 2.296e-31;
      const double G         = 1.930e-5 * pow(10.000, -9.000 * xf);
      constexpr double fac   = 4.0 * 3.141592653589793238462643383279502884 / 299792458.0;
      const double th        = 300.0 / a.T;
      const double pd        = a.P * (1.0000 - a.h2o);
      const double strength  = S * (pd * pd) * pow(th, xT);
      return a.n2 * fac * strength * (f * f) / (1.000 + G * pow(f, xf)) * a.n2;
    }
    case AB200_PREDEF_LIQUIDCLOUD_ELL07: {  // ELL07::compute, ELL07.cc:39-188 (range errors: predef_kernel)
      if (a.lwc < 1e-10) return 0.0;
      constexpr double two_pi = 6.283185307179586476925286766559005768, pi = 3.141592653589793238462643383279502884;
      constexpr double dB_km_to_1_m = 1e-3 / (10.0 * 0.434294481903251827651128918916605082);
      const double tc = a.T - 273.15, tc2 = tc * tc, tc3 = tc2 * tc;
      const double eps_s = 87.9144 - 0.404399 * tc - 9.58726e-4 * tc2 - 1.32802e-6 * tc3;
      // three Debye relaxations (Ellison 2007, table 2) ...
      const double del[3] = {79.23882 * exp(-0.004300598 * tc), 3.815866 * exp(-0.01117295 * tc), 1.634967 * exp(-0.006841548 * tc)};
      const double tau[3] = {1.382264e-13 * exp(652.7648 / (tc + 133.1383)), 3.510354e-16 * exp(1249.533 / (tc + 133.1383)),
                             6.30035e-15 * exp(405.5169 / (tc + 133.1383))};
      // ... and two resonances: amplitude, centre, relaxation time
      const double dr[2] = {0.8379692 + -0.006118594 * tc + -0.000012936798 * tc2, 0.6165532 + 0.007238532 * tc + -0.00009523366 * tc2};
      const double fr[2] = {4235901000000.0 + -14260880000.0 * tc + 273815700.0 * tc2 + -1246943.0 * tc3,
                            15983170000000.0 + -74413570000.0 * tc + 497448000.0 * tc2};
      const double tr[2] = {9.618642e-14 + 1.795786e-16 * tc + -9.310017E-18 * tc2 + 1.655473e-19 * tc3,
                            2.882476e-14 + -3.142118e-16 * tc + 3.528051e-18 * tc2};
      const double w = two_pi * f;
      double re_d = 0.0, im_d = 0.0;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double q = 1. + (w * tau[i]) * (w * tau[i]);
        re_d += (tau[i] * tau[i]) * del[i] / q;
        im_d += tau[i] * del[i] / q;
      }
      double re = eps_s - (w * w) * re_d, im = w * im_d;
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const double wp = two_pi * tr[i] * (fr[i] + f), wm = two_pi * tr[i] * (fr[i] - f);
        const double qp = 1. + wp * wp, qm = 1. + wm * wm;
        re -= (two_pi * tr[i]) * (two_pi * tr[i]) * dr[i] / 2. * (f * (fr[i] + f) / qp - f * (fr[i] - f) / qm);
        im += pi * f * tr[i] * dr[i] * (1. / qp + 1. / qm);
      }
      const double ImNw = 1.500 / 1.00e3 * (3.000 * im / ((re + 2.000) * (re + 2.000) + im * im));
      return a.lwc * 1.000e6 * dB_km_to_1_m * 0.1820 * (f * 1e-9) * ImNw;
    }
    default: {  // Standard::water_self
      constexpr double C = 1.796e-33, x = 4.5;
      const double dummy = C * pow(300. / a.T, x + 3) * (a.P * a.P) * a.h2o;
      return a.h2o * dummy * (f * f);
    }
  }
}

// ---- MT_CKD 4.x water continua: MT_CKD400.cc:37-92 (radiation term, four-point interpolation), :99-172 (foreign), :174-256 (self) ----
__device__ __forceinline__ double mtckd_radfn(double XVI, double XKT) {
  if (XKT > 0.0) {
    const double XVIOKT = XVI / XKT;
    if (XVIOKT <= 0.01) return 0.5 * XVIOKT * XVI;
    if (XVIOKT <= 10) {
      const double EXPVKT = expm1(-XVIOKT);
      return -XVI * EXPVKT / (2 + EXPVKT);
    }
    return XVI;
  }
  return XVI;
}
__device__ __forceinline__ int mtckd_lower_bound(const double* __restrict__ v, int n, double x) {  // first i with v[i] >= x
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (v[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// f: the frequency; f_first: the first frequency of the level's grid (where the reference's cursor starts, :139-143).
// ONE body for base and perturbed points (exact zeros for targets the model does not depend on).
__device__ __noinline__ double predef_mtckd(bool self, const MtckdDev& d, double f, double f_first, const PredefPoint& a) {
  // Conversion::freq2kaycm(x) = x / (100 c)
  const int n = d.n;
  if (f < 0) return 0.0;
  const double x = f / (100 * 299792458.0), last = d.v[n - 1];
  if (x > last || f_first / (100 * 299792458.0) > last) return 0.0;
  const double dvc = d.v[1] - d.v[0], recdvc = 1 / dvc;
  // the cursor of the reference: starts at lower_bound(x_first - 2 dvc) and advances while x > v[cur + 1]
  const int cur0 = mtckd_lower_bound(d.v, n, f_first / (100 * 299792458.0) - 2 * dvc);
  int cur = mtckd_lower_bound(d.v, n, x) - 1;
  cur = cur > cur0 ? cur : cur0;
  const double P0 = (1e-3 * d.ref_press) * 1e5, T0 = d.ref_temp, xkt = a.T / 1.4387752;
  const double rho_rat = (a.P / P0) * (T0 / a.T);
  const double num_den_cm2 = 1e-6 * a.h2o * a.P / (1.380649e-23 * a.T);
  const double r = T0 / a.T;
  double k[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    int i = cur - 1 + j;
    if (i < 0) i += 2;  // the zero-frequency mirror (:147-152; only reachable with cur == 0)
    if (i >= n) {
      k[j] = 0.0;
    } else if (self) {
      k[j] = d.self[i] * a.h2o * rho_rat * pow(r, d.texp[i]) * mtckd_radfn(d.v[i], xkt);
    } else {
      k[j] = d.fore[i] * (1.0 - a.h2o) * rho_rat * mtckd_radfn(d.v[i], xkt);
    }
  }
  const double P = recdvc * (x - d.v[cur]);
  const double C = (3 - 2 * P) * P * P, B = 0.5 * P * (1 - P), B1 = B * (1 - P), B2 = B * P;
  const double out = 1e2 * num_den_cm2 * (-k[0] * B1 + k[1] * (1 - C + B2) + k[2] * (C + B1) - k[3] * B2);
  return out >= 0 ? out : 0.0;
}
__host__ __device__ inline bool predef_is_mtckd(int m) { return m >= AB200_PREDEF_H2O_FOREIGNCONT_CKDMT400 && m <= AB200_PREDEF_H2O_SELFCONT_CKDMT430; }

__host__ __device__ inline int predef_species_of(int m, const ab200_predef_species& s) {
  switch (m) {
    case AB200_PREDEF_O2_SELFCONT_STANDARD: case AB200_PREDEF_O2_PWR98: case AB200_PREDEF_O2_MPM89: case AB200_PREDEF_O2_PWR2021:
    case AB200_PREDEF_O2_PWR2022: case AB200_PREDEF_O2_TRE05: case AB200_PREDEF_O2_MPM2020: return s.o2;
    case AB200_PREDEF_N2_SELFCONT_STANDARD: case AB200_PREDEF_N2_SELFCONT_MPM93: case AB200_PREDEF_N2_SELFCONT_PWR2021: return s.n2;
    case AB200_PREDEF_LIQUIDCLOUD_ELL07: return s.liquidcloud;
    default: return s.h2o;
  }
}
__host__ __device__ inline bool predef_is_line_list(int m) {
  return (m >= AB200_PREDEF_H2O_PWR98 && m <= AB200_PREDEF_O2_MPM89) || (m >= AB200_PREDEF_H2O_PWR2021 && m <= AB200_PREDEF_O2_PWR2022) ||
         m == AB200_PREDEF_O2_TRE05 || m == AB200_PREDEF_O2_MPM2020;
}

// ---- line-list models: per (level, state) tables in shared memory ----------------------------------------------------
constexpr int PD_MAX_LINES  = 49;
constexpr int PD_MAX_STATES = 1 + AB200_MAX_TARGETS;
constexpr int PD_REC = 8;
static_assert(AB200_PWR2021_O2_LINES == AB200_PWR2022_O2_LINES && AB200_PWR2022_O2_LINES <= PD_MAX_LINES, "O2 line lists of PWR2021 / PWR2022");
struct PredefTables {
  double line[PD_MAX_STATES][PD_MAX_LINES][PD_REC];  // centre [GHz], strength, width, mixing / shift / quadratic terms
  double scal[PD_MAX_STATES][6];                     // continuum and scale factors of the state
};
__device__ __forceinline__ int predef_nlines(int m) {
  switch (m) {
    case AB200_PREDEF_H2O_PWR98: return AB200_PWR98_H2O_LINES;
    case AB200_PREDEF_O2_PWR98: return AB200_PWR98_O2_LINES;
    case AB200_PREDEF_H2O_MPM89: return AB200_MPM89_H2O_LINES;
    case AB200_PREDEF_O2_MPM89: return AB200_MPM89_O2_LINES;
    case AB200_PREDEF_H2O_PWR2021: return AB200_PWR2021_H2O_LINES;
    case AB200_PREDEF_H2O_PWR2022: return AB200_PWR2022_H2O_LINES;
    case AB200_PREDEF_O2_PWR2021: return AB200_PWR2021_O2_LINES;
    case AB200_PREDEF_O2_TRE05: return AB200_TRE05_O2_LINES;
    case AB200_PREDEF_O2_MPM2020: return AB200_MPM2020_O2_LINES;
    default: return AB200_PWR2022_O2_LINES;
  }
}
// line l of model m at the point a: what depends on the level only
__device__ __forceinline__ void predef_line_record(int m, int l, const PredefPoint& a, double* __restrict__ r) {
  if (m == AB200_PREDEF_H2O_PWR98) {  // PWR98.cc:206-216
    const double* c  = ab200_pwr98_h2o + 7 * l;
    const double ti  = 300.0 / a.T;
    const double pvap = 1e-2 * a.P * a.h2o, pda = (1e-2 * a.P) - pvap;
    r[0] = c[0];
    r[1] = c[1] * pow(ti, 2.5) * exp(c[2] * (1.0 - ti));
    r[2] = (c[3] * pda * pow(ti, c[4])) + (c[5] * pvap * pow(ti, c[6]));
    r[3] = 0.0;
  } else if (m == AB200_PREDEF_O2_PWR98) {  // PWR98.cc:372-398
    const double* c = ab200_pwr98_o2 + 6 * l;
    const double TH = 3.0000e2 / a.T, TH1 = TH - 1.000e0, B = pow(TH, 0.80);
    const double PRESWV = 1e-2 * (a.P * a.h2o), PRESDA = 1e-2 * (a.P * (1.000e0 - a.h2o));
    const double DEN  = 0.001 * (PRESDA * B + 1.1 * PRESWV * TH);
    const double DENS = 0.001 * (PRESDA + 1.1 * PRESWV) * TH;
    r[0] = c[0];
    r[1] = c[1] * exp(-c[4] * TH1);
    r[2] = c[3] * ((fabs(c[0] - 118.75) < 0.10) ? DENS : DEN);
    r[3] = 0.001 * 0.01 * a.P * B * (c[2] + c[5] * TH1);
  } else if (m == AB200_PREDEF_H2O_MPM89) {  // MPM89.cc:160-170
    const double* c    = ab200_mpm89_h2o + 7 * l;
    const double theta = 300.0 / a.T, pwv_dummy = 1e-3 * a.P, pwv = pwv_dummy * a.h2o, pda = pwv_dummy - pwv;
    r[0] = c[0];
    r[1] = pwv_dummy * c[1] * pow(theta, 3.5) * exp(c[2] * (1.000 - theta));
    r[2] = c[3] * 0.001 * (c[5] * pwv * pow(theta, c[6]) + pda * pow(theta, c[4]));
    r[3] = 0.0;
  } else if (m == AB200_PREDEF_H2O_PWR2021 || m == AB200_PREDEF_H2O_PWR2022) {  // PWR20xx::compute_h2o, PWR20xx.cc:63-120
    const bool y21   = m == AB200_PREDEF_H2O_PWR2021;
    const double* c  = (y21 ? ab200_pwr2021_h2o : ab200_pwr2022_h2o) + 19 * l;
    const double* sc = y21 ? ab200_pwr2021_h2o_scalars : ab200_pwr2022_h2o_scalars;
    const double p_hpa = a.P * 1e-2, pvap_hpa = a.h2o * p_hpa, pdry_hpa = p_hpa - pvap_hpa;
    const double pvap_bar = pvap_hpa * 1e-3, pdry_bar = pdry_hpa * 1e-3;
    const double th = sc[0] / a.T, lth = log(th);
    const double xd_air = c[8] <= 0 ? c[4] : c[8], xd_self = c[10] <= 0 ? c[6] : c[10];  // missing exponents: the width's
    const double x2_air = c[14] <= 0 ? c[4] : c[14], x2_self = c[16] <= 0 ? c[6] : c[16];
    const double w0 = c[3] * pdry_bar * pow(th, c[4]) + c[5] * pvap_bar * pow(th, c[6]);
    r[0] = c[0];
    r[1] = c[1] * pow(th, 2.5) * exp(c[2] * (1.0 - th));
    r[2] = w0;
    r[3] = c[13] * pdry_bar * pow(th, x2_air) + c[15] * pvap_bar * pow(th, x2_self);                                      // w2
    r[4] = c[17] * pdry_bar + c[18] * pvap_bar;                                                                          // d2
    r[5] = c[7] * pdry_bar * (1.0 - c[11] * lth) * pow(th, xd_air) + c[9] * pvap_bar * (1.0 - c[12] * lth) * pow(th, xd_self);  // shift
    r[6] = w0 / (750.0 * 750.0 + w0 * w0);                                                                               // base
  } else if (m == AB200_PREDEF_O2_PWR2021 || m == AB200_PREDEF_O2_PWR2022) {  // PWR20xx::compute_o2, PWR20xx.cc:516-545
    const double* c = (m == AB200_PREDEF_O2_PWR2021 ? ab200_pwr2021_o2 : ab200_pwr2022_o2) + 10 * l;
    const double theta = 300.0 / a.T, tm1 = theta - 1.0, b = pow(theta, 0.754);
    const double pvap_pa = a.h2o * a.P, pdry_pa = a.P - pvap_pa;
    const double den = (pdry_pa * 1e-5) * b + 1.2 * (pvap_pa * 1e-5) * theta, pe2 = den * den;
    r[0] = c[0];
    r[1] = c[1] * exp(-c[2] * tm1);
    r[2] = c[3] * den;                   // width
    r[3] = 1.0 + pe2 * (c[6] + c[7] * tm1);  // g
    r[4] = den * (c[4] + c[5] * tm1);    // y
    r[5] = pe2 * (c[8] + c[9] * tm1);    // delta_nu
  } else if (m == AB200_PREDEF_O2_MPM2020) {  // MPM2020::compute, MPM2020.cc:113-139: the std::transform block
    const double* c = ab200_mpm2020_o2 + 10 * l;  // f0, c, a2, ga, y0, y1, g0, g1, dv0, dv1
    const double p = a.P * 1e-5, theta = 300. / a.T, dt = theta - 1, ta1 = pow(theta, 0.754) * p, ta2 = ta1 * ta1;
    r[0] = c[0];
    r[1] = (c[1] / c[0]) * ((theta * theta * theta) * p) * exp(-c[2] * dt);
    r[2] = c[3] * ta1;               // ga
    r[3] = (c[6] + c[7] * dt) * ta2;  // g
    r[4] = (c[4] + c[5] * dt) * ta1;  // y
    r[5] = (c[8] + c[9] * dt) * ta2;  // dv
  } else if (m == AB200_PREDEF_O2_TRE05) {  // TRE05::oxygen, TRE05.cc:274-284 (note: the mixing term scales with the TOTAL pressure)
    const double* c    = ab200_tre05_o2 + 7 * l;
    const double theta = 300.0 / a.T, pwv = 1.000000e-2 * a.P * a.h2o, pda = (1.000000e-2 * a.P) - pwv;
    r[0] = c[0];
    r[1] = 1.000e-6 * pda * c[1] / c[0] * (theta * theta * theta) * exp(c[2] * (1.0 - theta));
    r[2] = c[3] * 0.001 * ((pda * pow(theta, 0.8 - c[4])) + (1.10 * pwv * theta));
    r[3] = (c[5] + c[6] * theta) * (pda + pwv) * pow(theta, 0.8) * 0.001;
  } else {  // MPM89::oxygen, MPM89.cc:372-400
    const double* c    = ab200_mpm89_o2 + 7 * l;
    const double theta = 300.0 / a.T, pwv = 1e-3 * a.P * a.h2o, pda = (1e-3 * a.P) - pwv;
    r[0] = c[0];
    r[1] = c[1] * 1.000e-6 * pda * (theta * theta * theta) * exp(c[2] * (1.000 - theta)) / c[0];
    r[2] = c[3] * 1.000e-3 * ((pda * pow(theta, 0.80 - c[4])) + (1.10 * pwv * theta));
    r[3] = (c[5] + c[6] * theta) * 1.000e-3 * pda * pow(theta, 0.8);
  }
}
__device__ __forceinline__ void predef_state_scalars(int m, const PredefPoint& a, double* __restrict__ s) {
  if (m == AB200_PREDEF_H2O_PWR98) {
    const double ti = 300.0 / a.T, pvap_dummy = 1e-2 * a.P, pvap = 1e-2 * a.P * a.h2o, pda = (1e-2 * a.P) - pvap;
    s[0] = a.h2o;
    s[1] = 3.335e16 * (2.1667 * a.P / a.T);                                                                  // den_dummy
    s[2] = pvap_dummy * (ti * ti * ti) * 1.000e-9 * ((0.543 * pda) + (17.96 * pvap * pow(ti, 4.5)));         // con
  } else if (m == AB200_PREDEF_O2_PWR98) {
    const double TH = 3.0000e2 / a.T, B = pow(TH, 0.80);
    const double PRESWV = 1e-2 * (a.P * a.h2o), PRESDA = 1e-2 * (a.P * (1.000e0 - a.h2o));
    s[0] = a.o2;
    s[1] = 1.23e-10 * (TH * TH) * a.P;                               // CCONT
    s[2] = 0.56 * (0.001 * (PRESDA * B + 1.1 * PRESWV * TH));         // DFNR
    s[3] = a.P;
    s[4] = TH * TH * TH;
  } else if (m == AB200_PREDEF_H2O_MPM89) {
    const double theta = 300.0 / a.T, pwv_dummy = 1e-3 * a.P, pwv = pwv_dummy * a.h2o, pda = pwv_dummy - pwv;
    s[0] = a.h2o;
    s[1] = pwv_dummy * (theta * theta * theta) * 1.000e-5 * ((0.113 * pda) + (3.57 * pwv * pow(theta, 7.5)));  // Nppc
  } else if (m == AB200_PREDEF_H2O_PWR2021 || m == AB200_PREDEF_H2O_PWR2022) {
    const double* sc = m == AB200_PREDEF_H2O_PWR2021 ? ab200_pwr2021_h2o_scalars : ab200_pwr2022_h2o_scalars;
    const double p_hpa = a.P * 1e-2, pvap_hpa = a.h2o * p_hpa, pdry_hpa = p_hpa - pvap_hpa, thc = sc[1] / a.T;
    s[0] = a.h2o;
    s[1] = (sc[2] * pdry_hpa * pow(thc, sc[3]) + sc[4] * pvap_hpa * pow(thc, sc[5])) * pvap_hpa;  // continuum / (f^2 conv)
    s[2] = a.P;
    s[3] = a.T;
  } else if (m == AB200_PREDEF_O2_MPM2020) {
    s[0] = a.o2;
  } else if (m == AB200_PREDEF_O2_TRE05) {
    const double theta = 300.0 / a.T, pwv = 1.000000e-2 * a.P * a.h2o, pda = (1.000000e-2 * a.P) - pwv;
    s[0] = a.o2;
    s[1] = 6.140e-5 * pda * (theta * theta);           // strength_cont
    s[2] = 0.560e-3 * (pwv + pda) * pow(theta, 0.800);  // gam_cont
  } else if (m == AB200_PREDEF_O2_PWR2021 || m == AB200_PREDEF_O2_PWR2022) {
    const double theta = 300.0 / a.T, b = pow(theta, 0.754);
    const double pvap_pa = a.h2o * a.P, pdry_pa = a.P - pvap_pa;
    s[0] = a.o2;
    s[1] = 0.56 * ((pdry_pa * 1e-5) * b + 1.2 * (pvap_pa * 1e-5) * theta);  // df_cont
    s[2] = theta;
    s[3] = pdry_pa;
  } else {
    const double theta = 300.0 / a.T, pwv = 1e-3 * a.P * a.h2o, pda = (1e-3 * a.P) - pwv;
    s[0] = a.o2;
    s[1] = 6.140e-4 * pda * (theta * theta);           // strength_cont
    s[2] = 5.60e-3 * (pwv + pda) * pow(theta, 0.800);  // gam_cont
  }
}
// the model at frequency f [Hz] from the tables of one state.  ONE body (no inlining): the base and the perturbed evaluation of a
// difference quotient must round alike, so that a target the model does not depend on gives an exact zero.
__device__ __noinline__ double predef_line_model(int m, double f, const double (*__restrict__ L)[PD_REC], const double* __restrict__ s) {
  const double ff = f * 1e-9;
  if (m == AB200_PREDEF_H2O_PWR98) {
    double sum = 0.0;
    for (int l = 0; l < AB200_PWR98_H2O_LINES; l++) {
      const double fl = L[l][0], width = L[l][2], wsq = width * width;
      const double df0 = ff - fl, df1 = ff + fl;
      const double base = width / (wsq + 562500.000);
      double res = 0.0;
      if (fabs(df0) < 750.0) res += width / (df0 * df0 + wsq) - base;
      if (fabs(df1) < 750.0) res += width / (df1 * df1 + wsq) - base;
      const double q = ff / fl;
      sum += L[l][1] * res * (q * q);
    }
    const double absl = 0.3183e-4 * s[1] * sum;
    return s[0] * 1.000e-3 * (absl + (s[2] * ff * ff));
  }
  if (m == AB200_PREDEF_O2_PWR98) {
    if (s[0] == 0.) return 0.0;
    const double DFNR = s[2];
    const double CONT = s[1] * (ff * ff * DFNR / (ff * ff + DFNR * DFNR));
    double SUM = 0.0;
    for (int l = 0; l < AB200_PWR98_O2_LINES; l++) {
      const double F = L[l][0], DF = L[l][2], Y = L[l][3];
      const double SF1 = (DF + (ff - F) * Y) / ((ff - F) * (ff - F) + DF * DF);
      const double SF2 = (DF - (ff + F) * Y) / ((ff + F) * (ff + F) + DF * DF);
      SUM += L[l][1] * (SF1 + SF2) * (ff / F) * (ff / F);
    }
    return s[0] * (CONT + (2.414322e7 * SUM * s[3] * s[4] / 3.141592653589793238462643383279502884));
  }
  if (m == AB200_PREDEF_H2O_PWR2021 || m == AB200_PREDEF_H2O_PWR2022) {  // PWR20xx.cc:124-165
    const int nl = m == AB200_PREDEF_H2O_PWR2021 ? AB200_PWR2021_H2O_LINES : AB200_PWR2022_H2O_LINES;
    double line_sum = 0.0;
    for (int l = 0; l < nl; l++) {
      const double fl = L[l][0], w0 = L[l][2], w2 = L[l][3], d2 = L[l][4], shift = L[l][5], base = L[l][6];
      const double df_1 = ff - fl - shift, df_2 = ff + fl + shift;
      double resonant = 0.0;
      if ((w2 > 0) && (fabs(df_1) < (10.0 * w0))) {
        // speed-dependent Voigt core: sd = 2 (1 - sqrt(pi) xrt erfcx(xrt)) / (w2 - i d2), xrt = sqrt((w0 - 1.5 w2 + i (df + 1.5 d2)) / (w2 - i d2))
        const double nr = w0 - 1.5 * w2, ni = df_1 + 1.5 * d2, dn = w2 * w2 + d2 * d2;
        const double xr = (nr * w2 - ni * d2) / dn, xi = (ni * w2 + nr * d2) / dn;  // (nr + i ni) / (w2 - i d2)
        const double mod = hypot(xr, xi);
        double rr, ri;  // principal square root
        if (mod == 0.0) { rr = 0.0; ri = 0.0; }
        else if (xr >= 0.0) { rr = sqrt(0.5 * (mod + xr)); ri = xi / (2.0 * rr); }
        else { ri = copysign(sqrt(0.5 * (mod - xr)), xi); rr = xi / (2.0 * ri); }
        double er, ei;
        faddeeva_w(-ri, rr, er, ei);  // erfcx(z) = w(i z)
        constexpr double spi = 1.77245385090551603;
        const double pr = spi * (rr * er - ri * ei), pi_ = spi * (rr * ei + ri * er);
        const double qr = 2.0 * (1.0 - pr), qi = -2.0 * pi_;
        resonant += (qr * w2 - qi * d2) / dn - base;  // Re[(qr + i qi) / (w2 - i d2)]
      } else if (fabs(df_1) < 750.0) {
        resonant += w0 / (df_1 * df_1 + w0 * w0) - base;
      }
      if (fabs(df_2) < 750.0) resonant += w0 / (df_2 * df_2 + w0 * w0) - base;
      const double q = ff / fl;
      line_sum += L[l][1] * resonant * (q * q);
    }
    line_sum = 1e-13 * 0.318309886183790671537767526745028724 * line_sum * s[2] * s[0] / (cst::k * s[3]);
    return line_sum + s[1] * (ff * ff) * 1e-3;
  }
  if (m == AB200_PREDEF_O2_PWR2021 || m == AB200_PREDEF_O2_PWR2022) {  // PWR20xx.cc:547-570
    const double f2 = ff * ff, dfc = s[1], theta = s[2];
    const double cont = 1.584e-17 * f2 * dfc / (theta * (f2 + dfc * dfc));
    double sum = 0.0;
    for (int l = 0; l < AB200_PWR2021_O2_LINES; l++) {
      const double fl = L[l][0], width = L[l][2], g = L[l][3], y = L[l][4], dnu = L[l][5];
      const double df_1 = ff - fl - dnu, df_2 = ff + fl + dnu;
      const double sfac_1 = (width * g + df_1 * y) / (df_1 * df_1 + width * width);
      const double sfac_2 = (width * g - df_2 * y) / (df_2 * df_2 + width * width);
      const double q = ff / fl;
      sum += L[l][1] * (sfac_1 + sfac_2) * (q * q);
    }
    sum += cont;
    const double absorption = 1.004 * 1e-13 * s[0] * 0.318309886183790671537767526745028724 / (cst::k * 300.0) * sum * s[3] * (theta * theta * theta);
    return absorption > 0 ? absorption : 0.0;
  }
  constexpr double dB_km_to_1_m = 1e-3 / (10.0 * 0.434294481903251827651128918916605082);
  if (m == AB200_PREDEF_H2O_MPM89) {
    auto term = [&](int l) {
      const double fl = L[l][0], gam = L[l][2];
      const double f_minus = 1.000 / ((ff - fl) * (ff - fl) + gam * gam);
      const double f_plus  = 1.000 / ((ff + fl) * (ff + fl) + gam * gam);
      return L[l][1] * (fabs(ff / fl) * gam * (f_minus + f_plus));
    };
    double acc = 0.0;  // the reference's std::transform_reduce adds in groups of four (libstdc++ <numeric>)
    int l = 0;
    for (; l + 4 <= AB200_MPM89_H2O_LINES; l += 4) acc += (term(l) + term(l + 1)) + (term(l + 2) + term(l + 3));
    for (; l < AB200_MPM89_H2O_LINES; l++) acc += term(l);
    return s[0] * dB_km_to_1_m * 0.1820 * ff * (acc + (s[1] * ff));
  }
  if (m == AB200_PREDEF_O2_MPM2020) {  // sum_lines, MPM2020.cc:18-36, and :142-147
    double acc = 0.0;
    for (int l = 0; l < AB200_MPM2020_O2_LINES; l++) {
      const double f0 = L[l][0], ga = L[l][2], g = L[l][3], y = L[l][4], dv = L[l][5];
      const double d1 = ff - f0 - dv, d2 = ff + f0 + dv;
      acc += L[l][1] * ((ga * (1 + g) + y * d1) / (ga * ga + d1 * d1) + (ga * (1 + g) - y * d2) / (ga * ga + d2 * d2));
    }
    constexpr double conv = 0.1820 * 1e-7 / (2.0946 * 0.434294481903251827651128918916605082);
    return acc > 0 ? conv * s[0] * (ff * ff) * acc : 0.0;
  }
  if (m == AB200_PREDEF_O2_TRE05) {  // TRE05.cc:268-294: the MPM93 O2 form, lines added one after the other
    if (s[0] == 0.) return 0.0;
    double acc = 0.0;
    for (int l = 0; l < AB200_TRE05_O2_LINES; l++) {
      const double fl = L[l][0], gam = L[l][2], delta = L[l][3];
      const double f_minus = (gam - delta * (fl - ff)) / ((fl - ff) * (fl - ff) + gam * gam);
      const double f_plus  = (gam - delta * (fl + ff)) / ((fl + ff) * (fl + ff) + gam * gam);
      acc += L[l][1] * (ff * (f_minus + f_plus));
    }
    if (acc < 0.000) acc = 0.0;
    const double Nppc = s[1] * ff * s[2] / ((ff * ff) + (s[2] * s[2]));
    return s[0] * dB_km_to_1_m * 0.1820 * ff * (acc + Nppc) / 0.2085;
  }
  if (s[0] == 0.) return 0.0;
  auto term = [&](int l) {
    const double fl = L[l][0], gam = L[l][2], delta = L[l][3];
    const double f_minus = (gam - delta * (fl - ff)) / ((fl - ff) * (fl - ff) + gam * gam);
    const double f_plus  = (gam - delta * (fl + ff)) / ((fl + ff) * (fl + ff) + gam * gam);
    return L[l][1] * (ff * (f_minus + f_plus));
  };
  double acc = 0.0;
  for (int l = 0; l + 4 <= AB200_MPM89_O2_LINES; l += 4) acc += (term(l) + term(l + 1)) + (term(l + 2) + term(l + 3));
  const double Nppc = s[1] * ff * s[2] / ((ff * ff) + (s[2] * s[2]));
  return s[0] * dB_km_to_1_m * 0.1820 * ff * (((acc < 0.000) ? 0.0 : acc) + Nppc) / 0.2085;
}

__global__ void __launch_bounds__(128) predef_kernel(PredefParams p) {
  __shared__ PredefTables tab;
  __shared__ PredefPoint pts[PD_MAX_STATES];
  __shared__ int state_of[AB200_MAX_TARGETS];  // state of target q's perturbed point, 0 when the target does not move the point
  __shared__ int nstates;
  const int64_t iv = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int lev = blockIdx.y;
  const double* __restrict__ vmr = p.vmr + int64_t(lev) * p.n_species;
  auto v = [&](int idx) { return idx >= 0 ? vmr[idx] : 0.0; };
  const PredefPoint a{p.T[lev], p.P[lev], v(p.sp.o2), v(p.sp.n2), v(p.sp.h2o), v(p.sp.liquidcloud)};
  if (threadIdx.x == 0) {
    // perturbed points: temperature target (:267-279), then the VMR targets of CO2, O2, N2, H2O, liquidcloud (:237-241, :300-314)
    int n = 1;
    pts[0] = a;
    const int vmr_idx[5] = {p.sp.co2, p.sp.o2, p.sp.n2, p.sp.h2o, p.sp.liquidcloud};
    for (int q = 0; q < p.nq; q++) {
      state_of[q] = 0;
      PredefPoint b = a;
      bool moved = false, counted = false;
      if (q == p.it) {
        b.T += p.tg_d[q];
        moved = counted = true;
      } else if (p.tg_kind[q] == AB200_TARGET_VMR) {
        bool first = true;  // jac_targets.find: the first target of that species
        for (int q2 = 0; q2 < q; q2++) first = first && !(p.tg_kind[q2] == AB200_TARGET_VMR && p.tg_species[q2] == p.tg_species[q]);
        for (int j = 0; j < 5 && first; j++)
          if (vmr_idx[j] >= 0 && vmr_idx[j] == p.tg_species[q]) counted = true;
        if (counted) {
          const int idx = p.tg_species[q];
          if (idx == p.sp.o2) b.o2 += p.tg_d[q], moved = true;
          if (idx == p.sp.n2) b.n2 += p.tg_d[q], moved = true;
          if (idx == p.sp.h2o) b.h2o += p.tg_d[q], moved = true;
          if (idx == p.sp.liquidcloud) b.lwc += p.tg_d[q], moved = true;
        }
      }
      const bool wind = p.tg_kind[q] >= AB200_TARGET_WIND_U && p.tg_kind[q] <= AB200_TARGET_WIND_W;
      bool first_wind = wind;  // jac_targets.find: the first target of that component
      for (int q2 = 0; q2 < q && wind; q2++) first_wind = first_wind && p.tg_kind[q2] != p.tg_kind[q];
      if (moved) {
        pts[n] = b;
        state_of[q] = n++;
      } else if (first_wind) {
        state_of[q] = -3;  // frequency derivative by the target's perturbation (freq_jac, :280-296)
      } else if (counted) {
        state_of[q] = -1;  // a counted target that leaves the point alone: (model - model) / d = 0, nothing to add
      } else {
        state_of[q] = -2;  // not a target of this method
      }
    }
    nstates = n;
  }
  __syncthreads();
  const double f = iv < p.nf ? (p.ffac ? p.ffac[lev] : 1.0) * p.f[int64_t(lev) * p.f_stride + iv] : 1.0;
  double kacc = 0.0, dacc[AB200_MAX_TARGETS];
#pragma unroll
  for (int q = 0; q < AB200_MAX_TARGETS; q++) dacc[q] = 0.0;
  for (int k = 0; k < p.n_models; k++) {
    const int m = p.models[k];
    if (p.select_species != AB200_SPECIES_BATH && predef_species_of(m, p.sp) != p.select_species) continue;  // CTA-uniform
    const bool lines = predef_is_line_list(m);
    if (lines) {
      if ((m == AB200_PREDEF_O2_PWR98 || m == AB200_PREDEF_O2_MPM89 || m == AB200_PREDEF_O2_TRE05) && threadIdx.x == 0)
        for (int st = 0; st < nstates; st++)
          if (pts[st].o2 != 0. && pts[st].o2 < 1.000e-25) atomicOr(p.flags, 32);
      __syncthreads();  // the previous model's tables have been read
      const int nl = predef_nlines(m);
      for (int e = threadIdx.x; e < nstates * nl; e += blockDim.x) predef_line_record(m, e % nl, pts[e / nl], tab.line[e / nl][e % nl]);
      if (threadIdx.x < nstates) predef_state_scalars(m, pts[threadIdx.x], tab.scal[threadIdx.x]);
      __syncthreads();
    }
    if (m == AB200_PREDEF_LIQUIDCLOUD_ELL07) {  // the reference's user errors, ELL07.cc:99-117 (only where there is liquid water)
      bool bad = false;
      for (int st = 0; st < nstates; st++) {
        const PredefPoint& b = pts[st];
        if (b.lwc < 1e-10) continue;
        bad = bad || b.lwc > 5.00e-3 || b.T < 210 || b.T > 373 || (iv < p.nf && f > 25e12);
        for (int q = 0; q < p.nq && st == 0; q++)
          if (state_of[q] == -3) bad = bad || (iv < p.nf && f + p.tg_d[q] > 25e12);
      }
      if (bad) atomicOr(p.flags, 64);
    }
    if (iv >= p.nf) continue;
    if (predef_is_mtckd(m)) {  // table models: their own evaluator, same row rules
      const MtckdDev& cd = p.ckd[m >= AB200_PREDEF_H2O_FOREIGNCONT_CKDMT430];
      const bool self = m == AB200_PREDEF_H2O_SELFCONT_CKDMT400 || m == AB200_PREDEF_H2O_SELFCONT_CKDMT430;
      const double f_first = (p.ffac ? p.ffac[lev] : 1.0) * p.f[int64_t(lev) * p.f_stride];
      const double pm = predef_mtckd(self, cd, f, f_first, a);
      kacc += pm;
      for (int q = 0; q < p.nq; q++) {
        const int st = state_of[q];
        if (st == -3) {
          const double row = (predef_mtckd(self, cd, f + p.tg_d[q], f_first + p.tg_d[q], a) - pm) / p.tg_d[q];
          dacc[q] += p.wjac ? row * f * p.wjac[3 * lev + (p.tg_kind[q] - AB200_TARGET_WIND_U)] : row;
        } else if (st > 0) {
          dacc[q] += (predef_mtckd(self, cd, f, f_first, pts[st]) - pm) / p.tg_d[q];
        }
      }
      continue;
    }
    const double pm = lines ? predef_line_model(m, f, tab.line[0], tab.scal[0]) : predef_model(m, f, a);
    kacc += pm;
    for (int q = 0; q < p.nq; q++) {
      const int st = state_of[q];
      if (st == -3) {  // wind target: (model(f + d) - model(f)) / d, then the wind fix of the resident rows
        const double fd = f + p.tg_d[q];
        const double pq = lines ? predef_line_model(m, fd, tab.line[0], tab.scal[0]) : predef_model(m, fd, a);
        const double row = (pq - pm) / p.tg_d[q];
        dacc[q] += p.wjac ? row * f * p.wjac[3 * lev + (p.tg_kind[q] - AB200_TARGET_WIND_U)] : row;
        continue;
      }
      if (st <= 0) continue;
      const double pq = lines ? predef_line_model(m, f, tab.line[st], tab.scal[st]) : predef_model(m, f, pts[st]);
      dacc[q] += (pq - pm) / p.tg_d[q];
    }
  }
  if (iv >= p.nf) return;
  p.K[(int64_t(lev) * p.k_pitch + iv) * 7] += kacc;
  for (int q = 0; q < p.nq; q++)
    if (dacc[q] != 0.0) p.dK[((int64_t(lev) * p.nq + q) * p.k_pitch + iv) * 7] += dacc[q];
}

// fills the model / species / target part of the parameters and validates it; 0 or an error code with the message set
}  // namespace ab200
// device-resident MT_CKD tables: one allocation, four arrays per table
struct ab200_predef_data {
  int device = 0;
  double* d_buf = nullptr;
  ab200::MtckdDev ckd[2]{};
};
namespace ab200 {

int predef_setup(PredefParams& pp, const int32_t* models, int32_t n_models, const ab200_predef_species* sp, int32_t n_species, int32_t nq,
                 const int32_t* tg_kind, const int32_t* tg_species, const double* target_d, const ab200_predef_data* data) {
  if (n_models < 0 || n_models > 16 || (n_models > 0 && !models) || !sp)
    return set_error(AB200_ERR_INVALID, "predefined models: null argument or more than 16 models");
  pp.n_models = n_models;
  pp.sp = *sp;
  for (int idx : {sp->o2, sp->n2, sp->h2o, sp->co2, sp->liquidcloud})
    if (idx >= n_species) return set_error(AB200_ERR_INVALID, "predefined models: species index beyond the VMR vector");
  for (int k = 0; k < n_models; k++) {
    const int m = models[k];
    if (m < AB200_PREDEF_O2_SELFCONT_STANDARD || m > AB200_PREDEF_H2O_SELFCONT_CKDMT430)
      return set_error(AB200_ERR_UNSUPPORTED, "predefined model " + std::to_string(m) +
                                                  " is outside the GPU path (the StandardType continua, PWR98, MPM89, MPM93 N2 and "
                                                  "PWR2021 / PWR2022, TRE05, MPM2020, ELL07 and the MT_CKD 4.x water continua are; no CPU fallback)");
    if (predef_is_mtckd(m)) {  // check(data), MT_CKD400.cc:94-98
      const int which = m >= AB200_PREDEF_H2O_FOREIGNCONT_CKDMT430;
      if (!data || data->ckd[which].n == 0) return set_error(AB200_ERR_INVALID, "No data (MT_CKD water continuum without its model data)");
      int dev = 0;
      cudaGetDevice(&dev);
      if (data->device != dev) return set_error(AB200_ERR_INVALID, "predefined model data lives on another device");
      pp.ckd[which] = data->ckd[which];
    }
    const bool need_h2o = m != AB200_PREDEF_N2_SELFCONT_STANDARD && m != AB200_PREDEF_LIQUIDCLOUD_ELL07;
    if (predef_species_of(m, *sp) < 0 || (need_h2o && sp->h2o < 0))
      return set_error(AB200_ERR_INVALID, "predefined model " + std::to_string(m) + " needs a species the atmosphere does not carry");
    pp.models[k] = m;
  }
  pp.nq = nq;
  pp.it = -1;
  if (nq > 0 && !target_d) return set_error(AB200_ERR_INVALID, "predefined models: target_d is null with Jacobian targets");
  for (int q = 0; q < nq; q++) {
    pp.tg_kind[q] = tg_kind[q]; pp.tg_species[q] = tg_species[q]; pp.tg_d[q] = target_d[q];
    if (tg_kind[q] == AB200_TARGET_T && pp.it < 0) pp.it = q;
    if (!(target_d[q] != 0.0) || !std::isfinite(target_d[q]))
      return set_error(AB200_ERR_INVALID, "predefined models: target " + std::to_string(q) + " lacks a perturbation value");
  }
  return 0;
}

int launch_predef(const PredefParams& p, int nlev, cudaStream_t stream) {
  if (p.nf == 0 || nlev == 0 || p.n_models == 0) return 0;
  dim3 grid(static_cast<unsigned>((p.nf + 127) / 128), static_cast<unsigned>(nlev));
  predef_kernel<<<grid, 128, 0, stream>>>(p);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

int predef_on_path(const int32_t* models, int32_t n_models, const ab200_predef_species* sp, const double* target_d, int64_t nf,
                   const double* d_f, int64_t f_stride, const double* d_ffac, const double* d_T, const double* d_P, const double* d_vmr,
                   int32_t n_species, int32_t select_species, double* d_K, double* d_dK, int64_t k_pitch, int32_t nq,
                   const int32_t* tg_kind, const int32_t* tg_species, int np, int* d_flags, const double* d_wjac, cudaStream_t stream,
                   const ab200_predef_data* data) {
  PredefParams pp{};
  pp.flags = d_flags;
  pp.wjac = d_wjac;
  AB_TRY(predef_setup(pp, models, n_models, sp, n_species, nq, tg_kind, tg_species, target_d, data));
  pp.nf = nf; pp.f = d_f; pp.f_stride = f_stride; pp.ffac = d_ffac; pp.T = d_T; pp.P = d_P; pp.vmr = d_vmr;
  pp.n_species = n_species; pp.select_species = select_species; pp.K = d_K; pp.dK = d_dK; pp.k_pitch = k_pitch;
  return launch_predef(pp, np, stream);
}

}  // namespace ab200

using namespace ab200;

extern "C" int ab200_predef_data_create(const ab200_mtckd_water* ckdmt400, const ab200_mtckd_water* ckdmt430, int32_t device,
                                        ab200_predef_data** out) {
  if (!out) return set_error(AB200_ERR_INVALID, "ab200_predef_data_create: null output");
  *out = nullptr;
  const ab200_mtckd_water* src[2] = {ckdmt400, ckdmt430};
  size_t total = 0;
  for (const ab200_mtckd_water* w : src) {
    if (!w) continue;
    // abs_predef_dataAddWaterMTCKD400, m_predefined_absorption_models.cc:78-86 (the lengths are one n here)
    if (!w->wavenumbers || !w->self_absco_ref || !w->for_absco_ref || !w->self_texp)
      return set_error(AB200_ERR_INVALID, "ab200_predef_data_create: null table");
    if (w->n < 4) return set_error(AB200_ERR_INVALID, "It makes no sense to have input shorter than 4");
    for (int i = 1; i < w->n; i++)
      if (!(w->wavenumbers[i] >= w->wavenumbers[i - 1]))
        return set_error(AB200_ERR_INVALID, "The wavenumbers must be increasing in a regular manner");
    total += 4 * static_cast<size_t>(w->n);
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return set_error(AB200_ERR_CUDA, "ab200_predef_data_create: no such CUDA device (this library has no CPU fallback)");
  }
  AB_CUDA(cudaSetDevice(device));
  auto* d = new ab200_predef_data;
  d->device = device;
  if (cudaMalloc(&d->d_buf, (total ? total : 1) * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    delete d;
    return set_error(AB200_ERR_NOMEM, "ab200_predef_data_create: device allocation failed");
  }
  double* at = d->d_buf;
  for (int k = 0; k < 2; k++) {
    const ab200_mtckd_water* w = src[k];
    if (!w) continue;
    const double* cols[4] = {w->wavenumbers, w->self_absco_ref, w->for_absco_ref, w->self_texp};
    for (const double* c : cols) {
      if (cudaMemcpy(at, c, static_cast<size_t>(w->n) * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(d->d_buf);
        delete d;
        return set_error(AB200_ERR_CUDA, "ab200_predef_data_create: copy to the device failed");
      }
      at += w->n;
    }
    const double* base = at - 4 * static_cast<size_t>(w->n);
    d->ckd[k] = MtckdDev{w->n, w->ref_temp, w->ref_press, base, base + w->n, base + 2 * static_cast<size_t>(w->n), base + 3 * static_cast<size_t>(w->n)};
  }
  *out = d;
  return AB200_OK;
}

extern "C" void ab200_predef_data_destroy(ab200_predef_data* data) {
  if (!data) return;
  cudaSetDevice(data->device);
  cudaFree(data->d_buf);
  delete data;
}

extern "C" int ab200_predef_levels(const int32_t* models, int32_t n_models, const ab200_predef_species* species, int64_t nf, const double* f,
                                   int64_t f_level_stride, const ab200_atm_path* atm, int32_t n_species, int32_t select_species, int32_t nq,
                                   const ab200_target* targets, const double* target_d, double* K, double* dK) {
  return ab200_predef_levels_data(models, n_models, species, nf, f, f_level_stride, atm, n_species, select_species, nq, targets, target_d, K, dK,
                                  nullptr);
}

extern "C" int ab200_predef_levels_data(const int32_t* models, int32_t n_models, const ab200_predef_species* species, int64_t nf,
                                        const double* f, int64_t f_level_stride, const ab200_atm_path* atm, int32_t n_species,
                                        int32_t select_species, int32_t nq, const ab200_target* targets, const double* target_d, double* K,
                                        double* dK, const ab200_predef_data* data) {
  if (!atm || !K || (nf > 0 && !f)) return set_error(AB200_ERR_INVALID, "ab200_predef_levels: null argument");
  if (nf < 0 || atm->np < 0 || nq < 0 || nq > AB200_MAX_TARGETS || n_species <= 0) return set_error(AB200_ERR_INVALID, "ab200_predef_levels: bad size");
  if (nq > 0 && (!targets || !dK)) return set_error(AB200_ERR_INVALID, "ab200_predef_levels: null Jacobian argument with nq > 0");
  if (f_level_stride != 0 && f_level_stride != nf) return set_error(AB200_ERR_INVALID, "f_level_stride must be 0 or nf");
  int32_t kind[AB200_MAX_TARGETS], spc[AB200_MAX_TARGETS];
  for (int q = 0; q < nq; q++) { kind[q] = targets[q].kind; spc[q] = targets[q].species; }
  PredefParams pp{};
  if (data) AB_CUDA(cudaSetDevice(data->device));
  AB_TRY(predef_setup(pp, models, n_models, species, n_species, nq, kind, spc, target_d, data));
  const int np = atm->np;
  if (np == 0 || nf == 0 || n_models == 0) return AB200_OK;
  struct Buf {
    void* p = nullptr;
    ~Buf() { cudaFree(p); }
    int put(const void* src, size_t bytes) {
      if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) { cudaGetLastError(); return 1; }
      if (src && bytes && cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); return 1; }
      return 0;
    }
  } bf, bT, bP, bv, bK, bdK, bflag;
  const int zero = 0;
  if (bflag.put(&zero, sizeof(int))) return set_error(AB200_ERR_NOMEM, "ab200_predef_levels: device allocation failed");
  pp.flags = static_cast<int*>(bflag.p);
  const size_t nfl = static_cast<size_t>(nf) * (f_level_stride ? np : 1), nk = static_cast<size_t>(np) * nf * 7;
  if (bf.put(f, nfl * 8) || bT.put(atm->T, np * 8) || bP.put(atm->P, np * 8) || bv.put(atm->vmr, static_cast<size_t>(np) * n_species * 8) ||
      bK.put(K, nk * 8) || bdK.put(dK, nk * nq * 8))
    return set_error(AB200_ERR_NOMEM, "ab200_predef_levels: device allocation or copy failed");
  pp.nf = nf; pp.f = static_cast<double*>(bf.p); pp.f_stride = f_level_stride; pp.ffac = nullptr;
  pp.T = static_cast<double*>(bT.p); pp.P = static_cast<double*>(bP.p); pp.vmr = static_cast<double*>(bv.p);
  pp.n_species = n_species; pp.select_species = select_species;
  pp.K = static_cast<double*>(bK.p); pp.dK = static_cast<double*>(bdK.p); pp.k_pitch = nf;
  AB_TRY(launch_predef(pp, np, nullptr));
  int h_flag = 0;
  AB_CUDA(cudaMemcpy(&h_flag, bflag.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (h_flag & 32)
    return set_error(AB200_ERR_INVALID, "O2 full absorption model has detected a O2 volume mixing ratio which is below the threshold of "
                                        "1e-25.  Therefore no calculation is performed.");
  if (h_flag & 64) return set_error(AB200_ERR_INVALID, PREDEF_ELL07_RANGE_MSG);
  AB_CUDA(cudaMemcpy(K, bK.p, nk * 8, cudaMemcpyDeviceToHost));
  if (nq > 0) AB_CUDA(cudaMemcpy(dK, bdK.p, nk * nq * 8, cudaMemcpyDeviceToHost));
  return AB200_OK;
}
