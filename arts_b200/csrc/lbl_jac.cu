// lbl_jac.cu — Jacobian part of stage 1 (nq > 0): d(propagation matrix)/d(temperature | VMR).
//
//   lbl_prepare_jac_kernel   per (sub-line, level, target): ds, dz, dz_fac (+ the cutoff value of dX)
//       replaces ComputeData::dt_core_calc / dVMR_core_calc
//       (reference src/core/lbl/lbl_lineshape_voigt_lte.cpp:984-1033, :1153-1189) with
//       dline_strength_calc_dT :116-143, dline_strength_calc_dVMR :86-114, line::ds_dT lbl_data.h:138-142
//       and the model derivatives lbl_lineshape_model.cpp:92-148
//   lbl_sum_jac_kernel       sum over lines of dX = ds F + s (dz + dz_fac z) dF per frequency and target
//       replaces band_shape::dT / dVMR (:475-507, :655-720) and compute_derivative (:1463-1561)
//
// The reference differentiates w(z) by a FORWARD FINITE DIFFERENCE (single_shape::dF, :250-268:
// dz = max(1e-4 |.|, 1e-4) per component, one extra w per pair); parity means reproducing that, not
// the analytic -2 z w + 2i/sqrt(pi).  Both w(z) and w(z + dz) go through the same
// region map as the reference (far closed forms, continued fraction, series).
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "catalog.hpp"
#include "faddeeva.cuh"
#include "lbl.hpp"
#include "lbl_model.cuh"

namespace ab200 {

struct cplx {
  double re, im;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cplx cscale(double s, cplx a) { return {s * a.re, s * a.im}; }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  const double d = 1.0 / (b.re * b.re + b.im * b.im);
  return {(a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d};
}

// w(x + i y) for any finite x, y >= 0, with the reference's region map (Faddeeva.cc:689-741):
// nu == 1 (x+y > 1e7, :708-720), nu == 2 (:721-725), continued fraction, series.  E1 = series_E1(y).
__device__ __forceinline__ cplx w_any(double x, double y, double E1) {
  const double ax = fabs(x);
  double wr, wi;
  if (ax + y > 1e7) {
    if (ax > y) {
      const double yax   = y / ax;
      const double denom = fad::ISPI / (ax + yax * y);
      wr = denom * yax;
      wi = denom;
    } else {
      const double xya   = ax / y;
      const double denom = fad::ISPI / (xya * ax + y);
      wr = denom;
      wi = denom * xya;
    }
  } else if (ax + y > FAR_LIMIT) {
    // nu == 2 in the reference's own operand order: dr = x^2 - y^2 - 1/2, di = 2 x y
    const double dr = ax * ax - y * y - 0.5, di = 2 * ax * y;
    const double denom = fad::ISPI / (dr * dr + di * di);
    wr = denom * (ax * di - y * dr);
    wi = denom * (ax * dr + y * di);
  } else if (cf_region(ax, y)) {
    w_cf(ax, y, wr, wi);
  } else {
    w_series(ax, y, E1, wr, wi);
  }
  return {wr, x < 0.0 ? -wi : wi};
}

// single_shape::all(f): z, F = w(z), dF by the forward finite difference of :250-268
__device__ __forceinline__ void z_F_dF(double x, double y, double E1, double E1p, cplx& z, cplx& F, cplx& dF) {
  z = {x, y};
  const cplx dz{fmax(1e-4 * fabs(x), 1e-4), fmax(1e-4 * fabs(y), 1e-4)};
  const double ax = fabs(x);
  if (ax + y > MID_LIMIT && ax + y <= FAR_LIMIT) {
    // four-term closed form (faddeeva.cuh w_mid) at both points: its truncation error is a smooth function of z, so
    // the difference quotient carries it at ~8x its relative size (1e-11 at the limit), not at 1/1e-4.  The branch is
    // taken on the base point for both evaluations so a pair never mixes two approximations.
    const double xd = x + dz.re;
    double wr, wi, vr, vi;
    w_mid(ax, y, wr, wi);
    w_mid(fabs(xd), y + dz.im, vr, vi);
    F = {wr, x < 0.0 ? -wi : wi};
    const cplx F2{vr, xd < 0.0 ? -vi : vi};
    dF = cdiv(csub(F2, F), dz);
    return;
  }
  F = w_any(x, y, E1);
  const cplx F2 = w_any(x + dz.re, y + dz.im, E1p);
  dF = cdiv(csub(F2, F), dz);
}

// single_shape::dT / dVMR, :310-323
__device__ __forceinline__ cplx dX(cplx s, cplx ds, cplx dzq, double dz_fac, cplx z, cplx F, cplx dF) {
  const cplx t = cadd(dzq, cscale(dz_fac, z));
  return cadd(cmul(ds, F), cmul(cmul(s, t), dF));
}

// ---------------------------------------------------------------------------
// per (sub-line, level, target) derivative records
//   jac  [lev][tile][q][2][TL][4] : (ds_re, ds_im, dz_re, dz_im) | (dz_fac, dcut_re, dcut_im, 0)
//   jcom [lev][tile][TL]          : E1(y + dy), the series constant of the displaced point
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TL) lbl_prepare_jac_kernel(PrepareParams p, JacPrepParams jp) {
  const int64_t tile = blockIdx.x;
  const int lev      = blockIdx.y;
  const int lane     = threadIdx.x;
  const int64_t slot = tile * TL + lane;
  const int64_t par  = p.sub_parent[slot];
  const double* rec  = p.prep + (int64_t(lev) * p.ntiles + tile) * tile_doubles();
  const uint8_t sfl  = p.sub_flags[slot];
  const double f0rec = rec[(0 * TL + lane) * REC_GROUP + 0];
  const double f0s   = (sfl & SUB_TWIN) ? -f0rec : f0rec;  // the real centre f0' (a mirror twin's record holds -f0')
  const double igd   = rec[(1 * TL + lane) * REC_GROUP + 1];
  const double y     = rec[(1 * TL + lane) * REC_GROUP + 2];
  const double s_re  = rec[(1 * TL + lane) * REC_GROUP + 3];
  const double E1    = rec[(2 * TL + lane) * REC_GROUP + 0];
  // real merged segments keep the line's cutoff in the s_im slot (s_im == 0 there)
  const double s_im  = p.tile_mode[tile] == 0 ? 0.0 : rec[(2 * TL + lane) * REC_GROUP + 1];
  const bool live    = par >= 0 && igd != 0.0;  // padding and inactive cutoff lines have igd == 0

  const double yp  = y + fmax(1e-4 * fabs(y), 1e-4);
  const double E1p = (live && yp <= 7.0) ? series_E1(yp) : 0.0;
  jp.jcom[(int64_t(lev) * p.ntiles + tile) * TL + lane] = E1p;

  const double T = p.T[lev], P = p.P[lev];
  const double cut = p.sub_cut[slot];
  for (int q = 0; q < jp.nq; q++) {
    double o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (live) {
      const int isot = p.line_isot[par];
      const int spec = p.isot_species[isot];
      LineModel lm{p.ls_offset, p.ls_species, p.ls_type, p.ls_X, par, p.T0[par], T, P, p.vmr + int64_t(lev) * p.n_species};
      const double G = lm.mix(AB200_VAR_G, false), Y = lm.mix(AB200_VAR_Y, false);
      const double f0c = p.f0[par];
      const double Q   = p.Q[int64_t(lev) * p.n_isot + isot];
      const double ex  = exp(-p.e0[par] / (cst::k * T));
      const double sl  = p.a[par] * p.gu[par] * ex / (f0c * f0c * f0c * Q);  // line::s, lbl_data.h:66-68
      const double r   = p.isorat[int64_t(lev) * p.n_isot + isot];
      const double x   = p.vmr[int64_t(lev) * p.n_species + spec];
      const double Sz  = p.sub_Sz[slot];
      const cplx lmc{1 + G, -Y};
      double dD0, dDV, dG0, dG, dY;
      cplx ds;
      double dz_fac;
      if (jp.kind[q] == AB200_TARGET_T) {
        dD0 = lm.mix(AB200_VAR_D0, true); dDV = lm.mix(AB200_VAR_DV, true); dG0 = lm.mix(AB200_VAR_G0, true);
        dG  = lm.mix(AB200_VAR_G, true);  dY  = lm.mix(AB200_VAR_Y, true);
        const double dQ  = jp.dQdT[int64_t(lev) * p.n_isot + isot];
        const double dsl = p.a[par] * p.gu[par] * (p.e0[par] * Q - cst::k * (T * T) * dQ) * ex /
                           (f0c * f0c * f0c * cst::k * (T * T) * (Q * Q));  // line::ds_dT, lbl_data.h:138-142
        const double df0 = dD0 + dDV;
        const cplx dlm{dG, -dY};
        // dline_strength_calc_dT, :116-143
        const cplx num = csub(csub(cscale(2 * T * f0s, cadd(cscale(sl, dlm), cscale(dsl, lmc))), cscale(2 * T * df0 * sl, lmc)),
                              cscale(f0s * sl, lmc));
        ds     = cscale(Sz * (cst::inv_sqrt_pi * igd * r * x) / (2 * T * f0s), num);
        dz_fac = (-2 * T * dD0 - 2 * T * dDV - f0s) / (2 * T * f0s);  // :1013-1015
      } else if (jp.kind[q] == AB200_TARGET_ISORAT) {
        // compute_derivative(SpeciesIsotope) :1526-1544: scl shape / isorat for the bands of that isotopologue; as a
        // record (ds = s / isorat, nothing else) it also works inside species-merged segments and with cutoffs
        dD0 = dDV = dG0 = dG = dY = 0.0;
        ds     = isot == jp.species[q] ? cscale(1.0 / r, {s_re, s_im}) : cplx{0.0, 0.0};
        dz_fac = 0.0;
      } else if (jp.kind[q] >= AB200_TARGET_LINE_F0) {
        // line targets (line_key): only the sub-lines of jp.line[q] have a record, set_filter :1192-1201
        dD0 = dDV = dG0 = dG = dY = 0.0;
        ds     = {0.0, 0.0};
        dz_fac = 0.0;
        if (par == jp.line[q]) {
          const cplx s{s_re, s_im};
          const double pre = cst::inv_sqrt_pi * igd * r * x;
          if (jp.kind[q] == AB200_TARGET_LINE_F0) {
            // df0_core_calc :1204-1238, dline_strength_calc_df0 :66-84 with line::ds_df0_s_ratio = -3 / f0 (lbl_data.h:118)
            const double dsl = (-3.0 / f0c) * sl;
            ds     = cscale(Sz * pre * (f0s * dsl - sl) / f0s, lmc);
            dz_fac = -1.0 / f0s;
            dD0    = 1.0;  // dz = -inv_gd below
          } else if (jp.kind[q] == AB200_TARGET_LINE_E0) {
            ds = cscale(-1.0 / (cst::k * T), s);  // de0_core_calc :1241-1265, ds_de0_s_ratio lbl_data.h:101-103
          } else if (jp.kind[q] == AB200_TARGET_LINE_A) {
            ds = cscale(1.0 / p.a[par], s);  // da_core_calc :1268-1290
          } else {
            const double d = lm.dmix_dX(jp.ls_var[q], jp.species[q], jp.coeff[q]);
            switch (jp.ls_var[q]) {
              case AB200_VAR_G0: dG0 = d; break;  // dG0_core_calc :1293-1317: dz = i inv_gd dG0_dX
              case AB200_VAR_D0:                  // dD0_core_calc :1320-1349
              case AB200_VAR_DV:                  // dDV_core_calc :1419-1450
                dz_fac = -d / f0s;
                ds     = cscale(-d / f0s, s);
                dD0    = d;
                break;
              case AB200_VAR_Y: ds = cmul({0.0, -d * Sz * pre}, {sl, 0.0}); break;  // dline_strength_calc_dY :38-50
              case AB200_VAR_G: ds = {Sz * pre * d * sl, 0.0}; break;               // dline_strength_calc_dG :52-64
              default: break;
            }
          }
        }
      } else if (jp.kind[q] >= AB200_TARGET_MAG_U) {
        // single_shape::dH, :305-307: s dz dF, dz = -inv_gd dH/dmag_c Splitting (:1071-1078); lines without Zeeman
        // splitting have dz = 0 and pol = no segments are skipped in the sum kernel's epilogue (:1484-1486)
        dD0 = dDV = dG0 = dG = dY = 0.0;
        ds     = {0.0, 0.0};
        dz_fac = 0.0;
      } else if (jp.kind[q] >= AB200_TARGET_WIND_U) {
        // single_shape::df, :275: s inv_gd dF(f) -- dX with ds = 0, dz = inv_gd, dz_fac = 0 (for a mirror twin too:
        // zm = inv_gd (f + f0') has the same slope)
        dD0 = dDV = dG0 = dG = dY = 0.0;
        ds     = {0.0, 0.0};
        dz_fac = 0.0;
      } else {
        const int ts = jp.species[q];
        dD0 = lm.dmix_dvmr(AB200_VAR_D0, ts); dDV = lm.dmix_dvmr(AB200_VAR_DV, ts); dG0 = lm.dmix_dvmr(AB200_VAR_G0, ts);
        dG  = lm.dmix_dvmr(AB200_VAR_G, ts);  dY  = lm.dmix_dvmr(AB200_VAR_Y, ts);
        const double df0 = dD0 + dDV;
        const cplx dlm{dG, -dY};
        // dline_strength_calc_dVMR, :86-114
        const double pre = -cst::inv_sqrt_pi * igd * r * sl;
        if (ts == spec) {
          ds = cscale(Sz * pre, csub(cscale(x * (df0 / f0s), lmc), cadd(cscale(x, dlm), lmc)));
        } else {
          ds = cscale(Sz * pre * x, csub(cscale(df0 / f0s, lmc), dlm));
        }
        dz_fac = -(dD0 + dDV) / f0s;  // :1176
      }
      cplx dzq{igd * -(dD0 + dDV), igd * dG0};
      if (jp.kind[q] >= AB200_TARGET_LINE_F0) {
        // dzq above is the whole of it (line targets, isotopologue ratio)
      } else if (jp.kind[q] >= AB200_TARGET_MAG_U) {
        const double dzc = p.sub_dzc[slot];
        dzq = {dzc == 0.0 ? 0.0 : -igd * dzc, 0.0};  // times mag_c / |mag| in the sum kernels' epilogue (one row for u, v, w)
      } else if (jp.kind[q] >= AB200_TARGET_WIND_U) {
        dzq = {igd, 0.0};
      }
      if (sfl & SUB_MIRRORED) {
        // mirrored dT / dVMR (lbl_lineshape_voigt_lte_mirrored.cpp:305-325): s (dz + dz_fac z_) (dFp + dFm) with
        // z_ = zp - zm = -2 inv_gd f0', independent of the frequency: fold it into dz and drop the x-proportional part,
        // for the line and for its twin alike
        dzq.re += dz_fac * (-2.0 * igd * f0s);
        dz_fac = 0.0;
      }
      o[0] = ds.re; o[1] = ds.im; o[2] = dzq.re; o[3] = dzq.im; o[4] = dz_fac;
      if (cut < DBL_MAX) {  // band_shape::dT(dcut, ...) at f0' + cutoff, :475-486
        cplx z, F, dF;
        z_F_dF(igd * ((f0s + cut) - f0s), y, E1, E1p, z, F, dF);
        const cplx dc = dX({s_re, s_im}, ds, dzq, dz_fac, z, F, dF);
        o[5] = dc.re; o[6] = dc.im;
      }
    }
    double* out = jp.jac + ((int64_t(lev) * p.ntiles + tile) * jp.nq + q) * (2 * TL * 4);
    double2* o0 = reinterpret_cast<double2*>(out + (0 * TL + lane) * 4);
    double2* o1 = reinterpret_cast<double2*>(out + (1 * TL + lane) * 4);
    o0[0] = make_double2(o[0], o[1]); o0[1] = make_double2(o[2], o[3]);
    o1[0] = make_double2(o[4], o[5]); o1[1] = make_double2(o[6], o[7]);
  }
}

// ---------------------------------------------------------------------------
// line sum with derivatives
// ---------------------------------------------------------------------------
constexpr int JAC_NT = 128;
constexpr int JAC_R  = 2;
constexpr int JAC_F_TILE = JAC_NT * JAC_R;
constexpr int JAC_Q  = 4;   // targets per pass
constexpr int JAC_BASE_FIELDS = 10;  // f0', igd, y, s_re, s_im, E1, E1p, cut_re, cut_im | y2, y^2+1/2, y2^2+1/2, y y2-1/2, s_re/sqrt(pi), 2y, 2y2
constexpr int JAC_Q_FIELDS    = 7;   // ds_re, ds_im, dz_re, dz_im, dz_fac, dcut_re, dcut_im | dz_im + dz_fac y

// dscl(f) of df_core_calc, :1040-1049
__device__ __forceinline__ double line_scale_df(double f, double T, double P) {
  constexpr double c = cst::c * cst::c / (8 * cst::pi);
  const double N = P / (cst::k * T);
  const double r = (cst::h * f) / (cst::k * T);
  return N * (r * exp(-r) - expm1(-r)) * c;
}

// dscl(f) of dt_core_calc, :990-1000
__device__ __forceinline__ double line_scale_dT(double f, double T, double P) {
  constexpr double c = cst::c * cst::c / (8 * cst::pi);
  const double N  = P / (cst::k * T);
  const double dN = -P / (cst::k * (T * T));
  const double r  = (cst::h * f) / (cst::k * T);
  return -f * (N * r * exp(-r) / T + dN * expm1(-r)) * c;
}
__device__ __forceinline__ double line_scale_v(double f, double T, double P) {
  constexpr double c = cst::c * cst::c / (8 * cst::pi);
  const double N     = P / (cst::k * T);
  const double r     = (cst::h * f) / (cst::k * T);
  return -N * f * expm1(-r) * c;
}

// nu == 2 closed form of the reference (Faddeeva.cc:721-725) for x = |Re z| >= 0, y >= 0; branch free
__device__ __forceinline__ cplx w_far2(double ax, double y) {
  const double dr = ax * ax - y * y - 0.5, di = 2 * ax * y;
  const double denom = fad::ISPI * fast_rcp(dr * dr + di * di);
  return {denom * (ax * di - y * dr), denom * (ax * dr + y * di)};
}
// z, F, dF for a pair that is far together with its displaced point (tile / line level guarantee)
__device__ __forceinline__ void z_F_dF_far(double x, double y, cplx& z, cplx& F, cplx& dF) {
  z = {x, y};
  const double ax = fabs(x);
  F = w_far2(ax, y);
  const cplx dz{fmax(1e-4 * ax, 1e-4), fmax(1e-4 * y, 1e-4)};
  const double x2 = x + dz.re;
  cplx F2 = w_far2(fabs(x2), y + dz.im);
  if (x < 0.0) F.im = -F.im;
  if (x2 < 0.0) F2.im = -F2.im;
  const double d = fast_rcp(dz.re * dz.re + dz.im * dz.im);
  const cplx n = csub(F2, F);
  dF = {(n.re * dz.re + n.im * dz.im) * d, (n.im * dz.re - n.re * dz.im) * d};
}

// Closed form of a far pair of a REAL line, shared by all targets.  With c = i/sqrt(pi), D1 = z^2 - 1/2,
// D2 = (z+dz)^2 - 1/2, P = D1 D2:
//   F = c z / D1 = c z D2 conj(P) / |P|^2,      dF = (F(z+dz) - F(z)) / dz = -c (z (z+dz) + 1/2) conj(P) / |P|^2
// i.e. the reference's forward difference (lbl_lineshape_voigt_lte.cpp:250-268) of its nu = 2 form evaluated without the
// subtraction, ONE reciprocal per pair.  G1 = sqrt(pi) Re F, G2 = sqrt(pi) Re dF, G3 = -sqrt(pi) Im dF, G4 = x G2; a target adds
//   Re(ds F + (dz + dz_fac z) s dF) = A_q G1 + B_q G2 + C_q G4 + D_q G3 (+ E_q G5, G5 = sqrt(pi) Im F)
// with A_q = Re ds / sqrt(pi), B_q = sp Re dz, C_q = sp dz_fac, D_q = sp Im(dz + dz_fac z), E_q = -Im ds / sqrt(pi),
// sp = s / sqrt(pi): 35 FP64 instructions per pair + 4 per target (38 + 5 where some Im ds != 0).
struct FarLine {  // y-only constants of a line
  double y, y2, ty, ty2, cy1, cy2, yy;
  __device__ __forceinline__ void from_y(double y_) {
    y = y_; y2 = y + fmax(1e-4 * fabs(y), 1e-4); ty = 2.0 * y; ty2 = 2.0 * y2;
    cy1 = y * y + 0.5; cy2 = y2 * y2 + 0.5; yy = y * y2 - 0.5;
  }
};
template <bool IM>
__device__ __forceinline__ void far_pair(const FarLine& c, double x, double x2, double& G1, double& G2, double& G3, double& G4,
                                         double& G5) {
  const double A  = __fma_rn(x, x, -c.cy1), B = __dmul_rn(c.ty, x);
  const double A2 = __fma_rn(x2, x2, -c.cy2), B2 = __dmul_rn(c.ty2, x2);
  const double Pr = __fma_rn(A, A2, -__dmul_rn(B, B2)), Pi = __fma_rn(A, B2, __dmul_rn(B, A2));
  const double Mr = __fma_rn(x, x2, -c.yy), Mi = __fma_rn(x, c.y2, __dmul_rn(c.y, x2));
  const double n  = far_rcp(__fma_rn(Pr, Pr, __dmul_rn(Pi, Pi)));
  const double N2 = __fma_rn(Mi, Pr, -__dmul_rn(Mr, Pi));
  const double N3 = __fma_rn(Mr, Pr, __dmul_rn(Mi, Pi));
  const double Wr = __fma_rn(x, A2, -__dmul_rn(c.y, B2)), Wi = __fma_rn(x, B2, __dmul_rn(c.y, A2));
  const double N1 = __fma_rn(Wr, Pi, -__dmul_rn(Wi, Pr));
  G1 = __dmul_rn(N1, n); G2 = __dmul_rn(N2, n); G3 = __dmul_rn(N3, n); G4 = __dmul_rn(x, G2);
  // sqrt(pi) Im F, only where a strength derivative has an imaginary part: the reference's d/dVMR of an ABSENT
  // Y or G model is the sum of the broadeners' VMRs when the line has no bath broadener (lbl_lineshape_model.cpp:112, sic)
  G5 = IM ? __dmul_rn(__fma_rn(Wr, Pr, __dmul_rn(Wi, Pi)), n) : 0.0;
}

// compute_derivative's last step (:1474-1561) for one computed target at one frequency: dpm += npm (.) d.  The three
// magnetic-field components share ONE computed row (single_shape::dH with dz = -inv_gd Splitting, :305-307, :1071-1078:
// the component only scales it by mag_c / |mag|, a per-level number) and so do the three wind components (single_shape::df,
// :275; the component enters through freq_wind_shift_jac in spectral_propmat_jacWindFix): out_row lists the requested ones.
__device__ __forceinline__ void jac_store_rows(const JacSumParams& jp, int cq, int kind, int lev, int pol, int64_t i, int64_t k_pitch,
                                               const double* __restrict__ npm, double f, double scl, double scl_df, cplx shape, cplx d) {
  const int32_t* rows = jp.out_row[cq];
  if (kind == AB200_TARGET_MAG_U) {
    // compute_derivative :1484-1513 with zeeman::scale(npm, dnpm, scl shape, scl dshape), lbl_zeeman.h:442-453
    if (pol == POL_NO) return;
    const cplx F = cscale(scl, shape);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (rows[c] < 0) continue;
      const double* __restrict__ dn = jp.dnpm + ((int64_t(lev) * 3 + c) * 4 + pol) * 7;
      const cplx dc = cscale(jp.mag_ratio[3 * lev + c], d);
      double* o = jp.dK + ((int64_t(lev) * jp.nrows + rows[c]) * k_pitch + i) * 7;
      o[0] += dn[0] * F.re + npm[0] * dc.re; o[1] += dn[1] * F.re + npm[1] * dc.re;
      o[2] += dn[2] * F.re + npm[2] * dc.re; o[3] += dn[3] * F.re + npm[3] * dc.re;
      o[4] += dn[4] * F.im + npm[4] * dc.im; o[5] += dn[5] * F.im + npm[5] * dc.im;
      o[6] += dn[6] * F.im + npm[6] * dc.im;
    }
    return;
  }
  if (kind == AB200_TARGET_WIND_U) {
    // compute_derivative :1514-1523, then spectral_propmat_jacWindFix (m_frequency_grid.cc:106-182): x * f * df_du
    d = cadd(d, cscale(scl_df, shape));
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (rows[c] < 0) continue;
      // wind_jac null: AB200_FLAG_WIND_ROWS_DF, the caller's agenda applies the fix
      const cplx dc = jp.wind_jac ? cscale(jp.wind_jac[3 * lev + c], cscale(f, d)) : d;
      double* o = jp.dK + ((int64_t(lev) * jp.nrows + rows[c]) * k_pitch + i) * 7;
      o[0] += npm[0] * dc.re; o[1] += npm[1] * dc.re; o[2] += npm[2] * dc.re; o[3] += npm[3] * dc.re;
      o[4] += npm[4] * dc.im; o[5] += npm[5] * dc.im; o[6] += npm[6] * dc.im;
    }
    return;
  }
  double* o = jp.dK + ((int64_t(lev) * jp.nrows + rows[0]) * k_pitch + i) * 7;
  o[0] += npm[0] * d.re; o[1] += npm[1] * d.re; o[2] += npm[2] * d.re; o[3] += npm[3] * d.re;
  o[4] += npm[4] * d.im; o[5] += npm[5] * d.im; o[6] += npm[6] * d.im;
}

// ---------------------------------------------------------------------------
// Very far pairs of REAL lines: one rational function per pair, shared by every target.
//
// In line space, zeta = u + i g (u = f - f0', g = G0), zeta2 = k u + i g2 the displaced point of the reference's
// forward difference (:250-268: dz = 1e-4 (|x|, y) component-wise, so k = 1 + 1e-4 for u > 0 and 1 - 1e-4 for u < 0;
// g2 = GD y2), h = GD^2 / 2:
//   F  = c z / (z^2 - 1/2)                          = c GD zeta2 / (zeta zeta2 - h (1 + dz/z))
//   dF = -c (z z2 + 1/2) / ((z^2 - 1/2)(z2^2 - 1/2)) = -c GD^2 / (zeta zeta2 - 3 h) (1 + O(h^2 / |zeta|^4))
// Both are written over the ONE denominator Pi = zeta zeta2 - h = (k U - a0) + i c1 u (U = u^2, a0 = g g2 + h,
// c1 = g2 + k g).  That is F itself (relative error 1e-4 h / |Pi| < 1e-12: rows that hold F only - isotopologue ratio,
// line strength - keep the forward model's accuracy), and the forward difference dF with a relative error
// 2 h / |Pi| = 1 / |z|^2 <= 6.9e-9 for |x| > VFAR_LIMIT - of a quantity that is itself a 1e-4 finite-difference
// approximation of w'(z).  With n = 1 / |Pi|^2 = 1 / (k^2 U^2 + b1 U + a0^2):
//   sqrt(pi) Re F  = GD (k^2 g U + g2 a0) n,       sqrt(pi) Im F = GD u (k^2 U - k a0 + g2 c1) n
//   sqrt(pi) Re dF = -GD^2 c1 u n,                 sqrt(pi) Im dF = -GD^2 (k U - a0) n
// so Re(ds F + s (dz + dz_fac z) dF) of a target is (alpha_q + gamma_q u + beta_q U (+ delta_q u U)) n with per-line
// coefficients, and the pair costs DADD u, DMUL U, 2 DFMA |Pi|^2, MUFU + 2 DFMA n, 2 DFMA for the forward shape and
// 3 DFMA per target: 8 + 3 NQ FP64-pipe instructions (14 for T + VMR) against 35 + 4 NQ of far_pair and 7 of the
// forward kernel.
//
// WHICH pairs take this form is a property of the pair alone: |x| > VFAR_LIMIT with x = fl(igd fl(f - f0')), in tiles
// without cutoffs of real-line segments.  lbl_sum_jac_vfar_kernel sums exactly those pairs, lbl_sum_jac_kernel exactly the
// others, so a Jacobian row at a frequency does not depend on block boundaries or on the shard the frequency fell in.
// The tile summaries only decide how much testing a (tile, block) needs: every pair in (dist igd_min > VFAR_LIMIT: no
// test, one sign), no pair possible (skipped), or mixed (one predicated pass per sign of u).
// ---------------------------------------------------------------------------
constexpr double VFAR_LIMIT = 1.2e4;
constexpr int VF_NT = 128;
__device__ __forceinline__ bool vfar_all(const double* __restrict__ s4, double dist) {
  return dist > 0.0 && s4[2] * dist * (1.0 - 2e-4) > VFAR_LIMIT;
}
// the epilogue's frequency factors, out of line: inlined per frequency and target kind they are most of the kernel's code
__device__ __noinline__ void vf_scales(double f, double T, double P, bool need_T, bool need_df, double& scl, double& scl_dT,
                                       double& scl_df) {
  scl    = line_scale_v(f, T, P);
  scl_dT = need_T ? line_scale_dT(f, T, P) : 0.0;
  scl_df = need_df ? line_scale_df(f, T, P) : 0.0;
}
// every pair of a line with the frequencies of [fblk_min, fblk_max] passes the pair test (block-uniform shortcut)
__device__ __forceinline__ bool line_all_vfar(double f0s, double igd, double fblk_min, double fblk_max) {
  const double dl = fmax(fblk_min - f0s, f0s - fblk_max);
  return dl > 0.0 && igd * dl * (1.0 - 2e-4) > VFAR_LIMIT;
}
constexpr int vf_stride(int nq) { return (6 + 4 * nq + 1) & ~1; }  // doubles per staged line: f0', b1, a0^2, alpha0, beta0 | 4 per target | +-igd

#ifndef VF_MB4
#define VF_MB4 4
#endif
#ifndef VF_MB2
#define VF_MB2 5
#endif
#ifndef VF_UNROLL
#define VF_UNROLL 1
#endif
constexpr int VF_UNROLL_N = VF_UNROLL;
template <int NQ, int R>
__global__ void __launch_bounds__(VF_NT, (R == 4 ? VF_MB4 : VF_MB2)) lbl_sum_jac_vfar_kernel(SumParams p, JacSumParams jp) {
  constexpr int S = vf_stride(NQ);
  constexpr int F_TILE = VF_NT * R;
  extern __shared__ __align__(16) double sm[];  // [TL][S]
  __shared__ int wcnt[VF_NT / 32];
  const int tid = threadIdx.x;
  const int lev = blockIdx.y;
  const int64_t fblk = int64_t(blockIdx.x) * F_TILE;
  const double* __restrict__ fg = p.f + int64_t(lev) * p.f_stride;
  const double ffac = p.ffac[lev];
  double f[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int64_t i = fblk + r * VF_NT + tid;
    f[r] = ffac * fg[i < p.nf ? i : p.nf - 1];
  }
  const double fblk_min = ffac * fg[fblk];
  const double fblk_max = ffac * fg[(fblk + F_TILE - 1 < p.nf) ? fblk + F_TILE - 1 : p.nf - 1];
  const double* __restrict__ prep = p.prep + int64_t(lev) * p.ntiles * tile_doubles();
  const double* __restrict__ summ = p.summary + int64_t(lev) * p.ntiles * SUMMARY_DOUBLES;
  const double T = p.T[lev], P = p.P[lev];

  for (int is = 0; is < p.nsegs; is++) {
    const SegmentDev seg = p.segs[is];
    double shape[R], acc[NQ][R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      shape[r] = 0.0;
#pragma unroll
      for (int q = 0; q < NQ; q++) acc[q][r] = 0.0;
    }
    bool any_tile = false;
    for (int64_t t = seg.tile_begin; t < seg.tile_end; t++) {
      const double* __restrict__ s4 = summ + t * SUMMARY_DOUBLES;
      if (s4[0] > s4[1] || s4[4] < DBL_MAX) continue;  // no contributing line | lines with cutoffs: the other kernel
      const double dist = fmax(fblk_min - s4[1], s4[0] - fblk_max);
      const bool all = vfar_all(s4, dist);
      // u > 0 needs a line below a frequency of the block, u < 0 one above; |x| <= max igd * that distance
      const bool side_pos = all ? s4[1] < fblk_min : s4[7] * (fblk_max - s4[0]) * (1.0 + 1e-9) > VFAR_LIMIT;
      const bool side_neg = all ? s4[0] > fblk_max : s4[7] * (s4[1] - fblk_min) * (1.0 + 1e-9) > VFAR_LIMIT;
      const int count = p.tile_count[t];
      const double* g = prep + t * tile_doubles();
      const double* jt = jp.jac + ((int64_t(lev) * p.ntiles + t) * jp.nq + jp.q0) * (2 * TL * 4);
#pragma unroll 1
      for (int side = 0; side < 2; side++) {
        if (!(side == 0 ? side_pos : side_neg)) continue;
        any_tile = true;
        const double sg = side == 0 ? 1.0 : -1.0;
        const double k  = 1.0 + sg * 1e-4;
        const double k2 = k * k;
        // stage the lines of this side that can hold a very far pair in this block, in line order (in tiles that are
        // wholly very far: every contributing line)
        const double edge = side == 0 ? fblk_max : fblk_min;
        int nkeep = 0;
        bool any_im = false;
        __syncthreads();  // previous stage consumed
#pragma unroll 1
        for (int c0 = 0; c0 < count; c0 += VF_NT) {
          const int l = c0 + tid;
          double f0s = 0.0, igd = 0.0, y = 0.0, s_re = 0.0;
          if (l < count) {
            f0s = g[(0 * TL + l) * REC_GROUP];
            const double2 m = *reinterpret_cast<const double2*>(g + (1 * TL + l) * REC_GROUP);      // B1, igd
            const double2 n = *reinterpret_cast<const double2*>(g + (1 * TL + l) * REC_GROUP + 2);  // y, s_re
            igd = m.y; y = n.x; s_re = n.y;
          }
          const bool keep = igd != 0.0 && (all || sg * igd * (edge - f0s) * (1.0 + 1e-9) > VFAR_LIMIT);
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          if ((tid & 31) == 0) wcnt[tid >> 5] = __popc(bal);
          __syncthreads();
          int pos = nkeep + __popc(bal & ((1u << (tid & 31)) - 1u)), tot = 0;
#pragma unroll
          for (int w = 0; w < VF_NT / 32; w++) {
            pos += w < (tid >> 5) ? wcnt[w] : 0;
            tot += wcnt[w];
          }
          if (keep) {
            double* o = sm + pos * S;
            const double GD = 1.0 / igd, y2 = y + fmax(1e-4 * fabs(y), 1e-4);
            const double gg = y * GD, g2 = y2 * GD, h = 0.5 * GD * GD;
            const double a0 = gg * g2 + h, c1 = g2 + k * gg;
            const double sp = s_re * cst::inv_sqrt_pi;
            o[0] = f0s;
            o[1] = g2 * g2 + k2 * gg * gg - 2.0 * k * h;  // b1 = c1^2 - 2 k a0
            o[2] = a0 * a0;
            o[3] = sp * GD * g2 * a0;                     // alpha0
            o[4] = sp * GD * k2 * gg;                     // beta0
            o[5 + 4 * NQ] = sg * igd;
            const double spG2 = sp * GD * GD;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
              const double2* j0 = reinterpret_cast<const double2*>(jt + q * (2 * TL * 4) + (0 * TL + l) * 4);
              const double2 u0 = j0[0], u1 = j0[1];  // ds_re, ds_im | dz_re, dz_im
              const double dz_fac = jt[q * (2 * TL * 4) + (1 * TL + l) * 4];
              const double dsr = u0.x * cst::inv_sqrt_pi * GD, dsi = u0.y * cst::inv_sqrt_pi * GD;
              const double t_im = u1.y + dz_fac * y;
              any_im |= u0.y != 0.0;
              o[5 + 4 * q + 0] = dsr * g2 * a0 - spG2 * t_im * a0;                       // alpha
              o[5 + 4 * q + 1] = -spG2 * u1.x * c1 - dsi * (g2 * c1 - k * a0);           // gamma
              o[5 + 4 * q + 2] = dsr * k2 * gg + spG2 * (k * t_im - dz_fac * igd * c1);  // beta
              o[5 + 4 * q + 3] = -dsi * k2;                                               // delta (only where Im ds != 0)
            }
          }
          nkeep += tot;
          __syncthreads();  // wcnt may be rewritten
        }
        const bool tile_im = __syncthreads_or(any_im) != 0;
        auto loop = [&](auto im_tag, auto test_tag) {
          constexpr bool IM = decltype(im_tag)::value;
          constexpr bool TEST = decltype(test_tag)::value;
#pragma unroll VF_UNROLL_N
          for (int l = 0; l < nkeep; l++) {
            const double* __restrict__ o = sm + l * S;
            const double2 c01 = *reinterpret_cast<const double2*>(o), c23 = *reinterpret_cast<const double2*>(o + 2);
            const double b0 = o[4];
            const double sigd = TEST ? o[5 + 4 * NQ] : 0.0;
            double cq[NQ][4];
#pragma unroll
            for (int q = 0; q < NQ; q++) {
              cq[q][0] = o[5 + 4 * q]; cq[q][1] = o[6 + 4 * q]; cq[q][2] = o[7 + 4 * q];
              cq[q][3] = IM ? o[8 + 4 * q] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
              const double u = __dsub_rn(f[r], c01.x);
              const double U = __dmul_rn(u, u);
              double n = far_rcp(__fma_rn(__fma_rn(k2, U, c01.y), U, c23.x));
              if (TEST) n = __dmul_rn(sigd, u) > VFAR_LIMIT ? n : 0.0;  // the pair test: |fl(igd fl(f - f0'))| > VFAR_LIMIT
              shape[r] = __fma_rn(__fma_rn(b0, U, c23.y), n, shape[r]);
#pragma unroll
              for (int q = 0; q < NQ; q++) {
                double v = __fma_rn(cq[q][2], U, cq[q][0]);
                if (IM) v = __fma_rn(__dmul_rn(cq[q][3], u), U, v);
                acc[q][r] = __fma_rn(__fma_rn(cq[q][1], u, v), n, acc[q][r]);
              }
            }
          }
        };
        if (all) {
          if (tile_im) loop(std::true_type{}, std::false_type{});
          else loop(std::false_type{}, std::false_type{});
        } else {
          if (tile_im) loop(std::true_type{}, std::true_type{});
          else loop(std::false_type{}, std::true_type{});
        }
      }
    }
    if (!any_tile) continue;
    // compute_derivative :1474-1481, :1553-1560 (see lbl_sum_jac_kernel's epilogue): real strengths, only Re is used
    const double* __restrict__ npm = p.npm + (int64_t(lev) * 4 + seg.pol) * 7;
    if (npm[0] == 0 && npm[1] == 0 && npm[2] == 0 && npm[3] == 0) continue;
    bool need_T = false, need_df = false;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      need_T |= jp.kind[jp.q0 + q] == AB200_TARGET_T;
      need_df |= jp.kind[jp.q0 + q] == AB200_TARGET_WIND_U;
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int64_t i = fblk + r * VF_NT + tid;
      if (i >= p.nf) continue;
      double scl, scl_dT, scl_df;
      vf_scales(f[r], T, P, need_T, need_df, scl, scl_dT, scl_df);
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        double d = scl * acc[q][r];
        const int kind = jp.kind[jp.q0 + q];
        const int32_t* rows = jp.out_row[jp.q0 + q];
        if (kind == AB200_TARGET_T) d += scl_dT * shape[r];
        if (kind == AB200_TARGET_MAG_U) continue;  // pol = no: :1484-1486
        if (kind == AB200_TARGET_WIND_U) {         // one computed row d/df for the requested components
          d += scl_df * shape[r];
#pragma unroll
          for (int c = 0; c < 3; c++) {
            if (rows[c] < 0) continue;
            const double dc = jp.wind_jac ? jp.wind_jac[3 * lev + c] * (f[r] * d) : d;
            double* o = jp.dK + ((int64_t(lev) * jp.nrows + rows[c]) * p.k_pitch + i) * 7;
            o[0] += npm[0] * dc; o[1] += npm[1] * dc; o[2] += npm[2] * dc; o[3] += npm[3] * dc;
          }
          continue;
        }
        double* o = jp.dK + ((int64_t(lev) * jp.nrows + rows[0]) * p.k_pitch + i) * 7;
        o[0] += npm[0] * d; o[1] += npm[1] * d; o[2] += npm[2] * d; o[3] += npm[3] * d;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// The same for COMPLEX lines (Zeeman components, line mixing): strength, strength derivative and both parts of F and dF
// enter, so each of Re and Im of s F and of every target's dX is a cubic (alpha + gamma u + beta U + delta u U) n:
//   sqrt(pi) F  = [(A0 + A1 U) + i u (B0 + B1 U)] n,   A0 = GD g2 a0, A1 = GD k^2 g, B0 = GD (g2 c1 - k a0), B1 = GD k^2
//   sqrt(pi) dF = [C1 u + i (D0 + D1 U)] n,            C1 = -GD^2 c1, D0 = GD^2 a0, D1 = -GD^2 k
//   dX = ds F + s (dz + dz_fac z) dF,   s (dz + dz_fac z) = (w0 + w1 u),  w0 = s (dz + i dz_fac y), w1 = s dz_fac igd
// 7 shared FP64 instructions per pair + 8 for the shape + 8 per target (far_cplx path: ~45 + 12 per target).  The pair rule
// and the split with lbl_sum_jac_kernel are those of the real kernel above; a tile is staged in chunks of VC_CH lines.
// ---------------------------------------------------------------------------
constexpr int VC_CH = 128;
constexpr int vc_stride(int nq) { return 12 + 8 * nq; }

template <int NQ, int R>
__global__ void __launch_bounds__(VF_NT, (NQ <= 2 ? 4 : 3)) lbl_sum_jac_vfar_cplx_kernel(SumParams p, JacSumParams jp) {
  constexpr int S = vc_stride(NQ);
  constexpr int F_TILE = VF_NT * R;
  extern __shared__ __align__(16) double sm[];  // [VC_CH][S]
  __shared__ int wcnt[VF_NT / 32];
  const int tid = threadIdx.x;
  const int lev = blockIdx.y;
  const int64_t fblk = int64_t(blockIdx.x) * F_TILE;
  const double* __restrict__ fg = p.f + int64_t(lev) * p.f_stride;
  const double ffac = p.ffac[lev];
  double f[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int64_t i = fblk + r * VF_NT + tid;
    f[r] = ffac * fg[i < p.nf ? i : p.nf - 1];
  }
  const double fblk_min = ffac * fg[fblk];
  const double fblk_max = ffac * fg[(fblk + F_TILE - 1 < p.nf) ? fblk + F_TILE - 1 : p.nf - 1];
  const double* __restrict__ prep = p.prep + int64_t(lev) * p.ntiles * tile_doubles();
  const double* __restrict__ summ = p.summary + int64_t(lev) * p.ntiles * SUMMARY_DOUBLES;
  const double T = p.T[lev], P = p.P[lev];

  for (int is = 0; is < p.nsegs; is++) {
    const SegmentDev seg = p.segs[is];
    if (seg.has_cutoff) continue;  // segments with a cutoff: the other kernel
    cplx shape[R], acc[NQ][R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      shape[r] = {0.0, 0.0};
#pragma unroll
      for (int q = 0; q < NQ; q++) acc[q][r] = {0.0, 0.0};
    }
    bool any_tile = false;
    for (int64_t t = seg.tile_begin; t < seg.tile_end; t++) {
      const double* __restrict__ s4 = summ + t * SUMMARY_DOUBLES;
      if (s4[0] > s4[1]) continue;
      const double dist = fmax(fblk_min - s4[1], s4[0] - fblk_max);
      const bool all = vfar_all(s4, dist);
      const bool side_pos = all ? s4[1] < fblk_min : s4[7] * (fblk_max - s4[0]) * (1.0 + 1e-9) > VFAR_LIMIT;
      const bool side_neg = all ? s4[0] > fblk_max : s4[7] * (s4[1] - fblk_min) * (1.0 + 1e-9) > VFAR_LIMIT;
      const int count = p.tile_count[t];
      const double* g = prep + t * tile_doubles();
      const double* jt = jp.jac + ((int64_t(lev) * p.ntiles + t) * jp.nq + jp.q0) * (2 * TL * 4);
#pragma unroll 1
      for (int side = 0; side < 2; side++) {
        if (!(side == 0 ? side_pos : side_neg)) continue;
        any_tile = true;
        const double sg = side == 0 ? 1.0 : -1.0;
        const double k  = 1.0 + sg * 1e-4;
        const double k2 = k * k;
        const double edge = side == 0 ? fblk_max : fblk_min;
#pragma unroll 1
        for (int c0 = 0; c0 < count; c0 += VC_CH) {
          static_assert(VC_CH == VF_NT, "one line per thread and chunk");
          const int l = c0 + tid;
          double f0s = 0.0, igd = 0.0, y = 0.0, s_re = 0.0, s_im = 0.0;
          if (l < count) {
            f0s = g[(0 * TL + l) * REC_GROUP];
            const double2 m = *reinterpret_cast<const double2*>(g + (1 * TL + l) * REC_GROUP);      // B1, igd
            const double2 n = *reinterpret_cast<const double2*>(g + (1 * TL + l) * REC_GROUP + 2);  // y, s_re
            igd = m.y; y = n.x; s_re = n.y;
            s_im = g[(2 * TL + l) * REC_GROUP + 1];
          }
          // the lines of this side that can hold a very far pair in this block, compacted in line order
          const bool keep = igd != 0.0 && (all || sg * igd * (edge - f0s) * (1.0 + 1e-9) > VFAR_LIMIT);
          const unsigned bal = __ballot_sync(0xffffffffu, keep);
          __syncthreads();  // previous chunk consumed, wcnt free
          if ((tid & 31) == 0) wcnt[tid >> 5] = __popc(bal);
          __syncthreads();
          int pos = __popc(bal & ((1u << (tid & 31)) - 1u)), nc = 0;
#pragma unroll
          for (int w = 0; w < VF_NT / 32; w++) {
            pos += w < (tid >> 5) ? wcnt[w] : 0;
            nc += wcnt[w];
          }
          if (keep) {
            double* o = sm + pos * S;
            const double GD = 1.0 / igd, y2 = y + fmax(1e-4 * fabs(y), 1e-4);
            const double gg = y * GD, g2 = y2 * GD, h = 0.5 * GD * GD;
            const double a0 = gg * g2 + h, c1 = g2 + k * gg;
            const double A0 = GD * g2 * a0, A1 = GD * k2 * gg, B0 = GD * (g2 * c1 - k * a0), B1 = GD * k2;
            const double C1 = -GD * GD * c1, D0 = GD * GD * a0, D1 = -GD * GD * k;
            const double spr = s_re * cst::inv_sqrt_pi, spi = s_im * cst::inv_sqrt_pi;
            o[0] = f0s;
            o[1] = g2 * g2 + k2 * gg * gg - 2.0 * k * h;
            o[2] = a0 * a0;
            o[3] = sg * igd;
            o[4] = spr * A0; o[5] = -spi * B0; o[6] = spr * A1; o[7] = -spi * B1;   // Re s F: alpha, gamma, beta, delta
            o[8] = spi * A0; o[9] = spr * B0;  o[10] = spi * A1; o[11] = spr * B1;   // Im s F
#pragma unroll
            for (int q = 0; q < NQ; q++) {
              const double2* j0 = reinterpret_cast<const double2*>(jt + q * (2 * TL * 4) + (0 * TL + l) * 4);
              const double2 u0 = j0[0], u1 = j0[1];  // ds_re, ds_im | dz_re, dz_im
              const double dz_fac = jt[q * (2 * TL * 4) + (1 * TL + l) * 4];
              const double dr = u0.x * cst::inv_sqrt_pi, di = u0.y * cst::inv_sqrt_pi;
              const double t0 = u1.x, t1 = dz_fac * igd, ti = u1.y + dz_fac * y;
              const double w0r = spr * t0 - spi * ti, w1r = spr * t1, w0i = spi * t0 + spr * ti, w1i = spi * t1;
              double* oq = o + 12 + 8 * q;
              oq[0] = dr * A0 - w0i * D0;
              oq[1] = -di * B0 + w0r * C1 - w1i * D0;
              oq[2] = dr * A1 + w1r * C1 - w0i * D1;
              oq[3] = -di * B1 - w1i * D1;
              oq[4] = di * A0 + w0r * D0;
              oq[5] = dr * B0 + w1r * D0 + w0i * C1;
              oq[6] = di * A1 + w0r * D1 + w1i * C1;
              oq[7] = dr * B1 + w1r * D1;
            }
          }
          __syncthreads();
          auto loop = [&](auto test_tag) {
            constexpr bool TEST = decltype(test_tag)::value;
#pragma unroll 1
            for (int lc = 0; lc < nc; lc++) {
              const double* __restrict__ o = sm + lc * S;
              const double2 c01 = *reinterpret_cast<const double2*>(o), c23 = *reinterpret_cast<const double2*>(o + 2);
#pragma unroll
              for (int r = 0; r < R; r++) {
                const double u = __dsub_rn(f[r], c01.x);
                const double U = __dmul_rn(u, u);
                const double W = __dmul_rn(u, U);
                double n = far_rcp(__fma_rn(__fma_rn(k2, U, c01.y), U, c23.x));
                if (TEST) n = __dmul_rn(c23.y, u) > VFAR_LIMIT ? n : 0.0;
                auto cubic = [&](const double* c) { return __fma_rn(c[3], W, __fma_rn(c[2], U, __fma_rn(c[1], u, c[0]))); };
                shape[r].re = __fma_rn(cubic(o + 4), n, shape[r].re);
                shape[r].im = __fma_rn(cubic(o + 8), n, shape[r].im);
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                  acc[q][r].re = __fma_rn(cubic(o + 12 + 8 * q), n, acc[q][r].re);
                  acc[q][r].im = __fma_rn(cubic(o + 16 + 8 * q), n, acc[q][r].im);
                }
              }
            }
          };
          if (all) loop(std::false_type{});
          else loop(std::true_type{});
        }
      }
    }
    if (!any_tile) continue;
    // epilogue of lbl_sum_jac_kernel (compute_derivative :1474-1561) on this kernel's partial sums: it is linear in them
    const double* __restrict__ npm = p.npm + (int64_t(lev) * 4 + seg.pol) * 7;
    const bool all_zero = npm[0] == 0 && npm[1] == 0 && npm[2] == 0 && npm[3] == 0 && npm[4] == 0 && npm[5] == 0 &&
                          npm[6] == 0;
    if (all_zero) continue;
    bool need_T = false, need_df = false;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      need_T |= jp.kind[jp.q0 + q] == AB200_TARGET_T;
      need_df |= jp.kind[jp.q0 + q] == AB200_TARGET_WIND_U;
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int64_t i = fblk + r * VF_NT + tid;
      if (i >= p.nf) continue;
      double scl, scl_dT, scl_df;
      vf_scales(f[r], T, P, need_T, need_df, scl, scl_dT, scl_df);
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        cplx d = cscale(scl, acc[q][r]);
        const int kind = jp.kind[jp.q0 + q];
        if (kind == AB200_TARGET_T) d = cadd(d, cscale(scl_dT, shape[r]));
        jac_store_rows(jp, jp.q0 + q, kind, lev, seg.pol, i, p.k_pitch, npm, f[r], scl, scl_df, shape[r], d);
      }
    }
  }
}

// CTAs per SM the shared-memory footprint allows: (10 + 7 NQ) x 2 KB -> 34 / 48 / 62 / 76 KB
// EXT: the pass holds a wind or magnetic-field target (their epilogue costs the temperature / VMR passes 6 % in
// registers and spills when it is compiled in, so those keep an instantiation without it)
template <int NQ, bool EXT>
__global__ void __launch_bounds__(JAC_NT, (NQ <= 2 ? 4 : NQ == 3 ? 3 : 2)) lbl_sum_jac_kernel(SumParams p, JacSumParams jp) {
  extern __shared__ __align__(16) double sm[];  // [JAC_BASE_FIELDS + NQ * JAC_Q_FIELDS][TL]
  double* const sb = sm;
  double* const sq = sm + JAC_BASE_FIELDS * TL;
  const int tid = threadIdx.x;
  const int lev = blockIdx.y;
  const int64_t fblk = int64_t(blockIdx.x) * JAC_F_TILE;
  const double* __restrict__ fg = p.f + int64_t(lev) * p.f_stride;
  const double ffac = p.ffac[lev];
  double f[JAC_R];
#pragma unroll
  for (int r = 0; r < JAC_R; r++) {
    const int64_t i = fblk + r * JAC_NT + tid;
    f[r] = ffac * fg[i < p.nf ? i : p.nf - 1];
  }
  const double fblk_min = ffac * fg[fblk];
  const double fblk_max = ffac * fg[(fblk + JAC_F_TILE - 1 < p.nf) ? fblk + JAC_F_TILE - 1 : p.nf - 1];
  const double* __restrict__ prep = p.prep + int64_t(lev) * p.ntiles * tile_doubles();
  const double* __restrict__ summ = p.summary + int64_t(lev) * p.ntiles * SUMMARY_DOUBLES;
  const double* __restrict__ jcom = jp.jcom + int64_t(lev) * p.ntiles * TL;
  const double T = p.T[lev], P = p.P[lev];

  // a pass of line targets only visits the tiles that hold those lines: everything else has all-zero records and the
  // forward shape is not needed (no dscl term)
  bool line_only = true;
#pragma unroll
  for (int q = 0; q < NQ; q++) line_only &= jp.kind[jp.q0 + q] >= AB200_TARGET_LINE_F0 && jp.kind[jp.q0 + q] <= AB200_TARGET_LINE_LS;

  for (int is = 0; is < p.nsegs; is++) {
    const SegmentDev seg = p.segs[is];
    const double cutoff  = seg.has_cutoff ? seg.cutoff : DBL_MAX;
    int64_t t_lo = seg.tile_begin, t_hi = seg.tile_end;
    if (line_only) {
      int64_t lo = INT64_MAX, hi = -1;
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        const int64_t a = jp.line_tiles[jp.q0 + q][seg.pol][0], b = jp.line_tiles[jp.q0 + q][seg.pol][1];
        if (a >= 0) { lo = a < lo ? a : lo; hi = b > hi ? b : hi; }
      }
      t_lo = lo > t_lo ? lo : t_lo;
      t_hi = hi < t_hi ? hi : t_hi;
      if (t_lo >= t_hi) continue;  // none of the pass's lines in this segment
    }
    cplx shape[JAC_R], acc[NQ][JAC_R];
#pragma unroll
    for (int r = 0; r < JAC_R; r++) {
      shape[r] = {0.0, 0.0};
#pragma unroll
      for (int q = 0; q < NQ; q++) acc[q][r] = {0.0, 0.0};
    }
    for (int64_t t = t_lo; t < t_hi; t++) {
      if (line_only) {  // [t_lo, t_hi) is only the hull of the lines' tile ranges
        bool hit = false;
#pragma unroll
        for (int q = 0; q < NQ; q++)
          hit |= t >= jp.line_tiles[jp.q0 + q][seg.pol][0] && t < jp.line_tiles[jp.q0 + q][seg.pol][1];
        if (!hit) continue;
      }
      const double* __restrict__ s4 = summ + t * SUMMARY_DOUBLES;
      if (s4[0] > s4[1]) continue;  // no contributing line (CTA uniform)
      const double dist = fmax(0.0, fmax(fblk_min - s4[1], s4[0] - fblk_max));
      // real merged segments: per-line cutoffs (s4[4], s4[5] = min, max over the tile; DBL_MAX without)
      const double tile_cut    = jp.real_lines ? s4[5] : cutoff;
      const bool tile_has_cut  = jp.real_lines ? s4[4] < DBL_MAX : seg.has_cutoff != 0;
      if (dist > tile_cut * (1.0 + 1e-9)) continue;
      // pairs with |x| > VFAR_LIMIT of cutoff-free real-line tiles belong to lbl_sum_jac_vfar_kernel
      const bool vf_pairs = jp.skip_vfar && !tile_has_cut;
      if (vf_pairs && vfar_all(s4, fmax(fblk_min - s4[1], s4[0] - fblk_max))) continue;
      const int count = p.tile_count[t];
      // far for every pair of the tile and its displaced point (x shrinks by at most 1e-4 |x|); CTA uniform
      const bool far = !tile_has_cut && s4[2] * dist * (1.0 - 2e-4) + s4[3] > FAR_LIMIT * (1.0 + 1e-9);
      const bool far_closed = far && jp.real_lines;
      const bool far_cplx   = far && !jp.real_lines && jp.pair_far;  // complex lines (Zeeman, line mixing): see far_cplx_loop
      __syncthreads();  // previous tile fully consumed
      // stage the tile: the records of K1 + the derivative records of this pass, SoA in shared memory
      const double* g = prep + t * tile_doubles();
      const double* jt = jp.jac + ((int64_t(lev) * p.ntiles + t) * jp.nq + jp.q0) * (2 * TL * 4);
      bool any_im = false;
      if (far_closed) {
        // only what the closed form reads, with the per-line factors folded in (see far_pair)
        for (int l = tid; l < count; l += JAC_NT) {
          const double f0s = g[(0 * TL + l) * REC_GROUP];
          const double2 m = *reinterpret_cast<const double2*>(g + (1 * TL + l) * REC_GROUP);      // B1, igd
          const double2 n = *reinterpret_cast<const double2*>(g + (1 * TL + l) * REC_GROUP + 2);  // y, s_re
          const double y = n.x, y2 = y + fmax(1e-4 * fabs(y), 1e-4), sp = n.y * cst::inv_sqrt_pi;
          // far layout of the shared slots: 0 f0', 1 igd, 2 y | 3 y2, 4 y^2+1/2, 5 y2^2+1/2, 6 y y2-1/2, 7 sp, 8 2y, 9 2y2
          sb[0 * TL + l] = f0s; sb[1 * TL + l] = m.y; sb[2 * TL + l] = y; sb[3 * TL + l] = y2;
          sb[4 * TL + l] = y * y + 0.5; sb[5 * TL + l] = y2 * y2 + 0.5; sb[6 * TL + l] = y * y2 - 0.5;
          sb[7 * TL + l] = sp; sb[8 * TL + l] = 2.0 * y; sb[9 * TL + l] = 2.0 * y2;
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            const double2* j0 = reinterpret_cast<const double2*>(jt + q * (2 * TL * 4) + (0 * TL + l) * 4);
            const double2 u0 = j0[0], u1 = j0[1];                      // ds_re, ds_im | dz_re, dz_im
            const double dz_fac = jt[q * (2 * TL * 4) + (1 * TL + l) * 4];
            double* o = sq + q * JAC_Q_FIELDS * TL;
            o[0 * TL + l] = u0.x * cst::inv_sqrt_pi;                   // A_q: Re ds / sqrt(pi)
            o[1 * TL + l] = -u0.y * cst::inv_sqrt_pi;                  // E_q: -Im ds / sqrt(pi)
            any_im |= u0.y != 0.0;
            o[2 * TL + l] = sp * u1.x;                                 // B_q
            o[4 * TL + l] = sp * dz_fac;                               // C_q
            o[3 * TL + l] = sp * (u1.y + dz_fac * y);                  // D_q: Im(dz + dz_fac z) does not depend on f
          }
        }
      } else if (far_cplx) {
        // complex strengths: shape and dX need both parts.  With sp = s / sqrt(pi), tq = dz + dz_fac z = (dz + i dz_fac y)
        // + dz_fac x:  dX = ds F + (a + b x) dF,  a = s (dz + i dz_fac y),  b = s dz_fac  (all / sqrt(pi)), so a target
        // costs 12 FMAs on the shared numerators G1..G6 of the pair
        for (int l = tid; l < count; l += JAC_NT) {
          const double f0s = g[(0 * TL + l) * REC_GROUP];
          const double2 m = *reinterpret_cast<const double2*>(g + (1 * TL + l) * REC_GROUP);      // B1, igd
          const double2 n = *reinterpret_cast<const double2*>(g + (1 * TL + l) * REC_GROUP + 2);  // y, s_re
          const double s_im = g[(2 * TL + l) * REC_GROUP + 1];
          const double y = n.x, y2 = y + fmax(1e-4 * fabs(y), 1e-4);
          const cplx sp{n.y * cst::inv_sqrt_pi, s_im * cst::inv_sqrt_pi};
          // 0 f0', 1 igd, 2 y | 3 y2, 4 y^2+1/2, 5 y2^2+1/2, 6 y y2-1/2, 7 Re sp, 8 Im sp
          sb[0 * TL + l] = f0s; sb[1 * TL + l] = m.y; sb[2 * TL + l] = y; sb[3 * TL + l] = y2;
          sb[4 * TL + l] = y * y + 0.5; sb[5 * TL + l] = y2 * y2 + 0.5; sb[6 * TL + l] = y * y2 - 0.5;
          sb[7 * TL + l] = sp.re; sb[8 * TL + l] = sp.im;
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            const double2* j0 = reinterpret_cast<const double2*>(jt + q * (2 * TL * 4) + (0 * TL + l) * 4);
            const double2 u0 = j0[0], u1 = j0[1];                      // ds_re, ds_im | dz_re, dz_im
            const double dz_fac = jt[q * (2 * TL * 4) + (1 * TL + l) * 4];
            const cplx a = cmul(sp, {u1.x, u1.y + dz_fac * y});
            double* o = sq + q * JAC_Q_FIELDS * TL;
            o[0 * TL + l] = u0.x * cst::inv_sqrt_pi; o[1 * TL + l] = u0.y * cst::inv_sqrt_pi;
            o[2 * TL + l] = a.re; o[3 * TL + l] = a.im;
            o[4 * TL + l] = sp.re * dz_fac; o[5 * TL + l] = sp.im * dz_fac;
          }
        }
      } else {
        for (int l = tid; l < count; l += JAC_NT) {
          const double2 a = *reinterpret_cast<const double2*>(g + (0 * TL + l) * REC_GROUP);
          const double2 m = *reinterpret_cast<const double2*>(g + (1 * TL + l) * REC_GROUP);
          const double2 n = *reinterpret_cast<const double2*>(g + (1 * TL + l) * REC_GROUP + 2);
          const double2 h = *reinterpret_cast<const double2*>(g + (2 * TL + l) * REC_GROUP);
          const double2 k = *reinterpret_cast<const double2*>(g + (2 * TL + l) * REC_GROUP + 2);
          sb[0 * TL + l] = a.x; sb[1 * TL + l] = m.y; sb[2 * TL + l] = n.x; sb[3 * TL + l] = n.y;
          sb[4 * TL + l] = jp.real_lines ? 0.0 : h.y;  // s_im
          sb[5 * TL + l] = h.x; sb[6 * TL + l] = jcom[t * TL + l]; sb[7 * TL + l] = k.x;
          sb[8 * TL + l] = jp.real_lines ? h.y : k.y;  // cut_im | real lines: the line's cutoff [Hz]
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            const double2* j0 = reinterpret_cast<const double2*>(jt + q * (2 * TL * 4) + (0 * TL + l) * 4);
            const double2* j1 = reinterpret_cast<const double2*>(jt + q * (2 * TL * 4) + (1 * TL + l) * 4);
            const double2 u0 = j0[0], u1 = j0[1], u2 = j1[0], u3 = j1[1];
            double* o = sq + q * JAC_Q_FIELDS * TL;
            o[0 * TL + l] = u0.x; o[1 * TL + l] = u0.y; o[2 * TL + l] = u1.x; o[3 * TL + l] = u1.y;
            o[4 * TL + l] = u2.x; o[5 * TL + l] = u2.y; o[6 * TL + l] = u3.x;
          }
        }
      }
      const bool tile_im = __syncthreads_or(any_im) != 0;
      if (far_closed) {
        // closed-form far path for real lines (mode-0 segments: s, ds real; only Re dX is used since npm = (1,0,...))
        auto far_loop = [&](auto im_tag) {
          constexpr bool IM = decltype(im_tag)::value;
#pragma unroll 2
          for (int l = 0; l < count; l++) {
            const double f0s = sb[0 * TL + l], igd = sb[1 * TL + l];
            if (__double2hiint(igd) == 0) continue;  // igd == 0: inactive cutoff line (integer test, off the FP64 pipe)
            if (vf_pairs && line_all_vfar(f0s, igd, fblk_min, fblk_max)) continue;
            const FarLine c{sb[2 * TL + l], sb[3 * TL + l], sb[8 * TL + l], sb[9 * TL + l], sb[4 * TL + l], sb[5 * TL + l],
                            sb[6 * TL + l]};
            const double sp = sb[7 * TL + l];
            double Aq[NQ], Bq[NQ], Cq[NQ], Dq[NQ], Eq[NQ];
#pragma unroll
            for (int q = 0; q < NQ; q++) {
              const double* o = sq + q * JAC_Q_FIELDS * TL;
              Aq[q] = o[0 * TL + l]; Bq[q] = o[2 * TL + l]; Cq[q] = o[4 * TL + l]; Dq[q] = o[3 * TL + l];
              Eq[q] = IM ? o[1 * TL + l] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < JAC_R; r++) {
              const double x  = __dmul_rn(igd, __dsub_rn(f[r], f0s));
              if (vf_pairs && fabs(x) > VFAR_LIMIT) continue;
              const double x2 = __dadd_rn(x, fmax(__dmul_rn(1e-4, fabs(x)), 1e-4));
              double G1, G2, G3, G4, G5;
              far_pair<IM>(c, x, x2, G1, G2, G3, G4, G5);
              shape[r].re = __fma_rn(sp, G1, shape[r].re);
#pragma unroll
              for (int q = 0; q < NQ; q++) {
                double a = __fma_rn(Dq[q], G3, acc[q][r].re);
                if (IM) a = __fma_rn(Eq[q], G5, a);
                acc[q][r].re = __fma_rn(Aq[q], G1, __fma_rn(Bq[q], G2, __fma_rn(Cq[q], G4, a)));
              }
            }
          }
        };
        if (tile_im) far_loop(std::true_type{});
        else far_loop(std::false_type{});
        continue;
      }
      if (far_cplx) {
#pragma unroll 1
        for (int l = 0; l < count; l++) {
          const double f0s = sb[0 * TL + l], igd = sb[1 * TL + l];
          if (__double2hiint(igd) == 0) continue;
          if (vf_pairs && line_all_vfar(f0s, igd, fblk_min, fblk_max)) continue;
          const double y = sb[2 * TL + l], y2 = sb[3 * TL + l];
          const FarLine c{y, y2, 2.0 * y, 2.0 * y2, sb[4 * TL + l], sb[5 * TL + l], sb[6 * TL + l]};
          const double spr = sb[7 * TL + l], spi = sb[8 * TL + l];
#pragma unroll
          for (int r = 0; r < JAC_R; r++) {
            const double x  = __dmul_rn(igd, __dsub_rn(f[r], f0s));
            if (vf_pairs && fabs(x) > VFAR_LIMIT) continue;
            const double x2 = __dadd_rn(x, fmax(__dmul_rn(1e-4, fabs(x)), 1e-4));
            double G1, G2, G3, G4, G5;
            far_pair<true>(c, x, x2, G1, G2, G3, G4, G5);
            const double G6 = __dmul_rn(x, G3);
            shape[r].re = __fma_rn(spr, G1, __fma_rn(-spi, G5, shape[r].re));
            shape[r].im = __fma_rn(spr, G5, __fma_rn(spi, G1, shape[r].im));
#pragma unroll
            for (int q = 0; q < NQ; q++) {
              const double* o = sq + q * JAC_Q_FIELDS * TL;
              const double dr = o[0 * TL + l], di = o[1 * TL + l], ar = o[2 * TL + l], ai = o[3 * TL + l],
                           br = o[4 * TL + l], bi = o[5 * TL + l];
              // dF = (G2 - i G3) / sqrt(pi), F = (G1 + i G5) / sqrt(pi)
              acc[q][r].re = __fma_rn(dr, G1, __fma_rn(-di, G5, __fma_rn(ar, G2, __fma_rn(ai, G3,
                             __fma_rn(br, G4, __fma_rn(bi, G6, acc[q][r].re))))));
              acc[q][r].im = __fma_rn(dr, G5, __fma_rn(di, G1, __fma_rn(-ar, G3, __fma_rn(ai, G2,
                             __fma_rn(-br, G6, __fma_rn(bi, G4, acc[q][r].im))))));
            }
          }
        }
        continue;
      }
      for (int l = 0; l < count; l++) {
        const double f0s = sb[0 * TL + l], igd = sb[1 * TL + l], y = sb[2 * TL + l];
        if (igd == 0.0) continue;  // inactive cutoff line
        if (vf_pairs && line_all_vfar(f0s, igd, fblk_min, fblk_max)) continue;  // every pair of the line belongs to the other kernel
        const cplx s{sb[3 * TL + l], sb[4 * TL + l]};
        const double lcut = jp.real_lines ? sb[8 * TL + l] : cutoff;
        const bool lhas   = lcut < DBL_MAX;
        const cplx cutval{sb[7 * TL + l], jp.real_lines ? 0.0 : sb[8 * TL + l]};
        // real lines: most pairs of a near tile are still far; they take the closed form, per pair
        FarLine c;
        double sp = 0.0, Aq[NQ], Bq[NQ], Cq[NQ], Dq[NQ], Eq[NQ];
        if (jp.real_lines) {
          c.from_y(y);
          sp = s.re * cst::inv_sqrt_pi;
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            const double* o = sq + q * JAC_Q_FIELDS * TL;
            Aq[q] = o[0 * TL + l] * cst::inv_sqrt_pi; Bq[q] = sp * o[2 * TL + l]; Cq[q] = sp * o[4 * TL + l];
            Dq[q] = sp * (o[3 * TL + l] + o[4 * TL + l] * y);
            Eq[q] = -o[1 * TL + l] * cst::inv_sqrt_pi;
          }
        }
#pragma unroll
        for (int r = 0; r < JAC_R; r++) {
          if (lhas && !(f0s >= f[r] - lcut && f0s <= f[r] + lcut)) continue;
          if (vf_pairs && fabs(__dmul_rn(igd, __dsub_rn(f[r], f0s))) > VFAR_LIMIT) continue;
          if (jp.real_lines && jp.pair_far) {
            const double x  = __dmul_rn(igd, __dsub_rn(f[r], f0s));
            const double x2 = __dadd_rn(x, fmax(__dmul_rn(1e-4, fabs(x)), 1e-4));
            if (fmin(fabs(x), fabs(x2)) + y > FAR_LIMIT) {  // the pair and its displaced point are both far
              double G1, G2, G3, G4, G5;
              far_pair<true>(c, x, x2, G1, G2, G3, G4, G5);
              shape[r].re = __fma_rn(sp, G1, shape[r].re);
              if (lhas) shape[r].re -= cutval.re;
#pragma unroll
              for (int q = 0; q < NQ; q++) {
                acc[q][r].re = __fma_rn(Aq[q], G1, __fma_rn(Bq[q], G2, __fma_rn(Cq[q], G4, __fma_rn(Dq[q], G3,
                               __fma_rn(Eq[q], G5, acc[q][r].re)))));
                if (lhas) acc[q][r].re -= sq[(q * JAC_Q_FIELDS + 5) * TL + l];
              }
              continue;
            }
          }
          cplx z, F, dF;
          if (far) z_F_dF_far(igd * (f[r] - f0s), y, z, F, dF);
          else z_F_dF(igd * (f[r] - f0s), y, sb[5 * TL + l], sb[6 * TL + l], z, F, dF);
          cplx sh = cmul(s, F);
          if (lhas) sh = csub(sh, cutval);
          shape[r] = cadd(shape[r], sh);
          const cplx sdF = cmul(s, dF);
#pragma unroll
          for (int q = 0; q < NQ; q++) {
            const double* o = sq + q * JAC_Q_FIELDS * TL;
            const cplx ds{o[0 * TL + l], o[1 * TL + l]}, dzq{o[2 * TL + l], o[3 * TL + l]};
            const cplx tq = cadd(dzq, cscale(o[4 * TL + l], z));
            cplx d = cadd(cmul(ds, F), cmul(tq, sdF));  // dX, :310-323
            if (lhas) d = csub(d, {o[5 * TL + l], o[6 * TL + l]});
            acc[q][r] = cadd(acc[q][r], d);
          }
        }
      }
    }
    // compute_derivative :1474-1481 (T) and :1553-1560 (VMR): dpm += npm (.) (dscl shape + scl dshape)
    const double* __restrict__ npm = p.npm + (int64_t(lev) * 4 + seg.pol) * 7;
    const bool all_zero = npm[0] == 0 && npm[1] == 0 && npm[2] == 0 && npm[3] == 0 && npm[4] == 0 && npm[5] == 0 &&
                          npm[6] == 0;
    if (all_zero) continue;
#pragma unroll
    for (int r = 0; r < JAC_R; r++) {
      const int64_t i = fblk + r * JAC_NT + tid;
      if (i >= p.nf) continue;
      const double scl = line_scale_v(f[r], T, P);
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        cplx d = cscale(scl, acc[q][r]);
        const int kind = jp.kind[jp.q0 + q];
        if (kind == AB200_TARGET_T) d = cadd(d, cscale(line_scale_dT(f[r], T, P), shape[r]));
        if (!EXT && (kind == AB200_TARGET_MAG_U || kind == AB200_TARGET_WIND_U)) continue;  // unreachable: EXT is set for these passes
        const double scl_df = (EXT && kind == AB200_TARGET_WIND_U) ? line_scale_df(f[r], T, P) : 0.0;
        jac_store_rows(jp, jp.q0 + q, kind, lev, seg.pol, i, p.k_pitch, npm, f[r], scl, scl_df, shape[r], d);
      }
    }
  }
}

int launch_prepare_jac(const PrepareParams& p, const JacPrepParams& jp, int nlev, cudaStream_t stream) {
  if (p.ntiles == 0 || nlev == 0 || jp.nq == 0) return 0;
  dim3 grid(static_cast<unsigned>(p.ntiles), static_cast<unsigned>(nlev));
  lbl_prepare_jac_kernel<<<grid, TL, 0, stream>>>(p, jp);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

template <int NQ, bool EXT>
static int launch_sum_jac_ne(const SumParams& p, const JacSumParams& jp, dim3 grid, cudaStream_t stream) {
  const size_t smem = size_t(JAC_BASE_FIELDS + NQ * JAC_Q_FIELDS) * TL * sizeof(double);
  AB_CUDA(cudaFuncSetAttribute(lbl_sum_jac_kernel<NQ, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));  // per device
  lbl_sum_jac_kernel<NQ, EXT><<<grid, JAC_NT, smem, stream>>>(p, jp);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}
template <int NQ, int R>
static int launch_sum_jac_vfar_r(const SumParams& p, const JacSumParams& jp, int nlev, cudaStream_t stream) {
  const size_t smem = size_t(vf_stride(NQ)) * TL * sizeof(double);
  AB_CUDA(cudaFuncSetAttribute(lbl_sum_jac_vfar_kernel<NQ, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const dim3 grid(static_cast<unsigned>((p.nf + VF_NT * R - 1) / (VF_NT * R)), static_cast<unsigned>(nlev));
  lbl_sum_jac_vfar_kernel<NQ, R><<<grid, VF_NT, smem, stream>>>(p, jp);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}
// frequencies per thread: 4 (512-frequency blocks) unless the grid is so small that whole-wave quantisation costs more
// than the shorter blocks' extra staging (one path of configs[4]: 20 x 100 blocks on 148 SMs)
template <int NQ>
static int launch_sum_jac_vfar(const SumParams& p, const JacSumParams& jp, int nlev, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  auto eff = [&](int r, int per_sm) {
    const double items = double((p.nf + VF_NT * r - 1) / (VF_NT * r)) * nlev, slots = double(sms) * per_sm;
    return items / (std::ceil(items / slots) * slots);
  };
  int R = eff(2, VF_MB2) > eff(4, VF_MB4) + 0.05 ? 2 : 4;
  if (const char* e = getenv("AB200_JAC_VFAR_R")) R = atoi(e) == 2 ? 2 : 4;
  return R == 2 ? launch_sum_jac_vfar_r<NQ, 2>(p, jp, nlev, stream) : launch_sum_jac_vfar_r<NQ, 4>(p, jp, nlev, stream);
}
template <int NQ>
static int launch_sum_jac_vfar_cplx(const SumParams& p, const JacSumParams& jp, int nlev, cudaStream_t stream) {
  constexpr int R = 2;
  const size_t smem = size_t(vc_stride(NQ)) * VC_CH * sizeof(double);
  AB_CUDA(cudaFuncSetAttribute(lbl_sum_jac_vfar_cplx_kernel<NQ, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const dim3 grid(static_cast<unsigned>((p.nf + VF_NT * R - 1) / (VF_NT * R)), static_cast<unsigned>(nlev));
  lbl_sum_jac_vfar_cplx_kernel<NQ, R><<<grid, VF_NT, smem, stream>>>(p, jp);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}
template <int NQ>
static int launch_sum_jac_n(const SumParams& p, const JacSumParams& jp, dim3 grid, cudaStream_t stream) {
  bool ext = false;
  for (int q = 0; q < NQ; q++) ext |= jp.kind[jp.q0 + q] == AB200_TARGET_WIND_U || jp.kind[jp.q0 + q] == AB200_TARGET_MAG_U;
  return ext ? launch_sum_jac_ne<NQ, true>(p, jp, grid, stream) : launch_sum_jac_ne<NQ, false>(p, jp, grid, stream);
}

int launch_sum_jac(const SumParams& p, JacSumParams jp, int nlev, cudaStream_t stream) {
  if (p.nsegs == 0 || nlev == 0 || p.nf == 0 || jp.nq == 0) return 0;
  dim3 grid(static_cast<unsigned>((p.nf + JAC_F_TILE - 1) / JAC_F_TILE), static_cast<unsigned>(nlev));
  {
    const char* e = getenv("AB200_JAC_PAIR_FAR");
    jp.pair_far = e ? atoi(e) : 1;
  }
  // targets per pass: JAC_Q, or AB200_JAC_PASS (1..4) for experiments
  int per_pass = JAC_Q;
  if (const char* e = getenv(jp.real_lines ? "AB200_JAC_PASS_REAL" : "AB200_JAC_PASS_CPLX")) per_pass = std::max(1, std::min(JAC_Q, atoi(e)));
  const bool vfar_on = [] { const char* e = getenv("AB200_JAC_VFAR"); return e ? atoi(e) != 0 : true; }();
  for (int q0 = 0; q0 < jp.nq; q0 += per_pass) {
    jp.q0 = q0;
    const int nq_pass = std::min(per_pass, jp.nq - q0);
    // real lines: the very far pairs go to the one-rational-function kernel, unless the pass holds line
    // targets only (those visit a handful of tiles and need no forward shape)
    bool line_only = true;
    for (int q = 0; q < nq_pass; q++) line_only &= jp.kind[q0 + q] >= AB200_TARGET_LINE_F0 && jp.kind[q0 + q] <= AB200_TARGET_LINE_LS;
    jp.skip_vfar = (vfar_on && !line_only) ? 1 : 0;
    if (jp.skip_vfar && !jp.real_lines) {
      switch (nq_pass) {
        case 1: AB_TRY((launch_sum_jac_vfar_cplx<1>(p, jp, nlev, stream))); break;
        case 2: AB_TRY((launch_sum_jac_vfar_cplx<2>(p, jp, nlev, stream))); break;
        case 3: AB_TRY((launch_sum_jac_vfar_cplx<3>(p, jp, nlev, stream))); break;
        default: AB_TRY((launch_sum_jac_vfar_cplx<4>(p, jp, nlev, stream))); break;
      }
    }
    if (jp.skip_vfar && jp.real_lines) {
      switch (nq_pass) {
        case 1: AB_TRY(launch_sum_jac_vfar<1>(p, jp, nlev, stream)); break;
        case 2: AB_TRY(launch_sum_jac_vfar<2>(p, jp, nlev, stream)); break;
        case 3: AB_TRY(launch_sum_jac_vfar<3>(p, jp, nlev, stream)); break;
        default: AB_TRY(launch_sum_jac_vfar<4>(p, jp, nlev, stream)); break;
      }
    }
    switch (nq_pass) {
      case 1: AB_TRY(launch_sum_jac_n<1>(p, jp, grid, stream)); break;
      case 2: AB_TRY(launch_sum_jac_n<2>(p, jp, grid, stream)); break;
      case 3: AB_TRY(launch_sum_jac_n<3>(p, jp, grid, stream)); break;
      default: AB_TRY(launch_sum_jac_n<4>(p, jp, grid, stream)); break;
    }
  }
  return 0;
}

}  // namespace ab200
