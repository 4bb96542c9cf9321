// lbl_model.cuh — device versions of the line-shape model (values, d/dT, d/dVMR).
//
// reference: src/core/lbl/lbl_temperature_model.h:62-314 (models T0..T5, AER, DPL, POLY and their
// d/dT members), src/core/lbl/lbl_lineshape_model.cpp:14-35 (pressure scaling), :70-90 (VMR-weighted
// mixture, Bath = remainder), :92-113 (d/dVMR), :127-148 (d/dT).
#pragma once

#include "common.cuh"

namespace ab200 {

__device__ __forceinline__ double tm_value(int type, const double* __restrict__ x, double T0, double T) {
  switch (type) {
    case AB200_TM_T0: return x[0];
    case AB200_TM_T1: return x[0] * pow(T0 / T, x[1]);
    case AB200_TM_T2: return x[0] * pow(T0 / T, x[1]) * (1 + x[2] * log(T / T0));
    case AB200_TM_T3: return x[0] + x[1] * (T - T0);
    case AB200_TM_T4: return (x[0] + x[1] * (T0 / T - 1)) * pow(T0 / T, x[2]);
    case AB200_TM_T5: return x[0] * pow(T0 / T, 0.25 + 1.5 * x[1]);
    case AB200_TM_AER:
      if (T < 250.0) return x[0] + (T - 200.0) * (x[1] - x[0]) / (250.0 - 200.0);
      if (T > 296.0) return x[2] + (T - 296.0) * (x[3] - x[2]) / (340.0 - 296.0);
      return x[1] + (T - 250.0) * (x[2] - x[1]) / (296.0 - 250.0);
    case AB200_TM_DPL: return x[0] * pow(T0 / T, x[1]) + x[2] * pow(T0 / T, x[3]);
    case AB200_TM_POLY: return x[0] + T * (x[1] + T * (x[2] + T * x[3]));
    default: return 0.0;
  }
}

// The same values with the power written as the reference writes it, nonstd::pow(x, v) = exp(v log x)
// (lbl_temperature_model.h:36-60 through src/core/util/nonstd.h:25-31), and log(T0 / T) handed in: all the variables and
// broadeners of a line share T0 / T, so a line pays one log and one exp per power instead of a full pow() each.
__device__ __forceinline__ double tm_value_lq(int type, const double* __restrict__ x, double T0, double T, double q /* T0 / T */,
                                              double lq /* log(T0 / T) */) {
  switch (type) {
    case AB200_TM_T0: return x[0];
    case AB200_TM_T1: return x[0] * exp(x[1] * lq);
    case AB200_TM_T2: return x[0] * exp(x[1] * lq) * (1 + x[2] * log(T / T0));
    case AB200_TM_T3: return x[0] + x[1] * (T - T0);
    case AB200_TM_T4: return (x[0] + x[1] * (q - 1)) * exp(x[2] * lq);
    case AB200_TM_T5: return x[0] * exp((0.25 + 1.5 * x[1]) * lq);
    case AB200_TM_AER:
      if (T < 250.0) return x[0] + (T - 200.0) * (x[1] - x[0]) / (250.0 - 200.0);
      if (T > 296.0) return x[2] + (T - 296.0) * (x[3] - x[2]) / (340.0 - 296.0);
      return x[1] + (T - 250.0) * (x[2] - x[1]) / (296.0 - 250.0);
    case AB200_TM_DPL: return x[0] * exp(x[1] * lq) + x[2] * exp(x[3] * lq);
    case AB200_TM_POLY: return x[0] + T * (x[1] + T * (x[2] + T * x[3]));
    default: return 0.0;
  }
}

// d/dT of the above (lbl_temperature_model.h, the d*_dT members)
__device__ __forceinline__ double tm_dT(int type, const double* __restrict__ x, double T0, double T) {
  switch (type) {
    case AB200_TM_T0: return 0.0;
    case AB200_TM_T1: return -x[0] * x[1] * pow(T0 / T, x[1]) / T;
    case AB200_TM_T2:
      return -x[0] * x[1] * pow(T0 / T, x[1]) * (x[2] * log(T / T0) + 1.) / T + x[0] * x[2] * pow(T0 / T, x[1]) / T;
    case AB200_TM_T3: return x[1];
    case AB200_TM_T4:
      return -x[2] * pow(T0 / T, x[2]) * (x[0] + x[1] * (T0 / T - 1.)) / T - T0 * x[1] * pow(T0 / T, x[2]) / (T * T);
    case AB200_TM_T5: return -x[0] * pow(T0 / T, 1.5 * x[1] + 0.25) * (1.5 * x[1] + 0.25) / T;
    case AB200_TM_AER:
      if (T < 250.0) return (x[1] - x[0]) / (250.0 - 200.0);
      if (T > 296.0) return (x[3] - x[2]) / (340.0 - 296.0);
      return (x[2] - x[1]) / (296.0 - 250.0);
    case AB200_TM_DPL: return -x[0] * x[1] * pow(T0 / T, x[1]) / T + -x[2] * x[3] * pow(T0 / T, x[3]) / T;
    case AB200_TM_POLY: return x[1] + T * (2.0 * x[2] + T * 3.0 * x[3]);
    default: return 0.0;
  }
}

// d/dX_k of the above, k = 0..3 (lbl_temperature_model.h, the d*_dX0..dX3 members; 0 where a model has no such coefficient)
__device__ __forceinline__ double tm_dX(int type, int k, const double* __restrict__ x, double T0, double T) {
  const double q = T0 / T;
  switch (type) {
    case AB200_TM_T0: return k == 0 ? 1.0 : 0.0;
    case AB200_TM_T1: return k == 0 ? pow(q, x[1]) : k == 1 ? x[0] * pow(q, x[1]) * log(q) : 0.0;
    case AB200_TM_T2:
      return k == 0   ? pow(q, x[1]) * (1 + x[2] * log(T / T0))
             : k == 1 ? x[0] * pow(q, x[1]) * (x[2] * log(T / T0) + 1.) * log(q)
             : k == 2 ? x[0] * pow(q, x[1]) * log(T / T0)
                      : 0.0;
    case AB200_TM_T3: return k == 0 ? 1.0 : k == 1 ? T - T0 : 0.0;
    case AB200_TM_T4:
      return k == 0   ? pow(q, x[2])
             : k == 1 ? pow(q, x[2]) * (q - 1.)
             : k == 2 ? pow(q, x[2]) * (x[0] + x[1] * (q - 1)) * log(q)
                      : 0.0;
    case AB200_TM_T5:
      return k == 0 ? pow(q, 1.5 * x[1] + 0.25) : k == 1 ? 1.5 * x[0] * pow(q, 1.5 * x[1] + 0.25) * log(q) : 0.0;
    case AB200_TM_AER:
      if (k == 0) return T < 250.0 ? 1 - (T - 200.0) / (250.0 - 200.0) : 0.0;
      if (k == 1) return T < 250.0 ? (T - 200.0) / (250.0 - 200.0) : T > 296.0 ? 0.0 : 1 - (T - 250.0) / (296.0 - 250.0);
      if (k == 2) return T < 250.0 ? 0.0 : T > 296.0 ? 1 - (T - 296.0) / (340.0 - 296.0) : (T - 250.0) / (296.0 - 250.0);
      return T > 296.0 ? (T - 296.0) / (340.0 - 296.0) : 0.0;
    case AB200_TM_DPL:
      return k == 0   ? pow(q, x[1])
             : k == 1 ? x[0] * pow(q, x[1]) * log(q)
             : k == 2 ? pow(q, x[3])
                      : x[2] * pow(q, x[3]) * log(q);
    case AB200_TM_POLY: return k == 0 ? 1.0 : k == 1 ? T : k == 2 ? T * T : T * T * T;
    default: return 0.0;
  }
}

// view of one catalog line's broadening table on the device
struct LineModel {
  const int64_t* ls_offset;
  const int32_t* ls_species;
  const int32_t* ls_type;
  const double* ls_X;
  int64_t line;
  double T0, T, P;
  const double* vmr;  // this level's [n_species]

  __device__ __forceinline__ static double pscale(int var, double P) {
    return (var == AB200_VAR_G || var == AB200_VAR_DV) ? P * P : P;  // lbl_lineshape_model.cpp:27-35
  }
  __device__ __forceinline__ double single(int64_t i, int var, bool dT) const {
    const int type = ls_type[i * AB200_NVAR + var];
    if (type == AB200_TM_ABSENT) return 0.0;
    const double* x = ls_X + (i * AB200_NVAR + var) * 4;
    return pscale(var, P) * (dT ? tm_dT(type, x, T0, T) : tm_value(type, x, T0, T));
  }
  // model::VAR(atm) / model::dVAR_dT(atm), :70-90, :127-148
  __device__ __forceinline__ double mix(int var, bool dT) const {
    double vsum = 0.0, res = 0.0, bth = 0.0;
    bool has_bath = false;
    for (int64_t i = ls_offset[line]; i < ls_offset[line + 1]; i++) {
      const double r = single(i, var, dT);
      const int sp   = ls_species[i];
      if (sp != AB200_SPECIES_BATH) {
        const double v = vmr[sp];
        vsum += v;
        res += v * r;
      } else {
        bth      = r;
        has_bath = true;
      }
    }
    return has_bath ? res + (1.0 - vsum) * bth : res / vsum;
  }
  // model::dVAR_dVMR(atm, species), :92-113
  __device__ __forceinline__ double dmix_dvmr(int var, int species) const {
    int64_t ptr = -1, bth = -1;
    for (int64_t i = ls_offset[line]; i < ls_offset[line + 1]; i++) {
      if (ls_species[i] == species) ptr = i;
      if (ls_species[i] == AB200_SPECIES_BATH) bth = i;
    }
    if (ptr < 0) return 0.0;
    const double x = single(ptr, var, false);
    if (species == AB200_SPECIES_BATH) return -x;
    if (bth >= 0) return x - single(bth, var, false);
    double t = 0.0;
    for (int64_t i = ls_offset[line]; i < ls_offset[line + 1]; i++) t += vmr[ls_species[i]];
    return (t - x) / t * t;  // sic, lbl_lineshape_model.cpp:112
  }
  // model::dVAR_dX(atm, species, coeff), lbl_lineshape_model.cpp:150-246 (species_model::dVAR_dXk :38-63)
  __device__ __forceinline__ double dmix_dX(int var, int species, int coeff) const {
    int64_t ptr = -1, bth = -1;
    double vsum = 0.0, vall = 0.0;
    for (int64_t i = ls_offset[line]; i < ls_offset[line + 1]; i++) {
      const int sp = ls_species[i];
      if (sp == species) ptr = i;
      if (sp == AB200_SPECIES_BATH) bth = i;
      else vsum += vmr[sp];
    }
    if (ptr < 0) return 0.0;
    const int type = ls_type[ptr * AB200_NVAR + var];
    const double x = type == AB200_TM_ABSENT
                         ? 0.0
                         : pscale(var, P) * tm_dX(type, coeff, ls_X + (ptr * AB200_NVAR + var) * 4, T0, T);
    if (species == AB200_SPECIES_BATH) return (1 - vsum) * x;
    if (bth >= 0) return vmr[species] * x;
    vall = vsum;  // no bath entry: every broadener is a species
    return x * vmr[species] / vall;
  }
};

}  // namespace ab200
