// predef.cu — predefined continuum models added into the resident propagation matrix.
//
//   predef_kernel   spectral_propmatAddPredefined (src/m_predefined_absorption_models.cc:156-191) with
//                   Absorption::PredefinedModel::compute (src/core/absorption/predefined_absorption_models.cc:219-317) for
//                   every (frequency, level): the four "StandardType" continua of src/core/predefined/standard.cc
//                   (O2 :51-84, N2 :118-138, H2O foreign :166-184, H2O self :212-226), the temperature row and the
//                   CO2 / O2 / N2 / H2O / liquidcloud VMR rows by the reference's perturbation (model(x + d) - model(x)) / d.
//
// One thread per (frequency, level); HBM bound on K (16 B per element, + 16 B per affected Jacobian row).
#include <cmath>

#include "predef.hpp"

namespace ab200 {

struct PredefParams {
  int32_t n_models;
  int32_t models[8];
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
};

struct PredefPoint {
  double T, P, o2, n2, h2o;
};

__device__ __forceinline__ double predef_model(int m, double f, const PredefPoint& a) {
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
    default: {  // Standard::water_self
      constexpr double C = 1.796e-33, x = 4.5;
      const double dummy = C * pow(300. / a.T, x + 3) * (a.P * a.P) * a.h2o;
      return a.h2o * dummy * (f * f);
    }
  }
}

__host__ __device__ inline int predef_species_of(int m, const ab200_predef_species& s) {
  return m == AB200_PREDEF_O2_SELFCONT_STANDARD ? s.o2 : m == AB200_PREDEF_N2_SELFCONT_STANDARD ? s.n2 : s.h2o;
}

__global__ void __launch_bounds__(128) predef_kernel(PredefParams p) {
  const int64_t iv = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (iv >= p.nf) return;
  const int lev = blockIdx.y;
  const double f = (p.ffac ? p.ffac[lev] : 1.0) * p.f[int64_t(lev) * p.f_stride + iv];
  const double* __restrict__ vmr = p.vmr + int64_t(lev) * p.n_species;
  auto v = [&](int idx) { return idx >= 0 ? vmr[idx] : 0.0; };
  const PredefPoint a{p.T[lev], p.P[lev], v(p.sp.o2), v(p.sp.n2), v(p.sp.h2o)};
  double kacc = 0.0, dacc[AB200_MAX_TARGETS];
#pragma unroll
  for (int q = 0; q < AB200_MAX_TARGETS; q++) dacc[q] = 0.0;
  const int vmr_idx[5] = {p.sp.co2, p.sp.o2, p.sp.n2, p.sp.h2o, p.sp.liquidcloud};  // vmrs_jac order, :237-241
  for (int k = 0; k < p.n_models; k++) {
    const int m = p.models[k];
    if (p.select_species != AB200_SPECIES_BATH && predef_species_of(m, p.sp) != p.select_species) continue;
    const double pm = predef_model(m, f, a);
    kacc += pm;
    if (p.it >= 0) {
      PredefPoint b = a;
      b.T += p.tg_d[p.it];
      dacc[p.it] += (predef_model(m, f, b) - pm) / p.tg_d[p.it];
    }
    for (int j = 0; j < 5; j++) {
      const int idx = vmr_idx[j];
      if (idx < 0) continue;
      for (int q = 0; q < p.nq; q++)
        if (p.tg_kind[q] == AB200_TARGET_VMR && p.tg_species[q] == idx) {
          PredefPoint b = a;
          if (idx == p.sp.o2) b.o2 += p.tg_d[q];
          if (idx == p.sp.n2) b.n2 += p.tg_d[q];
          if (idx == p.sp.h2o) b.h2o += p.tg_d[q];
          dacc[q] += (predef_model(m, f, b) - pm) / p.tg_d[q];
          break;
        }
    }
  }
  p.K[(int64_t(lev) * p.k_pitch + iv) * 7] += kacc;
  for (int q = 0; q < p.nq; q++)
    if (dacc[q] != 0.0) p.dK[((int64_t(lev) * p.nq + q) * p.k_pitch + iv) * 7] += dacc[q];
}

// fills the model / species / target part of the parameters and validates it; 0 or an error code with the message set
int predef_setup(PredefParams& pp, const int32_t* models, int32_t n_models, const ab200_predef_species* sp, int32_t n_species, int32_t nq,
                 const int32_t* tg_kind, const int32_t* tg_species, const double* target_d) {
  if (n_models < 0 || n_models > 8 || (n_models > 0 && !models) || !sp)
    return set_error(AB200_ERR_INVALID, "predefined models: null argument or more than 8 models");
  pp.n_models = n_models;
  pp.sp = *sp;
  for (int idx : {sp->o2, sp->n2, sp->h2o, sp->co2, sp->liquidcloud})
    if (idx >= n_species) return set_error(AB200_ERR_INVALID, "predefined models: species index beyond the VMR vector");
  for (int k = 0; k < n_models; k++) {
    const int m = models[k];
    if (m < AB200_PREDEF_O2_SELFCONT_STANDARD || m > AB200_PREDEF_H2O_SELFCONT_STANDARD)
      return set_error(AB200_ERR_UNSUPPORTED, "predefined model " + std::to_string(m) +
                                                  " is outside the GPU path (only the four StandardType continua are; no CPU fallback)");
    const bool need_h2o = m != AB200_PREDEF_N2_SELFCONT_STANDARD;
    if ((m == AB200_PREDEF_O2_SELFCONT_STANDARD && sp->o2 < 0) || (m == AB200_PREDEF_N2_SELFCONT_STANDARD && sp->n2 < 0) ||
        (need_h2o && sp->h2o < 0))
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
                   const int32_t* tg_kind, const int32_t* tg_species, int np, cudaStream_t stream) {
  PredefParams pp{};
  AB_TRY(predef_setup(pp, models, n_models, sp, n_species, nq, tg_kind, tg_species, target_d));
  pp.nf = nf; pp.f = d_f; pp.f_stride = f_stride; pp.ffac = d_ffac; pp.T = d_T; pp.P = d_P; pp.vmr = d_vmr;
  pp.n_species = n_species; pp.select_species = select_species; pp.K = d_K; pp.dK = d_dK; pp.k_pitch = k_pitch;
  return launch_predef(pp, np, stream);
}

}  // namespace ab200

using namespace ab200;

extern "C" int ab200_predef_levels(const int32_t* models, int32_t n_models, const ab200_predef_species* species, int64_t nf, const double* f,
                                   int64_t f_level_stride, const ab200_atm_path* atm, int32_t n_species, int32_t select_species, int32_t nq,
                                   const ab200_target* targets, const double* target_d, double* K, double* dK) {
  if (!atm || !K || (nf > 0 && !f)) return set_error(AB200_ERR_INVALID, "ab200_predef_levels: null argument");
  if (nf < 0 || atm->np < 0 || nq < 0 || nq > AB200_MAX_TARGETS || n_species <= 0) return set_error(AB200_ERR_INVALID, "ab200_predef_levels: bad size");
  if (nq > 0 && (!targets || !dK)) return set_error(AB200_ERR_INVALID, "ab200_predef_levels: null Jacobian argument with nq > 0");
  if (f_level_stride != 0 && f_level_stride != nf) return set_error(AB200_ERR_INVALID, "f_level_stride must be 0 or nf");
  int32_t kind[AB200_MAX_TARGETS], spc[AB200_MAX_TARGETS];
  for (int q = 0; q < nq; q++) { kind[q] = targets[q].kind; spc[q] = targets[q].species; }
  PredefParams pp{};
  AB_TRY(predef_setup(pp, models, n_models, species, n_species, nq, kind, spc, target_d));
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
  } bf, bT, bP, bv, bK, bdK;
  const size_t nfl = static_cast<size_t>(nf) * (f_level_stride ? np : 1), nk = static_cast<size_t>(np) * nf * 7;
  if (bf.put(f, nfl * 8) || bT.put(atm->T, np * 8) || bP.put(atm->P, np * 8) || bv.put(atm->vmr, static_cast<size_t>(np) * n_species * 8) ||
      bK.put(K, nk * 8) || bdK.put(dK, nk * nq * 8))
    return set_error(AB200_ERR_NOMEM, "ab200_predef_levels: device allocation or copy failed");
  pp.nf = nf; pp.f = static_cast<double*>(bf.p); pp.f_stride = f_level_stride; pp.ffac = nullptr;
  pp.T = static_cast<double*>(bT.p); pp.P = static_cast<double*>(bP.p); pp.vmr = static_cast<double*>(bv.p);
  pp.n_species = n_species; pp.select_species = select_species;
  pp.K = static_cast<double*>(bK.p); pp.dK = static_cast<double*>(bdK.p); pp.k_pitch = nf;
  AB_TRY(launch_predef(pp, np, nullptr));
  AB_CUDA(cudaMemcpy(K, bK.p, nk * 8, cudaMemcpyDeviceToHost));
  if (nq > 0) AB_CUDA(cudaMemcpy(dK, bdK.p, nk * nq * 8, cudaMemcpyDeviceToHost));
  return AB200_OK;
}
