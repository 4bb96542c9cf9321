// stokes.hpp — kernel parameter block and launchers of stage 2 (stokes.cu).
#pragma once

#include "common.cuh"

namespace ab200 {

struct StokesParams {
  int32_t np;
  int64_t nf;
  const double* K;     // [np][k_pitch][7]; k_pitch % 128 == 0 so every 128-frequency row block is a full TMA box
  int64_t k_pitch;
  const double* f;     // [np][nf] or [nf]
  int64_t f_stride;
  const double* ffac;  // [np] wind-shift factor of each level's grid (1 without wind)
  const double* T;     // [np] level temperatures
  const double* invT;  // [np] 1 / T
  const double* r;     // [np-1] layer lengths
  const double* I_bkg; // [nf][4]
  double* I;           // [nf][4]
  double* I_lev;       // optional [np][nf][4]: radiance arriving at every level (for the Jacobian pass), or nullptr
  int32_t rte_option;
  int32_t tran_exact;
  int* flags;          // device error flags
  int32_t scalar;      // K is known to have only A != 0 (no polarised segment was summed): scalar fast path
  int32_t no_emission; // J = 0 at every level: pure transmission (AB200_FLAG_NO_EMISSION)
};

// fused Jacobian pass B (stokes_jac.cu)
struct StokesJacParams {
  int32_t np, nq;
  int64_t nf;
  const double* K;      // [np][k_pitch][7]
  const double* dK;     // [np][nq][k_pitch][7]
  int64_t k_pitch;
  const double* f;      // [np][nf] or [nf]
  int64_t f_stride;
  const double* ffac;   // [np]
  const double* T;      // [np]
  const double* r;      // [np-1]
  const double* dr;     // [2][np-1][nq]
  const double* I_lev;  // [np][nf][4] radiance arriving at each level (pass A)
  double* dI;           // [nf][np][nq][4], or nullptr when only the x-space Jacobian is wanted
  // observer epilogue (x-space accumulation inside the pass, spectral_rad_jacAddPathPropagation m_rad.cc:62-127 and
  // spectral_rad_jacFromBackground :26-60): Jx [nx][nf][4] zero on entry, or nullptr
  double* Jx;
  const int64_t* map_offset;  // [np*nq + 1] CSR over (level, target)
  const int32_t* map_x;
  const double* map_w;
  int32_t n_bkg;              // surface-temperature rows: Jx[bkg_x][f] += P[f][np-1] (bkg_w dB/dT(f, bkg_T), 0, 0, 0)
  const int32_t* bkg_x;
  const double* bkg_w;
  double bkg_T;
  int32_t it;           // index of the temperature target or -1
  int32_t rte_option;
  int* flags;
  int32_t no_emission;  // J = 0, dJ = 0 at every level
  int32_t scalar;       // K and dK are known to have only A != 0: scalar pass (no 4x4 algebra)
};

int launch_stokes_chain(const StokesParams& p, cudaStream_t stream);
int launch_stokes_jac(const StokesJacParams& p, cudaStream_t stream);
int launch_tramat_jac(int np, int64_t nf, int nq, const double* K, const double* dK, const double* r, const double* dr,
                      int linsrc, double* dT, double* dL, int linprop, cudaStream_t stream);
int launch_rte_emission_jac(int linsrc, int np, int64_t nf, int nq, const double* T, const double* L, const double* P,
                            const double* dT, const double* dL, const double* J, const double* dJ, const double* I_bkg,
                            double* I, double* dI, cudaStream_t stream);
int launch_planck_tb(int64_t nf, const double* f, double* I, cudaStream_t stream);
int launch_transmission_apply(int np, int64_t nf, const double* P, const double* I_bkg, double* I, cudaStream_t stream);

// observer epilogue (observer.cu)
int launch_background_planck(int64_t nf, const double* f, double T, double* I_bkg, cudaStream_t stream);
int launch_unit_transform(int64_t nf, int32_t nx, const double* f, int32_t unit, double n_real, double* I, double* Jx,
                          cudaStream_t stream);
int launch_sensor_sumup(int64_t nf, int32_t nx, int32_t n_channels, const int64_t* w_offset, const int64_t* w_freq,
                        const double* w_stokes, const double* I, const double* Jx, double* y, double* Jy,
                        cudaStream_t stream);
int launch_tramat(int np, int64_t nf, const double* K, const double* r, int linsrc, int exact, double* T, double* L,
                  double* P, int linprop, int* flags, cudaStream_t stream);
int launch_srcvec(int np, int64_t nf, int nq, const double* K, const double* f, int64_t f_stride, const double* Tlev,
                  int it, double* J, double* dJ, cudaStream_t stream);
int launch_rte_emission(int linsrc, int np, int64_t nf, const double* T, const double* L, const double* J,
                        const double* I_bkg, double* I, cudaStream_t stream);

int launch_dawson(int64_t n, const double* zr, const double* zi, double* dr, double* di, cudaStream_t stream);
}  // namespace ab200
