// stokes_jac.cu — Jacobian part of stage 2 (nq > 0): d/dq of the transmission matrices, of the
// linear-in-source operator and of the radiance recursion, for temperature / VMR targets.
//
//   tramat_jac_kernel        dT, dL of TransmittanceMatrix::{constant,linsrc}
//                            (reference src/core/rtepack/rtepack_transmission.cc:1114-1193 with
//                            tran::deriv :558-674 and tran::linsrc_deriv :277-447)
//   rte_emission_jac_kernel  spectral_rad_jac_path of rte_emission on materialised arrays
//                            (rtepack_rtestep.cc:265-372, the `if (nq)` blocks)
//   stokes_levels_kernel +   the FUSED version: pass A walks the path from the background and keeps the
//   stokes_jac_kernel        radiance arriving at every level (32 B per step instead of the reference's
//                            3 x 128 B + 4 x 128 B nq); pass B walks from the sensor with the running
//                            cumulative transmission P_i = T_1 ... T_i (rtepack_transmission.cc:1322-1327),
//                            re-forms T, Lambda and their derivatives in registers and accumulates dI.
//
// All arithmetic is the reference's literal form; AB200_FLAG_TRAN_EXACT is rejected
// together with Jacobian targets at the API.
#include "rtepack.cuh"
#include "stokes.hpp"

namespace ab200 {
using namespace rte;

__device__ __forceinline__ void mv(const double* __restrict__ m, const double* __restrict__ s, double* __restrict__ o) {
  mat_vec(m, s, o);
}
__device__ __forceinline__ void diag_of(double* __restrict__ m, double d) {
#pragma unroll
  for (int i = 0; i < 16; i++) m[i] = 0.0;
  m[0] = m[5] = m[10] = m[15] = d;
}
__device__ __forceinline__ double midpoint(double a, double b) { return (a + b) * 0.5; }  // std::midpoint for finite, non-huge doubles

// one layer's contribution to dI0 (level i) and dI1 (level i+1), BEFORE the multiplication with P_i
// constant: rtepack_rtestep.cc:293-306 ; linsrc: :348-367.  dj0 = dJ[i][q], dj1 = dJ[i+1][q] (Stokes vectors).
template <bool LINSRC>
__device__ __forceinline__ void layer_terms(const double* __restrict__ T, const double* __restrict__ L,
                                            const double* __restrict__ dT0, const double* __restrict__ dT1,
                                            const double* __restrict__ dL0, const double* __restrict__ dL1,
                                            const double* __restrict__ v /*Iv - J (const) | Iv - J0 (linsrc)*/,
                                            const double* __restrict__ jd /*J0 - J1 (linsrc)*/,
                                            const double* __restrict__ dj0, const double* __restrict__ dj1,
                                            double* __restrict__ c0, double* __restrict__ c1) {
  double a[4], b[4], x[4], y[4];
  if (LINSRC) {
    // dI0 += P (dJ1 - L dJ0 + dT0 ImJ0 + dL0 J0mJ1);  dI1 += P (dT1 ImJ0 + dL1 J0mJ1 + L dJ1 - T dJ0)
    mv(dT0, v, a);
    mv(dL0, jd, b);
    mv(L, dj0, x);
#pragma unroll
    for (int e = 0; e < 4; e++) c0[e] = dj1[e] - x[e] + a[e] + b[e];
    mv(dT1, v, a);
    mv(dL1, jd, b);
    mv(L, dj1, x);
    mv(T, dj0, y);
#pragma unroll
    for (int e = 0; e < 4; e++) c1[e] = a[e] + b[e] + x[e] - y[e];
  } else {
    // dI0 += P (dT0 Iv + avg(dJ0, -(T dJ0)));  dI1 += P (dT1 Iv + avg(dJ1, -(T dJ1)))
    mv(dT0, v, a);
    mv(T, dj0, x);
#pragma unroll
    for (int e = 0; e < 4; e++) c0[e] = a[e] + midpoint(dj0[e], -x[e]);
    mv(dT1, v, a);
    mv(T, dj1, x);
#pragma unroll
    for (int e = 0; e < 4; e++) c1[e] = a[e] + midpoint(dj1[e], -x[e]);
  }
}

// ---------------------------------------------------------------------------
// un-fused: dT, dL [2][nf][np][nq][16]; one thread per (frequency, layer i = 1..np-1)
// ---------------------------------------------------------------------------
__global__ void tramat_jac_kernel(int np, int64_t nf, int nq, const double* __restrict__ K, const double* __restrict__ dK,
                                  const double* __restrict__ r, const double* __restrict__ dr, int linsrc,
                                  double* __restrict__ dT, double* __restrict__ dL, int linprop) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= nf * (np - 1)) return;
  const int64_t iv = idx / (np - 1);
  const int i      = 1 + int(idx % (np - 1));
  const Propmat k1 = load_propmat(K + (int64_t(i - 1) * nf + iv) * 7);
  const Propmat k2 = load_propmat(K + (int64_t(i) * nf + iv) * 7);
  const double ri  = r[i - 1];
  Tran t;
  t.init(k1, k2, ri, false);
  double Tm[16], m[16];
  if (t.polarized) t.T(Tm); else diag_of(Tm, t.exp_a);
  const int64_t half = nf * int64_t(np) * nq * 16;
  for (int j = 0; j < nq; j++) {
    const Propmat dk0 = load_propmat(dK + ((int64_t(i - 1) * nq + j) * nf + iv) * 7);
    const Propmat dk1 = load_propmat(dK + ((int64_t(i) * nq + j) * nf + iv) * 7);
    const double dr0 = dr[int64_t(i - 1) * nq + j];
    const double dr1 = dr[(int64_t(np - 1) + (i - 1)) * nq + j];
    double* o0 = dT + ((iv * np + (i - 1)) * nq + j) * 16;
    double* o1 = dT + half + ((iv * np + i) * nq + j) * 16;
    t.deriv(Tm, k1, k2, dk0, ri, dr0, m);
    const double dt00_0 = m[0];
#pragma unroll
    for (int e = 0; e < 16; e++) o0[e] = m[e];
    t.deriv(Tm, k1, k2, dk1, ri, dr1, m);
    const double dt00_1 = m[0];
#pragma unroll
    for (int e = 0; e < 16; e++) o1[e] = m[e];
    if (linsrc) {
      double* l0 = dL + ((iv * np + (i - 1)) * nq + j) * 16;
      double* l1 = dL + half + ((iv * np + i) * nq + j) * 16;
      // linprop (TransmittanceMatrix::linprop :1225-1247) passes dr1 to BOTH derivative calls (:1238, sic)
      const int lc = linprop ? linprop_case(k1.A, k2.A, ri, t.polarized) : 0;
      double Lm[16];
      if (lc == 2) linprop_lambda_pol(Tm, k1, k2, ri, 4, Lm);  // the perturbed Lambda is compared with the layer's own, :545-555
      if (lc == 2) linprop_lambda_pol_deriv(Lm, k1, k2, dk0, ri, dr1, true, m);
      else if (lc == 1) diag_of(m, linprop_lambda_deriv(k1.A, k2.A, dk0.A, Tm[0], dt00_0, ri, dr1, true));
      else t.linsrc_deriv(dk0, ri, linprop ? dr1 : dr0, m);
#pragma unroll
      for (int e = 0; e < 16; e++) l0[e] = m[e];
      if (lc == 2) linprop_lambda_pol_deriv(Lm, k1, k2, dk1, ri, dr1, false, m);
      else if (lc == 1) diag_of(m, linprop_lambda_deriv(k1.A, k2.A, dk1.A, Tm[0], dt00_1, ri, dr1, false));
      else t.linsrc_deriv(dk1, ri, dr1, m);
#pragma unroll
      for (int e = 0; e < 16; e++) l1[e] = m[e];
    }
  }
}

int launch_tramat_jac(int np, int64_t nf, int nq, const double* K, const double* dK, const double* r, const double* dr,
                      int linsrc, double* dT, double* dL, int linprop, cudaStream_t stream) {
  if (nf == 0 || np < 2 || nq == 0) return 0;
  const int64_t n = nf * (np - 1);
  tramat_jac_kernel<<<static_cast<unsigned>((n + 63) / 64), 64, 0, stream>>>(np, nf, nq, K, dK, r, dr, linsrc, dT, dL, linprop);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// un-fused: rte_emission with Jacobians on materialised T, L, P, dT, dL, J, dJ; one thread per frequency.
// dI [nf][np][nq][4] must be zero on entry (m_spectral_radiance.cc:36-40).
// ---------------------------------------------------------------------------
template <bool LINSRC>
__global__ void rte_emission_jac_kernel(int np, int64_t nf, int nq, const double* __restrict__ T, const double* __restrict__ L,
                                        const double* __restrict__ P, const double* __restrict__ dT,
                                        const double* __restrict__ dL, const double* __restrict__ J,
                                        const double* __restrict__ dJ, const double* __restrict__ I_bkg,
                                        double* __restrict__ I, double* __restrict__ dI) {
  const int64_t iv = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (iv >= nf) return;
  const int64_t half = nf * int64_t(np) * nq * 16;
  double Iv[4];
#pragma unroll
  for (int e = 0; e < 4; e++) Iv[e] = I_bkg[iv * 4 + e];
  for (int i = np - 2; i >= 0; i--) {
    const double* Tm  = T + (iv * np + i + 1) * 16;
    const double* Lm  = LINSRC ? L + (iv * np + i + 1) * 16 : nullptr;
    const double* Pm  = P + (iv * np + i) * 16;
    const double* Ji  = J + (iv * np + i) * 4;
    const double* Ji1 = J + (iv * np + i + 1) * 4;
    double v[4], jd[4] = {0, 0, 0, 0}, Jm[4] = {0, 0, 0, 0};
    if (LINSRC) {
#pragma unroll
      for (int e = 0; e < 4; e++) { v[e] = Iv[e] - Ji1[e]; jd[e] = Ji1[e] - Ji[e]; }
    } else {
#pragma unroll
      for (int e = 0; e < 4; e++) { Jm[e] = (Ji[e] + Ji1[e]) * 0.5; v[e] = Iv[e] - Jm[e]; }
    }
    for (int q = 0; q < nq; q++) {
      const double* dT0 = dT + ((iv * np + i) * nq + q) * 16;
      const double* dT1 = dT + half + ((iv * np + i + 1) * nq + q) * 16;
      const double* dL0 = LINSRC ? dL + ((iv * np + i) * nq + q) * 16 : nullptr;
      const double* dL1 = LINSRC ? dL + half + ((iv * np + i + 1) * nq + q) * 16 : nullptr;
      const double* dj0 = dJ + ((iv * np + i) * nq + q) * 4;
      const double* dj1 = dJ + ((iv * np + i + 1) * nq + q) * 4;
      double c0[4], c1[4], o[4];
      layer_terms<LINSRC>(Tm, Lm, dT0, dT1, dL0, dL1, v, jd, dj0, dj1, c0, c1);
      double* d0 = dI + ((iv * np + i) * nq + q) * 4;
      double* d1 = dI + ((iv * np + i + 1) * nq + q) * 4;
      mv(Pm, c0, o);
#pragma unroll
      for (int e = 0; e < 4; e++) d0[e] += o[e];
      mv(Pm, c1, o);
#pragma unroll
      for (int e = 0; e < 4; e++) d1[e] += o[e];
    }
    double o[4];
    mv(Tm, v, o);
    if (LINSRC) {
      double l[4];
      mv(Lm, jd, l);
#pragma unroll
      for (int e = 0; e < 4; e++) Iv[e] = o[e] + l[e] + Ji[e];
    } else {
#pragma unroll
      for (int e = 0; e < 4; e++) Iv[e] = o[e] + Jm[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 4; e++) I[iv * 4 + e] = Iv[e];
}

int launch_rte_emission_jac(int linsrc, int np, int64_t nf, int nq, const double* T, const double* L, const double* P,
                            const double* dT, const double* dL, const double* J, const double* dJ, const double* I_bkg,
                            double* I, double* dI, cudaStream_t stream) {
  if (nf == 0) return 0;
  const unsigned grid = static_cast<unsigned>((nf + 63) / 64);
  if (linsrc)
    rte_emission_jac_kernel<true><<<grid, 64, 0, stream>>>(np, nf, nq, T, L, P, dT, dL, J, dJ, I_bkg, I, dI);
  else
    rte_emission_jac_kernel<false><<<grid, 64, 0, stream>>>(np, nf, nq, T, L, P, dT, dL, J, dJ, I_bkg, I, dI);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// fused pass B: forward from the sensor, one thread per frequency
//   I_lev [np][nf][4]: radiance arriving at level i from behind (written by stokes_chain_kernel, pass A)
//   dK    [np][nq][k_pitch... ] see StokesJacParams
// ---------------------------------------------------------------------------
// one finished element of spectral_rad_jac_path [nf][np][nq]: stored, and / or mapped to the state vector with the
// flat interpolation weights of path point i (m_rad.cc:107-125; the thread owns column iv of Jx: no atomics)
__device__ __forceinline__ void emit(const StokesJacParams& p, int64_t iv, int i, int q, double d0, double d1, double d2,
                                     double d3) {
  if (p.dI) {
    double* d = p.dI + ((iv * p.np + i) * p.nq + q) * 4;
    d[0] = d0; d[1] = d1; d[2] = d2; d[3] = d3;
  }
  if (p.Jx) {
    const int64_t row = int64_t(i) * p.nq + q;
    for (int64_t e = p.map_offset[row]; e < p.map_offset[row + 1]; e++) {
      const double w = p.map_w[e];
      if (w == 0.0) continue;
      double2* x = reinterpret_cast<double2*>(p.Jx + (int64_t(p.map_x[e]) * p.nf + iv) * 4);
      double2 a = x[0], b = x[1];
      a.x = __fma_rn(w, d0, a.x); a.y = __fma_rn(w, d1, a.y); b.x = __fma_rn(w, d2, b.x); b.y = __fma_rn(w, d3, b.y);
      x[0] = a; x[1] = b;
    }
  }
}

// LINPROP: the Dawson-function code (unpolarised closed form, polarised complex matrix form and its perturbation derivative)
// only exists in its own instantiation, so that the linsrc chain keeps its registers
template <bool LINSRC, bool LINPROP = false>
__global__ void __launch_bounds__(64, 1) stokes_jac_kernel(StokesJacParams p) {
  const int64_t iv = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (iv >= p.nf) return;
  const int np = p.np, nq = p.nq;
  double Pm[16];
  diag_of(Pm, 1.0);
  bool P_scalar = true;  // P is still a multiple of the identity (unpolarised so far)
  // carry of the dI1 contribution to level i from layer (i-1, i)
  double carry[AB200_MAX_TARGETS][4];
#pragma unroll
  for (int q = 0; q < AB200_MAX_TARGETS; q++) carry[q][0] = carry[q][1] = carry[q][2] = carry[q][3] = 0.0;

  Propmat k0 = load_propmat(p.K + (int64_t(0) * p.k_pitch + iv) * 7);
  double f0  = p.ffac[0] * p.f[iv];
  const bool emit_src = !p.no_emission;
  double j0  = (k0.is_rotational() || !emit_src) ? 0.0 : planck(f0, p.T[0]);
  for (int i = 0; i + 1 < np; i++) {
    const Propmat k1 = load_propmat(p.K + (int64_t(i + 1) * p.k_pitch + iv) * 7);
    const double f1  = p.ffac[i + 1] * p.f[int64_t(i + 1) * p.f_stride + iv];
    const double j1  = (k1.is_rotational() || !emit_src) ? 0.0 : planck(f1, p.T[i + 1]);
    const double ri  = p.r[i];
    Tran t;
    t.init(k0, k1, ri, false);
    double Tm[16], Lm[16];
    if (t.polarized) t.T(Tm); else diag_of(Tm, t.exp_a);
    const int lc = (LINSRC && LINPROP) ? linprop_case(k0.A, k1.A, ri, t.polarized) : 0;
    if (LINSRC) {
      if (LINPROP && lc == 2) linprop_lambda_pol(Tm, k0, k1, ri, 4, Lm);
      else if (LINPROP && lc == 1) diag_of(Lm, linprop_lambda(k0.A, k1.A, ri, t.exp_a));
      else if (t.polarized) t.L(Lm);
      else diag_of(Lm, func_F(t.a));
    }
    // radiance arriving at level i+1
    double v[4] = {0, 0, 0, 0}, jd[4] = {0, 0, 0, 0};
    if (nq > 0) {
      const double* Il = p.I_lev + (int64_t(i + 1) * p.nf + iv) * 4;
      if (LINSRC) {  // J0 = J[i+1], J1 = J[i] in the reference's naming
        v[0] = Il[0] - j1; v[1] = Il[1]; v[2] = Il[2]; v[3] = Il[3];
        jd[0] = j1 - j0;
      } else {
        const double jm = (j0 + j1) * 0.5;
        v[0] = Il[0] - jm; v[1] = Il[1]; v[2] = Il[2]; v[3] = Il[3];
      }
    }
    for (int q = 0; q < nq; q++) {
      const Propmat dk0 = load_propmat(p.dK + ((int64_t(i) * nq + q) * p.k_pitch + iv) * 7);
      const Propmat dk1 = load_propmat(p.dK + ((int64_t(i + 1) * nq + q) * p.k_pitch + iv) * 7);
      const double dr0 = p.dr[int64_t(i) * nq + q];
      const double dr1 = p.dr[(int64_t(np - 1) + i) * nq + q];
      const double dj0[4] = {(q == p.it && emit_src && !k0.is_rotational()) ? dplanck_dt(f0, p.T[i]) : 0.0, 0.0, 0.0, 0.0};
      const double dj1[4] = {(q == p.it && emit_src && !k1.is_rotational()) ? dplanck_dt(f1, p.T[i + 1]) : 0.0, 0.0, 0.0, 0.0};
      double dT0[16], dT1[16], dL0[16], dL1[16];
      t.deriv(Tm, k0, k1, dk0, ri, dr0, dT0);
      t.deriv(Tm, k0, k1, dk1, ri, dr1, dT1);
      if (LINSRC) {
        const bool lp = LINPROP;  // dr1 in both calls for linprop (:1238, sic)
        if (LINPROP && lc == 2) {
          linprop_lambda_pol_deriv(Lm, k0, k1, dk0, ri, dr1, true, dL0);
          linprop_lambda_pol_deriv(Lm, k0, k1, dk1, ri, dr1, false, dL1);
        } else if (LINPROP && lc == 1) {
          diag_of(dL0, linprop_lambda_deriv(k0.A, k1.A, dk0.A, Tm[0], dT0[0], ri, dr1, true));
          diag_of(dL1, linprop_lambda_deriv(k0.A, k1.A, dk1.A, Tm[0], dT1[0], ri, dr1, false));
        } else {
          t.linsrc_deriv(dk0, ri, lp ? dr1 : dr0, dL0);
          t.linsrc_deriv(dk1, ri, dr1, dL1);
        }
      }
      double c0[4], c1[4], o[4];
      layer_terms<LINSRC>(Tm, Lm, dT0, dT1, dL0, dL1, v, jd, dj0, dj1, c0, c1);
      mv(Pm, c0, o);
      emit(p, iv, i, q, carry[q][0] + o[0], carry[q][1] + o[1], carry[q][2] + o[2], carry[q][3] + o[3]);
      mv(Pm, c1, o);
      carry[q][0] = o[0]; carry[q][1] = o[1]; carry[q][2] = o[2]; carry[q][3] = o[3];
    }
    // P_{i+1} = P_i T_{i+1}, rtepack_transmission.cc:1322-1327
    if (P_scalar && !t.polarized) {
      const double s = Pm[0] * t.exp_a;
      Pm[0] = Pm[5] = Pm[10] = Pm[15] = s;
    } else {
      double o[16];
      mat_mul(Pm, Tm, o);
#pragma unroll
      for (int e = 0; e < 16; e++) Pm[e] = o[e];
      P_scalar = false;
    }
    k0 = k1; f0 = f1; j0 = j1;
  }
  for (int q = 0; q < nq; q++)  // last level only receives the dI1 term
    emit(p, iv, np - 1, q, carry[q][0], carry[q][1], carry[q][2], carry[q][3]);
  if (p.Jx && p.n_bkg > 0) {
    // spectral_rad_jacFromBackground (m_rad.cc:26-60) of spectral_radSurfaceBlackbody's Jacobian (m_background.cc:126-140):
    // P_{np-1} (w dB/dT, 0, 0, 0), with the sensor's frequency grid
    const double dB = dplanck_dt(p.f[iv], p.bkg_T);
    for (int b = 0; b < p.n_bkg; b++) {
      const double s = p.bkg_w[b] * dB;
      double* x = p.Jx + (int64_t(p.bkg_x[b]) * p.nf + iv) * 4;
      x[0] += Pm[0] * s; x[1] += Pm[4] * s; x[2] += Pm[8] * s; x[3] += Pm[12] * s;
    }
  }
}

// The same pass for an unpolarised path (every K and dK has only A != 0: all segments were real, pol = no): T, Lambda, their
// derivatives and the running P are scalars, the radiance may still carry Q, U, V from the background.  Same formulas as the
// unpolarised branches of tran::deriv (:566-569), linsrc_deriv (:289-292), linsrc_linprop_deriv (:493-541) and the
// recursion (rtepack_rtestep.cc:293-306, :348-367); ~60 registers instead of 255 and no local memory.
template <int OPT /* AB200_RTE_* */>
__global__ void __launch_bounds__(128) stokes_jac_scalar_kernel(StokesJacParams p) {
  constexpr bool LINSRC = OPT != AB200_RTE_CONSTANT;
  constexpr bool lp     = OPT == AB200_RTE_LINPROP;  // the Dawson-function code only exists in this instantiation
  const int64_t iv = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (iv >= p.nf) return;
  const int np = p.np, nq = p.nq;
  const bool emit_src = !p.no_emission;
  double P = 1.0;
  double carry[AB200_MAX_TARGETS][4];
#pragma unroll
  for (int q = 0; q < AB200_MAX_TARGETS; q++) carry[q][0] = carry[q][1] = carry[q][2] = carry[q][3] = 0.0;
  double k0 = p.K[(int64_t(0) * p.k_pitch + iv) * 7];
  double f0 = p.ffac[0] * p.f[iv];
  double j0 = (k0 == 0.0 || !emit_src) ? 0.0 : planck(f0, p.T[0]);
  for (int i = 0; i + 1 < np; i++) {
    const double k1 = p.K[(int64_t(i + 1) * p.k_pitch + iv) * 7];
    const double f1 = p.ffac[i + 1] * p.f[int64_t(i + 1) * p.f_stride + iv];
    const double j1 = (k1 == 0.0 || !emit_src) ? 0.0 : planck(f1, p.T[i + 1]);
    const double ri = p.r[i];
    const double a  = -0.5 * ri * (k0 + k1);
    const double T  = exp(a);
    const int lc    = lp ? linprop_case(k0, k1, ri, false) : 0;
    const double L  = !LINSRC ? 0.0 : ((lp && lc == 1) ? linprop_lambda(k0, k1, ri, T) : func_F(a));
    double v[4] = {0, 0, 0, 0}, jd = 0.0;
    if (nq > 0) {
      const double2* Il = reinterpret_cast<const double2*>(p.I_lev + (int64_t(i + 1) * p.nf + iv) * 4);
      const double2 i01 = Il[0], i23 = Il[1];
      v[0] = i01.x - (LINSRC ? j1 : (j0 + j1) * 0.5); v[1] = i01.y; v[2] = i23.x; v[3] = i23.y;
      if (LINSRC) jd = j1 - j0;
    }
    const double inv_r = (fabs(ri) > 1e-20) ? 1.0 / ri : 0.0;
    for (int q = 0; q < nq; q++) {
      const double dk0 = p.dK[((int64_t(i) * nq + q) * p.k_pitch + iv) * 7];
      const double dk1 = p.dK[((int64_t(i + 1) * nq + q) * p.k_pitch + iv) * 7];
      const double dr0 = p.dr[int64_t(i) * nq + q];
      const double dr1 = p.dr[(int64_t(np - 1) + i) * nq + q];
      const double dj0 = (q == p.it && emit_src && k0 != 0.0) ? dplanck_dt(f0, p.T[i]) : 0.0;
      const double dj1 = (q == p.it && emit_src && k1 != 0.0) ? dplanck_dt(f1, p.T[i + 1]) : 0.0;
      const double dT0 = -0.5 * (ri * dk0 + dr0 * (k0 + k1)) * T;
      const double dT1 = -0.5 * (ri * dk1 + dr1 * (k0 + k1)) * T;
      double c0[4], c1[4];
      if (LINSRC) {
        double dL0, dL1;
        if (lp && lc == 1) {
          dL0 = linprop_lambda_deriv(k0, k1, dk0, T, dT0, ri, dr1, true);
          dL1 = linprop_lambda_deriv(k0, k1, dk1, T, dT1, ri, dr1, false);
        } else {
          const double Fp = func_Fp(a);
          dL0 = Fp * (((lp ? dr1 : dr0) * inv_r) * a - 0.5 * ri * dk0);  // dr1 in both calls for linprop (:1238, sic)
          dL1 = Fp * ((dr1 * inv_r) * a - 0.5 * ri * dk1);
        }
        // dI0 += P (dJ1 - L dJ0 + dT0 ImJ0 + dL0 J0mJ1);  dI1 += P (dT1 ImJ0 + dL1 J0mJ1 + L dJ1 - T dJ0)
        c0[0] = dj1 - L * dj0 + dT0 * v[0] + dL0 * jd;
        c1[0] = dT1 * v[0] + dL1 * jd + L * dj1 - T * dj0;
      } else {
        c0[0] = dT0 * v[0] + midpoint(dj0, -(T * dj0));
        c1[0] = dT1 * v[0] + midpoint(dj1, -(T * dj1));
      }
#pragma unroll
      for (int e = 1; e < 4; e++) { c0[e] = dT0 * v[e]; c1[e] = dT1 * v[e]; }
      emit(p, iv, i, q, carry[q][0] + P * c0[0], carry[q][1] + P * c0[1], carry[q][2] + P * c0[2], carry[q][3] + P * c0[3]);
#pragma unroll
      for (int e = 0; e < 4; e++) carry[q][e] = P * c1[e];
    }
    P *= T;
    k0 = k1; f0 = f1; j0 = j1;
  }
  for (int q = 0; q < nq; q++) emit(p, iv, np - 1, q, carry[q][0], carry[q][1], carry[q][2], carry[q][3]);
  if (p.Jx && p.n_bkg > 0) {
    const double dB = dplanck_dt(p.f[iv], p.bkg_T);
    for (int b = 0; b < p.n_bkg; b++) p.Jx[(int64_t(p.bkg_x[b]) * p.nf + iv) * 4] += P * (p.bkg_w[b] * dB);
  }
}

int launch_stokes_jac(const StokesJacParams& p, cudaStream_t stream) {
  if (p.nf == 0 || p.np == 0 || (p.nq == 0 && !(p.Jx && p.n_bkg > 0))) return 0;
  if (p.scalar) {
    const unsigned grid = static_cast<unsigned>((p.nf + 127) / 128);
    if (p.rte_option == AB200_RTE_LINSRC)
      stokes_jac_scalar_kernel<AB200_RTE_LINSRC><<<grid, 128, 0, stream>>>(p);
    else if (p.rte_option == AB200_RTE_LINPROP)
      stokes_jac_scalar_kernel<AB200_RTE_LINPROP><<<grid, 128, 0, stream>>>(p);
    else
      stokes_jac_scalar_kernel<AB200_RTE_CONSTANT><<<grid, 128, 0, stream>>>(p);
    count_launch();
    AB_CUDA(cudaGetLastError());
    return 0;
  }
  const unsigned grid = static_cast<unsigned>((p.nf + 63) / 64);
  if (p.rte_option != AB200_RTE_CONSTANT)
    if (p.rte_option == AB200_RTE_LINPROP) stokes_jac_kernel<true, true><<<grid, 64, 0, stream>>>(p);
    else stokes_jac_kernel<true><<<grid, 64, 0, stream>>>(p);
  else
    stokes_jac_kernel<false><<<grid, 64, 0, stream>>>(p);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ab200
