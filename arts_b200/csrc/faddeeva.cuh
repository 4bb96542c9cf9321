// faddeeva.cuh — register-resident FP64 Faddeeva w(z) for the line-sum kernels.
//
// Replaces Faddeeva::w (reference 3rdparty/Faddeeva/Faddeeva.cc:659-935, called from
// src/core/lbl/lbl_lineshape_voigt_lte.cpp:239).  It is NOT a transcription: the
// reference's erfcx / w_im Chebyshev switch tables (100-way branches) and its
// data-dependent series loops are replaced by a branch-light design that keeps a warp
// converged:
//   * far wing  (|x|+y > 4000, the reference's nu<=2 region :707-725): closed form
//     evaluated in *line space* (Hz) with per-line constants prepared once per level,
//     7 FP64 instructions per (line, frequency) for the real part, one MUFU reciprocal.
//   * continued fraction (the reference's region :695-700 with its nu(z) fit :729) with
//     the same number of terms, reciprocal by MUFU + one cubic Newton step.
//   * core (everything else): the Zaghloul-Ali sampling sum (ACM TOMS 916, the method
//     behind Faddeeva.cc:786-931) rewritten as ONE fixed-length two-sided window of
//     2*12+1 terms centred at n0 = round(x/a); the y-only parts (erfcx(y) and
//     sum_n exp(-a^2n^2)/(a^2n^2+y^2)) are hoisted into a per-(line, level) constant E1(y)
//     computed by the prepare kernel, so the per-frequency work has no erfcx.
// Measured against mpmath (tools/proto_faddeeva.py): <= 5e-16 relative on both parts in
// the core region (the reference package itself is ~3.5e-14 there).
#pragma once

#include "common.cuh"

namespace ab200 {

namespace fad {
constexpr double A    = 0.518321480430085929872;  // pi / sqrt(-log(eps/2)), Faddeeva.cc:669
constexpr double A2   = 0.268657157075235951582;  // a^2
constexpr double CC   = 0.329973702884629072537;  // 2a/pi
constexpr double ISPI = 0.56418958354775628694807945156;
constexpr int W       = 12;  // half window; exp(-(a*(W+1/2))^2) ~ 5e-19
// exp(-a^2 k^2), k = 0..13
__device__ constexpr double T[14] = {
    1.0,
    7.64405281671221563e-01,
    3.41424527166548425e-01,
    8.91072646929412548e-02,
    1.35887299055460086e-02,
    1.21085455253437481e-03,
    6.30452613933449404e-05,
    1.91805156577114683e-06,
    3.40969447714832381e-08,
    3.54175089099469393e-10,
    2.14965079583260682e-12,
    7.62368911833724354e-15,
    1.57982797110681093e-17,
    1.91294189103582677e-20,
};
}  // namespace fad

// The reference's choice between continued fraction and series (Faddeeva.cc:695-700),
// for x = |Re z| >= 0 and y = Im z >= 0.
__device__ __forceinline__ bool cf_region(double x, double y) {
  return (y > 7.0) || (x > 6.0 && (y > 0.1 || (x > 8.0 && y > 1e-10) || x > 28.0));
}

// E1(y) = erfcx(y) - (2a/pi) y sum_{n>=1} exp(-a^2 n^2) / (a^2 n^2 + y^2): the y-only part
// of the core-region formula (coef1 / expx2 in Faddeeva.cc:869-886).  Per (line, level).
__device__ __forceinline__ double series_E1(double y) {
  const double y2 = y * y;
  double s1       = 0.0;
#pragma unroll
  for (int n = 13; n >= 1; --n) {
    const double t = fad::A * n;
    s1 = __fma_rn(fad::T[n], fast_rcp(__fma_rn(t, t, y2)), s1);  // (reciprocal to ~1 ulp: E1 is the same to 2e-16)
  }
  return erfcx(y) - fad::CC * y * s1;
}

// Core region, x >= 0, y >= 0, E1 = series_E1(y).
__device__ __forceinline__ void w_series(double x, double y, double E1, double& wr, double& wi) {
  const double expx2 = exp(-x * x);
  const double th    = x * y;
  double s, c;
  sincos(th, &s, &c);
  const double s2   = 2.0 * s * c;
  const double c2   = __fma_rn(-2.0 * s, s, 1.0);
  const double sinc = (th < 1e-4) ? __fma_rn(-th * th, 1.0 / 6.0, 1.0) : s / th;
  double re         = expx2 * __fma_rn(E1, c2, fad::CC * x * s * sinc);
  double im         = expx2 * __fma_rn(fad::CC * x * c, sinc, -E1 * s2);

  const double n0    = floor(__fma_rn(x, 1.0 / fad::A, 0.5));
  const double delta = __fma_rn(-fad::A, n0, x);
  const double g0    = exp(-delta * delta);
  const double p     = exp(2.0 * fad::A * delta);
  const double m     = fast_rcp(p);
  const double y2    = y * y;
  double sr = 0.0, si = 0.0;
  {
    const double t   = fad::A * n0;
    const bool valid = n0 != 0.0;
    const double den = valid ? __fma_rn(t, t, y2) : 1.0;
    const double v   = valid ? g0 * fast_rcp(den) : 0.0;
    sr               = v;
    si               = t * v;
  }
  double pk = g0, mk = g0;
#pragma unroll
  for (int k = 1; k <= fad::W; ++k) {
    pk *= p;
    mk *= m;
    const double tp = fad::A * (n0 + k);
    const double tm = fad::A * (n0 - k);
    const double vp = (fad::T[k] * pk) * fast_rcp(__fma_rn(tp, tp, y2));
    sr += vp;
    si = __fma_rn(tp, vp, si);
    const bool valid = (n0 - k) != 0.0;
    const double dm  = valid ? __fma_rn(tm, tm, y2) : 1.0;
    const double vm  = valid ? (fad::T[k] * mk) * fast_rcp(dm) : 0.0;
    sr += vm;
    si = __fma_rn(tm, vm, si);
  }
  if (x < 1e-2) {
    // n0 == 0 here: the +k / -k terms of the odd sum cancel down to O(x), which the running
    // products p^k, p^-k cannot resolve (Im w -> 0 linearly at the line centre; the
    // reference's Taylor branch, Faddeeva.cc:811-842).  Pair them analytically:
    // g0 (p^k - p^-k) = 2 g0 sinh(2 a x k), |2 a x k| <= 0.125 -> 5-term odd series (2e-17).
    si = 0.0;
#pragma unroll 1
    for (int k = 1; k <= fad::W; ++k) {
      const double t  = fad::A * k;
      const double u  = 2.0 * t * x;
      const double u2 = u * u;
      const double sh =
          u * __fma_rn(u2, __fma_rn(u2, __fma_rn(u2, __fma_rn(u2, 1.0 / 362880.0, 1.0 / 5040.0), 1.0 / 120.0), 1.0 / 6.0), 1.0);
      si = __fma_rn(t * fad::T[k], 2.0 * g0 * sh * fast_rcp(__fma_rn(t, t, y2)), si);
    }
  }
  wr = __fma_rn(0.5 * fad::CC * y, sr, re);
  wi = __fma_rn(0.5 * fad::CC, si, im);
}

// Continued fraction, x >= 0, y >= 0, same term count as Faddeeva.cc:726-741.
__device__ __forceinline__ void w_cf(double x, double y, double& wr_out, double& wi_out) {
  // The reference runs the backward recurrence w <- z - k / w for k = (nu-1)/2, ..., 1/2 from w = z with one complex
  // division per term (Faddeeva.cc:729-741).  The same finite continued fraction as a ratio N / D needs none:
  //   w = N / D,  z - k / w = (z N - k D) / N   =>   (N, D) <- (z N - k D, N),
  // six FMA-class instructions per term, and a single reciprocal at the end for w(z) = (i / sqrt(pi)) D / N.
  // |N| grows like |z|^nu <= 4000^4 or 6^20: no scaling needed.  Same term count nu(z) as the reference (:729).
  // (the quotient through the reciprocal: within 2 ulp of the division, so the floor differs only where 3.9 + q is that close to an
  // integer - and there one term more or less of a converged fraction changes nothing above 1e-16)
  const int nu = int(3.9 + 11.398 * fast_rcp(0.08254 * x + 0.1421 * y + 0.2023));  // floor of a positive number
  double Nr = x, Ni = y, Dr = 1.0, Di = 0.0;
  double k = 0.5 * double(nu - 1);
  for (int j = nu - 1; j > 0; j--) {
    const double tr = __fma_rn(x, Nr, __fma_rn(-y, Ni, -k * Dr));
    const double ti = __fma_rn(x, Ni, __fma_rn(y, Nr, -k * Di));
    Dr = Nr; Di = Ni; Nr = tr; Ni = ti;
    k -= 0.5;
  }
  const double den = fad::ISPI * fast_rcp(__fma_rn(Nr, Nr, Ni * Ni));
  wr_out           = den * __fma_rn(Dr, Ni, -Di * Nr);
  wi_out           = den * __fma_rn(Dr, Nr, Di * Ni);
}
// nu <= 2 closed form in z space (used only by the stand-alone evaluator below; the line
// kernels use the line-space form far_accumulate*()).
__device__ __forceinline__ void w_far_z(double x, double y, double& wr, double& wi) {
  const double q  = x * x;
  const double y2 = y * y;
  const double dr = q - (y2 + 0.5);
  const double d2 = __fma_rn(dr, dr, 4.0 * y2 * q);
  const double rc = fad::ISPI * fast_rcp(d2);
  wr              = y * (q + (y2 + 0.5)) * rc;
  wi              = x * (q + (y2 - 0.5)) * rc;
}

// w(x + i y) for non-far (x+y <= FAR_LIMIT) pairs with y >= 0; x may be negative.
__device__ __forceinline__ void w_near(double x, double y, double E1, double& wr, double& wi) {
  const double ax = fabs(x);
  if (cf_region(ax, y)) {
    w_cf(ax, y, wr, wi);
  } else {
    w_series(ax, y, E1, wr, wi);
  }
  if (x < 0.0) wi = -wi;  // w(-x + i y) = conj(w(x + i y))
}

// ---- far wing in line space --------------------------------------------------
// s*w(z) with z = igd*(u + i g), |z| large:  s*w = S*zeta/(zeta^2 - h), zeta = u + i g,
// S = i*s*GD/sqrt(pi), h = GD^2/2, GD = 1/igd (algebraically Faddeeva.cc:721-725).  With
// Q = u^2 + c3, c3 = g^2 - h (= |zeta|^2 - h, no cancellation in the far region |z| > 2000):
//   D2      = |zeta^2 - h|^2 = Q^2 + kappa,            kappa = 4 g^2 h >= 0
//   Re(s w) = [(A1 Q + B1) + (u A2) Q] / D2,           A1 = Si g,  B1 = 2 h A1, A2 = Sr
//   Im(s w) = [(A3 Q + B3) + (u A4) Q] / D2,           A3 = -Sr g, B3 = 2 h A3, A4 = Si
// The real part costs 7 FP64-pipe instructions per (line, frequency): DADD u, DFMA Q, DFMA D2,
// MUFU.RCP64H + 2 DFMA (one Newton step: 2^-40 ~ 9e-13 relative, three orders inside the 1e-9
// parity bound for sums of same-sign terms), DFMA numerator, DFMA accumulate — against the 28
// algorithmic flop of the reference's closed form.
// Explicit rounding intrinsics: the compiler may not re-associate or contract these, so the fast tile
// loop and the per-pair path of the general loop produce identical bits — this is what makes the
// result independent of the frequency tiling / sharding.
__device__ __forceinline__ double far_rcp(double D2) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(D2));
  const double e = __fma_rn(-D2, r, 1.0);
  return __fma_rn(r, e, r);
}

__device__ __forceinline__ double far_accumulate_re(double acc, double u, double c3, double kappa, double A1, double B1) {
  const double Q = __fma_rn(u, u, c3);
  const double r = far_rcp(__fma_rn(Q, Q, kappa));
  return __fma_rn(__fma_rn(A1, Q, B1), r, acc);
}

// complex part: cubic reciprocal step (the dispersive parts change sign, keep ~1e-16 per term)
__device__ __forceinline__ void far_accumulate_cplx(double& acc_re, double& acc_im, double u, double c3, double kappa,
                                                    double A1, double B1, double A2, double A3, double B3, double A4) {
  const double Q  = __fma_rn(u, u, c3);
  const double r  = fast_rcp(__fma_rn(Q, Q, kappa));
  const double nr = __fma_rn(Q, __fma_rn(u, A2, A1), B1);  // (A1 + u A2) Q + B1
  const double ni = __fma_rn(Q, __fma_rn(u, A4, A3), B3);  // (A3 + u A4) Q + B3
  acc_re          = __fma_rn(nr, r, acc_re);
  acc_im          = __fma_rn(ni, r, acc_im);
}

// real part only of the same sum: segments with pol = no scale with npm = (1, 0, ..., 0), their imaginary part is never used
__device__ __forceinline__ double far_accumulate_cplx_re(double acc_re, double u, double c3, double kappa, double A1, double B1,
                                                         double A2) {
  const double Q  = __fma_rn(u, u, c3);
  const double r  = fast_rcp(__fma_rn(Q, Q, kappa));
  const double nr = __fma_rn(Q, __fma_rn(u, A2, A1), B1);
  return __fma_rn(nr, r, acc_re);
}

// Four terms of the same continued fraction, w = (i/sqrt(pi)) / (z - (1/2)/(z - 1/(z - (3/2)/z))), collapsed to one
// rational function: with t = z^2,  w = (i/sqrt(pi)) z (t - 5/2) / (t^2 - 3 t + 3/4).  One reciprocal, ~24 FP64
// instructions; the truncation error is ~1.5 / |z|^8: <= 1.3e-12 of Re w and of Im w for |x| + y >= 48.  x >= 0, y >= 0.
__device__ __forceinline__ void w_mid(double x, double y, double& wr, double& wi) {
  const double tr = __fma_rn(x, x, -(y * y)), ti = 2.0 * x * y;
  const double a  = tr - 2.5;
  const double Nr = __fma_rn(x, a, -(y * ti)), Ni = __fma_rn(x, ti, y * a);
  const double Dr = __fma_rn(tr, tr - 3.0, __fma_rn(-ti, ti, 0.75)), Di = ti * __fma_rn(2.0, tr, -3.0);
  const double r  = fad::ISPI * fast_rcp(__fma_rn(Dr, Dr, Di * Di));
  wr = __fma_rn(Nr, Di, -(Ni * Dr)) * r;  // -Im(N conj D) / (sqrt(pi) |D|^2)
  wi = __fma_rn(Nr, Dr, Ni * Di) * r;     //  Re(N conj D) / (sqrt(pi) |D|^2)
}
// w(x + i y) for the near pairs of the FORWARD line sums (x + y <= the kernel's far limit): the closed form above from
// MID_LIMIT on, the reference's regions below
__device__ __forceinline__ void w_near_fast(double x, double y, double E1, double& wr, double& wi) {
  const double ax = fabs(x);
  if (ax + y > MID_LIMIT) {
    w_mid(ax, y, wr, wi);
  } else if (cf_region(ax, y)) {
    w_cf(ax, y, wr, wi);
  } else {
    w_series(ax, y, E1, wr, wi);
  }
  if (x < 0.0) wi = -wi;
}

// Stand-alone w(z) for arbitrary finite z (tests, ab200_faddeeva_w); y < 0 through the
// reflection w(z) = 2 exp(-z^2) - w(-z) like Faddeeva.cc:742-748.
__device__ inline void faddeeva_w(double zr, double zi, double& wr, double& wi) {
  const double ya = fabs(zi);
  const double xs = zi < 0.0 ? -zr : zr;
  const double ax = fabs(xs);
  double r, i;
  if (ax + ya > FAR_LIMIT) {
    w_far_z(ax, ya, r, i);
  } else if (cf_region(ax, ya)) {
    w_cf(ax, ya, r, i);
  } else {
    w_series(ax, ya, series_E1(ya), r, i);
  }
  if (xs < 0.0) i = -i;
  if (zi < 0.0) {
    // 2 exp(-z^2) = 2 exp((ya - xs)(xs + ya)) * (cos(2 xs y) + i sin(2 xs y)), y = zi
    const double mag = 2.0 * exp((ya - xs) * (xs + ya));
    double s, c;
    sincos(2.0 * xs * zi, &s, &c);
    r = mag * c - r;
    i = mag * s - i;
  }
  wr = r;
  wi = i;
}

}  // namespace ab200
