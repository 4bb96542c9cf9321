// rtepack.cuh — device versions of the rtepack value algebra used by stage 2.
//
// Follows the arithmetic of reference src/core/rtepack/rtepack_transmission.cc
// (tran ctor :22-116, operator() :118-150, linsrc :207-275), rtepack_source.cc:88-95,
// physics_funcs.cc:192-197 and rtepack_multitype.h:58-68.  Everything lives in registers:
// the fused kernel (stokes.cu) never materialises T, Lambda or J in memory.
#pragma once

#include "common.cuh"
#include "faddeeva.cuh"

namespace ab200 {
namespace rte {

constexpr double too_small = 1e-4;  // rtepack_transmission.cc:20

struct Propmat {
  double A, B, C, D, U, V, W;
  __device__ __forceinline__ bool is_rotational() const { return A == 0.0 && B == 0.0 && C == 0.0 && D == 0.0; }
  __device__ __forceinline__ bool is_polarized() const {
    return B != 0 || C != 0 || D != 0 || U != 0 || V != 0 || W != 0;
  }
};

__device__ __forceinline__ Propmat load_propmat(const double* __restrict__ p) {
  return Propmat{p[0], p[1], p[2], p[3], p[4], p[5], p[6]};
}

// planck, physics_funcs.cc:192-197
__device__ __forceinline__ double planck(double f, double t) {
  constexpr double a = 2 * cst::h / (cst::c * cst::c);
  constexpr double b = cst::h / cst::k;
  return a * (f * f * f) / expm1((b * f) / t);
}
// dplanck_dt, physics_funcs.cc:254-263
__device__ __forceinline__ double dplanck_dt(double f, double t) {
  constexpr double a = 2 * cst::h / (cst::c * cst::c);
  constexpr double b = cst::h / cst::k;
  const double inv_exp_t_m1 = 1.0 / expm1(b * f / t);
  const double f2 = f * f;
  return a * b * (f2 * f2) * inv_exp_t_m1 * (1 + inv_exp_t_m1) / (t * t);
}
// invplanck, physics_funcs.cc:153-158
__device__ __forceinline__ double invplanck(double i, double f) {
  constexpr double a = cst::h / cst::k;
  constexpr double b = 2 * cst::h / (cst::c * cst::c);
  return (a * f) / log1p((b * f * f * f) / i);
}

__device__ __forceinline__ double func_F(double z) {  // :208-210
  return fabs(z) < 1e-8 ? 1.0 + z * 0.5 + z * z / 6.0 : expm1(z) / z;
}
__device__ __forceinline__ double func_Fp(double z) {  // :212-216
  if (fabs(z) < too_small) return 0.5 + z / 3.0 + z * z / 8.0;
  const double ez = exp(z);
  return (ez * (z - 1.0) + 1.0) / (z * z);
}
__device__ __forceinline__ double func_Fpp(double z) {  // :291-295
  if (fabs(z) < too_small) return 1.0 / 3.0 + z / 4.0 + z * z / 10.0;
  const double ez = exp(z);
  return (ez * (z * z - 2.0 * z + 2.0) - 2.0) / (z * z * z);
}
__device__ __forceinline__ double func_F3p(double z) {  // :297-303
  if (fabs(z) < too_small) return 0.25 + z * 0.2;
  const double ez = exp(z);
  const double z2 = z * z;
  const double z3 = z2 * z;
  return (ez * (z3 - 3.0 * z2 + 6.0 * z - 6.0) + 6.0) / (z3 * z);
}
__device__ __forceinline__ double func_F4p(double z) {  // :305-313
  if (fabs(z) < too_small) return 0.2 + z / 6.0;
  const double ez = exp(z);
  const double z2 = z * z;
  const double z3 = z2 * z;
  const double z4 = z2 * z2;
  const double z5 = z4 * z;
  return (ez * (z4 - 4.0 * z3 + 12.0 * z2 - 24.0 * z + 24.0) - 24.0) / z5;
}

// Faddeeva::Dawson(x) for real x (reference 3rdparty/Faddeeva, used by tran::linsrc_linprop
// rtepack_transmission.cc:453): D(x) = sqrt(pi)/2 Im w(x + 0i), odd; through the register-resident w(z).
__device__ __forceinline__ double dawson(double x) {
  const double ax = fabs(x);
  double wr, wi;
  if (ax > FAR_LIMIT) w_far_z(ax, 0.0, wr, wi);
  else if (cf_region(ax, 0.0)) w_cf(ax, 0.0, wr, wi);
  else w_series(ax, 0.0, 1.0 /* E1(0) = erfcx(0) */, wr, wi);
  const double d = 0.886226925452758013649083741671 * wi;  // sqrt(pi)/2
  return x < 0.0 ? -d : d;
}

// rte_option = linprop (tran::linsrc_linprop, rtepack_transmission.cc:449-475): which branch a layer takes.
//   0: gradient below 1e-8 -> the linsrc operator (:456-457)   1: unpolarised Dawson form (:459-465)
//   2: polarised (complex matrix sqrt / dawson, :467-474) -> outside the GPU path, reported as an error
__device__ __forceinline__ int linprop_case(double k1A, double k2A, double r, bool polarized) {
  const double alpha2 = (k2A - k1A) / (2.0 * r);
  if (alpha2 < 1e-8) return 0;
  return polarized ? 2 : 1;
}
__device__ __forceinline__ double linprop_lambda(double k1A, double k2A, double r, double t00) {
  const double alpha = sqrt((k2A - k1A) / (2.0 * r));
  const double u0    = k1A / (2.0 * alpha);
  const double u1    = k2A / (2.0 * alpha);
  return (dawson(u1) - t00 * dawson(u0)) / (r * alpha);
}
// tran::linsrc_linprop_deriv, unpolarised closed form :493-541
__device__ __forceinline__ double linprop_lambda_deriv(double k1a, double k2a, double dk, double t00, double dt00, double r,
                                                       double dr, bool k1_deriv) {
  const double denom = 2.0 * r;
  const double alpha = sqrt(fmax(0.0, (k2a - k1a) / denom));
  const double u0 = k1a / (2.0 * alpha), u1 = k2a / (2.0 * alpha);
  const double D0 = dawson(u0), D1 = dawson(u1);
  const double dD0 = 1.0 - 2.0 * u0 * D0, dD1 = 1.0 - 2.0 * u1 * D1;
  double d_alpha, d_u0, d_u1;
  if (k1_deriv) {
    d_alpha = -0.5 * dk / (denom * alpha);
    d_u0    = (dk * 2.0 * alpha - k1a * 2.0 * d_alpha) / (4.0 * alpha * alpha);
    d_u1    = -k2a * d_alpha / (2.0 * alpha * alpha);
  } else {
    d_alpha = 0.5 * dk / (denom * alpha);
    d_u0    = -k1a * d_alpha / (2.0 * alpha * alpha);
    d_u1    = (dk * 2.0 * alpha - k2a * 2.0 * d_alpha) / (4.0 * alpha * alpha);
  }
  const double d_num       = dD1 * d_u1 - dt00 * D0 - t00 * dD0 * d_u0 - t00 * D0 * d_u0;  // sic, :531-532
  const double denom_val   = r * alpha;
  const double d_denom_val = dr * alpha + r * d_alpha;
  return (d_num * denom_val - (D1 - t00 * D0) * d_denom_val) / (denom_val * denom_val);
}

// 4x4 row-major helpers
__device__ __forceinline__ void mat_vec(const double* __restrict__ m, const double* __restrict__ s, double* __restrict__ o) {
  // rtepack_multitype.h:58-68 (same operand order)
  o[0] = m[0] * s[0] + m[1] * s[1] + m[2] * s[2] + m[3] * s[3];
  o[1] = m[4] * s[0] + m[5] * s[1] + m[6] * s[2] + m[7] * s[3];
  o[2] = m[9] * s[1] + m[10] * s[2] + m[11] * s[3] + m[8] * s[0];
  o[3] = m[12] * s[0] + m[13] * s[1] + m[14] * s[2] + m[15] * s[3];
}
__device__ __forceinline__ void mat_mul(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ o) {
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      o[4 * i + j] = a[4 * i + 0] * b[j] + a[4 * i + 1] * b[4 + j] + a[4 * i + 2] * b[8 + j] + a[4 * i + 3] * b[12 + j];
}

// rtepack::tran, rtepack_transmission.cc:22-116.  exact == false reproduces the reference's
// literal lines :67-70 (x2 = sqrt(t1), x = sqrt(x2)); exact == true uses x2 = t1 (the true
// eigenvalue squares), see DESIGN.md "reference quirks".
struct Tran {
  double a, exp_a;
  double b, c, d, u, v, w;
  double b2, c2, d2, u2, v2, w2;
  double B, C, S;
  double x2, y2, x, y, cy, sy, cx, sx;
  double ix, iy, inv_x2y2;
  double C0, C1, C2, C3;
  bool polarized, x_zero, y_zero, both_zero, either_zero;

  __device__ __forceinline__ void init(const Propmat& k1, const Propmat& k2, double r, bool exact) {
    a         = -0.5 * r * (k1.A + k2.A);
    exp_a     = exp(a);
    polarized = k1.is_polarized() || k2.is_polarized();
    if (!polarized) return;
    b = -0.5 * r * (k1.B + k2.B);
    c = -0.5 * r * (k1.C + k2.C);
    d = -0.5 * r * (k1.D + k2.D);
    u = -0.5 * r * (k1.U + k2.U);
    v = -0.5 * r * (k1.V + k2.V);
    w = -0.5 * r * (k1.W + k2.W);
    b2 = b * b; c2 = c * c; d2 = d * d; u2 = u * u; v2 = v * v; w2 = w * w;
    B = u2 + v2 + w2 - b2 - c2 - d2;
    const double t = d * u - c * v + b * w;
    C = -(t * t);
    const double disc = B * B - 4 * C;
    S = sqrt(fmax(0.0, disc));
    const double t1 = 0.5 * (S - B);
    const double t2 = 0.5 * (S + B);
    if (exact) {
      x2 = fmax(0.0, t1);
      y2 = fmax(0.0, t2);
    } else {
      x2 = sqrt(fmax(0.0, t1));
      y2 = sqrt(fmax(0.0, t2));
    }
    x = sqrt(x2);
    y = sqrt(y2);
    sincos(y, &sy, &cy);
    cx = cosh(x);
    sx = sinh(x);
    x_zero      = x < too_small;
    y_zero      = y < too_small;
    both_zero   = y_zero && x_zero;
    either_zero = y_zero || x_zero;
    ix       = x_zero ? 0.0 : 1.0 / x;
    iy       = y_zero ? 0.0 : 1.0 / y;
    inv_x2y2 = both_zero ? 1.0 : 1.0 / (x2 + y2);
    C0 = either_zero ? 1.0 : (cy * x2 + cx * y2) * inv_x2y2;
    C1 = either_zero ? 1.0 : (sy * x2 * iy + sx * y2 * ix) * inv_x2y2;
    C2 = both_zero ? 0.5 : (cx - cy) * inv_x2y2;
    C3 = both_zero ? 1.0 / 6.0 : ((x_zero ? 1.0 : sx * ix) - (y_zero ? 1.0 : sy * iy)) * inv_x2y2;
    polarized = isfinite(C0) && isfinite(C1) && isfinite(C2) && isfinite(C3);
  }

  // operator()(), :118-150 (polarised branch; callers handle !polarized with exp_a)
  __device__ __forceinline__ void T(double* __restrict__ m) const {
    const double C2b = C2 * (c * u + d * v);
    const double C2c = C2 * (b * u - d * w);
    const double C2d = C2 * (b * v + c * w);
    const double C2u = C2 * (b * c - v * w);
    const double C2v = C2 * (b * d + u * w);
    const double C2w = C2 * (c * d - u * v);
    const double C3b = C3 * (b * (B - w2) + w * (c * v - d * u));
    const double C3c = C3 * (c * (v2 - B) - v * (d * u + b * w));
    const double C3d = C3 * (d * (u2 - B) - u * (c * v - b * w));
    const double C3u = C3 * (d * (c * v - b * w) - u * (B + d2));
    const double C3v = C3 * (c * (d * u + b * w) - v * (B + c2));
    const double C3w = C3 * (b * (c * v - d * u) - w * (B + b2));
    m[0]  = exp_a * (C0 + C2 * (b2 + c2 + d2));
    m[1]  = exp_a * (C1 * b - C2b - C3b);
    m[2]  = exp_a * (C1 * c + C2c + C3c);
    m[3]  = exp_a * (C1 * d + C2d + C3d);
    m[4]  = exp_a * (C1 * b + C2b - C3b);
    m[5]  = exp_a * (C0 + C2 * (b2 - u2 - v2));
    m[6]  = exp_a * (C1 * u + C2u + C3u);
    m[7]  = exp_a * (C1 * v + C2v + C3v);
    m[8]  = exp_a * (C1 * c - C2c + C3c);
    m[9]  = exp_a * (-C1 * u + C2u - C3u);
    m[10] = exp_a * (C0 + C2 * (c2 - u2 - w2));
    m[11] = exp_a * (C1 * w + C2w + C3w);
    m[12] = exp_a * (C1 * d - C2d + C3d);
    m[13] = exp_a * (-C1 * v + C2v - C3v);
    m[14] = exp_a * (-C1 * w + C2w - C3w);
    m[15] = exp_a * (C0 + C2 * (d2 - v2 - w2));
  }

  // coefficients of Lambda = l0 I + l1 S + l2 S^2 + l3 S^3, :220-266 (polarised branch)
  __device__ __forceinline__ void linsrc_coeffs(double& l0, double& l1, double& l2, double& l3) const {
    if (both_zero) {
      l0 = func_F(a);
      l1 = func_Fp(a);
      if (fabs(a) < too_small) {
        l2 = 1.0 / 6.0 + a / 12.0;
        l3 = 1.0 / 24.0 + a / 60.0;
      } else {
        const double ez = exp(a);
        const double a2 = a * a;
        const double a3 = a2 * a;
        l2 = 0.5 * (ez * (a2 - 2.0 * a + 2.0) - 2.0) / a3;
        l3 = (ez * (a3 - 3.0 * a2 + 6.0 * a - 6.0) + 6.0) / (6.0 * a2 * a2);
      }
      return;
    }
    double Pp, Pm_div_x;
    if (x_zero) {
      Pp       = func_F(a);
      Pm_div_x = func_Fp(a);
    } else {
      const double f1 = func_F(a + x);
      const double f2 = func_F(a - x);
      Pp       = 0.5 * (f1 + f2);
      Pm_div_x = 0.5 * (f1 - f2) / x;
    }
    double Qp, q_im;
    if (y_zero) {
      Qp   = func_F(a);
      q_im = func_Fp(a);
    } else {
      const double denom       = a * a + y * y;
      const double ea_cy_m1    = exp_a * cy - 1.0;
      const double ea_sy       = exp_a * sy;
      Qp                       = (a * ea_cy_m1 + y * ea_sy) / denom;
      const double sin_y_div_y = (fabs(y) < 1e-6) ? 1.0 - y * y / 6.0 : sy / y;
      q_im                     = (a * exp_a * sin_y_div_y - ea_cy_m1) / denom;
    }
    l2 = (Pp - Qp) * inv_x2y2;
    l0 = Pp - l2 * x2;
    l3 = (Pm_div_x - q_im) * inv_x2y2;
    l1 = Pm_div_x - l3 * x2;
  }

  __device__ __forceinline__ void S_mat(double* __restrict__ s) const {
    s[0] = 0;  s[1] = b;   s[2] = c;   s[3] = d;
    s[4] = b;  s[5] = 0;   s[6] = u;   s[7] = v;
    s[8] = c;  s[9] = -u;  s[10] = 0;  s[11] = w;
    s[12] = d; s[13] = -v; s[14] = -w; s[15] = 0;
  }

  // full Lambda, :268-274 (polarised branch)
  __device__ __forceinline__ void L(double* __restrict__ m) const {
    double l0, l1, l2, l3;
    linsrc_coeffs(l0, l1, l2, l3);
    double s[16], s2[16], s3[16];
    S_mat(s);
    mat_mul(s, s, s2);
    mat_mul(s, s2, s3);
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = s[i] * l1 + s2[i] * l2 + s3[i] * l3;
    m[0] += l0; m[5] += l0; m[10] += l0; m[15] += l0;
  }

  // d/dq of operator()(), rtepack_transmission.cc:558-674: t = T of this layer, dk = dK/dq of ONE end
  // level, dr = d r / dq (hydrostatic term).  Literal (quirk 6) arithmetic only.
  __device__ __forceinline__ void deriv(const double* __restrict__ t, const Propmat& k1, const Propmat& k2, const Propmat& dk,
                                        double r, double dr, double* __restrict__ m) const {
    const double da = -0.5 * (r * dk.A + dr * (k1.A + k2.A));
    if (!polarized) {
#pragma unroll
      for (int i = 0; i < 16; i++) m[i] = 0.0;
      m[0] = m[5] = m[10] = m[15] = da * exp_a;
      return;
    }
    const double db  = -0.5 * (r * dk.B + dr * (k1.B + k2.B));
    const double dc  = -0.5 * (r * dk.C + dr * (k1.C + k2.C));
    const double dd  = -0.5 * (r * dk.D + dr * (k1.D + k2.D));
    const double du  = -0.5 * (r * dk.U + dr * (k1.U + k2.U));
    const double dv  = -0.5 * (r * dk.V + dr * (k1.V + k2.V));
    const double dw  = -0.5 * (r * dk.W + dr * (k1.W + k2.W));
    const double db2 = 2 * db * b, dc2 = 2 * dc * c, dd2 = 2 * dd * d, du2 = 2 * du * u, dv2 = 2 * dv * v, dw2 = 2 * dw * w;
    const double dB  = du2 + dv2 + dw2 - db2 - dc2 - dd2;
    const double dC  = -2 * (d * u - c * v + b * w) * (dd * u + d * du - dc * v - c * dv + db * w + b * dw);
    const double dS  = (B * dB - 2 * dC) / S;
    const double dx2 = 0.25 * (dS - dB) / x2;
    const double dy2 = 0.25 * (dS + dB) / y2;
    const double dx  = 0.5 * dx2 / x;
    const double dy  = 0.5 * dy2 / y;
    const double dcy = -sy * dy, dsy = cy * dy, dcx = sx * dx, dsx = cx * dx;
    const double dix = -dx * ix * ix, diy = -dy * iy * iy;
    const double dx2dy2 = dx2 + dy2;
    const double dC0 = either_zero ? 0.0 : (dcy * x2 + cy * dx2 + dcx * y2 + cx * dy2 - C0 * dx2dy2) * inv_x2y2;
    const double dC1 = either_zero ? 0.0
                                   : (dsy * x2 * iy + sy * dx2 * iy + sy * x2 * diy + dsx * y2 * ix + sx * dy2 * ix +
                                      sx * y2 * dix - C1 * dx2dy2) *
                                         inv_x2y2;
    const double dC2 = both_zero ? 0.0 : ((x_zero ? 0.0 : (dcx - C2 * dx2)) - (y_zero ? 0.0 : (dcy + C2 * dy2))) * inv_x2y2;
    const double dC3 = both_zero ? 0.0
                                 : ((x_zero ? 0.0 : (dsx * ix + sx * dix - C3 * dx2)) -
                                    (y_zero ? 0.0 : (dsy * iy + sy * diy + C3 * dy2))) *
                                       inv_x2y2;
    const double dC2b = dC2 * (c * u + d * v) + C2 * (dc * u + c * du + dd * v + d * dv);
    const double dC2c = dC2 * (b * u - d * w) + C2 * (db * u + b * du - dd * w - d * dw);
    const double dC2d = dC2 * (b * v + c * w) + C2 * (db * v + b * dv + dc * w + c * dw);
    const double dC2u = dC2 * (b * c - v * w) + C2 * (db * c + b * dc - dv * w - v * dw);
    const double dC2v = dC2 * (b * d + u * w) + C2 * (db * d + b * dd + du * w + u * dw);
    const double dC2w = dC2 * (c * d - u * v) + C2 * (dc * d + c * dd - du * v - u * dv);
    const double dC3b = dC3 * (b * (B - w2) + w * (c * v - d * u)) +
                        C3 * (db * (B - w2) + b * (dB - dw2) + dw * (c * v - d * u) + w * (dc * v + c * dv - dd * u - d * du));
    const double dC3c = dC3 * (c * (v2 - B) - v * (d * u + b * w)) +
                        C3 * (dc * (v2 - B) + c * (dv2 - dB) - dv * (d * u + b * w) - v * (dd * u + d * du + db * w + b * dw));
    const double dC3d = dC3 * (d * (u2 - B) - u * (c * v - b * w)) +
                        C3 * (dd * (u2 - B) + d * (du2 - dB) - du * (c * v - b * w) - u * (dc * v + c * dv - db * w - b * dw));
    const double dC3u = dC3 * (d * (c * v - b * w) - u * (B + d2)) +
                        C3 * (dd * (c * v - b * w) + d * (dc * v + c * dv - db * w - b * dw) - du * (B + d2) - u * (dB + dd2));
    const double dC3v = dC3 * (c * (d * u + b * w) - v * (B + c2)) +
                        C3 * (dc * (d * u + b * w) + c * (dd * u + d * du + db * w + b * dw) - dv * (B + c2) - v * (dB + dc2));
    const double dC3w = dC3 * (b * (c * v - d * u) - w * (B + b2)) +
                        C3 * (db * (c * v - d * u) + b * (dc * v + c * dv - dd * u - d * du) - dw * (B + b2) - w * (dB + db2));
    const double dM00 = dC0 + dC2 * (b2 + c2 + d2) + C2 * (db2 + dc2 + dd2);
    const double dM11 = dC0 + dC2 * (b2 - u2 - v2) + C2 * (db2 - du2 - dv2);
    const double dM22 = dC0 + dC2 * (c2 - u2 - w2) + C2 * (dc2 - du2 - dw2);
    const double dM33 = dC0 + dC2 * (d2 - v2 - w2) + C2 * (dd2 - dv2 - dw2);
    const double e[16] = {dM00,
                          dC1 * b + C1 * db - dC2b - dC3b,
                          dC1 * c + C1 * dc + dC2c + dC3c,
                          dC1 * d + C1 * dd + dC2d + dC3d,
                          dC1 * b + C1 * db + dC2b - dC3b,
                          dM11,
                          dC1 * u + C1 * du + dC2u + dC3u,
                          dC1 * v + C1 * dv + dC2v + dC3v,
                          dC1 * c + C1 * dc - dC2c + dC3c,
                          -dC1 * u - C1 * du + dC2u - dC3u,
                          dM22,
                          dC1 * w + C1 * dw + dC2w + dC3w,
                          dC1 * d + C1 * dd - dC2d + dC3d,
                          -dC1 * v - C1 * dv + dC2v - dC3v,
                          -dC1 * w - C1 * dw + dC2w - dC3w,
                          dM33};
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = da * t[i] + exp_a * e[i];
  }

  // d/dq of linsrc(), rtepack_transmission.cc:277-447
  __device__ __forceinline__ void linsrc_deriv(const Propmat& dk, double r, double dr, double* __restrict__ m) const {
    const double inv_r = (fabs(r) > 1e-20) ? 1.0 / r : 0.0;
    const double dr_r  = dr * inv_r;
    const double da    = dr_r * a - 0.5 * r * dk.A;
    if (!polarized) {
#pragma unroll
      for (int i = 0; i < 16; i++) m[i] = 0.0;
      m[0] = m[5] = m[10] = m[15] = func_Fp(a) * da;
      return;
    }
    const double db  = dr_r * b - 0.5 * r * dk.B;
    const double dc  = dr_r * c - 0.5 * r * dk.C;
    const double dd  = dr_r * d - 0.5 * r * dk.D;
    const double du  = dr_r * u - 0.5 * r * dk.U;
    const double dv  = dr_r * v - 0.5 * r * dk.V;
    const double dw  = dr_r * w - 0.5 * r * dk.W;
    const double db2 = 2.0 * db * b, dc2 = 2.0 * dc * c, dd2 = 2.0 * dd * d, du2 = 2.0 * du * u, dv2 = 2.0 * dv * v,
                 dw2 = 2.0 * dw * w;
    const double dB  = du2 + dv2 + dw2 - db2 - dc2 - dd2;
    const double dC  = -2.0 * (d * u - c * v + b * w) * (dd * u + d * du - dc * v - c * dv + db * w + b * dw);
    const double dS_val = (S > 1e-9) ? (B * dB - 2.0 * dC) / S : 0.0;
    const double dx2    = (x2 > 1e-9) ? 0.25 * (dS_val - dB) / x2 : 0.0;
    const double dy2    = (y2 > 1e-9) ? 0.25 * (dS_val + dB) / y2 : 0.0;
    const double dx     = (x > 1e-9) ? 0.5 * dx2 / x : 0.0;
    const double dy     = (y > 1e-9) ? 0.5 * dy2 / y : 0.0;
    double l1, l2, l3, dl0, dl1, dl2, dl3;
    if (both_zero) {
      const double fpa  = func_Fp(a);
      const double fppa = func_Fpp(a);
      dl0 = fpa * da;
      l1  = fpa;
      dl1 = fppa * da;
      if (fabs(a) < too_small) {
        l2  = 1.0 / 6.0 + a / 12.0;
        dl2 = da / 12.0;
        l3  = 1.0 / 24.0 + a / 60.0;
        dl3 = da / 60.0;
      } else {
        const double f3pa = func_F3p(a);
        const double f4pa = func_F4p(a);
        l2  = 0.5 * fppa;
        dl2 = 0.5 * f3pa * da;
        l3  = f3pa / 6.0;
        dl3 = f4pa / 6.0 * da;
      }
    } else {
      double Pp = 0.0, Pm_div_x = 0.0, Qp = 0.0, q_im = 0.0, dPp = 0.0, dPm_div_x = 0.0, dQp = 0.0, dq_im = 0.0;
      if (x_zero) {
        Pp        = func_F(a);
        dPp       = func_Fp(a) * da + 0.5 * func_Fpp(a) * dx2;
        Pm_div_x  = func_Fp(a);
        dPm_div_x = func_Fpp(a) * da + (func_F3p(a) / 6.0) * dx2;
      } else {
        const double f_apx = func_F(a + x), f_amx = func_F(a - x);
        const double fp_apx = func_Fp(a + x), fp_amx = func_Fp(a - x);
        Pp                   = 0.5 * (f_apx + f_amx);
        const double sum_fp  = fp_apx + fp_amx;
        const double diff_fp = fp_apx - fp_amx;
        dPp                  = 0.5 * sum_fp * da + 0.5 * diff_fp * dx;
        Pm_div_x             = 0.5 * (f_apx - f_amx) / x;
        const double dPm_da  = 0.5 * diff_fp / x;
        double dPm_dx;
        if (x < 1e-3) {
          dPm_dx = func_F3p(a) * x / 3.0;
        } else {
          const double diff_f = f_apx - f_amx;
          dPm_dx              = (x * sum_fp - diff_f) / (2.0 * x * x);
        }
        dPm_div_x = dPm_da * da + dPm_dx * dx;
      }
      if (y_zero) {
        Qp    = func_F(a);
        dQp   = func_Fp(a) * da - 0.5 * func_Fpp(a) * dy2;
        q_im  = func_Fp(a);
        dq_im = func_Fpp(a) * da - (func_F3p(a) / 6.0) * dy2;
      } else {
        const double denom       = a * a + y * y;
        const double ea_cy_m1    = exp_a * cy - 1.0;
        const double ea_sy       = exp_a * sy;
        Qp                       = (a * ea_cy_m1 + y * ea_sy) / denom;
        const double sin_y_div_y = (fabs(y) < 1e-6) ? 1.0 - y * y / 6.0 : sy / y;
        q_im                     = (a * exp_a * sin_y_div_y - ea_cy_m1) / denom;
        const double ImF         = (a * exp_a * sy - y * ea_cy_m1) / denom;
        const double A_val       = exp_a * ((a - 1.0) * cy - y * sy) + 1.0;
        const double B_val       = exp_a * ((a - 1.0) * sy + y * cy);
        const double C_val       = a * a - y * y;
        const double D_val       = 2.0 * a * y;
        const double denom2      = C_val * C_val + D_val * D_val;
        const double Fp_re       = (A_val * C_val + B_val * D_val) / denom2;
        const double Fp_im       = (B_val * C_val - A_val * D_val) / denom2;
        dQp                      = Fp_re * da - Fp_im * dy;
        const double dImF        = Fp_im * da + Fp_re * dy;
        dq_im                    = (y * dImF - ImF * dy) / (y * y);
      }
      const double inv  = inv_x2y2;
      const double dinv = -inv * inv * (dx2 + dy2);
      l2  = (Pp - Qp) * inv;
      dl2 = (dPp - dQp) * inv + (Pp - Qp) * dinv;
      dl0 = dPp - dl2 * x2 - l2 * dx2;
      l3  = (Pm_div_x - q_im) * inv;
      dl3 = (dPm_div_x - dq_im) * inv + (Pm_div_x - q_im) * dinv;
      l1  = Pm_div_x - l3 * x2;
      dl1 = dPm_div_x - dl3 * x2 - l3 * dx2;
    }
    double s[16], ds[16], s2[16], ds2[16], s3[16], ds3[16], t1[16], t2[16];
    S_mat(s);
    ds[0] = 0;   ds[1] = db;   ds[2] = dc;   ds[3] = dd;
    ds[4] = db;  ds[5] = 0;    ds[6] = du;   ds[7] = dv;
    ds[8] = dc;  ds[9] = -du;  ds[10] = 0;   ds[11] = dw;
    ds[12] = dd; ds[13] = -dv; ds[14] = -dw; ds[15] = 0;
    mat_mul(s, s, s2);
    mat_mul(ds, s, t1);
    mat_mul(s, ds, t2);
#pragma unroll
    for (int i = 0; i < 16; i++) ds2[i] = t1[i] + t2[i];
    mat_mul(s, s2, s3);
    mat_mul(ds, s2, t1);
    mat_mul(s, ds2, t2);
#pragma unroll
    for (int i = 0; i < 16; i++) ds3[i] = t1[i] + t2[i];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = s[i] * dl1 + ds[i] * l1 + s2[i] * dl2 + ds2[i] * l2 + s3[i] * dl3 + ds3[i] * l3;
    m[0] += dl0; m[5] += dl0; m[10] += dl0; m[15] += dl0;
  }

  // Lambda * (j,0,0,0)^T: first column of Lambda times j — all the LTE recursion needs,
  // because the LTE source vector has only an I component (rtepack_source.cc:88-95).
  __device__ __forceinline__ void L_col0(double j, double* __restrict__ o) const {
    double l0, l1, l2, l3;
    linsrc_coeffs(l0, l1, l2, l3);
    // v1 = S e0, v2 = S v1, v3 = S v2
    const double p1 = b, q1 = c, r1 = d;  // v1 = (0, b, c, d)
    const double v2_0 = b * p1 + c * q1 + d * r1;
    const double v2_1 = u * q1 + v * r1;
    const double v2_2 = -u * p1 + w * r1;
    const double v2_3 = -v * p1 - w * q1;
    const double v3_0 = b * v2_1 + c * v2_2 + d * v2_3;
    const double v3_1 = b * v2_0 + u * v2_2 + v * v2_3;
    const double v3_2 = c * v2_0 - u * v2_1 + w * v2_3;
    const double v3_3 = d * v2_0 - v * v2_1 - w * v2_2;
    o[0] = (l0 + l2 * v2_0 + l3 * v3_0) * j;
    o[1] = (l1 * p1 + l2 * v2_1 + l3 * v3_1) * j;
    o[2] = (l1 * q1 + l2 * v2_2 + l3 * v3_2) * j;
    o[3] = (l1 * r1 + l2 * v2_3 + l3 * v3_3) * j;
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// rte_option = linprop for POLARISED layers (tran::linsrc_linprop :467-474): Lambda = Re[alpha^-1 (D(u1) - T D(u0))] / r with
// alpha = sqrt((k2 - k1) / 2r) the complex square root of the absorption gradient (specmat sqrt(const propmat&), :872-1002),
// u = alpha^-1 k / 2 and D the ELEMENT-WISE Dawson function of a complex 4x4 matrix (rtepack_spectral_matrix.cc:6-26, sic).
// Rare and heavy (up to 32 complex Dawson evaluations through the register-resident w(z)); the functions are __noinline__ and
// keep their matrices in local memory so that the kernels' common paths keep their register budget.
// ---------------------------------------------------------------------------------------------------------------------
struct cx {
  double r, i;
};
__device__ __forceinline__ cx cmk(double r, double i = 0.0) { return cx{r, i}; }
__device__ __forceinline__ cx operator+(cx a, cx b) { return cx{a.r + b.r, a.i + b.i}; }
__device__ __forceinline__ cx operator-(cx a, cx b) { return cx{a.r - b.r, a.i - b.i}; }
__device__ __forceinline__ cx operator-(cx a) { return cx{-a.r, -a.i}; }
__device__ __forceinline__ cx operator*(cx a, cx b) { return cx{a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
__device__ __forceinline__ cx operator*(cx a, double b) { return cx{a.r * b, a.i * b}; }
__device__ __forceinline__ cx operator*(double b, cx a) { return cx{a.r * b, a.i * b}; }
__device__ __forceinline__ cx operator/(cx a, double b) { return cx{a.r / b, a.i / b}; }
__device__ __forceinline__ cx operator/(cx a, cx b) {  // Smith's algorithm (what the compiler's complex division does)
  if (fabs(b.r) >= fabs(b.i)) {
    const double q = b.i / b.r, d = b.r + b.i * q;
    return cx{(a.r + a.i * q) / d, (a.i - a.r * q) / d};
  }
  const double q = b.r / b.i, d = b.r * q + b.i;
  return cx{(a.r * q + a.i) / d, (a.i * q - a.r) / d};
}
// principal square root; the sign of a zero imaginary part picks the side of the cut like csqrt
__device__ __forceinline__ cx csqrt_(cx z) {
  if (z.r == 0.0 && z.i == 0.0) return cx{0.0, z.i};
  const double t = sqrt(0.5 * (fabs(z.r) + hypot(z.r, z.i)));
  if (z.r >= 0.0) return cx{t, z.i / (2.0 * t)};
  return cx{fabs(z.i) / (2.0 * t), copysign(t, z.i)};
}

// Faddeeva::Dawson(complex) (3rdparty/Faddeeva/Faddeeva.cc:461-570): D(z) = i sqrt(pi)/2 (exp(-z^2) - w(z)) with the package's
// series where that difference cancels (|z| small; small |y| and |xy| next to the real axis)
__device__ inline cx dawson_c(cx z) {
  constexpr double spi2 = 0.886226925452758013649083741671;
  const double x = z.r, y = z.i;
  if (y == 0.0) return cx{dawson(x), 0.0};
  if (x == 0.0) {
    const double y2 = y * y;
    if (y2 < 2.5e-5) return cx{x, y * (1. + y2 * (0.6666666666666666666666666666666666666667 + y2 * 0.26666666666666666666666666666666666667))};
    return cx{x, spi2 * (y >= 0 ? exp(y2) - erfcx(y) : erfcx(-y) - exp(y2))};
  }
  const double mRe = (y - x) * (x + y), mIm = -2 * x * y;  // -z^2
  if (fabs(y) < 5e-3) {
    if (fabs(x) < 5e-3) {  // dawson(z) = z - 2/3 z^3 + 4/15 z^5
      const cx mz2{mRe, mIm};
      const cx p = cmk(1.0) + mz2 * (cmk(0.6666666666666666666666666666666666666667) + mz2 * 0.2666666666666666666666666666666666666667);
      return z * p;
    }
    if (fabs(mIm) < 5e-3) {  // expansion in y about the real axis, :541-569
      const double x2 = x * x, y2 = y * y;
      if (x2 > 1600) {
        if (x2 > 25e14) {
          const double xy2 = (x * y) * (x * y);
          return cx{(0.5 + y2 * (0.5 + 0.25 * y2 - 0.16666666666666666667 * xy2)) / x,
                    y * (-1 + y2 * (-0.66666666666666666667 + 0.13333333333333333333 * xy2 - 0.26666666666666666667 * y2)) / (2 * x2 - 1)};
        }
        const double q = 1. / (-15 + x2 * (90 + x2 * (-60 + 8 * x2)));
        return cx{q * (x * (33 + x2 * (-28 + 4 * x2) + y2 * (18 - 4 * x2 + 4 * y2))),
                  q * (y * (-15 + x2 * (24 - 4 * x2) + y2 * (4 * x2 - 10 - 4 * y2)))};
      }
      const double D = dawson(x);
      return cx{D + y2 * (D + x - 2 * D * x2) +
                    y2 * y2 * (D * (0.5 - x2 * (2 - 0.66666666666666666667 * x2)) + x * (0.83333333333333333333 - 0.33333333333333333333 * x2)),
                y * (1 - 2 * D * x + y2 * 0.66666666666666666667 * (1 - x2 - D * x * (3 - 2 * x2)) +
                     y2 * y2 * (0.26666666666666666667 - x2 * (0.6 - 0.13333333333333333333 * x2) -
                                D * x * (1 - x2 * (1.3333333333333333333 - 0.26666666666666666667 * x2))))};
    }
  }
  double s, c, wr, wi;
  sincos(mIm, &s, &c);
  const double e = exp(mRe);
  cx res;
  if (y >= 0) {
    faddeeva_w(x, y, wr, wi);
    res = cx{e * c - wr, e * s - wi};
  } else {
    faddeeva_w(-x, -y, wr, wi);
    res = cx{wr - e * c, wi - e * s};
  }
  return cx{-spi2 * res.i, spi2 * res.r};
}

// specmat sqrt(const propmat&), :872-1002: Cayley-Hamilton coefficients d0..d3 of the principal square root, then the matrix
__device__ inline void sqrt_propmat(const Propmat& pm, cx* __restrict__ K) {
  constexpr double eps = 2.220446049250313e-16;
  const double a  = pm.A;
  const cx sqrt_a = csqrt_(cmk(a));
  for (int i = 0; i < 16; i++) K[i] = cmk(0.0);
  if (!pm.is_polarized()) {
    K[0] = K[5] = K[10] = K[15] = sqrt_a;
    return;
  }
  const double b = pm.B, c = pm.C, d = pm.D, u = pm.U, v = pm.V, w = pm.W;
  const double b2 = b * b, c2 = c * c, d2 = d * d, u2 = u * u, v2 = v * v, w2 = w * w;
  cx d0c = cmk(0.0), d1c = cmk(0.0), d2c = cmk(0.0), d3c = cmk(0.0);
  if (pm.is_rotational()) {
    const double rho = norm3d(u, v, w);
    if (rho <= eps) return;  // {0.0}
    const double r = sqrt(2.0 * rho);
    d1c = cmk(1.0 / r);
    d2c = cmk(-1.0 / (rho * r));
  } else {
    const double B = u2 + v2 + w2 - b2 - c2 - d2;
    const double t = d * u - c * v + b * w;
    const double C = -(t * t);
    const double S = sqrt(B * B - 4 * C);
    const double x2 = fmax(0.0, 0.5 * (S - B)), abs_y2 = fmax(0.0, 0.5 * (S + B));
    const double x = sqrt(x2), ys = sqrt(abs_y2);
    const cx sx = csqrt_(cmk(a + x)), dx = csqrt_(cmk(a - x));
    const cx sy = csqrt_(cx{a, ys}), dy = csqrt_(cx{a, 0.0 - ys});
    const cx Sx = sx + dx, Dx = sx - dx, Sy = sy + dy, Dy = sy - dy;
    if (x2 + abs_y2 <= eps) {
      d0c = sqrt_a;
      if (a <= eps) {  // sic, :934
        d1c = cmk(0.5) / sqrt_a;
        d2c = cmk(0.125) / (sqrt_a * a);
        d3c = cmk(0.0625) / (sqrt_a * (a * a));
      }
    } else {
      const double inv_sum_sq = 1.0 / (x2 + abs_y2);
      d0c = (abs_y2 * Sx + x2 * Sy) * (0.5 * inv_sum_sq);
      d2c = (Sx - Sy) * (0.5 * inv_sum_sq);
      const cx term1 = (x <= eps && a <= eps) ? cmk(0.0) : (x <= eps) ? cmk(0.5) / sqrt_a : (0.5 * Dx) / x;
      const cx term2 = (abs_y2 <= eps && a <= eps) ? cmk(0.0) : (abs_y2 <= eps) ? cmk(0.5) / sqrt_a : (0.5 * Dy) / cx{0.0, ys};
      d1c = (abs_y2 * term1 + x2 * term2) * inv_sum_sq;
      d3c = (term1 - term2) * inv_sum_sq;
    }
  }
  const double k2_00 = b2 + c2 + d2, k2_11 = b2 - u2 - v2, k2_22 = c2 - u2 - w2, k2_33 = d2 - v2 - w2;
  const double k2_01 = -(c * u + d * v), k2_02 = b * u - d * w, k2_03 = b * v + c * w;
  const double k2_12 = b * c - v * w, k2_13 = b * d + u * w, k2_23 = c * d - u * v;
  const double k3_01 = b * k2_00 - u * k2_02 - v * k2_03;
  const double k3_02 = c * k2_00 + u * k2_01 - w * k2_03;
  const double k3_03 = d * k2_00 + v * k2_01 + w * k2_02;
  const double k3_12 = -c * k2_01 + u * k2_11 - w * k2_13;
  const double k3_13 = -d * k2_01 + v * k2_11 + w * k2_12;
  const double k3_23 = -d * k2_02 + v * k2_12 + w * k2_22;
  K[0]  = d0c + d2c * k2_00;
  K[5]  = d0c + d2c * k2_11;
  K[10] = d0c + d2c * k2_22;
  K[15] = d0c + d2c * k2_33;
  K[1]  = d1c * b + d2c * k2_01 + d3c * k3_01;
  K[4]  = d1c * b - d2c * k2_01 + d3c * k3_01;
  K[2]  = d1c * c + d2c * k2_02 + d3c * k3_02;
  K[8]  = d1c * c - d2c * k2_02 + d3c * k3_02;
  K[3]  = d1c * d + d2c * k2_03 + d3c * k3_03;
  K[12] = d1c * d - d2c * k2_03 + d3c * k3_03;
  K[6]  = d1c * u + d2c * k2_12 + d3c * k3_12;
  K[9]  = d2c * k2_12 - d1c * u - d3c * k3_12;
  K[7]  = d1c * v + d2c * k2_13 + d3c * k3_13;
  K[13] = d2c * k2_13 - d1c * v - d3c * k3_13;
  K[11] = d1c * w + d2c * k2_23 + d3c * k3_23;
  K[14] = d2c * k2_23 - d1c * w - d3c * k3_23;
}

// inverse of a complex 4x4 matrix, adj(A) / det(A) (rtepack_spectral_matrix.h:203-242), through the twelve 2x2 minors of
// the Laplace expansion
__device__ inline void spec_inv(const cx* __restrict__ a, cx* __restrict__ o) {
  const cx s0 = a[0] * a[5] - a[4] * a[1], s1 = a[0] * a[6] - a[4] * a[2], s2 = a[0] * a[7] - a[4] * a[3];
  const cx s3 = a[1] * a[6] - a[5] * a[2], s4 = a[1] * a[7] - a[5] * a[3], s5 = a[2] * a[7] - a[6] * a[3];
  const cx c5 = a[10] * a[15] - a[14] * a[11], c4 = a[9] * a[15] - a[13] * a[11], c3 = a[9] * a[14] - a[13] * a[10];
  const cx c2 = a[8] * a[15] - a[12] * a[11], c1 = a[8] * a[14] - a[12] * a[10], c0 = a[8] * a[13] - a[12] * a[9];
  const cx det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
  const cx id  = cmk(1.0) / det;
  o[0]  = (a[5] * c5 - a[6] * c4 + a[7] * c3) * id;
  o[1]  = (a[2] * c4 - a[1] * c5 - a[3] * c3) * id;
  o[2]  = (a[13] * s5 - a[14] * s4 + a[15] * s3) * id;
  o[3]  = (a[10] * s4 - a[9] * s5 - a[11] * s3) * id;
  o[4]  = (a[6] * c2 - a[4] * c5 - a[7] * c1) * id;
  o[5]  = (a[0] * c5 - a[2] * c2 + a[3] * c1) * id;
  o[6]  = (a[14] * s2 - a[12] * s5 - a[15] * s1) * id;
  o[7]  = (a[8] * s5 - a[10] * s2 + a[11] * s1) * id;
  o[8]  = (a[4] * c4 - a[5] * c2 + a[7] * c0) * id;
  o[9]  = (a[1] * c2 - a[0] * c4 - a[3] * c0) * id;
  o[10] = (a[12] * s4 - a[13] * s2 + a[15] * s0) * id;
  o[11] = (a[9] * s2 - a[8] * s4 - a[11] * s0) * id;
  o[12] = (a[5] * c1 - a[4] * c3 - a[6] * c0) * id;
  o[13] = (a[0] * c3 - a[1] * c1 + a[2] * c0) * id;
  o[14] = (a[13] * s1 - a[12] * s3 - a[14] * s0) * id;
  o[15] = (a[8] * s3 - a[9] * s1 + a[10] * s0) * id;
}

// Lambda of a polarised layer with an absorption gradient >= 1e-8 (:467-474).  Tm = T of the layer; ncol = 4: the whole matrix
// L [16]; ncol = 1: only its first column L[0], L[4], L[8], L[12] (all the LTE recursion uses), which needs the first columns
// of u0, u1 and of the Dawson matrices only: 8 instead of 32 complex Dawson evaluations.
static __device__ __noinline__ void linprop_lambda_pol(const double* __restrict__ Tm, const Propmat& k1, const Propmat& k2, double r, int ncol,
                                                double* __restrict__ L) {
  const double dn = 2.0 * r;
  const Propmat a2{(k2.A - k1.A) / dn, (k2.B - k1.B) / dn, (k2.C - k1.C) / dn, (k2.D - k1.D) / dn,
                   (k2.U - k1.U) / dn, (k2.V - k1.V) / dn, (k2.W - k1.W) / dn};
  cx al[16], ai[16];
  sqrt_propmat(a2, al);
  spec_inv(al, ai);
  // M = dawson(u1) - T dawson(u0), u = alpha^-1 (k / 2) (specmat x propmat, rtepack_multitype.h:145-168)
  cx M[16];
  for (int pass = 0; pass < 2; pass++) {
    const Propmat& k = pass ? k2 : k1;
    const double a = k.A / 2.0, b = k.B / 2.0, c = k.C / 2.0, d = k.D / 2.0, u = k.U / 2.0, v = k.V / 2.0, w = k.W / 2.0;
    cx D[16];
    for (int i = 0; i < 4; i++) {
      const cx m1 = ai[4 * i], m2 = ai[4 * i + 1], m3 = ai[4 * i + 2], m4 = ai[4 * i + 3];
      D[4 * i] = dawson_c(a * m1 + b * m2 + c * m3 + d * m4);
      if (ncol > 1) {
        D[4 * i + 1] = dawson_c(a * m2 + b * m1 - m3 * u - m4 * v);
        D[4 * i + 2] = dawson_c(a * m3 + c * m1 + m2 * u - m4 * w);
        D[4 * i + 3] = dawson_c(a * m4 + d * m1 + m2 * v + m3 * w);
      }
    }
    if (pass == 0) {
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < ncol; j++)
          M[4 * i + j] = -(Tm[4 * i + 0] * D[j] + Tm[4 * i + 1] * D[4 + j] + Tm[4 * i + 2] * D[8 + j] + Tm[4 * i + 3] * D[12 + j]);
    } else {
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < ncol; j++) M[4 * i + j] = D[4 * i + j] + M[4 * i + j];
    }
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < ncol; j++) {
      double acc = 0.0;
      for (int k = 0; k < 4; k++) acc += ai[4 * i + k].r * M[4 * k + j].r - ai[4 * i + k].i * M[4 * k + j].i;
      L[4 * i + j] = acc / r;
    }
}

// tran::linsrc_linprop (:449-475), whichever branch the layer takes; t is the layer's tran state, Tm its T
static __device__ __noinline__ void linprop_lambda_any(const Tran& t, const double* __restrict__ Tm, const Propmat& k1, const Propmat& k2, double r,
                                                double* __restrict__ L) {
  const double alpha2 = (k2.A - k1.A) / (2.0 * r);
  if (alpha2 < 1e-8 || !t.polarized) {
    if (!t.polarized) {
      const double l = alpha2 < 1e-8 ? func_F(t.a) : linprop_lambda(k1.A, k2.A, r, Tm[0]);
      for (int i = 0; i < 16; i++) L[i] = 0.0;
      L[0] = L[5] = L[10] = L[15] = l;
    } else {
      t.L(L);
    }
    return;
  }
  linprop_lambda_pol(Tm, k1, k2, r, 4, L);
}

// tran::linsrc_linprop_deriv for a polarised layer with a gradient (:543-555): forward perturbation of 1e-6 (dk, dr) of the
// level the target belongs to ("These derivaties don't work so we use perturbations..."); lambda = the layer's Lambda
static __device__ __noinline__ void linprop_lambda_pol_deriv(const double* __restrict__ lambda, const Propmat& k1, const Propmat& k2, const Propmat& dk,
                                                      double r, double dr, bool k1_deriv, double* __restrict__ dL) {
  constexpr double eps = 1e-6;
  Propmat kp = k1_deriv ? k1 : k2;
  kp.A += dk.A * eps; kp.B += dk.B * eps; kp.C += dk.C * eps; kp.D += dk.D * eps; kp.U += dk.U * eps; kp.V += dk.V * eps; kp.W += dk.W * eps;
  const Propmat& q1 = k1_deriv ? kp : k1;
  const Propmat& q2 = k1_deriv ? k2 : kp;
  const double rp = r + dr * eps;
  Tran tp;
  tp.init(q1, q2, rp, false);
  double Tp[16], Lp[16];
  if (tp.polarized) {
    tp.T(Tp);
  } else {
    for (int i = 0; i < 16; i++) Tp[i] = 0.0;
    Tp[0] = Tp[5] = Tp[10] = Tp[15] = tp.exp_a;
  }
  linprop_lambda_any(tp, Tp, q1, q2, rp, Lp);
  for (int i = 0; i < 16; i++) dL[i] = (Lp[i] - lambda[i]) / eps;
}

}  // namespace rte
}  // namespace ab200
