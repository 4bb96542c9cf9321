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

}  // namespace rte
}  // namespace ab200
