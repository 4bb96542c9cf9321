// Observer epilogue on the device: the host glue the reference runs around the path.
//
//   background_planck_kernel   from_temp, src/m_background.cc:55-63 (spectral_radSurfaceBlackbody :113-141,
//                              spectral_radUniformCosmicBackground :65-72)
//   unit_transform_kernel      spectral_rad_transform_operator, spectral_radiance_transform_operator.cc:8-122,
//                              on spectral_rad [nf] and every row of spectral_rad_jac [nx][nf]
//   sensor_sumup_kernel        SensorObsel::sumup, src/core/sensor/obsel.cpp:246-279
//
// The x-space accumulation of spectral_rad_jacAddPathPropagation / spectral_rad_jacFromBackground (src/m_rad.cc:26-127)
// lives inside the fused Jacobian pass (stokes_jac.cu).  All three kernels are HBM bound: the transform reads and writes
// 32 (nx + 1) B per frequency, the sum-up reads 32 B per (entry, row).
#include "common.cuh"
#include "rtepack.cuh"
#include "stokes.hpp"

namespace ab200 {
using namespace rte;

__global__ void background_planck_kernel(int64_t nf, const double* __restrict__ f, double T, double* __restrict__ I_bkg) {
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= nf) return;
  double2* o = reinterpret_cast<double2*>(I_bkg + 4 * j);
  o[0] = make_double2(planck(f[j], T), 0.0);
  o[1] = make_double2(0.0, 0.0);
}

// dinvplanckdI, physics_funcs.cc:76-83
__device__ __forceinline__ double dinvplanckdI(double i, double f) {
  constexpr double a = cst::h / cst::k;
  constexpr double b = 2 * cst::h / (cst::c * cst::c);
  const double d     = b * f * f * f / i;
  const double binv  = a * f / log1p(d);
  return binv * binv / (a * f * i * (1 / d + 1));
}

__global__ void __launch_bounds__(128) unit_transform_kernel(int64_t nf, int32_t nx, const double* __restrict__ f,
                                                             int32_t unit, double n_real, double* __restrict__ I,
                                                             double* __restrict__ Jx) {
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= nf) return;
  double2* iv = reinterpret_cast<double2*>(I + 4 * j);
  const double2 a = iv[0], b = iv[1];
  double v[4] = {a.x, a.y, b.x, b.y};
  double dv[4] = {1.0, 1.0, 1.0, 1.0};
  const double n2 = n_real * n_real;
  const double fj = f[j];
  bool scale = true;
  switch (unit) {
    case AB200_UNIT_UNIT:  // :8-19
      scale = n_real != 1.0;
      if (scale) {
#pragma unroll
        for (int c = 0; c < 4; c++) { v[c] *= n2; dv[c] = n2; }
      }
      break;
    case AB200_UNIT_RJBT: {  // :21-44, invrayjean(1, f) physics_funcs.cc:172-176
      constexpr double k = cst::c * cst::c / (2 * cst::k);
      const double df = (k * 1.0) / (fj * fj);
#pragma unroll
      for (int c = 0; c < 4; c++) { v[c] *= df; dv[c] = df; }
    } break;
    case AB200_UNIT_PLANCKBT: {  // :46-87
      dv[0] = dinvplanckdI(v[0], fj);
      double n[4];
      n[0] = invplanck(v[0], fj);
#pragma unroll
      for (int c = 1; c < 4; c++) {
        dv[c] = dinvplanckdI(0.5 * (v[0] + v[c]), fj) - dinvplanckdI(0.5 * (v[0] - v[c]), fj);
        n[c]  = invplanck(0.5 * (v[0] + v[c]), fj) - invplanck(0.5 * (v[0] - v[c]), fj);
      }
#pragma unroll
      for (int c = 0; c < 4; c++) v[c] = n[c];
    } break;
    case AB200_UNIT_W_M2_M_SR: {  // :89-112
      const double df = (fj * (fj / cst::c));
#pragma unroll
      for (int c = 0; c < 4; c++) { v[c] *= df * n2; dv[c] = df * n2; }
    } break;
    default: {  // AB200_UNIT_W_M2_M1_SR :114-122
#pragma unroll
      for (int c = 0; c < 4; c++) { v[c] *= n2 * cst::c; dv[c] = n2 * cst::c; }
    } break;
  }
  iv[0] = make_double2(v[0], v[1]);
  iv[1] = make_double2(v[2], v[3]);
  if (!scale || Jx == nullptr) return;
  for (int32_t i = 0; i < nx; i++) {
    double2* x = reinterpret_cast<double2*>(Jx + (int64_t(i) * nf + j) * 4);
    double2 p = x[0], q = x[1];
    p.x *= dv[0]; p.y *= dv[1]; q.x *= dv[2]; q.y *= dv[3];
    x[0] = p; x[1] = q;
  }
}

// One warp per (channel, row): row nx is spectral_rad itself (y), rows 0..nx-1 the Jacobian (Jy).  Lanes stride over
// the channel's sparse entries; the butterfly reduction is a fixed tree, so the result is deterministic.
__global__ void __launch_bounds__(128) sensor_sumup_kernel(int64_t nf, int32_t nx, int32_t n_channels,
                                                           const int64_t* __restrict__ w_offset,
                                                           const int64_t* __restrict__ w_freq,
                                                           const double* __restrict__ w_stokes,
                                                           const double* __restrict__ I, const double* __restrict__ Jx,
                                                           double* __restrict__ y, double* __restrict__ Jy) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t rows = int64_t(nx) + 1;
  if (warp >= int64_t(n_channels) * rows) return;
  const int ch      = int(warp / rows);
  const int64_t row = warp % rows;
  const double* src = row == nx ? I : Jx + row * nf * 4;
  double sum = 0.0;
  for (int64_t e = w_offset[ch] + lane; e < w_offset[ch + 1]; e += 32) {
    const double2* a = reinterpret_cast<const double2*>(src + 4 * w_freq[e]);
    const double2* w = reinterpret_cast<const double2*>(w_stokes + 4 * e);
    const double2 a0 = a[0], a1 = a[1], w0 = w[0], w1 = w[1];
    sum += a0.x * w0.x + a0.y * w0.y + a1.x * w1.x + a1.y * w1.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) {
    if (row == nx) y[ch] = sum;
    else Jy[int64_t(ch) * nx + row] = sum;
  }
}

int launch_background_planck(int64_t nf, const double* f, double T, double* I_bkg, cudaStream_t stream) {
  if (nf == 0) return 0;
  background_planck_kernel<<<static_cast<unsigned>((nf + 255) / 256), 256, 0, stream>>>(nf, f, T, I_bkg);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

int launch_unit_transform(int64_t nf, int32_t nx, const double* f, int32_t unit, double n_real, double* I, double* Jx,
                          cudaStream_t stream) {
  if (nf == 0) return 0;
  unit_transform_kernel<<<static_cast<unsigned>((nf + 127) / 128), 128, 0, stream>>>(nf, nx, f, unit, n_real, I, Jx);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

int launch_sensor_sumup(int64_t nf, int32_t nx, int32_t n_channels, const int64_t* w_offset, const int64_t* w_freq,
                        const double* w_stokes, const double* I, const double* Jx, double* y, double* Jy,
                        cudaStream_t stream) {
  if (n_channels == 0) return 0;
  const int64_t warps = int64_t(n_channels) * (int64_t(Jx ? nx : 0) + 1);
  // without a Jacobian only the y rows exist: the kernel indexes rows as nx + 1 per channel, so pass nx = 0 then
  sensor_sumup_kernel<<<static_cast<unsigned>((warps * 32 + 127) / 128), 128, 0, stream>>>(
      nf, Jx ? nx : 0, n_channels, w_offset, w_freq, w_stokes, I, Jx, y, Jy);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ab200
