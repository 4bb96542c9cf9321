// common.cuh — shared device/host helpers of the arts_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/arts_b200.h"

namespace ab200 {

// ---- constants: same expressions as src/core/util/arts_constants.h:57-254 ----
namespace cst {
constexpr double pi          = 3.14159265358979323846264338327950288;
constexpr double inv_sqrt_pi = 0.564189583547756286948079451560772586;
constexpr double c           = 299792458;
constexpr double h           = 6.62607015e-34;
constexpr double k           = 1.380649e-23;
constexpr double NA          = 6.02214076e23;
constexpr double e           = 1.602176634e-19;
constexpr double alpha       = 7.2973525693e-3;
constexpr double R_inf       = 10973731.568160;
constexpr double inv_two_pi  = (1.0 / pi) / 2;
constexpr double h_bar       = h * inv_two_pi;
constexpr double m_e         = 2 * h * R_inf / (c * (alpha * alpha));
constexpr double bohr_magneton = e * h_bar / (2 * m_e);
constexpr double R           = k * NA;
constexpr double doppler_broadening_const_squared = 2000 * R / (c * c);
}  // namespace cst

// ---- geometry of the line catalog on the device ---------------------------
#ifndef AB200_TL
#define AB200_TL 256
#endif
constexpr int TL        = AB200_TL;  // (sub-)lines per tile; tiles never straddle a segment
constexpr int REC_GROUP = 4;    // doubles per record group (one LDS.128 pair)
constexpr int N_GROUPS  = 4;    // groups per line record -> 16 doubles = 128 B per (level, line)
constexpr int REC_DOUBLES = REC_GROUP * N_GROUPS;
// record layout of one tile (for one level): [group][line][4]
//   group 0: f0', c3 = g^2 - h, kappa = 4 g^2 h, A1 = Si*g          (far wing, real part)
//   group 1: B1 = 2 h A1, igd, y, s_re                              (far wing real | near evaluation)
//   group 2: E1(y), s_im, cut_re, cut_im                            (near evaluation, line mixing, cutoff value;
//            in real merged segments (mode 0) s_im == 0 and the slot holds the line's cutoff [Hz] instead)
//   group 3: A2 = Sr, A3 = -Sr*g, B3 = 2 h A3, A4 = Si              (far wing, complex part)
// with g = G0 [Hz], h = GD^2/2, S = i*s*GD/sqrt(pi) = Sr + i Si, (igd, y, s) the reference's
// single_shape (lbl_lineshape_voigt_lte.h:20-33).  The real-only kernel streams groups 0-1 for
// far tiles and 0-2 for near tiles; the complex kernel groups 0-1 + 3 or all four.
constexpr size_t tile_doubles() { return size_t(TL) * REC_DOUBLES; }

// tile summary written by the prepare kernel: f0'min, f0'max, min igd, min y | min cutoff, max cutoff,
// sum of the cutoff values ls(f0' + cutoff) (real part), max igd — over the contributing lines of the tile
constexpr int SUMMARY_DOUBLES = 8;

// far-wing boundary of the reference's Faddeeva: x + |y| > 4000 -> nu <= 2 closed form
// (3rdparty/Faddeeva/Faddeeva.cc:707-725)
constexpr double FAR_LIMIT = 4000.0;
// The closed form i z / (sqrt(pi) (z^2 - 1/2)) the reference uses above 4000 differs from w(z) by
// 1/(2 z^4) relative (real part for y << x: 2.5/x^4), i.e. <= 2.5e-12 for |z| >= 1000 — measured 2.7e-12 at
// x = 1024 in tests/test_gpu_propmat.py::test_mid_wing_closed_form_accuracy.
// The real line sum (same-sign terms, no cancellation) therefore uses it from |x|+y > 1000: four times
// fewer pairs in the expensive continued-fraction branch, parity bound 1e-9 untouched.
constexpr double FAR_LIMIT_REAL_SUM = 1000.0;
// |x| + y above which the forward line sums evaluate a near pair with the four-term continued fraction in closed form
// (faddeeva.cuh: w_mid, <= 1.3e-12 relative on both parts, measured against scipy's wofz) instead of the reference's
// nu(z)-term recurrence.  The Jacobian kernels do not use it: their forward difference amplifies w's error by 1e4.
constexpr double MID_LIMIT = 48.0;

// Far-field (multipole) sums of the real, cutoff-free segments (lbl_fmm.cu): per cluster and level a record of
// MOM_DOUBLES doubles
constexpr int MP_P = 16;
constexpr double MP_THETA = 6.0;
constexpr int MOM_DOUBLES = 24;
// record slots: centre, acceptance distance, radius, m_1 .. m_16, then for lines with ByLine cutoffs the distance up to
// which every line of the cluster is inside its window, the distance beyond which every line is outside, and the sum of
// the lines' cutoff values ls(f0' + cutoff)
constexpr int MOM_C = 0, MOM_RHO = 1, MOM_R = 2, MOM_M1 = 3, MOM_IN = 19, MOM_OUT = 20, MOM_CUT = 21;
constexpr int FMM_GROUP = 16;  // tiles per coarsest cluster
constexpr int64_t FMM_MIN_LINES = 1024;  // real segments with fewer (sub-)lines keep the line-by-line kernel (break-even ~800)

constexpr int AB200_MAX_TARGETS = 8;  // Jacobian targets per call (temperature + species VMRs)

enum Pol : int { POL_NO = 0, POL_PI = 1, POL_SM = 2, POL_SP = 3 };

// ---- error plumbing ---------------------------------------------------------
int set_error(int code, const std::string& msg);
// internal hooks between api.cu and multi.cu
int path_create_ex(const ab200_catalog* cat, int64_t nf, int32_t np, int32_t nq, bool stage2_only, ab200_path** out);
cudaStream_t path_stream(ab200_path* p);
double* path_K(ab200_path* p, int64_t* k_pitch);
int path_adopt_K(ab200_path* p);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int n = 1);

#define AB_CUDA(expr)                                                              \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) return ::ab200::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define AB_TRY(expr)        \
  do {                      \
    int _rc = (expr);       \
    if (_rc) return _rc;    \
  } while (0)

// ---- small device helpers -----------------------------------------------------
#ifdef __CUDACC__
// 1/d to ~1 ulp: MUFU.RCP64H seed (2^-20) + one cubic Newton step (error^3 = 2^-60).
// Valid for normal, finite, non-zero d (the callers' denominators are sums of squares > 0).
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  const double e = __fma_rn(-d, r, 1.0);
  const double t = __fma_rn(e, e, e);
  return __fma_rn(r, t, r);
}

// Minimum / maximum of a double over the lanes of `mask` with the integer warp reduction (REDUX): the bits of a double, with the
// sign bit flipped (all bits for a negative value), order like the value, so the 64-bit minimum is two 32-bit reductions - the high
// words, then the low words of the lanes that hold the minimal high word.  ~14 instructions instead of the ~70 of a five-stage
// butterfly of fmin over 64-bit shuffles; the result is the same double (a NaN orders above +inf, -0 below +0).  Every lane of
// `mask` must call with the same mask; the two half-warps may call together with their own halves.
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
  const long long b = __double_as_longlong(v);
  return static_cast<unsigned long long>(b ^ ((b >> 63) | static_cast<long long>(0x8000000000000000ull)));
}
__device__ __forceinline__ double ordered_value(unsigned hi, unsigned lo) {
  const unsigned long long k = (static_cast<unsigned long long>(hi) << 32) | lo;
  return __longlong_as_double(static_cast<long long>(k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull)));
}
__device__ __forceinline__ double lanes_min(double v, unsigned mask = 0xffffffffu) {
  const unsigned long long k = ordered_bits(v);
  const unsigned hi = static_cast<unsigned>(k >> 32), lo = static_cast<unsigned>(k);
  const unsigned mhi = __reduce_min_sync(mask, hi);
  return ordered_value(mhi, __reduce_min_sync(mask, hi == mhi ? lo : 0xffffffffu));
}
__device__ __forceinline__ double lanes_max(double v, unsigned mask = 0xffffffffu) {
  const unsigned long long k = ordered_bits(v);
  const unsigned hi = static_cast<unsigned>(k >> 32), lo = static_cast<unsigned>(k);
  const unsigned mhi = __reduce_max_sync(mask, hi);
  return ordered_value(mhi, __reduce_max_sync(mask, hi == mhi ? lo : 0u));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) ----------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// orders prior generic-proxy accesses to shared memory before subsequent async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
#endif

}  // namespace ab200
