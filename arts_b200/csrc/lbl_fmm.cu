// lbl_fmm.cu — the line sum of real segments (with or without ByLine cutoffs) as a hierarchical far-field (multipole) sum.
//
// Replaces, for merged real segments, the line-by-line K2 loop of lbl_sum_real_kernel (reference band_shape::operator(),
// src/core/lbl/lbl_lineshape_voigt_lte.cpp:431-436 and with cutoff :591-608, per pair s Faddeeva::w(z) :239).
//
// Beyond |x| + y = 48 the forward kernels already evaluate w(z) with four terms of its continued fraction collapsed to
// one rational function (faddeeva.cuh w_mid: w = (i/sqrt(pi)) z (t - 5/2) / (t^2 - 3 t + 3/4), t = z^2, 1.3e-12).  Its
// partial fractions are four simple poles on the real z axis - the 4-point Gauss-Hermite rule,
//     w(z) ~ (i/sqrt(pi)) sum_j w_j [1 / (z - r_j) + 1 / (z + r_j)],  r = 0.52465, 1.65068,  w = 0.454124, 0.045876,
// so in line space (zeta = u + i g, u = f - f0', z = zeta / GD) a line is S sum_j w_j [1/(zeta - a_j) + 1/(zeta + a_j)],
// S = i s GD / sqrt(pi), a_j = GD r_j.  About a centre c (v = f - c, delta = f0' - c) every pole is 1 / (v - p),
// p = delta +- a_j - i g, and a CLUSTER of lines collapses to
//     sum_l Re(s_l w(z_l)) = (1 / v) sum_{k>=1} m_k (R / v)^k,   m_k = -sum_l Si_l sum_j w_j Im[(p_lj+ / R)^k + (p_lj- / R)^k],
// valid for |v| > R = max |p|.  MP_P = 16 moments and |v| >= MP_THETA R = 6 R reproduce the pair-by-pair sum of the same
// rational function to 3e-13 (measured against 40-digit arithmetic for widths from 1e2 to 3e9 Hz); a cluster then costs
// MP_P + 8 FP64 instructions per frequency whatever its size.  Im(p^k) is carried by (A, B) <- (d A + g B, d B - g A),
// whose terms have equal signs: relative accuracy does not depend on g / |delta| (Doppler regime).
//
// Clusters: 16 lines, 64 lines, a tile (256), 16 tiles.  A cluster X is ACCEPTED by a frequency when |f - c_X| > rho_X,
//     rho_X = max(6 R_X, every line of X has |x| + y > 48 (1 + 1e-9) beyond it, rho_child + |c_child - c_X| for its children),
// so acceptance is nested (a frequency that accepts X accepts every cluster inside X) and depends on the frequency
// alone.  A line contributes through its coarsest accepted cluster (lbl_fmm_far_kernel); pairs whose 16-line cluster is
// not accepted are evaluated one by one with the per-pair arithmetic of lbl_sum_real_kernel (lbl_fmm_near_kernel).
// The split never depends on block, shard or GPU boundaries: spectra stay bit-identical under any frequency partition.
// ByLine cutoffs: a line contributes ls(f) - ls(f0' + cutoff) inside its window and nothing outside.  A cluster's
// expansion (minus the sum of its cutoff values) then serves the frequencies between rho and the distance IN up to which
// every line is inside its window; beyond OUT every line is outside and the cluster contributes 0; in between the
// window edges cut through the cluster and its lines are taken pair by pair with the per-pair window test.  IN and OUT are
// nested like rho.
//
// configs[3] (1e6 lines): per (frequency, level) ~700 cluster visits + ~100-400 pairs instead of 1e6 pairs.
#include <cfloat>

#include "catalog.hpp"
#include "faddeeva.cuh"
#include "lbl.hpp"

namespace ab200 {

namespace {
constexpr double GH_R0 = 0.52464762327529031788, GH_R1 = 1.65068012388578455588;   // sqrt((3 -+ sqrt 6) / 2)
constexpr double GH_W0 = 0.45412414523193150818, GH_W1 = 0.04587585476806849182;   // (t_j - 5/2) / (2 (t_j - t_other))

// maximum / minimum over the warp (width 32) or over the lane's half-warp (width 16): common.cuh lanes_max / lanes_min
__device__ __forceinline__ unsigned width_mask(int width) {
  return width == 32 ? 0xffffffffu : ((threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu);
}
__device__ __forceinline__ double warp_max(double v, int width) { return lanes_max(v, width_mask(width)); }
__device__ __forceinline__ double warp_min(double v, int width) { return lanes_min(v, width_mask(width)); }

// moments of one line about centre c with scale 1 / R: term[k] = -Si sum_j w_j (B_{j+} + B_{j-}) at power k + 1
__device__ __forceinline__ void line_terms(bool live, double d, double GD, double g, double Si, double iR, double* term) {
  const double gh = g * iR;
  const double dd[4] = {(d + GH_R0 * GD) * iR, (d - GH_R0 * GD) * iR, (d + GH_R1 * GD) * iR, (d - GH_R1 * GD) * iR};
  double A[4] = {1.0, 1.0, 1.0, 1.0}, B[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < MP_P; k++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const double A1 = __fma_rn(dd[j], A[j], gh * B[j]), B1 = __fma_rn(dd[j], B[j], -(gh * A[j]));
      A[j] = A1; B[j] = B1;
    }
    term[k] = live ? -Si * (GH_W0 * (B[0] + B[1]) + GH_W1 * (B[2] + B[3])) : 0.0;
  }
}
__device__ __forceinline__ double line_radius(double d, double GD, double g) { return fabs(d) + GH_R1 * GD + g; }
// |f - c| beyond which the line's |x| + y exceeds the mid limit with the margin of the per-pair test
__device__ __forceinline__ double line_reach(double d, double GD, double y) {
  return fabs(d) + fmax(0.0, MID_LIMIT * (1.0 + 1e-9) - y) * GD;
}

// A frequency at distance v from a cluster's centre is SERVED by the cluster's record when the expansion applies (beyond rho
// and, with ByLine cutoffs, while every line is still inside its window) or when every line is outside its window
// (contribution 0).  `ann` tells the first case.
__device__ __forceinline__ bool fmm_served(double v, double rho, double in, double out, bool& ann) {
  ann = v > rho && v <= in;
  return ann || v > out;
}

// Moments of a parent cluster from those of a child (multipole-to-multipole shift).  With a_j = m_j R_c^j the coefficients
// of 1 / v^(j+1) about the child's centre, the same poles seen from the parent's centre sit at p + dc (dc = c_child -
// c_parent, real), so a'_k = sum_{j<=k} C(k, j) dc^(k-j) a_j: exact, no truncation (a'_k needs only a_1 .. a_k; a_0 = 0
// because the strengths are imaginary).  In scaled form every factor is <= 1:
//     m'_k = sum_{j=1..k} C(k, j) (dc / R_p)^(k-j) (R_c / R_p)^j m_j,     |dc| + R_c <= R_p.
__constant__ double BINOM[MP_P + 1][MP_P + 1];
__device__ __forceinline__ double m2m_term(int k, double x, double r, const double* __restrict__ m /* m[j-1] = m_j */) {
  // Horner in x over j = 1 .. k: sum_j [C(k, j) r^j m_j] x^(k-j) (no table of powers: k differs from thread to thread, and an
  // array indexed by it would live in local memory)
  double sum = 0.0, rp = 1.0;
#pragma unroll
  for (int j = 1; j <= MP_P; j++) {
    rp *= r;
    if (j <= k) sum = __fma_rn(sum, x, BINOM[k][j] * rp * m[j - 1]);
  }
  return sum;
}
}  // namespace

// ---------------------------------------------------------------------------
// moments of the 16-line, 64-line and tile clusters: one CTA per (tile, level), one thread per line
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TL, 4) lbl_fmm_moments_tile_kernel(PrepareParams p, FmmBuffers fb) {
  static_assert(TL == 256, "cluster sizes 16 / 64 / 256");
  const int64_t tile = blockIdx.x;
  const int lev      = blockIdx.y;
  const int lane     = threadIdx.x;
  const double* __restrict__ rec = p.prep + (int64_t(lev) * p.ntiles + tile) * tile_doubles();
  double* __restrict__ o2 = fb.L2 + (int64_t(lev) * p.ntiles + tile) * MOM_DOUBLES;
  double* __restrict__ o1 = fb.L1 + (int64_t(lev) * p.ntiles + tile) * 4 * MOM_DOUBLES;
  double* __restrict__ o0 = fb.L0 + (int64_t(lev) * p.ntiles + tile) * 16 * MOM_DOUBLES;
  if (p.tile_mode[tile] != 0) {  // complex tile: never served
    for (int i = lane; i < 21 * MOM_DOUBLES; i += TL) {
      double* o = i < MOM_DOUBLES ? o2 + i : i < 5 * MOM_DOUBLES ? o1 + (i - MOM_DOUBLES) : o0 + (i - 5 * MOM_DOUBLES);
      const int k = i % MOM_DOUBLES;
      *o = (k == MOM_RHO || k == MOM_OUT) ? DBL_MAX : k == MOM_IN ? -1.0 : 0.0;
    }
    return;
  }
  const double f0s = rec[(0 * TL + lane) * REC_GROUP + 0];
  const double igd = rec[(1 * TL + lane) * REC_GROUP + 1];
  const double y   = rec[(1 * TL + lane) * REC_GROUP + 2];
  const double sre = rec[(1 * TL + lane) * REC_GROUP + 3];
  const bool live  = igd != 0.0;
  const double GD  = live ? 1.0 / igd : 0.0;
  const double g   = y * GD, Si = sre * GD * cst::inv_sqrt_pi;
  // ByLine cutoff of the line (real merged segments keep it in the s_im slot) and its cutoff value ls(f0' + cutoff)
  const double cutl = live ? rec[(2 * TL + lane) * REC_GROUP + 1] : DBL_MAX;
  const bool has_cut = live && cutl < DBL_MAX;
  const double cval = has_cut ? rec[(2 * TL + lane) * REC_GROUP + 2] : 0.0;

  __shared__ double sh[TL / 32][8];
  __shared__ double c1s[4], rho1s[4], in1s[4], out1s[4], cs1s[4], c0s[16], rho0s[16], in0s[16], out0s[16], cs0s[16];
  __shared__ double c2s;
  __shared__ double m0s[16][MP_P], m1s[4][MP_P], R0s[16], R1s[4];
  const int warp = lane >> 5;
  const int w1 = warp & ~1;

  // --- centres: midpoint of the live lines' centres per cluster
  const double lo = live ? f0s : DBL_MAX, hi = live ? f0s : -DBL_MAX;
  const double lo0 = warp_min(lo, 16), hi0 = warp_max(hi, 16);
  const double lo32 = warp_min(lo0, 32), hi32 = warp_max(hi0, 32);
  if ((lane & 31) == 0) { sh[warp][0] = lo32; sh[warp][1] = hi32; }
  __syncthreads();
  const double lo1 = fmin(sh[w1][0], sh[w1 + 1][0]), hi1 = fmax(sh[w1][1], sh[w1 + 1][1]);
  // over the tile's eight warps: lane l takes warp l & 7's value, one integer warp reduction (instead of eight loads and fmin each)
  static_assert(TL / 32 == 8, "eight warps per tile");
  const double lo2 = lanes_min(sh[lane & 7][0]), hi2 = lanes_max(sh[lane & 7][1]);
  __syncthreads();
  const bool e0 = lo0 > hi0, e1 = lo1 > hi1, e2 = lo2 > hi2;  // empty clusters
  const double c2 = e2 ? 0.0 : 0.5 * (lo2 + hi2);
  const double c1 = e1 ? c2 : 0.5 * (lo1 + hi1);
  const double c0 = e0 ? c1 : 0.5 * (lo0 + hi0);
  const double d0 = live ? f0s - c0 : 0.0, d1 = live ? f0s - c1 : 0.0, d2 = live ? f0s - c2 : 0.0;

  // --- per cluster: radius R, reach D of the mid form, window bounds (every line inside its window up to IN, every line
  // outside beyond OUT) and the sum of the cutoff values
  const double R0 = warp_max(live ? line_radius(d0, GD, g) : 0.0, 16), D0 = warp_max(live ? line_reach(d0, GD, y) : 0.0, 16);
  double R1 = warp_max(live ? line_radius(d1, GD, g) : 0.0, 32), D1 = warp_max(live ? line_reach(d1, GD, y) : 0.0, 32);
  double R2 = warp_max(live ? line_radius(d2, GD, g) : 0.0, 32), D2 = warp_max(live ? line_reach(d2, GD, y) : 0.0, 32);
  // Window bounds and cutoff sums: a tile without a ByLine cutoff (the common case) has IN = +inf, OUT = +inf where a cluster has
  // a line (0 where it is empty) and no cutoff values - the seven reductions below would only confirm that.  The vote is on the
  // cutoff itself (+inf for a dead lane), not on has_cut: with `__syncthreads_or(has_cut)` nvcc 12.9 folds the OUT operands below
  // to `live ? +inf : 0` (the cutoff case is dropped from the PTX), and every cutoff cluster falls back to the pair-by-pair pass
  // (tests/test_gpu_farfield.py::test_cutoff_case_is_served_by_the_far_field times exactly that)
  const bool tile_has_cut = __syncthreads_or(cutl < DBL_MAX) != 0;
  double I0 = DBL_MAX, O0 = e0 ? 0.0 : DBL_MAX, CS0 = 0.0, I1 = DBL_MAX, O1 = DBL_MAX, I2 = DBL_MAX, O2 = DBL_MAX;
  if (tile_has_cut) {
    I0 = warp_min(has_cut ? cutl - fabs(d0) : DBL_MAX, 16);
    O0 = warp_max(!live ? 0.0 : has_cut ? cutl + fabs(d0) : DBL_MAX, 16);
    CS0 = cval;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) CS0 += __shfl_xor_sync(0xffffffffu, CS0, o);
    I1 = warp_min(has_cut ? cutl - fabs(d1) : DBL_MAX, 32); O1 = warp_max(!live ? 0.0 : has_cut ? cutl + fabs(d1) : DBL_MAX, 32);
    I2 = warp_min(has_cut ? cutl - fabs(d2) : DBL_MAX, 32); O2 = warp_max(!live ? 0.0 : has_cut ? cutl + fabs(d2) : DBL_MAX, 32);
  }
  if ((lane & 31) == 0) {
    double* r = sh[warp];
    r[0] = R1; r[1] = D1; r[2] = I1; r[3] = O1; r[4] = R2; r[5] = D2; r[6] = I2; r[7] = O2;
  }
  __syncthreads();
  R1 = fmax(sh[w1][0], sh[w1 + 1][0]); D1 = fmax(sh[w1][1], sh[w1 + 1][1]);
  R2 = lanes_max(sh[lane & 7][4]); D2 = lanes_max(sh[lane & 7][5]);
  if (tile_has_cut) {
    I1 = fmin(sh[w1][2], sh[w1 + 1][2]); O1 = fmax(sh[w1][3], sh[w1 + 1][3]);
    I2 = lanes_min(sh[lane & 7][6]); O2 = lanes_max(sh[lane & 7][7]);
  } else {
    O1 = e1 ? 0.0 : DBL_MAX;
    O2 = e2 ? 0.0 : DBL_MAX;
  }
  __syncthreads();

  // --- nested service bounds, finest first: a frequency served by a cluster is served by every cluster inside it
  const double rho0 = e0 ? 0.0 : fmax(MP_THETA * R0, D0) * (1.0 + 1e-12);
  const double in0  = e0 ? DBL_MAX : (I0 < DBL_MAX ? I0 * (1.0 - 1e-12) : DBL_MAX);
  const double out0 = e0 ? DBL_MAX : (O0 < DBL_MAX ? O0 * (1.0 + 1e-12) : DBL_MAX);
  if ((lane & 15) == 0) {
    const int q = lane >> 4;
    c0s[q] = c0; rho0s[q] = e0 ? -1.0 : rho0; in0s[q] = in0; out0s[q] = out0; cs0s[q] = CS0; R0s[q] = R0;
  }
  __syncthreads();
  // (the bounds of a 64-line cluster are needed by its first lane only, those of the tile by lane 0)
  double rho1 = 0.0, in1 = DBL_MAX, out1 = DBL_MAX, cs1 = 0.0;
  if ((lane & 63) == 0) {
    rho1 = e1 ? 0.0 : fmax(MP_THETA * R1, D1);
    in1  = I1 < DBL_MAX ? I1 * (1.0 - 1e-12) : DBL_MAX;
    out1 = O1 < DBL_MAX ? O1 * (1.0 + 1e-12) : DBL_MAX;
    for (int q = 0; q < 4; q++) {
      const int qq = (lane >> 6) * 4 + q;
      if (rho0s[qq] < 0.0) continue;
      const double dist = fabs(c0s[qq] - c1);
      rho1 = fmax(rho1, rho0s[qq] + dist);
      in1  = fmin(in1, in0s[qq] < DBL_MAX ? in0s[qq] - dist : DBL_MAX);
      out1 = fmax(out1, out0s[qq] < DBL_MAX ? out0s[qq] + dist : DBL_MAX);
      cs1 += cs0s[qq];
    }
    rho1 = e1 ? 0.0 : rho1 * (1.0 + 1e-12);
    if (e1) { in1 = DBL_MAX; out1 = DBL_MAX; }
    const int sidx = lane >> 6;
    c1s[sidx] = c1; rho1s[sidx] = e1 ? -1.0 : rho1; in1s[sidx] = in1; out1s[sidx] = out1; cs1s[sidx] = cs1; R1s[sidx] = R1;
  }
  if (lane == 0) c2s = c2;
  __syncthreads();
  double rho2 = 0.0, in2 = DBL_MAX, out2 = DBL_MAX, cs2 = 0.0;
  if (lane == 0) {
    rho2 = e2 ? 0.0 : fmax(MP_THETA * R2, D2);
    in2  = I2 < DBL_MAX ? I2 * (1.0 - 1e-12) : DBL_MAX;
    out2 = O2 < DBL_MAX ? O2 * (1.0 + 1e-12) : DBL_MAX;
    for (int sidx = 0; sidx < 4; sidx++) {
      if (rho1s[sidx] < 0.0) continue;
      const double dist = fabs(c1s[sidx] - c2);
      rho2 = fmax(rho2, rho1s[sidx] + dist);
      in2  = fmin(in2, in1s[sidx] < DBL_MAX ? in1s[sidx] - dist : DBL_MAX);
      out2 = fmax(out2, out1s[sidx] < DBL_MAX ? out1s[sidx] + dist : DBL_MAX);
      cs2 += cs1s[sidx];
    }
    rho2 = e2 ? 0.0 : rho2 * (1.0 + 1e-12);
    if (e2) { in2 = DBL_MAX; out2 = DBL_MAX; }
  }

  // --- moments of the 16-line clusters from the lines
  double term[MP_P];
  line_terms(live, d0, GD, g, Si, R0 > 0.0 ? 1.0 / R0 : 0.0, term);
#pragma unroll
  for (int k = 0; k < MP_P; k++) {
    double v = term[k];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    term[k] = v;
  }
  if ((lane & 15) == 0) {
    double* o = o0 + (lane >> 4) * MOM_DOUBLES;
    o[MOM_C] = c0; o[MOM_RHO] = rho0; o[MOM_R] = R0; o[MOM_IN] = in0; o[MOM_OUT] = out0; o[MOM_CUT] = CS0;
#pragma unroll
    for (int k = 0; k < MP_P; k++) { o[MOM_M1 + k] = term[k]; m0s[lane >> 4][k] = term[k]; }
    o[22] = 0.0; o[23] = 0.0;
  }
  __syncthreads();
  // --- 64-line clusters and the tile: shifted sums of their children's moments (no second pass over the lines); one
  // (parent, k, child) per thread, children added in catalog order ((0 + 1) + (2 + 3))
  {
    const int q = lane & 3, k = ((lane >> 2) & 15) + 1, sidx = lane >> 6, qq = sidx * 4 + q;
    const double iR = R1s[sidx] > 0.0 ? 1.0 / R1s[sidx] : 0.0;
    double v = rho0s[qq] < 0.0 ? 0.0 : m2m_term(k, (c0s[qq] - c1s[sidx]) * iR, R0s[qq] * iR, m0s[qq]);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (q == 0) {
      m1s[sidx][k - 1] = v;
      o1[sidx * MOM_DOUBLES + MOM_M1 + k - 1] = v;
    }
  }
  if ((lane & 63) == 0) {
    double* o = o1 + (lane >> 6) * MOM_DOUBLES;
    o[MOM_C] = c1; o[MOM_RHO] = rho1; o[MOM_R] = R1; o[MOM_IN] = in1; o[MOM_OUT] = out1; o[MOM_CUT] = cs1; o[22] = 0.0; o[23] = 0.0;
  }
  __syncthreads();
  if (lane < 4 * MP_P) {
    const int sidx = lane & 3, k = (lane >> 2) + 1;
    const double iR = R2 > 0.0 ? 1.0 / R2 : 0.0;
    double v = rho1s[sidx] < 0.0 ? 0.0 : m2m_term(k, (c1s[sidx] - c2s) * iR, R1s[sidx] * iR, m1s[sidx]);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (sidx == 0) o2[MOM_M1 + k - 1] = v;
  }
  if (lane == 0) {
    o2[MOM_C] = c2s; o2[MOM_RHO] = rho2; o2[MOM_R] = R2; o2[MOM_IN] = in2; o2[MOM_OUT] = out2; o2[MOM_CUT] = cs2; o2[22] = 0.0; o2[23] = 0.0;
  }
}

// ---------------------------------------------------------------------------
// moments of the 16-tile groups: one warp per (group, level), from the 16 tile records
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32) lbl_fmm_moments_group_kernel(PrepareParams p, FmmBuffers fb, const int32_t* __restrict__ tile_seg) {
  const int64_t grp = blockIdx.x;
  const int lev     = blockIdx.y;
  const int lane    = threadIdx.x;
  const int64_t t0  = grp * FMM_GROUP, t1 = min(t0 + FMM_GROUP, p.ntiles);
  double* __restrict__ out = fb.L3 + (int64_t(lev) * fb.ngroups + grp) * MOM_DOUBLES;
  const double* __restrict__ m2 = fb.L2 + int64_t(lev) * p.ntiles * MOM_DOUBLES;
  const double* __restrict__ summ = p.summary + int64_t(lev) * p.ntiles * SUMMARY_DOUBLES;
  // a group is usable when it is complete and inside one real segment; its centre is the midpoint of its lines' centres,
  // its radius the largest reach of a child seen from there (every lane computes the same)
  bool good = t1 - t0 == FMM_GROUP;
  double lo = DBL_MAX, hi = -DBL_MAX;
  for (int64_t t = t0; t < t1; t++) {
    good = good && tile_seg[t] == tile_seg[t0] && tile_seg[t] >= 0;
    const double* s4 = summ + t * SUMMARY_DOUBLES;
    if (s4[0] <= s4[1]) { lo = fmin(lo, s4[0]); hi = fmax(hi, s4[1]); }
  }
  if (!(good && lo <= hi)) {
    if (lane < MOM_DOUBLES) out[lane] = (lane == MOM_RHO || lane == MOM_OUT) ? DBL_MAX : lane == MOM_IN ? -1.0 : 0.0;
    return;
  }
  const double c = 0.5 * (lo + hi);
  double R = 0.0, rho = 0.0, in = DBL_MAX, outb = 0.0, cs = 0.0;
  for (int64_t t = t0; t < t1; t++) {
    const double* s4 = summ + t * SUMMARY_DOUBLES;
    if (s4[0] > s4[1]) continue;  // empty tile
    const double* r = m2 + t * MOM_DOUBLES;
    const double dist = fabs(r[MOM_C] - c);
    R    = fmax(R, r[MOM_R] + dist);
    rho  = fmax(rho, r[MOM_RHO] + dist);
    in   = fmin(in, r[MOM_IN] < DBL_MAX ? r[MOM_IN] - dist : DBL_MAX);
    outb = fmax(outb, r[MOM_OUT] < DBL_MAX ? r[MOM_OUT] + dist : DBL_MAX);
    cs += r[MOM_CUT];
  }
  rho = fmax(rho, MP_THETA * R) * (1.0 + 1e-12);
  const double iR = R > 0.0 ? 1.0 / R : 0.0;
  if (lane < MP_P) {
    double v = 0.0;
    for (int64_t t = t0; t < t1; t++) {  // tiles in catalog order
      const double* s4 = summ + t * SUMMARY_DOUBLES;
      if (s4[0] > s4[1]) continue;
      const double* r = m2 + t * MOM_DOUBLES;
      v += m2m_term(lane + 1, (r[MOM_C] - c) * iR, r[MOM_R] * iR, r + MOM_M1);
    }
    out[MOM_M1 + lane] = v;
  }
  if (lane == 0) {
    out[MOM_C] = c; out[MOM_RHO] = rho; out[MOM_R] = R; out[MOM_IN] = in; out[MOM_OUT] = outb; out[MOM_CUT] = cs; out[22] = 0.0; out[23] = 0.0;
  }
}

// ---------------------------------------------------------------------------
// far-field pass: every frequency sums the expansions of its coarsest accepted clusters, in catalog order
// ---------------------------------------------------------------------------
constexpr int FF_NT = 128, FF_R = 4;

struct Mask4 {
  bool m[FF_R];
};

// adds the expansion of cluster `mo` for the frequencies with on[r] (tt = 0 switches a frequency off exactly)
__device__ __forceinline__ void ff_add(const double* __restrict__ mo, double c, const double* f, const bool* on, double* acc) {
  const double2 m01 = __ldg(reinterpret_cast<const double2*>(mo) + 1);  // R, m_1
  double tau[FF_R], tt[FF_R], s[FF_R];
#pragma unroll
  for (int r = 0; r < FF_R; r++) {
    tt[r]  = on[r] ? fast_rcp(__dsub_rn(f[r], c)) : 0.0;
    tau[r] = __dmul_rn(m01.x, tt[r]);
  }
  const double2* __restrict__ mk = reinterpret_cast<const double2*>(mo) + 2;  // (m_2, m_3) ... (m_16, pad)
  {
    const double2 top = __ldg(mk + (MP_P - 2) / 2);
#pragma unroll
    for (int r = 0; r < FF_R; r++) s[r] = top.x;
  }
#pragma unroll
  for (int j = (MP_P - 2) / 2 - 1; j >= 0; j--) {
    const double2 pr = __ldg(mk + j);
#pragma unroll
    for (int r = 0; r < FF_R; r++) s[r] = __fma_rn(__fma_rn(s[r], tau[r], pr.y), tau[r], pr.x);
  }
  const double cutsum = __ldg(mo + MOM_CUT);  // ls(f) - ls(f0' + cutoff), :591-608: the cluster's cutoff values at once
#pragma unroll
  for (int r = 0; r < FF_R; r++) {
    s[r]   = __fma_rn(s[r], tau[r], m01.y);
    acc[r] = __fma_rn(__dmul_rn(s[r], tau[r]), tt[r], acc[r]);
    acc[r] = __dsub_rn(acc[r], on[r] ? cutsum : 0.0);
  }
}

// One level of the descent.  `parent_on[r]`: frequency r accepts an ancestor (its contribution is already in).  Returns
// whether every frequency of the warp is served by this cluster or an ancestor, and updates on_out.
__device__ __forceinline__ bool ff_visit(const double* __restrict__ mo, const double* f, double fmin_b, double fmax_b, const bool* parent_on,
                                         bool* on_out, double* acc) {
  const double2 cr = __ldg(reinterpret_cast<const double2*>(mo));  // c, rho
  const double in = __ldg(mo + MOM_IN), out = __ldg(mo + MOM_OUT);
  bool add[FF_R];
  bool any = false;
#pragma unroll
  for (int r = 0; r < FF_R; r++) {
    bool ann;
    const bool served = fmm_served(fabs(__dsub_rn(f[r], cr.x)), cr.y, in, out, ann);  // the frequency's own test
    on_out[r] = parent_on[r] || served;
    add[r]    = ann && !parent_on[r];
    any |= add[r];
  }
  if (__any_sync(0xffffffffu, any)) ff_add(mo, cr.x, f, add, acc);
  bool all = true;
#pragma unroll
  for (int r = 0; r < FF_R; r++) all = all && on_out[r];
  return __all_sync(0xffffffffu, all);  // every frequency of the warp is served: nothing below this cluster is needed
}

__global__ void __launch_bounds__(FF_NT) lbl_fmm_far_kernel(SumParams p, FmmBuffers fb) {
  const int tid = threadIdx.x;
  const int lev = blockIdx.y;
  const int64_t fblk = int64_t(blockIdx.x) * (FF_NT * FF_R);
  const double* __restrict__ fg = p.f + int64_t(lev) * p.f_stride;
  const double ffac = p.ffac[lev];
  // a warp owns 32 * FF_R consecutive frequencies: its descent follows a narrow band of the spectrum
  const int64_t wbase = fblk + (tid >> 5) * (32 * FF_R) + (tid & 31);
  double f[FF_R];
#pragma unroll
  for (int r = 0; r < FF_R; r++) {
    const int64_t i = wbase + r * 32;
    f[r] = ffac * fg[i < p.nf ? i : p.nf - 1];
  }
  // bounds of the WARP's frequencies decide which clusters the warp descends into (uniform per warp; the values added
  // never depend on them)
  double fmin_b = fmin(fmin(f[0], f[1]), fmin(f[2], f[3])), fmax_b = fmax(fmax(f[0], f[1]), fmax(f[2], f[3]));
  fmin_b = warp_min(fmin_b, 32);
  fmax_b = warp_max(fmax_b, 32);
  const double* __restrict__ L3 = fb.L3 + int64_t(lev) * fb.ngroups * MOM_DOUBLES;
  const double* __restrict__ L2 = fb.L2 + int64_t(lev) * p.ntiles * MOM_DOUBLES;
  const double* __restrict__ L1 = fb.L1 + int64_t(lev) * p.ntiles * 4 * MOM_DOUBLES;
  const double* __restrict__ L0 = fb.L0 + int64_t(lev) * p.ntiles * 16 * MOM_DOUBLES;
  for (int is = 0; is < p.nsegs; is++) {
    const SegmentDev seg = p.segs[is];
    double acc[FF_R];
    bool none[FF_R];
#pragma unroll
    for (int r = 0; r < FF_R; r++) { acc[r] = 0.0; none[r] = false; }
    for (int64_t g = seg.tile_begin / FMM_GROUP; g * FMM_GROUP < seg.tile_end; g++) {
      bool onG[FF_R];
      if (ff_visit(L3 + g * MOM_DOUBLES, f, fmin_b, fmax_b, none, onG, acc)) continue;
      const int64_t ta = max(g * FMM_GROUP, seg.tile_begin), tb = min((g + 1) * FMM_GROUP, seg.tile_end);
      for (int64_t t = ta; t < tb; t++) {
        bool onT[FF_R];
        if (ff_visit(L2 + t * MOM_DOUBLES, f, fmin_b, fmax_b, onG, onT, acc)) continue;
#pragma unroll 1
        for (int s = 0; s < 4; s++) {
          bool onS[FF_R];
          if (ff_visit(L1 + (t * 4 + s) * MOM_DOUBLES, f, fmin_b, fmax_b, onT, onS, acc)) continue;
#pragma unroll 1
          for (int q = 0; q < 4; q++) {
            bool onQ[FF_R];
            ff_visit(L0 + (t * 16 + s * 4 + q) * MOM_DOUBLES, f, fmin_b, fmax_b, onS, onQ, acc);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < FF_R; r++) {
      const int64_t i = wbase + r * 32;
      if (i < p.k_pitch) fb.far_acc[(int64_t(is) * gridDim.y + lev) * p.k_pitch + i] = acc[r];
    }
  }
}

// ---------------------------------------------------------------------------
// near pass: one thread per frequency sums, pair by pair, the lines of the 16-line clusters it does not accept, adds the
// far-field part, scales, clamps per segment and writes K (K3: ComputeData ctor scale :944-953, clamp :1688-1692)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double fmm_line_scale(double f, double T, double P) {
  constexpr double c = cst::c * cst::c / (8 * cst::pi);
  const double N     = P / (cst::k * T);
  const double r     = (cst::h * f) / (cst::k * T);
  return -N * f * expm1(-r) * c;
}

__global__ void __launch_bounds__(128) lbl_fmm_near_kernel(SumParams p, FmmBuffers fb, int store_full) {
  __shared__ double sacc[4][32][32];  // [warp][frequency of the warp][lane]: every lane's partial sum of every frequency
  const int tid = threadIdx.x, lane = tid & 31;
  const int lev = blockIdx.y;
  const int64_t i = int64_t(blockIdx.x) * 128 + tid;  // this lane's frequency (epilogue); the pair sums are formed warp-wide
  const double* __restrict__ fg = p.f + int64_t(lev) * p.f_stride;
  const double ffac = p.ffac[lev];
  const double f_own = ffac * fg[i < p.nf ? i : p.nf - 1];
  const double* __restrict__ prep = p.prep + int64_t(lev) * p.ntiles * tile_doubles();
  const double* __restrict__ L2 = fb.L2 + int64_t(lev) * p.ntiles * MOM_DOUBLES;
  const double* __restrict__ L1 = fb.L1 + int64_t(lev) * p.ntiles * 4 * MOM_DOUBLES;
  const double* __restrict__ L0 = fb.L0 + int64_t(lev) * p.ntiles * 16 * MOM_DOUBLES;
  const double* __restrict__ scan = fb.scan + int64_t(lev) * p.ntiles * 2;
  const double T = p.T[lev], P = p.P[lev];
  double kacc = 0.0;
  for (int is = 0; is < p.nsegs; is++) {
    const SegmentDev seg = p.segs[is];
    double direct = 0.0;  // this lane's frequency
    // The warp's 32 frequencies share most of their near lines (the window of unaccepted clusters is a dozen grid steps
    // wide).  The warp therefore walks the 16-line clusters between the brackets of its lowest and highest frequency ONCE:
    // lane j tests the cluster against ITS frequency (acceptance data of the cluster and of its 64- and 256-line parents,
    // the same nested test as the far pass) and a ballot tells which frequencies need the cluster's lines pair by pair;
    // the lines are then loaded once - lane (h, k) holds line k of the cluster with q mod 2 == h - and applied to those
    // frequencies, each lane adding into its own partial sum of frequency j (shared memory).  A lane's partial sum of a
    // frequency thus collects the lines k mod 16 of every second needed cluster in catalog order, and a fixed shuffle tree
    // adds the 32 partial sums: the value depends on the frequency and the catalog only.  (Looping over the frequencies
    // outside, as this kernel did before, re-read every line and every cluster header up to 32 times per warp through L1
    // and L2: 33 GB of DRAM traffic per 33 levels, L2 hit rate 59 %.)
    double (*pacc)[32] = sacc[tid >> 5];
#pragma unroll
    for (int j = 0; j < 32; j++) pacc[j][lane] = 0.0;
    const double wf_lo = warp_min(f_own, 32), wf_hi = warp_max(f_own, 32);
    // tiles that may be needed by some frequency of the warp: from the bracket start of the lowest frequency to the bracket
    // end of the highest (scan: running max of c + U from the segment's first tile, running min of c - U from its last,
    // lbl_fmm_scan_kernel; both monotone).  Outside a frequency's own bracket the tile-level test below says "served".
    int64_t a = seg.tile_begin, b = seg.tile_end;
    while (a < b) {  // first tile with max_{t' <= t}(c + U) >= wf_lo
      const int64_t m = (a + b) >> 1;
      if (__ldg(scan + 2 * m) < wf_lo) a = m + 1;
      else b = m;
    }
    int64_t lo = a, hi = seg.tile_end;
    while (lo < hi) {  // first tile with min_{t' >= t}(c - U) > wf_hi
      const int64_t m = (lo + hi) >> 1;
      if (__ldg(scan + 2 * m + 1) <= wf_hi) lo = m + 1;
      else hi = m;
    }
    const int half = lane >> 4, k16 = lane & 15;
#pragma unroll 1
    for (int64_t t = a; t < lo; t++) {
      bool ann;
      const double* __restrict__ r2 = L2 + t * MOM_DOUBLES;
      const double2 c2 = __ldg(reinterpret_cast<const double2*>(r2));
      const bool n2 = !fmm_served(fabs(__dsub_rn(f_own, c2.x)), c2.y, __ldg(r2 + MOM_IN), __ldg(r2 + MOM_OUT), ann);
      if (!__any_sync(0xffffffffu, n2)) continue;  // every frequency of the warp has the tile in its far-field sum
      const double* __restrict__ g0 = prep + t * tile_doubles();
      const int count = p.tile_count[t];
#pragma unroll 1
      for (int s4 = 0; s4 < 4; s4++) {
        if (s4 * 64 >= count) break;
        const double* __restrict__ r1 = L1 + (t * 4 + s4) * MOM_DOUBLES;
        const double2 c1 = __ldg(reinterpret_cast<const double2*>(r1));
        const bool n1 = n2 && !fmm_served(fabs(__dsub_rn(f_own, c1.x)), c1.y, __ldg(r1 + MOM_IN), __ldg(r1 + MOM_OUT), ann);
        if (!__any_sync(0xffffffffu, n1)) continue;
#pragma unroll 1
        for (int qp = 0; qp < 2; qp++) {  // two 16-line clusters per step, one per half warp
          const int qa = s4 * 4 + qp * 2;
          const double* __restrict__ ra0 = L0 + (t * 16 + qa) * MOM_DOUBLES;
          const double* __restrict__ rb0 = ra0 + MOM_DOUBLES;
          const double2 ca = __ldg(reinterpret_cast<const double2*>(ra0)), cb = __ldg(reinterpret_cast<const double2*>(rb0));
          const bool na = n1 && !fmm_served(fabs(__dsub_rn(f_own, ca.x)), ca.y, __ldg(ra0 + MOM_IN), __ldg(ra0 + MOM_OUT), ann);
          const bool nb = n1 && !fmm_served(fabs(__dsub_rn(f_own, cb.x)), cb.y, __ldg(rb0 + MOM_IN), __ldg(rb0 + MOM_OUT), ann);
          const unsigned ma = __ballot_sync(0xffffffffu, na), mb = __ballot_sync(0xffffffffu, nb);
          if ((ma | mb) == 0u) continue;
          unsigned mine = half ? mb : ma;  // the frequencies that need this half warp's cluster
          const int l = (qa + half) * 16 + k16;
          const bool have = mine != 0u && l < count;
          double2 rc = make_double2(0.0, 0.0), ra = rc, rb = rc, rd = rc, re = rc;
          double cutv = 0.0;
          if (have) {  // the line's record, once for all the frequencies that need it
            const double* __restrict__ ga = g0 + (0 * TL + l) * REC_GROUP;
            const double* __restrict__ gb = g0 + (1 * TL + l) * REC_GROUP;
            const double* __restrict__ gc = g0 + (2 * TL + l) * REC_GROUP;
            rc = __ldg(reinterpret_cast<const double2*>(gb));      // B1, igd
            ra = __ldg(reinterpret_cast<const double2*>(ga));      // f0', c3
            rb = __ldg(reinterpret_cast<const double2*>(ga) + 1);  // kappa, A1
            rd = __ldg(reinterpret_cast<const double2*>(gb) + 1);  // y, s_re
            re = __ldg(reinterpret_cast<const double2*>(gc));      // E1(y), the line's cutoff
            if (seg.has_cutoff) cutv = __ldg(gc + 2);
          }
          const bool live = have && rc.y != 0.0;
          while (__any_sync(0xffffffffu, mine != 0u)) {
            const int j = mine ? __ffs(mine) - 1 : 0;
            const bool on = live && mine != 0u;
            mine &= mine - 1u;
            const double f = __shfl_sync(0xffffffffu, f_own, j);
            if (!on) continue;
            // the per-pair arithmetic of lbl_sum_real_kernel's near loop
            double acc = pacc[j][lane];
            if (seg.has_cutoff) {
              // frequency_spans (lbl_lineshape_voigt_lte.h:123-133) and ls(f) - ls(f0' + cutoff) (:591-608)
              const double lcut = re.y;
              if (lcut < DBL_MAX) {
                if (!(ra.x >= f - lcut && ra.x <= f + lcut)) continue;
                acc = __dsub_rn(acc, cutv);
              }
            }
            const double u  = __dsub_rn(f, ra.x);
            const double ax = __dmul_rn(fabs(u), rc.y);
            if (__dadd_rn(ax, rd.x) > FAR_LIMIT_REAL_SUM) {
              acc = far_accumulate_re(acc, u, ra.y, rb.x, rb.y, rc.x);
            } else {
              double wr, wi;
              w_near_fast(rc.y * u, rd.x, re.x, wr, wi);
              acc = __fma_rn(rd.y, wr, acc);
            }
            pacc[j][lane] = acc;
          }
        }
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int j = 0; j < 32; j++) {
      double acc = pacc[j][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == j) direct = acc;
    }
    const double tot = i < p.k_pitch ? fb.far_acc[(int64_t(is) * gridDim.y + lev) * p.k_pitch + i] + direct : 0.0;
    const double F = fmm_line_scale(f_own, T, P) * tot;
    if (!(p.no_negative_absorption && F < 0.0)) kacc += F;
  }
  if (i >= p.nf) {
    if (store_full && i < p.k_pitch) {
      double* o = p.K + (int64_t(lev) * p.k_pitch + i) * 7;
#pragma unroll
      for (int c = 0; c < 7; c++) o[c] = 0.0;
    }
    return;
  }
  double* o = p.K + (int64_t(lev) * p.k_pitch + i) * 7;
  if (store_full) {
    o[0] = kacc;
#pragma unroll
    for (int c = 1; c < 7; c++) o[c] = 0.0;
  } else {
    o[0] += kacc;
  }
}

// running bounds of the tiles' acceptance intervals per (segment, level): scan[t] = {max_{t' <= t}(c + rho), min_{t' >= t}(c - rho)}
__global__ void lbl_fmm_scan_kernel(SumParams p, FmmBuffers fb) {
  // one warp per (segment, level): 32 tiles per step, inclusive max / min scan by shuffles plus the carry of the steps before
  // (max and min are exact and associative: the same numbers as a serial sweep)
  const int lane = threadIdx.x;
  const SegmentDev seg = p.segs[blockIdx.x];
  const int lev = blockIdx.y;
  const double* __restrict__ L2 = fb.L2 + int64_t(lev) * p.ntiles * MOM_DOUBLES;
  double* __restrict__ scan = fb.scan + int64_t(lev) * p.ntiles * 2;
  // U: beyond it the tile serves every frequency (past its acceptance distance, or past every window if it has cutoffs)
  auto U = [&](int64_t t) {
    const double* r = L2 + t * MOM_DOUBLES;
    return r[MOM_IN] < DBL_MAX ? fmax(r[MOM_RHO], r[MOM_OUT]) : r[MOM_RHO];
  };
  double carry = -DBL_MAX;
  for (int64_t t0 = seg.tile_begin; t0 < seg.tile_end; t0 += 32) {
    const int64_t t = t0 + lane;
    double v = t < seg.tile_end ? L2[t * MOM_DOUBLES] + U(t) : -DBL_MAX;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double w = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v = fmax(v, w);
    }
    v = fmax(v, carry);
    if (t < seg.tile_end) scan[2 * t] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  carry = DBL_MAX;
  for (int64_t t1 = seg.tile_end; t1 > seg.tile_begin; t1 -= 32) {  // tiles t1 - 1 - lane, from the segment's last tile down
    const int64_t t = t1 - 1 - lane;
    double v = t >= seg.tile_begin ? L2[t * MOM_DOUBLES] - U(t) : DBL_MAX;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double w = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v = fmin(v, w);
    }
    v = fmin(v, carry);
    if (t >= seg.tile_begin) scan[2 * t + 1] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
}

// ---------------------------------------------------------------------------
int launch_fmm(const PrepareParams& pp, const SumParams& sp, const FmmBuffers& fb, const int32_t* tile_seg, int nlev, int store_full,
               cudaStream_t stream) {
  if (pp.ntiles == 0 || nlev == 0 || sp.nf == 0 || sp.nsegs == 0) return 0;
  {
    // binomial coefficients of the moment shift: per device, before the first launch on it
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !done[dev]) {
      double h[MP_P + 1][MP_P + 1] = {};
      for (int n = 0; n <= MP_P; n++) {
        h[n][0] = 1.0;
        for (int k = 1; k <= n; k++) h[n][k] = h[n - 1][k - 1] + (k <= n - 1 ? h[n - 1][k] : 0.0);
      }
      AB_CUDA(cudaMemcpyToSymbol(BINOM, h, sizeof(h)));
      done[dev] = true;
    }
  }
  lbl_fmm_moments_tile_kernel<<<dim3(static_cast<unsigned>(pp.ntiles), static_cast<unsigned>(nlev)), TL, 0, stream>>>(pp, fb);
  lbl_fmm_moments_group_kernel<<<dim3(static_cast<unsigned>(fb.ngroups), static_cast<unsigned>(nlev)), 32, 0, stream>>>(pp, fb, tile_seg);
  lbl_fmm_scan_kernel<<<dim3(static_cast<unsigned>(sp.nsegs), static_cast<unsigned>(nlev)), 32, 0, stream>>>(sp, fb);
  lbl_fmm_far_kernel<<<dim3(static_cast<unsigned>((sp.nf + FF_NT * FF_R - 1) / (FF_NT * FF_R)), static_cast<unsigned>(nlev)), FF_NT, 0,
                       stream>>>(sp, fb);
  lbl_fmm_near_kernel<<<dim3(static_cast<unsigned>((sp.k_pitch + 127) / 128), static_cast<unsigned>(nlev)), 128, 0, stream>>>(sp, fb, store_full);
  count_launch(5);
  AB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ab200
