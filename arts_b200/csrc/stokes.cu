// stokes.cu — stage 2 of the hot path: the rtepack Stokes chain.
//
//   stokes_chain_kernel    FUSED K4+K5+K6: per frequency, walk the path from the background
//                          to the sensor; the layer transmission matrix T = exp(-r (K_i+K_{i+1})/2),
//                          the linear-in-source operator Lambda and the Planck source are
//                          formed in registers and consumed immediately.  Replaces the
//                          materialising sequence TransmittanceMatrix::init
//                          (reference src/core/rtepack/rtepack_transmission.cc:1114-1193,1254-1328),
//                          SourceVector::init (rtepack_source.cc:52-105) and rte_emission
//                          (rtepack_rtestep.cc:265-404), which round-trips 3 x 128 B per
//                          (frequency, level) through memory.
//   tramat/srcvec/rte kernels   the un-fused compatibility entry points with the
//                          reference's array layouts (index 0 = identity, P cumulative).
//
// The fused kernel is HBM bound on reading K (56 B per frequency-level step): each CTA owns
// 128 consecutive frequencies and streams the K rows of successive levels through shared
// memory with 1-D TMA bulk copies (7168 B per level, 4 stages in flight), so DRAM sees long
// contiguous bursts while every thread reads its own 7 doubles conflict-free (stride 7).
#include "rtepack.cuh"
#include "stokes.hpp"

namespace ab200 {
using namespace rte;

constexpr int ST_NT     = 128;  // 4 warps, each an independent pipeline over 32 frequencies
constexpr int ST_STAGES = 5;    // 5 x 1792 B per warp in flight; 35 KB per CTA -> 6 CTAs per SM

// source of one level: J = B(f,T) e_I, or 0 if K is purely rotational (rtepack_source.cc:88-95, LTE)
__device__ __forceinline__ double source_I(const Propmat& k, double f, double T) {
  return k.is_rotational() ? 0.0 : planck(f, T);
}

// one step of the recursion over layer (i, i+1): k0 = K_i (sensor side), k1 = K_{i+1}.
// SCALAR: the caller guarantees that both levels are unpolarised (only A != 0), so the polarised
// branch is compiled out; the arithmetic of the branch taken is identical in both instantiations.
template <int OPT /* AB200_RTE_* */, bool SCALAR>
__device__ __forceinline__ void rte_step(double* __restrict__ I, const Propmat& k0, const Propmat& k1, double j0 /*J_i*/,
                                         double j1 /*J_{i+1}*/, double r, bool exact, int* __restrict__ flags) {
  constexpr bool LINSRC  = OPT != AB200_RTE_CONSTANT;  // linsrc and linprop share the linevo recursion
  constexpr bool linprop = OPT == AB200_RTE_LINPROP;
  Tran t;
  if (SCALAR) {
    t.a         = -0.5 * r * (k0.A + k1.A);
    t.exp_a     = exp(t.a);
    t.polarized = false;
  } else {
    t.init(k0, k1, r, exact);
  }
  if (SCALAR || !t.polarized) {
    if (LINSRC) {  // linevo :341-370 with scalar T, Lambda
      const double lam = (linprop && linprop_case(k0.A, k1.A, r, false) == 1) ? linprop_lambda(k0.A, k1.A, r, t.exp_a)
                                                                               : func_F(t.a);
      const double dj  = j1 - j0;
      I[0] = t.exp_a * (I[0] - j1) + lam * dj + j0;
      I[1] = t.exp_a * I[1];
      I[2] = t.exp_a * I[2];
      I[3] = t.exp_a * I[3];
    } else {  // constant :287-309
      const double jm = (j0 + j1) * 0.5;  // avg() = std::midpoint, rtepack_stokes_vector.h:120
      I[0] = t.exp_a * (I[0] - jm) + jm;
      I[1] = t.exp_a * I[1];
      I[2] = t.exp_a * I[2];
      I[3] = t.exp_a * I[3];
    }
    return;
  }
  double Tm[16];
  t.T(Tm);
  if (LINSRC) {
    const double v[4] = {I[0] - j1, I[1], I[2], I[3]};  // I - J_{i+1}
    double o[4], l[4];
    mat_vec(Tm, v, o);
    if (linprop && linprop_case(k0.A, k1.A, r, true) == 2) {  // polarised layer with an absorption gradient, :467-474
      double Lc[16];
      linprop_lambda_pol(Tm, k0, k1, r, 1, Lc);
      const double dj = j1 - j0;
      l[0] = Lc[0] * dj; l[1] = Lc[4] * dj; l[2] = Lc[8] * dj; l[3] = Lc[12] * dj;
    } else {
      t.L_col0(j1 - j0, l);
    }
    I[0] = o[0] + l[0] + j0;
    I[1] = o[1] + l[1];
    I[2] = o[2] + l[2];
    I[3] = o[3] + l[3];
  } else {
    const double jm   = (j0 + j1) * 0.5;
    const double v[4] = {I[0] - jm, I[1], I[2], I[3]};
    double o[4];
    mat_vec(Tm, v, o);
    I[0] = o[0] + jm;
    I[1] = o[1];
    I[2] = o[2];
    I[3] = o[3];
  }
}

// Scalar fast path of the FUSED chain (only A != 0 on the whole path): same formulas as rte_step's
// unpolarised branch with the transcendental count cut to two per step — exp(a) and F(a) = expm1(a)/a
// share one expm1, the two divisions become MUFU reciprocals + Newton (fast_rcp, ~1 ulp), and
// hf/kT uses the per-level 1/T.  Differences from the literal form are a few ulp (parity: 1e-9 on I).
template <int OPT>
__device__ __forceinline__ void rte_step_scalar(double* __restrict__ I, double A0, double A1, double j0, double j1, double r) {
  constexpr bool LINSRC  = OPT != AB200_RTE_CONSTANT;
  constexpr bool linprop = OPT == AB200_RTE_LINPROP;
  const double a = -0.5 * r * (A0 + A1);
  double ea;
  if (LINSRC) {
    const double em  = expm1(a);
    ea               = em + 1.0;
    double lam;
    if (linprop && linprop_case(A0, A1, r, false) == 1) lam = linprop_lambda(A0, A1, r, ea);
    else lam = fabs(a) < 1e-8 ? 1.0 + a * 0.5 + a * a / 6.0 : em * fast_rcp(a);
    I[0] = ea * (I[0] - j1) + lam * (j1 - j0) + j0;
  } else {
    ea              = exp(a);
    const double jm = (j0 + j1) * 0.5;
    I[0] = ea * (I[0] - jm) + jm;
  }
  I[1] = ea * I[1];
  I[2] = ea * I[2];
  I[3] = ea * I[3];
}
// B(f,T) = (2h/c^2) f^3 / expm1(h f / k T) with af3 = (2h/c^2) f^3 and invT = 1/T
__device__ __forceinline__ double planck_fast(double f, double af3, double invT) {
  constexpr double b = cst::h / cst::k;
  const double x = (b * f) * invT;
  return x > 700.0 ? 0.0 : af3 * fast_rcp(expm1(x));
}

// Every warp owns 32 consecutive frequencies and its own ring of ST_STAGES shared-memory stages;
// lane 0 refills a stage with one 1792-byte TMA bulk copy (the K rows of 32 frequencies at one
// level) as soon as the warp has moved that stage into registers.  No CTA-wide barrier in the loop.
template <int OPT, bool SCALAR>
__global__ void __launch_bounds__(ST_NT) stokes_chain_kernel(StokesParams p) {
  __shared__ __align__(128) double sK[ST_NT / 32][ST_STAGES][32 * 7];
  __shared__ uint64_t full[ST_NT / 32][ST_STAGES];   // TMA -> the warp's lanes
  __shared__ uint64_t empty[ST_NT / 32][ST_STAGES];  // the warp's 32 lanes -> lane 0 (consumer release)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t iv0 = int64_t(blockIdx.x) * ST_NT + warp * 32;
  if (iv0 >= p.nf) return;  // whole warp idle
  const int64_t iv  = iv0 + lane;
  const bool active = iv < p.nf;
  const int np      = p.np;
  constexpr uint32_t ROW_BYTES = 32 * 7 * sizeof(double);

  if (lane == 0) {
    for (int s = 0; s < ST_STAGES; s++) {
      mbar_init(&full[warp][s], 1);
      mbar_init(&empty[warp][s], 32);
    }
    mbar_fence_init();
  }
  __syncwarp();

  // levels are consumed from np-1 (background side) down to 0; step n reads level np-1-n
  auto issue = [&](int n) {
    const int lev     = np - 1 - n;
    const uint32_t st = n % ST_STAGES;
    mbar_expect_tx(&full[warp][st], ROW_BYTES);
    tma_load_1d(&sK[warp][st][0], p.K + (int64_t(lev) * p.k_pitch + iv0) * 7, ROW_BYTES, &full[warp][st]);
  };
  if (lane == 0)
    for (int n = 0; n < ST_STAGES && n < np; n++) issue(n);

  double I[4] = {0, 0, 0, 0};
  if (active) {
    const double2* b = reinterpret_cast<const double2*>(p.I_bkg + iv * 4);
    const double2 b0 = b[0], b1 = b[1];
    I[0] = b0.x; I[1] = b0.y; I[2] = b1.x; I[3] = b1.y;
  }
  const int64_t ivc = active ? iv : p.nf - 1;
  const double f_raw = p.f[ivc];
  double f = f_raw;
  constexpr double planck_a = 2 * cst::h / (cst::c * cst::c);
  double af3 = planck_a * (f * f * f);
  double ffac_prev = 1.0;  // shared grid: recompute f and f^3 only when the level's wind factor changes
  Propmat k_next{};
  double j_next = 0.0;
  for (int n = 0; n < np; n++) {
    const int lev     = np - 1 - n;
    const uint32_t st = n % ST_STAGES;
    mbar_wait(&full[warp][st], (n / ST_STAGES) & 1);
    Propmat k{};
    if (SCALAR) k.A = sK[warp][st][lane * 7];
    else k = load_propmat(&sK[warp][st][lane * 7]);
    // Stage st may only be refilled once every lane's shared-memory loads have RETURNED (a warp barrier orders
    // instruction issue, not the completion of LDS; a TMA write overtaking a queued LDS was observed on B200).
    // Consumer release, as in lbl_sum_real_kernel: every lane arrives on the stage's `empty` mbarrier after its loads
    // (mbarrier.arrive has release semantics: the lane's prior reads are performed before the arrival is visible);
    // lane 0 waits for all 32 arrivals, orders the generic-proxy reads before the async-proxy write, and refills.
    mbar_arrive(&empty[warp][st]);
    if (lane == 0 && n + ST_STAGES < np) {
      mbar_wait(&empty[warp][st], (n / ST_STAGES) & 1);
      fence_proxy_async_smem();
      issue(n + ST_STAGES);
    }
    const double ffac = p.ffac[lev];
    if (p.f_stride != 0 || ffac != ffac_prev) {
      f         = ffac * (p.f_stride != 0 ? p.f[int64_t(lev) * p.f_stride + ivc] : f_raw);
      af3       = planck_a * (f * f * f);
      ffac_prev = ffac;
    }
    const double j = p.no_emission ? 0.0 : SCALAR ? (k.A == 0.0 ? 0.0 : planck_fast(f, af3, p.invT[lev])) : source_I(k, f, p.T[lev]);
    if (n > 0) {
      if (p.I_lev && active) {  // radiance arriving at level lev+1 (Jacobian pass B)
        double2* o = reinterpret_cast<double2*>(p.I_lev + (int64_t(lev + 1) * p.nf + iv) * 4);
        o[0] = make_double2(I[0], I[1]);
        o[1] = make_double2(I[2], I[3]);
      }
      if (SCALAR) rte_step_scalar<OPT>(I, k.A, k_next.A, j, j_next, p.r[lev]);
      else rte_step<OPT, false>(I, k, k_next, j, j_next, p.r[lev], p.tran_exact != 0, p.flags);
    }
    k_next = k;
    j_next = j;
  }
  if (active) {
    double2* o = reinterpret_cast<double2*>(p.I + iv * 4);
    o[0] = make_double2(I[0], I[1]);
    o[1] = make_double2(I[2], I[3]);
  }
}

int launch_stokes_chain(const StokesParams& p, cudaStream_t stream) {
  if (p.nf == 0) return 0;
  const unsigned grid = static_cast<unsigned>((p.nf + ST_NT - 1) / ST_NT);
  auto go = [&](auto kernel) { kernel<<<grid, ST_NT, 0, stream>>>(p); };
  switch (p.rte_option * 2 + (p.scalar ? 1 : 0)) {
    case AB200_RTE_CONSTANT * 2 + 0: go(stokes_chain_kernel<AB200_RTE_CONSTANT, false>); break;
    case AB200_RTE_CONSTANT * 2 + 1: go(stokes_chain_kernel<AB200_RTE_CONSTANT, true>); break;
    case AB200_RTE_LINSRC * 2 + 0: go(stokes_chain_kernel<AB200_RTE_LINSRC, false>); break;
    case AB200_RTE_LINSRC * 2 + 1: go(stokes_chain_kernel<AB200_RTE_LINSRC, true>); break;
    case AB200_RTE_LINPROP * 2 + 0: go(stokes_chain_kernel<AB200_RTE_LINPROP, false>); break;
    default: go(stokes_chain_kernel<AB200_RTE_LINPROP, true>); break;
  }
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// spectral_planck_op, spectral_radiance_transform_operator.cc:46-87 (forward part)
// ---------------------------------------------------------------------------
__global__ void planck_tb_kernel(int64_t nf, const double* __restrict__ f, double* __restrict__ I) {
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= nf) return;
  double* v       = I + 4 * j;
  const double fj = f[j];
  const double v0 = v[0], v1 = v[1], v2 = v[2], v3 = v[3];
  v[0] = invplanck(v0, fj);
  v[1] = invplanck(0.5 * (v0 + v1), fj) - invplanck(0.5 * (v0 - v1), fj);
  v[2] = invplanck(0.5 * (v0 + v2), fj) - invplanck(0.5 * (v0 - v2), fj);
  v[3] = invplanck(0.5 * (v0 + v3), fj) - invplanck(0.5 * (v0 - v3), fj);
}

// rte_transmission forward part, rtepack_rtestep.cc:469-470: I = P[iv][np-1] * I0
__global__ void transmission_apply_kernel(int np, int64_t nf, const double* __restrict__ P, const double* __restrict__ I_bkg,
                                          double* __restrict__ I) {
  const int64_t iv = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (iv >= nf) return;
  double o[4];
  mat_vec(P + (iv * np + (np - 1)) * 16, I_bkg + iv * 4, o);
  I[iv * 4 + 0] = o[0]; I[iv * 4 + 1] = o[1]; I[iv * 4 + 2] = o[2]; I[iv * 4 + 3] = o[3];
}

int launch_transmission_apply(int np, int64_t nf, const double* P, const double* I_bkg, double* I, cudaStream_t stream) {
  if (nf == 0 || np == 0) return 0;
  transmission_apply_kernel<<<static_cast<unsigned>((nf + 127) / 128), 128, 0, stream>>>(np, nf, P, I_bkg, I);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

int launch_planck_tb(int64_t nf, const double* f, double* I, cudaStream_t stream) {
  if (nf == 0) return 0;
  planck_tb_kernel<<<static_cast<unsigned>((nf + 255) / 256), 256, 0, stream>>>(nf, f, I);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// un-fused compatibility kernels (reference array layouts)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void store16(double* __restrict__ o, const double* __restrict__ m) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) *reinterpret_cast<double2*>(o + i) = make_double2(m[i], m[i + 1]);
}
__device__ __forceinline__ void diag16(double* __restrict__ m, double d) {
#pragma unroll
  for (int i = 0; i < 16; i++) m[i] = 0.0;
  m[0] = m[5] = m[10] = m[15] = d;
}

// TransmittanceMatrix::init forward part: T, L [nf][np][16], index 0 = identity (:1300-1314),
// one thread per (frequency, level); constant :1114-1131, linsrc :1151-1169
__global__ void tramat_kernel(int np, int64_t nf, const double* __restrict__ K, const double* __restrict__ r, int linsrc,
                              int exact, double* __restrict__ T, double* __restrict__ L, int linprop, int* __restrict__ flags) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= nf * np) return;
  const int64_t iv = idx / np;
  const int i      = int(idx % np);
  double m[16];
  if (i == 0) {
    diag16(m, 1.0);
    store16(T + idx * 16, m);
    if (linsrc) store16(L + idx * 16, m);
    return;
  }
  const Propmat k1 = load_propmat(K + (int64_t(i - 1) * nf + iv) * 7);
  const Propmat k2 = load_propmat(K + (int64_t(i) * nf + iv) * 7);
  Tran t;
  t.init(k1, k2, r[i - 1], exact != 0);
  if (t.polarized) t.T(m); else diag16(m, t.exp_a);
  store16(T + idx * 16, m);
  if (linsrc) {
    const int lc = linprop ? linprop_case(k1.A, k2.A, r[i - 1], t.polarized) : 0;
    if (lc == 2) {
      double Tm[16];
#pragma unroll
      for (int e = 0; e < 16; e++) Tm[e] = m[e];
      linprop_lambda_pol(Tm, k1, k2, r[i - 1], 4, m);
    } else if (lc == 1) diag16(m, linprop_lambda(k1.A, k2.A, r[i - 1], t.exp_a));
    else if (t.polarized) t.L(m);
    else diag16(m, func_F(t.a));
    store16(L + idx * 16, m);
  }
}

// P[i,0] = 1, P[i,j] = P[i,j-1] T[i,j] (:1322-1327); one thread per frequency
__global__ void cumtran_kernel(int np, int64_t nf, const double* __restrict__ T, double* __restrict__ P) {
  const int64_t iv = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (iv >= nf) return;
  double acc[16], t[16], o[16];
  diag16(acc, 1.0);
  store16(P + (iv * np) * 16, acc);
  for (int j = 1; j < np; j++) {
#pragma unroll
    for (int e = 0; e < 16; e++) t[e] = T[(iv * np + j) * 16 + e];
    mat_mul(acc, t, o);
#pragma unroll
    for (int e = 0; e < 16; e++) acc[e] = o[e];
    store16(P + (iv * np + j) * 16, acc);
  }
}

// SourceVector::init, LTE: J [nf][np][4] (rtepack_source.cc:85-105)
__global__ void srcvec_kernel(int np, int64_t nf, int nq, const double* __restrict__ K, const double* __restrict__ f,
                              int64_t f_stride, const double* __restrict__ Tlev, int it, double* __restrict__ J,
                              double* __restrict__ dJ) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= nf * np) return;
  const int64_t j = idx / np;
  const int i     = int(idx % np);
  const Propmat k = load_propmat(K + (int64_t(i) * nf + j) * 7);
  const bool rot  = k.is_rotational();
  const double fj = f[int64_t(i) * f_stride + j];
  double* o = J + idx * 4;
  o[0] = rot ? 0.0 : planck(fj, Tlev[i]);
  o[1] = o[2] = o[3] = 0.0;
  for (int q = 0; q < nq; q++) {
    double* d = dJ + (idx * nq + q) * 4;
    d[0] = (!rot && it == q) ? dplanck_dt(fj, Tlev[i]) : 0.0;
    d[1] = d[2] = d[3] = 0.0;
  }
}

// rte_emission forward (nq == 0): constant :287-309 / linevo :341-370 on materialised T, L, J
__global__ void rte_emission_kernel(int linsrc, int np, int64_t nf, const double* __restrict__ T,
                                    const double* __restrict__ L, const double* __restrict__ J,
                                    const double* __restrict__ I_bkg, double* __restrict__ I) {
  const int64_t iv = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (iv >= nf) return;
  double Iv[4];
#pragma unroll
  for (int e = 0; e < 4; e++) Iv[e] = I_bkg[iv * 4 + e];
  for (int i = np - 2; i >= 0; i--) {
    double Tm[16], v[4], o[4];
#pragma unroll
    for (int e = 0; e < 16; e++) Tm[e] = T[(iv * np + i + 1) * 16 + e];
    const double* Ji  = J + (iv * np + i) * 4;
    const double* Ji1 = J + (iv * np + i + 1) * 4;
    if (!linsrc) {
      double Jm[4];
#pragma unroll
      for (int e = 0; e < 4; e++) { Jm[e] = (Ji[e] + Ji1[e]) * 0.5; v[e] = Iv[e] - Jm[e]; }
      mat_vec(Tm, v, o);
#pragma unroll
      for (int e = 0; e < 4; e++) Iv[e] = o[e] + Jm[e];
    } else {
      double Lm[16], dj[4], l[4];
#pragma unroll
      for (int e = 0; e < 16; e++) Lm[e] = L[(iv * np + i + 1) * 16 + e];
#pragma unroll
      for (int e = 0; e < 4; e++) { v[e] = Iv[e] - Ji1[e]; dj[e] = Ji1[e] - Ji[e]; }
      mat_vec(Tm, v, o);
      mat_vec(Lm, dj, l);
#pragma unroll
      for (int e = 0; e < 4; e++) Iv[e] = o[e] + l[e] + Ji[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 4; e++) I[iv * 4 + e] = Iv[e];
}

int launch_tramat(int np, int64_t nf, const double* K, const double* r, int linsrc, int exact, double* T, double* L,
                  double* P, int linprop, int* flags, cudaStream_t stream) {
  if (nf == 0 || np == 0) return 0;
  const int64_t n = nf * np;
  tramat_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(np, nf, K, r, linsrc, exact, T, L, linprop, flags);
  count_launch();
  AB_CUDA(cudaGetLastError());
  cumtran_kernel<<<static_cast<unsigned>((nf + 127) / 128), 128, 0, stream>>>(np, nf, T, P);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

int launch_srcvec(int np, int64_t nf, int nq, const double* K, const double* f, int64_t f_stride, const double* Tlev,
                  int it, double* J, double* dJ, cudaStream_t stream) {
  if (nf == 0 || np == 0) return 0;
  const int64_t n = nf * np;
  srcvec_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(np, nf, nq, K, f, f_stride, Tlev, it, J, dJ);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

int launch_rte_emission(int linsrc, int np, int64_t nf, const double* T, const double* L, const double* J,
                        const double* I_bkg, double* I, cudaStream_t stream) {
  if (nf == 0) return 0;
  rte_emission_kernel<<<static_cast<unsigned>((nf + 127) / 128), 128, 0, stream>>>(linsrc, np, nf, T, L, J, I_bkg, I);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

// stand-alone complex Dawson function (tests, ab200_dawson): the evaluator of the polarised linprop layers
__global__ void dawson_kernel(int64_t n, const double* zr, const double* zi, double* dr, double* di) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const rte::cx d = rte::dawson_c(rte::cx{zr[i], zi[i]});
  dr[i] = d.r;
  di[i] = d.i;
}
int launch_dawson(int64_t n, const double* zr, const double* zi, double* dr, double* di, cudaStream_t stream) {
  if (n == 0) return 0;
  dawson_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, stream>>>(n, zr, zi, dr, di);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ab200
