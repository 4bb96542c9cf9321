// lookup.cu — absorption lookup tables added into the resident propagation matrix.
//
//   lookup_kernel   _spectral_propmatAddLookup (src/m_lookup.cc:20-141) for every (frequency, level): table::absorption
//                   (src/core/lookup/lookup_map.cpp:190-238) of every selected species - Lagrange interpolation of the given
//                   orders in temperature offset, water ratio, log-pressure (descending grid) and frequency
//                   (pressure_/frequency_/water_/temperature_lagrange :133-188; stencil, weights and limits of
//                   src/core/matpack/lagrange_interp.h:160-248, 300-440, 572-650) times the species' number density - the
//                   no_negative_absorption filter and the Jacobian rows by re-extraction at the perturbed point.
//
// The tables themselves are what the line-by-line path produces: xsec[it][iw][ip][f] = K.A / number density for the reference
// profile with temperature offset t_pert[it] and water ratio w_pert[iw] (table ctor :22-131), i.e. ab200_propmat_levels over
// nt * nw * np levels.  One thread per (frequency, level); (to+1)(wo+1)(po+1)(fo+1) table reads per extraction, from L2.
#include <cmath>
#include <memory>
#include <vector>

#include "lookup.hpp"

struct ab200_lookup {
  int device = 0;
  int32_t n_tables = 0;
  std::vector<int32_t> h_meta;  // [n_tables][8]
  int32_t* d_meta = nullptr;
  int64_t* d_off = nullptr;
  double* d_pool = nullptr;
  ~ab200_lookup() { cudaFree(d_meta); cudaFree(d_off); cudaFree(d_pool); }
};

namespace ab200 {

struct Lag {
  int i0, order;
  double w[LUT_MAXP];
};

// lagrange_interp: check_limit, the fixed point of update_pos (both grid orders; nearest neighbour for order 0) and set_weights
__device__ __forceinline__ bool make_lag(Lag& l, const double* __restrict__ xi, int n, int order, double x, double limit) {
  const bool ascending = n <= 1 || xi[0] < xi[1];
  bool ok = true;
  if (order > 0 && limit > 0.0) {
    const double hi = ascending ? xi[n - 1] + limit * (xi[n - 1] - xi[n - 2]) : xi[0] + limit * (xi[0] - xi[1]);
    const double lo = ascending ? xi[0] - limit * (xi[1] - xi[0]) : xi[n - 1] - limit * (xi[n - 2] - xi[n - 1]);
    if (hi < x || lo > x) ok = false;
  }
  const int P = order + 1;
  l.order = order;
  if (n <= P) {
    l.i0 = 0;
  } else {
    const int Of = order / 2, xf = Of, xe = n - P / 2 - 1;
    // first index m whose value is not before x in grid order; the walk of update_pos ends at clamp(m - 1, xf, xe)
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const bool before = ascending ? xi[mid] < x : xi[mid] > x;
      if (before) lo = mid + 1; else hi = mid;
    }
    int xp = lo - 1;
    xp = xp < xf ? xf : (xp > xe ? xe : xp);
    if (order == 0) {
      const int xn = xp + 1;
      xp = (xn == n || fabs(x - xi[xn]) > fabs(x - xi[xp])) ? xp : xn;
      xp = xp < xf ? xf : (xp > xe ? xe : xp);
    }
    l.i0 = xp - xf;
  }
  for (int j = 0; j < order; j++) {
    const double xj = xi[l.i0 + j];
    double numer = 1.0, denom = 1.0;
    for (int k = 0; k < order; k++) {
      const int m = l.i0 + k + (k >= j);
      numer *= x - xi[m];
      denom *= xj - xi[m];
    }
    l.w[j] = numer / denom;
  }
  double last = 1.0;
  for (int j = 0; j < order; j++) last -= l.w[j];
  l.w[order] = last;
  return ok;
}

__device__ __forceinline__ double interp1(const double* __restrict__ field, const Lag& l) {
  double out = 0.0;
  for (int i = 0; i <= l.order; i++) out += field[l.i0 + i] * l.w[i];
  return out;
}

// What table::absorption needs of one (table, atmospheric state): the Lagrange stencils and weights in log-pressure, temperature
// offset and water ratio, and the number density of the species.  They depend on the level (and on the perturbed Jacobian
// point) only, so one thread of the CTA builds them in shared memory and every frequency reuses them; a thread's own work is
// the frequency stencil, once per table, and the (to+1)(wo+1)(po+1)(fo+1) products.
struct LutState {
  Lag pl, tl, wl;
  double scale;  // vmr(species) * P / (k T)
  int ok;
};
constexpr int LUT_MAX_STATES = 1 + AB200_MAX_TARGETS;

__device__ void lut_state(const LutParams& p, int k, double T, double P, const double* __restrict__ vmr, int sp_pert, double dv,
                          LutState& st) {
  const int32_t* m = p.t.meta + 8 * k;
  const int species = m[0], np = m[2], nt = m[3], nw = m[4], do_t = m[5], do_w = m[6];
  const int64_t* o = p.t.off + 7 * k;
  const double* pool = p.t.pool;
  auto v_of = [&](int s) { return vmr[s] + (s == sp_pert ? dv : 0.0); };
  bool ok = true;
  st.tl.i0 = st.wl.i0 = 0; st.tl.order = st.wl.order = 0; st.tl.w[0] = st.wl.w[0] = 1.0;
  ok &= make_lag(st.pl, pool + o[1], np, p.po, log(P), p.extpol);
  if (do_w) ok &= make_lag(st.wl, pool + o[3], nw, p.wo, v_of(p.h2o_species) / interp1(pool + o[5], st.pl), p.extpol);
  if (do_t) ok &= make_lag(st.tl, pool + o[2], nt, p.to, T - interp1(pool + o[4], st.pl), p.extpol);
  st.scale = v_of(species) * (P / (cst::k * T));
  st.ok = ok ? 1 : 0;
}

// table::absorption (lookup_map.cpp:190-238) for one frequency stencil and one prepared state
__device__ __forceinline__ double table_absorption(const LutParams& p, int k, const Lag& fl, const LutState& st) {
  const int32_t* m = p.t.meta + 8 * k;
  const int nf = m[1], np = m[2], nw = m[4], do_t = m[5], do_w = m[6];
  const double* __restrict__ xs = p.t.pool + p.t.off[7 * k + 6];
  double out = 0.0;
  for (int a = 0; a <= st.tl.order; a++)
    for (int b = 0; b <= st.wl.order; b++)
      for (int c = 0; c <= st.pl.order; c++) {
        const double* __restrict__ row = xs + ((int64_t(st.tl.i0 + a) * nw + st.wl.i0 + b) * np + st.pl.i0 + c) * nf + fl.i0;
        for (int d = 0; d <= fl.order; d++) {
          double v = row[d];
          if (do_t) v *= st.tl.w[a];
          if (do_w) v *= st.wl.w[b];
          v *= st.pl.w[c];
          v *= fl.w[d];
          out += v;
        }
      }
  return out * st.scale;
}

__global__ void __launch_bounds__(128) lookup_kernel(LutParams p) {
  __shared__ LutState sst[LUT_MAX_STATES];
  const int64_t iv = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int lev = blockIdx.y;
  const bool live = iv < p.nf;
  const double f = live ? (p.ffac ? p.ffac[lev] : 1.0) * p.f[int64_t(lev) * p.f_stride + iv] : 0.0;
  const double T = p.T[lev], P = p.P[lev];
  const double* __restrict__ vmr = p.vmr + int64_t(lev) * p.n_species;
  bool ok = true;
  double ab = 0.0, dab[AB200_MAX_TARGETS];
#pragma unroll
  for (int q = 0; q < AB200_MAX_TARGETS; q++) dab[q] = 0.0;
  for (int k = 0; k < p.t.n_tables; k++) {
    const int32_t* m = p.t.meta + 8 * k;
    if (p.select_species != AB200_SPECIES_BATH && m[0] != p.select_species) continue;  // CTA-uniform
    if (int64_t(m[1]) * m[2] * m[3] * m[4] == 0) continue;                              // xsec.empty(), lookup_map.cpp:201
    __syncthreads();  // the previous table's states have been read
    if (threadIdx.x <= p.nq) {  // state 0: the level itself; state 1 + q: the perturbed point of target q
      const int q = int(threadIdx.x) - 1;
      const bool is_T = q >= 0 && p.tg_kind[q] == AB200_TARGET_T;
      lut_state(p, k, (q >= 0 && is_T) ? T + p.tg_d[q] : T, P, vmr, (q >= 0 && !is_T) ? p.tg_species[q] : -1,
                (q >= 0 && !is_T) ? p.tg_d[q] : 0.0, sst[threadIdx.x]);
    }
    __syncthreads();
    if (!live) continue;
    Lag fl;
    ok &= make_lag(fl, p.t.pool + p.t.off[7 * k], m[1], p.fo, f, p.extpol);
    ok &= sst[0].ok != 0;
    ab += table_absorption(p, k, fl, sst[0]);
    for (int q = 0; q < p.nq; q++) {
      ok &= sst[1 + q].ok != 0;
      dab[q] += table_absorption(p, k, fl, sst[1 + q]);
    }
  }
  if (!live) return;
  if (p.no_neg == 0 || ab > 0.0) p.K[(int64_t(lev) * p.k_pitch + iv) * 7] += ab;  // m_lookup.cc:73-77
  for (int q = 0; q < p.nq; q++) {
    const double d = p.tg_d[q];
    if (p.no_neg == 0 || dab[q] > 0.0)  // the row is ASSIGNED (sic, :130-135)
      p.dK[((int64_t(lev) * p.nq + q) * p.k_pitch + iv) * 7] = (dab[q] - ab) * (1.0 / d);
  }
  if (!ok) atomicOr(p.flags, 16);
}

int launch_lookup(const LutParams& p, int nlev, cudaStream_t stream) {
  if (p.nf == 0 || nlev == 0) return 0;
  dim3 grid(static_cast<unsigned>((p.nf + 127) / 128), static_cast<unsigned>(nlev));
  lookup_kernel<<<grid, 128, 0, stream>>>(p);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

__global__ void xsec_from_K_kernel(int64_t nf, const double* __restrict__ K, int64_t k_pitch, const double* __restrict__ T,
                                   const double* __restrict__ P, const double* __restrict__ vmr, int n_species, int species,
                                   double* __restrict__ xsec) {
  const int64_t iv = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (iv >= nf) return;
  const int lev = blockIdx.y;
  const double nd     = vmr[int64_t(lev) * n_species + species] * (P[lev] / (cst::k * T[lev]));  // AtmPoint::number_density(species)
  const double inv_nd = 1.0 / nd;                                                               // lookup_map.cpp:108
  xsec[int64_t(lev) * nf + iv] = K[(int64_t(lev) * k_pitch + iv) * 7] * inv_nd;
}

int launch_xsec_from_K(int np, int64_t nf, const double* K, int64_t k_pitch, const double* T, const double* P, const double* vmr,
                       int n_species, int species, double* xsec, cudaStream_t stream) {
  if (np == 0 || nf == 0) return 0;
  dim3 grid(static_cast<unsigned>((nf + 255) / 256), static_cast<unsigned>(np));
  xsec_from_K_kernel<<<grid, 256, 0, stream>>>(nf, K, k_pitch, T, P, vmr, n_species, species, xsec);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

LutDev lut_dev(const ab200_lookup* l) { return LutDev{l->n_tables, l->d_meta, l->d_off, l->d_pool}; }
int lut_device(const ab200_lookup* l) { return l->device; }

int lut_check_call(const ab200_lookup* l, int32_t n_species, int32_t h2o_species, int32_t select_species, int po, int to, int wo,
                   int fo) {
  if (po < 0 || to < 0 || wo < 0 || fo < 0) return set_error(AB200_ERR_INVALID, "negative interpolation order");
  if (std::max(std::max(po, to), std::max(wo, fo)) >= LUT_MAXP)
    return set_error(AB200_ERR_UNSUPPORTED, "interpolation orders above 7 are outside the GPU path");
  bool found = select_species == AB200_SPECIES_BATH;
  for (int k = 0; k < l->n_tables; k++) {
    const int32_t* m = l->h_meta.data() + 8 * k;
    if (m[0] >= n_species) return set_error(AB200_ERR_INVALID, "a lookup table names a species the atmosphere does not have");
    if (select_species != AB200_SPECIES_BATH && m[0] != select_species) continue;
    found = true;
    if (int64_t(m[1]) * m[2] * m[3] * m[4] == 0) continue;
    // check_limit: "Too few grid points for the given polynomial order" (lagrange_interp.h:585-588)
    if (fo >= m[1]) return set_error(AB200_ERR_INVALID, "Error in check_limit for Frequency:\nToo few grid points for the given polynomial order");
    if (po >= m[2]) return set_error(AB200_ERR_INVALID, "Error in check_limit for Log-Pressure:\nToo few grid points for the given polynomial order");
    if (m[5] && to >= m[3]) return set_error(AB200_ERR_INVALID, "Error in check_limit for Temperature:\nToo few grid points for the given polynomial order");
    if (m[6] && wo >= m[4]) return set_error(AB200_ERR_INVALID, "Error in check_limit for Water VMR:\nToo few grid points for the given polynomial order");
    if (m[6] && (h2o_species < 0 || h2o_species >= n_species))
      return set_error(AB200_ERR_INVALID, "a lookup table has a water grid but the call names no H2O species");
  }
  if (!found) return set_error(AB200_ERR_INVALID, "no lookup table for the selected species");  // unordered_map::at, m_lookup.cc:50
  return 0;
}

}  // namespace ab200

using namespace ab200;

extern "C" int ab200_lookup_create(const ab200_lookup_table* tables, int32_t n_tables, ab200_lookup** out) {
  if (!out || (n_tables > 0 && !tables) || n_tables < 0) return set_error(AB200_ERR_INVALID, "ab200_lookup_create: null argument");
  *out = nullptr;
  std::unique_ptr<ab200_lookup> l(new ab200_lookup());
  AB_CUDA(cudaGetDevice(&l->device));
  l->n_tables = n_tables;
  std::vector<int64_t> off;
  std::vector<double> pool;
  auto push = [&](const double* p, size_t n) {
    off.push_back(static_cast<int64_t>(pool.size()));
    if (p && n) pool.insert(pool.end(), p, p + n);
  };
  for (int k = 0; k < n_tables; k++) {
    const ab200_lookup_table& t = tables[k];
    if (t.species < 0 || t.nf < 0 || t.np < 0 || t.nt < 1 || t.nw < 1)
      return set_error(AB200_ERR_INVALID, "ab200_lookup_create: table " + std::to_string(k) + " has bad sizes");
    if (!t.f_grid || !t.log_p_grid)  // table::check, lookup_map.cpp:240-247
      return set_error(AB200_ERR_INVALID, "Must have frequency and pressure grids.");
    if ((t.do_t && (!t.t_pert || !t.t_atmref)) || (t.do_w && (!t.w_pert || !t.water_atmref)) || !t.xsec)
      return set_error(AB200_ERR_INVALID, "ab200_lookup_create: table " + std::to_string(k) + " lacks an array its flags announce");
    if ((!t.do_t && t.nt != 1) || (!t.do_w && t.nw != 1))
      return set_error(AB200_ERR_INVALID, "The shape of the absorption cross section table is incorrect.");
    for (int i = 1; i < t.nf; i++)
      if (!(t.f_grid[i] > t.f_grid[i - 1])) return set_error(AB200_ERR_INVALID, "ab200_lookup_create: f_grid must ascend");
    for (int i = 1; i < t.np; i++)
      if (!(t.log_p_grid[i] < t.log_p_grid[i - 1])) return set_error(AB200_ERR_INVALID, "ab200_lookup_create: log_p_grid must descend");
    const int32_t meta[8] = {t.species, t.nf, t.np, t.nt, t.nw, t.do_t ? 1 : 0, t.do_w ? 1 : 0, 0};
    l->h_meta.insert(l->h_meta.end(), meta, meta + 8);
    push(t.f_grid, t.nf);
    push(t.log_p_grid, t.np);
    push(t.do_t ? t.t_pert : nullptr, t.do_t ? t.nt : 0);
    push(t.do_w ? t.w_pert : nullptr, t.do_w ? t.nw : 0);
    push(t.do_t ? t.t_atmref : nullptr, t.do_t ? t.np : 0);
    push(t.do_w ? t.water_atmref : nullptr, t.do_w ? t.np : 0);
    push(t.xsec, static_cast<size_t>(t.nt) * t.nw * t.np * t.nf);
  }
  if (n_tables) {
    AB_CUDA(cudaMalloc(reinterpret_cast<void**>(&l->d_meta), l->h_meta.size() * sizeof(int32_t)));
    AB_CUDA(cudaMemcpy(l->d_meta, l->h_meta.data(), l->h_meta.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    AB_CUDA(cudaMalloc(reinterpret_cast<void**>(&l->d_off), off.size() * sizeof(int64_t)));
    AB_CUDA(cudaMemcpy(l->d_off, off.data(), off.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    AB_CUDA(cudaMalloc(reinterpret_cast<void**>(&l->d_pool), std::max<size_t>(pool.size(), 1) * sizeof(double)));
    if (!pool.empty()) AB_CUDA(cudaMemcpy(l->d_pool, pool.data(), pool.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  *out = l.release();
  return AB200_OK;
}

extern "C" void ab200_lookup_destroy(ab200_lookup* lut) { delete lut; }

extern "C" int ab200_lookup_levels(const ab200_lookup* lut, int64_t nf, const double* f, int64_t f_level_stride,
                                   const ab200_atm_path* atm, int32_t n_species, int32_t h2o_species, int32_t select_species, int32_t nq,
                                   const ab200_target* targets, const double* target_d, int32_t no_negative_absorption,
                                   int32_t p_interp_order, int32_t t_interp_order, int32_t water_interp_order, int32_t f_interp_order,
                                   double extpolfac, double* K, double* dK) {
  if (!lut || !atm || !K || (nf > 0 && !f)) return set_error(AB200_ERR_INVALID, "ab200_lookup_levels: null argument");
  if (nf < 0 || atm->np < 0 || nq < 0 || nq > AB200_MAX_TARGETS || n_species <= 0)
    return set_error(AB200_ERR_INVALID, "ab200_lookup_levels: bad size");
  if (nq > 0 && (!targets || !dK || !target_d)) return set_error(AB200_ERR_INVALID, "ab200_lookup_levels: null Jacobian argument with nq > 0");
  if (f_level_stride != 0 && f_level_stride != nf) return set_error(AB200_ERR_INVALID, "f_level_stride must be 0 or nf");
  AB_TRY(lut_check_call(lut, n_species, h2o_species, select_species, p_interp_order, t_interp_order, water_interp_order, f_interp_order));
  LutParams lp{};
  for (int q = 0; q < nq; q++) {
    if (targets[q].kind != AB200_TARGET_T && targets[q].kind != AB200_TARGET_VMR)
      return set_error(AB200_ERR_UNSUPPORTED, "only temperature and VMR targets are on the GPU path");
    if (targets[q].kind == AB200_TARGET_VMR && (targets[q].species < 0 || targets[q].species >= n_species))
      return set_error(AB200_ERR_INVALID, "Jacobian target species out of range");
    if (!std::isnormal(target_d[q]))  // m_lookup.cc:84-87
      return set_error(AB200_ERR_INVALID, "The target " + std::to_string(q) + " is not good, it lacks a perturbation value.");
    lp.tg_kind[q] = targets[q].kind; lp.tg_species[q] = targets[q].species; lp.tg_d[q] = target_d[q];
  }
  const int np = atm->np;
  if (np == 0 || nf == 0) return AB200_OK;
  AB_CUDA(cudaSetDevice(lut->device));
  struct Buf {
    void* p = nullptr;
    ~Buf() { cudaFree(p); }
    int put(const void* src, size_t bytes) {
      if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) { cudaGetLastError(); return 1; }
      if (src && bytes && cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); return 1; }
      return 0;
    }
  } bf, bT, bP, bv, bK, bdK, bfl;
  const size_t nfl = static_cast<size_t>(nf) * (f_level_stride ? np : 1), nk = static_cast<size_t>(np) * nf * 7;
  const int zero = 0;
  if (bf.put(f, nfl * 8) || bT.put(atm->T, np * 8) || bP.put(atm->P, np * 8) || bv.put(atm->vmr, static_cast<size_t>(np) * n_species * 8) ||
      bK.put(K, nk * 8) || bdK.put(dK, nk * nq * 8) || bfl.put(&zero, sizeof(int)))
    return set_error(AB200_ERR_NOMEM, "ab200_lookup_levels: device allocation or copy failed");
  lp.t = lut_dev(lut);
  lp.nf = nf; lp.f = static_cast<double*>(bf.p); lp.f_stride = f_level_stride; lp.ffac = nullptr;
  lp.T = static_cast<double*>(bT.p); lp.P = static_cast<double*>(bP.p); lp.vmr = static_cast<double*>(bv.p);
  lp.n_species = n_species; lp.h2o_species = h2o_species; lp.select_species = select_species;
  lp.K = static_cast<double*>(bK.p); lp.dK = static_cast<double*>(bdK.p); lp.k_pitch = nf; lp.nq = nq;
  lp.no_neg = no_negative_absorption; lp.po = p_interp_order; lp.to = t_interp_order; lp.wo = water_interp_order; lp.fo = f_interp_order;
  lp.extpol = extpolfac; lp.flags = static_cast<int*>(bfl.p);
  AB_TRY(launch_lookup(lp, np, nullptr));
  int h = 0;
  AB_CUDA(cudaMemcpy(&h, bfl.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (h & 16)
    return set_error(AB200_ERR_INVALID, "Error in check_limit: a frequency, pressure, temperature offset or water ratio is outside "
                                        "the extrapolation limits of a lookup table grid (lagrange_interp.h:572-650)");
  AB_CUDA(cudaMemcpy(K, bK.p, nk * 8, cudaMemcpyDeviceToHost));
  if (nq > 0) AB_CUDA(cudaMemcpy(dK, bdK.p, nk * nq * 8, cudaMemcpyDeviceToHost));
  return AB200_OK;
}
