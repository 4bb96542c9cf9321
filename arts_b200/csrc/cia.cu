// cia.cu — collision-induced absorption added into the resident propagation matrix.
//
//   cia_kernel   spectral_propmatAddCIA (src/m_cia.cc:27-178) for every (frequency, level): CIARecord::Extract
//                (src/core/absorption/cia.cc:214-226) over the datasets of every species pair, cia_interpolation (:76-190)
//                with lagrange_interp's stencil and weights (src/core/matpack/lagrange_interp.h:160-248, 300-440), the
//                temperature Jacobian by perturbation and the VMR Jacobians (m_cia.cc:146-177).
//
// One thread per (frequency, level); the data sets are a few KB to MB of tables that stay in L2.  HBM bound on K:
// 8 B read + 8 B written per (frequency, level) (+ 16 B per affected Jacobian row) against ~100 flop per data set.
#include <memory>
#include <vector>

#include "cia.hpp"

struct ab200_cia {
  int device = 0;
  int32_t n_records = 0;
  // per record
  std::vector<int32_t> h_species1, h_species2, h_ds_begin;  // ds_begin [n_records + 1]
  int32_t *d_species1 = nullptr, *d_species2 = nullptr, *d_ds_begin = nullptr;
  // per dataset: nf, nT, offsets into the pools
  int32_t *d_nf = nullptr, *d_nT = nullptr;
  int64_t *d_f_off = nullptr, *d_T_off = nullptr, *d_data_off = nullptr;
  double* d_pool = nullptr;
  ~ab200_cia() {
    cudaFree(d_species1); cudaFree(d_species2); cudaFree(d_ds_begin); cudaFree(d_nf); cudaFree(d_nT);
    cudaFree(d_f_off); cudaFree(d_T_off); cudaFree(d_data_off); cudaFree(d_pool);
  }
};

namespace ab200 {

// lagrange_interp::update_pos, identity transform, ascending grid: first index of the (order + 1)-point stencil.
// The reference walks from a guess; the fixed point is clamp(m - 1, xf, xe) with m = the first grid index not below x.
__device__ __forceinline__ int stencil_start(const double* __restrict__ xi, int n, int order, double x) {
  const int Pn = order + 1;
  if (n <= Pn) return 0;
  const int Of = order / 2;
  const int xf = Of, xe = n - Pn / 2 - 1;
  int lo = 0, hi = n;  // lower_bound(xi, x)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (xi[mid] < x) lo = mid + 1; else hi = mid;
  }
  int xp = lo - 1;
  xp = xp < xf ? xf : (xp > xe ? xe : xp);
  return xp - xf;
}
// set_weights (non-cyclic): the last weight is one minus the others
__device__ __forceinline__ void lag_weights(double* __restrict__ w, const double* __restrict__ xi, int i0, int order, double x) {
  for (int j = 0; j < order; j++) {
    const double xj = xi[i0 + j];
    double numer = 1.0, denom = 1.0;
    for (int k = 0; k < order; k++) {
      const int m = i0 + k + (k >= j);
      numer *= x - xi[m];
      denom *= xj - xi[m];
    }
    w[j] = numer / denom;
  }
  double last = 1.0;
  for (int j = 0; j < order; j++) last -= w[j];
  w[order] = last;
}

// cia_interpolation for one frequency of one data set; ok = false: temperature outside the extrapolation range
// `grid`, `ngrid`, `gfac`: the level's frequency grid (ascending, times gfac): the reference raises the temperature error for
// the whole call as soon as ONE frequency of the grid lies inside the data set, and then every frequency of that data set is
// NaN under `robust` (cia.cc:95-118 return early only when none does; :184-189)
__device__ __forceinline__ double cia_interp(const CiaDev& c, int ds, double f, double T, double T_extrapolfac, bool& ok,
                                             const double* __restrict__ grid, int64_t ngrid, double gfac) {
  const int nf = c.nf[ds], nT = c.nT[ds];
  const double* __restrict__ fg = c.pool + c.f_off[ds];
  const double* __restrict__ Tg = c.pool + c.T_off[ds];
  const double* __restrict__ dat = c.pool + c.data_off[ds];
  const int T_order = nT - 1 < 3 ? nT - 1 : 3;
  if (T_order > 0 && T_extrapolfac > 0.0) {  // check_limit, lagrange_interp.h:591-605
    const double hi = Tg[nT - 1] + T_extrapolfac * (Tg[nT - 1] - Tg[nT - 2]);
    const double lo = Tg[0] - T_extrapolfac * (Tg[1] - Tg[0]);
    if (hi < T || lo > T) {
      int64_t a = 0, b = ngrid;  // first grid frequency not below the data
      while (a < b) {
        const int64_t mid = (a + b) >> 1;
        if (gfac * grid[mid] < fg[0]) a = mid + 1; else b = mid;
      }
      if (a < ngrid && gfac * grid[a] <= fg[nf - 1]) ok = false;
      return 0.0;
    }
  }
  if (!(f >= fg[0] && f <= fg[nf - 1])) return 0.0;  // outside the data: zero (:95-118)
  double wT[4] = {1.0, 0.0, 0.0, 0.0}, wf[4];
  const int iT = stencil_start(Tg, nT, T_order, T);
  if (T_order > 0) lag_weights(wT, Tg, iT, T_order, T);
  const int i0 = stencil_start(fg, nf, 3, f);
  lag_weights(wf, fg, i0, 3, f);
  double out = 0.0;
  for (int a = 0; a <= 3; a++) {
    if (T_order == 0) {
      out += dat[int64_t(i0 + a) * nT] * wf[a];
    } else {
      for (int b = 0; b <= T_order; b++) out += dat[int64_t(i0 + a) * nT + iT + b] * wf[a] * wT[b];
    }
  }
  return out < 0.0 ? 0.0 : out;  // :181-182
}

__global__ void __launch_bounds__(128) cia_kernel(CiaParams p) {
  const int64_t iv = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (iv >= p.nf) return;
  const int lev = blockIdx.y;
  const double gfac = p.ffac ? p.ffac[lev] : 1.0;
  const double* __restrict__ grid = p.f + int64_t(lev) * p.f_stride;
  const double f = gfac * grid[iv];
  const double T = p.T[lev], P = p.P[lev];
  const double* __restrict__ vmr = p.vmr + int64_t(lev) * p.n_species;
  const double nd = P / (cst::k * T), dnd_dt = -P / (cst::k * (T * T));  // physics_funcs.h:54-72
  double kacc = 0.0, dacc[AB200_MAX_TARGETS];
#pragma unroll
  for (int q = 0; q < AB200_MAX_TARGETS; q++) dacc[q] = 0.0;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  for (int r = 0; r < p.c.n_records; r++) {
    const int s1 = p.c.species1[r], s2 = p.c.species2[r];
    if (p.select_species != AB200_SPECIES_BATH && p.select_species != s1) continue;
    double xsec = 0.0, dxsec = 0.0;
    for (int ds = p.c.ds_begin[r]; ds < p.c.ds_begin[r + 1]; ds++) {
      bool ok = true;
      double v = cia_interp(p.c, ds, f, T, p.T_extrapolfac, ok, grid, p.nf, gfac);
      if (!ok) {
        if (p.ignore_errors) v = nan; else atomicOr(p.flags, 8);
      }
      xsec += v;
      if (p.it >= 0) {
        ok = true;
        double w = cia_interp(p.c, ds, f, T + p.dt, p.T_extrapolfac, ok, grid, p.nf, gfac);
        if (!ok) {
          if (p.ignore_errors) w = nan; else atomicOr(p.flags, 8);
        }
        dxsec += w;
      }
    }
    const double nd_sec = nd * vmr[s2];
    kacc += nd_sec * xsec * nd * vmr[s1];
    if (p.it >= 0) {
      const double dnd_dt_sec = dnd_dt * vmr[s2];
      dacc[p.it] += ((nd_sec * (dxsec - xsec) / p.dt + xsec * dnd_dt_sec) * nd + xsec * nd_sec * dnd_dt) * vmr[s1];
    }
    for (int side = 0; side < 2; side++) {  // jac_targets.find(spec1), find(spec2): the first target of that species
      const int sp = side == 0 ? s1 : s2;
      for (int q = 0; q < p.nq; q++)
        if (p.tg_kind[q] == AB200_TARGET_VMR && p.tg_species[q] == sp) {
          dacc[q] += nd_sec * xsec * nd;
          break;
        }
    }
  }
  p.K[(int64_t(lev) * p.k_pitch + iv) * 7] += kacc;
  for (int q = 0; q < p.nq; q++)
    if (dacc[q] != 0.0 || dacc[q] != dacc[q]) p.dK[((int64_t(lev) * p.nq + q) * p.k_pitch + iv) * 7] += dacc[q];
}

int launch_cia(const CiaParams& p, int nlev, cudaStream_t stream) {
  if (p.nf == 0 || nlev == 0 || p.c.n_records == 0) return 0;
  dim3 grid(static_cast<unsigned>((p.nf + 127) / 128), static_cast<unsigned>(nlev));
  cia_kernel<<<grid, 128, 0, stream>>>(p);
  count_launch();
  AB_CUDA(cudaGetLastError());
  return 0;
}

int cia_device(const ab200_cia* c) { return c->device; }
int cia_max_species(const ab200_cia* c) {
  int m = -1;
  for (int32_t v : c->h_species1) m = v > m ? v : m;
  for (int32_t v : c->h_species2) m = v > m ? v : m;
  return m;
}

CiaDev cia_dev(const ab200_cia* c) {
  return CiaDev{c->n_records, c->d_species1, c->d_species2, c->d_ds_begin, c->d_nf, c->d_nT, c->d_f_off, c->d_T_off, c->d_data_off,
                c->d_pool};
}

}  // namespace ab200

using namespace ab200;

extern "C" int ab200_cia_create(const ab200_cia_record* records, int32_t n_records, ab200_cia** out) {
  if (!out || (n_records > 0 && !records) || n_records < 0) return set_error(AB200_ERR_INVALID, "ab200_cia_create: null argument");
  *out = nullptr;
  std::unique_ptr<ab200_cia> c(new ab200_cia());
  AB_CUDA(cudaGetDevice(&c->device));
  c->n_records = n_records;
  std::vector<int32_t> nf, nT;
  std::vector<int64_t> f_off, T_off, data_off;
  std::vector<double> pool;
  c->h_ds_begin.push_back(0);
  for (int r = 0; r < n_records; r++) {
    const ab200_cia_record& rec = records[r];
    if (rec.species1 < 0 || rec.species2 < 0 || rec.n_datasets < 0 || (rec.n_datasets > 0 && !rec.datasets))
      return set_error(AB200_ERR_INVALID, "ab200_cia_create: record " + std::to_string(r) + " is malformed");
    c->h_species1.push_back(rec.species1);
    c->h_species2.push_back(rec.species2);
    for (int k = 0; k < rec.n_datasets; k++) {
      const ab200_cia_dataset& ds = rec.datasets[k];
      if (ds.nf < 4)  // cia.cc:128-134
        return set_error(AB200_ERR_INVALID, "Not enough frequency grid points in CIA data.\nYou have only " + std::to_string(ds.nf) +
                                                " grid points.\nBut need at least 4.");
      if (ds.nT < 1 || !ds.f_grid || !ds.T_grid || !ds.data) return set_error(AB200_ERR_INVALID, "ab200_cia_create: empty data set");
      for (int i = 1; i < ds.nf; i++)
        if (!(ds.f_grid[i] > ds.f_grid[i - 1])) return set_error(AB200_ERR_INVALID, "ab200_cia_create: frequency grid must ascend");
      for (int i = 1; i < ds.nT; i++)
        if (!(ds.T_grid[i] > ds.T_grid[i - 1])) return set_error(AB200_ERR_INVALID, "ab200_cia_create: temperature grid must ascend");
      nf.push_back(ds.nf);
      nT.push_back(ds.nT);
      f_off.push_back(static_cast<int64_t>(pool.size()));
      pool.insert(pool.end(), ds.f_grid, ds.f_grid + ds.nf);
      T_off.push_back(static_cast<int64_t>(pool.size()));
      pool.insert(pool.end(), ds.T_grid, ds.T_grid + ds.nT);
      data_off.push_back(static_cast<int64_t>(pool.size()));
      pool.insert(pool.end(), ds.data, ds.data + static_cast<size_t>(ds.nf) * ds.nT);
    }
    c->h_ds_begin.push_back(static_cast<int32_t>(nf.size()));
  }
  auto up = [&](auto** d, const auto& v) -> int {
    using T = std::remove_reference_t<decltype(**d)>;
    if (v.empty()) return 0;
    AB_CUDA(cudaMalloc(reinterpret_cast<void**>(d), v.size() * sizeof(T)));
    AB_CUDA(cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
  };
  AB_TRY(up(&c->d_species1, c->h_species1)); AB_TRY(up(&c->d_species2, c->h_species2)); AB_TRY(up(&c->d_ds_begin, c->h_ds_begin));
  AB_TRY(up(&c->d_nf, nf)); AB_TRY(up(&c->d_nT, nT)); AB_TRY(up(&c->d_f_off, f_off)); AB_TRY(up(&c->d_T_off, T_off));
  AB_TRY(up(&c->d_data_off, data_off)); AB_TRY(up(&c->d_pool, pool));
  *out = c.release();
  return AB200_OK;
}

extern "C" void ab200_cia_destroy(ab200_cia* cia) { delete cia; }

// host-buffer form: what the shim of spectral_propmatAddCIA calls (K, dK accumulated like m_cia.cc:146-177)
extern "C" int ab200_cia_levels(const ab200_cia* cia, int64_t nf, const double* f, int64_t f_level_stride, const ab200_atm_path* atm,
                                int32_t n_species, int32_t select_species, int32_t nq, const ab200_target* targets, double dT,
                                double T_extrapolfac, int32_t ignore_errors, double* K, double* dK) {
  if (!cia || !atm || !K || (nf > 0 && !f)) return set_error(AB200_ERR_INVALID, "ab200_cia_levels: null argument");
  if (nf < 0 || atm->np < 0 || nq < 0 || nq > AB200_MAX_TARGETS || n_species <= 0)
    return set_error(AB200_ERR_INVALID, "ab200_cia_levels: bad size");
  if (nq > 0 && (!targets || !dK)) return set_error(AB200_ERR_INVALID, "ab200_cia_levels: null Jacobian argument with nq > 0");
  if (f_level_stride != 0 && f_level_stride != nf) return set_error(AB200_ERR_INVALID, "f_level_stride must be 0 or nf");
  if (cia_max_species(cia) >= n_species) return set_error(AB200_ERR_INVALID, "ab200_cia_levels: a CIA record names a species beyond n_species");
  for (int q = 0; q < nq; q++)
    if (targets[q].kind >= AB200_TARGET_WIND_U && targets[q].kind <= AB200_TARGET_WIND_W)
      return set_error(AB200_ERR_UNSUPPORTED, "collision-induced absorption with a wind target (the re-extraction at f + df of "
                                              "src/m_cia.cc:78-81, :123-129) is outside the GPU path; no CPU fallback");
  const int np = atm->np;
  if (np == 0 || nf == 0) return AB200_OK;
  CiaParams cp{};
  cp.it = -1;
  for (int q = 0; q < nq; q++) {
    cp.tg_kind[q] = targets[q].kind;
    cp.tg_species[q] = targets[q].species;
    if (targets[q].kind == AB200_TARGET_T && cp.it < 0) cp.it = q;
  }
  if (cp.it >= 0 && !std::isnormal(dT)) return set_error(AB200_ERR_INVALID, "dt must be >0 and not NaN or Inf: " + std::to_string(dT));
  for (int ip = 0; ip < np; ip++) {  // m_cia.cc:56-57
    if (!(atm->T[ip] > 0)) return set_error(AB200_ERR_INVALID, "Non-positive temperature");
    if (!(atm->P[ip] > 0)) return set_error(AB200_ERR_INVALID, "Non-positive pressure");
  }
  AB_CUDA(cudaSetDevice(cia->device));
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
    return set_error(AB200_ERR_NOMEM, "ab200_cia_levels: device allocation or copy failed");
  cp.c = cia_dev(cia);
  cp.nf = nf; cp.f = static_cast<double*>(bf.p); cp.f_stride = f_level_stride; cp.ffac = nullptr;
  cp.T = static_cast<double*>(bT.p); cp.P = static_cast<double*>(bP.p); cp.vmr = static_cast<double*>(bv.p);
  cp.n_species = n_species; cp.select_species = select_species;
  cp.K = static_cast<double*>(bK.p); cp.dK = static_cast<double*>(bdK.p); cp.k_pitch = nf; cp.nq = nq;
  cp.dt = dT; cp.T_extrapolfac = T_extrapolfac; cp.ignore_errors = ignore_errors; cp.flags = static_cast<int*>(bfl.p);
  AB_TRY(launch_cia(cp, np, nullptr));
  int h = 0;
  AB_CUDA(cudaMemcpy(&h, bfl.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (h & 8)
    return set_error(AB200_ERR_INVALID, "Problem with CIA species: the temperature of a level is outside the extrapolation range of a "
                                        "data set (check_limit for Temperature, lagrange_interp.h:572-650)");
  AB_CUDA(cudaMemcpy(K, bK.p, nk * 8, cudaMemcpyDeviceToHost));
  if (nq > 0) AB_CUDA(cudaMemcpy(dK, bdK.p, nk * nq * 8, cudaMemcpyDeviceToHost));
  return AB200_OK;
}
