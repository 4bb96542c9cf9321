// api.cu — the C ABI of include/arts_b200.h: error plumbing, the device-resident path
// workspace and the host-buffer entry points the ARTS workspace-method shims call.
//
// Host orchestration only; the arithmetic is in lbl.cu (stage 1) and stokes.cu (stage 2).
// No CPU fallback anywhere: a missing device or a failing CUDA call is an error return.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "catalog.hpp"
#include "lbl.hpp"
#include "cia.hpp"
#include "lookup.hpp"
#include "predef.hpp"
#include "stokes.hpp"

namespace ab200 {

static thread_local std::string g_err;
static thread_local int64_t g_launches = 0;

int set_error(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" + std::to_string(line) + ")";
  cudaGetLastError();  // clear the sticky-less error state
  return AB200_ERR_CUDA;
}
void count_launch(int n) { g_launches += n; }

static constexpr size_t PREP_BUDGET_BYTES = size_t(8) << 30;  // line records kept resident per batch of levels

template <typename T>
static int dev_alloc(T** p, size_t n) {
  *p = nullptr;
  if (n == 0) return 0;
  AB_CUDA(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
  return 0;
}

}  // namespace ab200

using namespace ab200;

struct ab200_path {
  const ab200_catalog* cat = nullptr;
  int64_t nf = 0, k_pitch = 0;
  int32_t np = 0, nq = 0;
  int32_t np_cap = 0;  // levels the workspace was created for; an upload may bring fewer (ragged batches of paths)
  cudaStream_t stream = nullptr;
  bool own_stream = false;

  // inputs
  double* d_f = nullptr;      // [np][nf] capacity
  double* d_small = nullptr;  // packed small per-level arrays, see upload()
  double *d_T = nullptr, *d_P = nullptr, *d_H = nullptr, *d_vmr = nullptr, *d_isorat = nullptr, *d_Q = nullptr,
         *d_npm = nullptr, *d_frange = nullptr, *d_r = nullptr;
  double* h_small = nullptr;  // pinned staging of the same
  size_t small_doubles = 0;
  double* d_Ibkg = nullptr;
  SegmentDev* d_segs = nullptr;  // [3][nsegments] capacity: mode-0 list (line-by-line kernel), mode-1 list, mode-0 far-field list
  SegmentDev* h_segs = nullptr;
  int32_t nsegs[3] = {0, 0, 0};  // [2]: real segments summed by lbl_fmm.cu ([0]: by the line-by-line kernel, AB200_FARFIELD=0)
  FmmBuffers fmm{};              // one allocation (fmm.L0), sized for levels_per_batch levels
  int32_t* d_tile_seg = nullptr; // [ntiles] index of the tile's far-field segment in the catalog, -1 for the others
  // workspace
  double* d_prep = nullptr;
  double* d_summary = nullptr;
  int32_t levels_per_batch = 0;
  int* d_flags = nullptr;
  // outputs
  double* d_K = nullptr;  // [np][k_pitch][7]
  double* d_I = nullptr;  // [nf][4]
  // Jacobian targets (nq > 0)
  double* d_dK = nullptr;    // [np][nq][k_pitch][7]
  double* d_dI = nullptr;    // [nf][np][nq][4]
  double* d_Ilev = nullptr;  // [np][nf][4] radiance arriving at each level
  double* d_jac = nullptr;   // [levels_per_batch][ntiles][nq][2][TL][4]
  double* d_jcom = nullptr;  // [levels_per_batch][ntiles][TL]
  double *d_dQdT = nullptr, *d_dr = nullptr, *d_invT = nullptr, *d_ffac = nullptr, *d_wjac = nullptr, *d_dnpm = nullptr, *d_magr = nullptr;
  int32_t tg_kind[AB200_MAX_TARGETS] = {0}, tg_species[AB200_MAX_TARGETS] = {0};
  int64_t tg_line[AB200_MAX_TARGETS] = {0};
  int32_t tg_ls_var[AB200_MAX_TARGETS] = {0}, tg_coeff[AB200_MAX_TARGETS] = {0};
  int32_t it = -1;  // position of the temperature target
  bool dk_preloaded = false;

  // per-kernel timing (ab200_path_set_timing)
  bool timing = false;
  struct Pending { int cls; cudaEvent_t e0, e1; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> free_events;
  double t_ms[4] = {0, 0, 0, 0};
  int64_t t_n[4] = {0, 0, 0, 0};
  std::vector<double> grid_bounds;  // [np][2] bounds of the whole (unsharded) grid, empty = use the uploaded grid's
  int64_t f_stride = 0;
  int32_t rte_option = AB200_RTE_LINSRC, no_neg = 1, select_species = AB200_SPECIES_BATH;
  uint32_t flags = 0;
  cudaEvent_t ev_staged = nullptr;  // the pinned staging blocks may be refilled once this has completed
  bool uploaded = false, k_preloaded = false;
  bool stage2_only = false;  // no line-sum workspace: K comes from other workspaces (ab200::path_adopt_K)
  bool k_adopted = false;    // K was copied in from workspaces that ran the line sum for this catalog and species selection
  bool k_only_A = false;  // K / dK were cleared and filled by a term that only touches A (lookup tables): scalar Stokes path

  // observer epilogue (ab200_path_run_observer)
  struct DevBuf {  // grow-only device array
    void* p = nullptr;
    size_t bytes = 0;
    int reserve(size_t n) {
      if (n <= bytes) return 0;
      cudaFree(p);
      p = nullptr; bytes = 0;
      if (cudaMalloc(&p, n) != cudaSuccess) { cudaGetLastError(); return 1; }
      bytes = n;
      return 0;
    }
    ~DevBuf() { cudaFree(p); }
  };
  DevBuf o_Jx, o_y, o_Jy, o_map_offset, o_map_x, o_map_w, o_bkg_x, o_bkg_w, o_w_offset, o_w_freq, o_w_stokes;
  int32_t o_nx = 0, o_nch = 0;
  bool o_ran = false, o_has_jx = false;

  ~ab200_path() {
    cudaFree(d_f); cudaFree(d_small); cudaFree(d_Ibkg); cudaFree(d_segs); cudaFree(d_prep); cudaFree(d_summary); cudaFree(fmm.L0); cudaFree(d_tile_seg);
    cudaFree(d_flags); cudaFree(d_K); cudaFree(d_I);
    cudaFree(d_dK); cudaFree(d_dI); cudaFree(d_Ilev); cudaFree(d_jac); cudaFree(d_jcom);
    if (h_small) cudaFreeHost(h_small);
    if (h_segs) cudaFreeHost(h_segs);
    if (ev_staged) cudaEventDestroy(ev_staged);
    for (auto& q : pending) { cudaEventDestroy(q.e0); cudaEventDestroy(q.e1); }
    for (auto e : free_events) cudaEventDestroy(e);
    if (own_stream && stream) cudaStreamDestroy(stream);
  }
};

extern "C" {

const char* ab200_last_error(void) { return g_err.c_str(); }

int ab200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int ab200_set_device(int device) {
  AB_CUDA(cudaSetDevice(device));
  return AB200_OK;
}

int64_t ab200_launch_count(int reset) {
  const int64_t n = g_launches;
  if (reset) g_launches = 0;
  return n;
}

// ---------------------------------------------------------------------------
// path workspace
// ---------------------------------------------------------------------------
int ab200_path_create(const ab200_catalog* cat, int64_t nf, int32_t np, int32_t nq, ab200_path** out) {
  return ab200::path_create_ex(cat, nf, np, nq, false, out);
}

// stage2_only: a workspace that only ever runs the Stokes chain on a K handed over by other workspaces (multi.cu's
// level-sharded split): no line records, cluster moments or Jacobian scratch are allocated.
extern "C++" int ab200::path_create_ex(const ab200_catalog* cat, int64_t nf, int32_t np, int32_t nq, bool stage2_only, ab200_path** out) {
  if (!cat || !out) return set_error(AB200_ERR_INVALID, "ab200_path_create: null argument");
  *out = nullptr;
  if (nf < 0 || np < 0 || nq < 0) return set_error(AB200_ERR_INVALID, "ab200_path_create: negative size");
  if (nq > AB200_MAX_TARGETS)
    return set_error(AB200_ERR_UNSUPPORTED, "at most " + std::to_string(AB200_MAX_TARGETS) + " Jacobian targets per call");
  std::unique_ptr<ab200_path> p(new ab200_path());
  p->cat = cat;
  p->nf = nf;
  p->np = np;
  p->np_cap = np;
  p->nq = nq;
  p->k_pitch = (nf + 127) / 128 * 128;
  AB_CUDA(cudaSetDevice(cat->device));
  AB_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  p->own_stream = true;

  const size_t snp = static_cast<size_t>(np);
  AB_TRY(dev_alloc(&p->d_f, snp * nf));
  // packed small arrays: T, P, H [np] | vmr [np][ns] | isorat, Q [np][ni] | npm [np][4][7] | frange [np][2] | r [np]
  //                      | dQdT [np][ni] | dr [2][np][nq] | 1/T [np] | wind factor [np] | freq_wind_shift_jac [np][3]
  //                      | dnorm_view [np][3][4][7] | mag / |mag| [np][3]
  p->small_doubles = snp * (3 + cat->n_species + 2 * cat->n_isot + 28 + 2 + 1 + cat->n_isot + 2 * static_cast<size_t>(nq) + 2 + 3 + 84 + 3);
  AB_TRY(dev_alloc(&p->d_small, p->small_doubles));
  if (p->small_doubles) AB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&p->h_small), p->small_doubles * sizeof(double)));
  double* q = p->d_small;
  p->d_T = q; q += snp;
  p->d_P = q; q += snp;
  p->d_H = q; q += snp;
  p->d_vmr = q; q += snp * cat->n_species;
  p->d_isorat = q; q += snp * cat->n_isot;
  p->d_Q = q; q += snp * cat->n_isot;
  p->d_npm = q; q += snp * 28;
  p->d_frange = q; q += snp * 2;
  p->d_r = q; q += snp;
  p->d_dQdT = q; q += snp * cat->n_isot;
  p->d_dr = q; q += 2 * snp * nq;
  p->d_invT = q; q += snp;
  p->d_ffac = q; q += snp;
  p->d_wjac = q; q += 3 * snp;
  p->d_dnpm = q; q += 84 * snp;
  p->d_magr = q;
  AB_TRY(dev_alloc(&p->d_Ibkg, static_cast<size_t>(nf) * 4));
  AB_TRY(dev_alloc(&p->d_I, static_cast<size_t>(nf) * 4));
  AB_TRY(dev_alloc(&p->d_K, snp * p->k_pitch * 7));
  const size_t nseg = cat->segments.size();
  AB_TRY(dev_alloc(&p->d_segs, 3 * nseg));
  if (nseg) AB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&p->h_segs), 3 * nseg * sizeof(SegmentDev)));
  AB_TRY(dev_alloc(&p->d_flags, 1));
  AB_CUDA(cudaMemset(p->d_flags, 0, sizeof(int)));

  const size_t per_level = static_cast<size_t>(cat->ntiles) * (tile_doubles() + static_cast<size_t>(nq) * 2 * TL * 4 + (nq ? TL : 0)) * sizeof(double);
  int lpb = np;
  if (per_level > 0) lpb = static_cast<int>(std::max<size_t>(1, std::min<size_t>(snp, PREP_BUDGET_BYTES / per_level)));
  p->levels_per_batch = std::max(lpb, 1);
  p->stage2_only = stage2_only;
  if (stage2_only) {
    *out = p.release();
    return AB200_OK;
  }
  AB_TRY(dev_alloc(&p->d_prep, static_cast<size_t>(p->levels_per_batch) * cat->ntiles * tile_doubles()));
  AB_TRY(dev_alloc(&p->d_summary, static_cast<size_t>(p->levels_per_batch) * cat->ntiles * SUMMARY_DOUBLES));
  {
    // far-field sums of the real segments (lbl_fmm.cu): moment records of four cluster levels + scratch
    size_t nff = 0;
    std::vector<int32_t> tseg(static_cast<size_t>(cat->ntiles), -1);
    for (size_t i = 0; i < cat->segments.size(); i++) {
      const Segment& sg = cat->segments[i];
      if (sg.mode != 0) continue;
      nff++;
      for (int64_t t = sg.tile_begin; t < sg.tile_end; t++) tseg[t] = static_cast<int32_t>(i);
    }
    if (nff > 0 && cat->ntiles > 0) {
      const size_t L = static_cast<size_t>(p->levels_per_batch), nt = static_cast<size_t>(cat->ntiles);
      p->fmm.ngroups = (cat->ntiles + FMM_GROUP - 1) / FMM_GROUP;
      const size_t n0 = L * nt * 16 * MOM_DOUBLES, n1 = L * nt * 4 * MOM_DOUBLES, n2 = L * nt * MOM_DOUBLES,
                   n3 = L * static_cast<size_t>(p->fmm.ngroups) * MOM_DOUBLES, nsc = L * nt * 2,
                   nacc = nff * L * static_cast<size_t>(p->k_pitch);
      AB_TRY(dev_alloc(&p->fmm.L0, n0 + n1 + n2 + n3 + nsc + nacc));
      p->fmm.L1 = p->fmm.L0 + n0;
      p->fmm.L2 = p->fmm.L1 + n1;
      p->fmm.L3 = p->fmm.L2 + n2;
      p->fmm.scan = p->fmm.L3 + n3;
      p->fmm.far_acc = p->fmm.scan + nsc;
      AB_TRY(dev_alloc(&p->d_tile_seg, nt));
      AB_CUDA(cudaMemcpy(p->d_tile_seg, tseg.data(), nt * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
  }
  if (nq > 0) {
    AB_TRY(dev_alloc(&p->d_jac, static_cast<size_t>(p->levels_per_batch) * cat->ntiles * nq * 2 * TL * 4));
    AB_TRY(dev_alloc(&p->d_jcom, static_cast<size_t>(p->levels_per_batch) * cat->ntiles * TL));
    AB_TRY(dev_alloc(&p->d_dK, snp * nq * p->k_pitch * 7));
    AB_TRY(dev_alloc(&p->d_dI, static_cast<size_t>(nf) * snp * nq * 4));
    AB_TRY(dev_alloc(&p->d_Ilev, snp * nf * 4));
  }
  *out = p.release();
  return AB200_OK;
}

int ab200_path_create_stage2(const ab200_catalog* cat, int64_t nf, int32_t np, ab200_path** out) {
  return ab200::path_create_ex(cat, nf, np, 0, true, out);
}
int ab200_path_adopt_K(ab200_path* p) { return ab200::path_adopt_K(p); }

void ab200_path_destroy(ab200_path* p) { delete p; }

int ab200_path_set_stream(ab200_path* p, void* stream) {
  if (!p) return set_error(AB200_ERR_INVALID, "ab200_path_set_stream: null path");
  if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
  p->stream = static_cast<cudaStream_t>(stream);
  p->own_stream = false;
  return AB200_OK;
}

int ab200_path_set_grid_bounds(ab200_path* p, const double* bounds) {
  if (!p) return set_error(AB200_ERR_INVALID, "ab200_path_set_grid_bounds: null path");
  if (!bounds) {
    p->grid_bounds.clear();
    return AB200_OK;
  }
  for (int ip = 0; ip < p->np; ip++)
    if (!(bounds[2 * ip] <= bounds[2 * ip + 1]))
      return set_error(AB200_ERR_INVALID, "ab200_path_set_grid_bounds: bounds must be ascending and finite");
  p->grid_bounds.assign(bounds, bounds + 2 * static_cast<size_t>(p->np));
  return AB200_OK;
}

int ab200_path_upload(ab200_path* p, const double* f, int64_t f_level_stride, const ab200_atm_path* atm,
                      int32_t select_species, int32_t no_negative_absorption, const ab200_target* targets,
                      const double* r, int32_t hse_derivative, int32_t rte_option, const double* I_bkg,
                      uint32_t flags) {
  if (!p || !f || !atm) return set_error(AB200_ERR_INVALID, "ab200_path_upload: null argument");
  const ab200_catalog* cat = p->cat;
  if (atm->np < 0 || atm->np > p->np_cap)
    return set_error(AB200_ERR_INVALID, "atm path has " + std::to_string(atm->np) + " levels, the workspace was created for " +
                                            std::to_string(p->np_cap));
  p->np = atm->np;  // ragged batches: a workspace created for the longest path takes every shorter one
  const int np = p->np;
  if (f_level_stride != 0 && f_level_stride != p->nf)
    return set_error(AB200_ERR_INVALID, "f_level_stride must be 0 (shared grid) or nf (one grid per level)");
  if (rte_option != AB200_RTE_CONSTANT && rte_option != AB200_RTE_LINSRC && rte_option != AB200_RTE_LINPROP)
    return set_error(AB200_ERR_INVALID, "unknown rte_option");
  if (select_species != AB200_SPECIES_BATH && (select_species < 0 || select_species >= cat->n_species))
    return set_error(AB200_ERR_INVALID, "select_species out of range");
  if (!atm->T || !atm->P || !atm->vmr || !atm->isorat || !atm->Q)
    return set_error(AB200_ERR_INVALID, "atm path: T, P, vmr, isorat and Q are required");
  AB_CUDA(cudaSetDevice(cat->device));
  // the previous upload's asynchronous copies out of the pinned staging blocks must have run before they are refilled
  if (p->ev_staged) AB_CUDA(cudaEventSynchronize(p->ev_staged));
  else AB_CUDA(cudaEventCreateWithFlags(&p->ev_staged, cudaEventDisableTiming));
  p->it = -1;
  if (p->nq > 0) {
    if (!targets) return set_error(AB200_ERR_INVALID, "ab200_path_upload: targets is null with nq > 0");
    if (flags & AB200_FLAG_TRAN_EXACT)
      return set_error(AB200_ERR_UNSUPPORTED, "AB200_FLAG_TRAN_EXACT has no Jacobian (the reference differentiates its literal form)");
    for (int q = 0; q < p->nq; q++) {
      p->tg_kind[q]    = targets[q].kind;
      p->tg_species[q] = targets[q].species;
      if (targets[q].kind == AB200_TARGET_T) {
        if (!atm->dQdT) return set_error(AB200_ERR_INVALID, "atm path: dQdT is required with a temperature target");
        if (p->it < 0) p->it = q;
      } else if (targets[q].kind == AB200_TARGET_VMR) {
        if (targets[q].species < 0 || targets[q].species >= cat->n_species)
          return set_error(AB200_ERR_INVALID, "Jacobian target " + std::to_string(q) + ": species out of range");
      } else if (targets[q].kind == AB200_TARGET_WIND_U || targets[q].kind == AB200_TARGET_WIND_V ||
                 targets[q].kind == AB200_TARGET_WIND_W) {
        // frequency derivative of the lines times f * freq_wind_shift_jac; nothing to check
      } else if (targets[q].kind == AB200_TARGET_MAG_U || targets[q].kind == AB200_TARGET_MAG_V ||
                 targets[q].kind == AB200_TARGET_MAG_W) {
        // Zeeman splitting derivative and the derivative of the polarisation matrix; nothing to check
      } else if (targets[q].kind >= AB200_TARGET_LINE_F0 && targets[q].kind <= AB200_TARGET_LINE_LS) {
        const ab200_target& t = targets[q];
        if (t.line < 0 || t.line >= cat->n_lines)
          return set_error(AB200_ERR_INVALID, "Jacobian target " + std::to_string(q) + ": line out of range");
        if (!cat->line_target_ok[static_cast<size_t>(t.line)])
          return set_error(AB200_ERR_UNSUPPORTED, "Jacobian target " + std::to_string(q) +
                                                      ": line targets need a VP_LTE band without cutoff");
        if (t.kind == AB200_TARGET_LINE_LS) {
          if (t.ls_var < 0 || t.ls_var >= AB200_NVAR || t.coeff < 0 || t.coeff > 3)
            return set_error(AB200_ERR_INVALID, "Jacobian target " + std::to_string(q) + ": line-shape variable or coefficient out of range");
          if (t.species != AB200_SPECIES_BATH && (t.species < 0 || t.species >= cat->n_species))
            return set_error(AB200_ERR_INVALID, "Jacobian target " + std::to_string(q) + ": broadener species out of range");
        }
        p->tg_line[q] = t.line; p->tg_ls_var[q] = t.ls_var; p->tg_coeff[q] = t.coeff;
      } else if (targets[q].kind == AB200_TARGET_ISORAT) {
        if (targets[q].species < 0 || targets[q].species >= cat->n_isot)
          return set_error(AB200_ERR_INVALID, "Jacobian target " + std::to_string(q) + ": isotopologue out of range");
        for (int ip = 0; ip < np; ip++)
          if (atm->isorat[static_cast<size_t>(ip) * cat->n_isot + targets[q].species] == 0)
            return set_error(AB200_ERR_INVALID, "Does not support 0 for isotopologue ratios");  // :1539
      } else if (targets[q].kind == AB200_TARGET_P) {
        return set_error(AB200_ERR_UNSUPPORTED, "Not implemented, pressure derivative");  // lbl_lineshape_voigt_lte.cpp:1482
      } else {
        return set_error(AB200_ERR_UNSUPPORTED, "Jacobian target " + std::to_string(q) +
                                                    ": unknown target kind");
      }
    }
  }

  const size_t snp = static_cast<size_t>(np);
  const size_t scap = static_cast<size_t>(p->np_cap);  // the packed arrays are laid out for the capacity
  double* h = p->h_small;
  double *hT = h, *hP = hT + scap, *hH = hP + scap, *hv = hH + scap, *hi = hv + scap * cat->n_species,
         *hQ = hi + scap * cat->n_isot, *hn = hQ + scap * cat->n_isot, *hfr = hn + scap * 28, *hr = hfr + scap * 2,
         *hdQ = hr + scap, *hdr = hdQ + scap * cat->n_isot, *hiT = hdr + 2 * scap * p->nq, *hff = hiT + scap, *hwj = hff + scap, *hdn = hwj + 3 * scap,
         *hmr = hdn + 84 * scap;
  std::fill(hdr, hdr + 2 * scap * p->nq, 0.0);
  for (int ip = 0; ip < np; ip++) {
    if (!(atm->T[ip] > 0) || !(atm->P[ip] >= 0))
      return set_error(AB200_ERR_INVALID, "level " + std::to_string(ip) + ": temperature must be > 0 and pressure >= 0");
    hT[ip] = atm->T[ip];
    hP[ip] = atm->P[ip];
    hiT[ip] = 1.0 / atm->T[ip];
    double mag[3] = {0, 0, 0}, los[2] = {0, 0};
    if (atm->mag) std::copy(atm->mag + 3 * ip, atm->mag + 3 * ip + 3, mag);
    if (atm->los) std::copy(atm->los + 2 * ip, atm->los + 2 * ip + 2, los);
    hH[ip] = std::hypot(mag[0], mag[1], mag[2]);
    double fac = 1.0;
    const double calm[3] = {0, 0, 0};  // freq_grid_pathFromPath runs wind_shift at every point, windless ones included
    if (!wind_factor(atm->wind ? atm->wind + 3 * ip : calm, los, &fac, hwj + 3 * ip))
      return set_error(AB200_ERR_INVALID, "level " + std::to_string(ip) + ": Negative frequency scaling factor (wind_shift)");
    if (!atm->wind) fac = 1.0;
    hff[ip] = fac;
    for (int pol = 0; pol < 4; pol++) norm_view(pol, mag, los, hn + (static_cast<size_t>(ip) * 4 + pol) * 7);
    for (int c = 0; c < 3; c++) {
      for (int pol = 0; pol < 4; pol++) dnorm_view(pol, c, mag, los, hdn + ((static_cast<size_t>(ip) * 3 + c) * 4 + pol) * 7);
      hmr[3 * ip + c] = mag[c] / hH[ip];  // dH/dmag_c, lbl_lineshape_voigt_lte.cpp:1071-1072 (NaN without a field, like there)
    }
    const double* fl = f + ip * f_level_stride;
    if (!p->grid_bounds.empty()) {
      hfr[2 * ip]     = fac * p->grid_bounds[2 * ip];
      hfr[2 * ip + 1] = fac * p->grid_bounds[2 * ip + 1];
    } else {
      hfr[2 * ip]     = p->nf ? fac * fl[0] : 0.0;
      hfr[2 * ip + 1] = p->nf ? fac * fl[p->nf - 1] : 0.0;
    }
    hr[ip] = (r && ip < np - 1) ? r[ip] : 0.0;
    for (int s = 0; s < cat->n_species; s++) {
      const double v = atm->vmr[static_cast<size_t>(ip) * cat->n_species + s];
      if (!(v >= 0)) return set_error(AB200_ERR_INVALID, "level " + std::to_string(ip) + ": negative or NaN VMR");
      hv[static_cast<size_t>(ip) * cat->n_species + s] = v;
    }
    for (int i = 0; i < cat->n_isot; i++) {
      const double ir = atm->isorat[static_cast<size_t>(ip) * cat->n_isot + i];
      const double Q  = atm->Q[static_cast<size_t>(ip) * cat->n_isot + i];
      if (!(ir >= 0) || !(Q > 0))
        return set_error(AB200_ERR_INVALID,
                         "level " + std::to_string(ip) + ": isotopologue ratio must be >= 0 and partition function > 0");
      hi[static_cast<size_t>(ip) * cat->n_isot + i] = ir;
      hQ[static_cast<size_t>(ip) * cat->n_isot + i] = Q;
      hdQ[static_cast<size_t>(ip) * cat->n_isot + i] = atm->dQdT ? atm->dQdT[static_cast<size_t>(ip) * cat->n_isot + i] : 0.0;
    }
    // hydrostatic path-length derivative, m_tramat.cc:18-24: dr [2][np-1][nq]
    if (hse_derivative && p->it >= 0 && r && ip < np - 1) {
      hdr[static_cast<size_t>(ip) * p->nq + p->it]            = r[ip] / (2.0 * atm->T[ip]);
      hdr[(snp - 1 + ip) * p->nq + p->it]                     = r[ip] / (2.0 * atm->T[ip + 1]);
    }
  }
  if (p->small_doubles)
    AB_CUDA(cudaMemcpyAsync(p->d_small, h, p->small_doubles * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  const size_t nfl = static_cast<size_t>(p->nf) * (f_level_stride ? np : 1);
  if (nfl) AB_CUDA(cudaMemcpyAsync(p->d_f, f, nfl * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  if (I_bkg && p->nf)
    AB_CUDA(cudaMemcpyAsync(p->d_Ibkg, I_bkg, static_cast<size_t>(p->nf) * 4 * sizeof(double), cudaMemcpyHostToDevice, p->stream));

  // segments selected by species (lbl_lineshape.cpp:191), split by kernel
  const size_t nseg = cat->segments.size();
  // AB200_FARFIELD=0 routes every real segment through the line-by-line kernel (A/B measurements, tests); read per upload
  const int farfield_mode = [] { const char* e = getenv("AB200_FARFIELD"); return e ? atoi(e) : 1; }();
  const bool farfield = farfield_mode != 0;
  p->nsegs[0] = p->nsegs[1] = p->nsegs[2] = 0;
  for (const Segment& s : cat->segments) {
    if (!(select_species == AB200_SPECIES_BATH || select_species == s.species)) continue;
    SegmentDev d{s.tile_begin, s.tile_end, s.cutoff, s.pol, s.has_cutoff};
    // Far-field sums cost ~5e-10 s per (frequency, level, segment) almost whatever the segment's size, plus 2.8e-10 s per
    // (line, level) for the records and moments; the line-by-line kernel costs 6e-13 s per pair: they win from ~800 lines per
    // segment (measured on configs[1] shapes, 1e5 frequencies x 100 levels, line by line -> far field per step: 1000 lines per
    // species 25.3 -> 21.3 ms, 2000: 48.8 -> 24.5, 2e4 (configs[1]): 472.9 -> 49.4; one configs[4] path, 1e4 frequencies,
    // 2000 lines per species: 5.9 -> 3.0 ms; 500 lines per species: 13.7 -> 18.2 ms, stays line by line).  AB200_FARFIELD=2
    // forces them, = 0 switches them off.
    const bool big = s.nsub >= FMM_MIN_LINES || farfield_mode == 2;
    const int list = s.mode == 1 ? 1 : (farfield && big && p->fmm.L0) ? 2 : 0;
    p->h_segs[list * nseg + p->nsegs[list]++] = d;
  }
  if (nseg) AB_CUDA(cudaMemcpyAsync(p->d_segs, p->h_segs, 3 * nseg * sizeof(SegmentDev), cudaMemcpyHostToDevice, p->stream));

  AB_CUDA(cudaEventRecord(p->ev_staged, p->stream));
  p->f_stride   = f_level_stride;
  p->rte_option = rte_option;
  p->no_neg     = no_negative_absorption;
  p->select_species = select_species;
  p->flags      = flags;
  p->uploaded   = true;
  p->k_preloaded = false;
  p->dk_preloaded = false;
  p->k_only_A = false;
  p->k_adopted = false;
  return AB200_OK;
}

// ---- internal hooks for multi.cu (level-sharded line sum, frequency-sharded Stokes chain) ----
extern "C++" cudaStream_t ab200::path_stream(ab200_path* p) { return p->stream; }
extern "C++" double* ab200::path_K(ab200_path* p, int64_t* k_pitch) {
  *k_pitch = p->k_pitch;
  return p->d_K;
}
// K [np][k_pitch][7] of an uploaded workspace has been filled with rows that other workspaces of the same catalog, species
// selection and flags summed (copies queued on this workspace's stream): the Stokes chain runs as if it had summed them.
extern "C++" int ab200::path_adopt_K(ab200_path* p) {
  if (!p || !p->uploaded) return set_error(AB200_ERR_INVALID, "path_adopt_K: path not uploaded");
  p->k_preloaded = false;
  p->k_adopted = true;
  return AB200_OK;
}

namespace {
// RAII-free scoped timer: records an event pair around one launch when timing is on
struct LaunchTimer {
  ab200_path* p;
  int cls;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  LaunchTimer(ab200_path* p_, int cls_) : p(p_), cls(cls_) {
    if (!p->timing) return;
    auto get = [&]() {
      cudaEvent_t e = nullptr;
      if (!p->free_events.empty()) { e = p->free_events.back(); p->free_events.pop_back(); }
      else cudaEventCreate(&e);
      return e;
    };
    e0 = get();
    e1 = get();
    cudaEventRecord(e0, p->stream);
  }
  void stop() {
    if (!e0) return;
    cudaEventRecord(e1, p->stream);
    p->pending.push_back({cls, e0, e1});
    e0 = nullptr;
  }
};

void fill_params(const ab200_path* p, int lev0, PrepareParams& pp, SumParams& sp) {
  const ab200_catalog* cat = p->cat;
  pp = PrepareParams{};
  pp.f0 = cat->d_f0; pp.a = cat->d_a; pp.e0 = cat->d_e0; pp.gu = cat->d_gu; pp.T0 = cat->d_T0;
  pp.line_isot = cat->d_line_isot; pp.ls_offset = cat->d_ls_offset; pp.ls_species = cat->d_ls_species;
  pp.ls_type = cat->d_ls_type; pp.ls_X = cat->d_ls_X; pp.isot_species = cat->d_isot_species;
  pp.isot_mass = cat->d_isot_mass; pp.sub_parent = cat->d_sub_parent; pp.sub_Sz = cat->d_sub_Sz;
  pp.sub_dzc = cat->d_sub_dzc; pp.sub_cut = cat->d_sub_cut; pp.tile_mode = cat->d_tile_mode; pp.sub_flags = cat->d_sub_flags;
  pp.n_species = cat->n_species; pp.n_isot = cat->n_isot; pp.ntiles = cat->ntiles;
  pp.T = p->d_T + lev0; pp.P = p->d_P + lev0; pp.H = p->d_H + lev0;
  pp.vmr = p->d_vmr + static_cast<size_t>(lev0) * cat->n_species;
  pp.isorat = p->d_isorat + static_cast<size_t>(lev0) * cat->n_isot;
  pp.Q = p->d_Q + static_cast<size_t>(lev0) * cat->n_isot;
  pp.frange = p->d_frange + 2 * static_cast<size_t>(lev0);
  pp.prep = p->d_prep; pp.summary = p->d_summary; pp.flags = p->d_flags;

  sp = SumParams{};
  sp.f = p->d_f + static_cast<size_t>(lev0) * p->f_stride;
  sp.f_stride = p->f_stride; sp.nf = p->nf; sp.k_pitch = p->k_pitch;
  sp.ffac = p->d_ffac + lev0;
  sp.T = p->d_T + lev0; sp.P = p->d_P + lev0;
  sp.npm = p->d_npm + static_cast<size_t>(lev0) * 28;
  sp.prep = p->d_prep; sp.summary = p->d_summary; sp.tile_count = cat->d_tile_count; sp.tile_mode = cat->d_tile_mode; sp.ntiles = cat->ntiles;
  sp.no_negative_absorption = p->no_neg;
  sp.K = p->d_K + static_cast<size_t>(lev0) * p->k_pitch * 7;
}
}  // namespace

int ab200_path_run_propmat(ab200_path* p) {
  if (!p || !p->uploaded) return set_error(AB200_ERR_INVALID, "ab200_path_run_propmat: path not uploaded");
  if (p->stage2_only) return set_error(AB200_ERR_INVALID, "ab200_path_run_propmat: this workspace has no line-sum buffers");
  const ab200_catalog* cat = p->cat;
  AB_CUDA(cudaSetDevice(cat->device));
  p->k_only_A = false;
  const size_t kbytes = static_cast<size_t>(p->np) * p->k_pitch * 7 * sizeof(double);
  // With mode-0 (real) segments selected the real line sum writes whole K records itself (vector stores);
  // only without them, or when the caller's K is accumulated into, K is zeroed / kept and updated in place.
  const bool store_full = !p->k_preloaded && p->nsegs[0] + p->nsegs[2] > 0 && cat->ntiles > 0 && p->nf > 0;
  if (!p->k_preloaded && !store_full && kbytes) AB_CUDA(cudaMemsetAsync(p->d_K, 0, kbytes, p->stream));
  if (p->nq > 0 && !p->dk_preloaded && kbytes) AB_CUDA(cudaMemsetAsync(p->d_dK, 0, kbytes * p->nq, p->stream));
  if (cat->ntiles == 0 || p->nf == 0) return AB200_OK;
  const size_t nseg = cat->segments.size();
  for (int lev0 = 0; lev0 < p->np; lev0 += p->levels_per_batch) {
    const int nlev = std::min(p->levels_per_batch, p->np - lev0);
    PrepareParams pp;
    SumParams sp;
    fill_params(p, lev0, pp, sp);
    {
      LaunchTimer t(p, 0);
      AB_TRY(launch_prepare(pp, nlev, p->stream));
      t.stop();
    }
    // real segments: the line-by-line kernel (AB200_FARFIELD=0) writes whole K records when it runs; the far-field sums
    // add to them, or write the records themselves when they are alone
    sp.k_store_full = (store_full && p->nsegs[0] > 0) ? 1 : 0;
    for (int mode = 0; mode < 2; mode++) {
      if (mode == 1 && p->nsegs[2] > 0) {
        SumParams sf = sp;
        sf.segs = p->d_segs + 2 * nseg;
        sf.nsegs = p->nsegs[2];
        LaunchTimer t(p, 1);
        AB_TRY(launch_fmm(pp, sf, p->fmm, p->d_tile_seg, nlev, (store_full && p->nsegs[0] == 0) ? 1 : 0, p->stream));
        t.stop();
      }
      sp.segs = p->d_segs + mode * nseg;
      sp.nsegs = p->nsegs[mode];
      if (sp.nsegs == 0) continue;
      LaunchTimer t(p, 1 + mode);
      AB_TRY(launch_sum(sp, nlev, mode, p->stream));
      t.stop();
    }
    if (p->nq > 0) {
      JacPrepParams jp{};
      JacSumParams js{};
      // computed targets: the three magnetic-field components share one derivative record, the three wind components another
      int nc = 0, c_mag = -1, c_wind = -1;
      for (int q = 0; q < p->nq; q++) {
        const int kind = p->tg_kind[q];
        const bool mag = kind >= AB200_TARGET_MAG_U && kind <= AB200_TARGET_MAG_W;
        const bool wind = kind >= AB200_TARGET_WIND_U && kind <= AB200_TARGET_WIND_W;
        const int comp = mag ? kind - AB200_TARGET_MAG_U : wind ? kind - AB200_TARGET_WIND_U : 0;
        int c;
        if (mag && c_mag >= 0 && js.out_row[c_mag][comp] < 0) c = c_mag;  // (a component asked for twice gets its own entry)
        else if (wind && c_wind >= 0 && js.out_row[c_wind][comp] < 0) c = c_wind;
        else {
          c = nc++;
          js.out_row[c][0] = js.out_row[c][1] = js.out_row[c][2] = -1;
          jp.kind[c] = js.kind[c] = mag ? AB200_TARGET_MAG_U : wind ? AB200_TARGET_WIND_U : kind;
          jp.species[c] = p->tg_species[q];
          jp.line[c] = p->tg_line[q]; jp.ls_var[c] = p->tg_ls_var[q]; jp.coeff[c] = p->tg_coeff[q];
          if (kind >= AB200_TARGET_LINE_F0 && kind <= AB200_TARGET_LINE_LS)
            std::copy_n(cat->line_tiles.data() + p->tg_line[q] * 8, 8, &js.line_tiles[c][0][0]);
          if (mag) c_mag = c;
          if (wind) c_wind = c;
        }
        js.out_row[c][comp] = q;
      }
      jp.nq = js.nq = nc;
      js.nrows = p->nq;
      jp.dQdT = p->d_dQdT + static_cast<size_t>(lev0) * cat->n_isot;
      jp.jac = p->d_jac;
      jp.jcom = p->d_jcom;
      js.jac = p->d_jac;
      js.jcom = p->d_jcom;
      js.dK = p->d_dK + static_cast<size_t>(lev0) * p->nq * p->k_pitch * 7;
      js.mag_ratio = p->d_magr + 3 * static_cast<size_t>(lev0);
      js.dnpm = p->d_dnpm + 84 * static_cast<size_t>(lev0);
      js.wind_jac = (p->flags & AB200_FLAG_WIND_ROWS_DF) ? nullptr : p->d_wjac + 3 * static_cast<size_t>(lev0);
      AB_TRY(launch_prepare_jac(pp, jp, nlev, p->stream));
      for (int list = 0; list < 3; list++) {  // 0 and 2: real segments (with / without cutoffs), 1: complex
        sp.segs = p->d_segs + list * nseg;
        sp.nsegs = p->nsegs[list];
        if (sp.nsegs == 0) continue;
        js.real_lines = list != 1 ? 1 : 0;
        AB_TRY(launch_sum_jac(sp, js, nlev, p->stream));
      }
    }
  }
  return AB200_OK;
}

int ab200_path_set_timing(ab200_path* p, int on) {
  if (!p) return set_error(AB200_ERR_INVALID, "ab200_path_set_timing: null path");
  p->timing = on != 0;
  return AB200_OK;
}

int ab200_path_get_timings(ab200_path* p, double ms[4], int64_t launches[4]) {
  if (!p || !ms || !launches) return set_error(AB200_ERR_INVALID, "ab200_path_get_timings: null argument");
  AB_CUDA(cudaSetDevice(p->cat->device));
  AB_CUDA(cudaStreamSynchronize(p->stream));
  for (auto& q : p->pending) {
    float t = 0.f;
    AB_CUDA(cudaEventElapsedTime(&t, q.e0, q.e1));
    p->t_ms[q.cls] += t;
    p->t_n[q.cls] += 1;
    p->free_events.push_back(q.e0);
    p->free_events.push_back(q.e1);
  }
  p->pending.clear();
  for (int i = 0; i < 4; i++) {
    ms[i] = p->t_ms[i];
    launches[i] = p->t_n[i];
    p->t_ms[i] = 0;
    p->t_n[i] = 0;
  }
  return AB200_OK;
}

int ab200_path_region_histogram(ab200_path* p, int64_t samples_per_level, uint64_t seed, double out[8]) {
  if (!p || !out) return set_error(AB200_ERR_INVALID, "ab200_path_region_histogram: null argument");
  if (!p->uploaded) return set_error(AB200_ERR_INVALID, "ab200_path_region_histogram: path not uploaded");
  for (int i = 0; i < 8; i++) out[i] = 0.0;
  const ab200_catalog* cat = p->cat;
  if (cat->ntiles == 0 || p->nf == 0 || p->np == 0) return AB200_OK;
  AB_CUDA(cudaSetDevice(cat->device));
  double* d_out = nullptr;
  AB_TRY(dev_alloc(&d_out, 8));
  cudaError_t e = cudaMemsetAsync(d_out, 0, 8 * sizeof(double), p->stream);
  int rc = 0;
  const size_t nseg = cat->segments.size();
  for (int lev0 = 0; lev0 < p->np && !rc && e == cudaSuccess; lev0 += p->levels_per_batch) {
    const int nlev = std::min(p->levels_per_batch, p->np - lev0);
    PrepareParams pp;
    SumParams sp;
    fill_params(p, lev0, pp, sp);
    sp.segs  = p->d_segs + nseg;  // the mode-1 list carries the cutoff windows
    sp.nsegs = p->nsegs[1];
    rc = launch_prepare(pp, nlev, p->stream);
    if (!rc) rc = launch_region_histogram(sp, nlev, samples_per_level, seed + 7919ull * lev0, d_out, p->stream);
  }
  if (!rc && e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, 8 * sizeof(double), cudaMemcpyDeviceToHost, p->stream);
  if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
  cudaFree(d_out);
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "ab200_path_region_histogram", __FILE__, __LINE__);
  return AB200_OK;
}

static int check_flags(ab200_path* p);

static int run_stokes_impl(ab200_path* p, const ab200_observer* obs) {
  if (!p || !p->uploaded) return set_error(AB200_ERR_INVALID, "ab200_path_run_stokes: path not uploaded");
  if (p->np < 1) return set_error(AB200_ERR_INVALID, "ab200_path_run_stokes: empty path");
  AB_CUDA(cudaSetDevice(p->cat->device));
  StokesParams sp{};
  sp.np = p->np; sp.nf = p->nf; sp.K = p->d_K; sp.k_pitch = p->k_pitch; sp.f = p->d_f; sp.f_stride = p->f_stride;
  sp.T = p->d_T; sp.invT = p->d_invT; sp.ffac = p->d_ffac; sp.r = p->d_r; sp.I_bkg = p->d_Ibkg; sp.I = p->d_I; sp.rte_option = p->rte_option;
  sp.tran_exact = (p->flags & AB200_FLAG_TRAN_EXACT) ? 1 : 0;
  sp.I_lev = p->nq > 0 ? p->d_Ilev : nullptr;
  sp.no_emission = (p->flags & AB200_FLAG_NO_EMISSION) ? 1 : 0;
  sp.flags = p->d_flags;
  sp.scalar = ((p->nsegs[1] == 0 && !p->k_preloaded) || p->k_only_A) ? 1 : 0;  // only mode-0 (real, pol = no) segments wrote K
  {
    LaunchTimer t(p, 3);
    AB_TRY(launch_stokes_chain(sp, p->stream));
    t.stop();
  }
  if (p->nq > 0 || (obs && obs->n_bkg > 0)) {
    StokesJacParams jp{};
    jp.np = p->np; jp.nq = p->nq; jp.nf = p->nf; jp.K = p->d_K; jp.dK = p->d_dK; jp.k_pitch = p->k_pitch;
    jp.f = p->d_f; jp.f_stride = p->f_stride; jp.ffac = p->d_ffac; jp.T = p->d_T; jp.r = p->d_r; jp.dr = p->d_dr; jp.I_lev = p->d_Ilev;
    jp.dI = p->d_dI; jp.it = p->it; jp.rte_option = p->rte_option; jp.flags = p->d_flags;
    jp.no_emission = (p->flags & AB200_FLAG_NO_EMISSION) ? 1 : 0;
    jp.scalar = ((p->nsegs[1] == 0 && !p->k_preloaded && !p->dk_preloaded) || p->k_only_A) ? 1 : 0;
    if (obs) {  // x-space accumulation inside the pass; the per-level dI is not written
      jp.dI = nullptr;
      jp.Jx = static_cast<double*>(p->o_Jx.p);
      jp.map_offset = static_cast<const int64_t*>(p->o_map_offset.p);
      jp.map_x = static_cast<const int32_t*>(p->o_map_x.p);
      jp.map_w = static_cast<const double*>(p->o_map_w.p);
      jp.n_bkg = obs->bkg_kind == AB200_BKG_PLANCK ? obs->n_bkg : 0;
      jp.bkg_x = static_cast<const int32_t*>(p->o_bkg_x.p);
      jp.bkg_w = static_cast<const double*>(p->o_bkg_w.p);
      jp.bkg_T = obs->bkg_T;
    }
    AB_TRY(launch_stokes_jac(jp, p->stream));
  }
  return AB200_OK;
}

int ab200_path_run_stokes(ab200_path* p) { return run_stokes_impl(p, nullptr); }

int ab200_path_add_lookup(ab200_path* p, const ab200_lookup* lut, int32_t h2o_species, const double* target_d, int32_t po, int32_t to,
                          int32_t wo, int32_t fo, double extpolfac, int32_t zero_init) {
  if (!p || !p->uploaded) return set_error(AB200_ERR_INVALID, "ab200_path_add_lookup: path not uploaded");
  if (!lut) return set_error(AB200_ERR_INVALID, "ab200_path_add_lookup: null lookup data");
  if (lut_device(lut) != p->cat->device) return set_error(AB200_ERR_INVALID, "ab200_path_add_lookup: lookup data and path live on different devices");
  if (p->nq > 0 && !target_d) return set_error(AB200_ERR_INVALID, "ab200_path_add_lookup: target_d is null with Jacobian targets");
  AB_TRY(lut_check_call(lut, p->cat->n_species, h2o_species, p->select_species, po, to, wo, fo));
  LutParams lp{};
  for (int q = 0; q < p->nq; q++) {
    if (!std::isnormal(target_d[q]))
      return set_error(AB200_ERR_INVALID, "The target " + std::to_string(q) + " is not good, it lacks a perturbation value.");
    lp.tg_kind[q] = p->tg_kind[q]; lp.tg_species[q] = p->tg_species[q]; lp.tg_d[q] = target_d[q];
  }
  AB_CUDA(cudaSetDevice(p->cat->device));
  if (zero_init) {
    AB_CUDA(cudaMemsetAsync(p->d_K, 0, static_cast<size_t>(p->np) * p->k_pitch * 7 * sizeof(double), p->stream));
    if (p->nq > 0)
      AB_CUDA(cudaMemsetAsync(p->d_dK, 0, static_cast<size_t>(p->np) * p->nq * p->k_pitch * 7 * sizeof(double), p->stream));
    p->k_only_A = true;  // K holds only A: the scalar Stokes instantiations apply
  }
  lp.t = lut_dev(lut);
  lp.nf = p->nf; lp.f = p->d_f; lp.f_stride = p->f_stride; lp.ffac = p->d_ffac; lp.T = p->d_T; lp.P = p->d_P; lp.vmr = p->d_vmr;
  lp.n_species = p->cat->n_species; lp.h2o_species = h2o_species; lp.select_species = p->select_species;
  lp.K = p->d_K; lp.dK = p->d_dK; lp.k_pitch = p->k_pitch; lp.nq = p->nq;
  lp.no_neg = p->no_neg; lp.po = po; lp.to = to; lp.wo = wo; lp.fo = fo; lp.extpol = extpolfac; lp.flags = p->d_flags;
  AB_TRY(launch_lookup(lp, p->np, p->stream));
  return AB200_OK;
}

int ab200_path_add_predefined(ab200_path* p, const int32_t* models, int32_t n_models, const ab200_predef_species* species,
                              const double* target_d) {
  return ab200_path_add_predefined_data(p, models, n_models, species, target_d, nullptr);
}

int ab200_path_add_predefined_data(ab200_path* p, const int32_t* models, int32_t n_models, const ab200_predef_species* species,
                                   const double* target_d, const ab200_predef_data* data) {
  if (!p || !p->uploaded) return set_error(AB200_ERR_INVALID, "ab200_path_add_predefined: path not uploaded");
  AB_CUDA(cudaSetDevice(p->cat->device));
  return predef_on_path(models, n_models, species, target_d, p->nf, p->d_f, p->f_stride, p->d_ffac, p->d_T, p->d_P, p->d_vmr,
                        p->cat->n_species, p->select_species, p->d_K, p->d_dK, p->k_pitch, p->nq, p->tg_kind, p->tg_species, p->np,
                        p->d_flags, (p->flags & AB200_FLAG_WIND_ROWS_DF) ? nullptr : p->d_wjac, p->stream, data);
}

int ab200_path_add_cia(ab200_path* p, const ab200_cia* cia, double T_extrapolfac, int32_t ignore_errors, double dT) {
  if (!p || !p->uploaded) return set_error(AB200_ERR_INVALID, "ab200_path_add_cia: path not uploaded");
  if (!cia) return set_error(AB200_ERR_INVALID, "ab200_path_add_cia: null CIA data");
  if (cia_device(cia) != p->cat->device) return set_error(AB200_ERR_INVALID, "ab200_path_add_cia: CIA data and path live on different devices");
  for (int q = 0; q < p->nq; q++)
    if (p->tg_kind[q] >= AB200_TARGET_WIND_U && p->tg_kind[q] <= AB200_TARGET_WIND_W)
      return set_error(AB200_ERR_UNSUPPORTED, "collision-induced absorption with a wind target (the re-extraction at f + df of "
                                              "src/m_cia.cc:78-81, :123-129) is outside the GPU path; no CPU fallback");
  if (p->it >= 0 && !std::isnormal(dT))  // m_cia.cc:89-91
    return set_error(AB200_ERR_INVALID, "dt must be >0 and not NaN or Inf: " + std::to_string(dT));
  AB_CUDA(cudaSetDevice(p->cat->device));
  CiaParams cp{};
  cp.c = cia_dev(cia);
  cp.nf = p->nf; cp.f = p->d_f; cp.f_stride = p->f_stride; cp.ffac = p->d_ffac; cp.T = p->d_T; cp.P = p->d_P; cp.vmr = p->d_vmr;
  cp.n_species = p->cat->n_species; cp.select_species = p->select_species;
  if (cia_max_species(cia) >= cp.n_species)
    return set_error(AB200_ERR_INVALID, "ab200_path_add_cia: a CIA record names a species the catalog does not have");
  cp.K = p->d_K; cp.dK = p->d_dK; cp.k_pitch = p->k_pitch; cp.nq = p->nq; cp.it = p->it;
  for (int q = 0; q < p->nq; q++) { cp.tg_kind[q] = p->tg_kind[q]; cp.tg_species[q] = p->tg_species[q]; }
  cp.dt = dT; cp.T_extrapolfac = T_extrapolfac; cp.ignore_errors = ignore_errors; cp.flags = p->d_flags;
  AB_TRY(launch_cia(cp, p->np, p->stream));
  return AB200_OK;
}

int ab200_path_run_observer(ab200_path* p, const ab200_observer* o) {
  if (!p || !p->uploaded) return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: path not uploaded");
  if (!o) return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: null observer");
  if (p->np < 1) return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: empty path");
  if (p->f_stride != 0)
    return set_error(AB200_ERR_UNSUPPORTED,
                     "ab200_path_run_observer needs the sensor's frequency grid: upload one grid (f_level_stride = 0) and "
                     "pass the wind in ab200_atm_path instead of per-level grids");
  if (o->unit < AB200_UNIT_UNIT || o->unit > AB200_UNIT_W_M2_M1_SR)
    return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: unknown spectral radiance unit");
  if (o->bkg_kind != AB200_BKG_UPLOADED && o->bkg_kind != AB200_BKG_PLANCK)
    return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: unknown background kind");
  if (o->bkg_kind == AB200_BKG_PLANCK && !(o->bkg_T > 0))
    return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: the background temperature must be positive");
  if (o->nx < 0 || o->n_bkg < 0 || o->n_channels < 0) return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: negative size");
  const int32_t n_bkg = o->bkg_kind == AB200_BKG_PLANCK ? o->n_bkg : 0;
  const int64_t rows = int64_t(p->np) * p->nq;
  const bool has_jx = o->nx > 0 && (p->nq > 0 || n_bkg > 0);
  if (p->nq > 0 && o->nx > 0 && (!o->map_offset || o->map_offset[0] != 0))
    return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: map_offset must start at 0");
  const int64_t nmap = (p->nq > 0 && o->nx > 0) ? o->map_offset[rows] : 0;
  // Mismatched input sizes of spectral_rad_jacAddPathPropagation (m_rad.cc:77-99): every x index inside [0, nx)
  for (int64_t r = 0; r < (nmap ? rows : 0); r++)
    if (o->map_offset[r + 1] < o->map_offset[r]) return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: map_offset must ascend");
  for (int64_t e = 0; e < nmap; e++)
    if (o->map_x[e] < 0 || o->map_x[e] >= o->nx)
      return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: path map entry " + std::to_string(e) + " is outside the state vector");
  for (int32_t b = 0; b < n_bkg; b++)
    if (o->bkg_x[b] < 0 || o->bkg_x[b] >= o->nx)
      return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: background row " + std::to_string(b) + " is outside the state vector");
  const int64_t nnz = o->n_channels ? o->w_offset[o->n_channels] : 0;
  if (o->n_channels && o->w_offset[0] != 0) return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: w_offset must start at 0");
  for (int32_t c = 0; c < o->n_channels; c++)
    if (o->w_offset[c + 1] < o->w_offset[c]) return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: w_offset must ascend");
  for (int64_t e = 0; e < nnz; e++)
    if (o->w_freq[e] < 0 || o->w_freq[e] >= p->nf)
      return set_error(AB200_ERR_INVALID, "ab200_path_run_observer: sensor weight " + std::to_string(e) + " is outside the frequency grid");
  AB_CUDA(cudaSetDevice(p->cat->device));

  auto put = [&](ab200_path::DevBuf& b, const void* src, size_t bytes) -> int {
    if (bytes == 0) return 0;
    if (b.reserve(bytes)) return set_error(AB200_ERR_NOMEM, "ab200_path_run_observer: out of device memory");
    AB_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, p->stream));  // pageable source: staged before return
    return 0;
  };
  if (nmap) {
    AB_TRY(put(p->o_map_offset, o->map_offset, (rows + 1) * sizeof(int64_t)));
    AB_TRY(put(p->o_map_x, o->map_x, nmap * sizeof(int32_t)));
    AB_TRY(put(p->o_map_w, o->map_w, nmap * sizeof(double)));
  } else if (p->nq > 0 && has_jx) {  // no entries: an all-zero offset table
    std::vector<int64_t> zero(rows + 1, 0);
    AB_TRY(put(p->o_map_offset, zero.data(), zero.size() * sizeof(int64_t)));
    AB_CUDA(cudaStreamSynchronize(p->stream));
  }
  AB_TRY(put(p->o_bkg_x, o->bkg_x, n_bkg * sizeof(int32_t)));
  AB_TRY(put(p->o_bkg_w, o->bkg_w, n_bkg * sizeof(double)));
  AB_TRY(put(p->o_w_offset, o->w_offset, o->n_channels ? (o->n_channels + 1) * sizeof(int64_t) : 0));
  AB_TRY(put(p->o_w_freq, o->w_freq, nnz * sizeof(int64_t)));
  AB_TRY(put(p->o_w_stokes, o->w_stokes, nnz * 4 * sizeof(double)));
  if (has_jx) {
    const size_t bytes = static_cast<size_t>(o->nx) * p->nf * 4 * sizeof(double);
    if (p->o_Jx.reserve(bytes)) return set_error(AB200_ERR_NOMEM, "ab200_path_run_observer: out of device memory for spectral_rad_jac");
    AB_CUDA(cudaMemsetAsync(p->o_Jx.p, 0, bytes, p->stream));  // spectral_rad_jacEmpty, m_rad.cc:14-23
  }
  if (o->n_channels) {
    if (p->o_y.reserve(o->n_channels * sizeof(double)) ||
        p->o_Jy.reserve(std::max<size_t>(1, static_cast<size_t>(o->n_channels) * o->nx) * sizeof(double)))
      return set_error(AB200_ERR_NOMEM, "ab200_path_run_observer: out of device memory");
  }
  if (o->bkg_kind == AB200_BKG_PLANCK) AB_TRY(launch_background_planck(p->nf, p->d_f, o->bkg_T, p->d_Ibkg, p->stream));

  ab200_observer oo = *o;
  if (!has_jx) oo.n_bkg = 0;
  if (has_jx) {
    AB_TRY(run_stokes_impl(p, &oo));
  } else {
    AB_TRY(run_stokes_impl(p, nullptr));
  }
  double* Jx = has_jx ? static_cast<double*>(p->o_Jx.p) : nullptr;
  AB_TRY(launch_unit_transform(p->nf, o->nx, p->d_f, o->unit, o->n_real, p->d_I, Jx, p->stream));
  AB_TRY(launch_sensor_sumup(p->nf, o->nx, o->n_channels, static_cast<const int64_t*>(p->o_w_offset.p),
                             static_cast<const int64_t*>(p->o_w_freq.p), static_cast<const double*>(p->o_w_stokes.p), p->d_I,
                             Jx, static_cast<double*>(p->o_y.p), static_cast<double*>(p->o_Jy.p), p->stream));
  p->o_nx = o->nx; p->o_nch = o->n_channels; p->o_ran = true; p->o_has_jx = has_jx;
  return AB200_OK;
}

int ab200_path_download_observer(ab200_path* p, double* I, double* Jx, double* y, double* Jy) {
  if (!p || !p->o_ran) return set_error(AB200_ERR_INVALID, "ab200_path_download_observer: ab200_path_run_observer has not run");
  if (I && p->nf)
    AB_CUDA(cudaMemcpyAsync(I, p->d_I, static_cast<size_t>(p->nf) * 4 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  const size_t jx_bytes = static_cast<size_t>(p->o_nx) * p->nf * 4 * sizeof(double);
  if (Jx && jx_bytes) {
    if (p->o_has_jx) AB_CUDA(cudaMemcpyAsync(Jx, p->o_Jx.p, jx_bytes, cudaMemcpyDeviceToHost, p->stream));
    else std::memset(Jx, 0, jx_bytes);  // no target reaches the state vector: spectral_rad_jacEmpty
  }
  if (y && p->o_nch)
    AB_CUDA(cudaMemcpyAsync(y, p->o_y.p, p->o_nch * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (Jy && p->o_nch && p->o_nx) {
    const size_t bytes = static_cast<size_t>(p->o_nch) * p->o_nx * sizeof(double);
    if (p->o_has_jx) AB_CUDA(cudaMemcpyAsync(Jy, p->o_Jy.p, bytes, cudaMemcpyDeviceToHost, p->stream));
    else std::memset(Jy, 0, bytes);
  }
  return check_flags(p);
}

static int check_flags(ab200_path* p) {
  int h = 0;
  AB_CUDA(cudaMemcpyAsync(&h, p->d_flags, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
  AB_CUDA(cudaStreamSynchronize(p->stream));
  if (h) {
    cudaMemsetAsync(p->d_flags, 0, sizeof(int), p->stream);
    if (h & 2) return set_error(AB200_ERR_INVALID, "non-finite line-shape parameter (f0', 1/GD, G0 or strength) at some level");
    if (h & 32)  // PWR98.cc:363-370, MPM89.cc:345-352
      return set_error(AB200_ERR_INVALID, "O2 full absorption model has detected a O2 volume mixing ratio which is below the threshold "
                                          "of 1e-25.  Therefore no calculation is performed.");
    if (h & 64)  // ELL07.cc:99-117
      return set_error(AB200_ERR_INVALID, "Liquid cloud absorption model ELL07: liquid water content above 5e-3 kg/m3, temperature outside "
                                          "210-373 K or frequencies above 25 THz (only valid inside these ranges)");
    if (h & 16)
      return set_error(AB200_ERR_INVALID,
                       "Error in check_limit: a frequency, pressure, temperature offset or water ratio is outside the "
                       "extrapolation limits of a lookup table grid (lagrange_interp.h:572-650)");
    if (h & 8)
      return set_error(AB200_ERR_INVALID,
                       "Problem with CIA species: the temperature of a level is outside the extrapolation range of a data set "
                       "(check_limit for Temperature, lagrange_interp.h:572-650; pass ignore_errors to get NaN instead)");
    return set_error(AB200_ERR_UNSUPPORTED, "negative pressure broadening (G0 < 0) is outside the GPU path");
  }
  return AB200_OK;
}

int ab200_path_sync(ab200_path* p) {
  if (!p) return set_error(AB200_ERR_INVALID, "ab200_path_sync: null path");
  return check_flags(p);
}

int ab200_path_download(ab200_path* p, double* I, double* dI, double* K, double* dK) {
  if (!p) return set_error(AB200_ERR_INVALID, "ab200_path_download: null path");
  if ((dI || dK) && p->nq == 0) return set_error(AB200_ERR_INVALID, "ab200_path_download: the path has no Jacobian targets");
  if (dI && p->nf && p->np)
    AB_CUDA(cudaMemcpyAsync(dI, p->d_dI, static_cast<size_t>(p->nf) * p->np * p->nq * 4 * sizeof(double), cudaMemcpyDeviceToHost,
                            p->stream));
  if (dK && p->nf && p->np)
    AB_CUDA(cudaMemcpy2DAsync(dK, static_cast<size_t>(p->nf) * 56, p->d_dK, static_cast<size_t>(p->k_pitch) * 56,
                              static_cast<size_t>(p->nf) * 56, static_cast<size_t>(p->np) * p->nq, cudaMemcpyDeviceToHost,
                              p->stream));
  if (I && p->nf)
    AB_CUDA(cudaMemcpyAsync(I, p->d_I, static_cast<size_t>(p->nf) * 4 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (K && p->nf && p->np)
    AB_CUDA(cudaMemcpy2DAsync(K, static_cast<size_t>(p->nf) * 56, p->d_K, static_cast<size_t>(p->k_pitch) * 56,
                              static_cast<size_t>(p->nf) * 56, p->np, cudaMemcpyDeviceToHost, p->stream));
  return check_flags(p);
}

void* ab200_path_device_ptr(ab200_path* p, int which) {
  if (!p) return nullptr;
  switch (which) {
    case 0: return p->d_I;
    case 1: return p->d_K;
    case 2: return p->d_dI;
    case 3: return p->d_dK;
    default: return nullptr;
  }
}

// ---------------------------------------------------------------------------
// host-buffer entry points (what the WSM shims call)
// ---------------------------------------------------------------------------
namespace {
struct PathCache {
  ab200_path* path = nullptr;
  const ab200_catalog* cat = nullptr;
  uint64_t serial = 0;  // of the catalog the workspace was sized for: an address can be reused by a later catalog
  int64_t nf = -1, ntiles = -1;
  int32_t np = -1, nq = -1;
  int32_t n_species = -1, n_isot = -1;
  size_t nseg = 0;
};
thread_local PathCache t_cache;
thread_local bool t_stream_set = false;
thread_local void* t_stream = nullptr;
thread_local std::vector<double> t_bounds;  // [np][2] bounds of the whole grid the thread's next calls are a shard of (empty: none)

// one cached workspace per host thread: the shims are called repeatedly with the same
// shapes (once per (pos, los) under measurement_vecFromSensor, src/m_rad.cc:321-343)
int cached_path(const ab200_catalog* cat, int64_t nf, int32_t np, int32_t nq, ab200_path** out) {
  PathCache& c = t_cache;
  if (c.path && c.cat == cat && c.serial == cat->serial && c.nf == nf && c.np == np && c.nq == nq && c.ntiles == cat->ntiles &&
      c.n_species == cat->n_species && c.n_isot == cat->n_isot && c.nseg == cat->segments.size()) {
    *out = c.path;
    if (t_stream_set && c.path->stream != static_cast<cudaStream_t>(t_stream)) AB_TRY(ab200_path_set_stream(c.path, t_stream));
    return AB200_OK;
  }
  if (c.path) {
    ab200_path_destroy(c.path);
    c.path = nullptr;
  }
  AB_TRY(ab200_path_create(cat, nf, np, nq, out));
  c = PathCache{*out, cat, cat->serial, nf, cat->ntiles, np, nq, cat->n_species, cat->n_isot, cat->segments.size()};
  if (t_stream_set) AB_TRY(ab200_path_set_stream(*out, t_stream));
  return AB200_OK;
}
}  // namespace

// the host-buffer entry points of this thread take a SHARD of a frequency grid: ByLine cutoffs select their lines with the
// bounds of the whole grid (band_data::active_lines, lbl_data.cpp:61-68), whatever part of it a call carries
int ab200_set_thread_grid_bounds(int32_t np, const double* bounds) {
  if (np <= 0 || !bounds) {
    t_bounds.clear();
    return AB200_OK;
  }
  for (int ip = 0; ip < np; ip++)
    if (!(bounds[2 * ip] <= bounds[2 * ip + 1])) return set_error(AB200_ERR_INVALID, "ab200_set_thread_grid_bounds: bounds must be ascending and finite");
  t_bounds.assign(bounds, bounds + 2 * static_cast<size_t>(np));
  return AB200_OK;
}
static int apply_thread_bounds(ab200_path* p, int np) {
  if (t_bounds.empty()) {
    p->grid_bounds.clear();
    return AB200_OK;
  }
  if (t_bounds.size() != 2 * static_cast<size_t>(np))
    return set_error(AB200_ERR_INVALID, "ab200_set_thread_grid_bounds was given " + std::to_string(t_bounds.size() / 2) +
                                            " levels, the call has " + std::to_string(np));
  p->grid_bounds = t_bounds;
  return AB200_OK;
}

int ab200_set_thread_stream(void* stream) {
  t_stream_set = stream != nullptr;
  t_stream     = stream;
  if (!t_stream_set && t_cache.path && !t_cache.path->own_stream) {  // back to a private stream
    ab200_path_destroy(t_cache.path);
    t_cache = PathCache{};
  }
  return AB200_OK;
}

// drop the calling thread's cached workspace (call before destroying a catalog it was built on)
int ab200_release_thread_cache(void) {
  if (t_cache.path) ab200_path_destroy(t_cache.path);
  t_cache = PathCache{};
  return AB200_OK;
}

int ab200_propmat_levels(const ab200_catalog* cat, int64_t nf, const double* f, int64_t f_level_stride,
                         const ab200_atm_path* atm, int32_t select_species, int32_t no_negative_absorption,
                         int32_t nq, const ab200_target* targets, uint32_t flags, double* K, double* dK) {
  if (!cat || !atm || !f || !K) return set_error(AB200_ERR_INVALID, "ab200_propmat_levels: null argument");
  if (nq > 0 && !dK) return set_error(AB200_ERR_INVALID, "ab200_propmat_levels: dK is null with nq > 0");
  ab200_path* p = nullptr;
  AB_TRY(cached_path(cat, nf, atm->np, nq, &p));
  AB_TRY(apply_thread_bounds(p, atm->np));
  AB_TRY(ab200_path_upload(p, f, f_level_stride, atm, select_species, no_negative_absorption, targets, nullptr, 0,
                           AB200_RTE_LINSRC, nullptr, flags));
  if (!(flags & AB200_FLAG_K_ZERO_INIT) && nf && atm->np) {  // += into the caller's values
    AB_CUDA(cudaMemcpy2DAsync(p->d_K, static_cast<size_t>(p->k_pitch) * 56, K, static_cast<size_t>(nf) * 56,
                              static_cast<size_t>(nf) * 56, atm->np, cudaMemcpyHostToDevice, p->stream));
    p->k_preloaded = true;
    if (nq > 0) {
      AB_CUDA(cudaMemcpy2DAsync(p->d_dK, static_cast<size_t>(p->k_pitch) * 56, dK, static_cast<size_t>(nf) * 56,
                                static_cast<size_t>(nf) * 56, static_cast<size_t>(atm->np) * nq, cudaMemcpyHostToDevice,
                                p->stream));
      p->dk_preloaded = true;
    }
  }
  AB_TRY(ab200_path_run_propmat(p));
  return ab200_path_download(p, nullptr, nullptr, K, nq > 0 ? dK : nullptr);
}

int ab200_lookup_precompute(const ab200_catalog* cat, int64_t nf, const double* f, const ab200_atm_path* atm_ref, int32_t select_species,
                            int32_t h2o_species, int32_t nt, const double* t_pert, int32_t nw, const double* w_pert,
                            const ab200_partfun_table* partfun, double* xsec) {
  if (!cat || !atm_ref || !xsec || (nf > 0 && !f)) return set_error(AB200_ERR_INVALID, "ab200_lookup_precompute: null argument");
  if (nf < 0 || nt < 1 || nw < 1 || (t_pert == nullptr && nt != 1) || (w_pert == nullptr && nw != 1))
    return set_error(AB200_ERR_INVALID, "ab200_lookup_precompute: bad perturbation grid sizes");
  if (select_species < 0 || select_species >= cat->n_species)
    return set_error(AB200_ERR_INVALID, "ab200_lookup_precompute: the table needs one species (atm_point.number_density(species))");
  if (w_pert && (h2o_species < 0 || h2o_species >= cat->n_species))
    return set_error(AB200_ERR_INVALID, "ab200_lookup_precompute: a water grid needs the index of H2O in the VMR vector");
  const int np = atm_ref->np;
  for (int ip = 1; ip < np; ip++)  // DescendingGrid log_p_grid, lookup_map.h
    if (!(atm_ref->P[ip] < atm_ref->P[ip - 1]))
      return set_error(AB200_ERR_INVALID, "the reference profile of a lookup table must have descending pressures");
  if (np == 0 || nf == 0) return AB200_OK;
  ab200_path* p = nullptr;
  AB_TRY(cached_path(cat, nf, np, 0, &p));
  double* d_x = nullptr;
  AB_TRY(dev_alloc(&d_x, static_cast<size_t>(np) * nf));
  struct Free { double* q; ~Free() { cudaFree(q); } } guard{d_x};
  std::vector<double> Tp(np), vp(static_cast<size_t>(np) * cat->n_species), Qp(static_cast<size_t>(np) * cat->n_isot);
  ab200_atm_path a = *atm_ref;
  a.T = Tp.data();
  a.vmr = vp.data();
  a.dQdT = nullptr;
  for (int it = 0; it < nt; it++) {
    for (int ip = 0; ip < np; ip++) Tp[ip] = atm_ref->T[ip] + (t_pert ? t_pert[it] : 0.0);
    if (partfun) {  // PartitionFunctions::Q at the perturbed temperature, as lbl::calculate does (line_strength_calc :22-36)
      AB_TRY(ab200_partfun_eval(partfun, cat->n_isot, np, Tp.data(), Qp.data(), nullptr));
      a.Q = Qp.data();
    }
    for (int iw = 0; iw < nw; iw++) {
      for (int ip = 0; ip < np; ip++) {  // :93-96
        for (int s = 0; s < cat->n_species; s++) {
          double v = atm_ref->vmr[static_cast<size_t>(ip) * cat->n_species + s];
          if (w_pert && s == h2o_species) v *= w_pert[iw];
          vp[static_cast<size_t>(ip) * cat->n_species + s] = v;
        }
      }
      AB_TRY(ab200_path_upload(p, f, 0, &a, select_species, /*no_negative_absorption=*/1, nullptr, nullptr, 0, AB200_RTE_LINSRC, nullptr,
                               AB200_FLAG_K_ZERO_INIT));
      AB_TRY(ab200_path_run_propmat(p));
      AB_TRY(launch_xsec_from_K(np, nf, p->d_K, p->k_pitch, p->d_T, p->d_P, p->d_vmr, cat->n_species, select_species, d_x, p->stream));
      AB_CUDA(cudaMemcpyAsync(xsec + (static_cast<size_t>(it) * nw + iw) * np * nf, d_x, static_cast<size_t>(np) * nf * sizeof(double),
                              cudaMemcpyDeviceToHost, p->stream));
      AB_TRY(check_flags(p));  // synchronises: Tp / vp and d_x are reused by the next point
    }
  }
  return AB200_OK;
}

int ab200_clearsky_emission(const ab200_catalog* cat, int64_t nf, const double* f, int64_t f_level_stride,
                            const ab200_atm_path* atm, int32_t select_species, int32_t no_negative_absorption,
                            int32_t nq, const ab200_target* targets, const double* r, int32_t hse_derivative,
                            int32_t rte_option, const double* I_bkg, uint32_t flags, double* I, double* dI,
                            double* K_out) {
  if (!cat || !atm || !f || !I || !I_bkg) return set_error(AB200_ERR_INVALID, "ab200_clearsky_emission: null argument");
  if (atm->np > 1 && !r) return set_error(AB200_ERR_INVALID, "ab200_clearsky_emission: r is null");
  if (nq > 0 && !dI) return set_error(AB200_ERR_INVALID, "ab200_clearsky_emission: dI is null with nq > 0");
  ab200_path* p = nullptr;
  AB_TRY(cached_path(cat, nf, atm->np, nq, &p));
  AB_TRY(apply_thread_bounds(p, atm->np));
  AB_TRY(ab200_path_upload(p, f, f_level_stride, atm, select_species, no_negative_absorption, targets, r,
                           hse_derivative, rte_option, I_bkg, flags));
  AB_TRY(ab200_path_run_propmat(p));
  AB_TRY(ab200_path_run_stokes(p));
  return ab200_path_download(p, I, nq > 0 ? dI : nullptr, (flags & AB200_FLAG_RETURN_K) ? K_out : nullptr, nullptr);
}

// --- un-fused compatibility entry points -----------------------------------------
namespace {
struct DevBuf {
  double* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  int alloc(size_t n) { return dev_alloc(&p, n); }
};
}  // namespace

int ab200_tramat(int32_t np, int64_t nf, int32_t nq, const double* K, const double* dK, const double* r,
                 const double* dr, int32_t rte_option, uint32_t flags, double* T, double* L, double* P, double* dT,
                 double* dL) {
  if (rte_option != AB200_RTE_CONSTANT && rte_option != AB200_RTE_LINSRC && rte_option != AB200_RTE_LINPROP)
    return set_error(AB200_ERR_INVALID, "unknown rte_option");
  if (np < 0 || nf < 0 || nq < 0) return set_error(AB200_ERR_INVALID, "ab200_tramat: negative size");
  if (np == 0 || nf == 0) return AB200_OK;
  const bool linprop = rte_option == AB200_RTE_LINPROP;
  const bool linsrc  = rte_option == AB200_RTE_LINSRC || linprop;  // L, dL exist for both
  if (!K || !T || !P || (np > 1 && !r) || (linsrc && !L)) return set_error(AB200_ERR_INVALID, "ab200_tramat: null argument");
  if (nq > 0 && (!dK || !dT || (np > 1 && !dr) || (linsrc && !dL)))
    return set_error(AB200_ERR_INVALID, "ab200_tramat: null Jacobian argument with nq > 0");
  if (nq > 0 && (flags & AB200_FLAG_TRAN_EXACT))
    return set_error(AB200_ERR_UNSUPPORTED, "AB200_FLAG_TRAN_EXACT has no Jacobian (the reference differentiates its literal form)");
  const size_t nk = static_cast<size_t>(np) * nf * 7, nm = static_cast<size_t>(np) * nf * 16;
  DevBuf dK_, dr_, dT_, dL_, dP_, ddK_, ddr_, ddT_, ddL_;
  AB_TRY(dK_.alloc(nk)); AB_TRY(dr_.alloc(std::max(np - 1, 1))); AB_TRY(dT_.alloc(nm)); AB_TRY(dP_.alloc(nm));
  if (linsrc) AB_TRY(dL_.alloc(nm));
  AB_CUDA(cudaMemcpy(dK_.p, K, nk * sizeof(double), cudaMemcpyHostToDevice));
  if (np > 1) AB_CUDA(cudaMemcpy(dr_.p, r, (np - 1) * sizeof(double), cudaMemcpyHostToDevice));
  int* d_flag = nullptr;
  AB_TRY(dev_alloc(&d_flag, 1));
  struct FlagGuard { int* p; ~FlagGuard() { cudaFree(p); } } flag_guard{d_flag};
  AB_CUDA(cudaMemset(d_flag, 0, sizeof(int)));
  AB_TRY(launch_tramat(np, nf, dK_.p, dr_.p, linsrc, (flags & AB200_FLAG_TRAN_EXACT) ? 1 : 0, dT_.p, dL_.p, dP_.p, linprop,
                       d_flag, 0));
  int h_flag = 0;
  AB_CUDA(cudaMemcpy(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
  AB_CUDA(cudaMemcpy(T, dT_.p, nm * sizeof(double), cudaMemcpyDeviceToHost));
  AB_CUDA(cudaMemcpy(P, dP_.p, nm * sizeof(double), cudaMemcpyDeviceToHost));
  if (linsrc) AB_CUDA(cudaMemcpy(L, dL_.p, nm * sizeof(double), cudaMemcpyDeviceToHost));
  if (nq > 0) {
    const size_t nd = 2 * nm * nq;
    AB_TRY(ddK_.alloc(nk * nq)); AB_TRY(ddr_.alloc(2 * static_cast<size_t>(std::max(np - 1, 1)) * nq)); AB_TRY(ddT_.alloc(nd));
    if (linsrc) AB_TRY(ddL_.alloc(nd));
    AB_CUDA(cudaMemcpy(ddK_.p, dK, nk * nq * sizeof(double), cudaMemcpyHostToDevice));
    if (np > 1) AB_CUDA(cudaMemcpy(ddr_.p, dr, 2 * static_cast<size_t>(np - 1) * nq * sizeof(double), cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemset(ddT_.p, 0, nd * sizeof(double)));  // dT = muelmat::constant(0), rtepack_transmission.cc:1300-1314
    if (linsrc) AB_CUDA(cudaMemset(ddL_.p, 0, nd * sizeof(double)));
    AB_TRY(launch_tramat_jac(np, nf, nq, dK_.p, ddK_.p, dr_.p, ddr_.p, linsrc, ddT_.p, ddL_.p, linprop, 0));
    AB_CUDA(cudaMemcpy(dT, ddT_.p, nd * sizeof(double), cudaMemcpyDeviceToHost));
    if (linsrc) AB_CUDA(cudaMemcpy(dL, ddL_.p, nd * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return AB200_OK;
}

int ab200_srcvec(int32_t np, int64_t nf, int32_t nq, const double* K, const double* f, int64_t f_level_stride,
                 const double* T_level, int32_t it, double* J, double* dJ) {
  if (np < 0 || nf < 0 || nq < 0) return set_error(AB200_ERR_INVALID, "ab200_srcvec: negative size");
  if (np == 0 || nf == 0) return AB200_OK;
  if (!K || !f || !T_level || !J || (nq > 0 && !dJ)) return set_error(AB200_ERR_INVALID, "ab200_srcvec: null argument");
  if (f_level_stride != 0 && f_level_stride != nf) return set_error(AB200_ERR_INVALID, "f_level_stride must be 0 or nf");
  const size_t nk = static_cast<size_t>(np) * nf * 7, nj = static_cast<size_t>(np) * nf * 4;
  const size_t nfl = static_cast<size_t>(nf) * (f_level_stride ? np : 1);
  DevBuf dK_, df_, dT_, dJ_, ddJ_;
  AB_TRY(dK_.alloc(nk)); AB_TRY(df_.alloc(nfl)); AB_TRY(dT_.alloc(np)); AB_TRY(dJ_.alloc(nj)); AB_TRY(ddJ_.alloc(nj * nq));
  AB_CUDA(cudaMemcpy(dK_.p, K, nk * sizeof(double), cudaMemcpyHostToDevice));
  AB_CUDA(cudaMemcpy(df_.p, f, nfl * sizeof(double), cudaMemcpyHostToDevice));
  AB_CUDA(cudaMemcpy(dT_.p, T_level, np * sizeof(double), cudaMemcpyHostToDevice));
  AB_TRY(launch_srcvec(np, nf, nq, dK_.p, df_.p, f_level_stride, dT_.p, it, dJ_.p, ddJ_.p, 0));
  AB_CUDA(cudaMemcpy(J, dJ_.p, nj * sizeof(double), cudaMemcpyDeviceToHost));
  if (nq) AB_CUDA(cudaMemcpy(dJ, ddJ_.p, nj * nq * sizeof(double), cudaMemcpyDeviceToHost));
  return AB200_OK;
}

int ab200_rte_emission(int32_t rte_option, int32_t np, int64_t nf, int32_t nq, const double* T, const double* L,
                       const double* P, const double* dT, const double* dL, const double* J, const double* dJ,
                       const double* I_bkg, double* I, double* dI) {
  if (rte_option != AB200_RTE_CONSTANT && rte_option != AB200_RTE_LINSRC && rte_option != AB200_RTE_LINPROP)
    return set_error(AB200_ERR_INVALID, "unknown rte_option");
  if (np < 0 || nf < 0 || nq < 0) return set_error(AB200_ERR_INVALID, "ab200_rte_emission: negative size");
  if (nf == 0) return AB200_OK;
  const bool linsrc = rte_option != AB200_RTE_CONSTANT;  // linsrc and linprop share linevo (rtepack_rtestep.cc:392-401)
  if (!T || !J || !I_bkg || !I || (linsrc && !L)) return set_error(AB200_ERR_INVALID, "ab200_rte_emission: null argument");
  if (nq > 0 && (!P || !dT || !dJ || !dI || (linsrc && !dL)))
    return set_error(AB200_ERR_INVALID, "ab200_rte_emission: null Jacobian argument with nq > 0");
  const size_t nm = static_cast<size_t>(np) * nf * 16, nj = static_cast<size_t>(np) * nf * 4;
  DevBuf dT_, dL_, dJ_, dB_, dI_, dP_, ddT_, ddL_, ddJ_, ddI_;
  AB_TRY(dT_.alloc(nm)); AB_TRY(dJ_.alloc(nj)); AB_TRY(dB_.alloc(nf * 4)); AB_TRY(dI_.alloc(nf * 4));
  if (linsrc) AB_TRY(dL_.alloc(nm));
  AB_CUDA(cudaMemcpy(dT_.p, T, nm * sizeof(double), cudaMemcpyHostToDevice));
  if (linsrc) AB_CUDA(cudaMemcpy(dL_.p, L, nm * sizeof(double), cudaMemcpyHostToDevice));
  AB_CUDA(cudaMemcpy(dJ_.p, J, nj * sizeof(double), cudaMemcpyHostToDevice));
  AB_CUDA(cudaMemcpy(dB_.p, I_bkg, nf * 4 * sizeof(double), cudaMemcpyHostToDevice));
  if (nq == 0) {
    AB_TRY(launch_rte_emission(linsrc, np, nf, dT_.p, dL_.p, dJ_.p, dB_.p, dI_.p, 0));
  } else {
    const size_t nd = 2 * nm * nq;
    AB_TRY(dP_.alloc(nm)); AB_TRY(ddT_.alloc(nd)); AB_TRY(ddJ_.alloc(nj * nq)); AB_TRY(ddI_.alloc(nj * nq));
    if (linsrc) AB_TRY(ddL_.alloc(nd));
    AB_CUDA(cudaMemcpy(dP_.p, P, nm * sizeof(double), cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemcpy(ddT_.p, dT, nd * sizeof(double), cudaMemcpyHostToDevice));
    if (linsrc) AB_CUDA(cudaMemcpy(ddL_.p, dL, nd * sizeof(double), cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemcpy(ddJ_.p, dJ, nj * nq * sizeof(double), cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemset(ddI_.p, 0, nj * nq * sizeof(double)));  // spectral_rad_jac_path = 0, m_spectral_radiance.cc:36-40
    AB_TRY(launch_rte_emission_jac(linsrc, np, nf, nq, dT_.p, dL_.p, dP_.p, ddT_.p, ddL_.p, dJ_.p, ddJ_.p, dB_.p, dI_.p,
                                   ddI_.p, 0));
    AB_CUDA(cudaMemcpy(dI, ddI_.p, nj * nq * sizeof(double), cudaMemcpyDeviceToHost));
  }
  AB_CUDA(cudaMemcpy(I, dI_.p, nf * 4 * sizeof(double), cudaMemcpyDeviceToHost));
  return AB200_OK;
}

int ab200_rte_transmission(int32_t np, int64_t nf, int32_t nq, const double* T, const double* P, const double* dT,
                           const double* I_bkg, double* I, double* dI) {
  if (np < 0 || nf < 0 || nq < 0) return set_error(AB200_ERR_INVALID, "ab200_rte_transmission: negative size");
  if (nf == 0) return AB200_OK;
  if (!P || !I_bkg || !I) return set_error(AB200_ERR_INVALID, "ab200_rte_transmission: null argument");
  if (nq > 0 && (!T || !dT || !dI)) return set_error(AB200_ERR_INVALID, "ab200_rte_transmission: null Jacobian argument with nq > 0");
  if (np == 0) return AB200_OK;  // rtepack_rtestep.cc:465
  const size_t nm = static_cast<size_t>(np) * nf * 16, nj = static_cast<size_t>(np) * nf * 4;
  DevBuf dP_, dB_, dI_;
  AB_TRY(dP_.alloc(nm)); AB_TRY(dB_.alloc(nf * 4)); AB_TRY(dI_.alloc(nf * 4));
  AB_CUDA(cudaMemcpy(dP_.p, P, nm * sizeof(double), cudaMemcpyHostToDevice));
  AB_CUDA(cudaMemcpy(dB_.p, I_bkg, nf * 4 * sizeof(double), cudaMemcpyHostToDevice));
  if (nq > 0) {
    // the derivative of the transmitted radiance = the `constant` emission recursion with J = 0, dJ = 0
    const size_t nd = 2 * nm * nq;
    DevBuf dT_, ddT_, dJ_, ddJ_, ddI_;
    AB_TRY(dT_.alloc(nm)); AB_TRY(ddT_.alloc(nd)); AB_TRY(dJ_.alloc(nj)); AB_TRY(ddJ_.alloc(nj * nq)); AB_TRY(ddI_.alloc(nj * nq));
    AB_CUDA(cudaMemcpy(dT_.p, T, nm * sizeof(double), cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemcpy(ddT_.p, dT, nd * sizeof(double), cudaMemcpyHostToDevice));
    AB_CUDA(cudaMemset(dJ_.p, 0, nj * sizeof(double)));
    AB_CUDA(cudaMemset(ddJ_.p, 0, nj * nq * sizeof(double)));
    AB_CUDA(cudaMemset(ddI_.p, 0, nj * nq * sizeof(double)));  // dI = 0, rtepack_rtestep.cc:467
    AB_TRY(launch_rte_emission_jac(false, np, nf, nq, dT_.p, nullptr, dP_.p, ddT_.p, nullptr, dJ_.p, ddJ_.p, dB_.p, dI_.p,
                                   ddI_.p, 0));
    AB_CUDA(cudaMemcpy(dI, ddI_.p, nj * nq * sizeof(double), cudaMemcpyDeviceToHost));
  }
  AB_TRY(launch_transmission_apply(np, nf, dP_.p, dB_.p, dI_.p, 0));  // I = P[np-1] I0, :469-470
  AB_CUDA(cudaMemcpy(I, dI_.p, nf * 4 * sizeof(double), cudaMemcpyDeviceToHost));
  return AB200_OK;
}

int ab200_planck_tb(int64_t nf, const double* f, double* I) {
  if (nf < 0) return set_error(AB200_ERR_INVALID, "ab200_planck_tb: negative size");
  if (nf == 0) return AB200_OK;
  if (!f || !I) return set_error(AB200_ERR_INVALID, "ab200_planck_tb: null argument");
  DevBuf df_, dI_;
  AB_TRY(df_.alloc(nf)); AB_TRY(dI_.alloc(nf * 4));
  AB_CUDA(cudaMemcpy(df_.p, f, nf * sizeof(double), cudaMemcpyHostToDevice));
  AB_CUDA(cudaMemcpy(dI_.p, I, nf * 4 * sizeof(double), cudaMemcpyHostToDevice));
  AB_TRY(launch_planck_tb(nf, df_.p, dI_.p, 0));
  AB_CUDA(cudaMemcpy(I, dI_.p, nf * 4 * sizeof(double), cudaMemcpyDeviceToHost));
  return AB200_OK;
}

// ---------------------------------------------------------------------------
// measurement helpers
// ---------------------------------------------------------------------------
static int measure_fp64(int iters, bool mix, double* tflops, double* ms) {
  if (!tflops || !ms || iters <= 0) return set_error(AB200_ERR_INVALID, "ab200_measure_dfma_peak: bad argument");
  int dev = 0, sms = 0;
  AB_CUDA(cudaGetDevice(&dev));
  AB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8;
  DevBuf out;
  AB_TRY(out.alloc(1));
  cudaEvent_t e0, e1;
  AB_CUDA(cudaEventCreate(&e0));
  AB_CUDA(cudaEventCreate(&e1));
  const int sign = mix ? -1 : 1;
  AB_TRY(launch_dfma_peak(sign * (iters / 4 + 1), blocks, out.p, 0));  // warm-up
  AB_CUDA(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    AB_CUDA(cudaEventRecord(e0, 0));
    AB_TRY(launch_dfma_peak(sign * iters, blocks, out.p, 0));
    AB_CUDA(cudaEventRecord(e1, 0));
    AB_CUDA(cudaEventSynchronize(e1));
    float t = 0;
    AB_CUDA(cudaEventElapsedTime(&t, e0, e1));
    best = std::min(best, t);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // plain: 8 chains x 8 unrolled DFMA per iteration; mix: 8 chains x 7 DFMA (+ 1 MUFU.RCP64H each)
  const double fma = double(blocks) * 256.0 * double(iters) * (mix ? 56.0 : 64.0);
  *ms     = best;
  *tflops = 2.0 * fma / (best * 1e-3) / 1e12;
  return AB200_OK;
}

int ab200_measure_dfma_peak(int iters, double* tflops, double* ms) { return measure_fp64(iters, false, tflops, ms); }
int ab200_measure_dfma_mix(int iters, double* tflops, double* ms) { return measure_fp64(iters, true, tflops, ms); }

int ab200_faddeeva_w(int64_t n, const double* zr, const double* zi, double* wr, double* wi) {
  if (n < 0) return set_error(AB200_ERR_INVALID, "ab200_faddeeva_w: negative size");
  if (n == 0) return AB200_OK;
  if (!zr || !zi || !wr || !wi) return set_error(AB200_ERR_INVALID, "ab200_faddeeva_w: null argument");
  DevBuf a, b, c, d;
  AB_TRY(a.alloc(n)); AB_TRY(b.alloc(n)); AB_TRY(c.alloc(n)); AB_TRY(d.alloc(n));
  AB_CUDA(cudaMemcpy(a.p, zr, n * sizeof(double), cudaMemcpyHostToDevice));
  AB_CUDA(cudaMemcpy(b.p, zi, n * sizeof(double), cudaMemcpyHostToDevice));
  AB_TRY(launch_faddeeva(n, a.p, b.p, c.p, d.p, 0));
  AB_CUDA(cudaMemcpy(wr, c.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  AB_CUDA(cudaMemcpy(wi, d.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  return AB200_OK;
}

// Faddeeva::Dawson(complex) on the device (the element-wise function of rtepack::dawson(specmat)), for tests
int ab200_dawson(int64_t n, const double* zr, const double* zi, double* dr, double* di) {
  if (n < 0) return set_error(AB200_ERR_INVALID, "ab200_dawson: negative size");
  if (n == 0) return AB200_OK;
  if (!zr || !zi || !dr || !di) return set_error(AB200_ERR_INVALID, "ab200_dawson: null argument");
  DevBuf a, b, c, d;
  AB_TRY(a.alloc(n)); AB_TRY(b.alloc(n)); AB_TRY(c.alloc(n)); AB_TRY(d.alloc(n));
  AB_CUDA(cudaMemcpy(a.p, zr, n * sizeof(double), cudaMemcpyHostToDevice));
  AB_CUDA(cudaMemcpy(b.p, zi, n * sizeof(double), cudaMemcpyHostToDevice));
  AB_TRY(launch_dawson(n, a.p, b.p, c.p, d.p, 0));
  AB_CUDA(cudaMemcpy(dr, c.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  AB_CUDA(cudaMemcpy(di, d.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  return AB200_OK;
}

}  // extern "C"
