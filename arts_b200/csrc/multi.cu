// multi.cu — ONE host process driving several GPUs behind the C ABI.
//
// The reference is a single process whose frequency loop is an OpenMP team (src/m_lbl.cc:273-295: chunks of ONE
// freq_grid per thread; src/m_rad.cc:321-343 for the batch of paths).  A shim inside that process therefore sees one
// call per path with the whole grid; ab200_multi_* splits that call's frequencies over the visible devices — one
// catalog replica, one worker thread (with its own stream, pinned staging and cached workspace) per device — and every
// device copies its results straight into the caller's arrays.  No collective: the path is independent per frequency.
//
// ab200_multi_propmat_levels deals contiguous blocks of LEVELS (its outputs are level-major: no replication, no exchange).
// ab200_multi_clearsky_emission with Jacobian targets or per-level grids splits the frequencies:
// 512-frequency blocks dealt round-robin (block b -> device b mod N).  The cost of a frequency is not uniform
// (the number of near pairs grows with frequency where Doppler widths do, and with a ByLine cutoff so does the number of
// lines in window), so contiguous chunks like the reference's omp_offset_count are unbalanced: round 1 measured 1.71x
// on 2 GPUs for configs[3] with a 750 GHz cutoff.  512 is the block width of the line-sum kernels, so a device's CTAs see
// exactly the blocks of the single-device run, and because the value at a frequency does not depend on the block or
// shard it is computed in, the result is bit-identical to the one-device result.  Line selection of
// ByLine cutoffs uses the bounds of the WHOLE grid (band_data::active_lines, lbl_data.cpp:61-68), handed to each worker
// through ab200_set_thread_grid_bounds.
//
// Second split (default for forward calls on a shared grid, AB200_MULTI_SPLIT=freq switches it off): LEVELS for the line
// sum, frequencies for the Stokes chain.  The line records and the cluster moments of the far-field sums are work per
// (line, level) that every frequency shard would repeat (30 of 93 ms per configs[3] shard: 8 devices gave 5.3x).  Device d
// therefore sums levels d, d + N, d + 2N, ... for ALL frequencies (nothing is replicated), and the one exchange the path
// then has is a transpose of K: device d pulls its contiguous frequency slice of every peer's levels straight out of the
// peer's HBM over NVLink (one strided cudaMemcpy2DAsync per peer, ordered by events, no host staging, 0.6 GB per device at
// 1e6 x 100) and runs the fused Stokes chain on it.  The value of K at a (frequency, level) does not depend on the
// partition, so the radiances stay bit-identical to the one-device call.
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {
constexpr int64_t MB = 512;  // frequencies per dealt block (= F_TILE of the line-sum kernels)

struct Worker {
  int device = 0;
  ab200_catalog* cat = nullptr;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = false, quit = false;
  int rc = 0;
  std::string err;
  // gather / scatter buffers of this worker (host)
  std::vector<double> f, bkg, I, dI, K, dK;
  // level-sharded split: p1 sums this device's levels over the whole grid, p2 runs the Stokes chain on this device's
  // frequency slice of all levels
  ab200_path *p1 = nullptr, *p2 = nullptr;
  int64_t p1_nf = -1, p2_nf = -1;
  int32_t p1_np = -1, p2_np = -1;
  cudaEvent_t ev_k = nullptr;  // K of p1 is complete
  std::vector<double> aT, aP, avmr, aiso, aQ, amag, alos, awind, ar;  // this device's levels of the caller's atm path

  void loop() {
    ab200_set_device(device);
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      cv.wait(lk, [&] { return has_job || quit; });
      if (quit) break;
      lk.unlock();
      const int r = job();
      const std::string e = r ? ab200_last_error() : "";
      lk.lock();
      rc = r;
      err = e;
      has_job = false;
      done = true;
      cv.notify_all();
    }
    if (p1) ab200_path_destroy(p1);
    if (p2) ab200_path_destroy(p2);
    if (ev_k) cudaEventDestroy(ev_k);
    ab200_release_thread_cache();
  }
};
}  // namespace

struct ab200_multi {
  std::vector<Worker*> w;
  std::mutex call_mu;  // one multi-device call at a time per set (the workers' workspaces are per set)
  int32_t n_species = 0, n_isot = 0;
};

namespace {
// the frequencies of device d: blocks d, d + n, d + 2n, ... of MB frequencies
int64_t local_count(int64_t nf, int n, int d) {
  const int64_t nb = (nf + MB - 1) / MB;
  int64_t c = 0;
  for (int64_t b = d; b < nb; b += n) c += std::min(MB, nf - b * MB);
  return c;
}
template <class Fn>  // fn(global offset, local offset, count) for every block of device d
void for_blocks(int64_t nf, int n, int d, Fn fn) {
  const int64_t nb = (nf + MB - 1) / MB;
  int64_t lo = 0;
  for (int64_t b = d; b < nb; b += n) {
    const int64_t cnt = std::min(MB, nf - b * MB);
    fn(b * MB, lo, cnt);
    lo += cnt;
  }
}

int run_all(ab200_multi* m, const std::function<int(Worker&, int)>& body) {
  std::lock_guard<std::mutex> call(m->call_mu);
  const int n = static_cast<int>(m->w.size());
  for (int d = 0; d < n; d++) {
    Worker& w = *m->w[d];
    std::lock_guard<std::mutex> lk(w.mu);
    w.job = [&body, &w, d] { return body(w, d); };
    w.done = false;
    w.has_job = true;
    w.cv.notify_all();
  }
  int rc = 0;
  std::string err;
  for (int d = 0; d < n; d++) {
    Worker& w = *m->w[d];
    std::unique_lock<std::mutex> lk(w.mu);
    w.cv.wait(lk, [&] { return w.done; });
    if (w.rc && !rc) {
      rc = w.rc;
      err = "device " + std::to_string(w.device) + ": " + w.err;
    }
  }
  return rc ? ab200::set_error(rc, err) : AB200_OK;
}

// bounds of the whole grid per level, as ab200_path_upload derives them from a grid it is given in full
void grid_bounds(int64_t nf, const double* f, int64_t stride, int np, std::vector<double>& b) {
  b.resize(2 * static_cast<size_t>(np));
  for (int ip = 0; ip < np; ip++) {
    const double* g = f + static_cast<size_t>(ip) * stride;
    b[2 * ip] = g[0];
    b[2 * ip + 1] = g[nf - 1];
  }
}
}  // namespace

int ab200_multi_create(const ab200_catalog_desc* desc, int32_t n_devices, const int32_t* devices, ab200_multi** out) {
  if (!desc || !out) return ab200::set_error(AB200_ERR_INVALID, "ab200_multi_create: null argument");
  *out = nullptr;
  const int avail = ab200_device_count();
  if (n_devices <= 0) n_devices = avail;
  if (avail < 1 || (!devices && n_devices > avail))
    return ab200::set_error(AB200_ERR_CUDA, "ab200_multi_create: " + std::to_string(n_devices) + " devices requested, " +
                                                std::to_string(avail) + " visible (arts_b200 has no CPU fallback)");
  for (int i = 0; devices && i < n_devices; i++)  // an explicit list may name a device more than once (two workers on it)
    if (devices[i] < 0 || devices[i] >= avail)
      return ab200::set_error(AB200_ERR_INVALID, "ab200_multi_create: device " + std::to_string(devices[i]) + " is not one of the " +
                                                     std::to_string(avail) + " visible devices");
  int prev = 0;
  cudaGetDevice(&prev);
  ab200_multi* m = new ab200_multi;
  m->n_species = desc->n_species;
  m->n_isot = desc->n_isot;
  for (int i = 0; i < n_devices; i++) {
    Worker* w = new Worker;
    w->device = devices ? devices[i] : i;
    int rc = ab200_set_device(w->device);
    if (!rc) rc = ab200_catalog_create(desc, &w->cat);  // one replica per device
    if (rc) {
      delete w;
      cudaSetDevice(prev);
      ab200_multi_destroy(m);
      return rc;
    }
    m->w.push_back(w);
  }
  // peers read each other's K directly (NVLink); without peer access the same copies are staged by the driver
  for (Worker* a : m->w)
    for (Worker* b : m->w) {
      if (a->device == b->device) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, a->device, b->device) == cudaSuccess && can) {
        cudaSetDevice(a->device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
        if (e != cudaSuccess) cudaGetLastError();  // already enabled: fine
      }
    }
  cudaSetDevice(prev);
  for (Worker* w : m->w) w->th = std::thread([w] { w->loop(); });
  *out = m;
  return AB200_OK;
}

void ab200_multi_destroy(ab200_multi* m) {
  if (!m) return;
  for (Worker* w : m->w) {
    if (w->th.joinable()) {
      {
        std::lock_guard<std::mutex> lk(w->mu);
        w->quit = true;
        w->cv.notify_all();
      }
      w->th.join();
    }
    if (w->cat) ab200_catalog_destroy(w->cat);
    delete w;
  }
  delete m;
}

int32_t ab200_multi_device_count(const ab200_multi* m) { return m ? static_cast<int32_t>(m->w.size()) : 0; }

namespace {
bool level_split_wanted() {
  const char* e = getenv("AB200_MULTI_SPLIT");
  return !(e && std::string(e) == "freq");
}

// Forward clear-sky call, levels dealt over the devices for the line sum, contiguous frequency slices for the Stokes chain.
int clearsky_level_split(ab200_multi* m, int64_t nf, const double* f, const ab200_atm_path* atm, int32_t select_species,
                         int32_t no_neg, const double* r, int32_t rte_option, const double* I_bkg, uint32_t flags, double* I,
                         double* K_out) {
  const int n = static_cast<int>(m->w.size()), np = atm->np;
  // frequency slice of device d for stage 2: whole 128-frequency rows (the TMA boxes of the Stokes kernel)
  const int64_t chunk = ((nf + n - 1) / n + 127) / 128 * 128;
  auto slice = [&](int d, int64_t& off, int64_t& cnt) {
    off = std::min<int64_t>(nf, static_cast<int64_t>(d) * chunk);
    cnt = std::min<int64_t>(chunk, nf - off);
  };
  const bool want_K = (flags & AB200_FLAG_RETURN_K) && K_out;
  // phase A: every device uploads, sums its levels over the whole grid and records an event
  int rc = run_all(m, [&](Worker& w, int d) -> int {
    const int npd = (np - d + n - 1) / n;  // levels d, d + n, ...
    const int npd_cap = (np + n - 1) / n;
    if (!w.p1 || w.p1_nf != nf || w.p1_np != npd_cap) {
      if (w.p1) ab200_path_destroy(w.p1);
      w.p1 = nullptr;
      AB_TRY(ab200_path_create(w.cat, nf, npd_cap, 0, &w.p1));
      w.p1_nf = nf; w.p1_np = npd_cap;
    }
    int64_t off, cnt;
    slice(d, off, cnt);
    if (!w.p2 || w.p2_nf != cnt || w.p2_np != np) {
      if (w.p2) ab200_path_destroy(w.p2);
      w.p2 = nullptr;
      AB_TRY(ab200::path_create_ex(w.cat, cnt, np, 0, true, &w.p2));
      w.p2_nf = cnt; w.p2_np = np;
    }
    if (!w.ev_k) AB_CUDA(cudaEventCreateWithFlags(&w.ev_k, cudaEventDisableTiming));
    // this device's levels of the atmosphere
    const int nsp = m->n_species, nis = m->n_isot;
    w.aT.resize(npd); w.aP.resize(npd); w.avmr.resize(static_cast<size_t>(npd) * nsp); w.aiso.resize(static_cast<size_t>(npd) * nis);
    w.aQ.resize(static_cast<size_t>(npd) * nis); w.amag.resize(3 * static_cast<size_t>(npd)); w.alos.resize(2 * static_cast<size_t>(npd));
    w.awind.resize(3 * static_cast<size_t>(npd)); w.ar.assign(static_cast<size_t>(std::max(npd, 1)), 0.0);
    for (int j = 0; j < npd; j++) {
      const size_t ip = static_cast<size_t>(d) + static_cast<size_t>(j) * n;
      w.aT[j] = atm->T[ip]; w.aP[j] = atm->P[ip];
      std::copy(atm->vmr + ip * nsp, atm->vmr + (ip + 1) * nsp, w.avmr.begin() + static_cast<size_t>(j) * nsp);
      std::copy(atm->isorat + ip * nis, atm->isorat + (ip + 1) * nis, w.aiso.begin() + static_cast<size_t>(j) * nis);
      std::copy(atm->Q + ip * nis, atm->Q + (ip + 1) * nis, w.aQ.begin() + static_cast<size_t>(j) * nis);
      if (atm->mag) std::copy(atm->mag + 3 * ip, atm->mag + 3 * ip + 3, w.amag.begin() + 3 * static_cast<size_t>(j));
      if (atm->los) std::copy(atm->los + 2 * ip, atm->los + 2 * ip + 2, w.alos.begin() + 2 * static_cast<size_t>(j));
      if (atm->wind) std::copy(atm->wind + 3 * ip, atm->wind + 3 * ip + 3, w.awind.begin() + 3 * static_cast<size_t>(j));
    }
    ab200_atm_path sub{};
    sub.np = npd; sub.T = w.aT.data(); sub.P = w.aP.data(); sub.vmr = w.avmr.data(); sub.isorat = w.aiso.data(); sub.Q = w.aQ.data();
    sub.dQdT = nullptr; sub.mag = atm->mag ? w.amag.data() : nullptr; sub.los = atm->los ? w.alos.data() : nullptr;
    sub.wind = atm->wind ? w.awind.data() : nullptr;
    AB_TRY(ab200_path_set_grid_bounds(w.p1, nullptr));
    AB_TRY(ab200_path_upload(w.p1, f, 0, &sub, select_species, no_neg, nullptr, w.ar.data(), 0, rte_option, nullptr, flags));
    if (npd > 0) AB_TRY(ab200_path_run_propmat(w.p1));
    AB_CUDA(cudaEventRecord(w.ev_k, ab200::path_stream(w.p1)));
    // stage-2 workspace: this device's frequency slice, all levels
    if (cnt > 0) {
      AB_TRY(ab200_path_set_grid_bounds(w.p2, nullptr));
      AB_TRY(ab200_path_upload(w.p2, f + off, 0, atm, select_species, no_neg, nullptr, r, 0, rte_option, I_bkg + 4 * off, flags));
    }
    return AB200_OK;
  });
  if (rc) return rc;
  // phase B: pull the slice of every device's levels, run the Stokes chain, return the radiances
  return run_all(m, [&](Worker& w, int d) -> int {
    int64_t off, cnt;
    slice(d, off, cnt);
    AB_TRY(ab200_path_sync(w.p1));  // device error flags of the line sum (also keeps p1 alive until its K is complete)
    if (cnt == 0) return AB200_OK;
    int64_t pitch2 = 0;
    double* K2 = ab200::path_K(w.p2, &pitch2);
    cudaStream_t s2 = ab200::path_stream(w.p2);
    for (int e = 0; e < n; e++) {
      Worker& src = *m->w[(d + e) % n];  // start with the own rows, then round the ring: no two devices pull from one peer at once
      const int de = (d + e) % n;
      const int npe = (np - de + n - 1) / n;
      if (npe <= 0) continue;
      int64_t pitch1 = 0;
      const double* K1 = ab200::path_K(src.p1, &pitch1);
      AB_CUDA(cudaStreamWaitEvent(s2, src.ev_k, 0));
      // rows j = 0..npe-1 of the peer are levels de + j n of the path
      AB_CUDA(cudaMemcpy2DAsync(K2 + static_cast<size_t>(de) * pitch2 * 7, static_cast<size_t>(n) * pitch2 * 56, K1 + static_cast<size_t>(off) * 7,
                                static_cast<size_t>(pitch1) * 56, static_cast<size_t>(cnt) * 56, static_cast<size_t>(npe), cudaMemcpyDefault, s2));
    }
    AB_TRY(ab200::path_adopt_K(w.p2));
    AB_TRY(ab200_path_run_stokes(w.p2));
    if (want_K) {
      w.K.resize(static_cast<size_t>(np) * cnt * 7);
      AB_TRY(ab200_path_download(w.p2, I + 4 * off, nullptr, w.K.data(), nullptr));
      for (int ip = 0; ip < np; ip++)
        std::memcpy(K_out + (static_cast<size_t>(ip) * nf + off) * 7, &w.K[static_cast<size_t>(ip) * cnt * 7], 7 * cnt * sizeof(double));
      return AB200_OK;
    }
    return ab200_path_download(w.p2, I + 4 * off, nullptr, nullptr, nullptr);
  });
}
}  // namespace

int ab200_multi_clearsky_emission(ab200_multi* m, int64_t nf, const double* f, int64_t f_level_stride, const ab200_atm_path* atm,
                                  int32_t select_species, int32_t no_negative_absorption, int32_t nq, const ab200_target* targets,
                                  const double* r, int32_t hse_derivative, int32_t rte_option, const double* I_bkg, uint32_t flags,
                                  double* I, double* dI, double* K_out) {
  if (!m || !atm || !f || !I || !I_bkg) return ab200::set_error(AB200_ERR_INVALID, "ab200_multi_clearsky_emission: null argument");
  if (f_level_stride != 0 && f_level_stride != nf)
    return ab200::set_error(AB200_ERR_INVALID, "f_level_stride must be 0 (shared grid) or nf (one grid per level)");
  if (nq > 0 && !dI) return ab200::set_error(AB200_ERR_INVALID, "ab200_multi_clearsky_emission: dI is null with nq > 0");
  if (nf == 0) return AB200_OK;
  const int n = static_cast<int>(m->w.size()), np = atm->np;
  if (n > 1 && nq == 0 && f_level_stride == 0 && np >= n && nf >= 128 * static_cast<int64_t>(n) && level_split_wanted() && (np < 2 || r))
    return clearsky_level_split(m, nf, f, atm, select_species, no_negative_absorption, r, rte_option, I_bkg, flags, I, K_out);
  const int nlev_f = f_level_stride ? np : 1;
  const bool want_K = (flags & AB200_FLAG_RETURN_K) && K_out;
  std::vector<double> bounds;
  grid_bounds(nf, f, f_level_stride, f_level_stride ? np : 1, bounds);
  if (!f_level_stride) {
    bounds.resize(2 * static_cast<size_t>(std::max(np, 1)));
    for (int ip = 1; ip < np; ip++) bounds[2 * ip] = bounds[0], bounds[2 * ip + 1] = bounds[1];
  }
  const size_t row = static_cast<size_t>(np) * nq * 4;  // doubles of dI per frequency
  return run_all(m, [&](Worker& w, int d) -> int {
    const int64_t cnt = local_count(nf, n, d);
    if (cnt == 0) return AB200_OK;
    w.f.resize(static_cast<size_t>(cnt) * nlev_f);
    w.bkg.resize(static_cast<size_t>(cnt) * 4);
    w.I.resize(static_cast<size_t>(cnt) * 4);
    if (nq > 0) w.dI.resize(static_cast<size_t>(cnt) * row);
    if (want_K) w.K.resize(static_cast<size_t>(np) * cnt * 7);
    for_blocks(nf, n, d, [&](int64_t g, int64_t l, int64_t c) {
      for (int ip = 0; ip < nlev_f; ip++)
        std::memcpy(&w.f[static_cast<size_t>(ip) * cnt + l], f + static_cast<size_t>(ip) * f_level_stride + g, c * sizeof(double));
      std::memcpy(&w.bkg[4 * l], I_bkg + 4 * g, 4 * c * sizeof(double));
    });
    AB_TRY(ab200_set_thread_grid_bounds(np, bounds.data()));
    const int rc = ab200_clearsky_emission(w.cat, cnt, w.f.data(), f_level_stride ? cnt : 0, atm, select_species, no_negative_absorption,
                                           nq, targets, r, hse_derivative, rte_option, w.bkg.data(), flags, w.I.data(),
                                           nq > 0 ? w.dI.data() : nullptr, want_K ? w.K.data() : nullptr);
    ab200_set_thread_grid_bounds(0, nullptr);
    if (rc) return rc;
    for_blocks(nf, n, d, [&](int64_t g, int64_t l, int64_t c) {
      std::memcpy(I + 4 * g, &w.I[4 * l], 4 * c * sizeof(double));
      if (nq > 0) std::memcpy(dI + g * row, &w.dI[l * row], c * row * sizeof(double));
      if (want_K)
        for (int ip = 0; ip < np; ip++)
          std::memcpy(K_out + (static_cast<size_t>(ip) * nf + g) * 7, &w.K[(static_cast<size_t>(ip) * cnt + l) * 7], 7 * c * sizeof(double));
    });
    return AB200_OK;
  });
}

int ab200_multi_propmat_levels(ab200_multi* m, int64_t nf, const double* f, int64_t f_level_stride, const ab200_atm_path* atm,
                               int32_t select_species, int32_t no_negative_absorption, int32_t nq, const ab200_target* targets,
                               uint32_t flags, double* K, double* dK) {
  if (!m || !atm || !f || !K) return ab200::set_error(AB200_ERR_INVALID, "ab200_multi_propmat_levels: null argument");
  if (f_level_stride != 0 && f_level_stride != nf)
    return ab200::set_error(AB200_ERR_INVALID, "f_level_stride must be 0 (shared grid) or nf (one grid per level)");
  if (nq > 0 && !dK) return ab200::set_error(AB200_ERR_INVALID, "ab200_multi_propmat_levels: dK is null with nq > 0");
  if (nf == 0 || atm->np == 0) return AB200_OK;
  // K [np][nf][7] and dK [np][nq][nf][7] are level-major: device d takes a contiguous block of LEVELS for all frequencies and
  // writes its rows straight into the caller's arrays (`+=` included) - nothing is replicated (the line records and cluster
  // moments are work per level), nothing is exchanged, nothing is gathered on the host.  Levels do not interact in this call,
  // so every row is the row of the one-device call bit for bit.
  const int n = static_cast<int>(m->w.size()), np = atm->np;
  const int per = (np + n - 1) / n;
  return run_all(m, [&](Worker& w, int d) -> int {
    const int l0 = std::min(np, d * per), l1 = std::min(np, l0 + per);
    if (l1 <= l0) return AB200_OK;
    ab200_atm_path sub = *atm;
    sub.np = l1 - l0;
    sub.T = atm->T + l0; sub.P = atm->P + l0;
    sub.vmr = atm->vmr + static_cast<size_t>(l0) * m->n_species;
    sub.isorat = atm->isorat + static_cast<size_t>(l0) * m->n_isot;
    sub.Q = atm->Q + static_cast<size_t>(l0) * m->n_isot;
    sub.dQdT = atm->dQdT ? atm->dQdT + static_cast<size_t>(l0) * m->n_isot : nullptr;
    sub.mag = atm->mag ? atm->mag + 3 * static_cast<size_t>(l0) : nullptr;
    sub.los = atm->los ? atm->los + 2 * static_cast<size_t>(l0) : nullptr;
    sub.wind = atm->wind ? atm->wind + 3 * static_cast<size_t>(l0) : nullptr;
    return ab200_propmat_levels(w.cat, nf, f + static_cast<size_t>(l0) * f_level_stride, f_level_stride, &sub, select_species,
                                no_negative_absorption, nq, targets, flags, K + static_cast<size_t>(l0) * nf * 7,
                                nq > 0 ? dK + static_cast<size_t>(l0) * nq * nf * 7 : nullptr);
  });
}
