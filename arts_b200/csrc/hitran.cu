// hitran.cu — catalog ingest (host code): HITRAN .par records straight into the SoA of ab200_catalog_desc.
//
// Replaces, for file_formatter = ["par"], line_strength_option = "A", compute_zeeman_parameters = 0:
//   read_par_line        src/core/lbl/lbl_hitran.cpp:66-89   the 160-column record and its unit conversions
//   read_hitran_par      src/core/lbl/lbl_hitran.cpp:146-172 frequency window (skip below, stop at the first above)
//   hitran_record::from  src/core/lbl/lbl_hitran.cpp:180-237 lbl::line + line_shape::model (T0 = 296 K)
//   abs_bandsReadHITRAN  src/m_lbl.cc:302-338               bands keyed by the global state = the isotopologue
// Unit conversions: src/core/util/arts_conversions.h:51,85-88,126-128,136-138,146.
//
// The reference reads one std::string per record and builds an AoS of unordered_maps; here the file is one buffer, the
// record starts are found with memchr, the records are parsed by several host threads into flat per-line arrays and a
// stable counting sort groups them by isotopologue.  Every field is parsed by std::from_chars on the trimmed column,
// like the reference (fast_float::from_chars / std::from_chars, :33-49): same value for every correctly rounded input.
#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

struct ab200_hitran_catalog {
  ab200_catalog_desc desc{};
  std::vector<int32_t> isot_species, band_isot, band_lineshape, band_cutoff_type;
  std::vector<double> isot_mass, band_cutoff_value;
  std::vector<int64_t> band_offset;
  // per-line / per-broadener arrays: allocated uninitialised and written once, in parallel
  std::unique_ptr<double[]> f0, a, e0, gu, gl, T0, zero, ls_X;
  std::unique_ptr<int32_t[]> ls_species, ls_type, two_J;
  std::unique_ptr<int64_t[]> ls_offset;
  std::unique_ptr<uint8_t[]> z_on;
};

namespace ab200 {
namespace {

// arts_conversions.h: kaycm2freq(x) = x * (100 * c)
constexpr double kC = 299792458.0;
constexpr double kH = 6.62607015e-34;
constexpr double kaycm2freq(double x) { return x * (100 * kC); }
constexpr double kS_FACTOR     = kaycm2freq(1e-4);            // kaycm_per_cmsquared2hz_per_msquared :126-128
constexpr double kGAMMA_FACTOR = kaycm2freq(1 / 101'325.0);   // kaycm_per_atm2hz_per_pa :136-138 with pa2atm(1) :88
constexpr double kE_FACTOR     = kaycm2freq(kH);              // kaycm2joule :146

struct Record {
  int32_t isot;  // index into the caller's table, -1: skipped (below fmin)
  double f0, A, gamma_air, gamma_self, E, n, delta, g_upp, g_low;
};
enum Status : uint8_t { ST_OK = 0, ST_SKIP = 1, ST_ERROR = 2 };

struct Field {
  const char* b;
  const char* e;
};
inline Field trimmed(const char* p, int n) {  // reader::read_next :24-26
  const char *b = p, *e = p + n;
  while (b < e && *b == ' ') b++;
  while (e > b && e[-1] == ' ') e--;
  return {b, e};
}
template <typename T>
inline bool parse(const char* p, int n, T& x) {  // :28-49: the whole trimmed column must parse
  const Field f = trimmed(p, n);
  const auto res = std::from_chars(f.b, f.e, x);
  return res.ec == std::errc{} && res.ptr == f.e;
}

std::string quoted(const char* p, size_t n) { return "\"" + std::string(p, n) + "\""; }

// One record [p, p + n) (without the newline).  Returns the status and fills rec / err.
Status parse_record(const char* p, size_t n, double fmin, int32_t strength, const ab200_hitran_isotopologue* tab, int32_t ntab, Record& rec,
                    std::string& err) {
  auto fail = [&](const std::string& m) {
    err = m + "\n\nFailed to read HITRAN line record:\n\n" + std::string(p, n);
    return ST_ERROR;
  };
  // columns: M 2 | I 1 | nu 12 | S 10 | A 10 | gamma_air 5 | gamma_self 5 | E'' 10 | n 4 | delta 8 | 79 skipped | g' 7 | g'' 7
  if (n < 15) return fail("Unexpected end of string");
  int64_t M = 0;
  if (!parse(p, 2, M)) return fail("Failed to parse value from string " + quoted(p, 2));
  const char I = p[2];
  double nu = 0;
  if (!parse(p + 3, 12, nu)) return fail("Failed to parse value from string " + quoted(p + 3, 12));
  rec.f0 = kaycm2freq(nu);
  if (rec.f0 < fmin) {  // read_par_line :72-73, before anything else is looked at
    rec.isot = -1;
    return ST_SKIP;
  }
  rec.isot = -1;
  for (int32_t i = 0; i < ntab; i++)
    if (tab[i].M == M && tab[i].I == I) {
      rec.isot = i;
      break;
    }
  if (rec.isot < 0)  // Hitran::id_from_lookup throws for an unknown pair
    return fail("HITRAN molecule " + std::to_string(M) + " isotopologue '" + std::string(1, I) + "' is not in the isotopologue table");
  if (n < 160) return fail("Unexpected end of string");
  double S = 0;
  struct Col { int off, len; double* out; };
  const Col cols[] = {{15, 10, &S},          {25, 10, &rec.A}, {35, 5, &rec.gamma_air}, {40, 5, &rec.gamma_self},
                      {45, 10, &rec.E},      {55, 4, &rec.n},  {59, 8, &rec.delta},     {146, 7, &rec.g_upp},
                      {153, 7, &rec.g_low}};
  for (const Col& c : cols)
    if (!parse(p + c.off, c.len, *c.out)) return fail("Failed to parse value from string " + quoted(p + c.off, c.len));
  // read_hitran_par_record skips ONE separator character after the par block (`if (not data.end_of_string()) data.skip(1)`,
  // lbl_hitran.cpp:126) and only then reports a remainder (:133-135): a 161-byte record (CRLF files, one trailing
  // separator) loads, from 162 bytes on the text after the separator is the error
  if (n > 161) return fail("Part of the line was not parsed: '" + std::string(p + 161, n - 161) + "'");
  rec.gamma_air  = rec.gamma_air * kGAMMA_FACTOR;
  rec.gamma_self = rec.gamma_self * kGAMMA_FACTOR;
  rec.E          = rec.E * kE_FACTOR;
  rec.delta      = rec.delta * kGAMMA_FACTOR;
  if (strength == AB200_HITRAN_STRENGTH_S) {
    // hitran_record::from :193-200: line::hitran_a (lbl_data.cpp:164-169) = compute_a(S / Ia) = einstein_a(s, gu, e0, f0,
    // 296 K, Q(296)) (:34-40, :155-162); HitranLineStrengthOption::A keeps the file's coefficient
    if (rec.g_upp == 0.0) rec.g_upp = rec.g_low = -1.0;
    constexpr double kK = 1.380649e-23, kPI = 3.14159265358979323846, T0 = 296.0;
    const double s  = (S * kS_FACTOR) / tab[rec.isot].hitran_ratio;
    const double cf = kC / rec.f0;
    rec.A = -8.0 * kPI * tab[rec.isot].Q296 * s /
            (rec.g_upp * std::exp(-rec.E / (kK * T0)) * std::expm1(-(kH * rec.f0) / (kK * T0)) * (cf * cf));
  }
  // hitran_record::from :211-216
  if (!std::isnormal(rec.A) || !std::isnormal(rec.g_upp))
    return fail("Invalid Einstein coefficient " + std::to_string(rec.A) + " or gu " + std::to_string(rec.g_upp) +
                " for full HITRAN RECORD");
  return ST_OK;
}

int build(const char* text, int64_t len, double fmin, double fmax, int32_t strength, const ab200_hitran_isotopologue* tab, int32_t ntab,
          int32_t n_species, int32_t n_threads, ab200_hitran_catalog** out) {
  if (!out) return set_error(AB200_ERR_INVALID, "ab200_hitran_read_par: null output");
  *out = nullptr;
  if ((!text && len > 0) || len < 0 || !tab || ntab <= 0 || n_species <= 0)
    return set_error(AB200_ERR_INVALID, "ab200_hitran_read_par: null or empty argument");
  for (int32_t i = 0; i < ntab; i++)
    if (tab[i].species < 0 || tab[i].species >= n_species || !(tab[i].mass > 0))
      return set_error(AB200_ERR_INVALID, "ab200_hitran_read_par: isotopologue " + std::to_string(i) + " has a bad species or mass");
  if (strength != AB200_HITRAN_STRENGTH_S && strength != AB200_HITRAN_STRENGTH_A)
    return set_error(AB200_ERR_INVALID, "ab200_hitran_read_par: unknown line_strength_option");
  if (strength == AB200_HITRAN_STRENGTH_S)
    for (int32_t i = 0; i < ntab; i++)
      if (!(tab[i].hitran_ratio > 0) || !(tab[i].Q296 > 0))
        return set_error(AB200_ERR_INVALID, "ab200_hitran_read_par: line_strength_option S needs hitran_ratio and Q296 of isotopologue " +
                                                std::to_string(i));

  // record boundaries (std::getline: '\n' separated, a trailing newline does not make an empty record)
  std::vector<int64_t> start;
  start.reserve(static_cast<size_t>(len / 161 + 2));
  for (int64_t pos = 0; pos < len;) {
    start.push_back(pos);
    const void* nl = std::memchr(text + pos, '\n', static_cast<size_t>(len - pos));
    pos = nl ? static_cast<const char*>(nl) - text + 1 : len;
  }
  const int64_t nrec = static_cast<int64_t>(start.size());
  auto rec_len = [&](int64_t i) {
    const int64_t end = (i + 1 < nrec) ? start[i + 1] - 1 : ((len > 0 && text[len - 1] == '\n') ? len - 1 : len);
    return static_cast<size_t>(end - start[i]);
  };

  std::unique_ptr<Record[]> recs(new Record[static_cast<size_t>(nrec) + 1]);
  std::unique_ptr<uint8_t[]> status(new uint8_t[static_cast<size_t>(nrec) + 1]);
  int nt = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  nt = std::max(1, std::min<int>(nt, static_cast<int>(std::max<int64_t>(1, nrec / 4096))));
  std::vector<std::string> first_err(static_cast<size_t>(nt));
  std::vector<int64_t> first_err_at(static_cast<size_t>(nt), -1);
  auto work = [&](int t) {
    const int64_t lo = nrec * t / nt, hi = nrec * (t + 1) / nt;
    std::string err;
    for (int64_t i = lo; i < hi; i++) {
      status[i] = parse_record(text + start[i], rec_len(i), fmin, strength, tab, ntab, recs[i], err);
      if (status[i] == ST_ERROR && first_err_at[t] < 0) {
        first_err_at[t] = i;
        first_err[t]    = err;
      }
    }
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }

  // read_hitran_par :153-169 in file order: stop at the first accepted record above fmax; an error before that point
  // is the reference's exception, anything after it is never read
  int64_t stop = nrec;
  for (int64_t i = 0; i < nrec; i++)
    if (status[i] == ST_OK && recs[i].f0 > fmax) {
      stop = i;
      break;
    } else if (status[i] == ST_ERROR) {
      break;  // reported below
    }
  for (int t = 0; t < nt; t++)
    if (first_err_at[t] >= 0 && first_err_at[t] < stop)
      return set_error(AB200_ERR_INVALID, "record " + std::to_string(first_err_at[t] + 1) + ": " + first_err[t]);

  // group by isotopologue (stable): one band per isotopologue that has lines, in table order.  Every thread counts
  // its own range, a prefix sum over (isotopologue, thread) gives every thread its write cursors, and the scatter runs
  // in parallel again.
  auto run = [&](auto&& fn) {
    if (nt == 1) {
      fn(0);
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < nt; t++) th.emplace_back(fn, t);
      for (auto& x : th) x.join();
    }
  };
  auto range = [&](int t, int64_t& lo, int64_t& hi) {
    lo = std::min(stop, nrec * t / nt);
    hi = std::min(stop, nrec * (t + 1) / nt);
  };
  std::vector<int64_t> cnt(static_cast<size_t>(nt) * ntab, 0);
  run([&](int t) {
    int64_t lo, hi;
    range(t, lo, hi);
    int64_t* c_ = cnt.data() + static_cast<size_t>(t) * ntab;
    for (int64_t i = lo; i < hi; i++)
      if (status[i] == ST_OK) c_[recs[i].isot]++;
  });
  std::unique_ptr<ab200_hitran_catalog> c(new ab200_hitran_catalog());
  c->isot_species.resize(ntab);
  c->isot_mass.resize(ntab);
  for (int32_t i = 0; i < ntab; i++) {
    c->isot_species[i] = tab[i].species;
    c->isot_mass[i]    = tab[i].mass;
  }
  std::vector<int64_t> cursor(static_cast<size_t>(nt) * ntab, 0);
  c->band_offset.push_back(0);
  int64_t nl = 0;
  for (int32_t i = 0; i < ntab; i++) {
    const int64_t before = nl;
    for (int t = 0; t < nt; t++) {
      cursor[static_cast<size_t>(t) * ntab + i] = nl;
      nl += cnt[static_cast<size_t>(t) * ntab + i];
    }
    if (nl == before) continue;
    c->band_isot.push_back(i);
    c->band_offset.push_back(nl);
  }
  const size_t nb = c->band_isot.size();
  c->band_lineshape.assign(nb, AB200_LINESHAPE_VP_LTE);  // default_band, m_lbl.cc:312-316
  c->band_cutoff_type.assign(nb, AB200_CUTOFF_NONE);
  c->band_cutoff_value.assign(nb, INFINITY);
  const size_t snl = static_cast<size_t>(nl);
  c->f0.reset(new double[snl + 1]); c->a.reset(new double[snl + 1]); c->e0.reset(new double[snl + 1]);
  c->gu.reset(new double[snl + 1]); c->gl.reset(new double[snl + 1]); c->T0.reset(new double[snl + 1]);
  c->zero.reset(new double[snl + 1]); c->z_on.reset(new uint8_t[snl + 1]); c->two_J.reset(new int32_t[snl + 1]);
  c->ls_offset.reset(new int64_t[snl + 1]);
  c->ls_species.reset(new int32_t[2 * snl + 1]);
  c->ls_type.reset(new int32_t[2 * snl * AB200_NVAR + 1]);
  c->ls_X.reset(new double[2 * snl * AB200_NVAR * 4 + 1]);
  c->ls_offset[snl] = static_cast<int64_t>(2 * snl);
  run([&](int t) {
    int64_t lo, hi;
    range(t, lo, hi);
    int64_t* cur = cursor.data() + static_cast<size_t>(t) * ntab;
    for (int64_t i = lo; i < hi; i++) {
      if (status[i] != ST_OK) continue;
      const Record& r = recs[i];
      const size_t l  = static_cast<size_t>(cur[r.isot]++);
      c->f0[l] = r.f0; c->a[l] = r.A; c->e0[l] = r.E; c->gu[l] = r.g_upp; c->gl[l] = r.g_low;
      c->T0[l] = 296.0;  // :224
      c->zero[l] = 0.0; c->z_on[l] = 0; c->two_J[l] = 0;  // l.z.on = false, :222
      c->ls_offset[l] = static_cast<int64_t>(2 * l);
      // single_models[self] then [Bath], :227-236
      const double gam[2] = {r.gamma_self, r.gamma_air};
      for (int k = 0; k < 2; k++) {
        const size_t e = 2 * l + k;
        c->ls_species[e] = k == 0 ? tab[r.isot].species : AB200_SPECIES_BATH;
        int32_t* ty = c->ls_type.get() + e * AB200_NVAR;
        double* X   = c->ls_X.get() + e * AB200_NVAR * 4;
        for (int v = 0; v < AB200_NVAR; v++) ty[v] = AB200_TM_ABSENT;
        for (int v = 0; v < AB200_NVAR * 4; v++) X[v] = 0.0;
        ty[AB200_VAR_G0]        = AB200_TM_T1;
        X[AB200_VAR_G0 * 4 + 0] = gam[k];
        X[AB200_VAR_G0 * 4 + 1] = r.n;
        if (r.delta != 0) {
          ty[AB200_VAR_D0]        = AB200_TM_T0;
          X[AB200_VAR_D0 * 4 + 0] = r.delta;
        }
      }
    }
  });

  ab200_catalog_desc& d = c->desc;
  d.n_species = n_species; d.n_isot = ntab; d.n_bands = static_cast<int32_t>(nb); d.n_lines = nl; d.n_ls = 2 * nl;
  d.isot_species = c->isot_species.data(); d.isot_mass = c->isot_mass.data();
  d.band_isot = c->band_isot.data(); d.band_lineshape = c->band_lineshape.data();
  d.band_cutoff_type = c->band_cutoff_type.data(); d.band_cutoff_value = c->band_cutoff_value.data();
  d.band_offset = c->band_offset.data();
  d.f0 = c->f0.get(); d.a = c->a.get(); d.e0 = c->e0.get(); d.gu = c->gu.get(); d.gl = c->gl.get(); d.T0 = c->T0.get();
  d.z_on = c->z_on.get(); d.z_gu = c->zero.get(); d.z_gl = c->zero.get(); d.two_Ju = c->two_J.get(); d.two_Jl = c->two_J.get();
  d.ls_offset = c->ls_offset.get(); d.ls_species = c->ls_species.get(); d.ls_type = c->ls_type.get(); d.ls_X = c->ls_X.get();
  *out = c.release();
  return AB200_OK;
}

}  // namespace
}  // namespace ab200

extern "C" {

// PartitionFunctions::Q / dQdT as generated by src/partfun/make_auto_partfuns.cc:28-153 (literal formulas per table kind)
int ab200_partfun_eval(const ab200_partfun_table* tables, int32_t n_isot, int32_t np, const double* T, double* Q, double* dQdT) {
  using ab200::set_error;
  if (n_isot < 0 || np < 0) return set_error(AB200_ERR_INVALID, "ab200_partfun_eval: negative size");
  if ((n_isot > 0 && !tables) || (np > 0 && (!T || !Q))) return set_error(AB200_ERR_INVALID, "ab200_partfun_eval: null argument");
  for (int32_t i = 0; i < n_isot; i++) {
    const ab200_partfun_table& t = tables[i];
    const bool gridded = t.kind == AB200_PARTFUN_INTERP || t.kind == AB200_PARTFUN_STATIC_INTERP;
    if (t.kind < AB200_PARTFUN_INTERP || t.kind > AB200_PARTFUN_STATIC_INTERP)
      return set_error(AB200_ERR_INVALID, "ab200_partfun_eval: isotopologue " + std::to_string(i) + " has an unknown table kind");
    if (!t.coef || t.n < 1 || (gridded && (!t.grid || t.n < 2)))
      return set_error(AB200_ERR_INVALID, "ab200_partfun_eval: isotopologue " + std::to_string(i) + " has an empty table");
    if (gridded)
      for (int32_t k = 1; k < t.n; k++)
        if (!(t.grid[k] > t.grid[k - 1])) return set_error(AB200_ERR_INVALID, "Temperature grid must be increasing");  // :34-37
  }
  for (int32_t ip = 0; ip < np; ip++) {
    const double Tl = T[ip];
    for (int32_t i = 0; i < n_isot; i++) {
      const ab200_partfun_table& t = tables[i];
      double q = 0.0, dq = 0.0;
      switch (t.kind) {
        case AB200_PARTFUN_INTERP: {  // :46-61
          const int64_t i_low = std::lower_bound(t.grid, t.grid + t.n, Tl) - t.grid;
          const size_t k = std::min<size_t>(static_cast<size_t>(i_low - (i_low > 0)), static_cast<size_t>(t.n - 2));
          q  = t.coef[k] + (Tl - t.grid[k]) * (t.coef[k + 1] - t.coef[k]) / (t.grid[k + 1] - t.grid[k]);
          dq = (t.coef[k + 1] - t.coef[k]) / (t.grid[k + 1] - t.grid[k]);
        } break;
        case AB200_PARTFUN_COEFF: {  // :77-103
          double TN = 1.0;
          q = t.coef[0];
          for (int32_t k = 1; k < t.n; k++) {
            TN *= Tl;
            q += TN * t.coef[k];
          }
          dq = t.n > 1 ? t.coef[1] : 0.0;
          TN = 1.0;
          for (int32_t k = 2; k < t.n; k++) {
            TN *= Tl;
            dq += static_cast<double>(k) * TN * t.coef[k];
          }
        } break;
        case AB200_PARTFUN_CONST: q = t.coef[0]; break;  // :107-115
        default: {                                     // STATIC_INTERP :117-153
          const double r_dT = 1.0 / (t.grid[1] - t.grid[0]);
          const double Tx   = (Tl - t.grid[0]) * r_dT;
          const size_t iTx  = static_cast<size_t>(Tx);
          const size_t k    = iTx > static_cast<size_t>(t.n - 2) ? static_cast<size_t>(t.n - 2) : iTx;
          const double To   = Tx - static_cast<double>(k);
          q  = t.coef[k] + To * (t.coef[k + 1] - t.coef[k]);
          dq = (t.coef[k + 1] - t.coef[k]) * r_dT;
        } break;
      }
      Q[static_cast<size_t>(ip) * n_isot + i] = q;
      if (dQdT) dQdT[static_cast<size_t>(ip) * n_isot + i] = dq;
    }
  }
  return AB200_OK;
}

int ab200_hitran_read_par(const char* text, int64_t len, double fmin, double fmax, int32_t line_strength_option,
                          const ab200_hitran_isotopologue* isotopologues,
                          int32_t n_isot, int32_t n_species, int32_t n_threads, ab200_hitran_catalog** out) {
  return ab200::build(text, len, fmin, fmax, line_strength_option, isotopologues, n_isot, n_species, n_threads, out);
}

int ab200_hitran_read_par_file(const char* filename, double fmin, double fmax, int32_t line_strength_option,
                               const ab200_hitran_isotopologue* isotopologues,
                               int32_t n_isot, int32_t n_species, int32_t n_threads, ab200_hitran_catalog** out) {
  if (!filename) return ab200::set_error(AB200_ERR_INVALID, "ab200_hitran_read_par_file: null file name");
  std::FILE* f = std::fopen(filename, "rb");
  if (!f) return ab200::set_error(AB200_ERR_INVALID, std::string("Cannot open file: ") + filename);  // open_input_file
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<char> buf(static_cast<size_t>(std::max<long>(n, 0)));
  const size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
  std::fclose(f);
  if (got != buf.size()) return ab200::set_error(AB200_ERR_INVALID, std::string("Cannot read file: ") + filename);
  return ab200::build(buf.data(), static_cast<int64_t>(buf.size()), fmin, fmax, line_strength_option, isotopologues, n_isot, n_species, n_threads,
                      out);
}

const ab200_catalog_desc* ab200_hitran_desc(const ab200_hitran_catalog* cat) { return cat ? &cat->desc : nullptr; }
void ab200_hitran_destroy(ab200_hitran_catalog* cat) { delete cat; }

}  // extern "C"
