// catalog.cu — ab200_catalog_create/destroy: AbsorptionBands (flattened by the shim) ->
// device SoA with pre-expanded Zeeman sub-lines, grouped into segments and 256-line tiles.
//
// Replaces the per-call AoS-of-maps walk of band_shape_helper / lines_push_back /
// zeeman_push_back (reference src/core/lbl/lbl_lineshape_voigt_lte.cpp:338-429): what is
// static there (which sub-lines exist, their relative strengths and splitting
// coefficients) is done once here; what depends on the atmosphere is left to the prepare
// kernel (lbl.cu).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <numeric>

#include "catalog.hpp"

namespace ab200 {

// ---- Zeeman bookkeeping (host) ------------------------------------------------
static int zeeman_dM(int pol) { return pol == POL_SM ? -1 : (pol == POL_SP ? 1 : 0); }  // lbl_zeeman.h:18-26

int64_t zeeman_size(bool on, int two_Jl, int pol) {  // lbl_zeeman.cpp:298-309, lbl_zeeman.h:92-96
  if (on) return pol == POL_NO ? 0 : two_Jl + 1;
  return pol == POL_NO ? 1 : 0;
}

// |<j1 m1; 1 q | J M>|^2 / (2J+1) = (j1 1 J; m1 q -M)^2 in closed form; all arguments doubled.
static double threej_sq_j2_1(int tj1, int tJ, int tm1, int q) {
  const double j1 = 0.5 * tj1, M = 0.5 * tm1 + q;
  double cg2;
  if (tJ == tj1 + 2) {
    if (q == 1) cg2 = (j1 + M) * (j1 + M + 1) / ((2 * j1 + 1) * (2 * j1 + 2));
    else if (q == 0) cg2 = (j1 - M + 1) * (j1 + M + 1) / ((2 * j1 + 1) * (j1 + 1));
    else cg2 = (j1 - M) * (j1 - M + 1) / ((2 * j1 + 1) * (2 * j1 + 2));
  } else if (tJ == tj1) {
    if (tj1 == 0) return 0.0;
    if (q == 1) cg2 = (j1 + M) * (j1 - M + 1) / (2 * j1 * (j1 + 1));
    else if (q == 0) cg2 = M * M / (j1 * (j1 + 1));
    else cg2 = (j1 - M) * (j1 + M + 1) / (2 * j1 * (j1 + 1));
  } else if (tJ == tj1 - 2) {
    if (q == 1) cg2 = (j1 - M) * (j1 - M + 1) / (2 * j1 * (2 * j1 + 1));
    else if (q == 0) cg2 = (j1 - M) * (j1 + M) / (j1 * (2 * j1 + 1));
    else cg2 = (j1 + M + 1) * (j1 + M) / (2 * j1 * (2 * j1 + 1));
  } else {
    return 0.0;  // triangle rule
  }
  return cg2 / (tJ + 1.0);
}

double zeeman_strength(int two_Ju, int two_Jl, int pol, int64_t n) {  // lbl_zeeman.cpp:261-277
  if (pol == POL_NO) return 1.0;
  const int tml = -two_Jl + 2 * static_cast<int>(n);
  const int tmu = tml + 2 * zeeman_dM(pol);
  if (std::abs(tml) > two_Jl || std::abs(tmu) > two_Ju) return 0.0;
  const double C = pol == POL_PI ? 1.5 : 0.75;  // polarization_factor, lbl_zeeman.h:154-162
  return C * threej_sq_j2_1(two_Jl, two_Ju, tml, zeeman_dM(pol));
}

double zeeman_splitting(double gu, double gl, int two_Ju, int two_Jl, int pol, int64_t n) {  // lbl_zeeman.h:342-352
  (void)two_Ju;
  if (pol == POL_NO) return 0.0;
  constexpr double C = cst::bohr_magneton / cst::h;
  const int tml = -two_Jl + 2 * static_cast<int>(n);
  const int tmu = tml + 2 * zeeman_dM(pol);
  return C * (0.5 * tmu * gu - 0.5 * tml * gl);
}

void norm_view(int pol, const double mag[3], const double los[2], double npm[7]) {  // lbl_zeeman.cpp:321-331,413-455
  const double deg = cst::pi / 180;
  const double u = mag[0], v = mag[1], w = mag[2];
  const double sa = std::sin(los[1] * deg), ca = std::cos(los[1] * deg);
  const double sz = std::sin(los[0] * deg), cz = std::cos(los[0] * deg);
  const double H    = std::hypot(u, v, w);
  const double uct  = sz * sa * u + sz * ca * v + cz * w;
  const double duct = u * sa * cz + v * ca * cz - w * sz;
  const double theta = H == 0 ? 0 : std::acos(uct / H);
  const double eta   = -std::atan2(ca * u - sa * v, -duct);
  const double CT    = std::cos(theta);
  const double ST    = std::sin(theta);
  const double ST2   = ST * ST;
  const double Q     = ST2 * std::cos(2 * eta);
  const double U     = ST2 * std::sin(2 * eta);
  switch (pol) {
    case POL_PI: { const double t[7] = {ST2, -Q, U, 0, 0, U, Q}; std::copy(t, t + 7, npm); } break;
    case POL_SM: { const double t[7] = {2 - ST2, Q, -U, 2 * CT, -2 * CT, -U, -Q}; std::copy(t, t + 7, npm); } break;
    case POL_SP: { const double t[7] = {2 - ST2, Q, -U, -2 * CT, 2 * CT, -U, -Q}; std::copy(t, t + 7, npm); } break;
    default: { const double t[7] = {1, 0, 0, 0, 0, 0, 0}; std::copy(t, t + 7, npm); } break;
  }
}

// dnorm_view_du / dv / dw, lbl_zeeman.cpp:457-536 with magnetic_angles::dtheta_d*, deta_d* (:361-411)
void dnorm_view(int pol, int comp, const double mag[3], const double los[2], double dnpm[7]) {
  const double deg = cst::pi / 180;
  const double u = mag[0], v = mag[1], w = mag[2];
  const double sa = std::sin(los[1] * deg), ca = std::cos(los[1] * deg);
  const double sz = std::sin(los[0] * deg), cz = std::cos(los[0] * deg);
  const double H    = std::hypot(u, v, w);
  const double uct  = sz * sa * u + sz * ca * v + cz * w;
  const double duct = u * sa * cz + v * ca * cz - w * sz;
  const double theta = H == 0 ? 0 : std::acos(uct / H);
  const double eta   = -std::atan2(ca * u - sa * v, -duct);
  const double rat   = (uct / H) * (uct / H);
  const double H2 = H * H, H3 = H * H * H;
  const double nom   = comp == 0 ? u * uct - sa * sz * H2 : comp == 1 ? v * uct - ca * sz * H2 : w * uct - cz * H2;
  const double dtheta = (H == 0.0 || rat == 1.0) ? 0 : nom / (std::sqrt(1.0 - rat) * H3);
  const double b2 = ca * u - sa * v;
  const double den = b2 * b2 + duct * duct;
  const double deta = H == 0 ? 0 : (comp == 0 ? (cz * v - ca * sz * w) : comp == 1 ? (sa * sz * w - cz * u) : sz * b2) / den;
  const double CT = std::cos(theta), ST = std::sin(theta), CE = std::cos(2 * eta), SE = std::sin(2 * eta);
  const double ST2  = ST * ST;
  const double dST2 = 2 * dtheta * ST * CT;
  const double dQ   = 2 * dtheta * ST * CE * CT - 2 * deta * SE * ST2;
  const double dU   = 2 * deta * ST2 * CE + 2 * dtheta * SE * ST * CT;
  const double dCT  = -dtheta * ST;
  switch (pol) {
    case POL_PI: { const double t[7] = {dST2, -dQ, dU, 0, 0, dU, dQ}; std::copy(t, t + 7, dnpm); } break;
    case POL_SM: { const double t[7] = {-dST2, dQ, -dU, 2 * dCT, -2 * dCT, -dU, -dQ}; std::copy(t, t + 7, dnpm); } break;
    case POL_SP: { const double t[7] = {-dST2, dQ, -dU, -2 * dCT, 2 * dCT, -dU, -dQ}; std::copy(t, t + 7, dnpm); } break;
    default: std::fill(dnpm, dnpm + 7, 0.0); break;
  }
}

bool wind_factor(const double wind[3], const double los[2], double* fac_out, double* jac_out) {
  const double deg = cst::pi / 180;
  const double u = wind[0], v = wind[1], w = wind[2];
  double za = 180 - los[0], aa = los[1] + 180;  // path::mirror
  if (aa > 180) aa -= 360;
  const double u2v2 = u * u + v * v;
  const double w2   = w * w;
  const double f2   = u2v2 + w2;
  const double f    = std::sqrt(f2);
  const double za_f = f == w ? 0.0 : std::acos(w / f);
  const double aa_f = std::atan2(u, v);
  const double za_p = za * deg, aa_p = aa * deg;
  const double czaf = std::cos(za_f), szaf = std::sin(za_f), czap = std::cos(za_p), szap = std::sin(za_p);
  const double caa  = std::cos(aa_f - aa_p);
  const double dp   = czaf * czap + szaf * szap * caa;
  double fac        = 1.0 - (f * dp) / cst::c;
  if (fac <= 0) return false;
  if (std::isnan(fac)) {  // "Zero shift if nan", :40-44
    *fac_out = 1.0;
    if (jac_out) jac_out[0] = jac_out[1] = jac_out[2] = 0.0;
    return true;
  }
  *fac_out = fac;
  if (jac_out) {
    // freq_wind_shift_jac, :56-82 (with its values at zero wind: df = 1, every other derivative 0)
    const double scl = f2 * std::sqrt(f2 - w2);
    const double saa = std::sin(aa_p - aa_f);
    const double comp[3] = {u, v, w};
    for (int c = 0; c < 3; c++) {
      const double df = (f == 0) ? 1.0 : comp[c] / f;
      double dczaf, dszaf, dcaa;
      if (c < 2) {
        dczaf = (f2 == 0) ? 0.0 : (-w * df / f2);
        dszaf = (f2 == w2) ? 0.0 : (w2 * df / scl);
        dcaa  = (u2v2 == 0) ? 0.0 : ((c == 0 ? v : -u) * saa / u2v2);
      } else {
        dczaf = (f2 == 0) ? 0.0 : (-w * df / f2 + 1.0 / f);
        dszaf = (scl == 0) ? 0.0 : ((w2 * df - f * w) / scl);
        dcaa  = 0.0;
      }
      const double ddp = c < 2 ? czap * dczaf + szap * caa * dszaf + szap * szaf * dcaa : czap * dczaf + szap * caa * dszaf;
      jac_out[c] = -(dp * df + f * ddp) / cst::c;
      jac_out[c] /= fac;
    }
  }
  return true;
}

template <typename T>
static int upload(T** dst, const T* src, size_t n) {
  *dst = nullptr;
  if (n == 0) return 0;
  AB_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), n * sizeof(T)));
  AB_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace ab200

using namespace ab200;

ab200_catalog::~ab200_catalog() {
  cudaFree(d_f0); cudaFree(d_a); cudaFree(d_e0); cudaFree(d_gu); cudaFree(d_T0); cudaFree(d_line_isot);
  cudaFree(d_ls_offset); cudaFree(d_ls_species); cudaFree(d_ls_type); cudaFree(d_ls_X);
  cudaFree(d_isot_species); cudaFree(d_isot_mass);
  cudaFree(d_sub_parent); cudaFree(d_sub_Sz); cudaFree(d_sub_dzc);
  cudaFree(d_tile_count); cudaFree(d_sub_cut); cudaFree(d_tile_mode); cudaFree(d_sub_flags);
}

extern "C" int ab200_zeeman_components(int on, double gu, double gl, int two_Ju, int two_Jl, int pol, int64_t cap,
                                       double* strength, double* splitting) {
  const int64_t n = zeeman_size(on != 0, two_Jl, pol);
  for (int64_t i = 0; i < n && i < cap; i++) {
    strength[i]  = zeeman_strength(two_Ju, two_Jl, pol, i);
    splitting[i] = zeeman_splitting(gu, gl, two_Ju, two_Jl, pol, i);
  }
  return static_cast<int>(n);
}

extern "C" int ab200_norm_view(int pol, const double* mag, const double* los, double* npm) {
  norm_view(pol, mag, los, npm);
  return 0;
}

extern "C" int ab200_catalog_create(const ab200_catalog_desc* d, ab200_catalog** out) {
  if (!d || !out) return set_error(AB200_ERR_INVALID, "ab200_catalog_create: null argument");
  *out = nullptr;
  if (d->n_species <= 0 || d->n_isot <= 0 || d->n_bands < 0 || d->n_lines < 0)
    return set_error(AB200_ERR_INVALID, "ab200_catalog_create: negative or zero sizes");
  for (int b = 0; b < d->n_bands; b++) {
    if (d->band_lineshape[b] != AB200_LINESHAPE_VP_LTE && d->band_lineshape[b] != AB200_LINESHAPE_VP_LTE_MIRROR)
      return set_error(AB200_ERR_UNSUPPORTED, "band " + std::to_string(b) +
                                                  ": only the VP_LTE and VP_LTE_MIRROR line shapes are on the GPU path (no CPU fallback)");
    if (d->band_lineshape[b] == AB200_LINESHAPE_VP_LTE_MIRROR && d->band_cutoff_type[b] != AB200_CUTOFF_NONE)
      return set_error(AB200_ERR_UNSUPPORTED, "band " + std::to_string(b) +
                                                  ": VP_LTE_MIRROR with a cutoff is outside the GPU path (the window of the mirror "
                                                  "image follows its parent line, lbl_lineshape_voigt_lte_mirrored.cpp:591-608)");
    if (d->band_isot[b] < 0 || d->band_isot[b] >= d->n_isot)
      return set_error(AB200_ERR_INVALID, "band " + std::to_string(b) + ": isotopologue index out of range");
    if (d->band_offset[b + 1] < d->band_offset[b])
      return set_error(AB200_ERR_INVALID, "band_offset must be non-decreasing");
    if (d->band_cutoff_type[b] == AB200_CUTOFF_BYLINE) {
      if (!(d->band_cutoff_value[b] >= 0))
        return set_error(AB200_ERR_INVALID, "band " + std::to_string(b) + ": negative cutoff");
      for (int64_t l = d->band_offset[b] + 1; l < d->band_offset[b + 1]; l++)
        if (d->f0[l] < d->f0[l - 1])
          return set_error(AB200_ERR_INVALID, "band " + std::to_string(b) +
                                                  ": lines of a band with cutoff must be sorted by f0 (lbl_data.cpp:61-68)");
    } else if (d->band_cutoff_type[b] != AB200_CUTOFF_NONE) {
      return set_error(AB200_ERR_UNSUPPORTED, "unknown cutoff type");
    }
  }
  if (d->n_bands > 0 && d->band_offset[d->n_bands] != d->n_lines)
    return set_error(AB200_ERR_INVALID, "band_offset[n_bands] != n_lines");
  for (int i = 0; i < d->n_isot; i++)
    if (d->isot_species[i] < 0 || d->isot_species[i] >= d->n_species || !(d->isot_mass[i] > 0))
      return set_error(AB200_ERR_INVALID, "isotopologue " + std::to_string(i) + ": bad species id or mass");
  for (int64_t i = 0; i < d->n_ls; i++) {
    if (d->ls_species[i] != AB200_SPECIES_BATH && (d->ls_species[i] < 0 || d->ls_species[i] >= d->n_species))
      return set_error(AB200_ERR_INVALID, "broadener species id out of range");
    for (int v = 0; v < AB200_NVAR; v++) {
      const int t = d->ls_type[i * AB200_NVAR + v];
      if (t < AB200_TM_ABSENT || t > AB200_TM_POLY) return set_error(AB200_ERR_INVALID, "unknown temperature model");
    }
  }

  auto cat = new ab200_catalog();
  static std::atomic<uint64_t> next_serial{1};
  cat->serial = next_serial.fetch_add(1);
  cudaError_t e = cudaGetDevice(&cat->device);
  if (e != cudaSuccess) {
    delete cat;
    return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
  }
  cat->n_species = d->n_species;
  cat->n_isot    = d->n_isot;
  cat->n_bands   = d->n_bands;
  cat->n_lines   = d->n_lines;
  cat->n_ls      = d->n_ls;
  cat->isot_species.assign(d->isot_species, d->isot_species + d->n_isot);
  cat->isot_mass.assign(d->isot_mass, d->isot_mass + d->n_isot);

  // per-line isotopologue + "simple band" classification
  std::vector<int32_t> line_isot(d->n_lines);
  std::vector<char> band_simple(d->n_bands, 1);
  for (int b = 0; b < d->n_bands; b++) {
    for (int64_t l = d->band_offset[b]; l < d->band_offset[b + 1]; l++) {
      line_isot[l] = d->band_isot[b];
      if (d->z_on[l] || !(d->a[l] >= 0) || !(d->gu[l] >= 0)) band_simple[b] = 0;
      for (int64_t i = d->ls_offset[l]; i < d->ls_offset[l + 1]; i++)
        if (d->ls_type[i * AB200_NVAR + AB200_VAR_Y] != AB200_TM_ABSENT ||
            d->ls_type[i * AB200_NVAR + AB200_VAR_G] != AB200_TM_ABSENT)
          band_simple[b] = 0;
    }
  }

  std::vector<int64_t> sub_parent;
  std::vector<double> sub_Sz, sub_dzc, sub_cut;
  std::vector<uint8_t> tile_mode, sub_flags;
  constexpr double INF = std::numeric_limits<double>::infinity();
  auto line_cutoff = [&](int64_t l) {  // ByLine cutoff of the band line l belongs to
    const int64_t b = std::upper_bound(d->band_offset, d->band_offset + d->n_bands + 1, l) - d->band_offset - 1;
    return d->band_cutoff_type[b] == AB200_CUTOFF_BYLINE ? d->band_cutoff_value[b] : INF;
  };
  // VP_LTE_MIRROR (lbl_lineshape_voigt_lte_mirrored.cpp:220): F(f) = w(z(f)) + w(zm(f)), zm = inv_gd (f + f0') + i z_imag.
  // The mirror image is the same sub-line centred at -f0': every slot of a mirrored band gets a twin slot whose centre
  // the prepare kernel negates (SUB_TWIN); it sorts to the negative end of a merged segment and is always far.
  cat->line_tiles.assign(static_cast<size_t>(d->n_lines) * 8, -1);
  cat->line_target_ok.assign(static_cast<size_t>(d->n_lines), 0);
  for (int32_t b = 0; b < d->n_bands; b++)
    if (d->band_lineshape[b] == AB200_LINESHAPE_VP_LTE && d->band_cutoff_type[b] == AB200_CUTOFF_NONE)
      std::fill(cat->line_target_ok.begin() + d->band_offset[b], cat->line_target_ok.begin() + d->band_offset[b + 1], uint8_t{1});
  auto band_of = [&](int64_t l) {
    return std::upper_bound(d->band_offset, d->band_offset + d->n_bands + 1, l) - d->band_offset - 1;
  };
  std::vector<uint8_t> fl;  // per entry of the segment under construction: SUB_* flags
  auto push_mirror_twins = [&](std::vector<int64_t>& par, std::vector<double>& sz, std::vector<double>& dz) {
    const size_t n = par.size();
    fl.assign(n, 0);
    for (size_t i = 0; i < n; i++) {
      if (d->band_lineshape[band_of(par[i])] != AB200_LINESHAPE_VP_LTE_MIRROR) continue;
      fl[i] = SUB_MIRRORED;
      par.push_back(par[i]); sz.push_back(sz[i]); dz.push_back(dz[i]);
      fl.push_back(SUB_MIRRORED | SUB_TWIN);
    }
  };
  auto close_segment = [&](Segment seg, std::vector<int64_t>& par, std::vector<double>& sz, std::vector<double>& dz) {
    if (par.empty()) return;
    const int64_t n_real = static_cast<int64_t>(par.size());
    push_mirror_twins(par, sz, dz);
    // sort by catalog f0 (sub-lines of one parent stay adjacent: stable); mirror images by -f0
    std::vector<int64_t> order(par.size());
    std::iota(order.begin(), order.end(), int64_t{0});
    auto key = [&](int64_t i) { return (fl[i] & SUB_TWIN) ? -d->f0[par[i]] : d->f0[par[i]]; };
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key(a) < key(b); });
    seg.nsub       = static_cast<int64_t>(par.size());
    seg.tile_begin = cat->ntiles;
    const int64_t nt = (seg.nsub + TL - 1) / TL;
    for (int64_t t = 0; t < nt; t++) {
      const int64_t lo = t * TL, hi = std::min<int64_t>(seg.nsub, lo + TL);
      cat->tile_count.push_back(static_cast<int32_t>(hi - lo));
      tile_mode.push_back(static_cast<uint8_t>(seg.mode));
      for (int64_t i = lo; i < lo + TL; i++) {
        if (i < hi) {
          if (!(fl[order[i]] & SUB_TWIN)) {
            int64_t* lt = cat->line_tiles.data() + (par[order[i]] * 4 + seg.pol) * 2;
            if (lt[0] < 0) lt[0] = cat->ntiles + t;
            lt[1] = cat->ntiles + t + 1;
          }
          sub_parent.push_back(par[order[i]]);
          sub_Sz.push_back(sz[order[i]]);
          sub_dzc.push_back(dz[order[i]]);
          sub_cut.push_back(line_cutoff(par[order[i]]));
          sub_flags.push_back(fl[order[i]]);
        } else {
          sub_parent.push_back(-1);
          sub_Sz.push_back(0.0);
          sub_dzc.push_back(0.0);
          sub_cut.push_back(INF);
          sub_flags.push_back(0);
        }
      }
    }
    cat->ntiles += nt;
    seg.tile_end = cat->ntiles;
    cat->counts[seg.pol] += n_real;  // sub-lines of the reference; mirror twins are an implementation detail
    cat->segments.push_back(seg);
    par.clear();
    sz.clear();
    dz.clear();
  };

  std::vector<int64_t> par;
  std::vector<double> sz, dz;
  // (1) merged simple bands, one segment per species, pol = no
  for (int s = 0; s < d->n_species; s++) {
    for (int b = 0; b < d->n_bands; b++) {
      if (!band_simple[b] || d->isot_species[d->band_isot[b]] != s) continue;
      for (int64_t l = d->band_offset[b]; l < d->band_offset[b + 1]; l++) {
        par.push_back(l);
        sz.push_back(1.0);
        dz.push_back(0.0);
      }
    }
    Segment seg{};
    seg.band = -1; seg.isot = -1; seg.species = s; seg.pol = POL_NO; seg.mode = 0; seg.has_cutoff = 0;
    seg.cutoff = 0.0;  // largest cutoff of the merged lines (+inf as soon as one line has none): tile skipping only
    for (int64_t l : par) {
      const double c = line_cutoff(l);
      seg.cutoff = std::max(seg.cutoff, c);
      if (c < INF) seg.has_cutoff = 1;  // at least one line carries a window
    }
    if (par.empty()) seg.cutoff = INF;
    close_segment(seg, par, sz, dz);
  }
  // (2) every other band: one segment per polarisation, in the reference's order
  //     no, pi, sm, sp (lbl_lineshape.cpp:190-208)
  for (int pol : {POL_NO, POL_PI, POL_SM, POL_SP}) {
    for (int b = 0; b < d->n_bands; b++) {
      if (band_simple[b]) continue;
      for (int64_t l = d->band_offset[b]; l < d->band_offset[b + 1]; l++) {
        const bool on = d->z_on[l] != 0;
        if (!((on && pol != POL_NO) || (!on && pol == POL_NO))) continue;  // lines_push_back :376-377
        const int64_t nz = zeeman_size(on, d->two_Jl[l], pol);
        for (int64_t iz = 0; iz < nz; iz++) {
          const double S = zeeman_strength(d->two_Ju[l], d->two_Jl[l], pol, iz);
          if (S == 0.0) continue;  // popped by zeeman_push_back :354-357 (s == 0)
          par.push_back(l);
          sz.push_back(S);
          dz.push_back(zeeman_splitting(d->z_gu[l], d->z_gl[l], d->two_Ju[l], d->two_Jl[l], pol, iz));
        }
      }
      Segment seg{};
      seg.band = b; seg.isot = d->band_isot[b]; seg.species = d->isot_species[d->band_isot[b]];
      seg.pol = pol; seg.mode = 1;
      seg.has_cutoff = d->band_cutoff_type[b] == AB200_CUTOFF_BYLINE;
      seg.cutoff = seg.has_cutoff ? d->band_cutoff_value[b] : std::numeric_limits<double>::infinity();
      close_segment(seg, par, sz, dz);
    }
  }

  int rc = 0;
  rc = rc ? rc : upload(&cat->d_f0, d->f0, d->n_lines);
  rc = rc ? rc : upload(&cat->d_a, d->a, d->n_lines);
  rc = rc ? rc : upload(&cat->d_e0, d->e0, d->n_lines);
  rc = rc ? rc : upload(&cat->d_gu, d->gu, d->n_lines);
  rc = rc ? rc : upload(&cat->d_T0, d->T0, d->n_lines);
  rc = rc ? rc : upload(&cat->d_line_isot, line_isot.data(), line_isot.size());
  rc = rc ? rc : upload(&cat->d_ls_offset, d->ls_offset, static_cast<size_t>(d->n_lines + 1));
  rc = rc ? rc : upload(&cat->d_ls_species, d->ls_species, d->n_ls);
  rc = rc ? rc : upload(&cat->d_ls_type, d->ls_type, static_cast<size_t>(d->n_ls) * AB200_NVAR);
  rc = rc ? rc : upload(&cat->d_ls_X, d->ls_X, static_cast<size_t>(d->n_ls) * AB200_NVAR * 4);
  rc = rc ? rc : upload(&cat->d_isot_species, d->isot_species, d->n_isot);
  rc = rc ? rc : upload(&cat->d_isot_mass, d->isot_mass, d->n_isot);
  rc = rc ? rc : upload(&cat->d_sub_parent, sub_parent.data(), sub_parent.size());
  rc = rc ? rc : upload(&cat->d_sub_Sz, sub_Sz.data(), sub_Sz.size());
  rc = rc ? rc : upload(&cat->d_sub_dzc, sub_dzc.data(), sub_dzc.size());
  rc = rc ? rc : upload(&cat->d_tile_count, cat->tile_count.data(), cat->tile_count.size());
  rc = rc ? rc : upload(&cat->d_sub_cut, sub_cut.data(), sub_cut.size());
  rc = rc ? rc : upload(&cat->d_tile_mode, tile_mode.data(), tile_mode.size());
  rc = rc ? rc : upload(&cat->d_sub_flags, sub_flags.data(), sub_flags.size());
  if (rc) {
    delete cat;
    return rc;
  }
  *out = cat;
  return AB200_OK;
}

extern "C" void ab200_catalog_destroy(ab200_catalog* cat) { delete cat; }

extern "C" int ab200_catalog_counts(const ab200_catalog* cat, int64_t counts[4]) {
  if (!cat || !counts) return set_error(AB200_ERR_INVALID, "ab200_catalog_counts: null argument");
  for (int i = 0; i < 4; i++) counts[i] = cat->counts[i];
  return AB200_OK;
}
