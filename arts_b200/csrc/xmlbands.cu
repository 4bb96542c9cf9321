// AbsorptionBands XML straight into the SoA of ab200_catalog_desc.  Host code.
//
// The reference reads  <Map type="AbsorptionBand" key="QuantumIdentifier" nelem="N">  as N pairs of
//   <QuantumIdentifier version="1"> ISOTOPOLOGUE  KEY upper lower ... </QuantumIdentifier>
//   <AbsorptionBand lineshape=".." cutoff_type=".." cutoff_value=".." nelem="n"> n lines </AbsorptionBand>
// (xml_io_stream<AbsorptionBand>::read, src/core/lbl/lbl_data.cpp:435-470) and every line as a whitespace-separated
// token stream (operator>>(line) lbl_data.cpp:52-58):
//   f0 a e0 gu gl | on gu gl (zeeman::model, lbl_zeeman.cpp:311-319) |
//   T0 n_species { SPECIES n_vars { VAR TYPE [n for POLY] X... } } (lbl_lineshape_model.cpp:260-296,
//   lbl_temperature_model.cpp:28-43, model_size lbl_temperature_model.h:17-33) | n_qn { KEY upper lower } (quantum.cc:150-163)
// into an unordered_map of bands of vectors of lines of maps of maps.  Here the same token stream fills the flat arrays
// directly: one band per <AbsorptionBand> in file order, the broadeners of a line in the order of the file.
//
// Names are resolved through two caller tables (the shim knows SpeciesEnum / SpeciesIsotope): isotopologue tags
// ("H2O-161") and broadener names ("Nitrogen", "N2", "Bath", ...).  Only what the GPU path can represent is accepted:
// variables G0 D0 DV Y G (a G2 / D2 / FVC / ETA entry with non-zero coefficients is AB200_ERR_UNSUPPORTED, an all-zero
// one is dropped like model::clear_zeroes), POLY with at most four coefficients, and the line's local J when Zeeman is on.
#include <algorithm>
#include <charconv>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <string_view>
#include <vector>

#include "common.cuh"

struct ab200_xml_catalog {
  ab200_catalog_desc desc{};
  std::vector<int32_t> isot_species, band_isot, band_lineshape, band_cutoff_type, ls_species, ls_type, two_Ju, two_Jl;
  std::vector<double> isot_mass, band_cutoff_value, f0, a, e0, gu, gl, T0, z_gu, z_gl, ls_X;
  std::vector<int64_t> band_offset, ls_offset;
  std::vector<uint8_t> z_on;
};

namespace ab200 {
namespace {

struct Cursor {
  const char* p;
  const char* e;
  void skip_ws() {
    while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
  }
  // next whitespace-delimited token; inside a band it stops in front of '<'
  bool token(std::string_view& out) {
    skip_ws();
    if (p >= e || *p == '<') return false;
    const char* b = p;
    while (p < e && !(*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r' || *p == '<')) p++;
    out = std::string_view(b, static_cast<size_t>(p - b));
    return true;
  }
  // "<name attr="v" ...>" -> name, raw attribute text; false at the end of the text
  bool tag(std::string_view& name, std::string_view& attrs) {
    skip_ws();
    if (p >= e || *p != '<') return false;
    const char* b = ++p;
    while (p < e && *p != '>') p++;
    if (p >= e) return false;
    std::string_view all(b, static_cast<size_t>(p - b));
    p++;
    size_t i = 0;
    while (i < all.size() && all[i] != ' ' && all[i] != '\t' && all[i] != '\n') i++;
    name  = all.substr(0, i);
    attrs = i < all.size() ? all.substr(i) : std::string_view{};
    return true;
  }
};

bool attribute(std::string_view attrs, std::string_view key, std::string_view& value) {
  size_t pos = 0;
  while ((pos = attrs.find(key, pos)) != std::string_view::npos) {
    const size_t after = pos + key.size();
    const bool starts  = pos == 0 || attrs[pos - 1] == ' ' || attrs[pos - 1] == '\t' || attrs[pos - 1] == '\n';
    if (starts && after + 1 < attrs.size() && attrs[after] == '=' && attrs[after + 1] == '"') {
      const size_t q = attrs.find('"', after + 2);
      if (q == std::string_view::npos) return false;
      value = attrs.substr(after + 2, q - after - 2);
      return true;
    }
    pos = after;
  }
  return false;
}

bool to_double(std::string_view s, double& v) {  // double_imanip: plain numbers, inf and nan
  if (s.empty()) return false;
  const char* b = s.data();
  const char* e = s.data() + s.size();
  if (*b == '+') b++;
  auto r = std::from_chars(b, e, v);
  return r.ec == std::errc() && r.ptr == e;
}
bool to_int(std::string_view s, int64_t& v) {
  auto r = std::from_chars(s.data(), s.data() + s.size(), v);
  return r.ec == std::errc() && r.ptr == s.data() + s.size();
}
// Rational "p" or "p/q" -> 2 p / q, which must be an integer for J
bool to_two_J(std::string_view s, int32_t& out) {
  const size_t sl = s.find('/');
  int64_t p = 0, q = 1;
  if (!to_int(s.substr(0, sl), p)) return false;
  if (sl != std::string_view::npos && !to_int(s.substr(sl + 1), q)) return false;
  if (q == 0 || (2 * p) % q != 0) return false;
  out = static_cast<int32_t>(2 * p / q);
  return true;
}

int var_index(std::string_view v) {  // -2: a variable the reference knows and this path does not; -1: unknown
  if (v == "G0") return AB200_VAR_G0;
  if (v == "D0") return AB200_VAR_D0;
  if (v == "DV") return AB200_VAR_DV;
  if (v == "Y") return AB200_VAR_Y;
  if (v == "G") return AB200_VAR_G;
  if (v == "G2" || v == "D2" || v == "FVC" || v == "ETA") return -2;
  return -1;
}
int model_index(std::string_view t, int& n) {  // model_size, lbl_temperature_model.h:17-33 (POLY: n follows in the file)
  struct M { const char* name; int id; int n; };
  static const M tab[] = {{"T0", AB200_TM_T0, 1}, {"T1", AB200_TM_T1, 2}, {"T2", AB200_TM_T2, 3}, {"T3", AB200_TM_T3, 2},
                          {"T4", AB200_TM_T4, 3}, {"T5", AB200_TM_T5, 2}, {"AER", AB200_TM_AER, 4}, {"DPL", AB200_TM_DPL, 4},
                          {"POLY", AB200_TM_POLY, -1}};
  for (const M& m : tab)
    if (t == m.name) { n = m.n; return m.id; }
  return -100;
}

int fail_at(const char* what, int64_t band, int64_t line) {
  return set_error(AB200_ERR_INVALID, std::string("Error reading AbsorptionBand ") + std::to_string(band) + " line " +
                                          std::to_string(line) + ": " + what);
}

int build(const char* text, int64_t len, const ab200_xml_isotopologue* isots, int32_t n_isot, const ab200_xml_species* names,
          int32_t n_names, int32_t n_species, ab200_xml_catalog** out) {
  if (!out) return set_error(AB200_ERR_INVALID, "ab200_xml_read_bands: out is null");
  *out = nullptr;
  if (!text || len < 0 || !isots || n_isot <= 0 || (n_names > 0 && !names) || n_species <= 0)
    return set_error(AB200_ERR_INVALID, "ab200_xml_read_bands: null or empty argument");
  std::unique_ptr<ab200_xml_catalog> c(new ab200_xml_catalog);
  for (int i = 0; i < n_isot; i++) {
    if (!isots[i].name || isots[i].species < 0 || isots[i].species >= n_species || !(isots[i].mass > 0))
      return set_error(AB200_ERR_INVALID, "ab200_xml_read_bands: bad isotopologue table entry " + std::to_string(i));
    c->isot_species.push_back(isots[i].species);
    c->isot_mass.push_back(isots[i].mass);
  }
  Cursor cur{text, text + len};
  std::string_view name, attrs, v;
  // prologue: <?xml ..?>, <arts ..>, then the map
  int64_t n_bands = -1;
  while (cur.tag(name, attrs)) {
    if (name == "Map") {
      if (!attribute(attrs, "type", v) || v != "AbsorptionBand")
        return set_error(AB200_ERR_INVALID, "ab200_xml_read_bands: the Map does not hold AbsorptionBand values");
      if (!attribute(attrs, "nelem", v) || !to_int(v, n_bands) || n_bands < 0)
        return set_error(AB200_ERR_INVALID, "ab200_xml_read_bands: bad nelem on the Map tag");
      break;
    }
    if (name != "?xml" && name != "arts") return set_error(AB200_ERR_INVALID, "ab200_xml_read_bands: unexpected tag <" + std::string(name) + ">");
  }
  if (n_bands < 0) return set_error(AB200_ERR_INVALID, "ab200_xml_read_bands: no <Map type=\"AbsorptionBand\"> found");

  c->band_offset.push_back(0);
  {  // size hints (untouched reserve costs nothing; the vectors still grow past them): one allocation per array instead
     // of doubling copies.  In the reference's files a line takes ~150 characters and up, a broadener entry ~45
    const size_t nl_max = static_cast<size_t>(len) / 120 + 16, ns_max = static_cast<size_t>(len) / 45 + 16;
    for (auto* v : {&c->f0, &c->a, &c->e0, &c->gu, &c->gl, &c->T0, &c->z_gu, &c->z_gl}) v->reserve(nl_max);
    c->z_on.reserve(nl_max); c->two_Ju.reserve(nl_max); c->two_Jl.reserve(nl_max); c->ls_offset.reserve(nl_max + 1);
    c->ls_species.reserve(ns_max); c->ls_type.reserve(ns_max * AB200_NVAR); c->ls_X.reserve(ns_max * AB200_NVAR * 4);
  }
  for (int64_t ib = 0; ib < n_bands; ib++) {
    if (!cur.tag(name, attrs) || name != "QuantumIdentifier") return fail_at("expected <QuantumIdentifier>", ib, -1);
    if (!cur.token(v)) return fail_at("empty QuantumIdentifier", ib, -1);
    int isot = -1;
    for (int i = 0; i < n_isot; i++)
      if (v == isots[i].name) { isot = i; break; }
    if (isot < 0) return fail_at(("unknown isotopologue " + std::string(v)).c_str(), ib, -1);
    while (cur.token(v)) {}  // the band's global quantum numbers are not needed on the path
    if (!cur.tag(name, attrs) || name != "/QuantumIdentifier") return fail_at("expected </QuantumIdentifier>", ib, -1);

    if (!cur.tag(name, attrs) || name != "AbsorptionBand") return fail_at("expected <AbsorptionBand>", ib, -1);
    int64_t nl = -1;
    double cutv = 0;
    if (!attribute(attrs, "lineshape", v)) return fail_at("no lineshape attribute", ib, -1);
    c->band_lineshape.push_back(v == "VP_LTE" ? AB200_LINESHAPE_VP_LTE : v == "VP_LTE_MIRROR" ? AB200_LINESHAPE_VP_LTE_MIRROR : AB200_LINESHAPE_OTHER);
    if (!attribute(attrs, "cutoff_type", v) || (v != "None" && v != "ByLine")) return fail_at("bad cutoff_type attribute", ib, -1);
    c->band_cutoff_type.push_back(v == "ByLine" ? AB200_CUTOFF_BYLINE : AB200_CUTOFF_NONE);
    if (!attribute(attrs, "cutoff_value", v) || !to_double(v, cutv)) return fail_at("bad cutoff_value attribute", ib, -1);
    c->band_cutoff_value.push_back(cutv);
    if (!attribute(attrs, "nelem", v) || !to_int(v, nl) || nl < 0) return fail_at("bad nelem attribute", ib, -1);
    c->band_isot.push_back(isot);

    for (int64_t il = 0; il < nl; il++) {
      double head[5], zg[2], T0;
      int64_t on = 0, nsp = 0, nqn = 0;
      for (double& h : head)
        if (!cur.token(v) || !to_double(v, h)) return fail_at("bad f0 / a / e0 / gu / gl", ib, il);
      if (!cur.token(v) || !to_int(v, on)) return fail_at("bad Zeeman switch", ib, il);
      for (double& g : zg)
        if (!cur.token(v) || !to_double(v, g)) return fail_at("bad Zeeman g value", ib, il);
      if (!cur.token(v) || !to_double(v, T0)) return fail_at("bad T0", ib, il);
      if (!cur.token(v) || !to_int(v, nsp) || nsp < 0) return fail_at("bad broadener count", ib, il);
      c->f0.push_back(head[0]); c->a.push_back(head[1]); c->e0.push_back(head[2]); c->gu.push_back(head[3]); c->gl.push_back(head[4]);
      c->z_on.push_back(on != 0); c->z_gu.push_back(zg[0]); c->z_gl.push_back(zg[1]); c->T0.push_back(T0);
      c->ls_offset.push_back(static_cast<int64_t>(c->ls_species.size()));
      for (int64_t is = 0; is < nsp; is++) {
        if (!cur.token(v)) return fail_at("missing broadener name", ib, il);
        int sp = INT32_MIN;
        for (int k = 0; k < n_names; k++)
          if (v == names[k].name) { sp = names[k].species; break; }
        if (sp == INT32_MIN || (sp != AB200_SPECIES_BATH && (sp < 0 || sp >= n_species)))
          return fail_at(("unknown broadener " + std::string(v)).c_str(), ib, il);
        int64_t nv = 0;
        if (!cur.token(v) || !to_int(v, nv) || nv < 0) return fail_at("bad variable count", ib, il);
        c->ls_species.push_back(sp);
        const size_t e = c->ls_species.size() - 1;
        c->ls_type.resize((e + 1) * AB200_NVAR, AB200_TM_ABSENT);
        c->ls_X.resize((e + 1) * AB200_NVAR * 4, 0.0);
        for (int64_t iv = 0; iv < nv; iv++) {
          if (!cur.token(v)) return fail_at("missing variable name", ib, il);
          const int var = var_index(v);
          if (var == -1) return fail_at(("unknown line-shape variable " + std::string(v)).c_str(), ib, il);
          if (!cur.token(v)) return fail_at("missing temperature model", ib, il);
          int n = 0;
          const int tm = model_index(v, n);
          if (tm == -100) return fail_at(("unknown temperature model " + std::string(v)).c_str(), ib, il);
          if (n < 0) {
            int64_t np = 0;
            if (!cur.token(v) || !to_int(v, np) || np < 0) return fail_at("bad POLY size", ib, il);
            n = static_cast<int>(np);
          }
          double X[4] = {0, 0, 0, 0};
          bool nonzero = false;
          for (int k = 0; k < n; k++) {
            double x;
            if (!cur.token(v) || !to_double(v, x)) return fail_at("bad model coefficient", ib, il);
            nonzero |= x != 0;
            if (k < 4) X[k] = x;
            else if (x != 0) {
              set_error(AB200_ERR_UNSUPPORTED, "POLY with more than four coefficients is not on the GPU path");
              return AB200_ERR_UNSUPPORTED;
            }
          }
          if (var == -2) {
            if (nonzero) return set_error(AB200_ERR_UNSUPPORTED, "line-shape variables G2, D2, FVC and ETA are not used by VP_LTE and not on the GPU path");
            continue;
          }
          c->ls_type[e * AB200_NVAR + var] = tm;
          std::copy(X, X + 4, c->ls_X.begin() + (e * AB200_NVAR + var) * 4);
        }
      }
      // local quantum numbers: only J is used (Zeeman pattern, lbl_zeeman.cpp:296-303)
      if (!cur.token(v) || !to_int(v, nqn) || nqn < 0) return fail_at("bad quantum number count", ib, il);
      int32_t tJu = 0, tJl = 0;
      bool have_J = false;
      for (int64_t iq = 0; iq < nqn; iq++) {
        std::string_view key, up, lo;
        if (!cur.token(key) || !cur.token(up) || !cur.token(lo)) return fail_at("truncated quantum numbers", ib, il);
        if (key == "J") {
          if (!to_two_J(up, tJu) || !to_two_J(lo, tJl)) return fail_at("bad J value", ib, il);
          have_J = true;
        }
      }
      if (on != 0 && !have_J) return fail_at("Zeeman is on but the line has no local J", ib, il);  // qn.at(J) throws there
      c->two_Ju.push_back(tJu); c->two_Jl.push_back(tJl);
    }
    c->band_offset.push_back(static_cast<int64_t>(c->f0.size()));
    if (!cur.tag(name, attrs) || name != "/AbsorptionBand") return fail_at("expected </AbsorptionBand> (nelem does not match the lines)", ib, nl);
  }
  c->ls_offset.push_back(static_cast<int64_t>(c->ls_species.size()));

  ab200_catalog_desc& d = c->desc;
  d.n_species = n_species; d.n_isot = n_isot; d.n_bands = static_cast<int32_t>(n_bands);
  d.n_lines = static_cast<int64_t>(c->f0.size()); d.n_ls = static_cast<int64_t>(c->ls_species.size());
  d.isot_species = c->isot_species.data(); d.isot_mass = c->isot_mass.data();
  d.band_isot = c->band_isot.data(); d.band_lineshape = c->band_lineshape.data();
  d.band_cutoff_type = c->band_cutoff_type.data(); d.band_cutoff_value = c->band_cutoff_value.data();
  d.band_offset = c->band_offset.data();
  d.f0 = c->f0.data(); d.a = c->a.data(); d.e0 = c->e0.data(); d.gu = c->gu.data(); d.gl = c->gl.data(); d.T0 = c->T0.data();
  d.z_on = c->z_on.data(); d.z_gu = c->z_gu.data(); d.z_gl = c->z_gl.data(); d.two_Ju = c->two_Ju.data(); d.two_Jl = c->two_Jl.data();
  d.ls_offset = c->ls_offset.data(); d.ls_species = c->ls_species.data(); d.ls_type = c->ls_type.data(); d.ls_X = c->ls_X.data();
  *out = c.release();
  return AB200_OK;
}

}  // namespace
}  // namespace ab200

extern "C" {

int ab200_xml_read_bands(const char* text, int64_t len, const ab200_xml_isotopologue* isotopologues, int32_t n_isot,
                         const ab200_xml_species* names, int32_t n_names, int32_t n_species, ab200_xml_catalog** out) {
  return ab200::build(text, len, isotopologues, n_isot, names, n_names, n_species, out);
}

int ab200_xml_read_bands_file(const char* filename, const ab200_xml_isotopologue* isotopologues, int32_t n_isot,
                              const ab200_xml_species* names, int32_t n_names, int32_t n_species, ab200_xml_catalog** out) {
  if (!filename) return ab200::set_error(AB200_ERR_INVALID, "ab200_xml_read_bands_file: null file name");
  std::FILE* f = std::fopen(filename, "rb");
  if (!f) return ab200::set_error(AB200_ERR_INVALID, std::string("Cannot open file: ") + filename);
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<char> buf(static_cast<size_t>(std::max<long>(n, 0)));
  const size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
  std::fclose(f);
  if (got != buf.size()) return ab200::set_error(AB200_ERR_INVALID, std::string("Cannot read file: ") + filename);
  return ab200::build(buf.data(), static_cast<int64_t>(buf.size()), isotopologues, n_isot, names, n_names, n_species, out);
}

const ab200_catalog_desc* ab200_xml_desc(const ab200_xml_catalog* cat) { return cat ? &cat->desc : nullptr; }
void ab200_xml_destroy(ab200_xml_catalog* cat) { delete cat; }

}  // extern "C"
