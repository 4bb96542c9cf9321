// catalog.hpp — device-resident line catalog (flattened AbsorptionBands) and the
// per-path device workspace.  Host-side C++ of the library; not part of the C ABI.
#pragma once

#include <cstdint>
#include <limits>
#include <vector>

#include "common.cuh"

namespace ab200 {

// A segment is the unit the reference clamps and scales on: one (band, polarisation)
// pair (voigt::lte::calculate, lbl_lineshape_voigt_lte.cpp:1652-1692).  Bands whose lines
// have no line mixing and no Zeeman effect have a non-negative real sum — with a ByLine
// cutoff too, because every term ls(f) - ls(f0' + cutoff) is >= 0 inside its window — so
// the clamp can never trigger; all such bands of one species are merged into a single
// segment (mode 0) whose lines are sorted by f0 and carry their own cutoff.
// slot flags
constexpr uint8_t SUB_MIRRORED = 1;  // the slot's band is VP_LTE_MIRROR (its Jacobian uses the frequency-independent zp - zm)
constexpr uint8_t SUB_TWIN = 2;      // the slot is the mirror image: centre -f0'

struct Segment {
  int32_t band;     // band index, or -1 for a merged segment
  int32_t isot;     // isotopologue of the band (-1 if merged over several)
  int32_t species;  // species id (select_species filter, lbl_lineshape.cpp:191)
  int32_t pol;      // Pol
  int32_t mode;     // 0: real-only far/near sum into A;  1: complex sum, 7 components, clamp
  int32_t has_cutoff;
  double cutoff;    // Hz, +inf without cutoff
  int64_t tile_begin, tile_end;
  int64_t nsub;     // (sub-)lines before padding
};

// device view handed to kernels
struct SegmentDev {
  int64_t tile_begin, tile_end;
  double cutoff;
  int32_t pol;
  int32_t has_cutoff;
};

}  // namespace ab200

struct ab200_catalog {
  int device = 0;
  uint64_t serial = 0;  // unique per ab200_catalog_create: cached workspaces are keyed on it, not on the address
  int32_t n_species = 0, n_isot = 0, n_bands = 0;
  int64_t n_lines = 0, n_ls = 0;
  int64_t counts[4] = {0, 0, 0, 0};  // sub-lines per polarisation

  std::vector<int32_t> isot_species;
  std::vector<double> isot_mass;
  std::vector<ab200::Segment> segments;
  std::vector<int64_t> line_tiles;      // [n_lines][4 pol][2]: first and last + 1 tile holding sub-lines of the line (-1, -1: none)
  std::vector<uint8_t> line_target_ok;  // per line: its band is plain VP_LTE without cutoff (line-parameter Jacobians)
  std::vector<int32_t> tile_count;  // real (sub-)lines per tile
  int64_t ntiles = 0;

  // per parent line
  double *d_f0 = nullptr, *d_a = nullptr, *d_e0 = nullptr, *d_gu = nullptr, *d_T0 = nullptr;
  int32_t* d_line_isot = nullptr;
  int64_t* d_ls_offset = nullptr;
  int32_t* d_ls_species = nullptr;
  int32_t* d_ls_type = nullptr;
  double* d_ls_X = nullptr;
  // per isotopologue
  int32_t* d_isot_species = nullptr;
  double* d_isot_mass = nullptr;
  // per (sub-)line slot, padded to ntiles*TL
  int64_t* d_sub_parent = nullptr;  // -1 for padding
  double* d_sub_Sz = nullptr;       // zeeman::model::Strength (1 for pol = no)
  double* d_sub_dzc = nullptr;      // zeeman::model::Splitting [Hz/T]
  // per tile
  int32_t* d_tile_count = nullptr;
  double* d_sub_cut = nullptr;      // per slot: ByLine cutoff of the line's band [Hz] (+inf if none / padding)
  uint8_t* d_tile_mode = nullptr;   // per tile: mode of its segment (0 real, 1 complex)
  uint8_t* d_sub_flags = nullptr;   // per slot: SUB_* (VP_LTE_MIRROR twins)

  ~ab200_catalog();
};

namespace ab200 {
// lbl_zeeman.cpp:261-309 / lbl_zeeman.h:92-136,342-352 on the host (closed-form 3j for j2 = 1)
int64_t zeeman_size(bool on, int two_Jl, int pol);
double zeeman_strength(int two_Ju, int two_Jl, int pol, int64_t n);
double zeeman_splitting(double gu, double gl, int two_Ju, int two_Jl, int pol, int64_t n);
// zeeman::norm_view, lbl_zeeman.cpp:413-455
void norm_view(int pol, const double mag[3], const double los[2], double npm[7]);
// wind_shift frequency factor, src/m_frequency_grid.cc:4-55 with path::mirror (path_point.cpp:33-39); false if <= 0
void dnorm_view(int pol, int comp, const double mag[3], const double los[2], double dnpm[7]);
bool wind_factor(const double wind[3], const double los[2], double* fac, double* jac = nullptr);
}  // namespace ab200
