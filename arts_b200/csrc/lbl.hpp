// lbl.hpp — kernel parameter blocks and launchers of stage 1 (lbl.cu).
#pragma once

#include "catalog.hpp"

namespace ab200 {

struct PrepareParams {
  // catalog (device)
  const double *f0, *a, *e0, *gu, *T0;
  const int32_t* line_isot;
  const int64_t* ls_offset;
  const int32_t* ls_species;
  const int32_t* ls_type;
  const double* ls_X;
  const int32_t* isot_species;
  const double* isot_mass;
  const int64_t* sub_parent;
  const double *sub_Sz, *sub_dzc;
  const double* sub_cut;     // [slot] ByLine cutoff [Hz], +inf if none
  const uint8_t* tile_mode;  // [tile] 0: real merged segment (record slot 9 holds the line's cutoff), 1: complex
  const uint8_t* sub_flags;  // [slot] SUB_MIRRORED | SUB_TWIN (catalog.hpp)
  int32_t n_species, n_isot;
  int64_t ntiles;
  // levels of this batch (device, already offset to the first level of the batch)
  const double *T, *P, *vmr, *isorat, *Q, *H;
  const double* frange;  // [nlev][2] first / last frequency of the level's grid
  // outputs
  double* prep;     // [nlev][ntiles][N_GROUPS][TL][4]
  double* summary;  // [nlev][ntiles][4]
  int* flags;       // bit 0: negative G0, bit 1: non-finite shape parameter
};

struct SumParams {
  const double* f;   // [nlev][nf] (f_stride = nf) or [nf] (f_stride = 0), offset to the batch
  int64_t f_stride;
  int64_t nf;
  const double* ffac;   // [nlev] wind-shift factor of each level's grid (1 without wind)
  const double *T, *P;  // [nlev]
  const double* npm;    // [nlev][4][7] zeeman::norm_view per polarisation
  const double* prep;
  const double* summary;
  const int32_t* tile_count;
  const uint8_t* tile_mode;
  int64_t ntiles;
  const SegmentDev* segs;  // segments selected for this launch (device)
  int32_t nsegs;
  int32_t no_negative_absorption;
  double* K;  // [nlev][k_pitch][7], offset to the batch
  int64_t k_pitch;
  int32_t debug_skip_near;  // measurement only (AB200_DEBUG_SKIP_NEAR=1): near tiles contribute nothing
  int32_t k_store_full;  // real kernel: K is not initialised; write whole records {A,0,0,0,0,0,0} with vector stores
};

// Jacobian targets (lbl_jac.cu)
// far-field sums (lbl_fmm.cu): moment records of the four cluster levels for one batch of levels, and scratch
struct FmmBuffers {
  double* L0;       // [nlev][ntiles][16][MOM_DOUBLES] 16-line clusters
  double* L1;       // [nlev][ntiles][4][MOM_DOUBLES]  64-line clusters
  double* L2;       // [nlev][ntiles][MOM_DOUBLES]     tiles
  double* L3;       // [nlev][ngroups][MOM_DOUBLES]    groups of FMM_GROUP tiles
  double* scan;     // [nlev][ntiles][2] running bounds of the tiles' acceptance intervals
  double* far_acc;  // [segments][nlev][k_pitch] far-field part of each segment's sum
  int64_t ngroups;
};
int launch_fmm(const PrepareParams& pp, const SumParams& sp, const FmmBuffers& fb, const int32_t* tile_seg, int nlev, int store_full,
               cudaStream_t stream);

// Jacobian targets as COMPUTED: one entry per distinct derivative record.  The three magnetic-field components are one
// entry (kind AB200_TARGET_MAG_U) and so are the three wind components (AB200_TARGET_WIND_U); JacSumParams::out_row maps
// an entry to the rows of the caller's dK it feeds.
struct JacPrepParams {
  int32_t nq;
  int32_t kind[AB200_MAX_TARGETS];     // AB200_TARGET_*
  int32_t species[AB200_MAX_TARGETS];  // for VMR targets; line-shape targets: the broadener
  int64_t line[AB200_MAX_TARGETS];     // line targets: the catalog line
  int32_t ls_var[AB200_MAX_TARGETS], coeff[AB200_MAX_TARGETS];  // line-shape targets: AB200_VAR_*, X0..X3
  const double* dQdT;                  // [nlev][n_isot], offset to the batch
  double* jac;                         // [nlev][ntiles][nq][2][TL][4]
  double* jcom;                        // [nlev][ntiles][TL]
};
struct JacSumParams {
  int32_t nq, q0;      // computed targets (stride of the derivative records), first one of this pass
  int32_t nrows;       // rows of dK per level (the caller's targets)
  int32_t out_row[AB200_MAX_TARGETS][3];  // computed target -> dK row; magnetic / wind entries: rows of u, v, w (-1: not requested)
  const double* mag_ratio;  // [nlev][3] mag_c / |mag| of the batch's levels (magnetic-field targets)
  int32_t real_lines;  // the segments of this launch are mode 0: real strengths, pol = no (closed-form far path)
  int32_t skip_vfar;   // real lines: pairs with |x| > VFAR_LIMIT of cutoff-free tiles are summed by lbl_sum_jac_vfar_kernel
  int32_t pair_far;    // near tiles of real lines: far pairs take the closed form (AB200_JAC_PAIR_FAR=0 turns it off, debugging)
  int32_t kind[AB200_MAX_TARGETS];
  const double* jac;
  const double* jcom;
  double* dK;  // [nlev][nq][k_pitch][7], offset to the batch
  int64_t line_tiles[AB200_MAX_TARGETS][4][2];  // line targets: the tiles that hold the line's sub-lines, per polarisation
  const double* dnpm;      // [nlev][3][4][7] dnorm_view_d{u,v,w} per polarisation (magnetic-field targets)
  const double* wind_jac;  // [nlev][3] freq_wind_shift_jac of the batch's levels (wind targets); null = leave d/df
};

int launch_prepare(const PrepareParams& p, int nlev, cudaStream_t stream);
int launch_prepare_jac(const PrepareParams& p, const JacPrepParams& jp, int nlev, cudaStream_t stream);
int launch_sum_jac(const SumParams& p, JacSumParams jp, int nlev, cudaStream_t stream);
int launch_sum(const SumParams& p, int nlev, int mode, cudaStream_t stream);
// region histogram of the evaluations of one batch of levels (measurement helper, see arts_b200.h)
int launch_region_histogram(const SumParams& p, int nlev, int64_t samples_per_level, uint64_t seed, double* d_out,
                            cudaStream_t stream);
int launch_faddeeva(int64_t n, const double* zr, const double* zi, double* wr, double* wi, cudaStream_t stream);
int launch_dfma_peak(int iters, int blocks, double* d_out, cudaStream_t stream);

}  // namespace ab200
