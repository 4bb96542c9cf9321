// lookup.hpp — absorption lookup tables: device view and kernel parameters (lookup.cu)
#pragma once

#include "common.cuh"

struct ab200_lookup;

namespace ab200 {

constexpr int LUT_MAXP = 8;  // stencil points per dimension on the device: interpolation orders up to 7 (the defaults)

struct LutDev {
  int32_t n_tables;
  const int32_t* meta;    // [n_tables][8]: species, nf, np, nt, nw, do_t, do_w, unused
  const int64_t* off;     // [n_tables][7]: f_grid, log_p_grid, t_pert, w_pert, t_atmref, water_atmref, xsec (pool offsets)
  const double* pool;
};

struct LutParams {
  LutDev t;
  int64_t nf;
  const double* f;
  int64_t f_stride;
  const double* ffac;
  const double *T, *P, *vmr;
  int32_t n_species, h2o_species, select_species;
  double* K;
  double* dK;
  int64_t k_pitch;
  int32_t nq;
  int32_t tg_kind[AB200_MAX_TARGETS], tg_species[AB200_MAX_TARGETS];
  double tg_d[AB200_MAX_TARGETS];
  int32_t no_neg, po, to, wo, fo;
  double extpol;
  int* flags;  // bit 16: a coordinate outside the extrapolation limits of a table grid
};

int launch_lookup(const LutParams& p, int nlev, cudaStream_t stream);
// xsec [np][nf] = K [np][k_pitch][7].A / (vmr[lev][species] P / (k T)): the table constructor's division, lookup_map.cpp:108-111
int launch_xsec_from_K(int np, int64_t nf, const double* K, int64_t k_pitch, const double* T, const double* P, const double* vmr,
                       int n_species, int species, double* xsec, cudaStream_t stream);
LutDev lut_dev(const ab200_lookup* l);
int lut_device(const ab200_lookup* l);
// host-side checks of a call against the tables: 0 or an error code with the message set
int lut_check_call(const ab200_lookup* l, int32_t n_species, int32_t h2o_species, int32_t select_species, int po, int to, int wo,
                   int fo);

}  // namespace ab200
