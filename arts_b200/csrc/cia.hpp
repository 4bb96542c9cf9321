// cia.hpp — collision-induced absorption: device view of the tables and the kernel parameters (cia.cu)
#pragma once

#include "common.cuh"

struct ab200_cia;

namespace ab200 {

struct CiaDev {
  int32_t n_records;
  const int32_t *species1, *species2, *ds_begin, *nf, *nT;
  const int64_t *f_off, *T_off, *data_off;
  const double* pool;
};

struct CiaParams {
  CiaDev c;
  int64_t nf;
  const double* f;  // [np][nf] or [nf]
  int64_t f_stride;
  const double* ffac;  // [np] or nullptr
  const double *T, *P;  // [np]
  const double* vmr;    // [np][n_species]
  int32_t n_species, select_species;
  double* K;   // [np][k_pitch][7]
  double* dK;  // [np][nq][k_pitch][7]
  int64_t k_pitch;
  int32_t nq, it;
  int32_t tg_kind[AB200_MAX_TARGETS], tg_species[AB200_MAX_TARGETS];
  double dt, T_extrapolfac;
  int32_t ignore_errors;
  int* flags;  // bit 8: temperature outside the extrapolation range of a data set
};

int launch_cia(const CiaParams& p, int nlev, cudaStream_t stream);
CiaDev cia_dev(const ab200_cia* c);
int cia_device(const ab200_cia* c);
int cia_max_species(const ab200_cia* c);  // largest species index any record names, -1 without records

}  // namespace ab200
