// predef.hpp — predefined continuum models: kernel parameters (predef.cu)
#pragma once

#include "common.cuh"

namespace ab200 {
struct PredefParams;
// path entry point helper (api.cu): validates and launches on the path's resident arrays
int predef_on_path(const int32_t* models, int32_t n_models, const ab200_predef_species* sp, const double* target_d, int64_t nf,
                   const double* d_f, int64_t f_stride, const double* d_ffac, const double* d_T, const double* d_P, const double* d_vmr,
                   int32_t n_species, int32_t select_species, double* d_K, double* d_dK, int64_t k_pitch, int32_t nq,
                   const int32_t* tg_kind, const int32_t* tg_species, int np, int* d_flags, const double* d_wjac, cudaStream_t stream,
                   const ab200_predef_data* data = nullptr);
}  // namespace ab200
