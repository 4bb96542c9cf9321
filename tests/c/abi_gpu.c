/* The GPU entry points of the drop-in boundary from plain C, the way the reference-side shim calls them:
 * ab200_catalog_create + ab200_clearsky_emission with HOST buffers (forward and with T + VMR Jacobian rows),
 * ab200_propmat_levels accumulating into the caller's K, and the brightness-temperature transform.  The results are
 * checked against the CPU oracle (oracle/_ref/liboracle.so, linked here as the CHECKER only): propmat <= 1e-9 relative,
 * Tb <= 1e-6 K, Jacobian rows <= 2e-7 of the column maximum; a second call must return the same bits.
 * Compiled as C99 by tests/test_gpu_c_abi.py; exit code 0 and "abi gpu ok" on success. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "arts_b200.h"

/* the oracle's entry points used here (oracle/oracle.cpp; same flattened structs) */
int orc_clearsky_emission(const ab200_catalog_desc *d, int64_t nf, const double *f, int64_t f_level_stride,
                          const ab200_atm_path *atm, int32_t select_species, int32_t no_negative_absorption, int32_t nq,
                          const ab200_target *targets, const double *r, int32_t hse_derivative, int32_t rte_option,
                          const double *I_bkg, uint32_t flags, double *I, double *dI, double *K_out);
int orc_planck_tb(int64_t nf, const double *f, double *I);
const char *orc_last_error(void);

#define NL 300 /* lines */
#define NF 1200
#define NP 12
#define NS 2 /* species: 0 absorber, 1 foreign broadener */
#define NQ 2

static uint64_t rng_state = 88172645463325252ULL;
static double urand(void) { /* xorshift64*: the same inputs on every run */
  rng_state ^= rng_state >> 12;
  rng_state ^= rng_state << 25;
  rng_state ^= rng_state >> 27;
  return (double)((rng_state * 2685821657736338717ULL) >> 11) / 9007199254740992.0;
}
static int cmp_double(const void *a, const void *b) {
  const double x = *(const double *)a, y = *(const double *)b;
  return x < y ? -1 : x > y;
}
static double planck(double f, double T) {
  const double h = 6.62607015e-34, k = 1.380649e-23, c = 299792458.0;
  return 2.0 * h * f * f * f / (c * c) / expm1(h * f / (k * T));
}

int main(void) {
  static double f0[NL], a[NL], e0[NL], gu[NL], gl[NL], T0[NL], z_g[NL], lsX[NL * 2 * AB200_NVAR * 4];
  static uint8_t z_on[NL];
  static int32_t twoJ[NL], ls_species[NL * 2], ls_type[NL * 2 * AB200_NVAR];
  static int64_t ls_offset[NL + 1];
  static double f[NF], I_bkg[NF * 4], I[NF * 4], I2[NF * 4], Io[NF * 4], dI[NF * NP * NQ * 4], dIo[NF * NP * NQ * 4];
  static double K[NP * NF * 7], Ko[NP * NF * 7], K2[NP * NF * 7];
  static double Tl[NP], Pl[NP], vmr[NP * NS], isorat[NP], Q[NP], dQ[NP], r[NP - 1];
  int i, j, v, q;

  for (i = 0; i < NL; i++) f0[i] = 100e9 + 30e9 * urand();
  qsort(f0, NL, sizeof(double), cmp_double);
  for (i = 0; i < NL; i++) {
    a[i] = 4.479e-9 * (1.0 + urand());
    e0[i] = 1e-23 * (1.0 + urand());
    gu[i] = gl[i] = 3.0;
    T0[i] = 296.0;
    ls_offset[i] = 2 * i;
    ls_species[2 * i] = 0;
    ls_species[2 * i + 1] = AB200_SPECIES_BATH;
    for (j = 0; j < 2; j++) {
      for (v = 0; v < AB200_NVAR; v++) ls_type[(2 * i + j) * AB200_NVAR + v] = AB200_TM_ABSENT;
      ls_type[(2 * i + j) * AB200_NVAR + AB200_VAR_G0] = AB200_TM_T1;
      lsX[((2 * i + j) * AB200_NVAR + AB200_VAR_G0) * 4 + 0] = 1e4 + 2e4 * urand();
      lsX[((2 * i + j) * AB200_NVAR + AB200_VAR_G0) * 4 + 1] = 0.5 + 0.5 * urand();
      ls_type[(2 * i + j) * AB200_NVAR + AB200_VAR_D0] = AB200_TM_T1;
      lsX[((2 * i + j) * AB200_NVAR + AB200_VAR_D0) * 4 + 0] = -500.0 + 1000.0 * urand();
      lsX[((2 * i + j) * AB200_NVAR + AB200_VAR_D0) * 4 + 1] = 0.7;
    }
  }
  ls_offset[NL] = 2 * NL;

  const int32_t isot_species[1] = {0}, band_isot[1] = {0}, band_ls[1] = {AB200_LINESHAPE_VP_LTE}, band_ct[1] = {AB200_CUTOFF_NONE};
  const double isot_mass[1] = {31.9898}, band_cv[1] = {0.0};
  const int64_t band_offset[2] = {0, NL};
  ab200_catalog_desc d;
  memset(&d, 0, sizeof d);
  d.n_species = NS; d.n_isot = 1; d.n_bands = 1; d.n_lines = NL; d.n_ls = 2 * NL;
  d.isot_species = isot_species; d.isot_mass = isot_mass;
  d.band_isot = band_isot; d.band_lineshape = band_ls; d.band_cutoff_type = band_ct; d.band_cutoff_value = band_cv;
  d.band_offset = band_offset;
  d.f0 = f0; d.a = a; d.e0 = e0; d.gu = gu; d.gl = gl; d.T0 = T0;
  d.z_on = z_on; d.z_gu = z_g; d.z_gl = z_g; d.two_Ju = twoJ; d.two_Jl = twoJ;
  d.ls_offset = ls_offset; d.ls_species = ls_species; d.ls_type = ls_type; d.ls_X = lsX;

  for (i = 0; i < NP; i++) {
    const double z = 2.5 * i; /* km */
    Tl[i] = 288.0 - 6.0 * z + 0.12 * z * z;
    Pl[i] = 101325.0 * exp(-z / 7.0);
    vmr[i * NS + 0] = 0.21;
    vmr[i * NS + 1] = 0.01 * exp(-z / 2.0);
    isorat[i] = 0.995;
    Q[i] = 215.0 * Tl[i] / 296.0;
    dQ[i] = 215.0 / 296.0;
    if (i < NP - 1) r[i] = 2500.0;
  }
  ab200_atm_path atm;
  memset(&atm, 0, sizeof atm);
  atm.np = NP; atm.T = Tl; atm.P = Pl; atm.vmr = vmr; atm.isorat = isorat; atm.Q = Q; atm.dQdT = dQ;
  for (j = 0; j < NF; j++) {
    f[j] = 100e9 + 30e9 * j / (NF - 1.0);
    I_bkg[4 * j] = planck(f[j], 288.0);
    I_bkg[4 * j + 1] = I_bkg[4 * j + 2] = I_bkg[4 * j + 3] = 0.0;
  }

  if (ab200_device_count() < 1) { printf("no CUDA device: %s\n", ab200_last_error()); return 10; }
  ab200_catalog *cat = NULL;
  if (ab200_catalog_create(&d, &cat) != AB200_OK) { printf("catalog_create: %s\n", ab200_last_error()); return 1; }

  /* forward, K returned */
  if (ab200_clearsky_emission(cat, NF, f, 0, &atm, AB200_SPECIES_BATH, 1, 0, NULL, r, 0, AB200_RTE_LINSRC, I_bkg,
                              AB200_FLAG_RETURN_K, I, NULL, K) != AB200_OK) { printf("clearsky: %s\n", ab200_last_error()); return 2; }
  if (orc_clearsky_emission(&d, NF, f, 0, &atm, AB200_SPECIES_BATH, 1, 0, NULL, r, 0, AB200_RTE_LINSRC, I_bkg, 0, Io, NULL, Ko)) {
    printf("oracle: %s\n", orc_last_error()); return 3;
  }
  double kmax = 0.0, worst = 0.0;
  for (i = 0; i < NP * NF * 7; i++) kmax = fmax(kmax, fabs(Ko[i]));
  for (i = 0; i < NP * NF * 7; i++) {
    const double e = fabs(K[i] - Ko[i]) / fmax(fabs(Ko[i]), 1e-12 * kmax);
    worst = fmax(worst, e);
  }
  if (!(kmax > 0.0) || worst > 1e-9) { printf("propmat off by %.3e relative\n", worst); return 4; }
  memcpy(I2, I, sizeof I);
  if (ab200_planck_tb(NF, f, I) != AB200_OK || orc_planck_tb(NF, f, Io)) { printf("planck_tb\n"); return 5; }
  double tb_err = 0.0;
  for (j = 0; j < NF; j++) tb_err = fmax(tb_err, fabs(I[4 * j] - Io[4 * j]));
  if (!(tb_err <= 1e-6) || !(I[0] > 100.0 && I[0] < 300.0)) { printf("Tb off by %.3e K (Tb[0] = %g)\n", tb_err, I[0]); return 6; }

  /* T + VMR rows; a second call returns the same bits */
  const ab200_target tg[NQ] = {{AB200_TARGET_T, 0, 0, 0, 0}, {AB200_TARGET_VMR, 0, 0, 0, 0}};
  for (int rep = 0; rep < 2; rep++) {
    if (ab200_clearsky_emission(cat, NF, f, 0, &atm, AB200_SPECIES_BATH, 1, NQ, tg, r, 1, AB200_RTE_LINSRC, I_bkg, 0,
                                rep ? Io : I, rep ? dIo : dI, NULL) != AB200_OK) { printf("clearsky jac: %s\n", ab200_last_error()); return 7; }
  }
  if (memcmp(I, Io, sizeof I) || memcmp(dI, dIo, sizeof dI) || memcmp(I, I2, sizeof I)) { printf("repeat call differs\n"); return 8; }
  if (orc_clearsky_emission(&d, NF, f, 0, &atm, AB200_SPECIES_BATH, 1, NQ, tg, r, 1, AB200_RTE_LINSRC, I_bkg, 0, Io, dIo, NULL)) {
    printf("oracle jac: %s\n", orc_last_error()); return 3;
  }
  for (q = 0; q < NQ; q++) {
    double cmax = 0.0, e = 0.0;
    for (i = 0; i < NF * NP; i++) cmax = fmax(cmax, fabs(dIo[(i * NQ + q) * 4]));
    for (i = 0; i < NF * NP; i++) e = fmax(e, fabs(dI[(i * NQ + q) * 4] - dIo[(i * NQ + q) * 4]));
    if (!(cmax > 0.0) || e > 2e-7 * cmax) { printf("jacobian row %d off by %.3e of the column maximum\n", q, e / cmax); return 9; }
  }

  /* spectral_propmatAddLines semantics: += into the caller's K (one level) */
  ab200_atm_path one = atm;
  one.np = 1;
  for (i = 0; i < NF * 7; i++) K2[i] = 0.25;
  if (ab200_propmat_levels(cat, NF, f, 0, &one, AB200_SPECIES_BATH, 1, 0, NULL, 0, K2, NULL) != AB200_OK) {
    printf("propmat_levels: %s\n", ab200_last_error()); return 11;
  }
  for (j = 0; j < NF; j++)
    if (fabs((K2[7 * j] - 0.25) - K[7 * j]) > 1e-9 * fabs(K[7 * j]) + 1e-25 || K2[7 * j + 1] != 0.25) { printf("+= semantics\n"); return 12; }

  /* one host process, several devices: two workers (on device 0 twice when the box has one GPU), levels
   * dealt over them for the line sum, K transposed between them, the Stokes chain per frequency slice - the bits of one device */
  {
    const int32_t devs[2] = {0, ab200_device_count() > 1 ? 1 : 0};
    ab200_multi *mu = NULL;
    if (ab200_clearsky_emission(cat, NF, f, 0, &atm, AB200_SPECIES_BATH, 1, 0, NULL, r, 0, AB200_RTE_LINSRC, I_bkg, 0, Io, NULL, NULL) != AB200_OK) {
      printf("clearsky: %s\n", ab200_last_error()); return 13;
    }
    if (ab200_multi_create(&d, 2, devs, &mu) != AB200_OK || ab200_multi_device_count(mu) != 2) { printf("multi_create: %s\n", ab200_last_error()); return 14; }
    if (ab200_multi_clearsky_emission(mu, NF, f, 0, &atm, AB200_SPECIES_BATH, 1, 0, NULL, r, 0, AB200_RTE_LINSRC, I_bkg, 0, I, NULL, NULL) != AB200_OK) {
      printf("multi clearsky: %s\n", ab200_last_error()); return 15;
    }
    for (i = 0; i < NF * 4; i++)
      if (I[i] != Io[i]) { printf("multi-device radiance differs from the one-device radiance at %ld\n", (long)i); return 16; }
    for (i = 0; i < NF * 7; i++) K2[i] = 0.0;
    if (ab200_multi_propmat_levels(mu, NF, f, 0, &one, AB200_SPECIES_BATH, 1, 0, NULL, AB200_FLAG_K_ZERO_INIT, K2, NULL) != AB200_OK) {
      printf("multi propmat: %s\n", ab200_last_error()); return 17;
    }
    for (j = 0; j < NF; j++)
      if (K2[7 * j] != K[7 * j]) { printf("multi-device propmat differs at %ld\n", (long)j); return 18; }
    ab200_multi_destroy(mu);
  }

  ab200_catalog_destroy(cat);
  printf("abi gpu ok: propmat %.2e, Tb %.2e K\n", worst, tb_err);
  return 0;
}
