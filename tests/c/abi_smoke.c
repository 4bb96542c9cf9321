/* The drop-in boundary from plain C: include/arts_b200.h must compile as C99 and the host-only entry points must work
 * when called the way a cgo / JNI / ctypes stub would call them (no GPU needed for these). */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "arts_b200.h"

int main(void) {
  /* the reference's HITRAN fixture record (tests/hitran/single_line.par), `par` columns only */
  const char *rec =
      " 11    0.072049 1.875E-30 4.668E-12.09460.391 1922.82890.730.002760          0 1 0          0 1 0  4  2  2        5  "
      "1  5      534253807294713152     9.0   11.0\n";
  ab200_hitran_isotopologue tab[1] = {{1, '1', 0, 18.010565, 0.997317, 174.58}};
  ab200_hitran_catalog *h = NULL;
  int rc = ab200_hitran_read_par(rec, (int64_t)strlen(rec), -INFINITY, INFINITY, AB200_HITRAN_STRENGTH_A, tab, 1, 1, 1, &h);
  if (rc != AB200_OK) { printf("read_par failed: %s\n", ab200_last_error()); return 1; }
  const ab200_catalog_desc *d = ab200_hitran_desc(h);
  if (d->n_lines != 1 || d->n_bands != 1 || fabs(d->f0[0] - 2.15997e9) > 1e5 || d->gu[0] != 9.0) { printf("bad record\n"); return 2; }
  ab200_hitran_destroy(h);
  /* the reference's default option: the Einstein coefficient from the line strength column, within 2e-4 of the file's own */
  rc = ab200_hitran_read_par(rec, (int64_t)strlen(rec), -INFINITY, INFINITY, AB200_HITRAN_STRENGTH_S, tab, 1, 1, 1, &h);
  if (rc != AB200_OK) { printf("read_par S failed: %s\n", ab200_last_error()); return 1; }
  d = ab200_hitran_desc(h);
  if (fabs(d->a[0] / 4.668e-12 - 1.0) > 1e-3) { printf("bad Einstein coefficient from S: %g\n", d->a[0]); return 2; }
  ab200_hitran_destroy(h);

  rc = ab200_hitran_read_par("xx\n", 3, -INFINITY, INFINITY, AB200_HITRAN_STRENGTH_A, tab, 1, 1, 1, &h);
  if (rc != AB200_ERR_INVALID || !strstr(ab200_last_error(), "Unexpected end of string")) { printf("error path: %d %s\n", rc, ab200_last_error()); return 3; }

  const double grid[3] = {100.0, 200.0, 300.0}, q[3] = {10.0, 30.0, 60.0}, T[2] = {150.0, 250.0};
  ab200_partfun_table pt = {AB200_PARTFUN_INTERP, 3, grid, q};
  double Q[2], dQ[2];
  rc = ab200_partfun_eval(&pt, 1, 2, T, Q, dQ);
  if (rc != AB200_OK || Q[0] != 20.0 || Q[1] != 45.0 || dQ[0] != 0.2 || dQ[1] != 0.3) { printf("partfun\n"); return 4; }

  /* the reference's abs_bands XML (first band of its fixture tests/core/nlte/nlte_lines.xml, two of the six broadeners) */
  const char *xml =
      "<?xml version=\"1.0\"?>\n<arts format=\"ascii\" version=\"1\">\n"
      "<Map type=\"AbsorptionBand\" key=\"QuantumIdentifier\" nelem=\"1\">\n"
      "<QuantumIdentifier version=\"1\"> H2O-161 J 1 1 Ka 1 0 Kc 0 1 </QuantumIdentifier>\n"
      "<AbsorptionBand lineshape=\"VP_LTE\" cutoff_type=\"None\" cutoff_value=\"750000000000.0\" nelem=\"1\">\n"
      "556936000000.0 0.003458 4.7266199e-22 9.0 9.0 0 0.0 0.0 300.0 2 Nitrogen 2 G0 T1 30741.117 0.77 D0 T5 1940.0 0.77 "
      "Water 2 G0 T1 143793.866 0.75 D0 T5 0.0 0.75 0\n</AbsorptionBand>\n</Map>\n</arts>\n";
  ab200_xml_isotopologue iso[1] = {{"H2O-161", 0, 18.010565}};
  ab200_xml_species names[2] = {{"Water", 0}, {"Nitrogen", 1}};
  ab200_xml_catalog *x = NULL;
  rc = ab200_xml_read_bands(xml, (int64_t)strlen(xml), iso, 1, names, 2, 2, &x);
  if (rc != AB200_OK) { printf("xml read failed: %s\n", ab200_last_error()); return 6; }
  d = ab200_xml_desc(x);
  if (d->n_lines != 1 || d->n_ls != 2 || d->f0[0] != 556936000000.0 || d->T0[0] != 300.0 || d->ls_species[0] != 1 ||
      d->ls_type[0 * AB200_NVAR + AB200_VAR_G0] != AB200_TM_T1 || d->ls_X[(0 * AB200_NVAR + AB200_VAR_G0) * 4] != 30741.117 ||
      d->ls_type[1 * AB200_NVAR + AB200_VAR_D0] != AB200_TM_T5) { printf("bad xml band\n"); return 7; }
  ab200_xml_destroy(x);
  names[1].name = "Argon";
  rc = ab200_xml_read_bands(xml, (int64_t)strlen(xml), iso, 1, names, 2, 2, &x);
  if (rc != AB200_ERR_INVALID || !strstr(ab200_last_error(), "unknown broadener Nitrogen")) { printf("xml error path: %d %s\n", rc, ab200_last_error()); return 8; }

  /* compute entry points exist and fail loudly instead of falling back when there is no device or no input */
  if (ab200_catalog_create(NULL, NULL) != AB200_ERR_INVALID) { printf("null catalog accepted\n"); return 5; }
  printf("devices visible: %d\n", ab200_device_count());
  printf("abi smoke ok\n");
  return 0;
}
