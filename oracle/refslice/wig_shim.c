/* oracle/refslice/wig_shim.c - TEST INFRASTRUCTURE ONLY.
 *
 * Entry point over the reference's vendored wigxjpf (3rdparty/wigner/wigxjpf, compiled where it lies by
 * oracle/Makefile), called the way the reference's wigner3j() calls it (src/core/physics/wigner_functions.cc:41-71:
 * doubled integer arguments, per-call temp storage of 3/2 * the largest argument + 1, WIGNER3 = wig3jj when the
 * fastwigxj tables - the same numbers, precomputed - are not loaded).  The prime-factor table is what
 * make_wigner_ready(:101-121) sets up with wig_table_init(largest, 3); TABLE_TWO_J is larger than any 2J the
 * tests ask for.  Nothing here computes. */
#include <stdlib.h>

#include "wigxjpf.h"

#define TABLE_TWO_J 2000

static int g_table = 0;

static int imax(int a, int b) { return a > b ? a : b; }

int refwig_wigner3j(int tj1, int tj2, int tj3, int tm1, int tm2, int tm3, double* out) {
  const int big = imax(imax(imax(abs(tj1), abs(tj2)), imax(abs(tj3), abs(tm1))), imax(abs(tm2), abs(tm3)));
  if (big > TABLE_TWO_J) return 1;
#pragma omp critical(refwig_table)
  if (!g_table) {
    wig_table_init(TABLE_TWO_J, 3);
    g_table = 1;
  }
  wig_thread_temp_init(big * 3 / 2 + 1);
  *out = wig3jj(tj1, tj2, tj3, tm1, tm2, tm3);
  wig_temp_free();
  return 0;
}
