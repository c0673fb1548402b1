/*
 * ref_glue.c -- TEST INFRASTRUCTURE.  C side of the callback that the reference's Fortran makes into
 * its C++ initial-condition object (initialconditionatpoint_, ICInterface.C:36-57), so that the
 * transliterated setAccelerationBCs4D (oracle/_ref) can be driven from the tests with the same
 * callback type as the hand-written oracle.  The Fortran passes GLOBAL indices; the callback receives
 * 0-based data-box indices.
 */
#include <stdint.h>

typedef double (*ok_ic_fn)(void* ctx, int i1, int i2, int i3, int i4);
static ok_ic_fn g_fn = 0;
static void* g_ctx = 0;
static int g_lo[4] = {0, 0, 0, 0};

void loki_ref_set_ic(ok_ic_fn fn, void* ctx, const int* data_box_lower) {
  g_fn = fn;
  g_ctx = ctx;
  for (int k = 0; k < 4; ++k) g_lo[k] = data_box_lower[k];
}

double initialconditionatpoint_(int64_t* ic, int* i1, int* i2, int* i3, int* i4) {
  (void)ic;
  return g_fn(g_ctx, *i1 - g_lo[0], *i2 - g_lo[1], *i3 - g_lo[2], *i4 - g_lo[3]);
}
