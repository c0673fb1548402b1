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

/* ------------------------------------------------------------------------------------------------------
 * CPU baseline through the reference's OWN Fortran kernels (transliterated, oracle/_ref): one RK4 stage the
 * way the reference executes it -- zeroSolnData, chargeDensity (ReductionSchedule.C:421-444), periodic fill
 * (ParallelArray.H:580-606), computeadvectionderivatives4D, setphasespacevel4D (materialised vel3/vel4),
 * setaccelerationbcs4D, computeaccelerationderivatives4D, two xpby4d, one copySolnData
 * (RK4Integrator.H:149-171, VPSystem.C:372-476) -- on a periodic box with a synthetic Maxwellian.  The Fortran
 * routines are called with the reference's argument lists; the C++ pieces between them (whole-array zero / copy,
 * the 4D->2D sum, the periodic wrap) are restated here.  One thread: bench.py forks one process per core, each
 * owning an independent sub-box like the reference's MPI ranks.  Returns seconds per stage.
 * ---------------------------------------------------------------------------------------------------- */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "_ref/loki_ref_protos.h"

static double ic_none(void* c, int i1, int i2, int i3, int i4) { (void)c; (void)i1; (void)i2; (void)i3; (void)i4; return 0.0; }

double loki_ref_time_rk4_stage(const int* n, int order, const double* dx, int reps) {
  const int ng = (order == 4) ? 2 : 3;
  const int64_t n1d = n[0] + 2 * ng, n2d = n[1] + 2 * ng, n3d = n[2] + 2 * ng, n4d = n[3] + 2 * ng;
  const int64_t vol = n1d * n2d * n3d * n4d;
  int nd[8], ni[8];
  for (int k = 0; k < 4; ++k) { nd[2 * k] = -ng; nd[2 * k + 1] = n[k] - 1 + ng; ni[2 * k] = 0; ni[2 * k + 1] = n[k] - 1; }
#define B8(b) &b[0], &b[1], &b[2], &b[3], &b[4], &b[5], &b[6], &b[7]
#define IDX(i1, i2, i3, i4) ((i1) + n1d * ((i2) + n2d * ((i3) + n3d * (int64_t)(i4))))
  double* f = (double*)malloc(sizeof(double) * vol);
  double* fold = (double*)malloc(sizeof(double) * vol);
  double* rhs = (double*)malloc(sizeof(double) * vol);
  double* delta = (double*)calloc(vol, sizeof(double));
  double* vxface = (double*)calloc((n3d + 1) * n4d * 2, sizeof(double));
  double* vyface = (double*)calloc(n3d * (n4d + 1) * 2, sizeof(double));
  double* vel1 = (double*)calloc((n1d + 1) * n2d * n3d * n4d, sizeof(double));
  double* vel2 = (double*)calloc((n2d + 1) * n3d * n4d * n1d, sizeof(double));
  double* vel3 = (double*)calloc((n3d + 1) * n4d * n1d * n2d, sizeof(double));
  double* vel4 = (double*)calloc((n4d + 1) * n1d * n2d * n3d, sizeof(double));
  double* accel = (double*)calloc(n1d * n2d * 2, sizeof(double));
  double* rho = (double*)calloc(n1d * n2d, sizeof(double));
  const double vlo = -0.5 * n[2] * dx[2], vlo2 = -0.5 * n[3] * dx[3];
  /* buildVelocityArrays / initializeVelocity, non-relativistic (KineticSpecies.C:1656-1694, 2024-2047) */
  for (int64_t i4 = 0; i4 < n4d; ++i4)
    for (int64_t i3 = 0; i3 <= n3d; ++i3) {
      vxface[i3 + (n3d + 1) * i4] = vlo + (i3 - ng) * dx[2];
      vxface[i3 + (n3d + 1) * (i4 + n4d)] = vlo2 + ((i4 - ng) + 0.5) * dx[3];
    }
  for (int64_t i4 = 0; i4 <= n4d; ++i4)
    for (int64_t i3 = 0; i3 < n3d; ++i3) {
      vyface[i3 + n3d * i4] = vlo + ((i3 - ng) + 0.5) * dx[2];
      vyface[i3 + n3d * (i4 + (n4d + 1))] = vlo2 + (i4 - ng) * dx[3];
    }
  for (int64_t i4 = 0; i4 < n4d; ++i4)
    for (int64_t i3 = 0; i3 < n3d; ++i3) {
      const double vx = vlo + ((i3 - ng) + 0.5) * dx[2], vy = vlo2 + ((i4 - ng) + 0.5) * dx[3];
      for (int64_t i2 = 0; i2 < n2d; ++i2)
        for (int64_t i1 = 0; i1 <= n1d; ++i1) vel1[i1 + (n1d + 1) * (i2 + n2d * (i3 + n3d * i4))] = vx;
      for (int64_t i1 = 0; i1 < n1d; ++i1)
        for (int64_t i2 = 0; i2 <= n2d; ++i2) vel2[i2 + (n2d + 1) * (i3 + n3d * (i4 + n4d * i1))] = vy;
      for (int64_t i2 = 0; i2 < n2d; ++i2)
        for (int64_t i1 = 0; i1 < n1d; ++i1)
          f[IDX(i1, i2, i3, i4)] = exp(-0.5 * (vx * vx + vy * vy)) * (1.0 + 0.1 * cos(0.3 * i1) * cos(0.2 * i2)) / 6.283185307179586;
    }
  memcpy(fold, f, sizeof(double) * vol);
  for (int64_t k = 0; k < n1d * n2d; ++k) { accel[k] = 0.01 * sin(0.1 * (double)k); accel[k + n1d * n2d] = 0.01 * cos(0.07 * (double)k); }
  int lower[4] = {-ng, -ng, -ng, -ng};
  loki_ref_set_ic(ic_none, 0, lower);
  double norm = -1.0, bz = 0.0, b_w = 1e-3, ax, ay;
  int64_t ic = 0;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int r = 0; r < reps; ++r) {
    memset(rhs, 0, sizeof(double) * vol);                                        /* zeroSolnData */
    for (int64_t k = 0; k < n1d * n2d; ++k) rho[k] = 0.0;                         /* chargeDensity */
    for (int i4 = ng; i4 < ng + n[3]; ++i4)
      for (int i3 = ng; i3 < ng + n[2]; ++i3)
        for (int i2 = ng; i2 < ng + n[1]; ++i2)
          for (int i1 = ng; i1 < ng + n[0]; ++i1) rho[i1 + n1d * i2] += f[IDX(i1, i2, i3, i4)];
    for (int64_t k = 0; k < n1d * n2d; ++k) rho[k] *= dx[2] * dx[3];
    for (int64_t k = 0; k < n1d * n2d; ++k) rho[k] *= -1.0;
    for (int64_t row = 0; row < n2d * n3d * n4d; ++row) {                         /* periodic x, then y */
      double* p = f + row * n1d;
      for (int k = 0; k < ng; ++k) { p[k] = p[k + n[0]]; p[ng + n[0] + k] = p[ng + k]; }
    }
    for (int64_t pl = 0; pl < n3d * n4d; ++pl) {
      double* p = f + pl * n1d * n2d;
      for (int k = 0; k < ng; ++k)
        for (int64_t i1 = 0; i1 < n1d; ++i1) {
          p[i1 + n1d * k] = p[i1 + n1d * (k + n[1])];
          p[i1 + n1d * (ng + n[1] + k)] = p[i1 + n1d * (ng + k)];
        }
    }
    computeadvectionderivatives4d_(rhs, f, B8(nd), B8(ni), vel1, vel2, (double*)dx, &order);
    setphasespacevel4d_(vel3, vel4, B8(nd), B8(ni), vxface, vyface, &norm, &bz, accel, &nd[0], &nd[1], &nd[2], &nd[3], &ax, &ay);
    setaccelerationbcs4d_(f, B8(nd), B8(nd), B8(ni), &order, vel3, vel4, &ic);
    computeaccelerationderivatives4d_(rhs, f, B8(nd), B8(ni), vel3, vel4, (double*)dx, &order);
    xpby4d_(delta, rhs, &b_w, B8(nd), B8(ni));                                   /* addSolnData(delta) */
    memcpy(f, fold, sizeof(double) * vol);                                       /* copySolnData */
    xpby4d_(f, rhs, &b_w, B8(nd), B8(ni));                                       /* addSolnData(pred) */
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
#undef B8
#undef IDX
  free(f); free(fold); free(rhs); free(delta); free(vxface); free(vyface);
  free(vel1); free(vel2); free(vel3); free(vel4); free(accel); free(rho);
  return ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)) / reps;
}
