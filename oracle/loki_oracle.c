/*
 * loki_oracle.c -- CPU ORACLE (test infrastructure only; see loki_oracle.h).
 *
 * Hand-written C restatement of the live Fortran/C++ arithmetic of the LLNL/LOKI Vlasov RHS path.
 * Each function cites the reference file:line it restates.  Operation order follows the reference
 * so that `gcc -O2 -ffp-contract=off` reproduces a `gfortran -O2` build bit for bit (the pin test
 * tests/test_oracle_pin.py checks this against the transliterated reference source in oracle/_ref).
 */
#include "loki_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ND(d) (ok_nd(g, (d)))
#define F4(a, i1, i2, i3, i4) (a)[ok_idx(g, (i1), (i2), (i3), (i4))]

static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }

/* The heavy loop nests run on all host cores when built with -fopenmp (oracle/Makefile): only loops whose
 * iterations write disjoint outputs are split, so every output element executes the reference's operation
 * sequence and the results stay bit-identical to the serial run (tests/test_oracle_pin.py compares with the
 * transliterated Fortran).  OMP_NUM_THREADS=1 gives the serial oracle (bench.py's per-core baseline). */
#ifdef _OPENMP
#define OK_PRAGMA(x) _Pragma(#x)
#define OK_PARALLEL_FOR OK_PRAGMA(omp parallel for schedule(static))
#define OK_PARALLEL_FOR2 OK_PRAGMA(omp parallel for collapse(2) schedule(static))
#define OK_PARALLEL_FOR2_MAX(v) OK_PRAGMA(omp parallel for collapse(2) schedule(static) reduction(max : v))
#else
#define OK_PARALLEL_FOR
#define OK_PARALLEL_FOR2
#define OK_PARALLEL_FOR2_MAX(v)
#endif

/* ------------------------------------------------------------------------------------------
 * WENO43Fit4D  (KineticSpeciesF.f:723-790)
 * um2,um1,u0,up1 are f(i-1),f(i),f(i+1),f(i+2) for the face between i and i+1.
 * ------------------------------------------------------------------------------------------ */
double ok_weno43_fit(double um2, double um1, double u0, double up1, double vel) {
  const double eps = 1.e-10;
  double tmp = 1.0 / 6.0;
  double fl = tmp * (-um2 + 5.0 * um1 + 2.0 * u0);
  double fr = tmp * (2.0 * um1 + 5.0 * u0 - up1);

  double c1l = u0 - 2.0 * um1 + um2;
  double c2l = u0 - um2;
  double c1r = up1 - 2.0 * u0 + um1;
  double c2r = up1 - um1;
  /* Fortran evaluates a*b*c left to right and x**2 as x*x */
  double bl = 8.0 * tmp * (c1l * c1l) + 0.5 * c1l * c2l + 0.25 * (c2l * c2l);
  double br = 8.0 * tmp * (c1r * c1r) - 0.5 * c1r * c2r + 0.25 * (c2r * c2r);

  double al = 1.0 / ((eps + bl) * (eps + bl));
  double ar = 1.0 / ((eps + br) * (eps + br));
  tmp = 1.0 / (al + ar);
  double wl = tmp * al;
  double wr = tmp * ar;

  /* Henrick mapping */
  al = wl * (0.75 + wl * (wl - 1.5));
  ar = wr * (0.75 + wr * (wr - 1.5));
  tmp = 1.0 / (al + ar);
  wl = tmp * al;
  wr = tmp * ar;

  double wmax = dmax(wl, wr);
  double wmin = dmin(wl, wr);
  if (vel > 0.0) {
    wl = wmax;
    wr = wmin;
  } else {
    wl = wmin;
    wr = wmax;
  }
  return (wl * fl + wr * fr);
}

/* ------------------------------------------------------------------------------------------
 * WENO65Fit4D  (KineticSpeciesF.f:914-979).  The smoothness indicators are Maple-generated
 * quadratic forms; term order and the divisions by constants are kept as written.
 * ------------------------------------------------------------------------------------------ */
double ok_weno65_fit(double um3, double um2, double um1, double u0, double up1, double up2, double vel) {
  const double eps = 1.e-10;
  double fl = (2.0 * um3 - 13.0 * um2 + 47.0 * um1 + 27.0 * u0 - 3.0 * up1) / 60.0;
  double fr = (-3.0 * um2 + 27.0 * um1 + 47.0 * u0 - 13.0 * up1 + 2.0 * up2) / 60.0;

  double bl = 0.5489E4 / 0.105E3 * (um1 * um1) +
              (-0.2242428E7 * u0 - 0.1887108E7 * um2 + 0.410226E6 * um3 + 0.557646E6 * up1) * um1 / 0.30240E5 +
              0.75329E5 / 0.3780E4 * (um2 * um2) +
              (0.1259696E7 * u0 - 0.275318E6 * um3 - 0.302534E6 * up1) * um2 / 0.30240E5 +
              0.33727E5 / 0.30240E5 * (um3 * um3) +
              (-0.264314E6 * u0 + 0.61952E5 * up1) * um3 / 0.30240E5 +
              0.106409E6 / 0.3780E4 * (u0 * u0) -
              0.227749E6 / 0.15120E5 * u0 * up1 +
              0.69217E5 / 0.30240E5 * (up1 * up1);

  double br = 0.106409E6 / 0.3780E4 * (um1 * um1) +
              (-0.2242428E7 * u0 - 0.455498E6 * um2 + 0.1259696E7 * up1 - 0.264314E6 * up2) * um1 / 0.30240E5 +
              0.69217E5 / 0.30240E5 * (um2 * um2) +
              (0.557646E6 * u0 - 0.302534E6 * up1 + 0.61952E5 * up2) * um2 / 0.30240E5 +
              0.75329E5 / 0.3780E4 * (up1 * up1) +
              (-0.1887108E7 * u0 - 0.275318E6 * up2) * up1 / 0.30240E5 +
              0.5489E4 / 0.105E3 * (u0 * u0) +
              0.68371E5 / 0.5040E4 * u0 * up2 +
              0.33727E5 / 0.30240E5 * (up2 * up2);

  double al = 1.0 / ((eps + bl) * (eps + bl));
  double ar = 1.0 / ((eps + br) * (eps + br));
  double wl = al / (al + ar);
  double wr = ar / (al + ar);

  al = wl * (0.75 + wl * (wl - 1.5));
  ar = wr * (0.75 + wr * (wr - 1.5));
  wl = al / (al + ar);
  wr = ar / (al + ar);

  double wmax = dmax(wl, wr);
  double wmin = dmin(wl, wr);
  if (vel > 0.0) {
    wl = wmax;
    wr = wmin;
  } else {
    wl = wmin;
    wr = wmax;
  }
  return (wl * fl + wr * fr);
}

void ok_weno43_fit_v(const double* u4, const double* vel, double* face, int64_t count) {
  for (int64_t k = 0; k < count; ++k)
    face[k] = ok_weno43_fit(u4[4 * k], u4[4 * k + 1], u4[4 * k + 2], u4[4 * k + 3], vel[k]);
}
void ok_weno65_fit_v(const double* u6, const double* vel, double* face, int64_t count) {
  for (int64_t k = 0; k < count; ++k)
    face[k] = ok_weno65_fit(u6[6 * k], u6[6 * k + 1], u6[6 * k + 2], u6[6 * k + 3], u6[6 * k + 4],
                            u6[6 * k + 5], vel[k]);
}

double ok_ic_from_tables(void* ctx, int i1, int i2, int i3, int i4) {
  const ok_ic_tables* t = (const ok_ic_tables*)ctx;
  const int64_t pxy = i1 + (int64_t)t->n1d * i2, pv = i3 + (int64_t)t->n3d * i4;
  switch (t->kind) {
    case 1: return t->fnorm * t->fv[pv] * t->fx[pxy] * t->frac;
    case 2: return t->fx[pxy] * t->fv[pv] + t->fx2[pxy] * t->fv2[pv];
    case 4: return t->fv[pv] * t->fx[pxy] * t->fx2[pxy];
    case 3: return t->full[pxy + (int64_t)t->n1d * t->n2d * pv];
    default: return t->fv[pv] * t->fx[pxy];
  }
}

/* xpby4d (KineticSpeciesF.f:10-38): x += b*y on the interior only */
void ok_xpby4d(double* x, const double* y, double b, const ok_geom* g) {
  const int ng = g->ng;
  OK_PARALLEL_FOR2
  for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
    for (int i3 = ng; i3 < ng + g->n[2]; ++i3)
      for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
        for (int i1 = ng; i1 < ng + g->n[0]; ++i1)
          F4(x, i1, i2, i3, i4) = F4(x, i1, i2, i3, i4) + b * F4(y, i1, i2, i3, i4);
}

/* rotated face-array offsets (KineticSpecies.C:1569-1584; KineticSpeciesF.f:65-66) */
static inline int64_t v3idx(const ok_geom* g, int i3, int i4, int i1, int i2) {
  return (((int64_t)i2 * ND(0) + i1) * ND(3) + i4) * (ND(2) + 1) + i3;
}
static inline int64_t v4idx(const ok_geom* g, int i4, int i1, int i2, int i3) {
  return (((int64_t)i3 * ND(1) + i2) * ND(0) + i1) * (ND(3) + 1) + i4;
}
static inline int64_t v1idx(const ok_geom* g, int i1, int i2, int i3, int i4) {
  return (((int64_t)i4 * ND(2) + i3) * ND(1) + i2) * (ND(0) + 1) + i1;
}
static inline int64_t v2idx(const ok_geom* g, int i2, int i3, int i4, int i1) {
  return (((int64_t)i1 * ND(3) + i4) * ND(2) + i3) * (ND(1) + 1) + i2;
}

/* setphasespacevel4D (KineticSpeciesF.f:42-114) */
void ok_set_phase_space_vel_4d(double* vel3, double* vel4, const ok_geom* g, const double* vxface_vel,
                               const double* vyface_vel, double normalization, double bz_const,
                               const double* accel, double* axmax_out, double* aymax_out) {
  const int ng = g->ng;
  const int n1d = (int)ND(0), n2d = (int)ND(1), n3d = (int)ND(2), n4d = (int)ND(3);
  double axmax = 0.0;
  OK_PARALLEL_FOR2_MAX(axmax)
  for (int i4 = 0; i4 < n4d; ++i4)
    for (int i3 = 0; i3 <= n3d; ++i3) {
      double vy = vxface_vel[(int64_t)i3 + (int64_t)(n3d + 1) * (i4 + (int64_t)n4d * 1)];
      for (int i2 = 0; i2 < n2d; ++i2)
        for (int i1 = 0; i1 < n1d; ++i1) {
          double v = accel[i1 + (int64_t)n1d * (i2 + (int64_t)n2d * 0)] + normalization * vy * bz_const;
          vel3[v3idx(g, i3, i4, i1, i2)] = v;
          if (i1 >= ng && i1 < ng + g->n[0] && i2 >= ng && i2 < ng + g->n[1] && i3 >= ng &&
              i3 <= ng + g->n[2] && i4 >= ng && i4 < ng + g->n[3])
            axmax = dmax(axmax, fabs(v));
        }
    }
  double aymax = 0.0;
  OK_PARALLEL_FOR2_MAX(aymax)
  for (int i3 = 0; i3 < n3d; ++i3)
    for (int i2 = 0; i2 < n2d; ++i2)
      for (int i1 = 0; i1 < n1d; ++i1)
        for (int i4 = 0; i4 <= n4d; ++i4) {
          double vx = vyface_vel[(int64_t)i3 + (int64_t)n3d * (i4 + (int64_t)(n4d + 1) * 0)];
          double v = accel[i1 + (int64_t)n1d * (i2 + (int64_t)n2d * 1)] - normalization * vx * bz_const;
          vel4[v4idx(g, i4, i1, i2, i3)] = v;
          if (i1 >= ng && i1 < ng + g->n[0] && i2 >= ng && i2 < ng + g->n[1] && i3 >= ng &&
              i3 < ng + g->n[2] && i4 >= ng && i4 <= ng + g->n[3])
            aymax = dmax(aymax, fabs(v));
        }
  *axmax_out = axmax;
  *aymax_out = aymax;
}

/* setphasespacevelmaxwell4D (KineticSpeciesF.f:118-197) */
void ok_set_phase_space_vel_maxwell_4d(double* vel3, double* vel4, const ok_geom* g,
                                       const double* vxface_vel, const double* vyface_vel,
                                       double normalization, double bz_const, const double* em_vars,
                                       const double* vz, double* axmax_out, double* aymax_out) {
  const int ng = g->ng;
  const int n1d = (int)ND(0), n2d = (int)ND(1), n3d = (int)ND(2), n4d = (int)ND(3);
  const int64_t pl = (int64_t)n1d * n2d;
#define EM(i1, i2, c) em_vars[(i1) + (int64_t)n1d * (i2) + pl * ((c)-1)]
  double axmax = 0.0;
  OK_PARALLEL_FOR2_MAX(axmax)
  for (int i3 = 0; i3 <= n3d; ++i3)
    for (int i4 = 0; i4 < n4d; ++i4) {
      double vy = vxface_vel[(int64_t)i3 + (int64_t)(n3d + 1) * (i4 + (int64_t)n4d * 1)];
      for (int i1 = 0; i1 < n1d; ++i1)
        for (int i2 = 0; i2 < n2d; ++i2) {
          double a = normalization *
                     (EM(i1, i2, 1) + vy * EM(i1, i2, 6) + vy * bz_const - vz[i1 + (int64_t)n1d * i2] * EM(i1, i2, 5));
          vel3[v3idx(g, i3, i4, i1, i2)] = a;
          if (i1 >= ng && i1 < ng + g->n[0] && i2 >= ng && i2 < ng + g->n[1] && i3 >= ng &&
              i3 <= ng + g->n[2] && i4 >= ng && i4 < ng + g->n[3])
            axmax = dmax(axmax, fabs(a));
        }
    }
  double aymax = 0.0;
  OK_PARALLEL_FOR2_MAX(aymax)
  for (int i4 = 0; i4 <= n4d; ++i4)
    for (int i1 = 0; i1 < n1d; ++i1)
      for (int i2 = 0; i2 < n2d; ++i2)
        for (int i3 = 0; i3 < n3d; ++i3) {
          double vx = vyface_vel[(int64_t)i3 + (int64_t)n3d * (i4 + (int64_t)(n4d + 1) * 0)];
          double a = normalization * (EM(i1, i2, 2) + vz[i1 + (int64_t)n1d * i2] * EM(i1, i2, 4) -
                                      vx * EM(i1, i2, 6) - vx * bz_const);
          vel4[v4idx(g, i4, i1, i2, i3)] = a;
          if (i1 >= ng && i1 < ng + g->n[0] && i2 >= ng && i2 < ng + g->n[1] && i3 >= ng &&
              i3 < ng + g->n[2] && i4 >= ng && i4 <= ng + g->n[3])
            aymax = dmax(aymax, fabs(a));
        }
#undef EM
  *axmax_out = axmax;
  *aymax_out = aymax;
}

/* setAccelerationBCs4D (KineticSpeciesF.f:1036-1162): outflow -> quadratic extrapolation marching
 * outward; inflow -> initial condition.  Loops over the full data box in the other dimensions. */
void ok_set_acceleration_bcs_4d(double* u, const ok_geom* g, const double* vel3, const double* vel4,
                                int at_lo3, int at_hi3, int at_lo4, int at_hi4, ok_ic_fn ic,
                                void* ic_ctx) {
  const int ng = g->ng;
  const int n1d = (int)ND(0), n2d = (int)ND(1), n3d = (int)ND(2), n4d = (int)ND(3);
  const int n3a = ng, n3b = ng + g->n[2] - 1, n4a = ng, n4b = ng + g->n[3] - 1;
  if (at_hi3 || at_lo3) {
    for (int i4 = 0; i4 < n4d; ++i4) {
      if (at_hi3)
        for (int i2 = 0; i2 < n2d; ++i2)
          for (int i1 = 0; i1 < n1d; ++i1) {
            if (vel3[v3idx(g, n3b + 1, i4, i1, i2)] >= 0.0) {
              for (int ig = 1; ig <= ng; ++ig)
                F4(u, i1, i2, n3b + ig, i4) = 3.0 * F4(u, i1, i2, n3b + ig - 1, i4) -
                                              3.0 * F4(u, i1, i2, n3b + ig - 2, i4) +
                                              F4(u, i1, i2, n3b + ig - 3, i4);
            } else {
              for (int ig = 1; ig <= ng; ++ig) F4(u, i1, i2, n3b + ig, i4) = ic(ic_ctx, i1, i2, n3b + ig, i4);
            }
          }
      if (at_lo3)
        for (int i2 = 0; i2 < n2d; ++i2)
          for (int i1 = 0; i1 < n1d; ++i1) {
            if (vel3[v3idx(g, n3a, i4, i1, i2)] > 0.0) {
              for (int ig = 1; ig <= ng; ++ig) F4(u, i1, i2, n3a - ig, i4) = ic(ic_ctx, i1, i2, n3a - ig, i4);
            } else {
              for (int ig = 1; ig <= ng; ++ig)
                F4(u, i1, i2, n3a - ig, i4) = 3.0 * F4(u, i1, i2, n3a - ig + 1, i4) -
                                              3.0 * F4(u, i1, i2, n3a - ig + 2, i4) +
                                              F4(u, i1, i2, n3a - ig + 3, i4);
            }
          }
    }
  }
  if (at_hi4 || at_lo4) {
    for (int i3 = 0; i3 < n3d; ++i3) {
      if (at_hi4)
        for (int i2 = 0; i2 < n2d; ++i2)
          for (int i1 = 0; i1 < n1d; ++i1) {
            if (vel4[v4idx(g, n4b + 1, i1, i2, i3)] >= 0.0) {
              for (int ig = 1; ig <= ng; ++ig)
                F4(u, i1, i2, i3, n4b + ig) = 3.0 * F4(u, i1, i2, i3, n4b + ig - 1) -
                                              3.0 * F4(u, i1, i2, i3, n4b + ig - 2) +
                                              F4(u, i1, i2, i3, n4b + ig - 3);
            } else {
              for (int ig = 1; ig <= ng; ++ig) F4(u, i1, i2, i3, n4b + ig) = ic(ic_ctx, i1, i2, i3, n4b + ig);
            }
          }
      if (at_lo4)
        for (int i2 = 0; i2 < n2d; ++i2)
          for (int i1 = 0; i1 < n1d; ++i1) {
            if (vel4[v4idx(g, n4a, i1, i2, i3)] > 0.0) {
              for (int ig = 1; ig <= ng; ++ig) F4(u, i1, i2, i3, n4a - ig) = ic(ic_ctx, i1, i2, i3, n4a - ig);
            } else {
              for (int ig = 1; ig <= ng; ++ig)
                F4(u, i1, i2, i3, n4a - ig) = 3.0 * F4(u, i1, i2, i3, n4a - ig + 1) -
                                              3.0 * F4(u, i1, i2, i3, n4a - ig + 2) +
                                              F4(u, i1, i2, i3, n4a - ig + 3);
            }
          }
    }
  }
}

/* face value between cells (i, i+1) along a line with element stride s; p points at cell i */
static inline double fit_right(const double* p, int64_t s, int order, double vel) {
  if (order == 4) return ok_weno43_fit(p[-s], p[0], p[s], p[2 * s], vel);
  return ok_weno65_fit(p[-2 * s], p[-s], p[0], p[s], p[2 * s], p[3 * s], vel);
}

/* computeadvectionderivatives4D (KineticSpeciesF.f:1949-2089): x pass assigns, y pass adds */
void ok_advection_derivatives_4d(double* rhs, const double* f, const ok_geom* g, const double* vel1,
                                 const double* vel2) {
  const int ng = g->ng;
  const int n1a = ng, n1b = ng + g->n[0] - 1, n2a = ng, n2b = ng + g->n[1] - 1;
  const int64_t s1 = 1, s2 = ND(0);
  const double dx = g->dx[0], dy = g->dx[1];
  OK_PARALLEL_FOR2
  for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
    for (int i3 = ng; i3 < ng + g->n[2]; ++i3) {
      double vx = vel1[v1idx(g, n1a, n2a, i3, i4)];
      double vy = vel2[v2idx(g, n2a, i3, i4, n1a)];
      for (int i2 = n2a; i2 <= n2b; ++i2) {
        double uLeft = fit_right(&F4(f, n1a - 1, i2, i3, i4), s1, g->order, vx);
        for (int i1 = n1a; i1 <= n1b; ++i1) {
          double uRight = fit_right(&F4(f, i1, i2, i3, i4), s1, g->order, vx);
          F4(rhs, i1, i2, i3, i4) = -(vx * uRight - vx * uLeft) / dx;
          uLeft = uRight;
        }
      }
      for (int i1 = n1a; i1 <= n1b; ++i1) {
        double uLeft = fit_right(&F4(f, i1, n2a - 1, i3, i4), s2, g->order, vy);
        for (int i2 = n2a; i2 <= n2b; ++i2) {
          double uRight = fit_right(&F4(f, i1, i2, i3, i4), s2, g->order, vy);
          F4(rhs, i1, i2, i3, i4) = F4(rhs, i1, i2, i3, i4) - (vy * uRight - vy * uLeft) / dy;
          uLeft = uRight;
        }
      }
    }
}

/* computeaccelerationderivatives4D (KineticSpeciesF.f:2093-2245): accumulates into rhs.  The
 * coefficient of a cell is vel3/vel4 at its LOWER face and is also the upwind selector of the fit
 * for its UPPER face (:2145-2152). */
void ok_acceleration_derivatives_4d(double* rhs, const double* f, const ok_geom* g,
                                    const double* vel3, const double* vel4) {
  const int ng = g->ng;
  const int n3a = ng, n3b = ng + g->n[2] - 1, n4a = ng, n4b = ng + g->n[3] - 1;
  const int64_t s3 = ND(0) * ND(1), s4 = ND(0) * ND(1) * ND(2);
  const double dvx = g->dx[2], dvy = g->dx[3];
  OK_PARALLEL_FOR2
  for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
    for (int i1 = ng; i1 < ng + g->n[0]; ++i1) {
      for (int i4 = n4a; i4 <= n4b; ++i4) {
        double ax = vel3[v3idx(g, n3a, i4, i1, i2)];
        double uLeft = fit_right(&F4(f, i1, i2, n3a - 1, i4), s3, g->order, ax);
        for (int i3 = n3a; i3 <= n3b; ++i3) {
          ax = vel3[v3idx(g, i3, i4, i1, i2)];
          double uRight = fit_right(&F4(f, i1, i2, i3, i4), s3, g->order, ax);
          F4(rhs, i1, i2, i3, i4) = F4(rhs, i1, i2, i3, i4) - (ax * uRight - ax * uLeft) / dvx;
          uLeft = uRight;
        }
      }
      for (int i3 = n3a; i3 <= n3b; ++i3) {
        double ay = vel4[v4idx(g, n4a, i1, i2, i3)];
        double uLeft = fit_right(&F4(f, i1, i2, i3, n4a - 1), s4, g->order, ay);
        for (int i4 = n4a; i4 <= n4b; ++i4) {
          ay = vel4[v4idx(g, i4, i1, i2, i3)];
          double uRight = fit_right(&F4(f, i1, i2, i3, i4), s4, g->order, ay);
          F4(rhs, i1, i2, i3, i4) = F4(rhs, i1, i2, i3, i4) - (ay * uRight - ay * uLeft) / dvy;
          uLeft = uRight;
        }
      }
    }
}

/* ------------------------------------------------------------------------------------------------------
 * Flux-form diagnostics: WENO43Avg4D / WENO65Avg4D (KineticSpeciesF.f:630-720, 797-910) + computeFlux4D
 * (:2359-2396) as computeadvectionfluxes4D (:1838-1945) and computeaccelerationfluxes4D (:2249-2355) call them.
 * Direction d (0..3): face, vel and flux are the ROTATED arrays of KineticSpecies.C:1569-1584 -- extents
 * (nd[d]+1, nd[d+1], nd[d+2], nd[d+3]) with the indices taken mod 4, face index j0 = the face below cell j0 of
 * direction d.  The fit runs over faces j0 = w .. nd[d]-w (w = 2 at order 4, 3 at order 6: `f1a = nf1a+2`, `+3`) and
 * the whole data box in the other directions; the flux = vel * face runs over the faces 2 .. extent-3 of ALL four
 * rotated extents whatever the order (computeFlux4D's `+2 / -2`), so at order 6 it also multiplies two faces the fit
 * never wrote.  Entries outside those ranges are left as the caller gave them.
 * ---------------------------------------------------------------------------------------------------- */
void ok_face_fluxes_4d(double* flux, double* face, const double* u, const ok_geom* g, const double* vel, int d) {
  const int w = (g->order == 4) ? 2 : 3;
  int64_t e[4], cs[4];     /* rotated extents; cell strides of the rotated index positions */
  const int64_t s_cell[4] = {1, ND(0), ND(0) * ND(1), ND(0) * ND(1) * ND(2)};
  for (int k = 0; k < 4; ++k) {
    e[k] = ND((d + k) % 4) + (k == 0 ? 1 : 0);
    cs[k] = s_cell[(d + k) % 4];
  }
  OK_PARALLEL_FOR2
  for (int64_t j3 = 0; j3 < e[3]; ++j3)
    for (int64_t j2 = 0; j2 < e[2]; ++j2)
      for (int64_t j1 = 0; j1 < e[1]; ++j1)
        for (int64_t j0 = w; j0 <= e[0] - 1 - w; ++j0) {
          const int64_t fi = j0 + e[0] * (j1 + e[1] * (j2 + e[2] * j3));
          const double* c = u + j0 * cs[0] + j1 * cs[1] + j2 * cs[2] + j3 * cs[3];   /* cell j0: the one above the face */
          face[fi] = fit_right(c - cs[0], cs[0], g->order, vel[fi]);
        }
  OK_PARALLEL_FOR2
  for (int64_t j3 = 2; j3 <= e[3] - 3; ++j3)
    for (int64_t j2 = 2; j2 <= e[2] - 3; ++j2)
      for (int64_t j1 = 2; j1 <= e[1] - 3; ++j1)
        for (int64_t j0 = 2; j0 <= e[0] - 3; ++j0) {
          const int64_t fi = j0 + e[0] * (j1 + e[1] * (j2 + e[2] * j3));
          flux[fi] = vel[fi] * face[fi];
        }
}

/* accumfluxdiv4D (KineticSpeciesF.f:985-1032): rhs = -div(flux) on the interior; the vy term is divided by dvx as
 * in the reference (:1024) */
void ok_accum_flux_div_4d(double* rhs, const ok_geom* g, const double* flux1, const double* flux2, const double* flux3,
                          const double* flux4) {
  const int ng = g->ng;
  const double dx = g->dx[0], dy = g->dx[1], dvx = g->dx[2];
  OK_PARALLEL_FOR2
  for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
    for (int i3 = ng; i3 < ng + g->n[2]; ++i3)
      for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
        for (int i1 = ng; i1 < ng + g->n[0]; ++i1) {
          double temp = -(flux1[v1idx(g, i1 + 1, i2, i3, i4)] - flux1[v1idx(g, i1, i2, i3, i4)]) / dx -
                        (flux2[v2idx(g, i2 + 1, i3, i4, i1)] - flux2[v2idx(g, i2, i3, i4, i1)]) / dy -
                        (flux3[v3idx(g, i3 + 1, i4, i1, i2)] - flux3[v3idx(g, i3, i4, i1, i2)]) / dvx -
                        (flux4[v4idx(g, i4 + 1, i1, i2, i3)] - flux4[v4idx(g, i4, i1, i2, i3)]) / dvx;
          F4(rhs, i1, i2, i3, i4) = temp;
        }
}

/* computekeflux (KineticSpeciesF.f:2734-2893): kinetic-energy flux through the phase-space boundary (dir, side) of a
 * box that touches it (the caller decides: `n1a .eq. ng1a` etc. compare with the domain box); sequential sums in
 * the reference's loop order */
double ok_compute_ke_flux(const ok_geom* g, const double* flux1, const double* flux2, const double* flux3,
                          const double* flux4, const double* velocities, const double* vxface_velocities,
                          const double* vyface_velocities, int dir, int side, double mass) {
  const int ng = g->ng;
  const int a[4] = {ng, ng, ng, ng}, b[4] = {ng + g->n[0] - 1, ng + g->n[1] - 1, ng + g->n[2] - 1, ng + g->n[3] - 1};
  const int fidx = side == 0 ? a[dir] : b[dir] + 1;
  const int64_t n3d = ND(2), n4d = ND(3);
  double ke = 0.0, ddir;
  if (dir == 0) {
    ddir = g->dx[1] * g->dx[2] * g->dx[3];
    for (int i4 = a[3]; i4 <= b[3]; ++i4)
      for (int i3 = a[2]; i3 <= b[2]; ++i3) {
        const double vx = velocities[i3 + n3d * i4], vy = velocities[i3 + n3d * (i4 + n4d)];
        const double v2 = vx * vx + vy * vy;
        for (int i2 = a[1]; i2 <= b[1]; ++i2) ke = ke + 0.5 * flux1[v1idx(g, fidx, i2, i3, i4)] * v2;
      }
  } else if (dir == 1) {
    ddir = g->dx[0] * g->dx[2] * g->dx[3];
    for (int i4 = a[3]; i4 <= b[3]; ++i4)
      for (int i3 = a[2]; i3 <= b[2]; ++i3) {
        const double vx = velocities[i3 + n3d * i4], vy = velocities[i3 + n3d * (i4 + n4d)];
        const double v2 = vx * vx + vy * vy;
        for (int i1 = a[0]; i1 <= b[0]; ++i1) ke = ke + 0.5 * flux2[v2idx(g, fidx, i3, i4, i1)] * v2;
      }
  } else if (dir == 2) {
    ddir = g->dx[0] * g->dx[1] * g->dx[3];
    for (int i4 = a[3]; i4 <= b[3]; ++i4) {
      const double vx = vxface_velocities[fidx + (n3d + 1) * i4], vy = vxface_velocities[fidx + (n3d + 1) * (i4 + n4d)];
      const double v2 = vx * vx + vy * vy;
      for (int i2 = a[1]; i2 <= b[1]; ++i2)
        for (int i1 = a[0]; i1 <= b[0]; ++i1) ke = ke + 0.5 * flux3[v3idx(g, fidx, i4, i1, i2)] * v2;
    }
  } else {
    ddir = g->dx[0] * g->dx[1] * g->dx[2];
    for (int i3 = a[2]; i3 <= b[2]; ++i3) {
      const double vx = vyface_velocities[i3 + n3d * fidx], vy = vyface_velocities[i3 + n3d * (fidx + (n4d + 1))];
      const double v2 = vx * vx + vy * vy;
      for (int i2 = a[1]; i2 <= b[1]; ++i2)
        for (int i1 = a[0]; i1 <= b[0]; ++i1) ke = ke + 0.5 * flux4[v4idx(g, fidx, i1, i2, i3)] * v2;
    }
  }
  return ke * mass * ddir;
}

/* computekevelspaceflux (KineticSpeciesF.f:2897-2990): the same through a velocity boundary (dir 2 / 3), left as a
 * field over (x,y): ke_flux(n1d,n2d) accumulates */
void ok_compute_ke_vel_space_flux(double* ke_flux, const ok_geom* g, const double* flux3, const double* flux4,
                                  const double* vxface_velocities, const double* vyface_velocities, int dir, int side,
                                  double mass) {
  const int ng = g->ng;
  const int64_t n3d = ND(2), n4d = ND(3);
  if (dir == 2) {
    const int i3 = side == 0 ? ng : ng + g->n[2];
    const double ddir = g->dx[3];
    for (int i4 = ng; i4 < ng + g->n[3]; ++i4) {
      const double vx = vxface_velocities[i3 + (n3d + 1) * i4], vy = vxface_velocities[i3 + (n3d + 1) * (i4 + n4d)];
      const double v2 = vx * vx + vy * vy;
      for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
        for (int i1 = ng; i1 < ng + g->n[0]; ++i1)
          ke_flux[i1 + ND(0) * i2] = ke_flux[i1 + ND(0) * i2] + 0.5 * mass * flux3[v3idx(g, i3, i4, i1, i2)] * v2 * ddir;
    }
  } else if (dir == 3) {
    const int i4 = side == 0 ? ng : ng + g->n[3];
    const double ddir = g->dx[2];
    for (int i3 = ng; i3 < ng + g->n[2]; ++i3) {
      const double vx = vyface_velocities[i3 + n3d * i4], vy = vyface_velocities[i3 + n3d * (i4 + (n4d + 1))];
      const double v2 = vx * vx + vy * vy;
      for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
        for (int i1 = ng; i1 < ng + g->n[0]; ++i1)
          ke_flux[i1 + ND(0) * i2] = ke_flux[i1 + ND(0) * i2] + 0.5 * mass * flux4[v4idx(g, i4, i1, i2, i3)] * v2 * ddir;
    }
  }
}

/* computecurrents (KineticSpeciesF.f:2400-2443) */
void ok_compute_currents(const ok_geom* g, const double* velocities, const double* u, const double* vz,
                         double* Jx, double* Jy, double* Jz) {
  const int ng = g->ng;
  const int64_t n3d = ND(2), n4d = ND(3), n1d = ND(0);
  for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
    for (int i3 = ng; i3 < ng + g->n[2]; ++i3) {
      double vx = velocities[i3 + n3d * (i4 + n4d * 0)];
      double vy = velocities[i3 + n3d * (i4 + n4d * 1)];
      for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
        for (int i1 = ng; i1 < ng + g->n[0]; ++i1) {
          double uu = F4(u, i1, i2, i3, i4);
          F4(Jx, i1, i2, i3, i4) = uu * vx;
          F4(Jy, i1, i2, i3, i4) = uu * vy;
          F4(Jz, i1, i2, i3, i4) = uu * vz[i1 + n1d * i2];
        }
    }
}

/* computekeedot (KineticSpeciesF.f:2563-2602); the running sum starts from the incoming value */
double ok_compute_ke_e_dot(const ok_geom* g, const double* u, double charge, const double* velocities,
                           const double* ext_efield, double ke_e_dot) {
  const int ng = g->ng;
  const int64_t n3d = ND(2), n1d = ND(0);
  for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
    for (int i3 = ng; i3 < ng + g->n[2]; ++i3) {
      double vx = velocities[i3 + n3d * i4];
      for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
        for (int i1 = ng; i1 < ng + g->n[0]; ++i1)
          ke_e_dot = ke_e_dot + ext_efield[i1 + n1d * i2] * vx * F4(u, i1, i2, i3, i4);
    }
  ke_e_dot = ke_e_dot * charge * g->dx[0] * g->dx[1] * g->dx[2] * g->dx[3];
  return ke_e_dot;
}

/* ReductionSchedule::sum_reduce_4d_to_2d + execute scaling (ReductionSchedule.C:421-444, 86-89):
 * dst is a 2D array with ghosts, zeroed, sequential sum (i1 fastest ... i4 slowest), then *=dv, *=weight */
void ok_reduce_4d_to_2d(double* dst, const double* src, const ok_geom* g, double dv, double weight) {
  const int ng = g->ng;
  const int64_t n1d = ND(0), n2d = ND(1);
  for (int64_t k = 0; k < n1d * n2d; ++k) dst[k] = 0.0;
  /* every dst(i1,i2) is its own sequential sum, i3 inner / i4 outer: the threads split i2, the order stays */
  OK_PARALLEL_FOR
  for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
    for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
      for (int i3 = ng; i3 < ng + g->n[2]; ++i3)
        for (int i1 = ng; i1 < ng + g->n[0]; ++i1) dst[i1 + n1d * i2] += F4(src, i1, i2, i3, i4);
  /* ParallelArray::operator*= runs over the whole data box (ParallelArray.H:708-716) */
  for (int64_t k = 0; k < n1d * n2d; ++k) dst[k] *= dv;
  for (int64_t k = 0; k < n1d * n2d; ++k) dst[k] *= weight;
}

/* communicatePeriodicBoundaries on one rank (ParallelArray.H:580-606): x sweep then y sweep, each over
 * the full extent of the other dimensions */
void ok_periodic_fill_4d(double* u, const ok_geom* g, int periodic_x, int periodic_y) {
  const int ng = g->ng;
  const int n1d = (int)ND(0), n2d = (int)ND(1), n3d = (int)ND(2), n4d = (int)ND(3);
  if (periodic_x)
    for (int i4 = 0; i4 < n4d; ++i4)
      for (int i3 = 0; i3 < n3d; ++i3)
        for (int i2 = 0; i2 < n2d; ++i2)
          for (int k = 0; k < ng; ++k) {
            F4(u, k, i2, i3, i4) = F4(u, k + g->n[0], i2, i3, i4);
            F4(u, ng + g->n[0] + k, i2, i3, i4) = F4(u, ng + k, i2, i3, i4);
          }
  if (periodic_y)
    for (int i4 = 0; i4 < n4d; ++i4)
      for (int i3 = 0; i3 < n3d; ++i3)
        for (int k = 0; k < ng; ++k)
          for (int i1 = 0; i1 < n1d; ++i1) {
            F4(u, i1, k, i3, i4) = F4(u, i1, k + g->n[1], i3, i4);
            F4(u, i1, ng + g->n[1] + k, i3, i4) = F4(u, i1, ng + k, i3, i4);
          }
}

void ok_periodic_fill_2d(double* u, int n1, int n2, int ng, int ncomp, int periodic_x, int periodic_y) {
  const int64_t n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  for (int c = 0; c < ncomp; ++c) {
    double* a = u + (int64_t)c * n1d * n2d;
    if (periodic_x)
      for (int i2 = 0; i2 < n2d; ++i2)
        for (int k = 0; k < ng; ++k) {
          a[k + n1d * i2] = a[k + n1 + n1d * i2];
          a[ng + n1 + k + n1d * i2] = a[ng + k + n1d * i2];
        }
    if (periodic_y)
      for (int k = 0; k < ng; ++k)
        for (int i1 = 0; i1 < n1d; ++i1) {
          a[i1 + n1d * k] = a[i1 + n1d * (k + n2)];
          a[i1 + n1d * (ng + n2 + k)] = a[i1 + n1d * (ng + k)];
        }
  }
}

/* buildVelocityArrays, non-relativistic branch (KineticSpecies.C:2024-2047).  lo34 = global index of
 * data-box element 0 in V1,V2 (i.e. interior lower - ng). */
void ok_build_velocity_tables(const ok_geom* g, const int lo34[2], double vxlo, double vylo,
                              double* velocities, double* vxface_vel, double* vyface_vel) {
  const int64_t n3d = ND(2), n4d = ND(3);
  const double dvx = g->dx[2], dvy = g->dx[3];
  for (int i3 = 0; i3 < n3d; ++i3) {
    double vx = vxlo + ((i3 + lo34[0]) + 0.5) * dvx;
    for (int i4 = 0; i4 < n4d; ++i4) {
      velocities[i3 + n3d * (i4 + n4d * 0)] = vx;
      velocities[i3 + n3d * (i4 + n4d * 1)] = vylo + ((i4 + lo34[1]) + 0.5) * dvy;
    }
  }
  for (int i3 = 0; i3 <= n3d; ++i3) {
    double vx = vxlo + (i3 + lo34[0]) * dvx;
    for (int i4 = 0; i4 < n4d; ++i4) {
      vxface_vel[i3 + (n3d + 1) * (i4 + n4d * 0)] = vx;
      vxface_vel[i3 + (n3d + 1) * (i4 + n4d * 1)] = vylo + ((i4 + lo34[1]) + 0.5) * dvy;
    }
  }
  for (int i3 = 0; i3 < n3d; ++i3) {
    double vx = vxlo + ((i3 + lo34[0]) + 0.5) * dvx;
    for (int i4 = 0; i4 <= n4d; ++i4) {
      vyface_vel[i3 + n3d * (i4 + (n4d + 1) * 0)] = vx;
      vyface_vel[i3 + n3d * (i4 + (n4d + 1) * 1)] = vylo + (i4 + lo34[1]) * dvy;
    }
  }
}

/* initializeVelocity (KineticSpecies.C:1656-1694) */
void ok_initialize_velocity(const ok_geom* g, const double* velocities, double* vel1, double* vel2) {
  const int n1d = (int)ND(0), n2d = (int)ND(1), n3d = (int)ND(2), n4d = (int)ND(3);
  for (int i3 = 0; i3 < n3d; ++i3)
    for (int i4 = 0; i4 < n4d; ++i4) {
      double c3 = velocities[i3 + (int64_t)n3d * (i4 + (int64_t)n4d * 0)];
      double c4 = velocities[i3 + (int64_t)n3d * (i4 + (int64_t)n4d * 1)];
      for (int i2 = 0; i2 < n2d; ++i2)
        for (int i1 = 0; i1 <= n1d; ++i1) vel1[v1idx(g, i1, i2, i3, i4)] = c3;
      for (int i1 = 0; i1 < n1d; ++i1)
        for (int i2 = 0; i2 <= n2d; ++i2) vel2[v2idx(g, i2, i3, i4, i1)] = c4;
    }
}

/* ------------------------------------------------------------------------------------------
 * Poisson (PoissonF.f:10-123, LokiPoissonSolveFFT.C:31-170)
 * ------------------------------------------------------------------------------------------ */
void ok_neutralize_charge(double* rho, int n1, int n2, int ng) {
  const int64_t n1d = n1 + 2 * ng;
  int count = 0;
  double sum = 0.0;
  for (int i2 = ng; i2 < ng + n2; ++i2)
    for (int i1 = ng; i1 < ng + n1; ++i1) {
      sum = sum + rho[i1 + n1d * i2];
      count = count + 1;
    }
  sum = sum / count;
  for (int i2 = ng; i2 < ng + n2; ++i2)
    for (int i1 = ng; i1 < ng + n1; ++i1) rho[i1 + n1d * i2] = rho[i1 + n1d * i2] - sum;
}

/* symbols of the 4th/6th-order FD Laplacian, pre-multiplied by nx*ny (LokiPoissonSolveFFT.C:66-115).
 * sx has nx entries, sy has ny/2+1.  order -1 = spectral. */
void ok_poisson_symbols(int nx, int ny, double Lx, double Ly, int order, double* sx, double* sy) {
  const double pi = 4.0 * atan(1.0);
  double dx = Lx / nx, dy = Ly / ny;
  for (int i = 0; i < nx; ++i) {
    double kx = 0.0;
    if (i >= 1) kx = (2 * i < nx) ? (2.0 * pi / Lx) * i : (2.0 * pi / Lx) * (nx - i);
    double dpdm = (2.0 * cos(dx * kx) - 2.0) / pow(dx, 2.0);
    double s;
    if (order == 4)
      s = dpdm - pow(dx, 2.0) / 12.0 * pow(dpdm, 2.0);
    else if (order == 6)
      s = dpdm - pow(dx, 2.0) / 12.0 * pow(dpdm, 2.0) + pow(dx, 4.0) / 90.0 * pow(dpdm, 3.0);
    else
      s = -kx * kx;
    sx[i] = s * (nx * ny);
  }
  for (int i = 0; i < ny / 2 + 1; ++i) {
    double ky = (i >= 1) ? (2.0 * pi / Ly) * i : 0.0;
    double dpdm = (2.0 * cos(dy * ky) - 2.0) / pow(dy, 2.0);
    double s;
    if (order == 4)
      s = dpdm - pow(dy, 2.0) / 12.0 * pow(dpdm, 2.0);
    else if (order == 6)
      s = dpdm - pow(dy, 2.0) / 12.0 * pow(dpdm, 2.0) + pow(dy, 4.0) / 90.0 * pow(dpdm, 3.0);
    else
      s = -ky * ky;
    sy[i] = s * (nx * ny);
  }
}

/* The reference calls FFTW3 r2c/c2r (third-party, absent here; version unpinned, configure.in:350-364).
 * The published definition is restated as a plain O(N^2)-per-line real DFT: forward
 * X[i][j] = sum_{a,b} x[a][b] exp(-2 pi I (i a/nx + j b/ny)), unnormalised inverse; the symbols carry
 * the nx*ny normalisation.  The result agrees with FFTW to round-off, not bitwise ("parity
 * unpinned" at this third-party boundary; the GPU path is compared with a tolerance). */
void ok_poisson_fft_solve(double* phi, const double* rho, int nx, int ny, int ng, const double* sx,
                          const double* sy) {
  const double pi = 4.0 * atan(1.0);
  const int64_t n1d = nx + 2 * ng;
  const int nyh = ny / 2 + 1;
  double* cx = (double*)malloc(sizeof(double) * 2 * nx);
  double* cy = (double*)malloc(sizeof(double) * 2 * ny);
  for (int k = 0; k < nx; ++k) { cx[2 * k] = cos(2.0 * pi * k / nx); cx[2 * k + 1] = sin(2.0 * pi * k / nx); }
  for (int k = 0; k < ny; ++k) { cy[2 * k] = cos(2.0 * pi * k / ny); cy[2 * k + 1] = sin(2.0 * pi * k / ny); }
  /* stage 1: along y (real -> half complex): T[a][j] */
  double* T = (double*)calloc((size_t)2 * nx * nyh, sizeof(double));
  for (int a = 0; a < nx; ++a)
    for (int j = 0; j < nyh; ++j) {
      double re = 0.0, im = 0.0;
      for (int b = 0; b < ny; ++b) {
        int m = (int)(((int64_t)j * b) % ny);
        double v = rho[(a + ng) + n1d * (b + ng)];
        re += v * cy[2 * m];
        im -= v * cy[2 * m + 1];
      }
      T[2 * (a * nyh + j)] = re;
      T[2 * (a * nyh + j) + 1] = im;
    }
  /* stage 2: along x, divide by symbol */
  double* X = (double*)calloc((size_t)2 * nx * nyh, sizeof(double));
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < nyh; ++j) {
      double re = 0.0, im = 0.0;
      for (int a = 0; a < nx; ++a) {
        int m = (int)(((int64_t)i * a) % nx);
        double tr = T[2 * (a * nyh + j)], ti = T[2 * (a * nyh + j) + 1];
        /* (tr + I ti) * (c - I s) */
        re += tr * cx[2 * m] + ti * cx[2 * m + 1];
        im += ti * cx[2 * m] - tr * cx[2 * m + 1];
      }
      if (sx[i] != 0.0 || sy[j] != 0.0) {
        re /= sx[i] + sy[j];
        im /= sx[i] + sy[j];
      }
      X[2 * (i * nyh + j)] = re;
      X[2 * (i * nyh + j) + 1] = im;
    }
  /* inverse along x: U[a][j] = sum_i X[i][j] exp(+2 pi I i a / nx) */
  for (int a = 0; a < nx; ++a)
    for (int j = 0; j < nyh; ++j) {
      double re = 0.0, im = 0.0;
      for (int i = 0; i < nx; ++i) {
        int m = (int)(((int64_t)i * a) % nx);
        double xr = X[2 * (i * nyh + j)], xi = X[2 * (i * nyh + j) + 1];
        re += xr * cx[2 * m] - xi * cx[2 * m + 1];
        im += xi * cx[2 * m] + xr * cx[2 * m + 1];
      }
      T[2 * (a * nyh + j)] = re;
      T[2 * (a * nyh + j) + 1] = im;
    }
  /* inverse along y (half complex -> real): x[a][b] = sum_j' U[a][j'] e^{+..}, Hermitian completion */
  for (int a = 0; a < nx; ++a)
    for (int b = 0; b < ny; ++b) {
      double acc = 0.0;
      for (int j = 0; j < nyh; ++j) {
        int m = (int)(((int64_t)j * b) % ny);
        double ur = T[2 * (a * nyh + j)], ui = T[2 * (a * nyh + j) + 1];
        double term = ur * cy[2 * m] - ui * cy[2 * m + 1];
        int self_conj = (j == 0) || (2 * j == ny);
        acc += self_conj ? term : 2.0 * term;
      }
      phi[(a + ng) + n1d * (b + ng)] = acc;
    }
  free(cx); free(cy); free(T); free(X);
}

/* computeEFieldFromPotential (PoissonF.f:68-123): E = +grad(phi), centred 4th/6th order */
void ok_efield_from_potential(double* em, const double* phi, int n1, int n2, int ng, int order,
                              int em_vars_dim, const double* dx) {
  (void)em_vars_dim;
  const int64_t n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  const int64_t pl = n1d * n2d;
#define PH(i1, i2) phi[(i1) + n1d * (i2)]
  for (int i2 = ng; i2 < ng + n2; ++i2)
    for (int i1 = ng; i1 < ng + n1; ++i1) {
      if (order == 4) {
        em[i1 + n1d * i2] =
            (PH(i1 - 2, i2) - 8.0 * PH(i1 - 1, i2) + 8.0 * PH(i1 + 1, i2) - PH(i1 + 2, i2)) / (12.0 * dx[0]);
        em[i1 + n1d * i2 + pl] =
            (PH(i1, i2 - 2) - 8.0 * PH(i1, i2 - 1) + 8.0 * PH(i1, i2 + 1) - PH(i1, i2 + 2)) / (12.0 * dx[1]);
      } else {
        em[i1 + n1d * i2] = (-1.0 * PH(i1 - 3, i2) + 9.0 * PH(i1 - 2, i2) - 45.0 * PH(i1 - 1, i2) +
                             45.0 * PH(i1 + 1, i2) - 9.0 * PH(i1 + 2, i2) + 1.0 * PH(i1 + 3, i2)) /
                            (60.0 * dx[0]);
        em[i1 + n1d * i2 + pl] = (-1.0 * PH(i1, i2 - 3) + 9.0 * PH(i1, i2 - 2) - 45.0 * PH(i1, i2 - 1) +
                                  45.0 * PH(i1, i2 + 1) - 9.0 * PH(i1, i2 + 2) + 1.0 * PH(i1, i2 + 3)) /
                                 (60.0 * dx[1]);
      }
    }
#undef PH
}

/* ------------------------------------------------------------------------------------------
 * Maxwell (MaxwellF.f:62-93 xpby2d, 97-355 maxwellevalrhs, 442-469 maxwellevalvzrhs).  Supergrid
 * absorbing layers are not restated: without them SGMetricFunction returns nu = 1.0 exactly
 * (MaxwellF.f:413-421 with xi = 0), and a multiplication by 1.0 is exact.
 * Integer powers follow the compiler's repeated-squaring expansion (x**4 = (x*x)*(x*x) ...).
 * ------------------------------------------------------------------------------------------ */
void ok_xpby2d(double* x, const double* y, double b, int n1, int n2, int ng, int ncomp) {
  const int64_t n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  for (int c = 0; c < ncomp; ++c)
    for (int i2 = ng; i2 < ng + n2; ++i2)
      for (int i1 = ng; i1 < ng + n1; ++i1) {
        int64_t o = i1 + n1d * (i2 + n2d * c);
        x[o] = x[o] + b * y[o];
      }
}

static inline double p2(double x) { return x * x; }
static inline double p3(double x) { return (x * x) * x; }
static inline double p4(double x) { double t = x * x; return t * t; }
static inline double p5(double x) { double t = x * x; return (t * x) * t; }
static inline double p6(double x) { double t = (x * x) * x; return t * t; }

void ok_maxwell_eval_rhs(double* rhs, const double* em, const double* Jx, const double* Jy,
                         const double* Jz, int n1, int n2, int ng, int order, const double* dx,
                         double c, double avWeak, double avStrong) {
  const int64_t n1d = n1 + 2 * ng, n2d = n2 + 2 * ng, pl = n1d * n2d;
  const double csquared = p2(c);
#define EMV(i1, i2, k) em[(i1) + n1d * (i2) + pl * ((k)-1)]
#define DEM(i1, i2, k) rhs[(i1) + n1d * (i2) + pl * ((k)-1)]
#define D4X(k) ((EMV(i1 - 2, i2, k) - 8.0 * EMV(i1 - 1, i2, k) + 8.0 * EMV(i1 + 1, i2, k) - EMV(i1 + 2, i2, k)) / (12.0 * dx[0]))
#define D4Y(k) ((EMV(i1, i2 - 2, k) - 8.0 * EMV(i1, i2 - 1, k) + 8.0 * EMV(i1, i2 + 1, k) - EMV(i1, i2 + 2, k)) / (12.0 * dx[1]))
#define D6X(k) ((-1.0 * EMV(i1 - 3, i2, k) + 9.0 * EMV(i1 - 2, i2, k) - 45.0 * EMV(i1 - 1, i2, k) + 45.0 * EMV(i1 + 1, i2, k) - 9.0 * EMV(i1 + 2, i2, k) + 1.0 * EMV(i1 + 3, i2, k)) / (60.0 * dx[0]))
#define D6Y(k) ((-1.0 * EMV(i1, i2 - 3, k) + 9.0 * EMV(i1, i2 - 2, k) - 45.0 * EMV(i1, i2 - 1, k) + 45.0 * EMV(i1, i2 + 1, k) - 9.0 * EMV(i1, i2 + 2, k) + 1.0 * EMV(i1, i2 + 3, k)) / (60.0 * dx[1]))
  for (int i2 = ng; i2 < ng + n2; ++i2)
    for (int i1 = ng; i1 < ng + n1; ++i1) {
      double Exdy, Eydx, Ezdx, Ezdy, Bxdy, Bydx, Bzdx, Bzdy;
      const double nux = 1.0, nuy = 1.0;
      if (order == 4) {
        Exdy = nuy * D4Y(1); Eydx = nux * D4X(2); Ezdx = nux * D4X(3); Ezdy = nuy * D4Y(3);
        Bxdy = nuy * D4Y(4); Bydx = nux * D4X(5); Bzdx = nux * D4X(6); Bzdy = nuy * D4Y(6);
      } else {
        Exdy = nuy * D6Y(1); Eydx = nux * D6X(2); Ezdx = nux * D6X(3); Ezdy = nuy * D6Y(3);
        Bxdy = nuy * D6Y(4); Bydx = nux * D6X(5); Bzdx = nux * D6X(6); Bzdy = nuy * D6Y(6);
      }
      const int64_t o = i1 + n1d * i2;
      DEM(i1, i2, 1) = csquared * (Bzdy)-Jx[o];
      DEM(i1, i2, 2) = -csquared * (Bzdx)-Jy[o];
      DEM(i1, i2, 3) = csquared * (Bydx - Bxdy) - Jz[o];
      DEM(i1, i2, 4) = -Ezdy;
      DEM(i1, i2, 5) = Ezdx;
      DEM(i1, i2, 6) = Exdy - Eydx;
    }
  if (avWeak > 0.0 || avStrong > 0.0) {
    for (int k = 1; k <= 6; ++k)
      for (int i2 = ng; i2 < ng + n2; ++i2)
        for (int i1 = ng; i1 < ng + n1; ++i1) {
          if (order == 4) {
            double uxxxx = (1.0 * EMV(i1 - 2, i2, k) - 4.0 * EMV(i1 - 1, i2, k) + 6.0 * EMV(i1, i2, k) -
                            4.0 * EMV(i1 + 1, i2, k) + 1.0 * EMV(i1 + 2, i2, k)) / (p4(dx[0]));
            double uyyyy = (1.0 * EMV(i1, i2 - 2, k) - 4.0 * EMV(i1, i2 - 1, k) + 6.0 * EMV(i1, i2, k) -
                            4.0 * EMV(i1, i2 + 1, k) + 1.0 * EMV(i1, i2 + 2, k)) / (p4(dx[1]));
            DEM(i1, i2, k) = DEM(i1, i2, k) -
                             (avWeak * c * p4(dx[0]) + avStrong * c * p3(dx[0])) / 16.0 * uxxxx -
                             (avWeak * c * p4(dx[1]) + avStrong * c * p3(dx[1])) / 16.0 * uyyyy;
          } else {
            double u6x = (1.0 * EMV(i1 - 3, i2, k) - 6.0 * EMV(i1 - 2, i2, k) + 15.0 * EMV(i1 - 1, i2, k) -
                          20.0 * EMV(i1, i2, k) + 15.0 * EMV(i1 + 1, i2, k) - 6.0 * EMV(i1 + 2, i2, k) +
                          1.0 * EMV(i1 + 3, i2, k)) / (p6(dx[0]));
            double u6y = (1.0 * EMV(i1, i2 - 3, k) - 6.0 * EMV(i1, i2 - 2, k) + 15.0 * EMV(i1, i2 - 1, k) -
                          20.0 * EMV(i1, i2, k) + 15.0 * EMV(i1, i2 + 1, k) - 6.0 * EMV(i1, i2 + 2, k) +
                          1.0 * EMV(i1, i2 + 3, k)) / (p6(dx[1]));
            DEM(i1, i2, k) = DEM(i1, i2, k) +
                             (avWeak * c * p6(dx[0]) + avStrong * c * p5(dx[0])) / 64.0 * u6x +
                             (avWeak * c * p6(dx[1]) + avStrong * c * p5(dx[1])) / 64.0 * u6y;
          }
        }
  }
#undef EMV
#undef DEM
#undef D4X
#undef D4Y
#undef D6X
#undef D6Y
}

void ok_maxwell_eval_vz_rhs(double* rhs, const double* em, double charge_per_mass, int n1, int n2, int ng) {
  const int64_t n1d = n1 + 2 * ng, n2d = n2 + 2 * ng, pl = n1d * n2d;
  for (int i2 = ng; i2 < ng + n2; ++i2)
    for (int i1 = ng; i1 < ng + n1; ++i1) rhs[i1 + n1d * i2] = charge_per_mass * em[i1 + n1d * i2 + 2 * pl];
}

/* ------------------------------------------------------------------------------------------
 * setAdvectionBCs4D (KineticSpeciesF.f:1166-1297): physical x / y boundaries of a non-periodic direction.
 * Same rule as the velocity boundaries: outflow (sign of the face velocity vel1 / vel2 at the boundary
 * face) -> quadratic extrapolation marching outward, inflow -> the initial condition.  x first over the
 * full data box of the other three directions, then y over the full x extent (ghosts just set included).
 * vel1: (n1d+1,n2d,n3d,n4d), vel2: (n2d+1,n3d,n4d,n1d) as the reference holds them.
 * ------------------------------------------------------------------------------------------ */
void ok_set_advection_bcs_4d(double* u, const ok_geom* g, const double* vel1, const double* vel2, int at_lo1,
                             int at_hi1, int at_lo2, int at_hi2, int x_periodic, int y_periodic, ok_ic_fn ic,
                             void* ic_ctx) {
  const int ng = g->ng;
  const int64_t n1d = ND(0), n2d = ND(1), n3d = ND(2), n4d = ND(3);
  const int n1a = ng, n1b = ng + g->n[0] - 1, n2a = ng, n2b = ng + g->n[1] - 1;
#define V1(i1, i2, i3, i4) vel1[(i1) + (n1d + 1) * ((i2) + n2d * ((i3) + n3d * (int64_t)(i4)))]
#define V2(i2, i3, i4, i1) vel2[(i2) + (n2d + 1) * ((i3) + n3d * ((i4) + n4d * (int64_t)(i1)))]
  if ((x_periodic != 1) && (at_hi1 || at_lo1)) {
    for (int i4 = 0; i4 < n4d; ++i4)
      for (int i3 = 0; i3 < n3d; ++i3) {
        if (at_hi1)
          for (int i2 = 0; i2 < n2d; ++i2) {
            if (V1(n1b + 1, i2, i3, i4) >= 0.0) {
              for (int ig = 1; ig <= ng; ++ig)
                F4(u, n1b + ig, i2, i3, i4) = 3.0 * F4(u, n1b + ig - 1, i2, i3, i4) -
                                              3.0 * F4(u, n1b + ig - 2, i2, i3, i4) + F4(u, n1b + ig - 3, i2, i3, i4);
            } else {
              for (int ig = 1; ig <= ng; ++ig) F4(u, n1b + ig, i2, i3, i4) = ic(ic_ctx, n1b + ig, i2, i3, i4);
            }
          }
        if (at_lo1)
          for (int i2 = 0; i2 < n2d; ++i2) {
            if (V1(n1a, i2, i3, i4) > 0.0) {
              for (int ig = 1; ig <= ng; ++ig) F4(u, n1a - ig, i2, i3, i4) = ic(ic_ctx, n1a - ig, i2, i3, i4);
            } else {
              for (int ig = 1; ig <= ng; ++ig)
                F4(u, n1a - ig, i2, i3, i4) = 3.0 * F4(u, n1a - ig + 1, i2, i3, i4) -
                                              3.0 * F4(u, n1a - ig + 2, i2, i3, i4) + F4(u, n1a - ig + 3, i2, i3, i4);
            }
          }
      }
  }
  if ((y_periodic != 1) && (at_hi2 || at_lo2)) {
    for (int i4 = 0; i4 < n4d; ++i4)
      for (int i3 = 0; i3 < n3d; ++i3) {
        if (at_hi2)
          for (int i1 = 0; i1 < n1d; ++i1) {
            if (V2(n2b + 1, i3, i4, i1) >= 0.0) {
              for (int ig = 1; ig <= ng; ++ig)
                F4(u, i1, n2b + ig, i3, i4) = 3.0 * F4(u, i1, n2b + ig - 1, i3, i4) -
                                              3.0 * F4(u, i1, n2b + ig - 2, i3, i4) + F4(u, i1, n2b + ig - 3, i3, i4);
            } else {
              for (int ig = 1; ig <= ng; ++ig) F4(u, i1, n2b + ig, i3, i4) = ic(ic_ctx, i1, n2b + ig, i3, i4);
            }
          }
        if (at_lo2)
          for (int i1 = 0; i1 < n1d; ++i1) {
            if (V2(n2a, i3, i4, i1) > 0.0) {
              for (int ig = 1; ig <= ng; ++ig) F4(u, i1, n2a - ig, i3, i4) = ic(ic_ctx, i1, n2a - ig, i3, i4);
            } else {
              for (int ig = 1; ig <= ng; ++ig)
                F4(u, i1, n2a - ig, i3, i4) = 3.0 * F4(u, i1, n2a - ig + 1, i3, i4) -
                                              3.0 * F4(u, i1, n2a - ig + 2, i3, i4) + F4(u, i1, n2a - ig + 3, i3, i4);
            }
          }
      }
  }
#undef V1
#undef V2
}

/* ------------------------------------------------------------------------------------------
 * The "JB" boundary conditions (use_new_bcs = true): setAccelerationBCs4DJB (KineticSpeciesF.f:1301-1520) and
 * setAdvectionBCs4DJB (:1524-1733).  Inflow (lower side: face velocity > 0, upper side: < 0) samples the
 * initial condition; otherwise ghost ig is extrapolated with the binomial formula of order
 * extrapEq = min(interior extent, solution_order), accumulated from 0.0 in stencil order.  Sides in the
 * reference's order: lower then upper of the first direction, then of the second.  The reference detects a
 * global boundary from the cell coordinate (:1368-1371); here the caller passes at_* like for the default BCs.
 * ------------------------------------------------------------------------------------------ */
static const double JB_ECOEFFS[6][6] = {{1.0, 0, 0, 0, 0, 0},        {2.0, -1.0, 0, 0, 0, 0},
                                        {3.0, -3.0, 1.0, 0, 0, 0},   {4.0, -6.0, 4.0, -1.0, 0, 0},
                                        {5.0, -10.0, 10.0, -5.0, 1.0, 0}, {6.0, -15.0, 20.0, -15.0, 6.0, -1.0}};
/* one side of direction d (0..3).  vsel: which face-velocity array, indexed like the reference does */
static void jb_side(double* u, const ok_geom* g, int d, int hi, const double* vel, ok_ic_fn ic, void* ic_ctx) {
  const int ng = g->ng;
  const int64_t nd[4] = {ND(0), ND(1), ND(2), ND(3)};
  const int na = ng, nb = ng + g->n[d] - 1;
  int e = g->n[d] < g->order ? g->n[d] : g->order; /* extrapEq */
  int lo_[4] = {0, 0, 0, 0}, hi_[4] = {(int)nd[0], (int)nd[1], (int)nd[2], (int)nd[3]};
  lo_[d] = 0; hi_[d] = 1; /* the swept direction is fixed at the boundary */
  int i[4];
  for (i[3] = lo_[3]; i[3] < hi_[3]; ++i[3])
    for (i[2] = lo_[2]; i[2] < hi_[2]; ++i[2])
      for (i[1] = lo_[1]; i[1] < hi_[1]; ++i[1])
        for (i[0] = lo_[0]; i[0] < hi_[0]; ++i[0]) {
          int f = hi ? nb + 1 : na; /* face index whose velocity decides */
          double v;
          if (d == 0) v = vel[f + (nd[0] + 1) * (i[1] + nd[1] * (i[2] + nd[2] * (int64_t)i[3]))];         /* vel1(i1,i2,i3,i4) */
          else if (d == 1) v = vel[f + (nd[1] + 1) * (i[2] + nd[2] * (i[3] + nd[3] * (int64_t)i[0]))];    /* vel2(i2,i3,i4,i1) */
          else if (d == 2) v = vel[f + (nd[2] + 1) * (i[3] + nd[3] * (i[0] + nd[0] * (int64_t)i[1]))];    /* vel3(i3,i4,i1,i2) */
          else v = vel[f + (nd[3] + 1) * (i[0] + nd[0] * (i[1] + nd[1] * (int64_t)i[2]))];                /* vel4(i4,i1,i2,i3) */
          int inflow = hi ? (v < 0.0) : (v > 0.0);
          for (int ig = 1; ig <= ng; ++ig) {
            int c[4] = {i[0], i[1], i[2], i[3]};
            c[d] = hi ? nb + ig : na - ig;
            if (inflow) {
              F4(u, c[0], c[1], c[2], c[3]) = ic(ic_ctx, c[0], c[1], c[2], c[3]);
            } else {
              double acc = 0.0;
              for (int k = 1; k <= e; ++k) {
                int q[4] = {c[0], c[1], c[2], c[3]};
                q[d] = hi ? c[d] - k : c[d] + k;
                acc = acc + JB_ECOEFFS[e - 1][k - 1] * F4(u, q[0], q[1], q[2], q[3]);
              }
              F4(u, c[0], c[1], c[2], c[3]) = acc;
            }
          }
        }
}
void ok_set_acceleration_bcs_4d_jb(double* u, const ok_geom* g, const double* vel3, const double* vel4, int at_lo3,
                                   int at_hi3, int at_lo4, int at_hi4, ok_ic_fn ic, void* ic_ctx) {
  if (at_lo3) jb_side(u, g, 2, 0, vel3, ic, ic_ctx);
  if (at_hi3) jb_side(u, g, 2, 1, vel3, ic, ic_ctx);
  if (at_lo4) jb_side(u, g, 3, 0, vel4, ic, ic_ctx);
  if (at_hi4) jb_side(u, g, 3, 1, vel4, ic, ic_ctx);
}
void ok_set_advection_bcs_4d_jb(double* u, const ok_geom* g, const double* vel1, const double* vel2, int at_lo1,
                                int at_hi1, int at_lo2, int at_hi2, int x_periodic, int y_periodic, ok_ic_fn ic,
                                void* ic_ctx) {
  if (at_lo1 && x_periodic != 1) jb_side(u, g, 0, 0, vel1, ic, ic_ctx);
  if (at_hi1 && x_periodic != 1) jb_side(u, g, 0, 1, vel1, ic, ic_ctx);
  if (at_lo2 && y_periodic != 1) jb_side(u, g, 1, 0, vel2, ic, ic_ctx);
  if (at_hi2 && y_periodic != 1) jb_side(u, g, 1, 1, vel2, ic, ic_ctx);
}

/* ------------------------------------------------------------------------------------------
 * appendkrook (KineticSpeciesF.f:2995-3034; called from completeRHS, KineticSpecies.C:1049-1080): Krook-layer
 * damping towards the initial condition, rhs -= nu(x,y)/dt * (u - f0) wherever nu != 0.
 * ------------------------------------------------------------------------------------------ */
void ok_append_krook(double* rhs, const double* u, const ok_geom* g, const double* nu, double dt, ok_ic_fn ic,
                     void* ic_ctx) {
  const int ng = g->ng;
  const int64_t n1d = ND(0);
  for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
    for (int i3 = ng; i3 < ng + g->n[2]; ++i3)
      for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
        for (int i1 = ng; i1 < ng + g->n[0]; ++i1)
          if (nu[i1 + n1d * i2] != 0.0) {
            double f0 = ic(ic_ctx, i1, i2, i3, i4);
            F4(rhs, i1, i2, i3, i4) = F4(rhs, i1, i2, i3, i4) - nu[i1 + n1d * i2] / dt * (F4(u, i1, i2, i3, i4) - f0);
          }
}

/* ------------------------------------------------------------------------------------------
 * Twilight-zone (manufactured-solution) source of TrigTZSource (TZSourceF.f:10-75 settrigtzsource, :79-137
 * computetrigtzsourceerror; called from completeRHS, KineticSpecies.C:1077-1080, and putToRestart, :987-1004).
 * The exact solution is f = alpha/(2 pi) exp(-alpha v^2/2) (1 + A cos(kx x) cos(ky y) sin(kt t)) with kx = ky = kt = alpha = 1;
 * h is what must be added to the Vlasov-Poisson right-hand side so that f solves the forced equation.  Maple's
 * expression, evaluated in the Fortran's parse order.  lo: global index of array cell 0 (the dataBox lower bound).
 * ------------------------------------------------------------------------------------------ */
void ok_set_trig_tz_source(double* f, const ok_geom* g, const int* lo, const double* xlo, const double* dx, double time,
                           const double* velocities, double amp) {
  const int64_t n3d = ND(2), n4d = ND(3);
  const double kx = 1.0, ky = 1.0, kt = 1.0, alpha = 1.0, A = amp, t = time;
  const double pi = 4.0 * atan(1.0);
  for (int i4 = 0; i4 < n4d; ++i4)
    for (int i3 = 0; i3 < n3d; ++i3) {
      const double vx = velocities[i3 + n3d * (i4 + n4d * 0)];
      const double vy = velocities[i3 + n3d * (i4 + n4d * 1)];
      for (int i2 = 0; i2 < ND(1); ++i2) {
        const double y = xlo[1] + ((lo[1] + i2) + 0.5) * dx[1];
        for (int i1 = 0; i1 < ND(0); ++i1) {
          const double x = xlo[0] + ((lo[0] + i1) + 0.5) * dx[0];
          const double h =
              -0.1e1 / (kx * kx + ky * ky) * A * sin(kx * x) * kx * cos(ky * y) * sin(kt * t) * (alpha * alpha) / pi * vx *
                  exp(-(alpha * (vx * vx + vy * vy) / 0.2e1)) * (0.1e1 + A * cos(kx * x) * cos(ky * y) * sin(kt * t)) / 0.2e1 -
              0.1e1 / (kx * kx + ky * ky) * A * cos(kx * x) * sin(ky * y) * ky * sin(kt * t) * (alpha * alpha) / pi * vy *
                  exp(-(alpha * (vx * vx + vy * vy) / 0.2e1)) * (0.1e1 + A * cos(kx * x) * cos(ky * y) * sin(kt * t)) / 0.2e1 -
              alpha / pi * exp(-(alpha * (vx * vx + vy * vy) / 0.2e1)) * A * sin(kx * x) * kx * cos(ky * y) * sin(kt * t) * vx / 0.2e1 -
              alpha / pi * exp(-(alpha * (vx * vx + vy * vy) / 0.2e1)) * A * cos(kx * x) * sin(ky * y) * ky * sin(kt * t) * vy / 0.2e1 +
              alpha / pi * exp(-(alpha * (vx * vx + vy * vy) / 0.2e1)) * A * cos(kx * x) * cos(ky * y) * cos(kt * t) * kt / 0.2e1;
          F4(f, i1, i2, i3, i4) = F4(f, i1, i2, i3, i4) + h;
        }
      }
    }
}
void ok_compute_trig_tz_source_error(double* error, const double* soln, const ok_geom* g, const int* lo, const double* xlo,
                                     const double* dx, double time, const double* velocities, double amp) {
  const int64_t n3d = ND(2), n4d = ND(3);
  const double kx = 1.0, ky = 1.0, kt = 1.0, alpha = 1.0, A = amp, t = time;
  const double pi = 4.0 * atan(1.0);
  for (int i4 = 0; i4 < n4d; ++i4)
    for (int i3 = 0; i3 < n3d; ++i3) {
      const double vx = velocities[i3 + n3d * (i4 + n4d * 0)];
      const double vy = velocities[i3 + n3d * (i4 + n4d * 1)];
      for (int i2 = 0; i2 < ND(1); ++i2) {
        const double y = xlo[1] + ((lo[1] + i2) + 0.5) * dx[1];
        for (int i1 = 0; i1 < ND(0); ++i1) {
          const double x = xlo[0] + ((lo[0] + i1) + 0.5) * dx[0];
          const double fexact = alpha / pi * exp(-(alpha * (vx * vx + vy * vy) / 0.2e1)) *
                                (0.1e1 + A * cos(kx * x) * cos(ky * y) * sin(kt * t)) / 0.2e1;
          F4(error, i1, i2, i3, i4) = F4(soln, i1, i2, i3, i4) - fexact;
        }
      }
    }
}

/* ElectronTrigTZSource (ElectronTZSourceF.f:10-75 setelectrontrigtzsource, :79-143 computeelectrontrigtzsourceerror;
 * deck test/EPWTZ): the same manufactured solution with kx = ky = 4; Maple ordered the source's terms differently */
void ok_set_electron_trig_tz_source(double* f, const ok_geom* g, const int* lo, const double* xlo, const double* dx, double time,
                                    const double* velocities, double amp) {
  const int64_t n3d = ND(2), n4d = ND(3);
  const double kx = 4.0, ky = 4.0, kt = 1.0, alpha = 1.0, A = amp, t = time;
  const double pi = 4.0 * atan(1.0);
  for (int i4 = 0; i4 < n4d; ++i4)
    for (int i3 = 0; i3 < n3d; ++i3) {
      const double vx = velocities[i3 + n3d * (i4 + n4d * 0)];
      const double vy = velocities[i3 + n3d * (i4 + n4d * 1)];
      for (int i2 = 0; i2 < ND(1); ++i2) {
        const double y = xlo[1] + ((lo[1] + i2) + 0.5) * dx[1];
        for (int i1 = 0; i1 < ND(0); ++i1) {
          const double x = xlo[0] + ((lo[0] + i1) + 0.5) * dx[0];
          const double h =
              alpha / pi * exp(-(alpha * (vx * vx + vy * vy) / 0.2e1)) * A * cos(kx * x) * cos(ky * y) * kt * cos(kt * t) / 0.2e1 -
              vx * alpha / pi * exp(-(alpha * (vx * vx + vy * vy) / 0.2e1)) * A * kx * sin(kx * x) * cos(ky * y) * sin(kt * t) / 0.2e1 -
              vy * alpha / pi * exp(-(alpha * (vx * vx + vy * vy) / 0.2e1)) * A * cos(kx * x) * ky * sin(ky * y) * sin(kt * t) / 0.2e1 -
              A * kx * sin(kx * x) * cos(ky * y) * sin(kt * t) / (kx * kx + ky * ky) * (alpha * alpha) / pi * vx *
                  exp(-(alpha * (vx * vx + vy * vy) / 0.2e1)) * (0.1e1 + A * cos(kx * x) * cos(ky * y) * sin(kt * t)) / 0.2e1 -
              A * cos(kx * x) * ky * sin(ky * y) * sin(kt * t) / (kx * kx + ky * ky) * (alpha * alpha) / pi * vy *
                  exp(-(alpha * (vx * vx + vy * vy) / 0.2e1)) * (0.1e1 + A * cos(kx * x) * cos(ky * y) * sin(kt * t)) / 0.2e1;
          F4(f, i1, i2, i3, i4) = F4(f, i1, i2, i3, i4) + h;
        }
      }
    }
}
void ok_compute_electron_trig_tz_source_error(double* error, const double* soln, const ok_geom* g, const int* lo,
                                              const double* xlo, const double* dx, double time, const double* velocities,
                                              double amp) {
  const int64_t n3d = ND(2), n4d = ND(3);
  const double kx = 4.0, ky = 4.0, kt = 1.0, alpha = 1.0, A = amp, t = time;
  const double pi = 4.0 * atan(1.0);
  for (int i4 = 0; i4 < n4d; ++i4)
    for (int i3 = 0; i3 < n3d; ++i3) {
      const double vx = velocities[i3 + n3d * (i4 + n4d * 0)];
      const double vy = velocities[i3 + n3d * (i4 + n4d * 1)];
      for (int i2 = 0; i2 < ND(1); ++i2) {
        const double y = xlo[1] + ((lo[1] + i2) + 0.5) * dx[1];
        for (int i1 = 0; i1 < ND(0); ++i1) {
          const double x = xlo[0] + ((lo[0] + i1) + 0.5) * dx[0];
          const double fexact = alpha / pi * exp(-(alpha * (vx * vx + vy * vy) / 0.2e1)) *
                                (0.1e1 + A * cos(kx * x) * cos(ky * y) * sin(kt * t)) / 0.2e1;
          F4(error, i1, i2, i3, i4) = F4(soln, i1, i2, i3, i4) - fexact;
        }
      }
    }
}

/* TwoSpecies_ElectronTrigTZSource / TwoSpecies_IonTrigTZSource (TwoSpecies_ElectronTZSourceF.f:10-160, :165-262;
 * TwoSpecies_IonTZSourceF.f:10-139, :143-238; deck test/IAWTZ): electrons carry the wave (kxE, kyE) = (4, 2), ions
 * (kxI, kyI) = (2, 4) at twice the amplitude; alpha = sqrt(mass); each species' source couples to both through the field.
 * Only the last assignment of each routine is live in the Fortran (cg, cg1; the exact solutions cg0, cg2).
 * species: 0 = the electron routines, 1 = the ion routines; dparams = {amp, electron_mass, ion_mass}. */
#define TZ2_COMMON                                                                                                  \
  const int64_t n3d = ND(2), n4d = ND(3);                                                                           \
  const double me = dparams[1], mi = dparams[2], A = dparams[0], t = time;                                          \
  const double kxI = 0.2e1, kyI = 0.4e1, ktI = 0.1e1, kxE = 0.4e1, kyE = 0.2e1, ktE = 0.1e1;                        \
  const double pi = 4.0 * atan(1.0);                                                                                \
  const double alphaE = sqrt(me), alphaI = sqrt(mi);                                                                \
  (void)alphaE; (void)alphaI; (void)ktI; (void)ktE; (void)kxI; (void)kyI; (void)kxE; (void)kyE; (void)me; (void)mi;
static inline double tz_p2(double x) { return x * x; }
static inline double tz_p4(double x) { double t = x * x; return t * t; }   /* gfortran's x**4 */
void ok_set_two_species_trig_tz_source(double* f, const ok_geom* g, const int* lo, const double* xlo, const double* dx,
                                       double time, const double* velocities, const double* dparams, int species) {
  TZ2_COMMON
  for (int i4 = 0; i4 < n4d; ++i4)
    for (int i3 = 0; i3 < n3d; ++i3) {
      const double vx = velocities[i3 + n3d * (i4 + n4d * 0)];
      const double vy = velocities[i3 + n3d * (i4 + n4d * 1)];
      for (int i2 = 0; i2 < ND(1); ++i2) {
        const double y = xlo[1] + ((lo[1] + i2) + 0.5) * dx[1];
        for (int i1 = 0; i1 < ND(0); ++i1) {
          const double x = xlo[0] + ((lo[0] + i1) + 0.5) * dx[0];
          double cg;
          if (species == 0) {
            cg = tz_p2(alphaE) / pi * exp(-(tz_p2(alphaE) * (vx * vx + vy * vy) / 0.2e1)) * A * cos(kxE * x) * cos(kyE * y) * ktE *
                     cos(ktE * t) / 0.2e1 -
                 vx * tz_p2(alphaE) / pi * exp(-(tz_p2(alphaE) * (vx * vx + vy * vy) / 0.2e1)) * A * kxE * sin(kxE * x) * cos(kyE * y) *
                     sin(ktE * t) / 0.2e1 -
                 vy * tz_p2(alphaE) / pi * exp(-(tz_p2(alphaE) * (vx * vx + vy * vy) / 0.2e1)) * A * cos(kxE * x) * kyE * sin(kyE * y) *
                     sin(ktE * t) / 0.2e1 -
                 0.1e1 / me *
                     (sin(ktE * t) * sin(kxE * x) * kxE * (tz_p2(kxI) + tz_p2(kyI)) * cos(kyE * y) -
                      0.2e1 * cos(kyI * y) * sin(ktI * t) * sin(kxI * x) * kxI * (tz_p2(kxE) + tz_p2(kyE))) *
                     A / (tz_p2(kxI) + tz_p2(kyI)) / (tz_p2(kxE) + tz_p2(kyE)) * tz_p4(alphaE) / pi * vx *
                     exp(-(tz_p2(alphaE) * (vx * vx + vy * vy) / 0.2e1)) * (0.1e1 + A * cos(kxE * x) * cos(kyE * y) * sin(ktE * t)) / 0.2e1 -
                 0.1e1 / me * A *
                     (sin(ktE * t) * sin(kyE * y) * kyE * (tz_p2(kxI) + tz_p2(kyI)) * cos(kxE * x) -
                      0.2e1 * kyI * cos(kxI * x) * sin(ktI * t) * sin(kyI * y) * (tz_p2(kxE) + tz_p2(kyE))) /
                     (tz_p2(kxI) + tz_p2(kyI)) / (tz_p2(kxE) + tz_p2(kyE)) * tz_p4(alphaE) / pi * vy *
                     exp(-(tz_p2(alphaE) * (vx * vx + vy * vy) / 0.2e1)) * (0.1e1 + A * cos(kxE * x) * cos(kyE * y) * sin(ktE * t)) / 0.2e1;
          } else {
            cg = tz_p2(alphaI) / pi * exp(-(tz_p2(alphaI) * (vx * vx + vy * vy) / 0.2e1)) * A * cos(kxI * x) * cos(kyI * y) * ktI *
                     cos(ktI * t) -
                 vx * tz_p2(alphaI) / pi * exp(-(tz_p2(alphaI) * (vx * vx + vy * vy) / 0.2e1)) * A * kxI * sin(kxI * x) * cos(kyI * y) *
                     sin(ktI * t) -
                 vy * tz_p2(alphaI) / pi * exp(-(tz_p2(alphaI) * (vx * vx + vy * vy) / 0.2e1)) * A * cos(kxI * x) * kyI * sin(kyI * y) *
                     sin(ktI * t) +
                 0.1e1 / mi *
                     (sin(ktE * t) * sin(kxE * x) * kxE * (tz_p2(kxI) + tz_p2(kyI)) * cos(kyE * y) -
                      0.2e1 * cos(kyI * y) * sin(ktI * t) * sin(kxI * x) * kxI * (tz_p2(kxE) + tz_p2(kyE))) *
                     A / (tz_p2(kxI) + tz_p2(kyI)) / (tz_p2(kxE) + tz_p2(kyE)) * tz_p4(alphaI) / pi * vx *
                     exp(-(tz_p2(alphaI) * (vx * vx + vy * vy) / 0.2e1)) * (0.1e1 + 0.2e1 * A * cos(kxI * x) * cos(kyI * y) * sin(ktI * t)) /
                     0.2e1 +
                 0.1e1 / mi * A *
                     (sin(ktE * t) * sin(kyE * y) * kyE * (tz_p2(kxI) + tz_p2(kyI)) * cos(kxE * x) -
                      0.2e1 * kyI * cos(kxI * x) * sin(ktI * t) * sin(kyI * y) * (tz_p2(kxE) + tz_p2(kyE))) /
                     (tz_p2(kxI) + tz_p2(kyI)) / (tz_p2(kxE) + tz_p2(kyE)) * tz_p4(alphaI) / pi * vy *
                     exp(-(tz_p2(alphaI) * (vx * vx + vy * vy) / 0.2e1)) * (0.1e1 + 0.2e1 * A * cos(kxI * x) * cos(kyI * y) * sin(ktI * t)) /
                     0.2e1;
          }
          F4(f, i1, i2, i3, i4) = F4(f, i1, i2, i3, i4) + cg;
        }
      }
    }
}
void ok_compute_two_species_trig_tz_source_error(double* error, const double* soln, const ok_geom* g, const int* lo,
                                                 const double* xlo, const double* dx, double time, const double* velocities,
                                                 const double* dparams, int species) {
  TZ2_COMMON
  for (int i4 = 0; i4 < n4d; ++i4)
    for (int i3 = 0; i3 < n3d; ++i3) {
      const double vx = velocities[i3 + n3d * (i4 + n4d * 0)];
      const double vy = velocities[i3 + n3d * (i4 + n4d * 1)];
      for (int i2 = 0; i2 < ND(1); ++i2) {
        const double y = xlo[1] + ((lo[1] + i2) + 0.5) * dx[1];
        for (int i1 = 0; i1 < ND(0); ++i1) {
          const double x = xlo[0] + ((lo[0] + i1) + 0.5) * dx[0];
          const double fexact =
              species == 0 ? tz_p2(alphaE) / pi * exp(-(tz_p2(alphaE) * (vx * vx + vy * vy) / 0.2e1)) *
                                 (0.1e1 + A * cos(kxE * x) * cos(kyE * y) * sin(ktE * t)) / 0.2e1
                           : tz_p2(alphaI) / pi * exp(-(tz_p2(alphaI) * (vx * vx + vy * vy) / 0.2e1)) *
                                 (0.1e1 + 0.2e1 * A * cos(kxI * x) * cos(kyI * y) * sin(ktI * t)) / 0.2e1;
          F4(error, i1, i2, i3, i4) = F4(soln, i1, i2, i3, i4) - fexact;
        }
      }
    }
}
#undef TZ2_COMMON

/* ------------------------------------------------------------------------------------------
 * Time-history diagnostics.
 * computeke (KineticSpeciesF.f:2447-2500): out = {ke, ke_x, ke_y, px, py}; the running sums start from
 * the incoming values like the Fortran's (the caller zeroes them, KineticSpecies.C:1198-1213).
 * computekemaxwell (KineticSpeciesF.f:2504-2559): out = {ke, ke_x, ke_y} with ke = ke_x + ke_y + ke_z.
 * ------------------------------------------------------------------------------------------ */
void ok_compute_ke(const ok_geom* g, const double* u, double mass, const double* velocities, double* out5) {
  const int ng = g->ng;
  const int64_t n3d = ND(2), n4d = ND(3);
  double ke_x = out5[1], ke_y = out5[2], px = out5[3], py = out5[4];
  const double dx = g->dx[0], dy = g->dx[1], dvx = g->dx[2], dvy = g->dx[3];
  for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
    for (int i3 = ng; i3 < ng + g->n[2]; ++i3) {
      double vx = velocities[i3 + n3d * (i4 + n4d * 0)];
      double vx2 = vx * vx;
      double vy = velocities[i3 + n3d * (i4 + n4d * 1)];
      double vy2 = vy * vy;
      for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
        for (int i1 = ng; i1 < ng + g->n[0]; ++i1) {
          double this_u = F4(u, i1, i2, i3, i4);
          ke_x = ke_x + 0.5 * this_u * vx2;
          ke_y = ke_y + 0.5 * this_u * vy2;
          px = px + this_u * vx;
          py = py + this_u * vy;
        }
    }
  ke_x = ke_x * mass * dx * dy * dvx * dvy;
  ke_y = ke_y * mass * dx * dy * dvx * dvy;
  px = px * mass * dx * dy * dvx * dvy;
  py = py * mass * dx * dy * dvx * dvy;
  out5[0] = ke_x + ke_y;
  out5[1] = ke_x; out5[2] = ke_y; out5[3] = px; out5[4] = py;
}
void ok_compute_ke_maxwell(const ok_geom* g, const double* u, double mass, const double* velocities,
                           const double* vz_in, double* out3) {
  const int ng = g->ng;
  const int64_t n3d = ND(2), n4d = ND(3), n1d = ND(0);
  double ke_x = out3[1], ke_y = out3[2], ke_z = 0.0;
  const double dx = g->dx[0], dy = g->dx[1], dvx = g->dx[2], dvy = g->dx[3];
  for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
    for (int i3 = ng; i3 < ng + g->n[2]; ++i3) {
      double vx = velocities[i3 + n3d * (i4 + n4d * 0)];
      double vx2 = vx * vx;
      double vy = velocities[i3 + n3d * (i4 + n4d * 1)];
      double vy2 = vy * vy;
      for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
        for (int i1 = ng; i1 < ng + g->n[0]; ++i1) {
          double vz = vz_in[i1 + n1d * i2];
          double vz2 = vz * vz;
          double this_u = F4(u, i1, i2, i3, i4);
          ke_x = ke_x + 0.5 * this_u * vx2;
          ke_y = ke_y + 0.5 * this_u * vy2;
          ke_z = ke_z + 0.5 * this_u * vz2;
        }
    }
  ke_x = ke_x * mass * dx * dy * dvx * dvy;
  ke_y = ke_y * mass * dx * dy * dvx * dvy;
  ke_z = ke_z * mass * dx * dy * dvx * dvy;
  out3[0] = ke_x + ke_y + ke_z;
  out3[1] = ke_x; out3[2] = ke_y;
}
/* Poisson::accumulateSequences (Poisson.C:796-860, ncomp = 2): out = {e_max, e_tot, ex_max, ey_max, e_sum_tot};
 * Maxwell::accumulateSequences (Maxwell.C:753-875, ncomp = 6; note its loop nest has i1 outer): out =
 * {e_max, e_tot, ex_max, ey_max, ez_max, e_sum_tot, b_max, b_tot, bx_max, by_max, bz_max, b_sum_tot} */
void ok_field_history(const double* em, int n1, int n2, int ng, int ncomp, const double* dx, double* out) {
  const int64_t n1d = n1 + 2 * ng, pl = n1d * (n2 + 2 * ng);
  const double area = dx[0] * dx[1];
  if (ncomp == 2) {
    double e_sum_tot = 0.0, e_max = 0.0, e_tot = 0.0, ex_max = 0.0, ey_max = 0.0;
    for (int i2 = ng; i2 < ng + n2; ++i2)
      for (int i1 = ng; i1 < ng + n1; ++i1) {
        const double ex = em[i1 + n1d * i2], ey = em[i1 + n1d * i2 + pl];
        double tmp = ex * ex + ey * ey;
        double e_loc = sqrt(tmp);
        e_sum_tot += 0.5 * tmp;
        e_max = fmax(e_max, e_loc);
        e_tot += e_loc;
        ex_max = fmax(ex_max, fabs(ex));
        ey_max = fmax(ey_max, fabs(ey));
      }
    e_sum_tot *= area;
    e_tot *= area;
    out[0] = e_max; out[1] = e_tot; out[2] = ex_max; out[3] = ey_max; out[4] = e_sum_tot;
    return;
  }
  double s[2] = {0.0, 0.0}, mx[2] = {0.0, 0.0}, tot[2] = {0.0, 0.0}, cm[6] = {0, 0, 0, 0, 0, 0};
  for (int i1 = ng; i1 < ng + n1; ++i1)
    for (int i2 = ng; i2 < ng + n2; ++i2)
      for (int h = 0; h < 2; ++h) {
        const double a = em[i1 + n1d * i2 + pl * (3 * h)], b = em[i1 + n1d * i2 + pl * (3 * h + 1)],
                     c = em[i1 + n1d * i2 + pl * (3 * h + 2)];
        double tmp = a * a + b * b + c * c;
        double loc = sqrt(tmp);
        s[h] += 0.5 * tmp;
        mx[h] = fmax(mx[h], loc);
        tot[h] += loc;
        cm[3 * h] = fmax(cm[3 * h], fabs(a));
        cm[3 * h + 1] = fmax(cm[3 * h + 1], fabs(b));
        cm[3 * h + 2] = fmax(cm[3 * h + 2], fabs(c));
      }
  for (int h = 0; h < 2; ++h) {
    out[6 * h] = mx[h]; out[6 * h + 1] = tot[h] * area;
    out[6 * h + 2] = cm[3 * h]; out[6 * h + 3] = cm[3 * h + 1]; out[6 * h + 4] = cm[3 * h + 2];
    out[6 * h + 5] = s[h] * area;
  }
}
