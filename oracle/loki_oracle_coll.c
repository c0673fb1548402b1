/*
 * loki_oracle_coll.c -- CPU ORACLE (test infrastructure, see loki_oracle.h) for the pitch-angle collision operator
 * of LLNL/LOKI: PitchAngleCollisionOperator::evaluate (PitchAngleCollisionOperator.C:61-134) and the Fortran it calls
 * (PitchAngleCollisionOperatorF.f).
 *
 *   C(f) = d/dvx [ nu ( wy^2 df/dvx - wx wy df/dvy ) ] + d/dvy [ nu ( wx^2 df/dvy - wx wy df/dvx ) ],
 *   w = v - V(x,y),  nu = nuCoeff (vth(x,y) / max(|w|, vfloor))^3 alpha(vx) beta(vy)
 *
 * PINNING.  evaluateCollisionality (:11-93), the non-conservative 4th-order operator (:1470-1614) and the four moment
 * routines (:1706-1852) are restated statement by statement and pinned BIT FOR BIT against the transliterated Fortran
 * (tests/test_oracle_pin.py).  The two conservative operators (:97-519 order 4, :523-1462 order 6) are 1 350 lines of
 * Maple output in the reference; they are NOT restated line by line.  The scheme behind the generated code was
 * identified and is written here in operator form (D+ / D- / D0 / delta^2 acting on whole velocity planes):
 *
 *   order 4:  d(a df)   ~ D+[a4 D-f] - h^2/24 D+[a2 D- d2 f] - h^2/24 d2 D+[a2 D-f]
 *             d_x(b d_y f) ~ D0x[b D0y f] - k^2/6 D0x[b D0y d2y f] - h^2/6 D0x d2x [b D0y f]
 *   order 6:  the same expansions carried to h^4 (3/640, 1/576, 1/30, 1/36 terms below)
 *
 *   with a2 / a4 / a6 the 2 / 4 / 6-point interpolation of the coefficient to the face.  These forms are algebraically
 *   identical to the generated code; tests/test_oracle_pin.py holds them to 1e-13 of the largest term against the
 *   transliterated Fortran of BOTH generated routines (measured: 4e-16 and 7e-15), i.e. pinned to rounding, not to the bit.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "loki_oracle.h"

/* evaluateCollisionality, PitchAngleCollisionOperatorF.f:11-93 */
double ok_pitch_angle_collisionality(double vx, double vy, double vxgrid, double vygrid, const double* range_lo,
                                     const double* range_hi, double vxmin, double vxmax, double vymin, double vymax,
                                     double vfloor, double vthermal, double nu_coef, int order) {
  const double vxra = range_lo[0], vxrb = range_hi[0], vyra = range_lo[1], vyrb = range_hi[1];
  double v = fmax(sqrt(vx * vx + vy * vy), vfloor);
  double r = vthermal / v;
  double nuei = nu_coef * ((r * r) * r);
  double xi, eta, va, vb;
  if (vxgrid < vxra && vxgrid >= vxmin) { va = vxra; vb = vxmin; xi = (vxgrid - va) / (vb - va); }
  else if (vxgrid > vxrb && vxgrid <= vxmax) { va = vxrb; vb = vxmax; xi = (vxgrid - va) / (vb - va); }
  else if (vxgrid < vxmin || vxgrid > vxmax) xi = 1.0;
  else xi = 0.0;
  if (vygrid < vyra && vygrid >= vymin) { va = vyra; vb = vymin; eta = (vygrid - va) / (vb - va); }
  else if (vygrid > vyrb && vygrid <= vymax) { va = vyrb; vb = vymax; eta = (vygrid - va) / (vb - va); }
  else if (vygrid < vymin || vygrid > vymax) eta = 1.0;
  else eta = 0.0;
  double alpha, beta;
  if (order == 4) {
    double x2 = xi * xi, e2 = eta * eta;
    alpha = 1.0 + (x2 * x2) * (((20.0 * (x2 * xi) - 70.0 * x2) + 84.0 * xi) - 35.0);
    beta = 1.0 + (e2 * e2) * (((20.0 * (e2 * eta) - 70.0 * e2) + 84.0 * eta) - 35.0);
  } else {
    /* integer powers as gfortran -O2 expands them: x^3 = (x x) x, x^4 = (x^2)^2, x^5 = (x^2 x) x^2, x^6 = (x^3)^2 */
    double x2 = xi * xi, x3 = x2 * xi, e2 = eta * eta, e3 = e2 * eta;
    alpha = 1.0 + (x3 * x3) * (((((252.0 * (x3 * x2) - 1386.0 * (x2 * x2)) + 3080.0 * x3) - 3465.0 * x2) + 1980.0 * xi) - 462.0);
    beta = 1.0 + (e3 * e3) * (((((252.0 * (e3 * e2) - 1386.0 * (e2 * e2)) + 3080.0 * e3) - 3465.0 * e2) + 1980.0 * eta) - 462.0);
  }
  return (nuei * alpha) * beta;
}

/* PitchAngleCollisionOperator.C:69-117: computePitchAngleSpeciesMoments (:1706-1750), three ReductionSchedule4D sums
 * (one rank: the local sum times dvx*dvy, ReductionSchedule4D.C:51-70), ...ReducedFields (:1754-1781), ...KEC
 * (:1785-1824), one more reduction, ...Vthermal (:1828-1852).  All over the WHOLE configuration data box (ghosts too). */
void ok_pitch_angle_fields(double* IVx, double* IVy, double* IVth, const double* u, const ok_geom* g,
                           const double* velocities) {
  const int ng = g->ng;
  const int64_t n1d = ok_nd(g, 0), n2d = ok_nd(g, 1), n3d = ok_nd(g, 2), n4d = ok_nd(g, 3), pl = n1d * n2d;
  const double eps = 1.0e-10, measure = g->dx[2] * g->dx[3];
  double* rN = (double*)calloc(pl, sizeof(double));
  double* rGx = (double*)calloc(pl, sizeof(double));
  double* rGy = (double*)calloc(pl, sizeof(double));
  double* rK = (double*)calloc(pl, sizeof(double));
  /* chunks of configuration-space points on all host cores; every point keeps the reference's order of additions */
  const int64_t CH = 256, nch = (pl + CH - 1) / CH;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < nch; ++c) {
    const int64_t k0 = c * CH, k1 = (k0 + CH < pl) ? k0 + CH : pl;
    for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
      for (int i3 = ng; i3 < ng + g->n[2]; ++i3) {
        const double vx = velocities[i3 + n3d * i4], vy = velocities[i3 + n3d * (i4 + n4d)];
        const double* up = u + (int64_t)(i4 * n3d + i3) * pl;
        for (int64_t k = k0; k < k1; ++k) {
          double uval = fmax(fabs(up[k]), eps);
          rN[k] = rN[k] + uval;
          rGx[k] = rGx[k] + vx * uval;
          rGy[k] = rGy[k] + vy * uval;
        }
      }
    for (int64_t k = k0; k < k1; ++k) { rN[k] *= measure; rGx[k] *= measure; rGy[k] *= measure; }
    for (int64_t k = k0; k < k1; ++k) { IVx[k] = rGx[k] / rN[k]; IVy[k] = rGy[k] / rN[k]; }
    for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
      for (int i3 = ng; i3 < ng + g->n[2]; ++i3) {
        const double vx = velocities[i3 + n3d * i4], vy = velocities[i3 + n3d * (i4 + n4d)];
        const double* up = u + (int64_t)(i4 * n3d + i3) * pl;
        for (int64_t k = k0; k < k1; ++k) {
          double uval = fmax(fabs(up[k]), eps);
          double wx = vx - IVx[k], wy = vy - IVy[k];
          rK[k] = rK[k] + (wx * wx + wy * wy) * uval;
        }
      }
    for (int64_t k = k0; k < k1; ++k) rK[k] *= measure;
    for (int64_t k = k0; k < k1; ++k) IVth[k] = sqrt(0.5 * rK[k] / rN[k]);
  }
  free(rN); free(rGx); free(rGy); free(rK);
}

/* ---- whole-plane difference operators on a (n3d, n4d) velocity plane, periodic shifts: the interior never sees the wrap
 *      because every composite below reaches at most ng cells ---- */
typedef struct { int nx, ny; } pdim;
static inline double at(const pdim* d, const double* a, int i, int j) {
  i = (i % d->nx + d->nx) % d->nx;
  j = (j % d->ny + d->ny) % d->ny;
  return a[i + (int64_t)d->nx * j];
}
#define PLANE_LOOP for (int j = 0; j < d->ny; ++j) for (int i = 0; i < d->nx; ++i)
#define O o[i + (int64_t)d->nx * j]
#define SH(a, k) (ax == 0 ? at(d, a, i + (k), j) : at(d, a, i, j + (k)))
static void Dp(const pdim* d, double* o, const double* a, int ax, double h) { PLANE_LOOP O = (SH(a, 1) - SH(a, 0)) / h; }
static void Dm(const pdim* d, double* o, const double* a, int ax, double h) { PLANE_LOOP O = (SH(a, 0) - SH(a, -1)) / h; }
static void D0(const pdim* d, double* o, const double* a, int ax, double h) { PLANE_LOOP O = (SH(a, 1) - SH(a, -1)) / (2.0 * h); }
static void d2(const pdim* d, double* o, const double* a, int ax, double h) { PLANE_LOOP O = ((SH(a, 1) - 2.0 * SH(a, 0)) + SH(a, -1)) / (h * h); }
/* coefficient at the lower face (i - 1/2) of cell i from 2, 4, 6 cell values */
static void a2(const pdim* d, double* o, const double* a, int ax) { PLANE_LOOP O = 0.5 * (SH(a, 0) + SH(a, -1)); }
static void a4(const pdim* d, double* o, const double* a, int ax) { PLANE_LOOP O = (9.0 * (SH(a, 0) + SH(a, -1)) - (SH(a, 1) + SH(a, -2))) / 16.0; }
static void a6(const pdim* d, double* o, const double* a, int ax) {
  PLANE_LOOP O = ((150.0 * (SH(a, 0) + SH(a, -1)) - 25.0 * (SH(a, 1) + SH(a, -2))) + 3.0 * (SH(a, 2) + SH(a, -3))) / 256.0;
}
static void mul(const pdim* d, double* o, const double* a, const double* b) { PLANE_LOOP O = at(d, a, i, j) * at(d, b, i, j); }
static void axpy(const pdim* d, double* o, double c, const double* a) { PLANE_LOOP O += c * at(d, a, i, j); }
#undef SH
#undef O
#undef PLANE_LOOP

/* out += d/dv_ax ( co d f / dv_ax ), conservative, order 4 or 6 */
static void diag_term(const pdim* d, double* out, const double* co, const double* f, int ax, double h, int order,
                      double* w[8]) {
  double *cf = w[0], *e = w[1], *t = w[2], *s = w[3], *G = w[4], *q = w[5];
  const double h2 = h * h, h4 = h2 * h2;
  Dm(d, e, f, ax, h);                                   /* e = D- f at the lower faces */
  (order == 4 ? a4 : a6)(d, cf, co, ax);
  mul(d, t, cf, e); Dp(d, s, t, ax, h); axpy(d, out, 1.0, s);            /* D+[a D- f] with the full-order face value */
  (order == 4 ? a2 : a4)(d, cf, co, ax);
  d2(d, t, f, ax, h); Dm(d, s, t, ax, h); mul(d, t, cf, s); Dp(d, s, t, ax, h); axpy(d, out, -h2 / 24.0, s);
  mul(d, t, cf, e); Dp(d, G, t, ax, h); d2(d, s, G, ax, h); axpy(d, out, -h2 / 24.0, s);
  if (order == 6) {
    a2(d, cf, co, ax);
    d2(d, t, f, ax, h); d2(d, s, t, ax, h); Dm(d, t, s, ax, h); mul(d, s, cf, t); Dp(d, t, s, ax, h);
    axpy(d, out, 3.0 * h4 / 640.0, t);                                    /* D+[a2 D- d4 f] */
    mul(d, t, cf, e); Dp(d, G, t, ax, h); d2(d, s, G, ax, h); d2(d, t, s, ax, h);
    axpy(d, out, 3.0 * h4 / 640.0, t);                                    /* d4 D+[a2 D- f] */
    d2(d, t, f, ax, h); Dm(d, s, t, ax, h); mul(d, t, cf, s); Dp(d, q, t, ax, h); d2(d, s, q, ax, h);
    axpy(d, out, h4 / 576.0, s);                                          /* d2 D+[a2 D- d2 f] */
  }
}
/* out -= d/dv_ax ( b d f / dv_ay ), order 4 or 6 */
static void cross_term(const pdim* d, double* out, const double* b, const double* f, int ax, double h, int ay, double k,
                       int order, double* w[8]) {
  double *g = w[0], *t = w[1], *s = w[2], *q = w[3], *r = w[4];
  const double h2 = h * h, k2 = k * k;
  D0(d, t, f, ay, k); mul(d, g, b, t);                                    /* g = b D0y f */
  D0(d, s, g, ax, h); axpy(d, out, -1.0, s);
  d2(d, t, f, ay, k); D0(d, s, t, ay, k); mul(d, q, b, s);                /* q = b D0y d2y f */
  D0(d, s, q, ax, h); axpy(d, out, k2 / 6.0, s);
  d2(d, t, g, ax, h); D0(d, s, t, ax, h); axpy(d, out, h2 / 6.0, s);
  if (order == 6) {
    d2(d, t, f, ay, k); d2(d, s, t, ay, k); D0(d, t, s, ay, k); mul(d, r, b, t); D0(d, s, r, ax, h);
    axpy(d, out, -k2 * k2 / 30.0, s);                                     /* D0x[b D0y d4y f] */
    d2(d, t, g, ax, h); d2(d, s, t, ax, h); D0(d, t, s, ax, h);
    axpy(d, out, -h2 * h2 / 30.0, t);                                     /* D0x d4x [b D0y f] */
    d2(d, t, q, ax, h); D0(d, s, t, ax, h);
    axpy(d, out, -h2 * k2 / 36.0, s);                                     /* D0x d2x [b D0y d2y f] */
  }
}

/* appendPitchAngleCollision, PitchAngleCollisionOperatorF.f:1618-1702.  velocities (n3d,n4d,2); IVx, IVy, IVth (n1d,n2d);
 * vlo / vhi: the velocity domain (xlo(3:4), xhi(3:4)); non-relativistic.  rhs += C(f) on the interior. */
void ok_append_pitch_angle_collision(double* rhs, const double* f, const ok_geom* g, const double* velocities,
                                     const double* IVx, const double* IVy, const double* IVth, const double* vlo,
                                     const double* vhi, const double* range_lo, const double* range_hi, double vfloor,
                                     double nu_coef, int conservative) {
  const int ng = g->ng, order = g->order;
  const int64_t n1d = ok_nd(g, 0), n2d = ok_nd(g, 1), n3d = ok_nd(g, 2), n4d = ok_nd(g, 3), pl = n1d * n2d;
  const double dvx = g->dx[2], dvy = g->dx[3];
  const int vrolloff = (order == 4) ? 3 : 4;   /* :285, :949 */
  const double vxmin = vlo[0] + vrolloff * dvx, vxmax = vhi[0] - vrolloff * dvx;
  const double vymin = vlo[1] + vrolloff * dvy, vymax = vhi[1] - vrolloff * dvy;
  if (!conservative) {
    if (order != 4) return;   /* :1686-1699: nothing is applied */
    /* nonConservativePitchAngle_4th, :1470-1614 */
    for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
      for (int i3 = ng; i3 < ng + g->n[2]; ++i3)
        for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
          for (int i1 = ng; i1 < ng + g->n[0]; ++i1) {
            const int64_t c2 = i1 + n1d * i2;
            const double vxgrid = velocities[i3 + n3d * i4], vygrid = velocities[i3 + n3d * (i4 + n4d)];
            const double vx = vxgrid - IVx[c2], vy = vygrid - IVy[c2];
            const double nuei = ok_pitch_angle_collisionality(vx, vy, vxgrid, vygrid, range_lo, range_hi, vxmin, vxmax, vymin,
                                                              vymax, vfloor, IVth[c2], nu_coef, 4);
#define F(a, b) f[c2 + pl * ((i3 + (a)) + n3d * (int64_t)(i4 + (b)))]
#define D1X(b) ((((-1.0 * F(2, b) + 8.0 * F(1, b)) - 8.0 * F(-1, b)) + 1.0 * F(-2, b)) / (12.0 * dvx))
            const double fvxp2 = D1X(2), fvxp1 = D1X(1), fvxm1 = D1X(-1), fvxm2 = D1X(-2);
            const double fvxvy = (((-1.0 * fvxp2 + 8.0 * fvxp1) - 8.0 * fvxm1) + 1.0 * fvxm2) / (12.0 * dvy);
            const double fvxvx = ((((-1.0 * F(2, 0) + 16.0 * F(1, 0)) - 30.0 * F(0, 0)) + 16.0 * F(-1, 0)) - 1.0 * F(-2, 0)) / (12.0 * (dvx * dvx));
            const double fvyvy = ((((-1.0 * F(0, 2) + 16.0 * F(0, 1)) - 30.0 * F(0, 0)) + 16.0 * F(0, -1)) - 1.0 * F(0, -2)) / (12.0 * (dvy * dvy));
            const double fvx = D1X(0);
            const double fvy = (((-1.0 * F(0, 2) + 8.0 * F(0, 1)) - 8.0 * F(0, -1)) + 1.0 * F(0, -2)) / (12.0 * dvy);
#undef D1X
#undef F
            const double temp = nuei * (((((vx * vx) * fvyvy - ((2.0 * vx) * vy) * fvxvy) + (vy * vy) * fvxvx) - vy * fvy) - vx * fvx);
            const int64_t c = c2 + pl * (i3 + n3d * (int64_t)i4);
            rhs[c] = rhs[c] + temp;
          }
    return;
  }
  /* conservative: the operator form of the header comment, one configuration-space point at a time */
  pdim dd = {(int)n3d, (int)n4d};
  const int64_t pv = n3d * n4d;
  /* configuration-space points are independent (disjoint outputs): all host cores, each with its own work planes; the
   * bits do not depend on the split */
#pragma omp parallel
  {
  double* buf = (double*)malloc(sizeof(double) * pv * 13);
  double *F = buf, *A = buf + pv, *B = buf + 2 * pv, *Cc = buf + 3 * pv, *out = buf + 4 * pv, *w[8];
  for (int k = 0; k < 8; ++k) w[k] = buf + (5 + k) * pv;
#pragma omp for collapse(2) schedule(static)
  for (int i2 = ng; i2 < ng + g->n[1]; ++i2)
    for (int i1 = ng; i1 < ng + g->n[0]; ++i1) {
      const int64_t c2 = i1 + n1d * i2;
      for (int i4 = 0; i4 < n4d; ++i4)
        for (int i3 = 0; i3 < n3d; ++i3) {
          const int64_t p = i3 + n3d * i4;
          const double vxgrid = velocities[p], vygrid = velocities[p + pv];
          const double wx = vxgrid - IVx[c2], wy = vygrid - IVy[c2];
          const double nu = ok_pitch_angle_collisionality(wx, wy, vxgrid, vygrid, range_lo, range_hi, vxmin, vxmax, vymin, vymax,
                                                          vfloor, IVth[c2], nu_coef, order);
          F[p] = f[c2 + pl * p];
          A[p] = nu * (wy * wy);
          B[p] = (nu * wx) * wy;
          Cc[p] = nu * (wx * wx);
          out[p] = 0.0;
        }
      diag_term(&dd, out, A, F, 0, dvx, order, w);
      diag_term(&dd, out, Cc, F, 1, dvy, order, w);
      cross_term(&dd, out, B, F, 0, dvx, 1, dvy, order, w);
      cross_term(&dd, out, B, F, 1, dvy, 0, dvx, order, w);
      for (int i4 = ng; i4 < ng + g->n[3]; ++i4)
        for (int i3 = ng; i3 < ng + g->n[2]; ++i3) {
          const int64_t p = i3 + n3d * i4;
          rhs[c2 + pl * p] = rhs[c2 + pl * p] + out[p];
        }
    }
  free(buf);
  }
}

/* PitchAngleCollisionOperator::computeRealLam, PitchAngleCollisionOperator.C:137-144 */
double ok_pitch_angle_real_lam(const ok_geom* g, double nu_coef, double vthermal_dt, double vfloor) {
  const double dv = fmin(g->dx[2], g->dx[3]);
  const double pi = 4.0 * atan(1.0);
  return nu_coef * pow(vthermal_dt, 3.0) * pi * pi / (dv * dv * fmax(dv, vfloor));
}
