// lk_coll.cuh -- per-cell arithmetic of the pitch-angle collision operator:
// PitchAngleCollisionOperator::evaluate (PitchAngleCollisionOperator.C:61-134) and the Fortran it calls
// (PitchAngleCollisionOperatorF.f: evaluateCollisionality :11-93, conservativePitchAngle_4th :97-519, _6th :523-1462,
// nonConservativePitchAngle_4th :1470-1614).
//
//   C(f) = d/dvx [ nu ( wy^2 df/dvx - wx wy df/dvy ) ] + d/dvy [ nu ( wx^2 df/dvy - wx wy df/dvx ) ],
//   w = v - V(x,y),  nu = nuCoeff (vth(x,y) / max(|w|, vfloor))^3 alpha(vx) beta(vy)
//
// The two conservative routines are 1 350 lines of Maple output in the reference.  Here they are the finite-difference
// scheme that output encodes, written as difference operators on the (2 ng + 1)^2 velocity window of one cell:
//
//   d(a df)      ~ D+[a_O D-f] - h^2/24 ( D+[a_{O-2} D- d2 f] + d2 D+[a_{O-2} D-f] )            (+ the h^4 terms for O = 6)
//   d_x(b d_y f) ~ D0x[b D0y f] - k^2/6 D0x[b D0y d2y f] - h^2/6 D0x d2x [b D0y f]              (+ the h^4 terms for O = 6)
//
// a_2 / a_4 / a_6: 2 / 4 / 6-point interpolation of the cell coefficient to the face.  The collisionality is needed on
// the plus-shaped 4 ng + 1 cells of the window only.  Plain IEEE arithmetic without contraction (the translation unit
// is built with -fmad=false), operation for operation the CPU oracle's order (oracle/loki_oracle_coll.c).
//
// Everything here is LK_HD so that tests/ can also compile these functions for the host and compare them cell by cell
// with the oracle where there is no GPU (tests/test_cpu_coll_host.py); the library itself only runs them in kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define LK_HD __host__ __device__ __forceinline__
#define LK_HDM __host__ __device__ __forceinline__
#else
#define LK_HD static inline
#define LK_HDM inline
#endif

namespace lkcoll {

typedef long long i64;

struct Params {
  double range_lo[2], range_hi[2];  // collision_vel_range_lo / _hi
  double vmin[2], vmax[2];          // velocity domain shrunk by the roll-off width (3 cells order 4, 4 cells order 6)
  double vfloor, nu_coef;
  double dvx, dvy;
};

// one direction's roll-off coordinate: 0 inside the collisional range, 0..1 across the roll-off, 1 outside (:33-62)
LK_HD double rolloff_coord(double v, double ra, double rb, double vmin, double vmax) {
  if (v < ra && v >= vmin) return (v - ra) / (vmin - ra);
  if (v > rb && v <= vmax) return (v - rb) / (vmax - rb);
  if (v < vmin || v > vmax) return 1.0;
  return 0.0;
}

// evaluateCollisionality, PitchAngleCollisionOperatorF.f:11-93.  Integer powers as the compiled reference forms them.
template <int ORDER>
LK_HD double collisionality(double wx, double wy, double vxgrid, double vygrid, const Params& p, double vthermal) {
  const double v = fmax(sqrt(wx * wx + wy * wy), p.vfloor);
  const double r = vthermal / v;
  const double nuei = p.nu_coef * ((r * r) * r);
  const double xi = rolloff_coord(vxgrid, p.range_lo[0], p.range_hi[0], p.vmin[0], p.vmax[0]);
  const double eta = rolloff_coord(vygrid, p.range_lo[1], p.range_hi[1], p.vmin[1], p.vmax[1]);
  double alpha, beta;
  if constexpr (ORDER == 4) {
    const double x2 = xi * xi, e2 = eta * eta;
    alpha = 1.0 + (x2 * x2) * (((20.0 * (x2 * xi) - 70.0 * x2) + 84.0 * xi) - 35.0);
    beta = 1.0 + (e2 * e2) * (((20.0 * (e2 * eta) - 70.0 * e2) + 84.0 * eta) - 35.0);
  } else {
    const double x2 = xi * xi, x3 = x2 * xi, e2 = eta * eta, e3 = e2 * eta;
    alpha = 1.0 + (x3 * x3) * (((((252.0 * (x3 * x2) - 1386.0 * (x2 * x2)) + 3080.0 * x3) - 3465.0 * x2) + 1980.0 * xi) - 462.0);
    beta = 1.0 + (e3 * e3) * (((((252.0 * (e3 * e2) - 1386.0 * (e2 * e2)) + 3080.0 * e3) - 3465.0 * e2) + 1980.0 * eta) - 462.0);
  }
  return (nuei * alpha) * beta;
}

// ---- difference operators on a line of 2R+1 values stored with the centre at index R ----
// face k is the lower face of cell k (between cells k-1 and k)
template <int O>
LK_HD double face_avg(const double* c, int k) {  // c: centre pointer
  if constexpr (O == 2) return 0.5 * (c[k] + c[k - 1]);
  else if constexpr (O == 4) return (9.0 * (c[k] + c[k - 1]) - (c[k + 1] + c[k - 2])) / 16.0;
  else return ((150.0 * (c[k] + c[k - 1]) - 25.0 * (c[k + 1] + c[k - 2])) + 3.0 * (c[k + 2] + c[k - 3])) / 256.0;
}
LK_HD double dm(const double* a, int k, double h) { return (a[k] - a[k - 1]) / h; }
LK_HD double dp(const double* a, int k, double h) { return (a[k + 1] - a[k]) / h; }
LK_HD double d0(const double* a, int k, double h) { return (a[k + 1] - a[k - 1]) / (2.0 * h); }
LK_HD double dd(const double* a, int k, double h) { return ((a[k + 1] - 2.0 * a[k]) + a[k - 1]) / (h * h); }

// out + d/dv ( co df/dv ) along one line; f, co: centre pointers of lines of 2R+1 values, R = ORDER/2
template <int ORDER>
LK_HD double diag_line(double out, const double* f, const double* co, double h) {
  constexpr int R = ORDER / 2;
  const double h2 = h * h, h4 = h2 * h2;
  // every intermediate on offsets -R..R (centre at index R); only the entries a term reaches are formed
  double e_[2 * R + 1], t_[2 * R + 1], s_[2 * R + 1], g_[2 * R + 1], u_[2 * R + 1];
  double *e = e_ + R, *t = t_ + R, *s = s_ + R, *G = g_ + R, *u = u_ + R;
  for (int k = -R + 1; k <= R; ++k) e[k] = dm(f, k, h);  // D- f at the lower faces
  // D+[a_O D- f]
  for (int k = 0; k <= 1; ++k) t[k] = face_avg<ORDER>(co, k) * e[k];
  out = out + 1.0 * dp(t, 0, h);
  // -h^2/24 D+[a_{O-2} D- d2 f]
  for (int k = -1; k <= 1; ++k) u[k] = dd(f, k, h);
  for (int k = 0; k <= 1; ++k) t[k] = face_avg<ORDER - 2>(co, k) * dm(u, k, h);
  out = out + (-h2 / 24.0) * dp(t, 0, h);
  // -h^2/24 d2 D+[a_{O-2} D- f]
  for (int k = -1; k <= 2; ++k) t[k] = face_avg<ORDER - 2>(co, k) * e[k];
  for (int k = -1; k <= 1; ++k) G[k] = dp(t, k, h);
  out = out + (-h2 / 24.0) * dd(G, 0, h);
  if constexpr (ORDER == 6) {
    // 3h^4/640 D+[a_2 D- d4 f]
    for (int k = -2; k <= 2; ++k) u[k] = dd(f, k, h);
    for (int k = -1; k <= 1; ++k) s[k] = dd(u, k, h);
    for (int k = 0; k <= 1; ++k) t[k] = face_avg<2>(co, k) * dm(s, k, h);
    out = out + (3.0 * h4 / 640.0) * dp(t, 0, h);
    // 3h^4/640 d4 D+[a_2 D- f]
    for (int k = -2; k <= 3; ++k) t[k] = face_avg<2>(co, k) * e[k];
    for (int k = -2; k <= 2; ++k) G[k] = dp(t, k, h);
    for (int k = -1; k <= 1; ++k) s[k] = dd(G, k, h);
    out = out + (3.0 * h4 / 640.0) * dd(s, 0, h);
    // h^4/576 d2 D+[a_2 D- d2 f]
    for (int k = -1; k <= 2; ++k) t[k] = face_avg<2>(co, k) * dm(u, k, h);
    for (int k = -1; k <= 1; ++k) G[k] = dp(t, k, h);
    out = out + (h4 / 576.0) * dd(G, 0, h);
  }
  return out;
}

// out - d/dv_a ( b df/dv_c ): W(a, c) is the window value at offset a along the outer direction and c along the inner
// one; b: centre pointer of the coefficient on the outer line (inner offset 0); h, k: outer / inner cell size
template <int ORDER, class Win>
LK_HD double cross_line(double out, const Win& W, const double* b, double h, double k) {
  constexpr int R = ORDER / 2;
  const double h2 = h * h, k2 = k * k;
  double g_[2 * R + 1], q_[2 * R + 1], t_[2 * R + 1], s_[2 * R + 1], c_[2 * R + 1], c2_[2 * R + 1];
  double *g = g_ + R, *q = q_ + R, *t = t_ + R, *s = s_ + R, *col = c_ + R, *col2 = c2_ + R;
  // g = b D0c f on the outer line
  for (int a = -R; a <= R; ++a) g[a] = b[a] * ((W(a, 1) - W(a, -1)) / (2.0 * k));
  out = out + (-1.0) * d0(g, 0, h);
  // q = b D0c d2c f
  for (int a = -(R - 1); a <= R - 1; ++a) {
    for (int c = -1; c <= 1; ++c) col[c] = ((W(a, c + 1) - 2.0 * W(a, c)) + W(a, c - 1)) / (k * k);
    q[a] = b[a] * d0(col, 0, k);
  }
  out = out + (k2 / 6.0) * d0(q, 0, h);
  // D0a d2a [g]
  for (int a = -1; a <= 1; ++a) t[a] = dd(g, a, h);
  out = out + (h2 / 6.0) * d0(t, 0, h);
  if constexpr (ORDER == 6) {
    // -k^4/30 D0a[b D0c d4c f]
    for (int a = -1; a <= 1; a += 2) {
      for (int c = -2; c <= 2; ++c) col[c] = ((W(a, c + 1) - 2.0 * W(a, c)) + W(a, c - 1)) / (k * k);
      for (int c = -1; c <= 1; ++c) col2[c] = dd(col, c, k);
      s[a] = b[a] * d0(col2, 0, k);
    }
    out = out + (-k2 * k2 / 30.0) * d0(s, 0, h);
    // -h^4/30 D0a d4a [g]
    for (int a = -2; a <= 2; ++a) t[a] = dd(g, a, h);
    for (int a = -1; a <= 1; ++a) s[a] = dd(t, a, h);
    out = out + (-h2 * h2 / 30.0) * d0(s, 0, h);
    // -h^2 k^2/36 D0a d2a [q]
    for (int a = -1; a <= 1; ++a) t[a] = dd(q, a, h);
    out = out + (-h2 * k2 / 36.0) * d0(t, 0, h);
  }
  return out;
}

template <int R>
struct WinXY {  // outer = vx, inner = vy
  const double* w;  // (2R+1)^2 window, w[(c + R) * (2R+1) + (a + R)]
  LK_HDM double operator()(int a, int c) const { return w[(c + R) * (2 * R + 1) + (a + R)]; }
};
template <int R>
struct WinYX {  // outer = vy, inner = vx
  const double* w;
  LK_HDM double operator()(int a, int c) const { return w[(a + R) * (2 * R + 1) + (c + R)]; }
};

// The conservative operator at one cell.  fc: pointer to f at the cell; s3, s4: strides of the two velocity
// directions; vel: the (n3d, n4d, 2) table, pv = n3d * n4d, p = i3 + n3d * i4 of the cell; ivx, ivy, vth at (i1, i2).
template <int ORDER>
LK_HD double conservative_cell(const double* fc, i64 s3, i64 s4, const double* vel, i64 pv, i64 p, int n3d, double ivx,
                               double ivy, double vth, const Params& P) {
  constexpr int R = ORDER / 2, W = 2 * R + 1;
  double win[W * W];
  for (int c = -R; c <= R; ++c)
    for (int a = -R; a <= R; ++a) win[(c + R) * W + (a + R)] = fc[a * s3 + c * s4];
  double Ax_[W], Bx_[W], Cy_[W], By_[W], fx_[W], fy_[W];
  double *Ax = Ax_ + R, *Bx = Bx_ + R, *Cy = Cy_ + R, *By = By_ + R, *fx = fx_ + R, *fy = fy_ + R;
  for (int a = -R; a <= R; ++a) {
    {
      const i64 q = p + a;  // along vx
      const double vxg = vel[q], vyg = vel[q + pv];
      const double wx = vxg - ivx, wy = vyg - ivy;
      const double nu = collisionality<ORDER>(wx, wy, vxg, vyg, P, vth);
      Ax[a] = nu * (wy * wy);
      Bx[a] = (nu * wx) * wy;
      fx[a] = win[R * W + (a + R)];
    }
    {
      const i64 q = p + (i64)a * n3d;  // along vy
      const double vxg = vel[q], vyg = vel[q + pv];
      const double wx = vxg - ivx, wy = vyg - ivy;
      const double nu = collisionality<ORDER>(wx, wy, vxg, vyg, P, vth);
      Cy[a] = nu * (wx * wx);
      By[a] = (nu * wx) * wy;
      fy[a] = win[(a + R) * W + R];
    }
  }
  double out = 0.0;
  out = diag_line<ORDER>(out, fx, Ax, P.dvx);
  out = diag_line<ORDER>(out, fy, Cy, P.dvy);
  WinXY<R> wxy = {win};
  WinYX<R> wyx = {win};
  out = cross_line<ORDER>(out, wxy, Bx, P.dvx, P.dvy);
  out = cross_line<ORDER>(out, wyx, By, P.dvy, P.dvx);
  return out;
}

// nonConservativePitchAngle_4th at one cell, PitchAngleCollisionOperatorF.f:1470-1614 (statement for statement)
LK_HD double nonconservative4_cell(const double* fc, i64 s3, i64 s4, double vxgrid, double vygrid, double ivx, double ivy,
                                   double vth, const Params& P) {
  const double dvx = P.dvx, dvy = P.dvy;
  const double vx = vxgrid - ivx, vy = vygrid - ivy;
  const double nuei = collisionality<4>(vx, vy, vxgrid, vygrid, P, vth);
#define LKC_F(a, b) fc[(a) * s3 + (b) * s4]
#define LKC_D1X(b) ((((-1.0 * LKC_F(2, b) + 8.0 * LKC_F(1, b)) - 8.0 * LKC_F(-1, b)) + 1.0 * LKC_F(-2, b)) / (12.0 * dvx))
  const double fvxp2 = LKC_D1X(2), fvxp1 = LKC_D1X(1), fvxm1 = LKC_D1X(-1), fvxm2 = LKC_D1X(-2);
  const double fvxvy = (((-1.0 * fvxp2 + 8.0 * fvxp1) - 8.0 * fvxm1) + 1.0 * fvxm2) / (12.0 * dvy);
  const double fvxvx = ((((-1.0 * LKC_F(2, 0) + 16.0 * LKC_F(1, 0)) - 30.0 * LKC_F(0, 0)) + 16.0 * LKC_F(-1, 0)) - 1.0 * LKC_F(-2, 0)) / (12.0 * (dvx * dvx));
  const double fvyvy = ((((-1.0 * LKC_F(0, 2) + 16.0 * LKC_F(0, 1)) - 30.0 * LKC_F(0, 0)) + 16.0 * LKC_F(0, -1)) - 1.0 * LKC_F(0, -2)) / (12.0 * (dvy * dvy));
  const double fvx = LKC_D1X(0);
  const double fvy = (((-1.0 * LKC_F(0, 2) + 8.0 * LKC_F(0, 1)) - 8.0 * LKC_F(0, -1)) + 1.0 * LKC_F(0, -2)) / (12.0 * dvy);
#undef LKC_D1X
#undef LKC_F
  return nuei * (((((vx * vx) * fvyvy - ((2.0 * vx) * vy) * fvxvy) + (vy * vy) * fvxvx) - vy * fvy) - vx * fvx);
}

// C(f) at one interior cell: the dispatch of appendPitchAngleCollision (:1618-1702).  Returns 0 where the reference
// applies nothing (non-conservative, order 6).
template <int ORDER>
LK_HD double collision_cell(const double* f, i64 pl, int n3d, int n4d, i64 c2, int i3, int i4, const double* vel,
                            const double* ivx, const double* ivy, const double* vth, const Params& P, int conservative) {
  const i64 pv = (i64)n3d * n4d, p = i3 + (i64)n3d * i4;
  const i64 s3 = pl, s4 = pl * n3d;
  const double* fc = f + c2 + pl * p;
  if (conservative) return conservative_cell<ORDER>(fc, s3, s4, vel, pv, p, n3d, ivx[c2], ivy[c2], vth[c2], P);
  if constexpr (ORDER == 4) return nonconservative4_cell(fc, s3, s4, vel[p], vel[p + pv], ivx[c2], ivy[c2], vth[c2], P);
  return 0.0;
}

// ---- the reduced fields of PitchAngleCollisionOperator::evaluate (PitchAngleCollisionOperator.C:69-117) at one
//      configuration-space point; u: pointer to the point's first velocity cell, pl: the velocity stride.  The sums run
//      over the interior velocity cells with i3 fastest: the reference's accumulation order per (i1, i2). ----
// computePitchAngleSpeciesMoments, PitchAngleCollisionOperatorF.f:1706-1750 (sums added to rn, rgx, rgy)
LK_HD void moments_point(const double* u, i64 pl, int ng, int n3, int n4, int n3d, const double* vel, i64 pv, double& rn,
                         double& rgx, double& rgy) {
  const double eps = 1.0e-10;
  for (int i4 = ng; i4 < ng + n4; ++i4)
    for (int i3 = ng; i3 < ng + n3; ++i3) {
      const i64 p = i3 + (i64)n3d * i4;
      const double vx = vel[p], vy = vel[p + pv];
      const double uval = fmax(fabs(u[pl * p]), eps);
      rn = rn + uval;
      rgx = rgx + vx * uval;
      rgy = rgy + vy * uval;
    }
}
// computePitchAngleSpeciesKEC, :1785-1824 (sum added to rk)
LK_HD void kec_point(const double* u, i64 pl, int ng, int n3, int n4, int n3d, const double* vel, i64 pv, double vx0,
                     double vy0, double& rk) {
  const double eps = 1.0e-10;
  for (int i4 = ng; i4 < ng + n4; ++i4)
    for (int i3 = ng; i3 < ng + n3; ++i3) {
      const i64 p = i3 + (i64)n3d * i4;
      const double vx = vel[p], vy = vel[p + pv];
      const double uval = fmax(fabs(u[pl * p]), eps);
      const double wx = vx - vx0, wy = vy - vy0;
      rk = rk + (wx * wx + wy * wy) * uval;
    }
}
// the whole chain at one point: the three ReductionSchedule4D sums of a rank that holds all of velocity space are the
// local sums times dvx dvy (ReductionSchedule4D.C:51-70), then ...ReducedFields (:1754-1781), ...KEC, ...Vthermal (:1828-1852)
LK_HD void fields_point(const double* u, i64 pl, int ng, int n3, int n4, int n3d, const double* vel, i64 pv, double measure,
                        double& ivx, double& ivy, double& vth) {
  double rn = 0.0, rgx = 0.0, rgy = 0.0, rk = 0.0;
  moments_point(u, pl, ng, n3, n4, n3d, vel, pv, rn, rgx, rgy);
  rn *= measure;
  rgx *= measure;
  rgy *= measure;
  ivx = rgx / rn;
  ivy = rgy / rn;
  kec_point(u, pl, ng, n3, n4, n3d, vel, pv, ivx, ivy, rk);
  rk *= measure;
  vth = sqrt(0.5 * rk / rn);
}

}  // namespace lkcoll
