// lk_bcs.cu -- setAdvectionBCs4D (KineticSpeciesF.f:1166-1297): the physical x / y boundaries of a
// configuration-space direction that is NOT periodic (all five benchmark decks are periodic,
// where the reference skips this routine's body and only wraps).  Same rule as the velocity boundaries:
// outflow by the sign of the face velocity at the boundary face -> u_g = 3u_-1 - 3u_-2 + u_-3 marching
// outward, inflow -> the initial condition (the IC classes' cached tables, lk_inflow).  x boundaries first
// over the full data box of (y, vx, vy), then y boundaries over the full x extent, ghosts just set included.
// Compiled with -fmad=false: the extrapolation is the reference's expression, bit for bit in both modes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/loki_b200.h"
// the face-acceleration expressions of setphasespacevel4D / maxwell4D (strict flavour: this file is built
// with -fmad=false); only inline device functions and plain structs come from this header
#define LK_STRICT 1
#include "lk_device.cuh"

namespace lkbcs {

typedef long long i64;

struct Geo {
  int n[4], nd[4], ng;
  i64 s[4];
};

__device__ __forceinline__ double inflow_value(const lk_inflow& ic, const Geo& g, int i1, int i2, int i3, int i4) {
  const i64 pxy = i1 + (i64)g.nd[0] * i2;
  const i64 pv = i3 + (i64)g.nd[2] * i4;
  switch (ic.kind) {
    case 1:  // PerturbedMaxwellianIC.C:279-281
      return ic.fnorm * ic.fv[pv] * ic.fx[pxy] * ic.frac;
    case 2:  // InterpenetratingStreamIC.C:275-278
      return ic.fx[pxy] * ic.fv[pv] + ic.fx2[pxy] * ic.fv2[pv];
    case 4:  // InterpenetratingStreamIC.C:279-281
      return ic.fv[pv] * ic.fx[pxy] * ic.fx2[pxy];
    case 3: {  // velocity-ghost layers of a cached non-factorable IC (only meaningful at velocity boundaries)
      if (i3 < g.ng || i3 >= g.ng + g.n[2]) {
        const int layer = (i3 < g.ng) ? i3 : (i3 - g.n[2]);
        return ic.ghost3[pxy + (i64)g.nd[0] * g.nd[1] * (layer + (i64)2 * g.ng * i4)];
      }
      const int layer = (i4 < g.ng) ? i4 : (i4 - g.n[3]);
      return ic.ghost4[pxy + (i64)g.nd[0] * g.nd[1] * (i3 + (i64)g.nd[2] * layer)];
    }
    default:
      return 0.0;
  }
}

// pass 0: x boundaries, one thread per (i2,i3,i4); pass 1: y boundaries, one thread per (i1,i3,i4)
__global__ void k_advection_bcs(Geo g, const double* __restrict__ vel, lk_inflow ic, double* __restrict__ u, int pass,
                                int at_lo, int at_hi) {
  const int ng = g.ng;
  const int nfast = (pass == 0) ? g.nd[1] : g.nd[0];
  const i64 total = (i64)nfast * g.nd[2] * g.nd[3];
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int ia = (int)(t % nfast);
  const i64 r = t / nfast;
  const int i3 = (int)(r % g.nd[2]), i4 = (int)(r / g.nd[2]);
  // vel1(n1b+1,i2,i3,i4) = vx(i3,i4), vel2(n2b+1,i3,i4,i1) = vy(i3,i4): the face velocities of
  // initializeVelocity (KineticSpecies.C:1664-1693) do not depend on x, y
  const double v = vel[i3 + (i64)g.nd[2] * (i4 + (i64)(pass == 0 ? 0 : g.nd[3]))];
  const int d = pass;                       // boundary direction
  const int na = ng, nb = ng + g.n[d] - 1;
  const i64 s = g.s[d];
  const i64 base = (pass == 0) ? (g.s[1] * ia + g.s[2] * i3 + g.s[3] * i4) : ((i64)ia + g.s[2] * i3 + g.s[3] * i4);
  double* p = u + base;
  if (at_hi) {
    if (v >= 0.0) {
      for (int ig = 1; ig <= ng; ++ig)
        p[(nb + ig) * s] = 3.0 * p[(nb + ig - 1) * s] - 3.0 * p[(nb + ig - 2) * s] + p[(nb + ig - 3) * s];
    } else {
      for (int ig = 1; ig <= ng; ++ig)
        p[(nb + ig) * s] = (pass == 0) ? inflow_value(ic, g, nb + ig, ia, i3, i4) : inflow_value(ic, g, ia, nb + ig, i3, i4);
    }
  }
  if (at_lo) {
    if (v > 0.0) {
      for (int ig = 1; ig <= ng; ++ig)
        p[(na - ig) * s] = (pass == 0) ? inflow_value(ic, g, na - ig, ia, i3, i4) : inflow_value(ic, g, ia, na - ig, i3, i4);
    } else {
      for (int ig = 1; ig <= ng; ++ig)
        p[(na - ig) * s] = 3.0 * p[(na - ig + 1) * s] - 3.0 * p[(na - ig + 2) * s] + p[(na - ig + 3) * s];
    }
  }
}

// ---- the "JB" boundary conditions (use_new_bcs): setAccelerationBCs4DJB / setAdvectionBCs4DJB
// (KineticSpeciesF.f:1301-1520, 1524-1733).  One launch per side of a direction d, one thread per boundary
// line: inflow (lower: face velocity > 0, upper: < 0) -> IC tables, else the binomial extrapolation of order
// e = min(interior extent, solution_order), accumulated from 0.0 in stencil order like the Fortran.
__constant__ double JB_ECOEFFS[6][6] = {{1.0, 0, 0, 0, 0, 0},        {2.0, -1.0, 0, 0, 0, 0},
                                        {3.0, -3.0, 1.0, 0, 0, 0},   {4.0, -6.0, 4.0, -1.0, 0, 0},
                                        {5.0, -10.0, 10.0, -5.0, 1.0, 0}, {6.0, -15.0, 20.0, -15.0, 6.0, -1.0}};
__global__ void k_bcs_jb(Geo g, lkstrict::DGeo dg, lkstrict::DAccel a, const double* __restrict__ vel, lk_inflow ic,
                         double* __restrict__ u, int d, int hi, int order) {
  const int ng = g.ng;
  // the three directions other than d, fastest first
  int od[3], k = 0;
  for (int q = 0; q < 4; ++q)
    if (q != d) od[k++] = q;
  const i64 total = (i64)g.nd[od[0]] * g.nd[od[1]] * g.nd[od[2]];
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int i[4];
  i[od[0]] = (int)(t % g.nd[od[0]]);
  const i64 r = t / g.nd[od[0]];
  i[od[1]] = (int)(r % g.nd[od[1]]);
  i[od[2]] = (int)(r / g.nd[od[1]]);
  const int na = ng, nb = ng + g.n[d] - 1;
  const int face = hi ? nb + 1 : na;
  double v;
  if (d == 0) v = vel[i[2] + (i64)g.nd[2] * i[3]];                              // vel1 = vx(i3,i4)
  else if (d == 1) v = vel[i[2] + (i64)g.nd[2] * (i[3] + (i64)g.nd[3])];        // vel2 = vy(i3,i4)
  else if (d == 2) v = lkstrict::accel_x(a, dg, i[0], i[1], face, i[3]);         // vel3(face,i4,i1,i2)
  else v = lkstrict::accel_y(a, dg, i[0], i[1], i[2], face);                      // vel4(face,i1,i2,i3)
  const bool inflow = hi ? (v < 0.0) : (v > 0.0);
  const int e = min(g.n[d], order);
  i[d] = 0;
  double* p = u + (i64)i[0] + g.s[1] * i[1] + g.s[2] * i[2] + g.s[3] * i[3];
  const i64 s = g.s[d];
  for (int ig = 1; ig <= ng; ++ig) {
    const int c = hi ? nb + ig : na - ig;
    if (inflow) {
      int q[4] = {i[0], i[1], i[2], i[3]};
      q[d] = c;
      p[c * s] = inflow_value(ic, g, q[0], q[1], q[2], q[3]);
    } else {
      double acc = 0.0;
      for (int m = 1; m <= e; ++m) acc = acc + JB_ECOEFFS[e - 1][m - 1] * p[(hi ? c - m : c + m) * s];
      p[c * s] = acc;
    }
  }
}

static Geo make_geo(const lk_geom* g) {
  Geo d;
  d.ng = g->ng;
  i64 s = 1;
  for (int k = 0; k < 4; ++k) {
    d.n[k] = g->n[k];
    d.nd[k] = g->n[k] + 2 * g->ng;
    d.s[k] = s;
    s *= d.nd[k];
  }
  return d;
}
// sides[8] = {x lo, x hi, y lo, y hi, vx lo, vx hi, vy lo, vy hi}: which sides to set, in the reference's order
cudaError_t set_bcs_jb(double* f, const lk_geom* g, const lk_accel* a, const double* velocities, const lk_inflow* ic,
                       const int sides[8], cudaStream_t st, int64_t* launches) {
  Geo d = make_geo(g);
  lkstrict::DGeo dg;
  lkstrict::DAccel da;
  memset(&dg, 0, sizeof(dg));
  memset(&da, 0, sizeof(da));
  for (int k = 0; k < 4; ++k) { dg.n[k] = d.n[k]; dg.nd[k] = d.nd[k]; dg.s[k] = d.s[k]; dg.dx[k] = g->dx[k]; }
  dg.ng = g->ng; dg.order = g->order;
  if (a) {
    da.kind = a->kind; da.field = a->field; da.vz = a->vz; da.vxf = a->vxface_velocities; da.vyf = a->vyface_velocities;
    da.norm = a->normalization; da.bz = a->bz_const;
  }
  lk_inflow di;
  if (ic) di = *ic;
  else {
    lk_inflow z = {};
    di = z;
  }
  for (int dir = 0; dir < 4; ++dir)
    for (int hi = 0; hi < 2; ++hi) {
      if (!sides[2 * dir + hi]) continue;
      i64 total = 1;
      for (int q = 0; q < 4; ++q)
        if (q != dir) total *= d.nd[q];
      k_bcs_jb<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(d, dg, da, velocities, di, f, dir, hi, g->order);
      ++*launches;
    }
  return cudaGetLastError();
}

// ---- appendkrook (KineticSpeciesF.f:2995-3034): rhs -= nu(x,y)/dt * (u - f0) where nu != 0, f0 from the IC tables
__global__ void k_append_krook(Geo g, const double* __restrict__ nu, double dt, lk_inflow ic, const double* __restrict__ u,
                               double* __restrict__ rhs) {
  const i64 total = (i64)g.n[0] * g.n[1] * g.n[2] * g.n[3];
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const int i1 = (int)(t % g.n[0]) + g.ng;
    i64 r = t / g.n[0];
    const int i2 = (int)(r % g.n[1]) + g.ng;
    r /= g.n[1];
    const int i3 = (int)(r % g.n[2]) + g.ng, i4 = (int)(r / g.n[2]) + g.ng;
    const double v = nu[i1 + (i64)g.nd[0] * i2];
    if (v != 0.0) {
      const i64 o = i1 + g.s[1] * i2 + g.s[2] * i3 + g.s[3] * i4;
      const double f0 = inflow_value(ic, g, i1, i2, i3, i4);
      rhs[o] = rhs[o] - v / dt * (u[o] - f0);
    }
  }
}
cudaError_t append_krook(double* rhs, const double* u, const lk_geom* g, const double* nu, double dt, const lk_inflow* ic,
                         cudaStream_t st, int64_t* launches);

cudaError_t set_advection_bcs(double* f, const lk_geom* g, const double* velocities, const lk_inflow* ic, const int at[4],
                              int periodic_x, int periodic_y, cudaStream_t st, int64_t* launches) {
  Geo d;
  d.ng = g->ng;
  i64 s = 1;
  for (int k = 0; k < 4; ++k) {
    d.n[k] = g->n[k];
    d.nd[k] = g->n[k] + 2 * g->ng;
    d.s[k] = s;
    s *= d.nd[k];
  }
  lk_inflow di;
  if (ic) di = *ic;
  else {
    lk_inflow z = {};
    di = z;
  }
  if (!periodic_x && (at[0] || at[1])) {
    const i64 total = (i64)d.nd[1] * d.nd[2] * d.nd[3];
    k_advection_bcs<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(d, velocities, di, f, 0, at[0], at[1]);
    ++*launches;
  }
  if (!periodic_y && (at[2] || at[3])) {
    const i64 total = (i64)d.nd[0] * d.nd[2] * d.nd[3];
    k_advection_bcs<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(d, velocities, di, f, 1, at[2], at[3]);
    ++*launches;
  }
  return cudaGetLastError();
}

cudaError_t append_krook(double* rhs, const double* u, const lk_geom* g, const double* nu, double dt, const lk_inflow* ic,
                         cudaStream_t st, int64_t* launches) {
  Geo d = make_geo(g);
  lk_inflow di;
  if (ic) di = *ic;
  else {
    lk_inflow z = {};
    di = z;
  }
  const i64 total = (i64)d.n[0] * d.n[1] * d.n[2] * d.n[3];
  i64 blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_append_krook<<<(unsigned)blocks, 256, 0, st>>>(d, nu, dt, di, u, rhs);
  ++*launches;
  return cudaGetLastError();
}

// ---- Twilight-zone sources (TZSourceF.f, ElectronTZSourceF.f, TwoSpecies_ElectronTZSourceF.f, TwoSpecies_IonTZSourceF.f):
// the forcing h(x, y, vx, vy, t) added to a right-hand side over the whole data box, and error = soln - f_exact.  The
// transcendental factors are separable -- sines / cosines of x (per i1) and of y (per i2) at the electron and the ion wave
// numbers, exp(-alpha^2 v^2 / 2) (per (i3, i4)), sines / cosines of t (scalars) -- and come from host tables built with libm,
// so the kernel evaluates Maple's expression trees in the Fortran's parse order on the same operand bits (-fmad=false).
// kind 0 TrigTZSource (kx = ky = 1), 1 ElectronTrigTZSource (kx = ky = 4), 2 TwoSpecies_ElectronTrigTZSource,
// 3 TwoSpecies_IonTrigTZSource (kxE = 4, kyE = 2, kxI = 2, kyI = 4; alpha = sqrt(mass)).
// tab: {sin(kxE x), cos(kxE x), sin(kxI x), cos(kxI x)}[n1d] {sin(kyE y), cos(kyE y), sin(kyI y), cos(kyI y)}[n2d] exp(..)[n3d n4d];
// kinds 0 / 1 use the "E" tables for their single wave number
struct TzScalars {
  double A, mass, alpha, ste, cte, sti, cti, pi;
  int kind;
};
__device__ __forceinline__ double tz_powi2(double x) { return x * x; }
__device__ __forceinline__ double tz_powi4(double x) { const double t = x * x; return t * t; }   // gfortran's x**4
__global__ void k_trig_tz(Geo g, const double* __restrict__ tab, const double* __restrict__ velocities, TzScalars q,
                          const double* __restrict__ soln, double* __restrict__ out) {
  const int kind = q.kind;
  const double* sxe_t = tab;
  const double* cxe_t = sxe_t + g.nd[0];
  const double* sxi_t = cxe_t + g.nd[0];
  const double* cxi_t = sxi_t + g.nd[0];
  const double* sye_t = cxi_t + g.nd[0];
  const double* cye_t = sye_t + g.nd[1];
  const double* syi_t = cye_t + g.nd[1];
  const double* cyi_t = syi_t + g.nd[1];
  const double* ev = cyi_t + g.nd[1];
  const double a = q.A, pi = q.pi;
  const i64 total = (i64)g.nd[0] * g.nd[1] * g.nd[2] * g.nd[3];
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const int i1 = (int)(t % g.nd[0]);
    i64 r = t / g.nd[0];
    const int i2 = (int)(r % g.nd[1]);
    r /= g.nd[1];
    const int i3 = (int)(r % g.nd[2]), i4 = (int)(r / g.nd[2]);
    const i64 pv = i3 + (i64)g.nd[2] * i4;
    const double vx = velocities[pv], vy = velocities[pv + (i64)g.nd[2] * g.nd[3]];
    const double e = ev[pv];
    if (kind < 2) {
      const double kx = kind ? 4.0 : 1.0, ky = kx, kt = 1.0, alpha = 1.0;
      const double sinx = sxe_t[i1], cosx = cxe_t[i1], siny = sye_t[i2], cosy = cye_t[i2], st = q.ste, ct = q.cte;
      if (soln) {
        const double fexact = alpha / pi * e * (0.1e1 + a * cosx * cosy * st) / 0.2e1;
        out[t] = soln[t] - fexact;
      } else if (kind) {
        const double h =
            alpha / pi * e * a * cosx * cosy * kt * ct / 0.2e1 - vx * alpha / pi * e * a * kx * sinx * cosy * st / 0.2e1 -
            vy * alpha / pi * e * a * cosx * ky * siny * st / 0.2e1 -
            a * kx * sinx * cosy * st / (kx * kx + ky * ky) * (alpha * alpha) / pi * vx * e * (0.1e1 + a * cosx * cosy * st) / 0.2e1 -
            a * cosx * ky * siny * st / (kx * kx + ky * ky) * (alpha * alpha) / pi * vy * e * (0.1e1 + a * cosx * cosy * st) / 0.2e1;
        out[t] = out[t] + h;
      } else {
        const double h =
            -0.1e1 / (kx * kx + ky * ky) * a * sinx * kx * cosy * st * (alpha * alpha) / pi * vx * e * (0.1e1 + a * cosx * cosy * st) / 0.2e1 -
            0.1e1 / (kx * kx + ky * ky) * a * cosx * siny * ky * st * (alpha * alpha) / pi * vy * e * (0.1e1 + a * cosx * cosy * st) / 0.2e1 -
            alpha / pi * e * a * sinx * kx * cosy * st * vx / 0.2e1 - alpha / pi * e * a * cosx * siny * ky * st * vy / 0.2e1 +
            alpha / pi * e * a * cosx * cosy * ct * kt / 0.2e1;
        out[t] = out[t] + h;
      }
      continue;
    }
    // the two-species sources (TwoSpecies_ElectronTZSourceF.f:134-152, :245-247; TwoSpecies_IonTZSourceF.f:112-129, :221-223)
    const double kxi = 0.2e1, kyi = 0.4e1, kti = 0.1e1, kxe = 0.4e1, kye = 0.2e1, kte = 0.1e1;
    const double kI2 = tz_powi2(kxi) + tz_powi2(kyi), kE2 = tz_powi2(kxe) + tz_powi2(kye);
    const double sxe = sxe_t[i1], cxe = cxe_t[i1], sxi = sxi_t[i1], cxi = cxi_t[i1];
    const double sye = sye_t[i2], cye = cye_t[i2], syi = syi_t[i2], cyi = cyi_t[i2];
    const double ste = q.ste, cte = q.cte, sti = q.sti, cti = q.cti, m = q.mass;
    const double al2 = tz_powi2(q.alpha), al4 = tz_powi4(q.alpha);
    if (kind == 2) {
      const double wave = 0.1e1 + (((a * cxe) * cye) * ste);
      if (soln) {
        out[t] = soln[t] - ((((al2 / pi) * e) * wave) / 0.2e1);
        continue;
      }
      const double T1 = (((((((al2 / pi) * e) * a) * cxe) * cye) * kte) * cte) / 0.2e1;
      const double T2 = ((((((((vx * al2) / pi) * e) * a) * kxe) * sxe) * cye) * ste) / 0.2e1;
      const double T3 = ((((((((vy * al2) / pi) * e) * a) * cxe) * kye) * sye) * ste) / 0.2e1;
      const double T4 = ((((((((((0.1e1 / m) * (((((ste * sxe) * kxe) * kI2) * cye) - (((((0.2e1 * cyi) * sti) * sxi) * kxi) * kE2))) * a) / kI2) / kE2) * al4) / pi) * vx) * e) * wave) / 0.2e1;
      const double T5 = ((((((((((0.1e1 / m) * a) * (((((ste * sye) * kye) * kI2) * cxe) - (((((0.2e1 * kyi) * cxi) * sti) * syi) * kE2))) / kI2) / kE2) * al4) / pi) * vy) * e) * wave) / 0.2e1;
      out[t] = out[t] + ((((T1 - T2) - T3) - T4) - T5);
    } else {
      const double wave = 0.1e1 + ((((0.2e1 * a) * cxi) * cyi) * sti);
      if (soln) {
        out[t] = soln[t] - ((((al2 / pi) * e) * wave) / 0.2e1);
        continue;
      }
      const double U1 = ((((((al2 / pi) * e) * a) * cxi) * cyi) * kti) * cti;
      const double U2 = (((((((vx * al2) / pi) * e) * a) * kxi) * sxi) * cyi) * sti;
      const double U3 = (((((((vy * al2) / pi) * e) * a) * cxi) * kyi) * syi) * sti;
      const double U4 = ((((((((((0.1e1 / m) * (((((ste * sxe) * kxe) * kI2) * cye) - (((((0.2e1 * cyi) * sti) * sxi) * kxi) * kE2))) * a) / kI2) / kE2) * al4) / pi) * vx) * e) * wave) / 0.2e1;
      const double U5 = ((((((((((0.1e1 / m) * a) * (((((ste * sye) * kye) * kI2) * cxe) - (((((0.2e1 * kyi) * cxi) * sti) * syi) * kE2))) / kI2) / kE2) * al4) / pi) * vy) * e) * wave) / 0.2e1;
      out[t] = out[t] + ((((U1 - U2) - U3) + U4) + U5);
    }
  }
}

// params = {amp, electron_mass, ion_mass} (dparams of the Fortran; the masses only matter for kinds 2 / 3)
static double tz_alpha(int kind, const double* params) {
  return kind == 2 ? sqrt(params[1]) : kind == 3 ? sqrt(params[2]) : 1.0;
}
// host side of the tables: libm, the Fortran's argument expressions; vel_host: (n3d, n4d, 2)
void trig_tz_tables(double* tab, const lk_geom* g, const int lo[2], const double xlo[2], const double* vel_host, int kind,
                    const double* params) {
  const double kxe = kind >= 2 ? 0.4e1 : (kind ? 4.0 : 1.0), kye = kind >= 2 ? 0.2e1 : kxe, kxi = 0.2e1, kyi = 0.4e1;
  const int n1d = g->n[0] + 2 * g->ng, n2d = g->n[1] + 2 * g->ng, n3d = g->n[2] + 2 * g->ng, n4d = g->n[3] + 2 * g->ng;
  double* xs = tab;
  double* ys = tab + 4 * (size_t)n1d;
  double* ev = ys + 4 * (size_t)n2d;
  for (int i1 = 0; i1 < n1d; ++i1) {
    const double x = xlo[0] + ((lo[0] + i1) + 0.5) * g->dx[0];
    xs[i1] = sin(kxe * x);
    xs[n1d + i1] = cos(kxe * x);
    xs[2 * n1d + i1] = kind >= 2 ? sin(kxi * x) : 0.0;
    xs[3 * n1d + i1] = kind >= 2 ? cos(kxi * x) : 0.0;
  }
  for (int i2 = 0; i2 < n2d; ++i2) {
    const double y = xlo[1] + ((lo[1] + i2) + 0.5) * g->dx[1];
    ys[i2] = sin(kye * y);
    ys[n2d + i2] = cos(kye * y);
    ys[2 * n2d + i2] = kind >= 2 ? sin(kyi * y) : 0.0;
    ys[3 * n2d + i2] = kind >= 2 ? cos(kyi * y) : 0.0;
  }
  const double alpha = tz_alpha(kind, params);
  for (i64 pv = 0; pv < (i64)n3d * n4d; ++pv) {
    const double vx = vel_host[pv], vy = vel_host[pv + (i64)n3d * n4d];
    // kinds 0 / 1: exp(-alpha * v^2 / 2) with alpha = 1; kinds 2 / 3: exp(-alpha**2 * v^2 / 2)
    ev[pv] = kind >= 2 ? exp(-((alpha * alpha) * (vx * vx + vy * vy) / 0.2e1)) : exp(-(alpha * (vx * vx + vy * vy) / 0.2e1));
  }
}
size_t trig_tz_table_count(const lk_geom* g) {
  const size_t n1d = g->n[0] + 2 * g->ng, n2d = g->n[1] + 2 * g->ng, n3d = g->n[2] + 2 * g->ng, n4d = g->n[3] + 2 * g->ng;
  return 4 * n1d + 4 * n2d + n3d * n4d;
}

// out += h (soln == nullptr) or out = soln - f_exact; tab_dev: trig_tz_tables on the device
cudaError_t trig_tz(double* out, const double* soln, const lk_geom* g, const double* tab_dev, const double* velocities,
                    double time, int kind, const double* params, cudaStream_t st, int64_t* launches) {
  Geo d = make_geo(g);
  TzScalars q;
  q.kind = kind;
  q.A = params[0];
  q.mass = kind == 2 ? params[1] : kind == 3 ? params[2] : 1.0;
  q.alpha = tz_alpha(kind, params);
  q.pi = 4.0 * atan(1.0);
  const double kte = 1.0, kti = 1.0;
  q.ste = sin(kte * time);
  q.cte = cos(kte * time);
  q.sti = sin(kti * time);
  q.cti = cos(kti * time);
  const i64 total = (i64)d.nd[0] * d.nd[1] * d.nd[2] * d.nd[3];
  i64 blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_trig_tz<<<(unsigned)blocks, 256, 0, st>>>(d, tab_dev, velocities, q, soln, out);
  ++*launches;
  return cudaGetLastError();
}

// ---- MaxwellF.f boundary routines (zeroghost2d :10-58, maxwelladdantennasource :359-389, maxwellsetembcs :473-657,
// maxwellsetvzbcs :661-731).  Arrays (n1d, n2d, ncomp); the ghost layers are filled outward one after the other, each
// from the three cells inside it, so a thread owns one boundary line and walks its ghosts in the Fortran's order.
#define E3(a, i1, i2, c) (a)[(i1) + (i64)n1d * ((i2) + (i64)n2d * (c))]
__global__ void k_zero_ghost_2d(double* __restrict__ u, int n1, int n2, int ng, int dim) {
  const int n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  const i64 total = (i64)n1d * n2d * dim;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const int i1 = (int)(t % n1d), i2 = (int)((t / n1d) % n2d);
    if (i1 < ng || i1 >= ng + n1 || i2 < ng || i2 >= ng + n2) u[t] = 0.0;
  }
}
__global__ void k_antenna_source(double* __restrict__ dem, const double* __restrict__ src, int n1, int n2, int ng) {
  const int n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  const i64 total = (i64)n1 * n2 * 6;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const int i1 = (int)(t % n1) + ng, i2 = (int)((t / n1) % n2) + ng, c = (int)(t / ((i64)n1 * n2));
    E3(dem, i1, i2, c) = E3(dem, i1, i2, c) - E3(src, i1, i2, c);
  }
}
// the outgoing characteristic of the pair (s1 * a, b) is kept, the incoming one zeroed (MaxwellF.f:519-541)
__device__ __forceinline__ void em_characteristic(double* pa, double* pb, double s1, double c, int high) {
  double u1 = s1 * *pa, u2 = *pb;
  double w1 = +u1 / (2. * c) + u2 / 2.;
  double w2 = -u1 / (2. * c) + u2 / 2.;
  if (high) w2 = 0.0; else w1 = 0.0;
  u1 = c * (w1 - w2);
  u2 = w1 + w2;
  *pa = s1 * u1;
  *pb = u2;
}
// dir 0: x edges (one thread per interior row i2), dir 1: y edges (one thread per interior column i1)
__global__ void k_em_bcs(double* __restrict__ em, int n1, int n2, int ng, int dir, int at_lo, int at_hi, double c) {
  const int n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  const int lines = dir == 0 ? n2 : n1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * lines) return;
  const int high = t / lines, line = t % lines + ng;
  if (high ? !at_hi : !at_lo) return;
  const int d = high ? 1 : -1;
  if (dir == 0) {
    const int i1 = high ? ng + n1 - 1 : ng, i2 = line;
    for (int i4 = 1; i4 <= ng; ++i4) {
      const int ig = i1 + d * i4;
      for (int k = 0; k < 6; ++k)
        E3(em, ig, i2, k) = +3.0 * E3(em, ig - d, i2, k) - 3.0 * E3(em, ig - 2 * d, i2, k) + 1.0 * E3(em, ig - 3 * d, i2, k);
      em_characteristic(&E3(em, ig, i2, 1), &E3(em, ig, i2, 5), 1.0, c, high);   // Ey, Bz
      em_characteristic(&E3(em, ig, i2, 2), &E3(em, ig, i2, 4), -1.0, c, high);  // -Ez, By
    }
  } else {
    const int i2 = high ? ng + n2 - 1 : ng, i1 = line;
    for (int i4 = 1; i4 <= ng; ++i4) {
      const int ig = i2 + d * i4;
      for (int k = 0; k < 6; ++k)
        E3(em, i1, ig, k) = +3.0 * E3(em, i1, ig - d, k) - 3.0 * E3(em, i1, ig - 2 * d, k) + 1.0 * E3(em, i1, ig - 3 * d, k);
      em_characteristic(&E3(em, i1, ig, 0), &E3(em, i1, ig, 5), -1.0, c, high);  // -Ex, Bz
      em_characteristic(&E3(em, i1, ig, 2), &E3(em, i1, ig, 3), 1.0, c, high);   // Ez, Bx
    }
  }
}
// even reflection about the boundary cell; dir 0 over every row of the data box, dir 1 over every column
__global__ void k_vz_bcs(double* __restrict__ vz, int n1, int n2, int ng, int dir, int at_lo, int at_hi) {
  const int n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  const int lines = dir == 0 ? n2d : n1d;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * lines) return;
  const int high = t / lines, line = t % lines;
  if (high ? !at_hi : !at_lo) return;
  for (int i3 = 1; i3 <= ng; ++i3) {
    if (dir == 0) {
      const int i1 = high ? ng + n1 - 1 : ng;
      E3(vz, high ? i1 + i3 : i1 - i3, line, 0) = E3(vz, high ? i1 - i3 : i1 + i3, line, 0);
    } else {
      const int i2 = high ? ng + n2 - 1 : ng;
      E3(vz, line, high ? i2 + i3 : i2 - i3, 0) = E3(vz, line, high ? i2 - i3 : i2 + i3, 0);
    }
  }
}
#undef E3
cudaError_t zero_ghost_2d(double* u, int n1, int n2, int ng, int dim, cudaStream_t st, int64_t* launches) {
  const i64 total = (i64)(n1 + 2 * ng) * (n2 + 2 * ng) * dim;
  k_zero_ghost_2d<<<(unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256), 256, 0, st>>>(u, n1, n2, ng, dim);
  ++*launches;
  return cudaGetLastError();
}
cudaError_t antenna_source(double* dem, const double* src, int n1, int n2, int ng, cudaStream_t st, int64_t* launches) {
  const i64 total = (i64)n1 * n2 * 6;
  k_antenna_source<<<(unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256), 256, 0, st>>>(dem, src, n1, n2, ng);
  ++*launches;
  return cudaGetLastError();
}
cudaError_t em_bcs(double* em, int n1, int n2, int ng, const int at[4], int x_periodic, int y_periodic, double c, cudaStream_t st,
                   int64_t* launches) {
  if (x_periodic == 0 && (at[0] || at[1])) {
    k_em_bcs<<<(2 * n2 + 127) / 128, 128, 0, st>>>(em, n1, n2, ng, 0, at[0], at[1], c);
    ++*launches;
  }
  if (y_periodic == 0 && (at[2] || at[3])) {
    k_em_bcs<<<(2 * n1 + 127) / 128, 128, 0, st>>>(em, n1, n2, ng, 1, at[2], at[3], c);
    ++*launches;
  }
  return cudaGetLastError();
}
cudaError_t vz_bcs(double* vz, int n1, int n2, int ng, const int at[4], int x_periodic, int y_periodic, cudaStream_t st,
                   int64_t* launches) {
  if (x_periodic == 0 && (at[0] || at[1])) {
    k_vz_bcs<<<(2 * (n2 + 2 * ng) + 127) / 128, 128, 0, st>>>(vz, n1, n2, ng, 0, at[0], at[1]);
    ++*launches;
  }
  if (y_periodic == 0 && (at[2] || at[3])) {
    k_vz_bcs<<<(2 * (n1 + 2 * ng) + 127) / 128, 128, 0, st>>>(vz, n1, n2, ng, 1, at[2], at[3]);
    ++*launches;
  }
  return cudaGetLastError();
}

}  // namespace lkbcs
