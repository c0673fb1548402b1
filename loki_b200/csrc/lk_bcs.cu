// lk_bcs.cu -- setAdvectionBCs4D (KineticSpeciesF.f:1166-1297): the physical x / y boundaries of a
// configuration-space direction that is NOT periodic (all five benchmark decks are periodic,
// where the reference skips this routine's body and only wraps).  Same rule as the velocity boundaries:
// outflow by the sign of the face velocity at the boundary face -> u_g = 3u_-1 - 3u_-2 + u_-3 marching
// outward, inflow -> the initial condition (the IC classes' cached tables, lk_inflow).  x boundaries first
// over the full data box of (y, vx, vy), then y boundaries over the full x extent, ghosts just set included.
// Compiled with -fmad=false: the extrapolation is the reference's expression, bit for bit in both modes.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/loki_b200.h"

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

}  // namespace lkbcs
