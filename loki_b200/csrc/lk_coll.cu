// lk_coll.cu -- kernels of the pitch-angle collision operator (completeRHS, KineticSpecies.C:1036-1046):
// the per-cell arithmetic is lk_coll.cuh, this file maps threads to cells.
//
//   k_fields / k_moments / k_kec   one thread per configuration-space point (i1 fastest: every velocity-plane read of a
//                                  warp is one coalesced row segment), walking the interior velocity cells in the
//                                  reference's order, so the sums carry the reference's bits.  HBM-bound: f is read
//                                  twice (the thermal speed needs the flow first), 16 B per cell.
//   k_append<ORDER>                one thread per interior cell, i1 fastest; the (2 ng + 1)^2 window is read straight
//                                  from global memory: the velocity neighbours of a warp's row segment are row segments
//                                  again, so every load is coalesced and the window's re-reads are served by L1 / L2.
//
// Compiled with -fmad=false: no contraction, the oracle's operation order.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/loki_b200.h"
#include "lk_coll.cuh"

namespace lkcoll {

struct Geo {
  int n[4], nd[4], ng;
  i64 pl, pv;
};

static Geo make_geo(const lk_geom* g) {
  Geo d;
  d.ng = g->ng;
  for (int k = 0; k < 4; ++k) {
    d.n[k] = g->n[k];
    d.nd[k] = g->n[k] + 2 * g->ng;
  }
  d.pl = (i64)d.nd[0] * d.nd[1];
  d.pv = (i64)d.nd[2] * d.nd[3];
  return d;
}

__global__ void k_fields(Geo g, const double* __restrict__ u, const double* __restrict__ vel, double measure,
                         double* __restrict__ ivx, double* __restrict__ ivy, double* __restrict__ vth) {
  const i64 c2 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c2 >= g.pl) return;
  double a, b, c;
  fields_point(u + c2, g.pl, g.ng, g.n[2], g.n[3], g.nd[2], vel, g.pv, measure, a, b, c);
  ivx[c2] = a;
  ivy[c2] = b;
  vth[c2] = c;
}

// the Fortran-ABI pieces: raw sums (the reference zeroes rN / rGammax / rGammay inside the routine, and rKEC outside)
__global__ void k_moments(Geo g, const double* __restrict__ u, const double* __restrict__ vel, double* __restrict__ rn,
                          double* __restrict__ rgx, double* __restrict__ rgy) {
  const i64 c2 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c2 >= g.pl) return;
  double a = 0.0, b = 0.0, c = 0.0;
  moments_point(u + c2, g.pl, g.ng, g.n[2], g.n[3], g.nd[2], vel, g.pv, a, b, c);
  rn[c2] = a;
  rgx[c2] = b;
  rgy[c2] = c;
}
__global__ void k_kec(Geo g, const double* __restrict__ u, const double* __restrict__ vel, const double* __restrict__ vx0,
                      const double* __restrict__ vy0, double* __restrict__ rk) {
  const i64 c2 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c2 >= g.pl) return;
  double a = rk[c2];
  kec_point(u + c2, g.pl, g.ng, g.n[2], g.n[3], g.nd[2], vel, g.pv, vx0[c2], vy0[c2], a);
  rk[c2] = a;
}
// mode 0: vx = gx / n, vy = gy / n (computePitchAngleSpeciesReducedFields); mode 1: vx = sqrt(0.5 gx / n) (...Vthermal)
__global__ void k_reduced(i64 pl, int mode, double* __restrict__ vx, double* __restrict__ vy, const double* __restrict__ n,
                          const double* __restrict__ gx, const double* __restrict__ gy) {
  const i64 c2 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c2 >= pl) return;
  if (mode == 0) {
    vx[c2] = gx[c2] / n[c2];
    vy[c2] = gy[c2] / n[c2];
  } else {
    vx[c2] = sqrt(0.5 * gx[c2] / n[c2]);
  }
}

template <int ORDER>
__global__ void __launch_bounds__(128)
k_append(Geo g, Params P, int conservative, const double* __restrict__ f, const double* __restrict__ vel,
         const double* __restrict__ ivx, const double* __restrict__ ivy, const double* __restrict__ vth,
         double* __restrict__ rhs) {
  const i64 total = (i64)g.n[0] * g.n[1] * g.n[2] * g.n[3];
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const int i1 = (int)(t % g.n[0]) + g.ng;
    i64 r = t / g.n[0];
    const int i2 = (int)(r % g.n[1]) + g.ng;
    r /= g.n[1];
    const int i3 = (int)(r % g.n[2]) + g.ng, i4 = (int)(r / g.n[2]) + g.ng;
    const i64 c2 = i1 + (i64)g.nd[0] * i2;
    const double cf = collision_cell<ORDER>(f, g.pl, g.nd[2], g.nd[3], c2, i3, i4, vel, ivx, ivy, vth, P, conservative);
    const i64 c = c2 + g.pl * (i3 + (i64)g.nd[2] * i4);
    rhs[c] = rhs[c] + cf;
  }
}

static unsigned blocks_for(i64 total, int threads) { return (unsigned)((total + threads - 1) / threads); }

cudaError_t fields(double* ivx, double* ivy, double* vth, const double* u, const lk_geom* g, const double* velocities,
                   cudaStream_t st, int64_t* launches) {
  Geo d = make_geo(g);
  k_fields<<<blocks_for(d.pl, 64), 64, 0, st>>>(d, u, velocities, g->dx[2] * g->dx[3], ivx, ivy, vth);
  ++*launches;
  return cudaGetLastError();
}
cudaError_t moments(double* rn, double* rgx, double* rgy, const double* u, const lk_geom* g, const double* velocities,
                    cudaStream_t st, int64_t* launches) {
  Geo d = make_geo(g);
  k_moments<<<blocks_for(d.pl, 64), 64, 0, st>>>(d, u, velocities, rn, rgx, rgy);
  ++*launches;
  return cudaGetLastError();
}
cudaError_t kec(double* rk, const double* vx0, const double* vy0, const double* u, const lk_geom* g,
                const double* velocities, cudaStream_t st, int64_t* launches) {
  Geo d = make_geo(g);
  k_kec<<<blocks_for(d.pl, 64), 64, 0, st>>>(d, u, velocities, vx0, vy0, rk);
  ++*launches;
  return cudaGetLastError();
}
cudaError_t reduced(int mode, double* vx, double* vy, const double* n, const double* gx, const double* gy, int64_t pl,
                    cudaStream_t st, int64_t* launches) {
  k_reduced<<<blocks_for(pl, 128), 128, 0, st>>>(pl, mode, vx, vy, n, gx, gy);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t append(double* rhs, const double* f, const lk_geom* g, const double* velocities, const double* ivx,
                   const double* ivy, const double* vth, const double vlo[2], const double vhi[2],
                   const lk_pitch_angle* pa, cudaStream_t st, int64_t* launches) {
  if (!pa->conservative && g->order != 4) return cudaSuccess;  // appendPitchAngleCollision :1686-1699 applies nothing
  Geo d = make_geo(g);
  Params P;
  const int vrolloff = (g->order == 4) ? 3 : 4;  // :285, :949
  for (int k = 0; k < 2; ++k) {
    P.range_lo[k] = pa->range_lo[k];
    P.range_hi[k] = pa->range_hi[k];
    P.vmin[k] = vlo[k] + vrolloff * g->dx[2 + k];
    P.vmax[k] = vhi[k] - vrolloff * g->dx[2 + k];
  }
  P.vfloor = pa->vfloor;
  P.nu_coef = pa->nu_coef;
  P.dvx = g->dx[2];
  P.dvy = g->dx[3];
  const i64 total = (i64)d.n[0] * d.n[1] * d.n[2] * d.n[3];
  i64 blocks = (total + 127) / 128;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (g->order == 4)
    k_append<4><<<(unsigned)blocks, 128, 0, st>>>(d, P, pa->conservative, f, velocities, ivx, ivy, vth, rhs);
  else
    k_append<6><<<(unsigned)blocks, 128, 0, st>>>(d, P, pa->conservative, f, velocities, ivx, ivy, vth, rhs);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace lkcoll
