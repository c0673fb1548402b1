// lk_kernels.cu -- sm_100a kernels of the Vlasov RHS path (everything except the fused marching
// stage kernel, which lives in lk_march.cuh).  Compiled twice: production (lkfast) and strict
// (lkstrict, -fmad=false); see lk_device.cuh.
#include "lk_device.cuh"
#include "lk_launch.h"
#include "lk_march.cuh"
#include "lk_pipe.cuh"

namespace LK_NS {

static int64_t g_launches = 0;
int64_t launches() { return g_launches; }
static int64_t g_pipe_launches = 0;
int64_t pipe_launches() { return g_pipe_launches; }
#define LK_LAUNCHED() (++g_launches, cudaGetLastError())

static DGeo make_geo(const lk_geom* g) {
  DGeo d;
  for (int k = 0; k < 4; ++k) {
    d.n[k] = g->n[k];
    d.nd[k] = g->n[k] + 2 * g->ng;
    d.dx[k] = g->dx[k];
  }
  d.ng = g->ng;
  d.order = g->order;
  d.s[0] = 1;
  d.s[1] = d.nd[0];
  d.s[2] = (i64)d.nd[0] * d.nd[1];
  d.s[3] = (i64)d.nd[0] * d.nd[1] * d.nd[2];
  return d;
}
static DAccel make_accel(const lk_accel* a) {
  DAccel d;
  d.kind = a->kind;
  d.field = a->field;
  d.vz = a->vz;
  d.vxf = a->vxface_velocities;
  d.vyf = a->vyface_velocities;
  d.norm = a->normalization;
  d.bz = a->bz_const;
  return d;
}
static DUpd make_upd(const lk_rk_update* u) {
  DUpd d;
  memset(&d, 0, sizeof(d));
  if (u) {
    d.f_old = u->f_old;
    d.delta_in = u->delta_in;
    d.delta_out = u->delta_out;
    d.pred = u->pred;
    d.w_delta = u->w_delta;
    d.c_pred = u->c_pred;
    d.use_delta = u->use_delta;
    d.active = 1;
    d.n_prev = u->n_prev;
    d.wrap = u->wrap;
    for (int j = 0; j < 7; ++j) {
      d.k_prev[j] = u->k_prev[j];
      d.c_prev[j] = u->c_prev[j];
    }
    if (u->krook_nu && u->krook_ic) {
      const lk_inflow* ic = u->krook_ic;
      d.krook_nu = u->krook_nu;
      d.krook_dt = u->krook_dt;
      d.krook_ic.kind = ic->kind; d.krook_ic.fx = ic->fx; d.krook_ic.fv = ic->fv; d.krook_ic.fx2 = ic->fx2; d.krook_ic.fv2 = ic->fv2;
      d.krook_ic.ghost3 = ic->ghost3; d.krook_ic.ghost4 = ic->ghost4; d.krook_ic.fnorm = ic->fnorm; d.krook_ic.frac = ic->frac;
    }
  }
  return d;
}
static inline unsigned nblk(i64 n, int t) { return (unsigned)((n + t - 1) / t); }

// ---------------------------------------------------------------------------------------------
// a1/a2 test hook
// ---------------------------------------------------------------------------------------------
__global__ void k_weno_fit(int order, const double* __restrict__ u, const double* __restrict__ vel,
                           double* __restrict__ face, i64 count) {
  i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  if (order == 4) {
    const double* p = u + 4 * k;
    face[k] = weno43(p[0], p[1], p[2], p[3], vel[k] > 0.0);
  } else {
    const double* p = u + 6 * k;
    face[k] = weno65(p[0], p[1], p[2], p[3], p[4], p[5], vel[k] > 0.0);
  }
}
cudaError_t weno_fit(int order, const double* u, const double* vel, double* face, int64_t count, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  k_weno_fit<<<nblk(count, 256), 256, 0, st>>>(order, u, vel, face, count);
  return LK_LAUNCHED();
}

// ---------------------------------------------------------------------------------------------
// a10 xpby4d: interior only (KineticSpeciesF.f:27-35).  One thread per 2 x-cells of a row.
// ---------------------------------------------------------------------------------------------
__global__ void k_xpby4d(DGeo g, double* __restrict__ x, const double* __restrict__ y, double b) {
  const i64 rows = (i64)g.n[1] * g.n[2] * g.n[3];
  const i64 total = rows * g.n[0];
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    int i1 = (int)(t % g.n[0]);
    i64 r = t / g.n[0];
    int i2 = (int)(r % g.n[1]);
    r /= g.n[1];
    int i3 = (int)(r % g.n[2]);
    int i4 = (int)(r / g.n[2]);
    i64 idx = gidx(g, i1 + g.ng, i2 + g.ng, i3 + g.ng, i4 + g.ng);
    x[idx] = x[idx] + b * y[idx];
  }
}
cudaError_t xpby4d(double* x, const double* y, double b, const lk_geom* g, cudaStream_t st) {
  DGeo d = make_geo(g);
  i64 total = (i64)g->n[0] * g->n[1] * g->n[2] * g->n[3];
  if (total <= 0) return cudaSuccess;
  unsigned blocks = (unsigned)min((i64)nblk(total, 256), (i64)148 * 32);
  k_xpby4d<<<blocks, 256, 0, st>>>(d, x, y, b);
  return LK_LAUNCHED();
}

// ---------------------------------------------------------------------------------------------
// a4/a5: max |a| over interior faces, and (optionally) the materialised rotated vel3/vel4 arrays.
// max is order independent, so a tree + atomicMax on the bit pattern of non-negative doubles is exact.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__global__ void k_zero2(double* out) { out[0] = 0.0; out[1] = 0.0; }

// thread per (i1, i2, i3face, i4) over the data box (+1 face); writes vel3 if non-null
__global__ void k_vel3(DGeo g, DAccel a, double* __restrict__ vel3, double* out2) {
  const i64 total = (i64)(g.nd[2] + 1) * g.nd[3] * g.nd[0] * g.nd[1];
  double m = 0.0;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    // vel3 layout (i3,i4,i1,i2): i3 fastest
    int i3 = (int)(t % (g.nd[2] + 1));
    i64 r = t / (g.nd[2] + 1);
    int i4 = (int)(r % g.nd[3]);
    r /= g.nd[3];
    int i1 = (int)(r % g.nd[0]);
    int i2 = (int)(r / g.nd[0]);
    double v = accel_x(a, g, i1, i2, i3, i4);
    if (vel3) vel3[t] = v;
    if (i1 >= g.ng && i1 < g.ng + g.n[0] && i2 >= g.ng && i2 < g.ng + g.n[1] && i3 >= g.ng &&
        i3 <= g.ng + g.n[2] && i4 >= g.ng && i4 < g.ng + g.n[3])
      m = fmax(m, fabs(v));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(out2 + 0, m);
}
__global__ void k_vel4(DGeo g, DAccel a, double* __restrict__ vel4, double* out2) {
  const i64 total = (i64)(g.nd[3] + 1) * g.nd[0] * g.nd[1] * g.nd[2];
  double m = 0.0;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    // vel4 layout (i4,i1,i2,i3): i4 fastest
    int i4 = (int)(t % (g.nd[3] + 1));
    i64 r = t / (g.nd[3] + 1);
    int i1 = (int)(r % g.nd[0]);
    r /= g.nd[0];
    int i2 = (int)(r % g.nd[1]);
    int i3 = (int)(r / g.nd[1]);
    double v = accel_y(a, g, i1, i2, i3, i4);
    if (vel4) vel4[t] = v;
    if (i1 >= g.ng && i1 < g.ng + g.n[0] && i2 >= g.ng && i2 < g.ng + g.n[1] && i3 >= g.ng &&
        i3 < g.ng + g.n[2] && i4 >= g.ng && i4 <= g.ng + g.n[3])
      m = fmax(m, fabs(v));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(out2 + 1, m);
}
// max only.  The acceleration is a monotone (affine) function of the one velocity-table value it reads
// (setphasespacevel4D, KineticSpeciesF.f:78-80, 98-100; Maxwell :154-158, 176-180), so its largest
// magnitude over the interior faces is attained at the smallest or the largest table value: reduce the
// two tables to {min, max} once (O(Nvx Nvy)), then scan (i1,i2) only.  Same expression, same bits.
__global__ void k_table_minmax(DGeo g, DAccel a, double* out4) {  // one CTA
  __shared__ double sh[4][32];
  const int ng = g.ng;
  double lo3 = 1e300, hi3 = -1e300, lo4 = 1e300, hi4 = -1e300;
  const int nf3 = (g.n[2] + 1) * g.n[3], nf4 = g.n[2] * (g.n[3] + 1);
  for (int t = threadIdx.x; t < nf3; t += blockDim.x) {  // vx faces: table component 1 (vy)
    const int i3 = t % (g.n[2] + 1) + ng, i4 = t / (g.n[2] + 1) + ng;
    const double v = a.vxf[i3 + (i64)(g.nd[2] + 1) * (i4 + (i64)g.nd[3])];
    lo3 = fmin(lo3, v); hi3 = fmax(hi3, v);
  }
  for (int t = threadIdx.x; t < nf4; t += blockDim.x) {  // vy faces: table component 0 (vx)
    const int i3 = t % g.n[2] + ng, i4 = t / g.n[2] + ng;
    const double v = a.vyf[i3 + (i64)g.nd[2] * i4];
    lo4 = fmin(lo4, v); hi4 = fmax(hi4, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo3 = fmin(lo3, __shfl_xor_sync(0xffffffffu, lo3, o)); hi3 = fmax(hi3, __shfl_xor_sync(0xffffffffu, hi3, o));
    lo4 = fmin(lo4, __shfl_xor_sync(0xffffffffu, lo4, o)); hi4 = fmax(hi4, __shfl_xor_sync(0xffffffffu, hi4, o));
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = lo3; sh[1][threadIdx.x >> 5] = hi3; sh[2][threadIdx.x >> 5] = lo4; sh[3][threadIdx.x >> 5] = hi4;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      sh[0][0] = fmin(sh[0][0], sh[0][k]); sh[1][0] = fmax(sh[1][0], sh[1][k]);
      sh[2][0] = fmin(sh[2][0], sh[2][k]); sh[3][0] = fmax(sh[3][0], sh[3][k]);
    }
    for (int k = 0; k < 4; ++k) out4[k] = sh[k][0];
  }
}
__global__ void k_max_accel(DGeo g, DAccel a, const double* __restrict__ mm, double* out2) {
  const i64 nxy = (i64)g.n[0] * g.n[1];
  double mx = 0.0, my = 0.0;
  const double vylo = mm[0], vyhi = mm[1], vxlo = mm[2], vxhi = mm[3];
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < nxy; t += (i64)gridDim.x * blockDim.x) {
    const i64 p = (t % g.n[0] + g.ng) + (i64)g.nd[0] * (t / g.n[0] + g.ng);
    mx = fmax(mx, fmax(fabs(accel_x_v(a, g, p, vylo)), fabs(accel_x_v(a, g, p, vyhi))));
    my = fmax(my, fmax(fabs(accel_y_v(a, g, p, vxlo)), fabs(accel_y_v(a, g, p, vxhi))));
  }
  mx = warp_max(mx);
  my = warp_max(my);
  if ((threadIdx.x & 31) == 0) {
    atomic_max_nonneg(out2 + 0, mx);
    atomic_max_nonneg(out2 + 1, my);
  }
}
cudaError_t max_accel(const lk_geom* g, const lk_accel* a, double* out2, double* scratch4, cudaStream_t st) {
  DGeo d = make_geo(g);
  DAccel da = make_accel(a);
  k_zero2<<<1, 1, 0, st>>>(out2);
  k_table_minmax<<<1, 1024, 0, st>>>(d, da, scratch4);
  g_launches += 2;
  const i64 total = (i64)g->n[0] * g->n[1];
  unsigned blocks = (unsigned)min((i64)nblk(total, 256), (i64)148 * 4);
  k_max_accel<<<blocks, 256, 0, st>>>(d, da, scratch4, out2);
  return LK_LAUNCHED();
}
cudaError_t set_phase_space_vel(double* vel3, double* vel4, const lk_geom* g, const lk_accel* a, double* out2,
                                cudaStream_t st) {
  DGeo d = make_geo(g);
  DAccel da = make_accel(a);
  k_zero2<<<1, 1, 0, st>>>(out2);
  ++g_launches;
  i64 t3 = (i64)(d.nd[2] + 1) * d.nd[3] * d.nd[0] * d.nd[1];
  i64 t4 = (i64)(d.nd[3] + 1) * d.nd[0] * d.nd[1] * d.nd[2];
  k_vel3<<<(unsigned)min((i64)nblk(t3, 256), (i64)148 * 32), 256, 0, st>>>(d, da, vel3, out2);
  ++g_launches;
  k_vel4<<<(unsigned)min((i64)nblk(t4, 256), (i64)148 * 32), 256, 0, st>>>(d, da, vel4, out2);
  return LK_LAUNCHED();
}

// ---------------------------------------------------------------------------------------------
// a6 setAccelerationBCs4D (KineticSpeciesF.f:1036-1162).  Pass 0: vx boundaries, one thread per
// (i1,i2,i4) of the data box; pass 1: vy boundaries, one thread per (i1,i2,i3) of the data box (it
// reads the vx ghosts written by pass 0, as the reference's second loop nest does).
// ---------------------------------------------------------------------------------------------
__global__ void k_accel_bcs(DGeo g, DAccel a, DInflow ic, double* __restrict__ u, int pass, int at_lo, int at_hi) {
  const int ng = g.ng;
  const int nother = (pass == 0) ? g.nd[3] : g.nd[2];
  const i64 total = (i64)g.nd[0] * g.nd[1] * nother;
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int i1 = (int)(t % g.nd[0]);
  i64 r = t / g.nd[0];
  int i2 = (int)(r % g.nd[1]);
  int io = (int)(r / g.nd[1]);
  if (pass == 0) {
    const int i4 = io, n3a = ng, n3b = ng + g.n[2] - 1;
    const i64 s = g.s[2];
    if (at_hi) {
      if (accel_x(a, g, i1, i2, n3b + 1, i4) >= 0.0) {
        double* p = u + gidx(g, i1, i2, n3b, i4);
        for (int ig = 1; ig <= ng; ++ig) p[ig * s] = bc_extrap(p[(ig - 1) * s], p[(ig - 2) * s], p[(ig - 3) * s]);
      } else {
        for (int ig = 1; ig <= ng; ++ig) u[gidx(g, i1, i2, n3b + ig, i4)] = inflow_value(ic, g, i1, i2, n3b + ig, i4, 3);
      }
    }
    if (at_lo) {
      if (accel_x(a, g, i1, i2, n3a, i4) > 0.0) {
        for (int ig = 1; ig <= ng; ++ig) u[gidx(g, i1, i2, n3a - ig, i4)] = inflow_value(ic, g, i1, i2, n3a - ig, i4, 3);
      } else {
        double* p = u + gidx(g, i1, i2, n3a, i4);
        for (int ig = 1; ig <= ng; ++ig) p[-ig * s] = bc_extrap(p[(1 - ig) * s], p[(2 - ig) * s], p[(3 - ig) * s]);
      }
    }
  } else {
    const int i3 = io, n4a = ng, n4b = ng + g.n[3] - 1;
    const i64 s = g.s[3];
    if (at_hi) {
      if (accel_y(a, g, i1, i2, i3, n4b + 1) >= 0.0) {
        double* p = u + gidx(g, i1, i2, i3, n4b);
        for (int ig = 1; ig <= ng; ++ig) p[ig * s] = bc_extrap(p[(ig - 1) * s], p[(ig - 2) * s], p[(ig - 3) * s]);
      } else {
        for (int ig = 1; ig <= ng; ++ig) u[gidx(g, i1, i2, i3, n4b + ig)] = inflow_value(ic, g, i1, i2, i3, n4b + ig, 4);
      }
    }
    if (at_lo) {
      if (accel_y(a, g, i1, i2, i3, n4a) > 0.0) {
        for (int ig = 1; ig <= ng; ++ig) u[gidx(g, i1, i2, i3, n4a - ig)] = inflow_value(ic, g, i1, i2, i3, n4a - ig, 4);
      } else {
        double* p = u + gidx(g, i1, i2, i3, n4a);
        for (int ig = 1; ig <= ng; ++ig) p[-ig * s] = bc_extrap(p[(1 - ig) * s], p[(2 - ig) * s], p[(3 - ig) * s]);
      }
    }
  }
}
cudaError_t set_accel_bcs(double* f, const lk_geom* g, const lk_accel* a, const lk_inflow* ic, const int at[4],
                          cudaStream_t st) {
  DGeo d = make_geo(g);
  DAccel da = make_accel(a);
  DInflow di;
  memset(&di, 0, sizeof(di));
  if (ic) {
    di.kind = ic->kind; di.fx = ic->fx; di.fv = ic->fv; di.fx2 = ic->fx2; di.fv2 = ic->fv2;
    di.ghost3 = ic->ghost3; di.ghost4 = ic->ghost4; di.fnorm = ic->fnorm; di.frac = ic->frac;
  }
  if (at[0] || at[1]) {
    i64 total = (i64)d.nd[0] * d.nd[1] * d.nd[3];
    k_accel_bcs<<<nblk(total, 128), 128, 0, st>>>(d, da, di, f, 0, at[0], at[1]);
    ++g_launches;
  }
  if (at[2] || at[3]) {
    i64 total = (i64)d.nd[0] * d.nd[1] * d.nd[2];
    k_accel_bcs<<<nblk(total, 128), 128, 0, st>>>(d, da, di, f, 1, at[2], at[3]);
    ++g_launches;
  }
  return cudaGetLastError();
}

// RK stage update from a materialised rhs (RK4Integrator.H:149-171, RK6Integrator.H:96-131: the addSolnData /
// copySolnData sequence in the reference's order of roundings): what the fused stage kernel does in its epilogue, for
// callers that put something between the rhs and the update (the Krook layer of completeRHS).  Interior cells.
__global__ void k_rk_update(DGeo g, DUpd upd, const double* __restrict__ rhs) {
  const i64 total = (i64)g.n[0] * g.n[1] * g.n[2] * g.n[3];
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const int i1 = (int)(t % g.n[0]);
    i64 r = t / g.n[0];
    const int i2 = (int)(r % g.n[1]);
    r /= g.n[1];
    const int i3 = (int)(r % g.n[2]), i4 = (int)(r / g.n[2]);
    const i64 idx = gidx(g, i1 + g.ng, i2 + g.ng, i3 + g.ng, i4 + g.ng);
    rk_update(upd, idx, rhs[idx]);
  }
}
cudaError_t rk_stage_update(const double* rhs, const lk_geom* g, const lk_rk_update* upd, cudaStream_t st) {
  DGeo d = make_geo(g);
  DUpd du = make_upd(upd);
  const i64 total = (i64)d.n[0] * d.n[1] * d.n[2] * d.n[3];
  k_rk_update<<<(unsigned)min((i64)nblk(total, 256), (i64)148 * 32), 256, 0, st>>>(d, du, rhs);
  ++g_launches;
  return cudaGetLastError();
}

// The inflow sample in EVERY velocity ghost cell of the data box (the value setAccelerationBCs4D gives a ghost whose
// face has the acceleration pointing inward; it depends on the position only).  The pipelined stage kernel relies on
// it (lk_rk_update.inflow_preset): it extrapolates the outflow ghosts on the fly and leaves memory alone.
__global__ void k_preset_inflow(DGeo g, DInflow ic, double* __restrict__ u) {
  const int ng = g.ng;
  const i64 nfull = (i64)2 * ng * g.nd[2];                 // the 2 ng whole vy ghost planes
  const i64 nv = nfull + (i64)2 * ng * g.n[3];             // + 2 ng vx ghost cells of every interior vy plane
  const i64 total = (i64)g.nd[0] * g.nd[1] * nv;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const int i1 = (int)(t % g.nd[0]);
    const i64 r = t / g.nd[0];
    const int i2 = (int)(r % g.nd[1]);
    const i64 j = r / g.nd[1];
    int i3, i4, dir;
    if (j < nfull) {
      i3 = (int)(j % g.nd[2]);
      const int l = (int)(j / g.nd[2]);
      i4 = (l < ng) ? l : l - ng + ng + g.n[3];
      dir = 4;
      if (i3 < ng || i3 >= ng + g.n[2]) dir = 3;           // corner cells: never read by a stencil; any table will do
    } else {
      const i64 k = j - nfull;
      const int l = (int)(k % (2 * ng));
      i4 = ng + (int)(k / (2 * ng));
      i3 = (l < ng) ? l : l - ng + ng + g.n[2];
      dir = 3;
    }
    u[gidx(g, i1, i2, i3, i4)] = inflow_value(ic, g, i1, i2, i3, i4, dir);
  }
}
cudaError_t preset_inflow(double* f, const lk_geom* g, const lk_inflow* ic, cudaStream_t st) {
  DGeo d = make_geo(g);
  DInflow di;
  memset(&di, 0, sizeof(di));
  if (ic) {
    di.kind = ic->kind; di.fx = ic->fx; di.fv = ic->fv; di.fx2 = ic->fx2; di.fv2 = ic->fv2;
    di.ghost3 = ic->ghost3; di.ghost4 = ic->ghost4; di.fnorm = ic->fnorm; di.frac = ic->frac;
  }
  const i64 total = (i64)d.nd[0] * d.nd[1] * ((i64)2 * d.ng * d.nd[2] + (i64)2 * d.ng * d.n[3]);
  k_preset_inflow<<<(unsigned)min((i64)nblk(total, 256), (i64)148 * 32), 256, 0, st>>>(d, di, f);
  ++g_launches;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// a14: periodic wrap (x sweep then y sweep over the full data box of the other dims) and halo slabs
// ---------------------------------------------------------------------------------------------
__global__ void k_periodic_x(DGeo g, double* __restrict__ u) {
  const int ng = g.ng;
  const i64 rows = (i64)g.nd[1] * g.nd[2] * g.nd[3];
  const i64 total = rows * 2 * ng;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    int k = (int)(t % (2 * ng));
    i64 row = t / (2 * ng);
    double* p = u + row * g.s[1];
    if (k < ng)
      p[k] = p[k + g.n[0]];
    else
      p[g.n[0] + k] = p[k];  // ghost ng+n+(k-ng) <- interior ng+(k-ng)
  }
}
__global__ void k_periodic_y(DGeo g, double* __restrict__ u) {
  const int ng = g.ng;
  const i64 total = (i64)g.nd[0] * 2 * ng * g.nd[2] * g.nd[3];
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    int i1 = (int)(t % g.nd[0]);
    i64 r = t / g.nd[0];
    int k = (int)(r % (2 * ng));
    i64 plane = r / (2 * ng);  // (i3,i4) flattened
    double* p = u + plane * g.s[2] + i1;
    if (k < ng)
      p[(i64)k * g.s[1]] = p[(i64)(k + g.n[1]) * g.s[1]];
    else
      p[(i64)(g.n[1] + k) * g.s[1]] = p[(i64)k * g.s[1]];
  }
}
cudaError_t periodic_fill_4d(double* f, const lk_geom* g, int px, int py, cudaStream_t st) {
  DGeo d = make_geo(g);
  if (px) {
    i64 total = (i64)d.nd[1] * d.nd[2] * d.nd[3] * 2 * d.ng;
    k_periodic_x<<<(unsigned)min((i64)nblk(total, 256), (i64)148 * 32), 256, 0, st>>>(d, f);
    ++g_launches;
  }
  if (py) {
    i64 total = (i64)d.nd[0] * 2 * d.ng * d.nd[2] * d.nd[3];
    k_periodic_y<<<(unsigned)min((i64)nblk(total, 256), (i64)148 * 32), 256, 0, st>>>(d, f);
    ++g_launches;
  }
  return cudaGetLastError();
}
// x slabs: (ng, n2, n3d, n4d) interior rows only in y (y ghosts are filled by the later y exchange,
// which carries full x rows incl. the x ghosts just received: same sequencing as the reference's
// per-dimension sweeps); y slabs: (n1d, ng, n3d, n4d)
__global__ void k_halo_x(DGeo g, double* __restrict__ buf, double* __restrict__ f, int side, int unpack) {
  const int ng = g.ng;
  const i64 total = (i64)ng * g.n[1] * g.nd[2] * g.nd[3];
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    int k = (int)(t % ng);
    i64 r = t / ng;
    int i2 = (int)(r % g.n[1]) + ng;
    i64 pl = r / g.n[1];
    int i1;
    if (!unpack)
      i1 = side ? (g.n[0] + k) : (ng + k);  // interior layers next to that side
    else
      i1 = side ? (ng + g.n[0] + k) : k;  // ghost layers on that side
    i64 idx = i1 + g.s[1] * i2 + g.s[2] * pl;
    if (unpack) f[idx] = buf[t]; else buf[t] = f[idx];
  }
}
__global__ void k_halo_y(DGeo g, double* __restrict__ buf, double* __restrict__ f, int side, int unpack) {
  const int ng = g.ng;
  const i64 total = (i64)g.nd[0] * ng * g.nd[2] * g.nd[3];
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    int i1 = (int)(t % g.nd[0]);
    i64 r = t / g.nd[0];
    int k = (int)(r % ng);
    i64 pl = r / ng;
    int i2;
    if (!unpack)
      i2 = side ? (g.n[1] + k) : (ng + k);
    else
      i2 = side ? (ng + g.n[1] + k) : k;
    i64 idx = i1 + g.s[1] * i2 + g.s[2] * pl;
    if (unpack) f[idx] = buf[t]; else buf[t] = f[idx];
  }
}
cudaError_t halo_pack(double* buf, const double* f, const lk_geom* g, int dir, int side, cudaStream_t st) {
  DGeo d = make_geo(g);
  i64 total = dir == 0 ? (i64)d.ng * d.n[1] * d.nd[2] * d.nd[3] : (i64)d.nd[0] * d.ng * d.nd[2] * d.nd[3];
  unsigned blocks = (unsigned)min((i64)nblk(total, 256), (i64)148 * 32);
  if (dir == 0) k_halo_x<<<blocks, 256, 0, st>>>(d, buf, (double*)f, side, 0);
  else k_halo_y<<<blocks, 256, 0, st>>>(d, buf, (double*)f, side, 0);
  return LK_LAUNCHED();
}
cudaError_t halo_unpack(double* f, const double* buf, const lk_geom* g, int dir, int side, cudaStream_t st) {
  DGeo d = make_geo(g);
  i64 total = dir == 0 ? (i64)d.ng * d.n[1] * d.nd[2] * d.nd[3] : (i64)d.nd[0] * d.ng * d.nd[2] * d.nd[3];
  unsigned blocks = (unsigned)min((i64)nblk(total, 256), (i64)148 * 32);
  if (dir == 0) k_halo_x<<<blocks, 256, 0, st>>>(d, (double*)buf, f, side, 1);
  else k_halo_y<<<blocks, 256, 0, st>>>(d, (double*)buf, f, side, 1);
  return LK_LAUNCHED();
}

// ---------------------------------------------------------------------------------------------
// a3+a8 (+a10/a11): fused Vlasov RHS.  variant 1 = one thread per cell straight from global memory
// (cross-check kernel); variant 0 = tiled shared-memory kernel (lk_stencil.cuh).
// ---------------------------------------------------------------------------------------------
template <int ORDER>
__global__ void __launch_bounds__(256)
k_rhs_naive(DGeo g, const double* __restrict__ f, const double* __restrict__ vel, DAccel a, DUpd upd,
            double* __restrict__ rhs_out, int flags) {
  const i64 total = (i64)g.n[0] * g.n[1] * g.n[2] * g.n[3];
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int ng = g.ng;
  int i1 = (int)(t % g.n[0]) + ng;
  i64 r = t / g.n[0];
  int i2 = (int)(r % g.n[1]) + ng;
  r /= g.n[1];
  int i3 = (int)(r % g.n[2]) + ng;
  int i4 = (int)(r / g.n[2]) + ng;
  const i64 idx = gidx(g, i1, i2, i3, i4);
  const double* p = f + idx;
  constexpr int NG = (ORDER == 4) ? 2 : 3, W = 2 * NG;
  const double FS = FaceScale<ORDER>::v;
  // faces above (R) and below (L) the cell along stride s
  auto faces = [&](i64 s, bool posR, bool posL, double& uR, double& uL) {
    double w[W + 1];
#pragma unroll
    for (int k = 0; k <= W; ++k) w[k] = p[(k - NG) * s];
    uL = fit_face<ORDER>(w, posL);
    uR = fit_face<ORDER>(w + 1, posR);
  };
  double rhs = 0.0, uR, uL;
  if (flags & 4) rhs = rhs_out[idx];
  if (flags & 1) {
    const double vx = __ldg(vel + i3 + (i64)g.nd[2] * i4);
    const double vy = __ldg(vel + i3 + (i64)g.nd[2] * (i4 + (i64)g.nd[3]));
    faces(1, vx > 0.0, vx > 0.0, uR, uL);
    rhs = sub_flux((flags & 4) ? rhs : 0.0, vx, uR, uL, g.dx[0], (1.0 / g.dx[0]) * FS);  // the x pass assigns (KineticSpeciesF.f:1999)
    faces(g.s[1], vy > 0.0, vy > 0.0, uR, uL);
    rhs = sub_flux(rhs, vy, uR, uL, g.dx[1], (1.0 / g.dx[1]) * FS);
  }
  if (flags & 2) {
    const double ax = accel_x(a, g, i1, i2, i3, i4);
    const double axl = (i3 > ng) ? accel_x(a, g, i1, i2, i3 - 1, i4) : ax;  // face reuse uLeft=uRight
    faces(g.s[2], ax > 0.0, axl > 0.0, uR, uL);
    rhs = sub_flux(rhs, ax, uR, uL, g.dx[2], (1.0 / g.dx[2]) * FS);
    const double ay = accel_y(a, g, i1, i2, i3, i4);
    const double ayl = (i4 > ng) ? accel_y(a, g, i1, i2, i3, i4 - 1) : ay;
    faces(g.s[3], ay > 0.0, ayl > 0.0, uR, uL);
    rhs = sub_flux(rhs, ay, uR, uL, g.dx[3], (1.0 / g.dx[3]) * FS);
  }
  if (upd.krook_nu && (flags & 2)) rhs = krook_term(upd, g, rhs, p[0], i1, i2, i3, i4);  // completeRHS, after the acceleration pass
  if (rhs_out) rhs_out[idx] = rhs;
  if (upd.active) rk_update(upd, idx, rhs);
}

cudaError_t vlasov_rhs(double* rhs_out, const double* f, const lk_geom* g, const double* velocities,
                       const lk_accel* a, const lk_rk_update* upd, int flags, int variant, double* mom_part,
                       int nmom, cudaStream_t st) {
  DGeo d = make_geo(g);
  DAccel da;
  memset(&da, 0, sizeof(da));
  if (a) da = make_accel(a);
  DUpd du = make_upd(upd);
  i64 total = (i64)g->n[0] * g->n[1] * g->n[2] * g->n[3];
  if (total <= 0) return cudaSuccess;
  if (variant == 0 || variant == 2) {
    DMom dm;
    dm.part = mom_part;
    dm.nmom = (mom_part && upd) ? nmom : 0;
    dm.nparts = march_moment_parts(d);
#if !LK_STRICT
    // aligned grids, RK4-shaped update, acceleration independent of the swept velocity: the pipelined kernel
    static const bool no_pipe = getenv("LK_NO_PIPE") != nullptr;
    if (!no_pipe && variant == 0 && pipe_eligible(d, da, du, rhs_out, flags)) {
      bool used = false;
      // velocity-boundary fill folded into the boundary tiles: f's ghosts hold the inflow sample (k_preset_inflow)
      int bcfold = (LK_PIPE_FOLD && upd && upd->accel_bcs && upd->inflow_preset) ? 3 : 0;
      if (upd && upd->tile_set) bcfold |= ((upd->tile_set & 3) << 4) | ((upd->cut_dirs & 3) << 6);
      cudaError_t e = (d.order == 4) ? launch_pipe<4>(d, f, velocities, da, du, dm, bcfold, st, &used)
                                     : launch_pipe<6>(d, f, velocities, da, du, dm, bcfold, st, &used);
      if (used) {
        if (e == cudaSuccess) { ++g_launches; ++g_pipe_launches; }
        return e;
      }
    }
#endif
    if (upd && upd->tile_set) return cudaErrorNotSupported;   // tile subsets exist in the pipelined kernel only
    cudaError_t e = launch_stage_march(d, f, velocities, da, du, rhs_out, flags, dm, st);
    if (e == cudaSuccess) ++g_launches;
    return e;
  }
  if (mom_part) return cudaErrorNotSupported;  // moments come from the marching kernel only
  if (g->order == 4)
    k_rhs_naive<4><<<nblk(total, 256), 256, 0, st>>>(d, f, velocities, da, du, rhs_out, flags);
  else
    k_rhs_naive<6><<<nblk(total, 256), 256, 0, st>>>(d, f, velocities, da, du, rhs_out, flags);
  return LK_LAUNCHED();
}
int stage_moment_parts(const lk_geom* g) { return march_moment_parts(make_geo(g)); }
bool stage_uses_pipe(const lk_geom* g, const lk_accel* a, const lk_rk_update* upd, double* rhs_out, int flags, int variant) {
#if LK_STRICT
  (void)g; (void)a; (void)upd; (void)rhs_out; (void)flags; (void)variant;
  return false;
#else
  static const bool no_pipe = getenv("LK_NO_PIPE") != nullptr;
  if (no_pipe || variant != 0 || !a || !upd) return false;
  DGeo d = make_geo(g);
  // the TMA path needs a 16-byte aligned array base and an even row length (get_map); every cudaMalloc'ed array has them
  return pipe_eligible(d, make_accel(a), make_upd(upd), rhs_out, flags) && (d.nd[0] % 2 == 0);
#endif
}
bool stage_folds_bcs(const lk_geom* g, const lk_accel* a, const lk_rk_update* upd, double* rhs_out, int flags, int variant) {
#if LK_STRICT
  (void)g; (void)a; (void)upd; (void)rhs_out; (void)flags; (void)variant;
  return false;
#else
  // ... and the extrapolations of the two ends of the march must not feed each other
  return LK_PIPE_FOLD && upd && upd->inflow_preset && stage_uses_pipe(g, a, upd, rhs_out, flags, variant) &&
         g->n[3] >= 2 * g->ng;
#endif
}
#if defined(LK_PIPE_TRACE) && !LK_STRICT
extern "C" int lk_debug_pipe_trace(long long* stamps, int* smid) {
  if (cudaMemcpyFromSymbol(stamps, g_pipe_trace, sizeof(long long) * 296 * 8 * 4 * 8) != cudaSuccess) return 1;
  if (cudaMemcpyFromSymbol(smid, g_pipe_smid, sizeof(int) * 296) != cudaSuccess) return 1;
  return 0;
}
#endif

// ---------------------------------------------------------------------------------------------
// a12/a13: velocity moments.  Grid (x-blocks, i2, chunk): each thread owns one (i1,i2) and sums its
// chunk of vy planes sequentially (i3 inner, i4 outer -- the reference's order, ReductionSchedule.C:
// 434-441), coalesced across i1.  chunks==1 reproduces the reference sum bit for bit; chunks>1 writes
// partials that k_moment_finish adds in chunk order (deterministic).
// NMOM = 1: sum f ; NMOM = 3: sum f*vx, f*vy, f*vz(i1,i2) (computecurrents, KineticSpeciesF.f:2430-2437)
// ---------------------------------------------------------------------------------------------
template <int NMOM>
__global__ void k_moment_partial(DGeo g, const double* __restrict__ f, const double* __restrict__ vel,
                                 const double* __restrict__ vz, double* __restrict__ part, int chunks) {
  const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i1 >= g.n[0]) return;
  const int i2 = blockIdx.y, c = blockIdx.z;
  const int per = (g.n[3] + chunks - 1) / chunks;
  const int j0 = c * per, j1 = min(g.n[3], j0 + per);
  const int ng = g.ng;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  double vzv = 0.0;
  if (NMOM == 3) vzv = vz[(i1 + ng) + (i64)g.nd[0] * (i2 + ng)];
  for (int j = j0; j < j1; ++j) {
    const double* p = f + gidx(g, i1 + ng, i2 + ng, ng, j + ng);
    const double* v = vel + ng + (i64)g.nd[2] * (j + ng);
#pragma unroll 4
    for (int k = 0; k < g.n[2]; ++k) {
      double u = p[(i64)k * g.s[2]];
      if (NMOM == 1) {
        s0 += u;
      } else {
        s0 += u * __ldg(v + k);
        s1 += u * __ldg(v + k + (i64)g.nd[2] * g.nd[3]);
        s2 += u * vzv;
      }
    }
  }
  const i64 nxy = (i64)g.n[0] * g.n[1];
  const i64 o = i1 + (i64)g.n[0] * i2 + nxy * c;
  part[o] = s0;
  if (NMOM == 3) {
    part[o + nxy * chunks] = s1;
    part[o + 2 * nxy * chunks] = s2;
  }
}
// dst(n1d,n2d): zero, ordered sum of partials, then *dv, then *weight (ReductionSchedule.C:86-89)
__global__ void k_moment_finish(DGeo g, const double* __restrict__ part, int chunks, double dv, double w,
                                double* __restrict__ d0, double* __restrict__ d1, double* __restrict__ d2, int nmom) {
  const i64 tot = (i64)g.nd[0] * g.nd[1];
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= tot) return;
  int i1 = (int)(t % g.nd[0]) - g.ng, i2 = (int)(t / g.nd[0]) - g.ng;
  const bool inside = i1 >= 0 && i1 < g.n[0] && i2 >= 0 && i2 < g.n[1];
  const i64 nxy = (i64)g.n[0] * g.n[1];
  for (int m = 0; m < nmom; ++m) {
    double s = 0.0;
    if (inside)
      for (int c = 0; c < chunks; ++c) s += part[i1 + (i64)g.n[0] * i2 + nxy * (c + (i64)chunks * m)];
    s *= dv;
    s *= w;
    (m == 0 ? d0 : (m == 1 ? d1 : d2))[t] = s;
  }
}
cudaError_t reduce_4d_to_2d(double* dst, const double* f, const lk_geom* g, double dv, double w, double* scratch,
                            int chunks, cudaStream_t st) {
  DGeo d = make_geo(g);
  dim3 grid(nblk(g->n[0], 64), g->n[1], chunks);
  k_moment_partial<1><<<grid, 64, 0, st>>>(d, f, nullptr, nullptr, scratch, chunks);
  ++g_launches;
  k_moment_finish<<<nblk((i64)d.nd[0] * d.nd[1], 128), 128, 0, st>>>(d, scratch, chunks, dv, w, dst, nullptr, nullptr, 1);
  return LK_LAUNCHED();
}
cudaError_t current_density(double* Jx, double* Jy, double* Jz, const double* f, const lk_geom* g,
                            const double* velocities, const double* vz, double dv, double w, double* scratch,
                            int chunks, cudaStream_t st) {
  DGeo d = make_geo(g);
  dim3 grid(nblk(g->n[0], 64), g->n[1], chunks);
  k_moment_partial<3><<<grid, 64, 0, st>>>(d, f, velocities, vz, scratch, chunks);
  ++g_launches;
  k_moment_finish<<<nblk((i64)d.nd[0] * d.nd[1], 128), 128, 0, st>>>(d, scratch, chunks, dv, w, Jx, Jy, Jz, 3);
  return LK_LAUNCHED();
}

// finish of the moments produced by the marching kernel's epilogue: part[(m*nparts + p)*nxy + xy]
cudaError_t moments_finish(double* d0, double* d1, double* d2, const double* part, int nparts, int nmom,
                           const lk_geom* g, double dv, double w, cudaStream_t st) {
  DGeo d = make_geo(g);
  k_moment_finish<<<nblk((i64)d.nd[0] * d.nd[1], 128), 128, 0, st>>>(d, part, nparts, dv, w, d0, d1, d2, nmom);
  return LK_LAUNCHED();
}
// ke_e_dot from the first vx-moment: charge*dx*dy*dvx*dvy * sum_xy ext(x,y) * (sum_p part1[p][xy]);
// fixed-order two-level tree (KineticSpeciesF.f:2563-2602 sums cell by cell; tolerance in the tests)
constexpr int KE_BLOCKS = 148;
__device__ double g_ke_part[KE_BLOCKS];
__device__ unsigned int g_ke_count = 0;
__global__ void k_ke_from_moment(DGeo g, const double* __restrict__ part1, int nparts, const double* __restrict__ ext,
                                 double scale, double* out) {
  __shared__ double sh[32];
  __shared__ bool last;
  const int nxy = g.n[0] * g.n[1];
  double s = 0.0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nxy; t += gridDim.x * blockDim.x) {
    double m = 0.0;
    for (int p = 0; p < nparts; ++p) m += part1[(i64)p * nxy + t];
    const int i1 = t % g.n[0] + g.ng, i2 = t / g.n[0] + g.ng;
    s += ext[i1 + (i64)g.nd[0] * i2] * m;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) b += sh[k];
    g_ke_part[blockIdx.x] = b;
    __threadfence();
    last = (atomicAdd(&g_ke_count, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  // the last block to finish adds the per-block sums in block order: deterministic whatever the schedule
  if (last && threadIdx.x == 0) {
    __threadfence();
    double b = 0.0;
    for (int k = 0; k < (int)gridDim.x; ++k) b += ((volatile double*)g_ke_part)[k];
    out[0] = b * scale;
    g_ke_count = 0;
  }
}
cudaError_t ke_from_moment(double* out, const double* part1, int nparts, const lk_geom* g, double charge,
                           const double* ext, cudaStream_t st) {
  DGeo d = make_geo(g);
  const double scale = charge * g->dx[0] * g->dx[1] * g->dx[2] * g->dx[3];
  const int nxy = g->n[0] * g->n[1];
  const int blocks = (int)min((i64)KE_BLOCKS, (i64)nblk(nxy, 256));
  k_ke_from_moment<<<blocks, 256, 0, st>>>(d, part1, nparts, ext, scale, out);
  return LK_LAUNCHED();
}

// a9 computekeedot: sum_{cells} ext(i1,i2,0)*vx*u, then *charge*dx*dy*dvx*dvy.  Deterministic two-level
// tree (the reference's fully sequential sum is not reproduced bit for bit; tolerance in the tests).
__global__ void k_ke_partial(DGeo g, const double* __restrict__ f, const double* __restrict__ vel,
                             const double* __restrict__ ext, double* __restrict__ part) {
  __shared__ double sh[8];
  const i64 total = (i64)g.n[0] * g.n[1] * g.n[2] * g.n[3];
  double s = 0.0;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    int i1 = (int)(t % g.n[0]) + g.ng;
    i64 r = t / g.n[0];
    int i2 = (int)(r % g.n[1]) + g.ng;
    r /= g.n[1];
    int i3 = (int)(r % g.n[2]) + g.ng;
    int i4 = (int)(r / g.n[2]) + g.ng;
    s += ext[i1 + (i64)g.nd[0] * i2] * __ldg(vel + i3 + (i64)g.nd[2] * i4) * f[gidx(g, i1, i2, i3, i4)];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) b += sh[k];
    part[blockIdx.x] = b;
  }
}
__global__ void k_ke_finish(const double* part, int n, double scale, double* out) {
  double s = 0.0;
  for (int k = 0; k < n; ++k) s += part[k];
  out[0] = s * scale;
}
cudaError_t ke_e_dot(double* out, const double* f, const lk_geom* g, double charge, const double* velocities,
                     const double* ext, double* scratch, int nblocks, cudaStream_t st) {
  DGeo d = make_geo(g);
  k_ke_partial<<<nblocks, 256, 0, st>>>(d, f, velocities, ext, scratch);
  ++g_launches;
  // ke_e_dot*charge*dx(1)*dx(2)*dx(3)*dx(4), left to right (KineticSpeciesF.f:2599)
  double scale = charge * g->dx[0] * g->dx[1] * g->dx[2] * g->dx[3];
  k_ke_finish<<<1, 1, 0, st>>>(scratch, nblocks, scale, out);
  return LK_LAUNCHED();
}

// ---------------------------------------------------------------------------------------------
// a16: Poisson.  The 2D problem is tiny (<= 512^2), every kernel is one thread per output value with a
// sequential inner sum so that the strict build reproduces the oracle's DFT bit for bit.
// ---------------------------------------------------------------------------------------------
__global__ void k_neutralize(double* rho, int n1, int n2, int ng) {
  // neutralizeCharge4D (PoissonF.f:41-61).  Strict: the reference's sequential sum; production: block tree.
  __shared__ double sh[32];
  __shared__ double mean;
  const i64 n1d = n1 + 2 * ng;
  const int total = n1 * n2;
#if LK_STRICT
  if (threadIdx.x == 0) {
    double sum = 0.0;
    for (int i2 = ng; i2 < ng + n2; ++i2)
      for (int i1 = ng; i1 < ng + n1; ++i1) sum = sum + rho[i1 + n1d * i2];
    mean = sum / total;
  }
  (void)sh;
#else
  double s = 0.0;
  for (int t = threadIdx.x; t < total; t += blockDim.x) s += rho[(t % n1 + ng) + n1d * (t / n1 + ng)];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) b += sh[k];
    mean = b / total;
  }
#endif
  __syncthreads();
  const double m = mean;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    i64 o = (t % n1 + ng) + n1d * (t / n1 + ng);
    rho[o] = rho[o] - m;
  }
}
cudaError_t neutralize(double* rho, int n1, int n2, int ng, cudaStream_t st) {
  k_neutralize<<<1, 1024, 0, st>>>(rho, n1, n2, ng);
  return LK_LAUNCHED();
}

// four DFT passes (same algebra as the FFTW r2c -> divide by symbol -> c2r of LokiPoissonSolveFFT.C:128-170)
__global__ void k_dft_y_fwd(const double* __restrict__ rho, int nx, int ny, int ng, const double* __restrict__ cy,
                            double* __restrict__ T) {
  const int nyh = ny / 2 + 1;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nx * nyh) return;
  int a = t % nx, j = t / nx;  // a fastest: coalesced reads of rho rows
  const i64 n1d = nx + 2 * ng;
  double re = 0.0, im = 0.0;
  int m = 0;
  for (int b = 0; b < ny; ++b) {
    double v = rho[(a + ng) + n1d * (b + ng)];
    re += v * cy[2 * m];
    im -= v * cy[2 * m + 1];
    m += j;
    if (m >= ny) m -= ny;
  }
  T[2 * (a * nyh + j)] = re;
  T[2 * (a * nyh + j) + 1] = im;
}
__global__ void k_dft_x_fwd(const double* __restrict__ T, int nx, int ny, const double* __restrict__ cx,
                            const double* __restrict__ sx, const double* __restrict__ sy, double* __restrict__ X) {
  const int nyh = ny / 2 + 1;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nx * nyh) return;
  int j = t % nyh, i = t / nyh;
  double re = 0.0, im = 0.0;
  int m = 0;
  for (int a = 0; a < nx; ++a) {
    double tr = T[2 * (a * nyh + j)], ti = T[2 * (a * nyh + j) + 1];
    re += tr * cx[2 * m] + ti * cx[2 * m + 1];
    im += ti * cx[2 * m] - tr * cx[2 * m + 1];
    m += i;
    if (m >= nx) m -= nx;
  }
  if (sx[i] != 0.0 || sy[j] != 0.0) {
    re /= sx[i] + sy[j];
    im /= sx[i] + sy[j];
  }
  X[2 * (i * nyh + j)] = re;
  X[2 * (i * nyh + j) + 1] = im;
}
__global__ void k_dft_x_inv(const double* __restrict__ X, int nx, int ny, const double* __restrict__ cx,
                            double* __restrict__ T) {
  const int nyh = ny / 2 + 1;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nx * nyh) return;
  int j = t % nyh, a = t / nyh;
  double re = 0.0, im = 0.0;
  int m = 0;
  for (int i = 0; i < nx; ++i) {
    double xr = X[2 * (i * nyh + j)], xi = X[2 * (i * nyh + j) + 1];
    re += xr * cx[2 * m] - xi * cx[2 * m + 1];
    im += xi * cx[2 * m] + xr * cx[2 * m + 1];
    m += a;
    if (m >= nx) m -= nx;
  }
  T[2 * (a * nyh + j)] = re;
  T[2 * (a * nyh + j) + 1] = im;
}
__global__ void k_dft_y_inv(const double* __restrict__ T, int nx, int ny, int ng, const double* __restrict__ cy,
                            double* __restrict__ phi) {
  const int nyh = ny / 2 + 1;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nx * ny) return;
  int a = t % nx, b = t / nx;
  const i64 n1d = nx + 2 * ng;
  double acc = 0.0;
  int m = 0;
  for (int j = 0; j < nyh; ++j) {
    double ur = T[2 * (a * nyh + j)], ui = T[2 * (a * nyh + j) + 1];
    double term = ur * cy[2 * m] - ui * cy[2 * m + 1];
    bool self_conj = (j == 0) || (2 * j == ny);
    acc += self_conj ? term : 2.0 * term;
    m += b;
    if (m >= ny) m -= ny;
  }
  phi[(a + ng) + n1d * (b + ng)] = acc;
}
cudaError_t poisson_dft(double* phi, const double* rho, int nx, int ny, int ng, const double* sx, const double* sy,
                        const double* cx, const double* cy, double* T, double* X, cudaStream_t st) {
  const int nyh = ny / 2 + 1;
  k_dft_y_fwd<<<nblk((i64)nx * nyh, 128), 128, 0, st>>>(rho, nx, ny, ng, cy, T);
  k_dft_x_fwd<<<nblk((i64)nx * nyh, 128), 128, 0, st>>>(T, nx, ny, cx, sx, sy, X);
  k_dft_x_inv<<<nblk((i64)nx * nyh, 128), 128, 0, st>>>(X, nx, ny, cx, T);
  k_dft_y_inv<<<nblk((i64)nx * ny, 128), 128, 0, st>>>(T, nx, ny, ng, cy, phi);
  g_launches += 4;
  return cudaGetLastError();
}

// computeEFieldFromPotential (PoissonF.f:68-123): E = +grad(phi); em comps 0,1 interior only
__global__ void k_efield(double* __restrict__ em, const double* __restrict__ phi, int n1, int n2, int ng, int order,
                         double dx, double dy) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n1 * n2) return;
  const i64 n1d = n1 + 2 * ng, pl = n1d * (n2 + 2 * ng);
  const int i1 = t % n1 + ng, i2 = t / n1 + ng;
  const double* p = phi + i1 + n1d * i2;
  if (order == 4) {
    em[i1 + n1d * i2] = (p[-2] - 8.0 * p[-1] + 8.0 * p[1] - p[2]) / (12.0 * dx);
    em[i1 + n1d * i2 + pl] = (p[-2 * n1d] - 8.0 * p[-n1d] + 8.0 * p[n1d] - p[2 * n1d]) / (12.0 * dy);
  } else {
    em[i1 + n1d * i2] =
        (-1.0 * p[-3] + 9.0 * p[-2] - 45.0 * p[-1] + 45.0 * p[1] - 9.0 * p[2] + 1.0 * p[3]) / (60.0 * dx);
    em[i1 + n1d * i2 + pl] = (-1.0 * p[-3 * n1d] + 9.0 * p[-2 * n1d] - 45.0 * p[-n1d] + 45.0 * p[n1d] -
                              9.0 * p[2 * n1d] + 1.0 * p[3 * n1d]) / (60.0 * dy);
  }
}
cudaError_t efield_from_phi(double* em, const double* phi, int n1, int n2, int ng, int order, double dx, double dy,
                            cudaStream_t st) {
  k_efield<<<nblk((i64)n1 * n2, 128), 128, 0, st>>>(em, phi, n1, n2, ng, order, dx, dy);
  return LK_LAUNCHED();
}

__global__ void k_periodic2d(double* __restrict__ u, int n1, int n2, int ng, int ncomp, int pass) {
  const i64 n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (pass == 0) {
    i64 total = (i64)2 * ng * n2d * ncomp;
    if (t >= total) return;
    int k = (int)(t % (2 * ng));
    i64 row = t / (2 * ng);  // (i2, comp)
    double* p = u + row * n1d;
    if (k < ng) p[k] = p[k + n1]; else p[n1 + k] = p[k];
  } else {
    i64 total = n1d * 2 * ng * ncomp;
    if (t >= total) return;
    int i1 = (int)(t % n1d);
    i64 r = t / n1d;
    int k = (int)(r % (2 * ng));
    int c = (int)(r / (2 * ng));
    double* p = u + (i64)c * n1d * n2d + i1;
    if (k < ng) p[(i64)k * n1d] = p[(i64)(k + n2) * n1d]; else p[(i64)(n2 + k) * n1d] = p[(i64)k * n1d];
  }
}
cudaError_t periodic_fill_2d(double* u, int n1, int n2, int ng, int ncomp, int px, int py, cudaStream_t st) {
  const i64 n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  if (px) { k_periodic2d<<<nblk((i64)2 * ng * n2d * ncomp, 128), 128, 0, st>>>(u, n1, n2, ng, ncomp, 0); ++g_launches; }
  if (py) { k_periodic2d<<<nblk(n1d * 2 * ng * ncomp, 128), 128, 0, st>>>(u, n1, n2, ng, ncomp, 1); ++g_launches; }
  return cudaGetLastError();
}

__global__ void k_xpby2d(double* __restrict__ x, const double* __restrict__ y, double b, int n1, int n2, int ng, int ncomp) {
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (i64)n1 * n2 * ncomp) return;
  const i64 n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  int i1 = (int)(t % n1) + ng;
  i64 r = t / n1;
  int i2 = (int)(r % n2) + ng;
  int c = (int)(r / n2);
  i64 o = i1 + n1d * (i2 + n2d * c);
  x[o] = x[o] + b * y[o];
}
cudaError_t xpby2d(double* x, const double* y, double b, int n1, int n2, int ng, int ncomp, cudaStream_t st) {
  k_xpby2d<<<nblk((i64)n1 * n2 * ncomp, 128), 128, 0, st>>>(x, y, b, n1, n2, ng, ncomp);
  return LK_LAUNCHED();
}

// computeAcceleration glue (KineticSpecies.C:697-755): accel = 0; accel <- E (whole box on one rank);
// accel += driver field; accel *= normalization
__global__ void k_form_accel(double* __restrict__ accel, const double* __restrict__ em, const double* __restrict__ ext,
                             double norm, i64 count) {
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  double v = em[t];
  if (ext) v += ext[t];
  accel[t] = v * norm;
}
cudaError_t form_accel(double* accel, const double* em, const double* ext, double norm, int n1, int n2, int ng,
                       cudaStream_t st) {
  i64 count = (i64)(n1 + 2 * ng) * (n2 + 2 * ng) * 2;
  k_form_accel<<<nblk(count, 128), 128, 0, st>>>(accel, em, ext, norm, count);
  return LK_LAUNCHED();
}

// ---------------------------------------------------------------------------------------------
// a17: maxwellevalrhs (MaxwellF.f:97-355) without supergrid layers (metric nu == 1).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double d1c(const double* p, i64 s, int order, double h) {
  if (order == 4) return (p[-2 * s] - 8.0 * p[-s] + 8.0 * p[s] - p[2 * s]) / (12.0 * h);
  return (-1.0 * p[-3 * s] + 9.0 * p[-2 * s] - 45.0 * p[-s] + 45.0 * p[s] - 9.0 * p[2 * s] + 1.0 * p[3 * s]) / (60.0 * h);
}
__global__ void k_maxwell_rhs(double* __restrict__ rhs, const double* __restrict__ em, const double* __restrict__ Jx,
                              const double* __restrict__ Jy, const double* __restrict__ Jz, int n1, int n2, int ng,
                              int order, double dx, double dy, double c, double avw, double avs) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n1 * n2) return;
  const i64 n1d = n1 + 2 * ng, pl = n1d * (n2 + 2 * ng);
  const int i1 = t % n1 + ng, i2 = t / n1 + ng;
  const i64 o = i1 + n1d * i2;
  const double csq = c * c;
  const double* E1 = em + o; const double* E2 = em + o + pl; const double* E3 = em + o + 2 * pl;
  const double* B1 = em + o + 3 * pl; const double* B2 = em + o + 4 * pl; const double* B3 = em + o + 5 * pl;
  double Exdy = d1c(E1, n1d, order, dy), Eydx = d1c(E2, 1, order, dx);
  double Ezdx = d1c(E3, 1, order, dx), Ezdy = d1c(E3, n1d, order, dy);
  double Bxdy = d1c(B1, n1d, order, dy), Bydx = d1c(B2, 1, order, dx);
  double Bzdx = d1c(B3, 1, order, dx), Bzdy = d1c(B3, n1d, order, dy);
  double r[6];
  r[0] = csq * (Bzdy)-Jx[o];
  r[1] = -csq * (Bzdx)-Jy[o];
  r[2] = csq * (Bydx - Bxdy) - Jz[o];
  r[3] = -Ezdy;
  r[4] = Ezdx;
  r[5] = Exdy - Eydx;
  if (avw > 0.0 || avs > 0.0) {
    for (int k = 0; k < 6; ++k) {
      const double* p = em + o + k * pl;
      if (order == 4) {
        double dx4 = (dx * dx) * (dx * dx), dy4 = (dy * dy) * (dy * dy);
        double uxxxx = (1.0 * p[-2] - 4.0 * p[-1] + 6.0 * p[0] - 4.0 * p[1] + 1.0 * p[2]) / dx4;
        double uyyyy = (1.0 * p[-2 * n1d] - 4.0 * p[-n1d] + 6.0 * p[0] - 4.0 * p[n1d] + 1.0 * p[2 * n1d]) / dy4;
        r[k] = r[k] - (avw * c * dx4 + avs * c * ((dx * dx) * dx)) / 16.0 * uxxxx -
               (avw * c * dy4 + avs * c * ((dy * dy) * dy)) / 16.0 * uyyyy;
      } else {
        double tx = (dx * dx) * dx, ty = (dy * dy) * dy, sx2 = dx * dx, sy2 = dy * dy;
        double dx6 = tx * tx, dy6 = ty * ty, dx5 = (sx2 * dx) * sx2, dy5 = (sy2 * dy) * sy2;
        double ux6 = (1.0 * p[-3] - 6.0 * p[-2] + 15.0 * p[-1] - 20.0 * p[0] + 15.0 * p[1] - 6.0 * p[2] + 1.0 * p[3]) / dx6;
        double uy6 = (1.0 * p[-3 * n1d] - 6.0 * p[-2 * n1d] + 15.0 * p[-n1d] - 20.0 * p[0] + 15.0 * p[n1d] -
                      6.0 * p[2 * n1d] + 1.0 * p[3 * n1d]) / dy6;
        r[k] = r[k] + (avw * c * dx6 + avs * c * dx5) / 64.0 * ux6 + (avw * c * dy6 + avs * c * dy5) / 64.0 * uy6;
      }
    }
  }
  for (int k = 0; k < 6; ++k) rhs[o + k * pl] = r[k];
}
cudaError_t maxwell_rhs(double* rhs, const double* em, const double* Jx, const double* Jy, const double* Jz, int n1,
                        int n2, int ng, int order, double dx, double dy, double c, double av_weak, double av_strong,
                        cudaStream_t st) {
  k_maxwell_rhs<<<nblk((i64)n1 * n2, 128), 128, 0, st>>>(rhs, em, Jx, Jy, Jz, n1, n2, ng, order, dx, dy, c, av_weak, av_strong);
  return LK_LAUNCHED();
}

}  // namespace LK_NS
