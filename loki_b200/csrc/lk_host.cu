// lk_host.cu -- host-side mirror of the reference's operator interface for the Vlasov-Poisson path
// (include/loki_b200_host.h).  Class and method names follow the reference so that the call order of
// VPSystem::evalRHS / RK4Integrator::stageAdvance can be checked line by line; every numerical step is
// a kernel of the C ABI in loki_b200.h (this file contains only 2D glue kernels: tile pack/scatter,
// species sum, driver field, scalar RK state).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <string>
#include <vector>

#include "../../include/loki_b200_host.h"

namespace loki {

typedef long long i64;

#define LKH_CHECK(expr)                 \
  do {                                  \
    int s__ = (expr);                   \
    if (s__ != LK_OK) return s__;       \
  } while (0)
#define LKH_CUDA(expr)                                   \
  do {                                                   \
    cudaError_t e__ = (expr);                            \
    if (e__ != cudaSuccess) return LK_ERR_CUDA;          \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    release();
    if (count == 0) return LK_OK;
    if (cudaMalloc((void**)&p, sizeof(T) * count) != cudaSuccess) { p = nullptr; return LK_ERR_CUDA; }
    n = count;
    return LK_OK;
  }
  int upload(const std::vector<T>& h) {
    LKH_CHECK(alloc(h.size()));
    LKH_CUDA(cudaMemcpy(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    return LK_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

// ---------------------------------------------------------------------------------------------
// 2D glue kernels (explicit round-to-nearest ops: identical in both arithmetic modes)
// ---------------------------------------------------------------------------------------------
__global__ void k_add_inplace(double* __restrict__ x, const double* __restrict__ y, i64 n) {  // ParallelArray::operator+=
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) x[t] = __dadd_rn(x[t], y[t]);
}
// local rho (n1d,n2d with ghosts) -> dense interior tile
__global__ void k_tile_pack(double* __restrict__ tile, const double* __restrict__ rho, int n1, int n2, int ng) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n1 * n2) return;
  tile[t] = rho[(t % n1 + ng) + (i64)(n1 + 2 * ng) * (t / n1 + ng)];
}
// dense tile of rank r -> global rho (ng ghosts, zeroed beforehand)
__global__ void k_tile_scatter(double* __restrict__ rho_g, const double* __restrict__ tile, int lo0, int lo1, int n0,
                               int n1, int n1d_g, int ng) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n0 * n1) return;
  rho_g[(lo0 + t % n0 + ng) + (i64)n1d_g * (lo1 + t / n0 + ng)] = tile[t];
}
// global em_vars (n1d_g,n2d_g,2) -> this rank's (n1d,n2d,2) window incl. ghosts (the expansion schedule)
__global__ void k_extract_window(double* __restrict__ loc, const double* __restrict__ glob, int lo0, int lo1, int n1d,
                                 int n2d, int n1d_g, int n2d_g, int ncomp) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n1d * n2d * ncomp) return;
  int i1 = t % n1d, i2 = (t / n1d) % n2d, c = t / (n1d * n2d);
  loc[t] = glob[(lo0 + i1) + (i64)n1d_g * ((lo1 + i2) + (i64)n2d_g * c)];
}
// ext_efield(i1,i2,0) = 0 + envel*h(i2)*g(i1); component 1 = 0 (evaluateShapedRampedDriver, :172-186)
__global__ void k_driver_field(double* __restrict__ ext, const double* __restrict__ g, const double* __restrict__ h,
                               double envel, int n1d, int n2d) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n1d * n2d) return;
  int i1 = t % n1d, i2 = t / n1d;
  ext[t] = __dmul_rn(__dmul_rn(envel, h[i2]), g[i1]);
  ext[t + n1d * n2d] = 0.0;
}
// scalar RK state m_integrated_ke_e_dot (KineticSpecies.C:282-284): v = {state, old, delta, rhs}
__global__ void k_scalar_rk4(double* v, double w_delta, double c_pred, int first, int use_delta) {
  double d = __dmul_rn(w_delta, v[3]);
  if (!first) d = __dadd_rn(v[2], d);
  v[2] = d;
  v[0] = __dadd_rn(v[1], __dmul_rn(c_pred, use_delta ? d : v[3]));
}
struct Rk6Coef {
  double c[8];
};
__global__ void k_scalar_rk6(double* v, double* k, int stage, Rk6Coef coef, int ncoef) {
  // k[stage] = rhs; state = old + sum coef[j]*k[j]; the coefficients travel as launch arguments (no host copy per stage)
  k[stage] = v[3];
  double p = v[1];
  for (int j = 0; j < ncoef; ++j) p = __dadd_rn(p, __dmul_rn(coef.c[j], k[j]));
  v[0] = p;
}
__global__ void k_copy_scalar(double* v, int dst, int src) { v[dst] = v[src]; }

static inline unsigned nb(i64 n, int t) { return (unsigned)((n + t - 1) / t); }

// ---------------------------------------------------------------------------------------------
// ShapedRampedCosineDriver (ShapedRampedCosineDriverF.f:10-189), separable: E = envel(t) h(y) g(x,t).
// Evaluated on the host with libm exactly as the Fortran does; only the O(Nx)+O(Ny) factors cross
// to the device.
// ---------------------------------------------------------------------------------------------
struct ShapedRampedCosineDriver {
  double p[16];
  double phase;
  int shape_type;
  bool active(double t) const { return p[5] <= t && t < p[5] + p[6] + p[7] + p[8]; }
  double envelope(double t) const {
    const double t0 = p[5], t_rampup = p[6], t_hold = p[7], t_rampdown = p[8], E_0 = p[4];
    if ((t < t0) || (t >= t0 + t_rampup + t_hold + t_rampdown)) return 0.0;
    if (t < t0 + t_rampup) return E_0 * (0.5 + 0.5 * tanh(4.0 * (2.0 * (t - t0) / t_rampup - 1.0)));
    if (t < t0 + t_rampup + t_hold) return E_0 * (0.5 + 0.5 * tanh(4.0));
    return E_0 * (0.5 - 0.5 * tanh(4.0 * (2.0 * (t - t0 - t_rampup - t_hold) / t_rampdown - 1.0)));
  }
  double ghat(double x, double t, double pi) const {
    const double xwidth = p[0], omega = p[3], t0 = p[5], x_shape = p[9], lwidth = p[10], x0 = p[11], alpha = p[12],
                 t_res = p[13];
    double gh;
    if (shape_type == 0) {
      if (fabs(x - x0) < 0.5 * lwidth) {
        double s = sin(pi * (x - x0) / lwidth);
        gh = 1.0 - x_shape * (s * s);
      } else {
        gh = 1.0 - x_shape;
      }
    } else {
      if (lwidth >= 0) {
        gh = (x <= x0) ? 1.0 : 1.0 - x_shape * (1.0 - exp(-(x - x0) / lwidth));
      } else {
        gh = (x <= x0) ? 1.0 - x_shape * (1.0 - exp(-(x - x0) / lwidth)) : 1.0;
      }
    }
    const double tt = t - t0 - t_res;
    return gh * cos(pi * x / xwidth - omega * (t - t0) + phase - 0.5 * alpha * (tt * tt));
  }
  double hfun(double y, double pi) const {
    const double ywidth = p[1], shape = p[2];
    if (fabs(y) < 0.5 * ywidth) {
      double s = sin(pi * y / ywidth);
      return 1.0 - shape * (s * s);
    }
    return 1.0 - shape;
  }
};

// ---------------------------------------------------------------------------------------------
// KineticSpecies (KineticSpecies.H/C): one species' share of phase space on this rank
// ---------------------------------------------------------------------------------------------
struct KineticSpecies {
  lk_geom g;
  double mass, charge, bz_const;
  double vlo[2], vhi[2];
  i64 vol;
  int n1d, n2d, n3d, n4d;
  // RK state: f[0..2] rotate between {state/old, pred A, pred B}; delta; k[0..7] for RK6
  DevBuf<double> farr[3], delta, karr[8], rhs_tmp;
  int i_state = 0, i_a = 1, i_b = 2;
  double* f_eval = nullptr;  // what the next evalRHS reads
  // tables (KineticSpecies.C:1953-2048)
  DevBuf<double> velocities, vxface, vyface;
  // per-stage 2D work arrays
  DevBuf<double> rho_s, accel, ext_efield, lam, drv_g, drv_h, ke;  // ke: {state, old, delta, rhs} + k[8]
  DevBuf<double> ke_k, ke_coef;
  // velocity moments of f_eval left behind by the fused stage kernel (production arithmetic)
  DevBuf<double> mom_part;
  lk_stage_moments mom;
  bool mom_valid = false;
  // Vlasov-Maxwell: per-species current densities (n1d,n2d) and the explicit inflow ghost tables of a
  // non-factorable initial condition (PerturbedMaxwellianIC.C:176-246 caches the full m_f; only its
  // velocity-ghost layers are ever read by setaccelerationbcs4d)
  DevBuf<double> Jx_s, Jy_s, Jz_s, ic_ghost3, ic_ghost4;
  // inflow (initial condition) tables
  DevBuf<double> ic_fx, ic_fv, ic_fx2, ic_fv2;
  lk_inflow inflow;
  bool has_driver = false;
  // periodic x/y ghost cells already written by the fused stage kernel (lk_rk_update::wrap) for this array
  const double* wrap_ptr = nullptr;
  int wrap_bits = 0;
  // periodic wrap of f in the uncut directions `dirs` (bit 0 x, bit 1 y), skipping what the kernel did
  int periodicFill(double* f, int dirs, void* st) {
    const int need = (f == wrap_ptr) ? (dirs & ~wrap_bits) : dirs;
    if (!need) return LK_OK;
    return lk_periodic_fill_4d(f, &g, need & 1, (need >> 1) & 1, st);
  }
  int wrapFor(int dirs) const {
    int w = 0;
    if ((dirs & 1) && g.n[0] >= 2 * g.ng) w |= 1;
    if ((dirs & 2) && g.n[1] >= 2 * g.ng) w |= 2;
    return w;
  }
  ShapedRampedCosineDriver driver;
  double lambda_max[4] = {0, 0, 0, 0};

  double* state() { return farr[i_state].p; }
  // which of the rotating arrays currently hold the inflow sample in their velocity ghost layers
  // (lk_rk_update.inflow_preset: the pipelined stage kernel then needs no separate velocity-boundary fill)
  // deck options beyond the benchmark decks: Krook layer nu(n1d,n2d) (KineticSpecies.C:1049-1062),
  // the "JB" boundary conditions (use_new_bcs, VPSystem.C:819-821), non-periodic x / y (KineticSpecies.H:998-1031)
  DevBuf<double> krook_nu;
  bool has_krook = false, use_new_bcs = false;
  // a pitch-angle collision operator (KineticSpecies.C:1036-1046; PitchAngleCollisionOperator.C): its parameters and
  // the three reduced fields IVx, IVy, IVth (n1d,n2d) every evaluation recomputes
  bool has_coll = false;
  // TrigTZSource (KineticSpecies.C:1077-1080): the manufactured-solution forcing; tables of lk_trig_tz_tables
  bool has_tz = false;
  int tz_kind = 0;   // 0 TrigTZSource, 1 ElectronTrigTZSource, 2 / 3 TwoSpecies_Electron / IonTrigTZSource
  double tz_params[3] = {0.0, 1.0, 1.0};   // the Fortran's dparams: amp, electron_mass, ion_mass
  DevBuf<double> tz_tab;
  lk_pitch_angle coll;
  DevBuf<double> coll_iv;
  // completeRHS's collision term on a materialised rhs (PitchAngleCollisionOperator::evaluate)
  int appendCollision(double* rhs, const double* f, void* st) {
    const size_t pl = (size_t)n1d * n2d;
    if (!coll_iv.p) LKH_CHECK(coll_iv.alloc(3 * pl));
    double* iv = coll_iv.p;
    LKH_CHECK(lk_pitch_angle_fields(iv, iv + pl, iv + 2 * pl, f, &g, velocities.p, st));
    return lk_append_pitch_angle_collision(rhs, f, &g, velocities.p, iv, iv + pl, iv + 2 * pl, vlo, vhi, &coll, st);
  }
  int nonperiodic = 0;              // bit 0 x, bit 1 y
  int at_xy[4] = {1, 1, 1, 1};      // this rank's tile touches x-lo, x-hi, y-lo, y-hi of the domain
  bool preset[3] = {false, false, false};
  // a stage launched for the cut-face tiles only (stageFinish part 1): what part 2 needs to launch the rest
  bool pending = false, pending_mom = false;
  // streaming the state through pinned host memory (lk_vp_upload_next / adopt / download_state)
  cudaEvent_t ev_up = nullptr, ev_down = nullptr;
  int next_idx = -1;               // array an upload of the next state is landing in
  bool down_pending = false;       // a download of the state array is in flight
  cudaEvent_t ev_face = nullptr;   // the face tiles of the last two-part stage are done
  bool face_in_flight = false;     // ... and `st` has not been made to wait for them yet
  lk_rk_update pending_u;
  lk_accel pending_a;
  double* pending_rhs = nullptr;
  const double* pending_f = nullptr;
  int arrayIndex(const double* p) const { return (p == farr[0].p) ? 0 : ((p == farr[1].p) ? 1 : 2); }
  void forgetPresets() { preset[0] = preset[1] = preset[2] = false; }

  lk_accel accelDesc() {
    lk_accel a;
    a.kind = 0;
    a.field = accel.p;
    a.vz = nullptr;
    a.vxface_velocities = vxface.p;
    a.vyface_velocities = vyface.p;
    a.normalization = charge / mass;  // KineticSpecies.C:752-754 (non-relativistic)
    a.bz_const = bz_const;
    return a;
  }

  // buildVelocityArrays, non-relativistic branch (KineticSpecies.C:2024-2047)
  int buildVelocityArrays() {
    const double dvx = g.dx[2], dvy = g.dx[3];
    const int ng = g.ng;
    std::vector<double> v((size_t)n3d * n4d * 2), vxf((size_t)(n3d + 1) * n4d * 2), vyf((size_t)n3d * (n4d + 1) * 2);
    for (int i3 = 0; i3 < n3d; ++i3) {
      double vx = vlo[0] + ((i3 - ng) + 0.5) * dvx;
      for (int i4 = 0; i4 < n4d; ++i4) {
        v[i3 + (size_t)n3d * i4] = vx;
        v[i3 + (size_t)n3d * (i4 + (size_t)n4d)] = vlo[1] + ((i4 - ng) + 0.5) * dvy;
      }
    }
    for (int i3 = 0; i3 <= n3d; ++i3) {
      double vx = vlo[0] + (i3 - ng) * dvx;
      for (int i4 = 0; i4 < n4d; ++i4) {
        vxf[i3 + (size_t)(n3d + 1) * i4] = vx;
        vxf[i3 + (size_t)(n3d + 1) * (i4 + (size_t)n4d)] = vlo[1] + ((i4 - ng) + 0.5) * dvy;
      }
    }
    for (int i3 = 0; i3 < n3d; ++i3) {
      double vx = vlo[0] + ((i3 - ng) + 0.5) * dvx;
      for (int i4 = 0; i4 <= n4d; ++i4) {
        vyf[i3 + (size_t)n3d * i4] = vx;
        vyf[i3 + (size_t)n3d * (i4 + (size_t)(n4d + 1))] = vlo[1] + (i4 - ng) * dvy;
      }
    }
    LKH_CHECK(velocities.upload(v));
    LKH_CHECK(vxface.upload(vxf));
    LKH_CHECK(vyface.upload(vyf));
    // m_lambda_max[X1], [X2] (KineticSpecies.C:1547-1555; the upper value is vhi + dv/2 as written there)
    lambda_max[0] = std::max(fabs(vlo[0] + 0.5 * (vhi[0] - vlo[0]) / g.n[2]), fabs(vhi[0] + 0.5 * (vhi[0] - vlo[0]) / g.n[2]));
    lambda_max[1] = std::max(fabs(vlo[1] + 0.5 * (vhi[1] - vlo[1]) / g.n[3]), fabs(vhi[1] + 0.5 * (vhi[1] - vlo[1]) / g.n[3]));
    return LK_OK;
  }

  // chargeDensity (KineticSpecies.H:265-270; schedule set-up KineticSpecies.C:1736-1753)
  int chargeDensity(const double* f, void* st) {
    // production: the stage kernel that wrote f also summed it over velocity space
    if (mom_valid && f == f_eval && !lk_get_strict())
      return lk_moments_finish(rho_s.p, nullptr, nullptr, &momFirst(), &g, g.dx[2] * g.dx[3], charge, st);
    return lk_reduce_4d_to_2d(rho_s.p, f, &g, g.dx[2] * g.dx[3], charge, st);
  }
  // the charge-density finish only needs moment 0 of the partial buffer
  const lk_stage_moments& momFirst() {
    mom0 = mom;
    mom0.nmom = 1;
    return mom0;
  }
  lk_stage_moments mom0;

  // m_ext_efield = 0.0, then the driver summed into it (KineticSpecies.C:735-751, :2109-2130)
  int evaluateDriver(double time, const double xlo[2], const int tile_lo[2], void* st) {
    LKH_CUDA(cudaMemsetAsync(ext_efield.p, 0, sizeof(double) * 2 * n1d * n2d, (cudaStream_t)st));  // m_ext_efield = 0.0
    if (driver.active(time)) {
      const double pi = 4.0 * atan(1.0);
      const double envel = driver.envelope(time);
      std::vector<double> gh(n1d), hh(n2d);
      for (int i1 = 0; i1 < n1d; ++i1) gh[i1] = driver.ghat(xlo[0] + g.dx[0] * (0.5 + (tile_lo[0] + i1 - g.ng)), time, pi);
      for (int i2 = 0; i2 < n2d; ++i2) hh[i2] = driver.hfun(xlo[1] + g.dx[1] * (0.5 + (tile_lo[1] + i2 - g.ng)), pi);
      // stream-ordered copies from pageable memory are synchronous with respect to the host buffer
      LKH_CUDA(cudaMemcpyAsync(drv_g.p, gh.data(), sizeof(double) * n1d, cudaMemcpyHostToDevice, (cudaStream_t)st));
      LKH_CUDA(cudaMemcpyAsync(drv_h.p, hh.data(), sizeof(double) * n2d, cudaMemcpyHostToDevice, (cudaStream_t)st));
      k_driver_field<<<nb((i64)n1d * n2d, 128), 128, 0, (cudaStream_t)st>>>(ext_efield.p, drv_g.p, drv_h.p, envel, n1d, n2d);
    }
    return LK_OK;
  }

  // computeAcceleration (KineticSpecies.C:697-774).  em_local: this rank's window of E incl. ghosts.
  int computeAcceleration(const double* em_local, double time, const double xlo[2], const int tile_lo[2], bool want_max,
                          void* st) {
    const double* ext = nullptr;
    if (has_driver) {
      LKH_CHECK(evaluateDriver(time, xlo, tile_lo, st));
      ext = ext_efield.p;
    }
    // m_accel = 0; expansion of E; drivers add; m_accel *= normalization
    LKH_CHECK(lk_form_accel(accel.p, em_local, ext, charge / mass, g.n[0], g.n[1], g.ng, st));
    if (want_max) {
      lk_accel a = accelDesc();
      LKH_CHECK(lk_max_accel(&g, &a, lam.p, st));  // axmax, aymax -> m_lambda_max[V1], [V2]
    }
    return LK_OK;
  }

  // setAccelerationBCs (KineticSpecies.H:421-453): velocity space is whole on every rank
  int setAccelerationBCs(double* f, void* st) {
    preset[arrayIndex(f)] = false;
    lk_accel a = accelDesc();
    const int at[4] = {1, 1, 1, 1};
    if (use_new_bcs) return lk_set_acceleration_bcs_4d_jb(f, &g, &a, &inflow, at, st);
    return lk_set_acceleration_bcs_4d(f, &g, &a, &inflow, at, st);
  }
  // fillAdvectionGhostCells on this rank (KineticSpecies.H:404-412, 998-1031): the physical boundary conditions of a
  // non-periodic direction, then the periodic wrap of the periodic directions this rank is not cut in (`dirs`)
  int fillAdvectionGhosts(double* f, int dirs, void* st) {
    if (nonperiodic) {
      const int xper = !(nonperiodic & 1), yper = !(nonperiodic & 2);
      if (use_new_bcs) {
        LKH_CHECK(lk_set_advection_bcs_4d_jb(f, &g, velocities.p, &inflow, at_xy, xper, yper, st));
      } else {
        LKH_CHECK(lk_set_advection_bcs_4d(f, &g, velocities.p, &inflow, at_xy, xper, yper, st));
      }
    }
    return periodicFill(f, dirs & ~nonperiodic, st);
  }

  // computeDt (KineticSpecies.C:647-694); the real eigenvalue is the collision operator's (:666-672)
  double computeDt(int rk_order) const {
    const double pi = 4.0 * atan(1.0);
    double imLam = 0.0, reLam = 0.0;
    for (int dir = 0; dir < 4; ++dir) imLam += pi * lambda_max[dir] / g.dx[dir];
    if (has_coll) {
      const double thisReLam = lk_pitch_angle_real_lam(&g, &coll);
      if (fabs(thisReLam) > reLam) reLam = thisReLam;
    }
    double alpha = (rk_order == 4) ? 2.6 : 4.95, beta = (rk_order == 4) ? 2.6 : 3.168;
    return sqrt(1.0 / (reLam * reLam / (alpha * alpha) + imLam * imLam / (beta * beta)));
  }
};

// RK6Integrator's tableau (RK6Integrator.H:77-103), shared by the Vlasov-Poisson and Vlasov-Maxwell stage loops
namespace rk6tab {
static const double A[8][8] = {
    {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {1.0 / 9.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {1.0 / 24.0, 1.0 / 8.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {1.0 / 6.0, -1.0 / 2.0, 2.0 / 3.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {935.0 / 2536.0, -2781.0 / 2536.0, 309.0 / 317.0, 321.0 / 1268.0, 0.0, 0.0, 0.0, 0.0},
    {-12710.0 / 951.0, 8287.0 / 317.0, -40.0 / 317.0, -6335.0 / 317.0, 8.0, 0.0, 0.0, 0.0},
    {5840285.0 / 3104064.0, -7019.0 / 2536.0, -52213.0 / 86224.0, 1278709.0 / 517344.0, -433.0 / 2448.0,
     33.0 / 1088.0, 0.0, 0.0},
    {-5101675.0 / 1767592.0, 112077.0 / 25994.0, 334875.0 / 441898.0, -973617.0 / 883796.0, -1421.0 / 1394.0,
     333.0 / 5576.0, 36.0 / 41.0, 0.0}};
static const double b[8] = {41.0 / 840.0, 0.0, 9.0 / 35.0, 9.0 / 280.0, 34.0 / 105.0, 9.0 / 280.0, 9.0 / 35.0, 41 / 840.0};
static const double c[8] = {0.0, 1.0 / 9.0, 1.0 / 6.0, 1.0 / 3.0, 1.0 / 2.0, 2.0 / 3.0, 5.0 / 6.0, 1.0};
}  // namespace rk6tab

// ---------------------------------------------------------------------------------------------
// VPSystem (+ VPState, Poisson, the RK integrators): one rank
// ---------------------------------------------------------------------------------------------
struct VPSystem {
  lk_vp_desc desc;
  std::vector<lk_species_desc> sdesc;
  std::vector<KineticSpecies*> species;
  cudaStream_t st = nullptr;
  int ng = 2;
  int n1d_g = 0, n2d_g = 0;     // global 2D extents incl. ghosts
  double dxg[4] = {0, 0, 1, 1};
  lk_poisson_plan* poisson = nullptr;
  DevBuf<double> rho_local, rho_tile_own, rho_gather_own, rho_g, phi_g, em_g, em_local;
  double* rho_tile_p = nullptr;
  double* rho_gather_p = nullptr;
  size_t rho_tile_n = 0;
  double time = 0.0, dt = 0.0;
  bool lambda_stale = true;

  // second stream for the cut-face tiles of a two-part stage (stageFinish part 1): the remaining tiles are launched on
  // `st` without waiting for them, so the two launches share the SMs with no idle tail between them
  cudaStream_t st_face = nullptr;
  cudaEvent_t ev_pre = nullptr;
  DevBuf<double> hist_scratch, flux_scratch;   // device side of lk_vp_time_history / lk_vp_flux_history
  ~VPSystem() {
    for (auto* s : species) {
      if (s->ev_face) cudaEventDestroy(s->ev_face);
      if (s->ev_up) cudaEventDestroy(s->ev_up);
      if (s->ev_down) cudaEventDestroy(s->ev_down);
      delete s;
    }
    if (poisson) lk_poisson_plan_destroy(poisson);
    if (st_face) cudaStreamDestroy(st_face);
    if (ev_pre) cudaEventDestroy(ev_pre);
  }
  int faceStream() {
    if (st_face) return LK_OK;
    int lo = 0, hi = 0;
    LKH_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LKH_CUDA(cudaStreamCreateWithPriority(&st_face, cudaStreamNonBlocking, hi));
    LKH_CUDA(cudaEventCreateWithFlags(&ev_pre, cudaEventDisableTiming));
    return LK_OK;
  }

  int nstages() const { return desc.rk_order == 4 ? 4 : 8; }

  int create(const lk_vp_desc* d, void* stream) {
    desc = *d;
    sdesc.assign(d->species, d->species + d->nspecies);
    desc.species = sdesc.data();
    st = (cudaStream_t)stream;
    if (!(d->order == 4 || d->order == 6) || !(d->rk_order == 4 || d->rk_order == 6) || d->nspecies < 1) return LK_ERR_ARG;
    ng = d->order == 4 ? 2 : 3;
    for (int k = 0; k < 2; ++k) {
      if (d->nglobal[k] < 1 || d->tile_n[k] < ng || d->tile_lo[k] < 0 || d->tile_lo[k] + d->tile_n[k] > d->nglobal[k]) return LK_ERR_ARG;
      dxg[k] = (d->xhi[k] - d->xlo[k]) / d->nglobal[k];  // ProblemDomain.C:20-27
    }
    n1d_g = d->nglobal[0] + 2 * ng;
    n2d_g = d->nglobal[1] + 2 * ng;
    LKH_CHECK(lk_poisson_plan_create(&poisson, d->nglobal[0], d->nglobal[1], ng, d->order, d->xhi[0] - d->xlo[0],
                                     d->xhi[1] - d->xlo[1]));
    const int n1d = d->tile_n[0] + 2 * ng, n2d = d->tile_n[1] + 2 * ng;
    LKH_CHECK(rho_local.alloc((size_t)n1d * n2d));
    rho_tile_n = (size_t)d->tile_n[0] * d->tile_n[1];
    LKH_CHECK(rho_tile_own.alloc(rho_tile_n));
    LKH_CHECK(rho_gather_own.alloc((size_t)d->nglobal[0] * d->nglobal[1]));
    rho_tile_p = rho_tile_own.p;
    rho_gather_p = rho_gather_own.p;
    LKH_CHECK(rho_g.alloc((size_t)n1d_g * n2d_g));
    LKH_CHECK(phi_g.alloc((size_t)n1d_g * n2d_g));
    LKH_CHECK(em_g.alloc((size_t)n1d_g * n2d_g * 2));
    LKH_CHECK(em_local.alloc((size_t)n1d * n2d * 2));
    LKH_CUDA(cudaMemset(rho_g.p, 0, sizeof(double) * rho_g.n));
    for (int s = 0; s < d->nspecies; ++s) {
      const lk_species_desc& sd = sdesc[s];
      KineticSpecies* ks = new KineticSpecies();
      species.push_back(ks);
      ks->g.n[0] = d->tile_n[0]; ks->g.n[1] = d->tile_n[1]; ks->g.n[2] = sd.nv[0]; ks->g.n[3] = sd.nv[1];
      ks->g.ng = ng; ks->g.order = d->order;
      ks->g.dx[0] = dxg[0]; ks->g.dx[1] = dxg[1];
      ks->g.dx[2] = (sd.vhi[0] - sd.vlo[0]) / sd.nv[0];
      ks->g.dx[3] = (sd.vhi[1] - sd.vlo[1]) / sd.nv[1];
      ks->mass = sd.mass; ks->charge = sd.charge; ks->bz_const = sd.bz_const;
      ks->vlo[0] = sd.vlo[0]; ks->vlo[1] = sd.vlo[1]; ks->vhi[0] = sd.vhi[0]; ks->vhi[1] = sd.vhi[1];
      ks->n1d = n1d; ks->n2d = n2d; ks->n3d = sd.nv[0] + 2 * ng; ks->n4d = sd.nv[1] + 2 * ng;
      ks->vol = (i64)n1d * n2d * ks->n3d * ks->n4d;
      for (int k = 0; k < 3; ++k) LKH_CHECK(ks->farr[k].alloc(ks->vol));
      if (d->rk_order == 4) {
        LKH_CHECK(ks->delta.alloc(ks->vol));
      } else {
        for (int k = 0; k < 8; ++k) LKH_CHECK(ks->karr[k].alloc(ks->vol));
      }
      LKH_CUDA(cudaMemset(ks->farr[0].p, 0, sizeof(double) * ks->vol));
      LKH_CHECK(ks->buildVelocityArrays());
      LKH_CHECK(ks->rho_s.alloc((size_t)n1d * n2d));
      LKH_CHECK(ks->accel.alloc((size_t)n1d * n2d * 2));
      LKH_CHECK(ks->lam.alloc(2));
      LKH_CHECK(ks->ke.alloc(4));
      LKH_CHECK(ks->ke_k.alloc(8));
      LKH_CHECK(ks->ke_coef.alloc(8));
      LKH_CUDA(cudaMemset(ks->ke.p, 0, sizeof(double) * 4));
      LKH_CUDA(cudaMemset(ks->lam.p, 0, sizeof(double) * 2));
      memset(&ks->inflow, 0, sizeof(ks->inflow));
      ks->has_driver = sd.has_driver != 0;
      {
        const int nmom = ks->has_driver ? 3 : 1;
        const int parts = lk_stage_moment_parts(&ks->g);
        if (parts < 1) return LK_ERR_ARG;
        const size_t cap = (size_t)nmom * parts * d->tile_n[0] * d->tile_n[1];
        LKH_CHECK(ks->mom_part.alloc(cap));
        ks->mom.nmom = nmom;
        ks->mom.partial = ks->mom_part.p;
        ks->mom.capacity = (int64_t)cap;
      }
      if (ks->has_driver) {
        memcpy(ks->driver.p, sd.driver, sizeof(double) * 16);
        ks->driver.phase = sd.driver_phase;
        ks->driver.shape_type = sd.driver_shape_type;
        LKH_CHECK(ks->ext_efield.alloc((size_t)n1d * n2d * 2));
        LKH_CHECK(ks->drv_g.alloc(n1d));
        LKH_CHECK(ks->drv_h.alloc(n2d));
      }
      ks->f_eval = ks->state();
    }
    return LK_OK;
  }

  // ---- stage pieces, in the order of VPSystem::evalRHS (VPSystem.C:372-476) ----
  // (1) chargeDensity of every species, summed into the local net charge density
  int momentsOf(bool use_eval) {
    for (size_t s = 0; s < species.size(); ++s) {
      KineticSpecies* ks = species[s];
      LKH_CHECK(ks->chargeDensity(use_eval ? ks->f_eval : ks->state(), st));
      if (s == 0) {
        // m_net_charge_density = 0.0; += first species  (0 + x == x)
        LKH_CUDA(cudaMemcpyAsync(rho_local.p, ks->rho_s.p, sizeof(double) * rho_local.n, cudaMemcpyDeviceToDevice, st));
      } else {
        k_add_inplace<<<nb((i64)rho_local.n, 128), 128, 0, st>>>(rho_local.p, ks->rho_s.p, (i64)rho_local.n);
      }
    }
    k_tile_pack<<<nb((i64)rho_tile_n, 128), 128, 0, st>>>(rho_tile_p, rho_local.p, desc.tile_n[0], desc.tile_n[1], ng);
    return LK_OK;
  }
  // (2) global charge density from the gathered tiles, EMSolverBase::electricField, expansion to the tile
  int fieldSolve(const int* tiles) {
    if (tiles == nullptr) {
      const int one[4] = {desc.tile_lo[0], desc.tile_lo[1], desc.tile_n[0], desc.tile_n[1]};
      k_tile_scatter<<<nb((i64)rho_tile_n, 128), 128, 0, st>>>(rho_g.p, rho_tile_p, one[0], one[1], one[2], one[3], n1d_g, ng);
    } else {
      size_t off = 0;
      for (int r = 0; r < desc.ntiles; ++r) {
        const int* t = tiles + 4 * r;
        k_tile_scatter<<<nb((i64)t[2] * t[3], 128), 128, 0, st>>>(rho_g.p, rho_gather_p + off, t[0], t[1], t[2], t[3], n1d_g, ng);
        off += (size_t)t[2] * t[3];
      }
    }
    LKH_CHECK(lk_electric_field(poisson, rho_g.p, phi_g.p, em_g.p, dxg, st));
    const int n1d = desc.tile_n[0] + 2 * ng, n2d = desc.tile_n[1] + 2 * ng;
    k_extract_window<<<nb((i64)n1d * n2d * 2, 128), 128, 0, st>>>(em_local.p, em_g.p, desc.tile_lo[0], desc.tile_lo[1], n1d, n2d,
                                                                n1d_g, n2d_g, 2);
    return LK_OK;
  }
  // (3a) fillAdvectionGhostCells on one rank: periodic wrap (the multi-rank exchange is the caller's)
  int uncutDirs() const {
    return (desc.tile_n[0] == desc.nglobal[0] ? 1 : 0) | (desc.tile_n[1] == desc.nglobal[1] ? 2 : 0);
  }
  int fillAdvectionGhostCellsLocal() {
    for (auto* ks : species) LKH_CHECK(ks->fillAdvectionGhosts(ks->f_eval, 3, st));
    return LK_OK;
  }

  // ---- RK4Integrator::stageAdvance / RK6 stage, fused behind the RHS evaluation ----
  // `only`: restrict to one species (the multi-rank driver interleaves the species' halo exchanges with
  // the other species' stage kernels); nullptr = all species in order
  // `part`: 0 = the whole stage.  1 / 2 = the same in two launches for a rank whose configuration space is cut: 1 does
  // everything but restricts the stage kernel to the tiles on a cut face (lk_rk_update.tile_set), so that the caller
  // can start the halo exchange of the new predictor; 2 launches the remaining tiles.  When the kernel cannot split
  // (not the pipelined instantiation) part 1 is the whole stage and part 2 does nothing.
  int stageFinish(int stage, KineticSpecies* only = nullptr, int part = 0) {
    if (part == 2) {
      for (auto* ks : species) {
        if ((only && ks != only) || !ks->pending) continue;
        ks->pending = false;
        ks->pending_u.tile_set = 2;
        LKH_CHECK(lk_vlasov_stage(ks->pending_rhs, ks->pending_f, &ks->g, ks->velocities.p, &ks->pending_a, &ks->pending_u,
                                  ks->pending_mom ? &ks->mom : nullptr, st));
        // everything queued on the main stream from here on sees the whole stage
        LKH_CUDA(cudaStreamWaitEvent(st, ks->ev_face, 0));
        ks->face_in_flight = false;
      }
      return LK_OK;
    }
    const bool rk4 = desc.rk_order == 4;
    const int last = nstages() - 1;
    const auto& A6 = rk6tab::A;
    const double* b6 = rk6tab::b;
    const double* c6 = rk6tab::c;
    // stage time (RK4Integrator.H:84-120, RK6Integrator.H:109-121)
    double t_stage;
    if (rk4) {
      const double dtOn2 = 0.5 * dt;
      t_stage = (stage == 0) ? time : (stage == 3 ? time + dt : time + dtOn2);
    } else {
      t_stage = time + c6[stage] * dt;
    }
    for (auto* ks : species) {
      if (only && ks != only) continue;
      if (ks->pending) return LK_ERR_ARG;  // part 2 of the previous two-part stage of this species was never issued
      // (4) acceleration; the maxima are only consumed by the next stableDt -> last stage
      LKH_CHECK(ks->computeAcceleration(em_local.p, t_stage, desc.xlo, desc.tile_lo, stage == last, st));
      // (5) velocity-boundary fill + advection + acceleration derivatives + RK update in one pass: the stage does
      // setAccelerationBCs itself (folded into the pipelined kernel's boundary tiles, or a separate fill first)
      lk_accel a = ks->accelDesc();
      lk_rk_update u;
      memset(&u, 0, sizeof(u));
      // the "JB" fill and a Krook-layer species take the separate passes below
      const bool plain = !ks->use_new_bcs && !ks->has_krook && !ks->has_coll && !ks->has_tz;
      static const bool no_fold = getenv("LOKI_NO_FOLD") != nullptr;  // A/B aid: the separate fill before every stage
      if (plain) {
        u.accel_bcs = &ks->inflow;
        u.inflow_preset = no_fold ? 0 : 1;
      } else {
        LKH_CHECK(ks->setAccelerationBCs(ks->f_eval, st));
      }
      double* pred = (ks->f_eval == ks->farr[ks->i_a].p) ? ks->farr[ks->i_b].p : ks->farr[ks->i_a].p;
      u.f_old = ks->state();
      u.pred = pred;
      double* rhs_out = nullptr;
      double ke_coef[8];
      int ke_ncoef = 0;
      if (rk4) {
        static const double THIRD = 1.0 / 3.0;
        const double dtOn2 = 0.5 * dt, dtOn3 = THIRD * dt, dtOn6 = 0.5 * dtOn3;
        const double w_eval[4] = {dtOn6, dtOn3, dtOn3, dtOn6};
        const double w_upd[4] = {dtOn2, dtOn2, dt, 1.0};
        u.delta_in = (stage == 0) ? nullptr : ks->delta.p;   // zeroSolnData(m_delta)
        u.delta_out = (stage == 3) ? nullptr : ks->delta.p;  // nothing reads m_delta after the last stage
        u.w_delta = w_eval[stage];
        u.c_pred = w_upd[stage];
        u.use_delta = (stage == 3);
      } else {
        rhs_out = ks->karr[stage].p;  // m_k[i]
        const double* coef = (stage == last) ? b6 : A6[stage + 1];
        int np = 0;
        for (int j = 0; j < stage; ++j) {
          ke_coef[ke_ncoef++] = dt * coef[j];
          u.k_prev[np] = ks->karr[j].p;
          u.c_prev[np] = dt * coef[j];
          ++np;
        }
        u.n_prev = np;
        u.c_pred = dt * coef[stage];
        ke_coef[ke_ncoef++] = dt * coef[stage];
      }
      static const bool no_fuse = getenv("LK_NO_FUSED_MOMENTS") != nullptr;  // debugging aid
      const bool fused_moments = !lk_get_strict() && !no_fuse;
      // completeRHS: the driver's energy input rate, integrated with the state (KineticSpecies.C:1084-1093).
      // Production: from the vx moment of f_eval that the previous stage kernel left behind (consumed
      // here, before this stage's kernel overwrites the partial buffer).
      if (ks->has_driver) {
        if (fused_moments && ks->mom_valid)
          LKH_CHECK(lk_ke_e_dot_from_moments(ks->ke.p + 3, &ks->mom, &ks->g, ks->charge, ks->ext_efield.p, st));
        else
          LKH_CHECK(lk_ke_e_dot(ks->ke.p + 3, ks->f_eval, &ks->g, ks->charge, ks->velocities.p, ks->ext_efield.p, st));
      }
      // production: the kernel also writes pred's periodic ghost copies in the directions this rank wraps itself
      u.wrap = fused_moments ? ks->wrapFor(uncutDirs() & ~ks->nonperiodic) : 0;
      static const bool krook_passes = getenv("LOKI_KROOK_PASSES") != nullptr;  // debugging aid: the three-pass form
      const bool passes = ks->has_coll || ks->has_tz || (ks->has_krook && krook_passes);  // the rhs is materialised between the passes
      if (passes) u.wrap = 0;  // lk_rk_stage_update writes interior cells only: the next fill wraps pred itself
      if (ks->has_krook && !passes) {
        // completeRHS's Krook layer (KineticSpecies.C:1049-1062) inside the fused stage: the per-cell epilogue of the
        // generic kernel subtracts nu/dt (f - f_IC) from the rhs before the update (and before m_k[stage] is stored)
        u.krook_nu = ks->krook_nu.p;
        u.krook_dt = dt;
        u.krook_ic = &ks->inflow;
        LKH_CHECK(lk_vlasov_stage(rhs_out, ks->f_eval, &ks->g, ks->velocities.p, &a, &u, fused_moments ? &ks->mom : nullptr, st));
      } else if (passes) {
        // completeRHS on a materialised rhs (RK6: in m_k[stage], RK4: in a scratch array): collision operator, then the
        // Krook layer (KineticSpecies.C:1036-1062), then the update alone
        if (!rhs_out) {
          if (!ks->rhs_tmp.p) LKH_CHECK(ks->rhs_tmp.alloc(ks->vol));
          rhs_out = ks->rhs_tmp.p;
        }
        LKH_CHECK(lk_vlasov_rhs(rhs_out, ks->f_eval, &ks->g, ks->velocities.p, &a, nullptr, st));
        if (ks->has_coll) LKH_CHECK(ks->appendCollision(rhs_out, ks->f_eval, st));
        if (ks->has_krook) LKH_CHECK(lk_append_krook(rhs_out, ks->f_eval, &ks->g, ks->krook_nu.p, dt, &ks->inflow, st));
        if (ks->has_tz) LKH_CHECK(lk_set_trig_tz_source(rhs_out, &ks->g, ks->tz_tab.p, ks->velocities.p, t_stage, ks->tz_kind, ks->tz_params, st));
        LKH_CHECK(lk_rk_stage_update(rhs_out, &ks->g, &u, st));
      } else {
        const int ie = ks->arrayIndex(ks->f_eval);
        if (plain && lk_vlasov_stage_folds_bcs(rhs_out, &ks->g, &a, &u)) {
          if (!ks->preset[ie]) LKH_CHECK(lk_preset_inflow_ghosts_4d(ks->f_eval, &ks->g, &ks->inflow, st));
          ks->preset[ie] = true;
        } else {
          u.inflow_preset = 0;     // the stage runs the separate fill: f_eval's ghosts get the extrapolations as well
          ks->preset[ie] = false;
        }
        const int cut = 3 & ~uncutDirs();
        if (part == 1 && cut && lk_vlasov_stage_can_split(rhs_out, &ks->g, &a, &u)) {
          u.tile_set = 1;
          u.cut_dirs = cut;
          ks->pending = true;
          ks->pending_u = u;
          ks->pending_a = a;
          ks->pending_rhs = rhs_out;
          ks->pending_f = ks->f_eval;
          ks->pending_mom = fused_moments;
          // the face tiles go to the high-priority stream, behind everything the main stream holds now
          LKH_CHECK(faceStream());
          if (!ks->ev_face) LKH_CUDA(cudaEventCreateWithFlags(&ks->ev_face, cudaEventDisableTiming));
          LKH_CUDA(cudaEventRecord(ev_pre, st));
          LKH_CUDA(cudaStreamWaitEvent(st_face, ev_pre, 0));
          LKH_CHECK(lk_vlasov_stage(rhs_out, ks->f_eval, &ks->g, ks->velocities.p, &a, &u, fused_moments ? &ks->mom : nullptr, st_face));
          LKH_CUDA(cudaEventRecord(ks->ev_face, st_face));
          ks->face_in_flight = true;
        } else {
          LKH_CHECK(lk_vlasov_stage(rhs_out, ks->f_eval, &ks->g, ks->velocities.p, &a, &u, fused_moments ? &ks->mom : nullptr, st));
        }
      }
      ks->wrap_ptr = pred;
      ks->wrap_bits = u.wrap;
      ks->mom_valid = fused_moments && !passes;  // moments of `pred`, the next stage's input
      if (ks->has_driver) {
        if (rk4) {
          static const double THIRD = 1.0 / 3.0;
          const double dtOn2 = 0.5 * dt, dtOn3 = THIRD * dt, dtOn6 = 0.5 * dtOn3;
          const double w_eval[4] = {dtOn6, dtOn3, dtOn3, dtOn6};
          const double w_upd[4] = {dtOn2, dtOn2, dt, 1.0};
          k_scalar_rk4<<<1, 1, 0, st>>>(ks->ke.p, w_eval[stage], w_upd[stage], stage == 0, stage == 3);
        } else {
          Rk6Coef cf;
          for (int j = 0; j < 8; ++j) cf.c[j] = (j < ke_ncoef) ? ke_coef[j] : 0.0;
          k_scalar_rk6<<<1, 1, 0, st>>>(ks->ke.p, ks->ke_k.p, stage, cf, ke_ncoef);
        }
      }
      ks->f_eval = pred;  // a_evalSoln of the next stage is the predictor
    }
    lambda_stale = true;
    return LK_OK;
  }

  int beginStep(double a_dt) {
    dt = a_dt;
    for (auto* ks : species)
      if (ks->next_idx >= 0) return LK_ERR_ARG;  // an uploaded state waits for lk_vp_adopt_next
    for (auto* ks : species) {
      if (ks->down_pending) {  // the array a download still reads becomes a predictor buffer of this step
        LKH_CUDA(cudaStreamWaitEvent(st, ks->ev_down, 0));
        ks->down_pending = false;
      }
      ks->f_eval = ks->state();  // stage 1 evaluates the old solution
      if (ks->has_driver) k_copy_scalar<<<1, 1, 0, st>>>(ks->ke.p, 1, 0);  // copySolnData(old, state)
    }
    return LK_OK;
  }
  int endStep() {
    // the last predictor is the new state; the previous state's buffer becomes a free predictor buffer
    for (auto* ks : species) {
      int i_new = (ks->f_eval == ks->farr[ks->i_a].p) ? ks->i_a : ks->i_b;
      int i_other = (i_new == ks->i_a) ? ks->i_b : ks->i_a;
      int i_old = ks->i_state;
      ks->i_state = i_new;
      ks->i_a = i_old;
      ks->i_b = i_other;
      ks->f_eval = ks->state();
    }
    time += dt;
    return LK_OK;
  }
  int advance(double a_dt) {
    if (desc.ntiles != 1) return LK_ERR_UNSUPPORTED;
    LKH_CHECK(beginStep(a_dt));
    for (int stage = 0; stage < nstages(); ++stage) {
      LKH_CHECK(momentsOf(true));
      LKH_CHECK(fieldSolve(nullptr));
      LKH_CHECK(fillAdvectionGhostCellsLocal());
      LKH_CHECK(stageFinish(stage));
    }
    return endStep();
  }

  int refreshLambda() {
    if (!lambda_stale) return LK_OK;
    LKH_CUDA(cudaStreamSynchronize(st));
    for (auto* ks : species) {
      double l[2];
      LKH_CUDA(cudaMemcpy(l, ks->lam.p, sizeof(l), cudaMemcpyDeviceToHost));
      ks->lambda_max[2] = l[0];
      ks->lambda_max[3] = l[1];
    }
    lambda_stale = false;
    return LK_OK;
  }

  // VPSystem::evalRHS in the reference's unfused order (parity hook); single rank
  int evalRHS(double** rhs_dev, double t) {
    if (desc.ntiles != 1) return LK_ERR_UNSUPPORTED;
    LKH_CHECK(momentsOf(false));
    LKH_CHECK(fieldSolve(nullptr));
    for (size_t s = 0; s < species.size(); ++s) {
      KineticSpecies* ks = species[s];
      double* f = ks->state();
      ks->mom_valid = false;
      ks->wrap_ptr = nullptr;
      LKH_CHECK(ks->fillAdvectionGhosts(f, 3, st));
      LKH_CHECK(lk_advection_derivatives_4d(rhs_dev[s], f, &ks->g, ks->velocities.p, st));
      LKH_CHECK(ks->computeAcceleration(em_local.p, t, desc.xlo, desc.tile_lo, true, st));
      LKH_CHECK(ks->setAccelerationBCs(f, st));
      lk_accel a = ks->accelDesc();
      LKH_CHECK(lk_acceleration_derivatives_4d(rhs_dev[s], f, &ks->g, &a, st));
      if (ks->has_coll) LKH_CHECK(ks->appendCollision(rhs_dev[s], f, st));
      if (ks->has_krook) LKH_CHECK(lk_append_krook(rhs_dev[s], f, &ks->g, ks->krook_nu.p, dt, &ks->inflow, st));
      if (ks->has_tz) LKH_CHECK(lk_set_trig_tz_source(rhs_dev[s], &ks->g, ks->tz_tab.p, ks->velocities.p, t, ks->tz_kind, ks->tz_params, st));
      if (ks->has_driver)
        LKH_CHECK(lk_ke_e_dot(ks->ke.p + 3, f, &ks->g, ks->charge, ks->velocities.p, ks->ext_efield.p, st));
    }
    lambda_stale = true;
    return LK_OK;
  }
};

// ---------------------------------------------------------------------------------------------
// VMSystem (+ VMState, Maxwell, RK4Integrator): one rank, whole configuration space
// (VMSystem.C:407-581, Maxwell.C:299-353, 562-623, Maxwell.H:199-204, 371-381)
// ---------------------------------------------------------------------------------------------
__global__ void k_mul2d(double* __restrict__ dst, const double* __restrict__ a, const double* __restrict__ b, i64 n) {
  i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[t] = __dmul_rn(a[t], b[t]);
}
// maxwellevalvzrhs (MaxwellF.f:442-469): dvz = (q/m) Ez on the interior
__global__ void k_vz_rhs(double* __restrict__ dvz, const double* __restrict__ em, double qm, int n1, int n2, int ng) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n1 * n2) return;
  const i64 n1d = n1 + 2 * ng, pl = n1d * (n2 + 2 * ng);
  const i64 o = (t % n1 + ng) + n1d * (t / n1 + ng);
  dvz[o] = __dmul_rn(qm, em[o + 2 * pl]);
}

extern "C" int lk_maxwell_vz_rhs(double* dvz, const double* em, int n1, int n2, int ng, double charge_per_mass, void* stream) {
  if (!dvz || !em || n1 < 1 || n2 < 1 || ng < 0) return LK_ERR_ARG;
  k_vz_rhs<<<nb((i64)n1 * n2, 128), 128, 0, (cudaStream_t)stream>>>(dvz, em, charge_per_mass, n1, n2, ng);
  return cudaGetLastError() == cudaSuccess ? LK_OK : LK_ERR_CUDA;
}

struct VMSystem {
  lk_vm_desc desc;
  std::vector<lk_species_desc> sdesc;
  std::vector<KineticSpecies*> species;
  DevBuf<double> hist_scratch;   // device side of lk_vm_time_history
  cudaStream_t st = nullptr;
  int ng = 2, n1 = 0, n2 = 0, n1d = 0, n2d = 0;
  i64 pl = 0;
  double dxg[2] = {0, 0};
  // Maxwell state: em_vars (n1d,n2d,6) and vz per species; [0] = state/old, [1] = new (predictor)
  DevBuf<double> em[2], em_rhs, em_delta, Jx, Jy, Jz, sumf;
  std::vector<DevBuf<double>*> vz[2], vz_rhs, vz_delta;
  // RK6Integrator's m_k[8] of the Maxwell part of the state (RK6Integrator.H:135-150); species keep theirs in karr
  DevBuf<double> em_k[8];
  std::vector<DevBuf<double>*> vz_k[8];
  int i_state = 0;
  double time = 0.0, dt = 0.0;
  bool lambda_stale = true;

  ~VMSystem() {
    for (auto* s : species) delete s;
    for (int k = 0; k < 2; ++k) for (auto* b : vz[k]) delete b;
    for (auto* b : vz_rhs) delete b;
    for (auto* b : vz_delta) delete b;
    for (int k = 0; k < 8; ++k) for (auto* b : vz_k[k]) delete b;
  }
  double* emState() { return em[i_state].p; }
  double* emNew() { return em[1 - i_state].p; }
  double* vzState(int s) { return vz[i_state][s]->p; }
  double* vzNew(int s) { return vz[1 - i_state][s]->p; }

  int create(const lk_vm_desc* d, void* stream) {
    desc = *d;
    const lk_vp_desc& b = d->base;
    sdesc.assign(b.species, b.species + b.nspecies);
    desc.base.species = sdesc.data();
    st = (cudaStream_t)stream;
    if (!(b.order == 4 || b.order == 6) || !(b.rk_order == 4 || b.rk_order == 6) || b.nspecies < 1) return LK_ERR_ARG;
    if (b.ntiles != 1 || b.tile_lo[0] != 0 || b.tile_lo[1] != 0 || b.tile_n[0] != b.nglobal[0] || b.tile_n[1] != b.nglobal[1])
      return LK_ERR_UNSUPPORTED;  // the Maxwell path runs on one GPU
    ng = b.order == 4 ? 2 : 3;
    n1 = b.nglobal[0]; n2 = b.nglobal[1];
    if (n1 < b.order + 1 || n2 < b.order + 1) return LK_ERR_ARG;  // stencil width (KineticSpecies.C:495-504)
    n1d = n1 + 2 * ng; n2d = n2 + 2 * ng;
    pl = (i64)n1d * n2d;
    for (int k = 0; k < 2; ++k) dxg[k] = (b.xhi[k] - b.xlo[k]) / b.nglobal[k];
    for (int k = 0; k < 2; ++k) {
      LKH_CHECK(em[k].alloc(pl * 6));
      LKH_CUDA(cudaMemset(em[k].p, 0, sizeof(double) * pl * 6));
    }
    LKH_CHECK(em_rhs.alloc(pl * 6));
    LKH_CHECK(em_delta.alloc(pl * 6));
    if (b.rk_order == 6)
      for (int k = 0; k < 8; ++k) {
        LKH_CHECK(em_k[k].alloc(pl * 6));
        LKH_CUDA(cudaMemset(em_k[k].p, 0, sizeof(double) * pl * 6));
      }
    LKH_CUDA(cudaMemset(em_rhs.p, 0, sizeof(double) * pl * 6));
    LKH_CUDA(cudaMemset(em_delta.p, 0, sizeof(double) * pl * 6));
    LKH_CHECK(Jx.alloc(pl)); LKH_CHECK(Jy.alloc(pl)); LKH_CHECK(Jz.alloc(pl)); LKH_CHECK(sumf.alloc(pl));
    for (int s = 0; s < b.nspecies; ++s) {
      const lk_species_desc& sd = sdesc[s];
      if (sd.has_driver) return LK_ERR_UNSUPPORTED;
      KineticSpecies* ks = new KineticSpecies();
      species.push_back(ks);
      ks->g.n[0] = n1; ks->g.n[1] = n2; ks->g.n[2] = sd.nv[0]; ks->g.n[3] = sd.nv[1];
      ks->g.ng = ng; ks->g.order = b.order;
      ks->g.dx[0] = dxg[0]; ks->g.dx[1] = dxg[1];
      ks->g.dx[2] = (sd.vhi[0] - sd.vlo[0]) / sd.nv[0];
      ks->g.dx[3] = (sd.vhi[1] - sd.vlo[1]) / sd.nv[1];
      ks->mass = sd.mass; ks->charge = sd.charge; ks->bz_const = sd.bz_const;
      ks->vlo[0] = sd.vlo[0]; ks->vlo[1] = sd.vlo[1]; ks->vhi[0] = sd.vhi[0]; ks->vhi[1] = sd.vhi[1];
      ks->n1d = n1d; ks->n2d = n2d; ks->n3d = sd.nv[0] + 2 * ng; ks->n4d = sd.nv[1] + 2 * ng;
      ks->vol = (i64)n1d * n2d * ks->n3d * ks->n4d;
      for (int k = 0; k < 3; ++k) LKH_CHECK(ks->farr[k].alloc(ks->vol));
      LKH_CHECK(ks->delta.alloc(ks->vol));
      LKH_CUDA(cudaMemset(ks->farr[0].p, 0, sizeof(double) * ks->vol));
      LKH_CHECK(ks->buildVelocityArrays());
      LKH_CHECK(ks->Jx_s.alloc(pl)); LKH_CHECK(ks->Jy_s.alloc(pl)); LKH_CHECK(ks->Jz_s.alloc(pl));
      LKH_CHECK(ks->lam.alloc(2));
      LKH_CUDA(cudaMemset(ks->lam.p, 0, sizeof(double) * 2));
      memset(&ks->inflow, 0, sizeof(ks->inflow));
      {
        const int parts = lk_stage_moment_parts(&ks->g);
        if (parts < 1) return LK_ERR_ARG;
        const size_t cap = (size_t)3 * parts * n1 * n2;
        LKH_CHECK(ks->mom_part.alloc(cap));
        ks->mom.nmom = 3;
        ks->mom.partial = ks->mom_part.p;
        ks->mom.capacity = (int64_t)cap;
      }
      ks->f_eval = ks->state();
      for (int k = 0; k < 2; ++k) {
        vz[k].push_back(new DevBuf<double>());
        LKH_CHECK(vz[k].back()->alloc(pl));
        LKH_CUDA(cudaMemset(vz[k].back()->p, 0, sizeof(double) * pl));
      }
      if (b.rk_order == 6) {
        for (int k = 0; k < 8; ++k) {
          LKH_CHECK(ks->karr[k].alloc(ks->vol));
          vz_k[k].push_back(new DevBuf<double>());
          LKH_CHECK(vz_k[k].back()->alloc(pl));
          LKH_CUDA(cudaMemset(vz_k[k].back()->p, 0, sizeof(double) * pl));   // ghosts of a rhs stay zero
        }
      }
      vz_rhs.push_back(new DevBuf<double>());
      vz_delta.push_back(new DevBuf<double>());
      LKH_CHECK(vz_rhs.back()->alloc(pl));
      LKH_CHECK(vz_delta.back()->alloc(pl));
      LKH_CUDA(cudaMemset(vz_rhs.back()->p, 0, sizeof(double) * pl));
      LKH_CUDA(cudaMemset(vz_delta.back()->p, 0, sizeof(double) * pl));
    }
    return LK_OK;
  }

  lk_accel accelDesc(KineticSpecies* ks, const double* em_eval, const double* vz_eval) {
    lk_accel a;
    a.kind = 1;
    a.field = em_eval;        // the expansion of em_vars / vz to the species covers the whole box on one rank
    a.vz = vz_eval;
    a.vxface_velocities = ks->vxface.p;
    a.vyface_velocities = ks->vyface.p;
    a.normalization = ks->charge / ks->mass;
    a.bz_const = ks->bz_const;
    return a;
  }

  // (1) currentDensity of every species (KineticSpecies.C:853-895) + (2) Maxwell::fillGhostCells and the
  // net current sums (VMSystem.C:453-470)
  int currentsOf(const double* const* f_of, double* em_eval, double* const* vz_eval, bool allow_fused) {
    for (size_t s = 0; s < species.size(); ++s) {
      KineticSpecies* ks = species[s];
      const double dv = ks->g.dx[2] * ks->g.dx[3];
      if (allow_fused && ks->mom_valid && f_of[s] == ks->f_eval && !lk_get_strict()) {
        // production: sum f, sum vx f, sum vy f were left behind by the stage kernel that wrote f;
        // Jz = q dv vz(x,y) sum f (the reference sums f*vz(x,y): same value up to rounding)
        LKH_CHECK(lk_moments_finish(sumf.p, ks->Jx_s.p, ks->Jy_s.p, &ks->mom, &ks->g, dv, ks->charge, st));
        k_mul2d<<<nb(pl, 128), 128, 0, st>>>(ks->Jz_s.p, sumf.p, vz_eval[s], pl);
      } else {
        LKH_CHECK(lk_current_density(ks->Jx_s.p, ks->Jy_s.p, ks->Jz_s.p, f_of[s], &ks->g, ks->velocities.p, vz_eval[s], dv,
                                     ks->charge, st));
      }
    }
    LKH_CHECK(lk_periodic_fill_2d(em_eval, n1, n2, ng, 6, 1, 1, st));
    for (size_t s = 0; s < species.size(); ++s) LKH_CHECK(lk_periodic_fill_2d(vz_eval[s], n1, n2, ng, 1, 1, 1, st));
    for (size_t s = 0; s < species.size(); ++s) {
      KineticSpecies* ks = species[s];
      if (s == 0) {
        LKH_CUDA(cudaMemcpyAsync(Jx.p, ks->Jx_s.p, sizeof(double) * pl, cudaMemcpyDeviceToDevice, st));
        LKH_CUDA(cudaMemcpyAsync(Jy.p, ks->Jy_s.p, sizeof(double) * pl, cudaMemcpyDeviceToDevice, st));
        LKH_CUDA(cudaMemcpyAsync(Jz.p, ks->Jz_s.p, sizeof(double) * pl, cudaMemcpyDeviceToDevice, st));
      } else {
        k_add_inplace<<<nb(pl, 128), 128, 0, st>>>(Jx.p, ks->Jx_s.p, pl);
        k_add_inplace<<<nb(pl, 128), 128, 0, st>>>(Jy.p, ks->Jy_s.p, pl);
        k_add_inplace<<<nb(pl, 128), 128, 0, st>>>(Jz.p, ks->Jz_s.p, pl);
      }
    }
    return LK_OK;
  }
  // (6) Maxwell::evalRHS (Maxwell.C:562-623)
  int maxwellRHS(double* rhs_em, double* const* rhs_vz, const double* em_eval) {
    LKH_CHECK(lk_maxwell_rhs(rhs_em, em_eval, Jx.p, Jy.p, Jz.p, n1, n2, ng, desc.base.order, dxg, desc.light_speed,
                             desc.av_weak, desc.av_strong, st));
    for (size_t s = 0; s < species.size(); ++s)
      k_vz_rhs<<<nb((i64)n1 * n2, 128), 128, 0, st>>>(rhs_vz[s], em_eval, species[s]->charge / species[s]->mass, n1, n2, ng);
    return LK_OK;
  }

  // one RK4 stage (RK4Integrator.H:149-171) over the whole VMState
  int stage(int stg) {
    static const double THIRD = 1.0 / 3.0;
    const double dtOn2 = 0.5 * dt, dtOn3 = THIRD * dt, dtOn6 = 0.5 * dtOn3;
    const double w_eval[4] = {dtOn6, dtOn3, dtOn3, dtOn6};
    const double w_upd[4] = {dtOn2, dtOn2, dt, 1.0};
    double* em_eval = (stg == 0) ? emState() : emNew();
    std::vector<double*> vz_eval(species.size()), rhs_vz(species.size());
    std::vector<const double*> f_of(species.size());
    for (size_t s = 0; s < species.size(); ++s) {
      vz_eval[s] = (stg == 0) ? vzState((int)s) : vzNew((int)s);
      rhs_vz[s] = vz_rhs[s]->p;
      f_of[s] = species[s]->f_eval;
    }
    LKH_CHECK(currentsOf(f_of.data(), em_eval, vz_eval.data(), true));
    static const bool no_fuse = getenv("LK_NO_FUSED_MOMENTS") != nullptr;
    const bool fused_moments = !lk_get_strict() && !no_fuse;
    for (size_t s = 0; s < species.size(); ++s) {
      KineticSpecies* ks = species[s];
      LKH_CHECK(ks->periodicFill(ks->f_eval, 3, st));  // fillAdvectionGhostCells
      lk_accel a = accelDesc(ks, em_eval, vz_eval[s]);
      if (stg == 3) LKH_CHECK(lk_max_accel(&ks->g, &a, ks->lam.p, st));  // consumed by the next stableDt only
      const int at[4] = {1, 1, 1, 1};
      LKH_CHECK(lk_set_acceleration_bcs_4d(ks->f_eval, &ks->g, &a, &ks->inflow, at, st));
      lk_rk_update u;
      memset(&u, 0, sizeof(u));
      double* pred = (ks->f_eval == ks->farr[ks->i_a].p) ? ks->farr[ks->i_b].p : ks->farr[ks->i_a].p;
      u.f_old = ks->state();
      u.pred = pred;
      u.delta_in = (stg == 0) ? nullptr : ks->delta.p;
      u.delta_out = (stg == 3) ? nullptr : ks->delta.p;
      u.w_delta = w_eval[stg];
      u.c_pred = w_upd[stg];
      u.use_delta = (stg == 3);
      u.wrap = fused_moments ? ks->wrapFor(3) : 0;
      LKH_CHECK(lk_vlasov_stage(nullptr, ks->f_eval, &ks->g, ks->velocities.p, &a, &u, fused_moments ? &ks->mom : nullptr, st));
      ks->wrap_ptr = pred;
      ks->wrap_bits = u.wrap;
      ks->mom_valid = fused_moments;
      ks->f_eval = pred;
    }
    // Maxwell: rhs (ghosts of m_rhs stay zero), delta += w rhs, new = old (ghosts too) + c (rhs | delta)
    LKH_CHECK(maxwellRHS(em_rhs.p, rhs_vz.data(), em_eval));
    if (stg == 0) {
      LKH_CUDA(cudaMemsetAsync(em_delta.p, 0, sizeof(double) * pl * 6, st));
      for (size_t s = 0; s < species.size(); ++s) LKH_CUDA(cudaMemsetAsync(vz_delta[s]->p, 0, sizeof(double) * pl, st));
    }
    LKH_CHECK(lk_xpby2d(em_delta.p, em_rhs.p, w_eval[stg], n1, n2, ng, 6, st));
    LKH_CUDA(cudaMemcpyAsync(emNew(), emState(), sizeof(double) * pl * 6, cudaMemcpyDeviceToDevice, st));
    LKH_CHECK(lk_xpby2d(emNew(), (stg < 3) ? em_rhs.p : em_delta.p, w_upd[stg], n1, n2, ng, 6, st));
    for (size_t s = 0; s < species.size(); ++s) {
      LKH_CHECK(lk_xpby2d(vz_delta[s]->p, vz_rhs[s]->p, w_eval[stg], n1, n2, ng, 1, st));
      LKH_CUDA(cudaMemcpyAsync(vzNew((int)s), vzState((int)s), sizeof(double) * pl, cudaMemcpyDeviceToDevice, st));
      LKH_CHECK(lk_xpby2d(vzNew((int)s), (stg < 3) ? vz_rhs[s]->p : vz_delta[s]->p, w_upd[stg], n1, n2, ng, 1, st));
    }
    lambda_stale = true;
    return LK_OK;
  }

  // one RK6 stage (RK6Integrator.H:105-133) over the whole VMState: k[stg] = rhs(predictor), then the next predictor
  // (or, after the last stage, the new state) = old + dt * sum_j coef_j k[j], added in the order j = 0 .. stg
  int stage6(int stg) {
    const int last = 7;
    const double* coef = (stg == last) ? rk6tab::b : rk6tab::A[stg + 1];
    double* em_eval = (stg == 0) ? emState() : emNew();
    std::vector<double*> vz_eval(species.size()), rhs_vz(species.size());
    std::vector<const double*> f_of(species.size());
    for (size_t s = 0; s < species.size(); ++s) {
      vz_eval[s] = (stg == 0) ? vzState((int)s) : vzNew((int)s);
      rhs_vz[s] = vz_k[stg][s]->p;
      f_of[s] = species[s]->f_eval;
    }
    LKH_CHECK(currentsOf(f_of.data(), em_eval, vz_eval.data(), true));
    static const bool no_fuse = getenv("LK_NO_FUSED_MOMENTS") != nullptr;
    const bool fused_moments = !lk_get_strict() && !no_fuse;
    for (size_t s = 0; s < species.size(); ++s) {
      KineticSpecies* ks = species[s];
      LKH_CHECK(ks->periodicFill(ks->f_eval, 3, st));
      lk_accel a = accelDesc(ks, em_eval, vz_eval[s]);
      if (stg == last) LKH_CHECK(lk_max_accel(&ks->g, &a, ks->lam.p, st));
      const int at[4] = {1, 1, 1, 1};
      LKH_CHECK(lk_set_acceleration_bcs_4d(ks->f_eval, &ks->g, &a, &ks->inflow, at, st));
      lk_rk_update u;
      memset(&u, 0, sizeof(u));
      double* pred = (ks->f_eval == ks->farr[ks->i_a].p) ? ks->farr[ks->i_b].p : ks->farr[ks->i_a].p;
      u.f_old = ks->state();
      u.pred = pred;
      int np = 0;
      for (int j = 0; j < stg; ++j) {
        u.k_prev[np] = ks->karr[j].p;
        u.c_prev[np] = dt * coef[j];
        ++np;
      }
      u.n_prev = np;
      u.c_pred = dt * coef[stg];
      u.wrap = fused_moments ? ks->wrapFor(3) : 0;
      LKH_CHECK(lk_vlasov_stage(ks->karr[stg].p, ks->f_eval, &ks->g, ks->velocities.p, &a, &u, fused_moments ? &ks->mom : nullptr, st));
      ks->wrap_ptr = pred;
      ks->wrap_bits = u.wrap;
      ks->mom_valid = fused_moments;
      ks->f_eval = pred;
    }
    LKH_CHECK(maxwellRHS(em_k[stg].p, rhs_vz.data(), em_eval));
    LKH_CUDA(cudaMemcpyAsync(emNew(), emState(), sizeof(double) * pl * 6, cudaMemcpyDeviceToDevice, st));
    for (int j = 0; j <= stg; ++j) LKH_CHECK(lk_xpby2d(emNew(), em_k[j].p, dt * coef[j], n1, n2, ng, 6, st));
    for (size_t s = 0; s < species.size(); ++s) {
      LKH_CUDA(cudaMemcpyAsync(vzNew((int)s), vzState((int)s), sizeof(double) * pl, cudaMemcpyDeviceToDevice, st));
      for (int j = 0; j <= stg; ++j) LKH_CHECK(lk_xpby2d(vzNew((int)s), vz_k[j][s]->p, dt * coef[j], n1, n2, ng, 1, st));
    }
    lambda_stale = true;
    return LK_OK;
  }

  int advance(double a_dt) {
    dt = a_dt;
    for (auto* ks : species) ks->f_eval = ks->state();
    if (desc.base.rk_order == 6) {
      for (int stg = 0; stg < 8; ++stg) LKH_CHECK(stage6(stg));
    } else {
      for (int stg = 0; stg < 4; ++stg) LKH_CHECK(stage(stg));
    }
    for (auto* ks : species) {
      int i_new = (ks->f_eval == ks->farr[ks->i_a].p) ? ks->i_a : ks->i_b;
      int i_other = (i_new == ks->i_a) ? ks->i_b : ks->i_a;
      int i_old = ks->i_state;
      ks->i_state = i_new;
      ks->i_a = i_old;
      ks->i_b = i_other;
      ks->f_eval = ks->state();
    }
    i_state = 1 - i_state;
    time += dt;
    return LK_OK;
  }

  // VMSystem::evalRHS in the reference's unfused order (parity hook)
  int evalRHS(double** rhs_dev, double* rhs_em, double** rhs_vz, double t) {
    (void)t;
    std::vector<double*> vz_eval(species.size());
    std::vector<const double*> f_of(species.size());
    for (size_t s = 0; s < species.size(); ++s) {
      vz_eval[s] = vzState((int)s);
      f_of[s] = species[s]->state();
      species[s]->mom_valid = false;
    }
    LKH_CHECK(currentsOf(f_of.data(), emState(), vz_eval.data(), false));
    for (size_t s = 0; s < species.size(); ++s) {
      KineticSpecies* ks = species[s];
      double* f = ks->state();
      LKH_CHECK(lk_periodic_fill_4d(f, &ks->g, 1, 1, st));
      LKH_CHECK(lk_advection_derivatives_4d(rhs_dev[s], f, &ks->g, ks->velocities.p, st));
      lk_accel a = accelDesc(ks, emState(), vz_eval[s]);
      LKH_CHECK(lk_max_accel(&ks->g, &a, ks->lam.p, st));
      const int at[4] = {1, 1, 1, 1};
      LKH_CHECK(lk_set_acceleration_bcs_4d(f, &ks->g, &a, &ks->inflow, at, st));
      LKH_CHECK(lk_acceleration_derivatives_4d(rhs_dev[s], f, &ks->g, &a, st));
    }
    LKH_CHECK(maxwellRHS(rhs_em, rhs_vz, emState()));
    lambda_stale = true;
    return LK_OK;
  }

  int refreshLambda() {
    if (!lambda_stale) return LK_OK;
    LKH_CUDA(cudaStreamSynchronize(st));
    for (auto* ks : species) {
      double l[2];
      LKH_CUDA(cudaMemcpy(l, ks->lam.p, sizeof(l), cudaMemcpyDeviceToHost));
      ks->lambda_max[2] = l[0];
      ks->lambda_max[3] = l[1];
    }
    lambda_stale = false;
    return LK_OK;
  }
};

}  // namespace loki

using loki::VPSystem;
struct lk_vp_system {
  VPSystem sys;
};

extern "C" {

int lk_vp_create(lk_vp_system** out, const lk_vp_desc* desc, void* stream) {
  if (!out || !desc || !desc->species) return LK_ERR_ARG;
  lk_vp_system* h = new lk_vp_system();
  int s = h->sys.create(desc, stream);
  if (s != LK_OK) {
    delete h;
    return s;
  }
  *out = h;
  return LK_OK;
}
void lk_vp_destroy(lk_vp_system* h) { delete h; }
int lk_vp_species_geom(const lk_vp_system* h, int s, lk_geom* g) {
  if (!h || !g || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  *g = h->sys.species[s]->g;
  return LK_OK;
}
int lk_vp_set_state(lk_vp_system* h, int s, const double* f_host) {
  if (!h || !f_host || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  auto* ks = h->sys.species[s];
  if (cudaMemcpy(ks->state(), f_host, sizeof(double) * ks->vol, cudaMemcpyHostToDevice) != cudaSuccess) return LK_ERR_CUDA;
  ks->f_eval = ks->state();
  ks->mom_valid = false;
  ks->wrap_ptr = nullptr;
  ks->preset[ks->i_state] = false;
  return LK_OK;
}
int lk_vp_get_state(lk_vp_system* h, int s, double* f_host) {
  if (!h || !f_host || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  auto* ks = h->sys.species[s];
  if (cudaStreamSynchronize(h->sys.st) != cudaSuccess) return LK_ERR_CUDA;
  if (cudaMemcpy(f_host, ks->state(), sizeof(double) * ks->vol, cudaMemcpyDeviceToHost) != cudaSuccess) return LK_ERR_CUDA;
  return LK_OK;
}
/* ---- streaming the state through the host without stalling (double buffering over the rotating arrays) ---- */
int lk_vp_download_state(lk_vp_system* h, int s, double* f_host, void* stream) {
  if (!h || !f_host || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  auto& S = h->sys;
  auto* ks = S.species[s];
  cudaStream_t cs = (cudaStream_t)stream;
  if (S.faceStream() != LK_OK) return LK_ERR_CUDA;
  if (!ks->ev_down && cudaEventCreateWithFlags(&ks->ev_down, cudaEventDisableTiming) != cudaSuccess) return LK_ERR_CUDA;
  if (cudaEventRecord(S.ev_pre, S.st) != cudaSuccess || cudaStreamWaitEvent(cs, S.ev_pre, 0) != cudaSuccess) return LK_ERR_CUDA;
  if (cudaMemcpyAsync(f_host, ks->state(), sizeof(double) * ks->vol, cudaMemcpyDeviceToHost, cs) != cudaSuccess) return LK_ERR_CUDA;
  if (cudaEventRecord(ks->ev_down, cs) != cudaSuccess) return LK_ERR_CUDA;
  ks->down_pending = true;
  return LK_OK;
}
int lk_vp_upload_next(lk_vp_system* h, int s, const double* f_host, void* stream) {
  if (!h || !f_host || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  auto& S = h->sys;
  auto* ks = S.species[s];
  cudaStream_t cs = (cudaStream_t)stream;
  if (S.faceStream() != LK_OK) return LK_ERR_CUDA;
  if (!ks->ev_up && cudaEventCreateWithFlags(&ks->ev_up, cudaEventDisableTiming) != cudaSuccess) return LK_ERR_CUDA;
  // the previous state's array: free once the step the main stream holds is done; never the array a download reads
  ks->next_idx = ks->i_a;
  if (cudaEventRecord(S.ev_pre, S.st) != cudaSuccess || cudaStreamWaitEvent(cs, S.ev_pre, 0) != cudaSuccess) return LK_ERR_CUDA;
  if (cudaMemcpyAsync(ks->farr[ks->next_idx].p, f_host, sizeof(double) * ks->vol, cudaMemcpyHostToDevice, cs) != cudaSuccess) return LK_ERR_CUDA;
  if (cudaEventRecord(ks->ev_up, cs) != cudaSuccess) return LK_ERR_CUDA;
  return LK_OK;
}
int lk_vp_adopt_next(lk_vp_system* h) {
  if (!h) return LK_ERR_ARG;
  auto& S = h->sys;
  for (auto* ks : S.species) {
    if (ks->next_idx < 0) continue;
    if (cudaStreamWaitEvent(S.st, ks->ev_up, 0) != cudaSuccess) return LK_ERR_CUDA;
    // roles: uploaded array -> state; the old state's array (a download may still read it) -> written last
    const int x = ks->i_state, y = ks->next_idx, z = (ks->i_a == y) ? ks->i_b : ks->i_a;
    ks->i_state = y;
    ks->i_a = z;
    ks->i_b = x;
    ks->next_idx = -1;
    ks->f_eval = ks->state();
    ks->mom_valid = false;
    ks->wrap_ptr = nullptr;
    ks->preset[y] = false;
  }
  return LK_OK;
}
double* lk_vp_state_ptr(lk_vp_system* h, int s) { return (h && s >= 0 && s < (int)h->sys.species.size()) ? h->sys.species[s]->state() : nullptr; }
double* lk_vp_eval_ptr(lk_vp_system* h, int s) { return (h && s >= 0 && s < (int)h->sys.species.size()) ? h->sys.species[s]->f_eval : nullptr; }
int lk_vp_set_inflow(lk_vp_system* h, int s, const double* fx, const double* fv, double fnorm, double frac) {
  if (!h || !fx || !fv || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  auto* ks = h->sys.species[s];
  std::vector<double> a(fx, fx + (size_t)ks->n1d * ks->n2d), b(fv, fv + (size_t)ks->n3d * ks->n4d);
  int st = ks->ic_fx.upload(a);
  if (st == LK_OK) st = ks->ic_fv.upload(b);
  if (st != LK_OK) return st;
  ks->inflow.kind = 1;
  ks->inflow.fx = ks->ic_fx.p;
  ks->inflow.fv = ks->ic_fv.p;
  ks->inflow.fnorm = fnorm;
  ks->inflow.frac = frac;
  ks->forgetPresets();
  return LK_OK;
}
int lk_vp_set_inflow2(lk_vp_system* h, int s, int kind, const double* fx, const double* fv, const double* fx2,
                      const double* fv2) {
  if (!h || !fx || !fv || !fx2 || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  if (!(kind == 2 || kind == 4) || (kind == 2 && !fv2)) return LK_ERR_ARG;
  auto* ks = h->sys.species[s];
  const size_t nxy = (size_t)ks->n1d * ks->n2d, nv = (size_t)ks->n3d * ks->n4d;
  int st = ks->ic_fx.upload(std::vector<double>(fx, fx + nxy));
  if (st == LK_OK) st = ks->ic_fv.upload(std::vector<double>(fv, fv + nv));
  if (st == LK_OK) st = ks->ic_fx2.upload(std::vector<double>(fx2, fx2 + nxy));
  if (st == LK_OK && kind == 2) st = ks->ic_fv2.upload(std::vector<double>(fv2, fv2 + nv));
  if (st != LK_OK) return st;
  memset(&ks->inflow, 0, sizeof(ks->inflow));
  ks->inflow.kind = kind;
  ks->inflow.fx = ks->ic_fx.p;
  ks->inflow.fv = ks->ic_fv.p;
  ks->inflow.fx2 = ks->ic_fx2.p;
  ks->inflow.fv2 = (kind == 2) ? ks->ic_fv2.p : nullptr;
  ks->forgetPresets();
  return LK_OK;
}
int lk_vp_set_boundary_options(lk_vp_system* h, int nonperiodic_x, int nonperiodic_y, int use_new_bcs) {
  if (!h) return LK_ERR_ARG;
  auto& S = h->sys;
  if ((nonperiodic_x || nonperiodic_y) && S.desc.ntiles != 1) return LK_ERR_UNSUPPORTED;  // the exchanger wraps periodically
  for (auto* ks : S.species) {
    if ((nonperiodic_x || nonperiodic_y) && ks->inflow.kind == 3) return LK_ERR_UNSUPPORTED;  // ghost tables hold v ghosts only
    ks->nonperiodic = (nonperiodic_x ? 1 : 0) | (nonperiodic_y ? 2 : 0);
    ks->use_new_bcs = use_new_bcs != 0;
    for (int k = 0; k < 2; ++k) {
      ks->at_xy[2 * k] = (S.desc.tile_lo[k] == 0);
      ks->at_xy[2 * k + 1] = (S.desc.tile_lo[k] + S.desc.tile_n[k] == S.desc.nglobal[k]);
    }
    ks->forgetPresets();
    ks->wrap_ptr = nullptr;
  }
  return LK_OK;
}
int lk_vp_set_krook(lk_vp_system* h, int s, const double* nu_host) {
  if (!h || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  auto* ks = h->sys.species[s];
  ks->has_krook = false;
  if (!nu_host) return LK_OK;
  std::vector<double> a(nu_host, nu_host + (size_t)ks->n1d * ks->n2d);
  int st = ks->krook_nu.upload(a);
  if (st != LK_OK) return st;
  ks->has_krook = true;
  ks->mom_valid = false;
  return LK_OK;
}
int lk_vp_set_pitch_angle(lk_vp_system* h, int s, const lk_pitch_angle* p) {
  if (!h || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  auto* ks = h->sys.species[s];
  ks->has_coll = false;
  if (!p) return LK_OK;
  int st = lk_pitch_angle_check(&ks->g, ks->vlo, ks->vhi, p);
  if (st != LK_OK) return st;
  ks->coll = *p;
  ks->has_coll = true;
  ks->mom_valid = false;
  ks->forgetPresets();
  return LK_OK;
}
int lk_vp_set_time(lk_vp_system* h, double t) {
  if (!h) return LK_ERR_ARG;
  h->sys.time = t;
  return LK_OK;
}
double lk_vp_time(const lk_vp_system* h) { return h ? h->sys.time : 0.0; }
int lk_vp_stable_dt(lk_vp_system* h, double* dt) {
  if (!h || !dt) return LK_ERR_ARG;
  int s = h->sys.refreshLambda();
  if (s != LK_OK) return s;
  double v = std::numeric_limits<double>::max();
  for (auto* ks : h->sys.species) v = std::min(v, ks->computeDt(h->sys.desc.rk_order));
  *dt = v;
  return LK_OK;
}
int lk_vp_lambda_max(lk_vp_system* h, int s, double out[2]) {
  if (!h || !out || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  int st = h->sys.refreshLambda();
  if (st != LK_OK) return st;
  out[0] = h->sys.species[s]->lambda_max[2];
  out[1] = h->sys.species[s]->lambda_max[3];
  return LK_OK;
}
int lk_vp_set_lambda_max(lk_vp_system* h, int s, const double in[2]) {
  if (!h || !in || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  int st = h->sys.refreshLambda();  // drain the device values first so that they do not overwrite the global ones
  if (st != LK_OK) return st;
  h->sys.species[s]->lambda_max[2] = in[0];
  h->sys.species[s]->lambda_max[3] = in[1];
  return LK_OK;
}
int lk_vp_advance(lk_vp_system* h, double dt) { return h ? h->sys.advance(dt) : LK_ERR_ARG; }
int lk_vp_nstages(const lk_vp_system* h) { return h ? h->sys.nstages() : 0; }
int lk_vp_begin_step(lk_vp_system* h, double dt) { return h ? h->sys.beginStep(dt) : LK_ERR_ARG; }
int lk_vp_stage_moments(lk_vp_system* h, int stage) {
  (void)stage;
  return h ? h->sys.momentsOf(true) : LK_ERR_ARG;
}
double* lk_vp_rho_tile_ptr(lk_vp_system* h) { return h ? h->sys.rho_tile_p : nullptr; }
double* lk_vp_rho_gather_ptr(lk_vp_system* h) { return h ? h->sys.rho_gather_p : nullptr; }
int lk_vp_set_comm_buffers(lk_vp_system* h, double* tile, double* gather) {
  if (!h || !tile || !gather) return LK_ERR_ARG;
  h->sys.rho_tile_p = tile;
  h->sys.rho_gather_p = gather;
  return LK_OK;
}
int lk_vp_stage_field(lk_vp_system* h, int stage, const int* tiles) {
  (void)stage;
  if (!h) return LK_ERR_ARG;
  int s = h->sys.fieldSolve(tiles);
  if (s != LK_OK) return s;
  if (tiles == nullptr || h->sys.desc.ntiles == 1) return h->sys.fillAdvectionGhostCellsLocal();
  return LK_OK;
}
int lk_vp_local_fill_needed(lk_vp_system* h, int s, int dir) {
  if (!h || s < 0 || s >= (int)h->sys.species.size() || dir < 0 || dir > 1) return 0;
  auto* ks = h->sys.species[s];
  const int dirs = 1 << dir;
  return ((ks->f_eval == ks->wrap_ptr) ? (dirs & ~ks->wrap_bits) : dirs) ? 1 : 0;
}
int lk_vp_local_fill(lk_vp_system* h, int s, int dir) {
  if (!h || s < 0 || s >= (int)h->sys.species.size() || dir < 0 || dir > 1) return LK_ERR_ARG;
  auto* ks = h->sys.species[s];
  return ks->periodicFill(ks->f_eval, 1 << dir, h->sys.st);
}
int lk_vp_stage_finish(lk_vp_system* h, int stage) {
  if (!h || stage < 0 || stage >= h->sys.nstages()) return LK_ERR_ARG;
  return h->sys.stageFinish(stage);
}
int lk_vp_stage_finish_species(lk_vp_system* h, int stage, int s) {
  if (!h || stage < 0 || stage >= h->sys.nstages() || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  return h->sys.stageFinish(stage, h->sys.species[s]);
}
int lk_vp_stage_finish_species_part(lk_vp_system* h, int stage, int s, int part) {
  if (!h || stage < 0 || stage >= h->sys.nstages() || s < 0 || s >= (int)h->sys.species.size() || part < 1 || part > 2) return LK_ERR_ARG;
  return h->sys.stageFinish(stage, h->sys.species[s], part);
}
int lk_vp_wait_faces(lk_vp_system* h, int s, void* stream) {
  // make `stream` wait until everything a neighbour needs of species s's new predictor has been written: the face
  // tiles of a two-part stage (their own stream), else whatever the main stream holds now
  if (!h || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  auto& S = h->sys;
  auto* ks = S.species[s];
  if (ks->face_in_flight) return cudaStreamWaitEvent((cudaStream_t)stream, ks->ev_face, 0) == cudaSuccess ? LK_OK : LK_ERR_CUDA;
  if (S.faceStream() != LK_OK) return LK_ERR_CUDA;
  if (cudaEventRecord(S.ev_pre, S.st) != cudaSuccess) return LK_ERR_CUDA;
  return cudaStreamWaitEvent((cudaStream_t)stream, S.ev_pre, 0) == cudaSuccess ? LK_OK : LK_ERR_CUDA;
}
int lk_vp_end_step(lk_vp_system* h) { return h ? h->sys.endStep() : LK_ERR_ARG; }
int lk_vp_eval_rhs(lk_vp_system* h, double** rhs_dev, double time) { return (h && rhs_dev) ? h->sys.evalRHS(rhs_dev, time) : LK_ERR_ARG; }
const double* lk_vp_em_vars_ptr(const lk_vp_system* h) { return h ? h->sys.em_g.p : nullptr; }
const double* lk_vp_rho_ptr(const lk_vp_system* h) { return h ? h->sys.rho_g.p : nullptr; }
int lk_vp_time_history(lk_vp_system* h, double* out, int capacity, int* written) {
  // VPSystem::accumulateSequences (VPSystem.C:591-636) without probes / particles / flux histories:
  // Poisson's five field histories of the field of the last evalRHS, then per species computeke's five
  // and the integrated driver work
  if (!h || !out || !written) return LK_ERR_ARG;
  auto& S = h->sys;
  const int ns = (int)S.species.size(), count = 5 + 6 * ns;
  *written = 0;
  if (capacity < count) return LK_ERR_ARG;
  loki::DevBuf<double>& d = S.hist_scratch;   // kept with the system: no cudaMalloc / cudaFree (implicit sync) per record
  int st = (d.p && d.n >= (size_t)(5 + 5 * ns)) ? LK_OK : d.alloc(5 + 5 * ns);
  if (st != LK_OK) return st;
  st = lk_field_history(d.p, S.em_g.p, S.desc.nglobal[0], S.desc.nglobal[1], S.ng, 2, S.dxg, S.st);
  for (int s = 0; s < ns && st == LK_OK; ++s) {
    auto* ks = S.species[s];
    st = lk_compute_ke(d.p + 5 + 5 * s, ks->state(), &ks->g, ks->mass, ks->velocities.p, nullptr, S.st);
  }
  if (st != LK_OK) return st;
  if (cudaStreamSynchronize(S.st) != cudaSuccess) return LK_ERR_CUDA;
  std::vector<double> hbuf(5 + 5 * ns);
  if (cudaMemcpy(hbuf.data(), d.p, sizeof(double) * hbuf.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return LK_ERR_CUDA;
  for (int k = 0; k < 5; ++k) out[k] = hbuf[k];
  for (int s = 0; s < ns; ++s) {
    for (int k = 0; k < 5; ++k) out[5 + 6 * s + k] = hbuf[5 + 5 * s + k];
    double v = 0.0;
    if (S.species[s]->has_driver && cudaMemcpy(&v, S.species[s]->ke.p, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
      return LK_ERR_CUDA;
    out[5 + 6 * s + 5] = v;
  }
  *written = count;
  return LK_OK;
}
int lk_vp_probe_history(lk_vp_system* h, int nprobes, const double* frac_x, const double* frac_y, double* out) {
  // Poisson::accumulateSequences, the probe part (Poisson.C:852-860): E at the cell floor(frac * N) of each probe, from
  // the field of the last evalRHS; a rank reports the probes inside its own tile and 0 for the others (summed over
  // ranks by the caller, Poisson.C:878-887)
  if (!h || nprobes < 0 || (nprobes > 0 && (!frac_x || !frac_y || !out))) return LK_ERR_ARG;
  auto& S = h->sys;
  if (cudaStreamSynchronize(S.st) != cudaSuccess) return LK_ERR_CUDA;
  const size_t pl = (size_t)S.n1d_g * S.n2d_g;
  for (int k = 0; k < nprobes; ++k) {
    const int ip = (int)floor(frac_x[k] * S.desc.nglobal[0]), jp = (int)floor(frac_y[k] * S.desc.nglobal[1]);
    out[2 * k] = out[2 * k + 1] = 0.0;
    if (ip < S.desc.tile_lo[0] || ip >= S.desc.tile_lo[0] + S.desc.tile_n[0] || jp < S.desc.tile_lo[1] ||
        jp >= S.desc.tile_lo[1] + S.desc.tile_n[1])
      continue;
    const size_t o = (size_t)(ip + S.ng) + (size_t)S.n1d_g * (jp + S.ng);
    if (cudaMemcpy(&out[2 * k], S.em_g.p + o, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(&out[2 * k + 1], S.em_g.p + pl + o, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
      return LK_ERR_CUDA;
  }
  return LK_OK;
}
int lk_vp_flux_history(lk_vp_system* h, double* out, int capacity, int* written) {
  // KineticSpecies::accumulateSequencesCommon (KineticSpecies.C:2052-2097): the kinetic-energy flux of every species
  // through the eight phase-space boundaries, straight from the state (lk_ke_flux_boundaries: no face / flux arrays).
  // Ghosts as there: x / y refreshed in the directions this rank wraps itself (a cut direction needs the caller's halo
  // exchange of lk_vp_state_ptr first), then the velocity-boundary fill with the acceleration of the last evalRHS.
  if (!h || !out || !written) return LK_ERR_ARG;
  auto& S = h->sys;
  const int ns = (int)S.species.size();
  *written = 0;
  if (capacity < 8 * ns) return LK_ERR_ARG;
  loki::DevBuf<double>& d = S.flux_scratch;
  int st = (d.p && d.n >= (size_t)(8 * ns)) ? LK_OK : d.alloc(8 * ns);
  if (st != LK_OK) return st;
  for (int s = 0; s < ns; ++s) {
    auto* ks = S.species[s];
    double* f = ks->state();
    st = ks->fillAdvectionGhosts(f, S.uncutDirs(), S.st);
    if (st == LK_OK) st = ks->setAccelerationBCs(f, S.st);
    if (st != LK_OK) return st;
    lk_accel a = ks->accelDesc();
    int at[8];
    for (int k = 0; k < 2; ++k) {
      at[2 * k] = (S.desc.tile_lo[k] == 0);
      at[2 * k + 1] = (S.desc.tile_lo[k] + S.desc.tile_n[k] == S.desc.nglobal[k]);
    }
    at[4] = at[5] = at[6] = at[7] = 1;  // velocity space is whole on every rank
    st = lk_ke_flux_boundaries(d.p + 8 * s, f, &ks->g, ks->velocities.p, &a, ks->mass, at, S.st);
    if (st != LK_OK) return st;
  }
  if (cudaStreamSynchronize(S.st) != cudaSuccess) return LK_ERR_CUDA;
  if (cudaMemcpy(out, d.p, sizeof(double) * 8 * ns, cudaMemcpyDeviceToHost) != cudaSuccess) return LK_ERR_CUDA;
  *written = 8 * ns;
  return LK_OK;
}
int lk_vp_ke_e_dot(lk_vp_system* h, int s, double* value) {
  if (!h || !value || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  if (cudaStreamSynchronize(h->sys.st) != cudaSuccess) return LK_ERR_CUDA;
  if (cudaMemcpy(value, h->sys.species[s]->ke.p, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return LK_ERR_CUDA;
  return LK_OK;
}

int lk_vp_set_trig_tz(lk_vp_system* h, int s, int on, double amp, double electron_mass, double ion_mass) {
  // kinetic_species.N.tz.name / tz.amp / tz.electron_mass / tz.ion_mass (TZSourceFactory.C:22-56): the forcing of completeRHS
  if (!h || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  auto& S = h->sys;
  auto* ks = S.species[s];
  ks->has_tz = false;
  if (!on) return LK_OK;
  if (on < 1 || on > 4) return LK_ERR_ARG;
  const int kind = on - 1;
  const double params[3] = {amp, kind >= 2 ? electron_mass : 1.0, kind >= 2 ? ion_mass : 1.0};
  int64_t count = 0;
  int st = lk_trig_tz_table_count(&ks->g, &count);
  if (st != LK_OK) return st;
  st = ks->tz_tab.alloc((size_t)count);
  if (st != LK_OK) return st;
  const int lo[2] = {S.desc.tile_lo[0] - ks->g.ng, S.desc.tile_lo[1] - ks->g.ng};
  st = lk_trig_tz_tables(ks->tz_tab.p, &ks->g, lo, S.desc.xlo, ks->velocities.p, kind, params, S.st);
  if (st != LK_OK) return st;
  ks->has_tz = true;
  ks->tz_kind = kind;
  for (int k = 0; k < 3; ++k) ks->tz_params[k] = params[k];
  return LK_OK;
}
int lk_vp_trig_tz_error(lk_vp_system* h, int s, double time, double* error_host) {
  // TrigTZSource::computeError (TrigTZSource.C:63-82) of the state: what putToRestart dumps instead of the
  // distribution when a twilight zone is defined (KineticSpecies.C:987-1004)
  if (!h || !error_host || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  auto& S = h->sys;
  auto* ks = S.species[s];
  if (!ks->has_tz) return LK_ERR_ARG;
  if (!ks->rhs_tmp.p) LKH_CHECK(ks->rhs_tmp.alloc(ks->vol));
  LKH_CHECK(lk_compute_trig_tz_source_error(ks->rhs_tmp.p, ks->state(), &ks->g, ks->tz_tab.p, ks->velocities.p, time, ks->tz_kind, ks->tz_params, S.st));
  LKH_CUDA(cudaStreamSynchronize(S.st));
  LKH_CUDA(cudaMemcpy(error_host, ks->rhs_tmp.p, sizeof(double) * ks->vol, cudaMemcpyDeviceToHost));
  return LK_OK;
}
int lk_vp_update_ghosts(lk_vp_system* h) {
  // VPSystem::updateGhosts (VPSystem.C:779-797), the species part: fillAdvectionGhostCells of the state on this rank
  if (!h) return LK_ERR_ARG;
  return h->sys.fillAdvectionGhostCellsLocal();
}
int lk_vp_set_ke_e_dot(lk_vp_system* h, int s, double value) {
  // KineticSpecies::getFromRestart (KineticSpecies.C:925): m_integrated_ke_e_dot as a restart dump recorded it
  if (!h || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  if (!h->sys.species[s]->has_driver) return LK_OK;   // the reference only records it for driven species (KineticSpecies.C:1011-1015)
  if (cudaStreamSynchronize(h->sys.st) != cudaSuccess) return LK_ERR_CUDA;
  if (cudaMemcpy(h->sys.species[s]->ke.p, &value, sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) return LK_ERR_CUDA;
  return LK_OK;
}
int lk_vp_driver_history(lk_vp_system* h, int s, double time, double* ke_e_dot, double* envel) {
  // KineticSpecies::accumulateSequencesCommon, the driver part (KineticSpecies.C:2099-2150): the driver evaluated at
  // `time` into m_ext_efield, computekeedot of the state against it, and the driver's time envelope
  if (!h || !ke_e_dot || !envel || s < 0 || s >= (int)h->sys.species.size()) return LK_ERR_ARG;
  auto& S = h->sys;
  auto* ks = S.species[s];
  *ke_e_dot = 0.0;
  *envel = 0.0;
  if (!ks->has_driver) return LK_OK;
  loki::DevBuf<double>& d = S.hist_scratch;
  int st = (d.p && d.n >= 1) ? LK_OK : d.alloc(16);
  if (st != LK_OK) return st;
  st = ks->evaluateDriver(time, S.desc.xlo, S.desc.tile_lo, S.st);
  if (st != LK_OK) return st;
  st = lk_ke_e_dot(d.p, ks->state(), &ks->g, ks->charge, ks->velocities.p, ks->ext_efield.p, S.st);
  if (st != LK_OK) return st;
  if (cudaStreamSynchronize(S.st) != cudaSuccess) return LK_ERR_CUDA;
  if (cudaMemcpy(ke_e_dot, d.p, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return LK_ERR_CUDA;
  *envel = ks->driver.active(time) ? ks->driver.envelope(time) : 0.0;
  return LK_OK;
}

/* ---- Vlasov-Maxwell ---- */
struct lk_vm_system {
  loki::VMSystem sys;
};
#define VM_SP_OK(h, s) ((h) && (s) >= 0 && (s) < (int)(h)->sys.species.size())
int lk_vm_create(lk_vm_system** out, const lk_vm_desc* desc, void* stream) {
  if (!out || !desc || !desc->base.species) return LK_ERR_ARG;
  lk_vm_system* h = new lk_vm_system();
  int s = h->sys.create(desc, stream);
  if (s != LK_OK) {
    delete h;
    return s;
  }
  *out = h;
  return LK_OK;
}
void lk_vm_destroy(lk_vm_system* h) { delete h; }
int lk_vm_species_geom(const lk_vm_system* h, int s, lk_geom* g) {
  if (!VM_SP_OK(h, s) || !g) return LK_ERR_ARG;
  *g = h->sys.species[s]->g;
  return LK_OK;
}
int lk_vm_set_state(lk_vm_system* h, int s, const double* f_host) {
  if (!VM_SP_OK(h, s) || !f_host) return LK_ERR_ARG;
  auto* ks = h->sys.species[s];
  if (cudaMemcpy(ks->state(), f_host, sizeof(double) * ks->vol, cudaMemcpyHostToDevice) != cudaSuccess) return LK_ERR_CUDA;
  ks->f_eval = ks->state();
  ks->mom_valid = false;
  ks->wrap_ptr = nullptr;
  return LK_OK;
}
int lk_vm_get_state(lk_vm_system* h, int s, double* f_host) {
  if (!VM_SP_OK(h, s) || !f_host) return LK_ERR_ARG;
  auto* ks = h->sys.species[s];
  if (cudaStreamSynchronize(h->sys.st) != cudaSuccess) return LK_ERR_CUDA;
  if (cudaMemcpy(f_host, ks->state(), sizeof(double) * ks->vol, cudaMemcpyDeviceToHost) != cudaSuccess) return LK_ERR_CUDA;
  return LK_OK;
}
double* lk_vm_state_ptr(lk_vm_system* h, int s) { return VM_SP_OK(h, s) ? h->sys.species[s]->state() : nullptr; }
int lk_vm_set_fields(lk_vm_system* h, const double* em_host) {
  if (!h || !em_host) return LK_ERR_ARG;
  return cudaMemcpy(h->sys.emState(), em_host, sizeof(double) * h->sys.pl * 6, cudaMemcpyHostToDevice) == cudaSuccess ? LK_OK : LK_ERR_CUDA;
}
int lk_vm_get_fields(lk_vm_system* h, double* em_host) {
  if (!h || !em_host) return LK_ERR_ARG;
  if (cudaStreamSynchronize(h->sys.st) != cudaSuccess) return LK_ERR_CUDA;
  return cudaMemcpy(em_host, h->sys.emState(), sizeof(double) * h->sys.pl * 6, cudaMemcpyDeviceToHost) == cudaSuccess ? LK_OK : LK_ERR_CUDA;
}
int lk_vm_set_vz(lk_vm_system* h, int s, const double* vz_host) {
  if (!VM_SP_OK(h, s) || !vz_host) return LK_ERR_ARG;
  return cudaMemcpy(h->sys.vzState(s), vz_host, sizeof(double) * h->sys.pl, cudaMemcpyHostToDevice) == cudaSuccess ? LK_OK : LK_ERR_CUDA;
}
int lk_vm_get_vz(lk_vm_system* h, int s, double* vz_host) {
  if (!VM_SP_OK(h, s) || !vz_host) return LK_ERR_ARG;
  if (cudaStreamSynchronize(h->sys.st) != cudaSuccess) return LK_ERR_CUDA;
  return cudaMemcpy(vz_host, h->sys.vzState(s), sizeof(double) * h->sys.pl, cudaMemcpyDeviceToHost) == cudaSuccess ? LK_OK : LK_ERR_CUDA;
}
const double* lk_vm_fields_ptr(lk_vm_system* h) { return h ? h->sys.emState() : nullptr; }
const double* lk_vm_current_ptr(lk_vm_system* h, int comp) {
  if (!h) return nullptr;
  return comp == 0 ? h->sys.Jx.p : (comp == 1 ? h->sys.Jy.p : (comp == 2 ? h->sys.Jz.p : nullptr));
}
int lk_vm_set_inflow(lk_vm_system* h, int s, const double* fx, const double* fv, double fnorm, double frac) {
  if (!VM_SP_OK(h, s) || !fx || !fv) return LK_ERR_ARG;
  auto* ks = h->sys.species[s];
  std::vector<double> a(fx, fx + (size_t)ks->n1d * ks->n2d), b(fv, fv + (size_t)ks->n3d * ks->n4d);
  int st = ks->ic_fx.upload(a);
  if (st == LK_OK) st = ks->ic_fv.upload(b);
  if (st != LK_OK) return st;
  memset(&ks->inflow, 0, sizeof(ks->inflow));
  ks->inflow.kind = 1;
  ks->inflow.fx = ks->ic_fx.p;
  ks->inflow.fv = ks->ic_fv.p;
  ks->inflow.fnorm = fnorm;
  ks->inflow.frac = frac;
  return LK_OK;
}
int lk_vm_set_inflow_ghosts(lk_vm_system* h, int s, const double* ghost3, const double* ghost4) {
  if (!VM_SP_OK(h, s) || !ghost3 || !ghost4) return LK_ERR_ARG;
  auto* ks = h->sys.species[s];
  const size_t n3 = (size_t)ks->n1d * ks->n2d * 2 * ks->g.ng * ks->n4d, n4 = (size_t)ks->n1d * ks->n2d * ks->n3d * 2 * ks->g.ng;
  std::vector<double> a(ghost3, ghost3 + n3), b(ghost4, ghost4 + n4);
  int st = ks->ic_ghost3.upload(a);
  if (st == LK_OK) st = ks->ic_ghost4.upload(b);
  if (st != LK_OK) return st;
  memset(&ks->inflow, 0, sizeof(ks->inflow));
  ks->inflow.kind = 3;
  ks->inflow.ghost3 = ks->ic_ghost3.p;
  ks->inflow.ghost4 = ks->ic_ghost4.p;
  return LK_OK;
}
int lk_vm_set_time(lk_vm_system* h, double t) {
  if (!h) return LK_ERR_ARG;
  h->sys.time = t;
  return LK_OK;
}
double lk_vm_time(const lk_vm_system* h) { return h ? h->sys.time : 0.0; }
int lk_vm_advance(lk_vm_system* h, double dt) { return h ? h->sys.advance(dt) : LK_ERR_ARG; }
int lk_vm_stable_dt(lk_vm_system* h, double* dt) {
  if (!h || !dt) return LK_ERR_ARG;
  int s = h->sys.refreshLambda();
  if (s != LK_OK) return s;
  double v = std::numeric_limits<double>::max();
  for (auto* ks : h->sys.species) v = std::min(v, ks->computeDt(h->sys.desc.base.rk_order));
  const double dt_maxwell = 1.0 / (h->sys.desc.light_speed * (1.0 / h->sys.dxg[0] + 1.0 / h->sys.dxg[1]));  // Maxwell.H:199-204
  *dt = std::min(v, dt_maxwell);
  return LK_OK;
}
int lk_vm_lambda_max(lk_vm_system* h, int s, double out[2]) {
  if (!VM_SP_OK(h, s) || !out) return LK_ERR_ARG;
  int st = h->sys.refreshLambda();
  if (st != LK_OK) return st;
  out[0] = h->sys.species[s]->lambda_max[2];
  out[1] = h->sys.species[s]->lambda_max[3];
  return LK_OK;
}
int lk_vm_time_history(lk_vm_system* h, double* out, int capacity, int* written) {
  // VMSystem's histories without probes / particles: Maxwell's twelve field histories of the current
  // em_vars, then per species computekemaxwell's {ke, ke_x, ke_y, px = 0, py = 0}
  if (!h || !out || !written) return LK_ERR_ARG;
  auto& S = h->sys;
  const int ns = (int)S.species.size(), count = 12 + 5 * ns;
  *written = 0;
  if (capacity < count) return LK_ERR_ARG;
  loki::DevBuf<double>& d = S.hist_scratch;
  int st = (d.p && d.n >= (size_t)count) ? LK_OK : d.alloc(count);
  if (st != LK_OK) return st;
  st = lk_field_history(d.p, S.emState(), S.n1, S.n2, S.ng, 6, S.dxg, S.st);
  for (int s = 0; s < ns && st == LK_OK; ++s) {
    auto* ks = S.species[s];
    st = lk_compute_ke(d.p + 12 + 5 * s, ks->state(), &ks->g, ks->mass, ks->velocities.p, S.vzState(s), S.st);
  }
  if (st != LK_OK) return st;
  if (cudaStreamSynchronize(S.st) != cudaSuccess) return LK_ERR_CUDA;
  if (cudaMemcpy(out, d.p, sizeof(double) * count, cudaMemcpyDeviceToHost) != cudaSuccess) return LK_ERR_CUDA;
  *written = count;
  return LK_OK;
}
int lk_vm_eval_rhs(lk_vm_system* h, double** rhs_dev, double* rhs_em_dev, double** rhs_vz_dev, double time) {
  return (h && rhs_dev && rhs_em_dev && rhs_vz_dev) ? h->sys.evalRHS(rhs_dev, rhs_em_dev, rhs_vz_dev, time) : LK_ERR_ARG;
}

}  // extern "C"
