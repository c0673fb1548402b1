// lk_capi.cu -- the extern "C" boundary declared in include/loki_b200.h.  Validates arguments,
// picks the arithmetic build (production / strict) and owns the small device scratch buffers.
// No CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "lk_launch.h"

namespace {
thread_local std::string g_err;
int g_strict = 0;
int g_variant = 0;
std::mutex g_mu;

int fail(int code, const char* what) {
  g_err = what;
  return code;
}
int cuda_fail(cudaError_t e, const char* where) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s: %s", where, cudaGetErrorString(e));
  g_err = buf;
  return LK_ERR_CUDA;
}
bool geom_ok(const lk_geom* g) {
  if (!g) return false;
  if (!((g->order == 4 && g->ng == 2) || (g->order == 6 && g->ng == 3))) return false;
  for (int k = 0; k < 4; ++k)
    if (g->n[k] < 1 || !(g->dx[k] > 0.0)) return false;
  return true;
}

// scratch buffers keyed by (device, stream, slot), grown on demand and kept: two systems on different streams or
// devices of one process never share one.  A buffer that has to grow is replaced after its stream has drained.
struct ScratchKey {
  int device;
  void* stream;
  int slot;
  bool operator<(const ScratchKey& o) const {
    if (device != o.device) return device < o.device;
    if (stream != o.stream) return stream < o.stream;
    return slot < o.slot;
  }
};
struct Scratch {
  void* p = nullptr;
  size_t bytes = 0;
};
std::map<ScratchKey, Scratch> g_scratch;
double* scratch(int slot, size_t bytes, void* stream) {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  Scratch& s = g_scratch[ScratchKey{dev, stream, slot}];
  if (s.bytes < bytes) {
    if (s.p) {
      cudaStreamSynchronize((cudaStream_t)stream);  // earlier launches on this stream may still use the old buffer
      cudaFree(s.p);
    }
    s.p = nullptr;
    s.bytes = 0;
    if (cudaMalloc(&s.p, bytes) != cudaSuccess) return nullptr;
    s.bytes = bytes;
  }
  return (double*)s.p;
}
int moment_chunks(const lk_geom* g) {
  if (g_strict) return 1;  // the reference's sequential sum order
  // enough CTAs to fill 148 SMs a few times over: grid = ceil(Nx/64) * Ny * chunks
  long long base = (long long)((g->n[0] + 63) / 64) * g->n[1];
  int c = (int)((148LL * 16 + base - 1) / base);
  if (c < 1) c = 1;
  if (c > g->n[3]) c = g->n[3];
  return c;
}
}  // namespace

#define DISPATCH(call) (g_strict ? lkstrict::call : lkfast::call)
#define CHECK_LAUNCH(expr, name)                       \
  do {                                                 \
    cudaError_t e__ = (expr);                          \
    if (e__ != cudaSuccess) return cuda_fail(e__, name); \
    return LK_OK;                                      \
  } while (0)

// ---- optional event timing of the fused stencil launches (bench.py's roofline.achieved) ----
namespace {
bool g_prof = false;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
size_t g_prof_used = 0;
}  // namespace

struct lk_poisson_plan {
  int nx, ny, ng, order;
  double *sx, *sy, *cx, *cy, *T, *X;  // device
  double *F1, *F2, *npart;            // device: FFT work arrays (power-of-two grids, production arithmetic)
};

// lk_fft.cu: O(N log N) production solve for power-of-two grids
namespace lkfft {
bool supported(int nx, int ny);
int neutralize_scratch_doubles();
cudaError_t neutralize(double* rho, int n1, int n2, int ng, double* part, cudaStream_t st, int64_t* launches);
cudaError_t poisson_fft(double* phi, const double* rho, int nx, int ny, int ng, const double* sx, const double* sy,
                        const double* cx, const double* cy, double* F1, double* F2, cudaStream_t st, int64_t* launches);
}
static int64_t g_fft_launches = 0;
// lk_bcs.cu: non-periodic x / y physical boundaries
namespace lkbcs {
cudaError_t set_advection_bcs(double* f, const lk_geom* g, const double* velocities, const lk_inflow* ic, const int at[4],
                              int periodic_x, int periodic_y, cudaStream_t st, int64_t* launches);
}
namespace lkbcs {
cudaError_t append_krook(double* rhs, const double* u, const lk_geom* g, const double* nu, double dt, const lk_inflow* ic,
                         cudaStream_t st, int64_t* launches);
cudaError_t set_bcs_jb(double* f, const lk_geom* g, const lk_accel* a, const double* velocities, const lk_inflow* ic,
                       const int sides[8], cudaStream_t st, int64_t* launches);
}
namespace lkbcs {
void trig_tz_tables(double* tab, const lk_geom* g, const int lo[2], const double xlo[2], const double* vel_host, int kind,
                    const double* params);
size_t trig_tz_table_count(const lk_geom* g);
cudaError_t trig_tz(double* out, const double* soln, const lk_geom* g, const double* tab_dev, const double* velocities,
                    double time, int kind, const double* params, cudaStream_t st, int64_t* launches);
}
namespace lkbcs {
cudaError_t zero_ghost_2d(double* u, int n1, int n2, int ng, int dim, cudaStream_t st, int64_t* launches);
cudaError_t antenna_source(double* dem, const double* src, int n1, int n2, int ng, cudaStream_t st, int64_t* launches);
cudaError_t em_bcs(double* em, int n1, int n2, int ng, const int at[4], int x_periodic, int y_periodic, double c, cudaStream_t st,
                   int64_t* launches);
cudaError_t vz_bcs(double* vz, int n1, int n2, int ng, const int at[4], int x_periodic, int y_periodic, cudaStream_t st,
                   int64_t* launches);
}
// lk_coll.cu: pitch-angle collision operator
namespace lkcoll {
cudaError_t fields(double* ivx, double* ivy, double* vth, const double* u, const lk_geom* g, const double* velocities,
                   cudaStream_t st, int64_t* launches);
cudaError_t append(double* rhs, const double* f, const lk_geom* g, const double* velocities, const double* ivx,
                   const double* ivy, const double* vth, const double vlo[2], const double vhi[2],
                   const lk_pitch_angle* pa, cudaStream_t st, int64_t* launches);
}
// lk_diag.cu: time-history diagnostics
namespace lkdiag {
int ke_scratch_doubles();
cudaError_t compute_ke(double* out5, const double* f, const lk_geom* g, double mass, const double* velocities,
                       const double* vz, double* scratch, cudaStream_t st, int64_t* launches);
cudaError_t field_history(double* out, const double* em, int n1, int n2, int ng, int ncomp, double dx, double dy,
                          cudaStream_t st, int64_t* launches);
}

extern "C" {

int lk_version(void) { return 100; }
const char* lk_last_error(void) { return g_err.c_str(); }
int lk_set_strict(int strict) {
  int old = g_strict;
  g_strict = strict ? 1 : 0;
  return old;
}
int lk_get_strict(void) { return g_strict; }
int lk_set_rhs_variant(int variant) {
  int old = g_variant;
  g_variant = (variant == 1 || variant == 2) ? variant : 0;
  return old;
}
int lk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}
int64_t lk_pipe_launch_count(void) { return lkfast::pipe_launches(); }
int64_t lk_launch_count(void) { return lkfast::launches() + lkstrict::launches() + g_fft_launches; }

int lk_weno_fit(int order, const double* u, const double* vel, double* face, int64_t count, void* stream) {
  if ((order != 4 && order != 6) || count < 0 || (count > 0 && (!u || !vel || !face))) return fail(LK_ERR_ARG, "lk_weno_fit: bad argument");
  CHECK_LAUNCH(DISPATCH(weno_fit)(order, u, vel, face, count, (cudaStream_t)stream), "lk_weno_fit");
}

int lk_xpby4d(double* x, const double* y, double b, const lk_geom* g, void* stream) {
  if (!geom_ok(g) || !x || !y) return fail(LK_ERR_ARG, "lk_xpby4d: bad argument");
  CHECK_LAUNCH(DISPATCH(xpby4d)(x, y, b, g, (cudaStream_t)stream), "lk_xpby4d");
}

static bool accel_ok(const lk_accel* a) {
  if (!a || !a->field) return false;
  if (a->kind == 2) return a->vz != nullptr;  // materialised vel3 / vel4
  if (!a->vxface_velocities || !a->vyface_velocities) return false;
  if (a->kind != 0 && a->kind != 1) return false;
  if (a->kind == 1 && !a->vz) return false;
  return true;
}

int lk_max_accel(const lk_geom* g, const lk_accel* a, double* out, void* stream) {
  if (!geom_ok(g) || !accel_ok(a) || !out || a->kind == 2) return fail(LK_ERR_ARG, "lk_max_accel: bad argument");
  double* s4 = scratch(2, sizeof(double) * 4, stream);
  if (!s4) return cuda_fail(cudaGetLastError(), "lk_max_accel: scratch");
  CHECK_LAUNCH(DISPATCH(max_accel)(g, a, out, s4, (cudaStream_t)stream), "lk_max_accel");
}
int lk_set_phase_space_vel_4d(double* vel3, double* vel4, const lk_geom* g, const lk_accel* a, double* out, void* stream) {
  if (!geom_ok(g) || !accel_ok(a) || !out || a->kind == 2) return fail(LK_ERR_ARG, "lk_set_phase_space_vel_4d: bad argument");
  CHECK_LAUNCH(DISPATCH(set_phase_space_vel)(vel3, vel4, g, a, out, (cudaStream_t)stream), "lk_set_phase_space_vel_4d");
}
int lk_set_acceleration_bcs_4d(double* f, const lk_geom* g, const lk_accel* a, const lk_inflow* ic, const int at[4],
                               void* stream) {
  if (!geom_ok(g) || !accel_ok(a) || !f || !at) return fail(LK_ERR_ARG, "lk_set_acceleration_bcs_4d: bad argument");
  if (g->n[2] < 3 || g->n[3] < 3) return fail(LK_ERR_ARG, "lk_set_acceleration_bcs_4d: need >= 3 velocity cells");
  if (ic) {
    if (ic->kind < 0 || ic->kind > 4) return fail(LK_ERR_ARG, "lk_set_acceleration_bcs_4d: bad inflow kind");
    if (ic->kind == 4 && (!ic->fx || !ic->fv || !ic->fx2)) return fail(LK_ERR_ARG, "lk_set_acceleration_bcs_4d: missing inflow tables");
    if ((ic->kind == 1 || ic->kind == 2) && (!ic->fx || !ic->fv)) return fail(LK_ERR_ARG, "lk_set_acceleration_bcs_4d: missing inflow tables");
    if (ic->kind == 2 && (!ic->fx2 || !ic->fv2)) return fail(LK_ERR_ARG, "lk_set_acceleration_bcs_4d: missing second inflow term");
    if (ic->kind == 3 && (!ic->ghost3 || !ic->ghost4)) return fail(LK_ERR_ARG, "lk_set_acceleration_bcs_4d: missing ghost tables");
  }
  CHECK_LAUNCH(DISPATCH(set_accel_bcs)(f, g, a, ic, at, (cudaStream_t)stream), "lk_set_acceleration_bcs_4d");
}
int lk_set_advection_bcs_4d(double* f, const lk_geom* g, const double* velocities, const lk_inflow* ic, const int at[4],
                            int periodic_x, int periodic_y, void* stream) {
  if (!geom_ok(g) || !f || !velocities || !at) return fail(LK_ERR_ARG, "lk_set_advection_bcs_4d: bad argument");
  if ((!periodic_x && g->n[0] < 3) || (!periodic_y && g->n[1] < 3)) return fail(LK_ERR_ARG, "lk_set_advection_bcs_4d: need >= 3 cells");
  if (ic) {
    if (ic->kind == 3) return fail(LK_ERR_UNSUPPORTED, "lk_set_advection_bcs_4d: ghost-table inflow holds velocity ghosts only");
    if (ic->kind < 0 || ic->kind > 4) return fail(LK_ERR_ARG, "lk_set_advection_bcs_4d: bad inflow kind");
    if (ic->kind != 0 && (!ic->fx || !ic->fv)) return fail(LK_ERR_ARG, "lk_set_advection_bcs_4d: missing inflow tables");
    if ((ic->kind == 2 && (!ic->fx2 || !ic->fv2)) || (ic->kind == 4 && !ic->fx2)) return fail(LK_ERR_ARG, "lk_set_advection_bcs_4d: missing second inflow term");
  }
  CHECK_LAUNCH(lkbcs::set_advection_bcs(f, g, velocities, ic, at, periodic_x, periodic_y, (cudaStream_t)stream, &g_fft_launches),
               "lk_set_advection_bcs_4d");
}
static int inflow_tables_ok(const lk_inflow* ic) {
  if (!ic) return 1;
  if (ic->kind < 0 || ic->kind > 4) return 0;
  if ((ic->kind == 1 || ic->kind == 2 || ic->kind == 4) && (!ic->fx || !ic->fv)) return 0;
  if ((ic->kind == 2 && (!ic->fx2 || !ic->fv2)) || (ic->kind == 4 && !ic->fx2)) return 0;
  if (ic->kind == 3 && (!ic->ghost3 || !ic->ghost4)) return 0;
  return 1;
}
int lk_pitch_angle_check(const lk_geom* g, const double* vlo, const double* vhi, const lk_pitch_angle* p) {
  if (!geom_ok(g) || !vlo || !vhi || !p) return fail(LK_ERR_ARG, "lk_pitch_angle_check: bad argument");
  if (p->conservative != 1 && g->order == 6)
    return fail(LK_ERR_ARG, "Non-conservative operator in 6th order not supported.");  // PitchAngleCollisionOperator.C:216-218
  // PitchAngleCollisionOperator.C:253-269
  const double rolloff = g->order == 4 ? 3 : 4;
  for (int k = 0; k < 2; ++k) {
    const double vmin = vlo[k] + rolloff * g->dx[2 + k], vmax = vhi[k] - rolloff * g->dx[2 + k];
    if (p->range_lo[k] <= vmin || p->range_hi[k] >= vmax)
      return fail(LK_ERR_ARG, k == 0 ? "x collision_vel_range box too large" : "y collision_vel_range box too large");
    if (p->range_lo[k] >= p->range_hi[k])
      return fail(LK_ERR_ARG, k == 0 ? "x collision_vel_range_lo exceeds x collision_vel_range_hi"
                                     : "y collision_vel_range_lo exceeds y collision_vel_range_hi");
  }
  return LK_OK;
}
double lk_pitch_angle_real_lam(const lk_geom* g, const lk_pitch_angle* p) {
  if (!g || !p) return 0.0;
  const double dv = fmin(g->dx[2], g->dx[3]);
  const double pi = 4.0 * atan(1.0);
  return p->nu_coef * pow(p->vthermal_dt, 3.0) * pi * pi / (dv * dv * fmax(dv, p->vfloor));
}
int lk_pitch_angle_fields(double* IVx, double* IVy, double* IVth, const double* u, const lk_geom* g, const double* velocities,
                          void* stream) {
  if (!geom_ok(g) || !IVx || !IVy || !IVth || !u || !velocities) return fail(LK_ERR_ARG, "lk_pitch_angle_fields: bad argument");
  CHECK_LAUNCH(lkcoll::fields(IVx, IVy, IVth, u, g, velocities, (cudaStream_t)stream, &g_fft_launches), "lk_pitch_angle_fields");
}
int lk_append_pitch_angle_collision(double* rhs, const double* f, const lk_geom* g, const double* velocities, const double* IVx,
                                    const double* IVy, const double* IVth, const double* vlo, const double* vhi,
                                    const lk_pitch_angle* p, void* stream) {
  if (!geom_ok(g) || !rhs || !f || !velocities || !IVx || !IVy || !IVth || !vlo || !vhi || !p || rhs == f)
    return fail(LK_ERR_ARG, "lk_append_pitch_angle_collision: bad argument");
  CHECK_LAUNCH(lkcoll::append(rhs, f, g, velocities, IVx, IVy, IVth, vlo, vhi, p, (cudaStream_t)stream, &g_fft_launches),
               "lk_append_pitch_angle_collision");
}
int lk_trig_tz_table_count(const lk_geom* g, int64_t* count) {
  if (!geom_ok(g) || !count) return fail(LK_ERR_ARG, "lk_trig_tz_table_count: bad argument");
  *count = (int64_t)lkbcs::trig_tz_table_count(g);
  return LK_OK;
}
static bool tz_args_ok(int kind, const double* params) {
  if (kind < 0 || kind > 3 || !params) return false;
  if (kind == 2 && !(params[1] > 0.0)) return false;
  if (kind == 3 && !(params[2] > 0.0)) return false;
  return true;
}
int lk_trig_tz_tables(double* tables, const lk_geom* g, const int lo[2], const double xlo[2], const double* velocities,
                      int kind, const double* params, void* stream) {
  // the time-independent factors of the source, built on the host with libm and left on the device in `tables`
  if (!geom_ok(g) || !tables || !lo || !xlo || !velocities || !tz_args_ok(kind, params))
    return fail(LK_ERR_ARG, "lk_trig_tz_tables: bad argument");
  const size_t nv = (size_t)(g->n[2] + 2 * g->ng) * (g->n[3] + 2 * g->ng) * 2;
  std::vector<double> vel(nv), tab(lkbcs::trig_tz_table_count(g));
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaMemcpy(vel.data(), velocities, sizeof(double) * nv, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return cuda_fail(e, "lk_trig_tz_tables");
  lkbcs::trig_tz_tables(tab.data(), g, lo, xlo, vel.data(), kind, params);
  e = cudaMemcpy(tables, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cuda_fail(e, "lk_trig_tz_tables");
  return LK_OK;
}
int lk_set_trig_tz_source(double* rhs, const lk_geom* g, const double* tables, const double* velocities, double time, int kind,
                          const double* params, void* stream) {
  if (!geom_ok(g) || !rhs || !tables || !velocities || !tz_args_ok(kind, params))
    return fail(LK_ERR_ARG, "lk_set_trig_tz_source: bad argument");
  CHECK_LAUNCH(lkbcs::trig_tz(rhs, nullptr, g, tables, velocities, time, kind, params, (cudaStream_t)stream, &g_fft_launches),
               "lk_set_trig_tz_source");
}
int lk_compute_trig_tz_source_error(double* error, const double* soln, const lk_geom* g, const double* tables,
                                    const double* velocities, double time, int kind, const double* params, void* stream) {
  if (!geom_ok(g) || !error || !soln || !tables || !velocities || !tz_args_ok(kind, params))
    return fail(LK_ERR_ARG, "lk_compute_trig_tz_source_error: bad argument");
  CHECK_LAUNCH(lkbcs::trig_tz(error, soln, g, tables, velocities, time, kind, params, (cudaStream_t)stream, &g_fft_launches),
               "lk_compute_trig_tz_source_error");
}
int lk_append_krook(double* rhs, const double* u, const lk_geom* g, const double* nu, double dt, const lk_inflow* ic, void* stream) {
  if (!geom_ok(g) || !rhs || !u || !nu || !(dt != 0.0) || !inflow_tables_ok(ic)) return fail(LK_ERR_ARG, "lk_append_krook: bad argument");
  if (ic && ic->kind == 3) return fail(LK_ERR_UNSUPPORTED, "lk_append_krook: ghost-table inflow holds velocity ghosts only");
  CHECK_LAUNCH(lkbcs::append_krook(rhs, u, g, nu, dt, ic, (cudaStream_t)stream, &g_fft_launches), "lk_append_krook");
}
int lk_set_acceleration_bcs_4d_jb(double* f, const lk_geom* g, const lk_accel* a, const lk_inflow* ic, const int at[4],
                                  void* stream) {
  if (!geom_ok(g) || !accel_ok(a) || !f || !at || !inflow_tables_ok(ic)) return fail(LK_ERR_ARG, "lk_set_acceleration_bcs_4d_jb: bad argument");
  const int sides[8] = {0, 0, 0, 0, at[0], at[1], at[2], at[3]};
  CHECK_LAUNCH(lkbcs::set_bcs_jb(f, g, a, nullptr, ic, sides, (cudaStream_t)stream, &g_fft_launches), "lk_set_acceleration_bcs_4d_jb");
}
int lk_set_advection_bcs_4d_jb(double* f, const lk_geom* g, const double* velocities, const lk_inflow* ic, const int at[4],
                               int periodic_x, int periodic_y, void* stream) {
  if (!geom_ok(g) || !f || !velocities || !at || !inflow_tables_ok(ic)) return fail(LK_ERR_ARG, "lk_set_advection_bcs_4d_jb: bad argument");
  if (ic && ic->kind == 3) return fail(LK_ERR_UNSUPPORTED, "lk_set_advection_bcs_4d_jb: ghost-table inflow holds velocity ghosts only");
  const int sides[8] = {!periodic_x && at[0], !periodic_x && at[1], !periodic_y && at[2], !periodic_y && at[3], 0, 0, 0, 0};
  CHECK_LAUNCH(lkbcs::set_bcs_jb(f, g, nullptr, velocities, ic, sides, (cudaStream_t)stream, &g_fft_launches), "lk_set_advection_bcs_4d_jb");
}
int lk_periodic_fill_4d(double* f, const lk_geom* g, int px, int py, void* stream) {
  if (!geom_ok(g) || !f) return fail(LK_ERR_ARG, "lk_periodic_fill_4d: bad argument");
  if ((px && g->n[0] < g->ng) || (py && g->n[1] < g->ng)) return fail(LK_ERR_ARG, "lk_periodic_fill_4d: box thinner than the ghost width");
  CHECK_LAUNCH(DISPATCH(periodic_fill_4d)(f, g, px, py, (cudaStream_t)stream), "lk_periodic_fill_4d");
}
int64_t lk_halo_count(const lk_geom* g, int dir) {
  if (!geom_ok(g) || dir < 0 || dir > 1) return -1;
  const int64_t n3d = g->n[2] + 2 * g->ng, n4d = g->n[3] + 2 * g->ng;
  return dir == 0 ? (int64_t)g->ng * g->n[1] * n3d * n4d : (int64_t)(g->n[0] + 2 * g->ng) * g->ng * n3d * n4d;
}
int lk_halo_pack(double* buf, const double* f, const lk_geom* g, int dir, int side, void* stream) {
  if (!geom_ok(g) || !buf || !f || dir < 0 || dir > 1 || side < 0 || side > 1) return fail(LK_ERR_ARG, "lk_halo_pack: bad argument");
  CHECK_LAUNCH(DISPATCH(halo_pack)(buf, f, g, dir, side, (cudaStream_t)stream), "lk_halo_pack");
}
int lk_halo_unpack(double* f, const double* buf, const lk_geom* g, int dir, int side, void* stream) {
  if (!geom_ok(g) || !buf || !f || dir < 0 || dir > 1 || side < 0 || side > 1) return fail(LK_ERR_ARG, "lk_halo_unpack: bad argument");
  CHECK_LAUNCH(DISPATCH(halo_unpack)(f, buf, g, dir, side, (cudaStream_t)stream), "lk_halo_unpack");
}

int lk_advection_derivatives_4d(double* rhs, const double* f, const lk_geom* g, const double* velocities, void* stream) {
  if (!geom_ok(g) || !rhs || !f || !velocities) return fail(LK_ERR_ARG, "lk_advection_derivatives_4d: bad argument");
  CHECK_LAUNCH(DISPATCH(vlasov_rhs)(rhs, f, g, velocities, nullptr, nullptr, 1, g_variant, nullptr, 0, (cudaStream_t)stream),
               "lk_advection_derivatives_4d");
}
int lk_acceleration_derivatives_4d(double* rhs, const double* f, const lk_geom* g, const lk_accel* a, void* stream) {
  if (!geom_ok(g) || !rhs || !f || !accel_ok(a)) return fail(LK_ERR_ARG, "lk_acceleration_derivatives_4d: bad argument");
  CHECK_LAUNCH(DISPATCH(vlasov_rhs)(rhs, f, g, nullptr, a, nullptr, 2 | 4, g_variant, nullptr, 0, (cudaStream_t)stream),
               "lk_acceleration_derivatives_4d");
}
static int stage_impl(double* rhs_out, const double* f, const lk_geom* g, const double* velocities, const lk_accel* a,
                      const lk_rk_update* upd, const lk_stage_moments* mom, void* stream, const char* who) {
  if (!geom_ok(g) || !f || !velocities || !accel_ok(a)) return fail(LK_ERR_ARG, "lk_vlasov_rhs: bad argument");
  if (!rhs_out && !upd) return fail(LK_ERR_ARG, "lk_vlasov_rhs: nothing to write");
  if (upd) {
    if (!upd->f_old || !upd->pred) return fail(LK_ERR_ARG, "lk_vlasov_rhs: rk update needs f_old and pred");
    if (upd->pred == f) return fail(LK_ERR_ARG, "lk_vlasov_rhs: pred must not alias the evaluated state");
    if (upd->n_prev < 0 || upd->n_prev > 7) return fail(LK_ERR_ARG, "lk_vlasov_rhs: n_prev out of range");
    if (upd->wrap && g_variant == 1) return fail(LK_ERR_UNSUPPORTED, "lk_vlasov_rhs: wrap needs the marching kernel (variant 0)");
    if (upd->wrap & ~3) return fail(LK_ERR_ARG, "lk_vlasov_rhs: bad wrap bits");
    for (int j = 0; j < upd->n_prev; ++j)
      if (!upd->k_prev[j]) return fail(LK_ERR_ARG, "lk_vlasov_rhs: missing k_prev");
    if (upd->delta_out && upd->delta_out == f) return fail(LK_ERR_ARG, "lk_vlasov_rhs: delta must not alias the evaluated state");
  }
  if (rhs_out == f) return fail(LK_ERR_ARG, "lk_vlasov_rhs: rhs must not alias the evaluated state");
  double* mpart = nullptr;
  int nmom = 0;
  if (mom) {
    if (!upd || !mom->partial || (mom->nmom != 1 && mom->nmom != 3)) return fail(LK_ERR_ARG, "lk_vlasov_stage: bad moments request");
    if (g_variant == 1) return fail(LK_ERR_UNSUPPORTED, "lk_vlasov_stage: moments need the marching kernel (variant 0)");
    const int64_t need = (int64_t)mom->nmom * DISPATCH(stage_moment_parts)(g) * g->n[0] * g->n[1];
    if (mom->capacity < need) return fail(LK_ERR_ARG, "lk_vlasov_stage: moment partial buffer too small");
    mpart = mom->partial;
    nmom = mom->nmom;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // setaccelerationbcs4d_ on behalf of the caller: folded into the pipelined kernel's boundary tiles, or run here
  lk_rk_update upd_local;
  if (upd && upd->krook_nu) {
    if (!upd->krook_ic || !(upd->krook_dt != 0.0) || !inflow_tables_ok(upd->krook_ic) || upd->krook_ic->kind == 3 || upd->krook_ic->kind == 0)
      return fail(LK_ERR_ARG, "lk_vlasov_stage: the Krook term needs krook_dt != 0 and initial-condition tables of kind 1, 2 or 4");
  }
  if (upd && upd->tile_set) {
    if (upd->tile_set < 0 || upd->tile_set > 2 || !(upd->cut_dirs & 3)) return fail(LK_ERR_ARG, "lk_vlasov_stage: bad tile_set / cut_dirs");
    if (!lk_vlasov_stage_can_split(rhs_out, g, a, upd)) return fail(LK_ERR_UNSUPPORTED, "lk_vlasov_stage: tile_set needs the pipelined kernel");
  }
  if (upd && upd->accel_bcs) {
    if (!lk_vlasov_stage_folds_bcs(rhs_out, g, a, upd)) {
      const int at[4] = {1, 1, 1, 1};
      int s = lk_set_acceleration_bcs_4d(const_cast<double*>(f), g, a, upd->accel_bcs, at, stream);
      if (s != LK_OK) return s;
      upd_local = *upd;
      upd_local.accel_bcs = nullptr;
      upd = &upd_local;
    }
  }
  if (!g_prof) CHECK_LAUNCH(DISPATCH(vlasov_rhs)(rhs_out, f, g, velocities, a, upd, 3, g_variant, mpart, nmom, st), who);
  if (g_prof_used == g_prof_events.size()) {
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return cuda_fail(cudaGetLastError(), "lk_vlasov_rhs: event");
    g_prof_events.push_back({e0, e1});
  }
  auto& ev = g_prof_events[g_prof_used++];
  cudaEventRecord(ev.first, st);
  cudaError_t e = DISPATCH(vlasov_rhs)(rhs_out, f, g, velocities, a, upd, 3, g_variant, mpart, nmom, st);
  cudaEventRecord(ev.second, st);
  if (e != cudaSuccess) return cuda_fail(e, who);
  return LK_OK;
}
int lk_vlasov_rhs(double* rhs_out, const double* f, const lk_geom* g, const double* velocities, const lk_accel* a,
                  const lk_rk_update* upd, void* stream) {
  return stage_impl(rhs_out, f, g, velocities, a, upd, nullptr, stream, "lk_vlasov_rhs");
}
int lk_rk_stage_update(const double* rhs, const lk_geom* g, const lk_rk_update* upd, void* stream) {
  if (!geom_ok(g) || !rhs || !upd || !upd->f_old || !upd->pred) return fail(LK_ERR_ARG, "lk_rk_stage_update: bad argument");
  if (upd->n_prev < 0 || upd->n_prev > 7) return fail(LK_ERR_ARG, "lk_rk_stage_update: n_prev out of range");
  if (upd->use_delta && !upd->delta_in) return fail(LK_ERR_ARG, "lk_rk_stage_update: use_delta needs delta_in");
  CHECK_LAUNCH(DISPATCH(rk_stage_update)(rhs, g, upd, (cudaStream_t)stream), "lk_rk_stage_update");
}
int lk_vlasov_stage_can_split(const double* rhs_out, const lk_geom* g, const lk_accel* a, const lk_rk_update* upd) {
  if (!g || !a || !upd || g_strict) return 0;
  return lkfast::stage_uses_pipe(g, a, upd, const_cast<double*>(rhs_out), 3, g_variant) ? 1 : 0;
}
int lk_vlasov_stage_folds_bcs(const double* rhs_out, const lk_geom* g, const lk_accel* a, const lk_rk_update* upd) {
  if (!g || !a || !upd || !upd->accel_bcs || g_strict) return 0;
  return lkfast::stage_folds_bcs(g, a, upd, const_cast<double*>(rhs_out), 3, g_variant) ? 1 : 0;
}
int lk_preset_inflow_ghosts_4d(double* f, const lk_geom* g, const lk_inflow* ic, void* stream) {
  if (!geom_ok(g) || !f || !ic) return fail(LK_ERR_ARG, "lk_preset_inflow_ghosts_4d: bad argument");
  if (ic->kind < 0 || ic->kind > 4) return fail(LK_ERR_ARG, "lk_preset_inflow_ghosts_4d: bad inflow kind");
  if ((ic->kind == 1 || ic->kind == 2 || ic->kind == 4) && (!ic->fx || !ic->fv)) return fail(LK_ERR_ARG, "lk_preset_inflow_ghosts_4d: missing inflow tables");
  if ((ic->kind == 2 || ic->kind == 4) && !ic->fx2) return fail(LK_ERR_ARG, "lk_preset_inflow_ghosts_4d: missing second inflow term");
  if (ic->kind == 2 && !ic->fv2) return fail(LK_ERR_ARG, "lk_preset_inflow_ghosts_4d: missing second inflow term");
  if (ic->kind == 3 && (!ic->ghost3 || !ic->ghost4)) return fail(LK_ERR_ARG, "lk_preset_inflow_ghosts_4d: missing ghost tables");
  CHECK_LAUNCH(DISPATCH(preset_inflow)(f, g, ic, (cudaStream_t)stream), "lk_preset_inflow_ghosts_4d");
}
int lk_vlasov_stage(double* rhs_out, const double* f, const lk_geom* g, const double* velocities, const lk_accel* a,
                    const lk_rk_update* upd, const lk_stage_moments* mom, void* stream) {
  return stage_impl(rhs_out, f, g, velocities, a, upd, mom, stream, "lk_vlasov_stage");
}
int lk_stage_moment_parts(const lk_geom* g) {
  if (!geom_ok(g)) return -1;
  return DISPATCH(stage_moment_parts)(g);
}
int lk_moments_finish(double* d0, double* d1, double* d2, const lk_stage_moments* mom, const lk_geom* g, double dv,
                      double weight, void* stream) {
  if (!geom_ok(g) || !mom || !mom->partial || !d0 || (mom->nmom == 3 && (!d1 || !d2)) || (mom->nmom != 1 && mom->nmom != 3))
    return fail(LK_ERR_ARG, "lk_moments_finish: bad argument");
  CHECK_LAUNCH(DISPATCH(moments_finish)(d0, d1, d2, mom->partial, DISPATCH(stage_moment_parts)(g), mom->nmom, g, dv, weight,
                                        (cudaStream_t)stream), "lk_moments_finish");
}
int lk_ke_e_dot_from_moments(double* out, const lk_stage_moments* mom, const lk_geom* g, double charge, const double* ext,
                             void* stream) {
  if (!geom_ok(g) || !mom || !mom->partial || mom->nmom != 3 || !out || !ext) return fail(LK_ERR_ARG, "lk_ke_e_dot_from_moments: bad argument");
  const int nparts = DISPATCH(stage_moment_parts)(g);
  const double* part1 = mom->partial + (size_t)nparts * g->n[0] * g->n[1];
  CHECK_LAUNCH(DISPATCH(ke_from_moment)(out, part1, nparts, g, charge, ext, (cudaStream_t)stream), "lk_ke_e_dot_from_moments");
}

int lk_reduce_4d_to_2d(double* dst, const double* f, const lk_geom* g, double dv, double weight, void* stream) {
  if (!geom_ok(g) || !dst || !f) return fail(LK_ERR_ARG, "lk_reduce_4d_to_2d: bad argument");
  const int chunks = moment_chunks(g);
  double* s = scratch(0, sizeof(double) * (size_t)g->n[0] * g->n[1] * chunks, stream);
  if (!s) return cuda_fail(cudaGetLastError(), "lk_reduce_4d_to_2d: scratch");
  CHECK_LAUNCH(DISPATCH(reduce_4d_to_2d)(dst, f, g, dv, weight, s, chunks, (cudaStream_t)stream), "lk_reduce_4d_to_2d");
}
int lk_current_density(double* Jx, double* Jy, double* Jz, const double* f, const lk_geom* g, const double* velocities,
                       const double* vz, double dv, double weight, void* stream) {
  if (!geom_ok(g) || !Jx || !Jy || !Jz || !f || !velocities || !vz) return fail(LK_ERR_ARG, "lk_current_density: bad argument");
  const int chunks = moment_chunks(g);
  double* s = scratch(0, sizeof(double) * 3 * (size_t)g->n[0] * g->n[1] * chunks, stream);
  if (!s) return cuda_fail(cudaGetLastError(), "lk_current_density: scratch");
  CHECK_LAUNCH(DISPATCH(current_density)(Jx, Jy, Jz, f, g, velocities, vz, dv, weight, s, chunks, (cudaStream_t)stream),
               "lk_current_density");
}
int lk_ke_e_dot(double* out, const double* f, const lk_geom* g, double charge, const double* velocities,
                const double* ext, void* stream) {
  if (!geom_ok(g) || !out || !f || !velocities || !ext) return fail(LK_ERR_ARG, "lk_ke_e_dot: bad argument");
  const int nblocks = 148 * 4;
  double* s = scratch(1, sizeof(double) * nblocks, stream);
  if (!s) return cuda_fail(cudaGetLastError(), "lk_ke_e_dot: scratch");
  CHECK_LAUNCH(DISPATCH(ke_e_dot)(out, f, g, charge, velocities, ext, s, nblocks, (cudaStream_t)stream), "lk_ke_e_dot");
}

// ---- Poisson ----
int lk_poisson_plan_create(lk_poisson_plan** plan, int nx, int ny, int ng, int order, double Lx, double Ly) {
  if (!plan || nx < 1 || ny < 1 || !(Lx > 0) || !(Ly > 0) || !(order == 4 || order == 6 || order == -1) || ng < 1)
    return fail(LK_ERR_ARG, "lk_poisson_plan_create: bad argument");
  // symbols of the 4th/6th-order FD Laplacian, pre-multiplied by nx*ny (LokiPoissonSolveFFT.C:66-115)
  const double pi = 4.0 * atan(1.0);
  const int nyh = ny / 2 + 1;
  std::vector<double> sx(nx), sy(nyh), cx(2 * nx), cy(2 * ny);
  const double dx = Lx / nx, dy = Ly / ny;
  auto sym = [&](double h, double k) {
    double dpdm = (2.0 * cos(h * k) - 2.0) / pow(h, 2.0);
    if (order == 4) return dpdm - pow(h, 2.0) / 12.0 * pow(dpdm, 2.0);
    if (order == 6) return dpdm - pow(h, 2.0) / 12.0 * pow(dpdm, 2.0) + pow(h, 4.0) / 90.0 * pow(dpdm, 3.0);
    return -k * k;
  };
  for (int i = 0; i < nx; ++i) {
    double kx = 0.0;
    if (i >= 1) kx = (2 * i < nx) ? (2.0 * pi / Lx) * i : (2.0 * pi / Lx) * (nx - i);
    sx[i] = sym(dx, kx) * (nx * ny);
  }
  for (int i = 0; i < nyh; ++i) {
    double ky = (i >= 1) ? (2.0 * pi / Ly) * i : 0.0;
    sy[i] = sym(dy, ky) * (nx * ny);
  }
  for (int k = 0; k < nx; ++k) { cx[2 * k] = cos(2.0 * pi * k / nx); cx[2 * k + 1] = sin(2.0 * pi * k / nx); }
  for (int k = 0; k < ny; ++k) { cy[2 * k] = cos(2.0 * pi * k / ny); cy[2 * k + 1] = sin(2.0 * pi * k / ny); }
  lk_poisson_plan* p = new lk_poisson_plan();
  memset(p, 0, sizeof(*p));
  p->nx = nx; p->ny = ny; p->ng = ng; p->order = order;
  cudaError_t e = cudaSuccess;
  auto up = [&](double** d, const std::vector<double>& h) {
    if (e != cudaSuccess) return;
    e = cudaMalloc((void**)d, sizeof(double) * h.size());
    if (e == cudaSuccess) e = cudaMemcpy(*d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice);
  };
  up(&p->sx, sx); up(&p->sy, sy); up(&p->cx, cx); up(&p->cy, cy);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->T, sizeof(double) * 2 * nx * nyh);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->X, sizeof(double) * 2 * nx * nyh);
  if (lkfft::supported(nx, ny)) {
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->F1, sizeof(double) * 2 * (size_t)nx * ny);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->F2, sizeof(double) * 2 * (size_t)nx * ny);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->npart, sizeof(double) * lkfft::neutralize_scratch_doubles());
  }
  if (e != cudaSuccess) {
    lk_poisson_plan_destroy(p);
    return cuda_fail(e, "lk_poisson_plan_create");
  }
  *plan = p;
  return LK_OK;
}
void lk_poisson_plan_destroy(lk_poisson_plan* p) {
  if (!p) return;
  cudaFree(p->sx); cudaFree(p->sy); cudaFree(p->cx); cudaFree(p->cy); cudaFree(p->T); cudaFree(p->X);
  cudaFree(p->F1); cudaFree(p->F2); cudaFree(p->npart);
  delete p;
}
int lk_electric_field(lk_poisson_plan* p, double* rho, double* phi, double* em, const double* dx, void* stream) {
  if (!p || !rho || !phi || !em || !dx) return fail(LK_ERR_ARG, "lk_electric_field: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  // EMSolverBase::electricField (EMSolverBase.C:270-371), single EM "processor" branch
  if (!g_strict && p->F1) {
    // production arithmetic on a power-of-two grid: two-level neutralisation sum + FFT passes
    e = lkfft::neutralize(rho, p->nx, p->ny, p->ng, p->npart, st, &g_fft_launches);
    if (e == cudaSuccess) e = lkfft::poisson_fft(phi, rho, p->nx, p->ny, p->ng, p->sx, p->sy, p->cx, p->cy, p->F1, p->F2, st, &g_fft_launches);
  } else {
    e = DISPATCH(neutralize)(rho, p->nx, p->ny, p->ng, st);
    if (e == cudaSuccess) e = DISPATCH(poisson_dft)(phi, rho, p->nx, p->ny, p->ng, p->sx, p->sy, p->cx, p->cy, p->T, p->X, st);
  }
  if (e == cudaSuccess) e = DISPATCH(periodic_fill_2d)(phi, p->nx, p->ny, p->ng, 1, 1, 1, st);
  const size_t pl = (size_t)(p->nx + 2 * p->ng) * (p->ny + 2 * p->ng);
  if (e == cudaSuccess) e = cudaMemsetAsync(em, 0, sizeof(double) * 2 * pl, st);  // m_em_vars = 0.0
  if (e == cudaSuccess) e = DISPATCH(efield_from_phi)(em, phi, p->nx, p->ny, p->ng, p->order == -1 ? 4 : p->order, dx[0], dx[1], st);
  if (e == cudaSuccess) e = DISPATCH(periodic_fill_2d)(em, p->nx, p->ny, p->ng, 2, 1, 1, st);
  if (e != cudaSuccess) return cuda_fail(e, "lk_electric_field");
  return LK_OK;
}
int lk_neutralize_charge(double* rho, int n1, int n2, int ng, void* stream) {
  if (!rho || n1 < 1 || n2 < 1 || ng < 0) return fail(LK_ERR_ARG, "lk_neutralize_charge: bad argument");
  CHECK_LAUNCH(DISPATCH(neutralize)(rho, n1, n2, ng, (cudaStream_t)stream), "lk_neutralize_charge");
}
int lk_efield_from_potential(double* em, const double* phi, int n1, int n2, int ng, int order, double dx, double dy,
                             void* stream) {
  if (!em || !phi || n1 < 1 || n2 < 1 || !((order == 4 && ng >= 2) || (order == 6 && ng >= 3)) || !(dx > 0.0) || !(dy > 0.0))
    return fail(LK_ERR_ARG, "lk_efield_from_potential: bad argument");
  CHECK_LAUNCH(DISPATCH(efield_from_phi)(em, phi, n1, n2, ng, order, dx, dy, (cudaStream_t)stream), "lk_efield_from_potential");
}
int lk_compute_ke(double* out5, const double* f, const lk_geom* g, double mass, const double* velocities, const double* vz,
                  void* stream) {
  if (!geom_ok(g) || !out5 || !f || !velocities) return fail(LK_ERR_ARG, "lk_compute_ke: bad argument");
  double* s = scratch(3, sizeof(double) * lkdiag::ke_scratch_doubles(), stream);
  if (!s) return cuda_fail(cudaGetLastError(), "lk_compute_ke: scratch");
  CHECK_LAUNCH(lkdiag::compute_ke(out5, f, g, mass, velocities, vz, s, (cudaStream_t)stream, &g_fft_launches), "lk_compute_ke");
}
int lk_field_history(double* out, const double* em, int n1, int n2, int ng, int ncomp, const double* dx, void* stream) {
  if (!out || !em || !dx || n1 < 1 || n2 < 1 || ng < 0 || !(ncomp == 2 || ncomp == 6)) return fail(LK_ERR_ARG, "lk_field_history: bad argument");
  CHECK_LAUNCH(lkdiag::field_history(out, em, n1, n2, ng, ncomp, dx[0], dx[1], (cudaStream_t)stream, &g_fft_launches), "lk_field_history");
}
int lk_periodic_fill_2d(double* u, int n1, int n2, int ng, int ncomp, int px, int py, void* stream) {
  if (!u || n1 < 1 || n2 < 1 || ng < 1 || ncomp < 1) return fail(LK_ERR_ARG, "lk_periodic_fill_2d: bad argument");
  CHECK_LAUNCH(DISPATCH(periodic_fill_2d)(u, n1, n2, ng, ncomp, px, py, (cudaStream_t)stream), "lk_periodic_fill_2d");
}
int lk_xpby2d(double* x, const double* y, double b, int n1, int n2, int ng, int ncomp, void* stream) {
  if (!x || !y || n1 < 1 || n2 < 1 || ng < 0 || ncomp < 1) return fail(LK_ERR_ARG, "lk_xpby2d: bad argument");
  CHECK_LAUNCH(DISPATCH(xpby2d)(x, y, b, n1, n2, ng, ncomp, (cudaStream_t)stream), "lk_xpby2d");
}
int lk_zero_ghost_2d(double* u, int n1, int n2, int ng, int ncomp, void* stream) {
  if (!u || n1 < 1 || n2 < 1 || ng < 0 || ncomp < 1) return fail(LK_ERR_ARG, "lk_zero_ghost_2d: bad argument");
  CHECK_LAUNCH(lkbcs::zero_ghost_2d(u, n1, n2, ng, ncomp, (cudaStream_t)stream, &g_fft_launches), "lk_zero_ghost_2d");
}
int lk_maxwell_add_antenna_source(double* dem, const double* antenna_source, int n1, int n2, int ng, void* stream) {
  if (!dem || !antenna_source || n1 < 1 || n2 < 1 || ng < 0) return fail(LK_ERR_ARG, "lk_maxwell_add_antenna_source: bad argument");
  CHECK_LAUNCH(lkbcs::antenna_source(dem, antenna_source, n1, n2, ng, (cudaStream_t)stream, &g_fft_launches), "lk_maxwell_add_antenna_source");
}
int lk_maxwell_set_em_bcs(double* em, int n1, int n2, int order, const int at[4], int x_periodic, int y_periodic, double light_speed,
                          void* stream) {
  if (!em || !at || (order != 4 && order != 6) || n1 < 3 || n2 < 3 || !(light_speed > 0.0))
    return fail(LK_ERR_ARG, "lk_maxwell_set_em_bcs: bad argument");
  CHECK_LAUNCH(lkbcs::em_bcs(em, n1, n2, order == 4 ? 2 : 3, at, x_periodic, y_periodic, light_speed, (cudaStream_t)stream, &g_fft_launches),
               "lk_maxwell_set_em_bcs");
}
int lk_maxwell_set_vz_bcs(double* vz, int n1, int n2, int order, const int at[4], int x_periodic, int y_periodic, void* stream) {
  if (!vz || !at || (order != 4 && order != 6) || n1 < 4 || n2 < 4) return fail(LK_ERR_ARG, "lk_maxwell_set_vz_bcs: bad argument");
  CHECK_LAUNCH(lkbcs::vz_bcs(vz, n1, n2, order == 4 ? 2 : 3, at, x_periodic, y_periodic, (cudaStream_t)stream, &g_fft_launches),
               "lk_maxwell_set_vz_bcs");
}
int lk_form_accel(double* accel, const double* em, const double* ext, double normalization, int n1, int n2, int ng,
                  void* stream) {
  if (!accel || !em || n1 < 1 || n2 < 1 || ng < 0) return fail(LK_ERR_ARG, "lk_form_accel: bad argument");
  CHECK_LAUNCH(DISPATCH(form_accel)(accel, em, ext, normalization, n1, n2, ng, (cudaStream_t)stream), "lk_form_accel");
}
int lk_maxwell_rhs(double* rhs, const double* em, const double* Jx, const double* Jy, const double* Jz, int n1, int n2,
                   int ng, int order, const double* dx, double c, double av_weak, double av_strong, void* stream) {
  if (!rhs || !em || !Jx || !Jy || !Jz || !dx || n1 < 1 || n2 < 1 || !((order == 4 && ng == 2) || (order == 6 && ng == 3)))
    return fail(LK_ERR_ARG, "lk_maxwell_rhs: bad argument");
  CHECK_LAUNCH(DISPATCH(maxwell_rhs)(rhs, em, Jx, Jy, Jz, n1, n2, ng, order, dx[0], dx[1], c, av_weak, av_strong,
                                     (cudaStream_t)stream),
               "lk_maxwell_rhs");
}

// ---- event timing of the fused stencil kernel ----
int lk_profile_enable(int on) {
  g_prof = on != 0;
  g_prof_used = 0;
  return LK_OK;
}
int lk_profile_summary(int64_t* launches, double* total_ms) {
  if (!launches || !total_ms) return fail(LK_ERR_ARG, "lk_profile_summary: bad argument");
  double tot = 0.0;
  for (size_t k = 0; k < g_prof_used; ++k) {
    if (cudaEventSynchronize(g_prof_events[k].second) != cudaSuccess) return cuda_fail(cudaGetLastError(), "lk_profile_summary");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_prof_events[k].first, g_prof_events[k].second);
    tot += ms;
  }
  *launches = (int64_t)g_prof_used;
  *total_ms = tot;
  return LK_OK;
}

// ---- memory helpers ----
int lk_malloc(void** p, int64_t bytes) {
  if (!p || bytes < 0) return fail(LK_ERR_ARG, "lk_malloc: bad argument");
  cudaError_t e = cudaMalloc(p, (size_t)bytes);
  if (e != cudaSuccess) return cuda_fail(e, "lk_malloc");
  return LK_OK;
}
int lk_free(void* p) {
  cudaError_t e = cudaFree(p);
  if (e != cudaSuccess) return cuda_fail(e, "lk_free");
  return LK_OK;
}
int lk_memcpy_h2d(void* dst, const void* src, int64_t bytes) {
  cudaError_t e = cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cuda_fail(e, "lk_memcpy_h2d");
  return LK_OK;
}
int lk_memcpy_d2h(void* dst, const void* src, int64_t bytes) {
  cudaError_t e = cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return cuda_fail(e, "lk_memcpy_d2h");
  return LK_OK;
}
int lk_memset(void* p, int value, int64_t bytes) {
  cudaError_t e = cudaMemset(p, value, (size_t)bytes);
  if (e != cudaSuccess) return cuda_fail(e, "lk_memset");
  return LK_OK;
}
int lk_sync(void* stream) {
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "lk_sync");
  return LK_OK;
}

}  // extern "C"
