// lk_f77.cu -- Level 0 of the drop-in boundary (include/loki_b200_f77.h): the reference's own Fortran-77
// symbols (xpby4d_, computeadvectionderivatives4d_, ...; KineticSpeciesF.H, PoissonF.H, MaxwellF.H) with the
// reference's own by-reference argument lists, on device arrays.  Each entry point rebuilds the geometry from the
// box integers the C++ wrappers pass (BOX4D_TO_FORT, tbox/Box.H:893-897) and forwards to the kernels behind
// loki_b200.h; the arrays the Fortran ABI materialises (vel3 / vel4 / vel1 / vel2) are read as they are.
// Compiled with -fmad=false: the two small kernels defined here are part of the bit-identical (strict) path.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/loki_b200.h"
#include "../../include/loki_b200_f77.h"

// lk_coll.cu: the pieces of PitchAngleCollisionOperator::evaluate the Fortran ABI exposes one by one
namespace lkcoll {
cudaError_t moments(double* rn, double* rgx, double* rgy, const double* u, const lk_geom* g, const double* velocities,
                    cudaStream_t st, int64_t* launches);
cudaError_t kec(double* rk, const double* vx0, const double* vy0, const double* u, const lk_geom* g,
                const double* velocities, cudaStream_t st, int64_t* launches);
cudaError_t reduced(int mode, double* vx, double* vy, const double* n, const double* gx, const double* gy, int64_t pl,
                    cudaStream_t st, int64_t* launches);
}

namespace {

typedef long long i64;
int g_status = LK_OK;

void fail(const char* who, const char* what) {
  g_status = LK_ERR_ARG;
  fprintf(stderr, "loki_b200 %s: %s\n", who, what);
}
bool check(int st) {
  g_status = st;
  return st == LK_OK;
}

// dataBox / interiorBox pair -> lk_geom (dx filled by the caller when the routine has it)
bool geom_from(const char* who, const int* const nd[8], const int* const n[8], int order_hint, lk_geom* g) {
  int ng = -1;
  for (int k = 0; k < 4; ++k) {
    const int lo = *n[2 * k] - *nd[2 * k], hi = *nd[2 * k + 1] - *n[2 * k + 1];
    g->n[k] = *n[2 * k + 1] - *n[2 * k] + 1;
    if (lo != hi || lo < 0 || (ng >= 0 && lo != ng) || g->n[k] < 1) {
      fail(who, "boxes are not an interior box grown by the same ghost width in every direction");
      return false;
    }
    ng = lo;
    g->dx[k] = 1.0;
  }
  if (order_hint == 0) order_hint = (ng == 3) ? 6 : 4;
  if (!((order_hint == 4 && ng == 2) || (order_hint == 6 && ng == 3))) {
    fail(who, "ghost width does not match solution_order (2 for order 4, 3 for order 6; KineticSpecies.C:155-160)");
    return false;
  }
  g->ng = ng;
  g->order = order_hint;
  return true;
}
bool geom2_from(const char* who, const int* const nd[4], const int* const n[4], int* n1, int* n2, int* ng) {
  int w = -1;
  int ext[2];
  for (int k = 0; k < 2; ++k) {
    const int lo = *n[2 * k] - *nd[2 * k], hi = *nd[2 * k + 1] - *n[2 * k + 1];
    ext[k] = *n[2 * k + 1] - *n[2 * k] + 1;
    if (lo != hi || lo < 0 || (w >= 0 && lo != w) || ext[k] < 1) {
      fail(who, "boxes are not an interior box grown by the same ghost width in both directions");
      return false;
    }
    w = lo;
  }
  *n1 = ext[0]; *n2 = ext[1]; *ng = w;
  return true;
}

struct DevTmp {  // small synchronous device scratch
  void* p = nullptr;
  explicit DevTmp(size_t bytes) { if (cudaMalloc(&p, bytes) != cudaSuccess) p = nullptr; }
  ~DevTmp() { if (p) cudaFree(p); }
};

// vel1(n1a, n2a, i3, i4) / vel2(n2a, i3, i4, n1a) -> the cell-centre velocity table (n3d, n4d, 2): the advection
// routines read the x / y face velocity at one (x,y) per (i3,i4) (KineticSpeciesF.f:1990, 2009)
__global__ void k_gather_velocities(double* __restrict__ tab, const double* __restrict__ vel1, const double* __restrict__ vel2,
                                    int n1d, int n2d, int n3d, int n4d, int ng) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n3d * n4d) return;
  const int i3 = t % n3d, i4 = t / n3d;
  tab[t] = vel1[ng + (i64)(n1d + 1) * (ng + (i64)n2d * (i3 + (i64)n3d * i4))];
  tab[t + n3d * n4d] = vel2[ng + (i64)(n2d + 1) * (i3 + (i64)n3d * (i4 + (i64)n4d * ng))];
}
// computecurrents (KineticSpeciesF.f:2430-2437): three 4D arrays over the interior
__global__ void k_currents_4d(double* __restrict__ jx, double* __restrict__ jy, double* __restrict__ jz,
                              const double* __restrict__ u, const double* __restrict__ vel, const double* __restrict__ vz,
                              int n1, int n2, int n3, int n4, int ng) {
  const i64 total = (i64)n1 * n2 * n3 * n4;
  const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int n1d = n1 + 2 * ng, n2d = n2 + 2 * ng, n3d = n3 + 2 * ng, n4d = n4 + 2 * ng;
  const int i1 = (int)(t % n1) + ng;
  i64 r = t / n1;
  const int i2 = (int)(r % n2) + ng;
  r /= n2;
  const int i3 = (int)(r % n3) + ng, i4 = (int)(r / n3) + ng;
  const i64 o = i1 + (i64)n1d * (i2 + (i64)n2d * (i3 + (i64)n3d * i4));
  const double f = u[o];
  jx[o] = f * vel[i3 + (i64)n3d * i4];
  jy[o] = f * vel[i3 + (i64)n3d * (i4 + (i64)n4d)];
  jz[o] = f * vz[i1 + (i64)n1d * i2];
}

lk_accel tables_accel(const double* vel3, const double* vel4) {
  lk_accel a;
  memset(&a, 0, sizeof(a));
  a.kind = 2;
  a.field = vel3;
  a.vz = vel4;
  return a;
}

}  // namespace

// Level-0 helpers of lk_capi.cu that have no 4D-geometry public face
extern "C" int lk_neutralize_charge(double* rho, int n1, int n2, int ng, void* stream);
extern "C" int lk_efield_from_potential(double* em, const double* phi, int n1, int n2, int ng, int order, double dx, double dy,
                                        void* stream);

extern "C" {

int lk_f77_status(void) { return g_status; }

#define BOX8(p) {p##1lo, p##1hi, p##2lo, p##2hi, p##3lo, p##3hi, p##4lo, p##4hi}

void xpby4d_(double* x, const double* y, const double* b, const int* nd1lo, const int* nd1hi, const int* nd2lo,
             const int* nd2hi, const int* nd3lo, const int* nd3hi, const int* nd4lo, const int* nd4hi, const int* n1lo,
             const int* n1hi, const int* n2lo, const int* n2hi, const int* n3lo, const int* n3hi, const int* n4lo,
             const int* n4hi) {
  const int* const nd[8] = BOX8(nd);
  const int* const n[8] = BOX8(n);
  lk_geom g;
  if (!geom_from("xpby4d_", nd, n, 0, &g)) return;
  if (check(lk_xpby4d(x, y, *b, &g, nullptr))) check(lk_sync(nullptr));
}

static void phase_space_vel(const char* who, int kind, double* vel3, double* vel4, const int* const nv[8],
                            const int* const ni[8], const double* vxf, const double* vyf, double norm, double bz,
                            const double* field, const double* vz, double* axmax, double* aymax) {
  lk_geom g;
  if (!geom_from(who, nv, ni, 0, &g)) return;
  lk_accel a;
  memset(&a, 0, sizeof(a));
  a.kind = kind;
  a.field = field;
  a.vz = vz;
  a.vxface_velocities = vxf;
  a.vyface_velocities = vyf;
  a.normalization = norm;
  a.bz_const = bz;
  DevTmp out(2 * sizeof(double));
  if (!out.p) return fail(who, "device scratch");
  if (!check(lk_set_phase_space_vel_4d(vel3, vel4, &g, &a, (double*)out.p, nullptr))) return;
  double h[2];
  if (!check(lk_memcpy_d2h(h, out.p, sizeof(h)))) return;
  *axmax = h[0];
  *aymax = h[1];
}
void setphasespacevel4d_(double* vel3, double* vel4, const int* nv1a, const int* nv1b, const int* nv2a, const int* nv2b,
                         const int* nv3a, const int* nv3b, const int* nv4a, const int* nv4b, const int* ni1a,
                         const int* ni1b, const int* ni2a, const int* ni2b, const int* ni3a, const int* ni3b,
                         const int* ni4a, const int* ni4b, const double* vxface_velocities,
                         const double* vyface_velocities, const double* normalization, const double* bz_const,
                         const double* accel, const int* na1a, const int* na1b, const int* na2a, const int* na2b,
                         double* axmax, double* aymax) {
  const int* const nv[8] = {nv1a, nv1b, nv2a, nv2b, nv3a, nv3b, nv4a, nv4b};
  const int* const ni[8] = {ni1a, ni1b, ni2a, ni2b, ni3a, ni3b, ni4a, ni4b};
  if (*na1a != *nv1a || *na1b != *nv1b || *na2a != *nv2a || *na2b != *nv2b)
    return fail("setphasespacevel4d_", "accel must live on the species' (x,y) data box");
  phase_space_vel("setphasespacevel4d_", 0, vel3, vel4, nv, ni, vxface_velocities, vyface_velocities, *normalization, *bz_const,
                  accel, nullptr, axmax, aymax);
}
void setphasespacevelmaxwell4d_(double* vel3, double* vel4, const int* nv1a, const int* nv1b, const int* nv2a,
                                const int* nv2b, const int* nv3a, const int* nv3b, const int* nv4a, const int* nv4b,
                                const int* ni1a, const int* ni1b, const int* ni2a, const int* ni2b, const int* ni3a,
                                const int* ni3b, const int* ni4a, const int* ni4b, const double* vxface_velocities,
                                const double* vyface_velocities, const double* normalization, const double* bz_const,
                                const double* em_vars, const double* vz, double* axmax, double* aymax) {
  const int* const nv[8] = {nv1a, nv1b, nv2a, nv2b, nv3a, nv3b, nv4a, nv4b};
  const int* const ni[8] = {ni1a, ni1b, ni2a, ni2b, ni3a, ni3b, ni4a, ni4b};
  phase_space_vel("setphasespacevelmaxwell4d_", 1, vel3, vel4, nv, ni, vxface_velocities, vyface_velocities, *normalization,
                  *bz_const, em_vars, vz, axmax, aymax);
}

void setaccelerationbcs4d_(double* u, const int* ng1a, const int* ng1b, const int* ng2a, const int* ng2b, const int* ng3a,
                           const int* ng3b, const int* ng4a, const int* ng4b, const int* nl1a, const int* nl1b,
                           const int* nl2a, const int* nl2b, const int* nl3a, const int* nl3b, const int* nl4a,
                           const int* nl4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b, const int* n3a,
                           const int* n3b, const int* n4a, const int* n4b, const int* solution_order, const double* vel3,
                           const double* vel4, const int64_t* ic) {
  (void)ng1a; (void)ng1b; (void)ng2a; (void)ng2b;
  const int* const nl[8] = {nl1a, nl1b, nl2a, nl2b, nl3a, nl3b, nl4a, nl4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  lk_geom g;
  if (!geom_from("setaccelerationbcs4d_", nl, n, *solution_order, &g)) return;
  // does this box touch the global velocity boundaries (KineticSpeciesF.f:1072, 1117)
  const int at[4] = {*ng3a + g.ng == *n3a, *ng3b - g.ng == *n3b, *ng4a + g.ng == *n4a, *ng4b - g.ng == *n4b};
  const lk_accel a = tables_accel(vel3, vel4);
  const lk_inflow* inflow = ic ? (const lk_inflow*)(intptr_t)*ic : nullptr;
  if (check(lk_set_acceleration_bcs_4d(u, &g, &a, inflow, at, nullptr))) check(lk_sync(nullptr));
}

void setadvectionbcs4d_(double* u, const int* ng1a, const int* ng1b, const int* ng2a, const int* ng2b, const int* ng3a,
                        const int* ng3b, const int* ng4a, const int* ng4b, const int* nl1a, const int* nl1b,
                        const int* nl2a, const int* nl2b, const int* nl3a, const int* nl3b, const int* nl4a,
                        const int* nl4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b, const int* n3a,
                        const int* n3b, const int* n4a, const int* n4b, const int* solution_order, const double* vel1,
                        const double* vel2, const int* xperiodic, const int* yperiodic, const int64_t* ic) {
  (void)ng3a; (void)ng3b; (void)ng4a; (void)ng4b;
  const int* const nl[8] = {nl1a, nl1b, nl2a, nl2b, nl3a, nl3b, nl4a, nl4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  lk_geom g;
  if (!geom_from("setadvectionbcs4d_", nl, n, *solution_order, &g)) return;
  const int at[4] = {*ng1a + g.ng == *n1a, *ng1b - g.ng == *n1b, *ng2a + g.ng == *n2a, *ng2b - g.ng == *n2b};
  const int n1d = g.n[0] + 2 * g.ng, n2d = g.n[1] + 2 * g.ng, n3d = g.n[2] + 2 * g.ng, n4d = g.n[3] + 2 * g.ng;
  DevTmp tab(sizeof(double) * 2 * n3d * n4d);
  if (!tab.p) return fail("setadvectionbcs4d_", "device scratch");
  k_gather_velocities<<<(n3d * n4d + 127) / 128, 128>>>((double*)tab.p, vel1, vel2, n1d, n2d, n3d, n4d, g.ng);
  const lk_inflow* inflow = ic ? (const lk_inflow*)(intptr_t)*ic : nullptr;
  if (check(lk_set_advection_bcs_4d(u, &g, (const double*)tab.p, inflow, at, *xperiodic, *yperiodic, nullptr))) check(lk_sync(nullptr));
}

void computeadvectionderivatives4d_(double* rhs, const double* f, const int* nd1a, const int* nd1b, const int* nd2a,
                                    const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b,
                                    const int* n1a, const int* n1b, const int* n2a, const int* n2b, const int* n3a,
                                    const int* n3b, const int* n4a, const int* n4b, const double* vel1,
                                    const double* vel2, const double* deltax, const int* solution_order) {
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  lk_geom g;
  if (!geom_from("computeadvectionderivatives4d_", nd, n, *solution_order, &g)) return;
  for (int k = 0; k < 4; ++k) g.dx[k] = deltax[k];
  const int n1d = g.n[0] + 2 * g.ng, n2d = g.n[1] + 2 * g.ng, n3d = g.n[2] + 2 * g.ng, n4d = g.n[3] + 2 * g.ng;
  DevTmp tab(sizeof(double) * 2 * n3d * n4d);
  if (!tab.p) return fail("computeadvectionderivatives4d_", "device scratch");
  k_gather_velocities<<<(n3d * n4d + 127) / 128, 128>>>((double*)tab.p, vel1, vel2, n1d, n2d, n3d, n4d, g.ng);
  if (check(lk_advection_derivatives_4d(rhs, f, &g, (const double*)tab.p, nullptr))) check(lk_sync(nullptr));
}

void computeaccelerationderivatives4d_(double* rhs, const double* f, const int* nd1a, const int* nd1b, const int* nd2a,
                                       const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a,
                                       const int* nd4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b,
                                       const int* n3a, const int* n3b, const int* n4a, const int* n4b, const double* vel3,
                                       const double* vel4, const double* dx, const int* solution_order) {
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  lk_geom g;
  if (!geom_from("computeaccelerationderivatives4d_", nd, n, *solution_order, &g)) return;
  for (int k = 0; k < 4; ++k) g.dx[k] = dx[k];
  const lk_accel a = tables_accel(vel3, vel4);
  if (check(lk_acceleration_derivatives_4d(rhs, f, &g, &a, nullptr))) check(lk_sync(nullptr));
}

void computecurrents_(const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b,
                      const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b,
                      const int* n3a, const int* n3b, const int* n4a, const int* n4b, const double* velocities,
                      const double* u, const double* vz, double* jx, double* jy, double* jz) {
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  lk_geom g;
  if (!geom_from("computecurrents_", nd, n, 0, &g)) return;
  const i64 total = (i64)g.n[0] * g.n[1] * g.n[2] * g.n[3];
  k_currents_4d<<<(unsigned)((total + 255) / 256), 256>>>(jx, jy, jz, u, velocities, vz, g.n[0], g.n[1], g.n[2], g.n[3], g.ng);
  g_status = (cudaGetLastError() == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess) ? LK_OK : LK_ERR_CUDA;
}

void computekeedot_(const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b,
                    const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b,
                    const int* n3a, const int* n3b, const int* n4a, const int* n4b, const double* xlo, const double* xhi,
                    const double* dx, const double* u, const double* charge, const double* velocities,
                    const double* ext_efield, double* ke_e_dot) {
  (void)xlo; (void)xhi;
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  lk_geom g;
  if (!geom_from("computekeedot_", nd, n, 0, &g)) return;
  for (int k = 0; k < 4; ++k) g.dx[k] = dx[k];
  DevTmp out(sizeof(double));
  if (!out.p) return fail("computekeedot_", "device scratch");
  if (!check(lk_ke_e_dot((double*)out.p, u, &g, *charge, velocities, ext_efield, nullptr))) return;
  check(lk_memcpy_d2h(ke_e_dot, out.p, sizeof(double)));
}

// geometry from the data box alone (the flux routines get no interior box): ghost width from solution_order
static bool geom_from_data(const char* who, const int* const nd[8], int order, const double* dx, lk_geom* g) {
  if (order != 4 && order != 6) {
    fail(who, "solution_order must be 4 or 6");
    return false;
  }
  g->order = order;
  g->ng = (order == 4) ? 2 : 3;
  for (int k = 0; k < 4; ++k) {
    g->n[k] = *nd[2 * k + 1] - *nd[2 * k] + 1 - 2 * g->ng;
    g->dx[k] = dx ? dx[k] : 1.0;
    if (g->n[k] < 1) {
      fail(who, "data box narrower than the ghost layers");
      return false;
    }
  }
  return true;
}

void computeadvectionfluxes4d_(double* flux1, double* flux2, const int* nd1a, const int* nd1b, const int* nd2a,
                               const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b,
                               const double* vel1, const double* vel2, double* face1, double* face2, const double* u,
                               const double* dx, const int* solution_order) {
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  lk_geom g;
  if (!geom_from_data("computeadvectionfluxes4d_", nd, *solution_order, dx, &g)) return;
  if (!check(lk_face_fluxes_4d(flux1, face1, u, &g, vel1, 0, nullptr))) return;
  if (!check(lk_face_fluxes_4d(flux2, face2, u, &g, vel2, 1, nullptr))) return;
  check(cudaDeviceSynchronize() == cudaSuccess ? LK_OK : LK_ERR_CUDA);
}

void computeaccelerationfluxes4d_(double* flux3, double* flux4, const int* nd1a, const int* nd1b, const int* nd2a,
                                  const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b,
                                  const double* vel3, const double* vel4, double* face3, double* face4, const double* u,
                                  const double* dx, const int* solution_order) {
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  lk_geom g;
  if (!geom_from_data("computeaccelerationfluxes4d_", nd, *solution_order, dx, &g)) return;
  if (!check(lk_face_fluxes_4d(flux3, face3, u, &g, vel3, 2, nullptr))) return;
  if (!check(lk_face_fluxes_4d(flux4, face4, u, &g, vel4, 3, nullptr))) return;
  check(cudaDeviceSynchronize() == cudaSuccess ? LK_OK : LK_ERR_CUDA);
}

void accumfluxdiv4d_(double* rhs, const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a,
                     const int* nd3b, const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a,
                     const int* n2b, const int* n3a, const int* n3b, const int* n4a, const int* n4b, const double* fluxx1,
                     const double* fluxx2, const double* fluxx3, const double* fluxx4, const double* deltax) {
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  lk_geom g;
  if (!geom_from("accumfluxdiv4d_", nd, n, 0, &g)) return;
  for (int k = 0; k < 4; ++k) g.dx[k] = deltax[k];
  if (!check(lk_accum_flux_div_4d(rhs, &g, fluxx1, fluxx2, fluxx3, fluxx4, nullptr))) return;
  check(cudaDeviceSynchronize() == cudaSuccess ? LK_OK : LK_ERR_CUDA);
}

void computekeflux_(const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b,
                    const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b,
                    const int* n3a, const int* n3b, const int* n4a, const int* n4b, const int* ng1a, const int* ng1b,
                    const int* ng2a, const int* ng2b, const int* ng3a, const int* ng3b, const int* ng4a, const int* ng4b,
                    const double* dx, const double* face_flux1, const double* face_flux2, const double* face_flux3,
                    const double* face_flux4, const double* velocities, const double* vxface_velocities,
                    const double* vyface_velocities, const int* dir, const int* side, const double* mass, double* ke_flux) {
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  const int* const dom[8] = {ng1a, ng1b, ng2a, ng2b, ng3a, ng3b, ng4a, ng4b};
  lk_geom g;
  if (!geom_from("computekeflux_", nd, n, 0, &g)) return;
  for (int k = 0; k < 4; ++k) g.dx[k] = dx[k];
  if (*dir < 0 || *dir > 3 || *side < 0 || *side > 1) return fail("computekeflux_", "dir / side out of range");
  // `dosum`: only a box that touches that boundary of the domain box sums (KineticSpeciesF.f:2781-2793); either way the
  // incoming value is scaled by mass * ddir (:2889)
  const bool touches = (*side == 0) ? (*n[2 * *dir] == *dom[2 * *dir]) : (*n[2 * *dir + 1] == *dom[2 * *dir + 1]);
  double ddir = 1.0;
  {
    bool first = true;
    for (int d = 0; d < 4; ++d)
      if (d != *dir) {
        ddir = first ? dx[d] : ddir * dx[d];
        first = false;
      }
  }
  if (!touches) {
    *ke_flux = *ke_flux * *mass * ddir;
    g_status = LK_OK;
    return;
  }
  if (*ke_flux != 0.0) return fail("computekeflux_", "ke_flux must come in as 0 (the reference's callers zero it, KineticSpecies.C:2076)");
  const double* fl[4] = {face_flux1, face_flux2, face_flux3, face_flux4};
  DevTmp out(sizeof(double));
  if (!out.p) return fail("computekeflux_", "device scratch");
  if (!check(lk_ke_flux_from_fluxes((double*)out.p, &g, fl[*dir], velocities, vxface_velocities, vyface_velocities, *dir, *side,
                                    *mass, nullptr)))
    return;
  check(lk_memcpy_d2h(ke_flux, out.p, sizeof(double)));
}

void computekevelspaceflux_(const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a,
                            const int* nd3b, const int* nd4a, const int* nd4b, const int* n1a, const int* n1b,
                            const int* n2a, const int* n2b, const int* n3a, const int* n3b, const int* n4a, const int* n4b,
                            const int* ng1a, const int* ng1b, const int* ng2a, const int* ng2b, const int* ng3a,
                            const int* ng3b, const int* ng4a, const int* ng4b, const double* dx, const double* face_flux3,
                            const double* face_flux4, double* ke_flux, const double* mass, const double* vxface_velocities,
                            const double* vyface_velocities, const int* side, const int* dir) {
  (void)ng1a; (void)ng1b; (void)ng2a; (void)ng2b;
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  lk_geom g;
  if (!geom_from("computekevelspaceflux_", nd, n, 0, &g)) return;
  for (int k = 0; k < 4; ++k) g.dx[k] = dx[k];
  g_status = LK_OK;
  if (*dir != 2 && *dir != 3) return;   // the routine has no other branch (:2930, 2959)
  const bool touches = (*dir == 2) ? ((*side == 0) ? (*n3a == *ng3a) : (*n3b == *ng3b)) : ((*side == 0) ? (*n4a == *ng4a) : (*n4b == *ng4b));
  if (!touches) return;
  if (!check(lk_ke_vel_space_flux(ke_flux, &g, (*dir == 2) ? face_flux3 : face_flux4, vxface_velocities, vyface_velocities, *dir,
                                  *side, *mass, nullptr)))
    return;
  check(cudaDeviceSynchronize() == cudaSuccess ? LK_OK : LK_ERR_CUDA);
}

// The twilight-zone routines (TZSourceF.f, ElectronTZSourceF.f, TwoSpecies_ElectronTZSourceF.f, TwoSpecies_IonTZSourceF.f)
// share one argument list; only the data box is passed and the kernels run over all of it, so the split into interior and
// ghosts does not matter (order 4's is assumed).  f, soln, error and velocities are device arrays; xlo, xhi, dx, dparams
// are the host-side heads PROBLEMDOMAIN_TO_FORT and the source classes pass.
static void tz_call(const char* who, int kind, double* out, const double* soln, const int* const nd[8], const double* xlo,
                    const double* dx, const double* time, const double* velocities, const double* dparams) {
  lk_geom g;
  if (!geom_from_data(who, nd, 4, dx, &g)) return;
  const double params[3] = {dparams[0], kind >= 2 ? dparams[1] : 1.0, kind >= 2 ? dparams[2] : 1.0};
  int64_t count = 0;
  if (!check(lk_trig_tz_table_count(&g, &count))) return;
  double* tab = nullptr;
  if (!check(lk_malloc((void**)&tab, sizeof(double) * count))) return;
  const int lo[2] = {*nd[0], *nd[2]};
  const double x0[2] = {xlo[0], xlo[1]};
  if (check(lk_trig_tz_tables(tab, &g, lo, x0, velocities, kind, params, nullptr))) {
    const int st = soln ? lk_compute_trig_tz_source_error(out, soln, &g, tab, velocities, *time, kind, params, nullptr)
                        : lk_set_trig_tz_source(out, &g, tab, velocities, *time, kind, params, nullptr);
    if (check(st)) check(lk_sync(nullptr));
  }
  lk_free(tab);
}
#define LK_TZ_PAIR(set_name, err_name, kind)                                                                                  \
  void set_name(double* f, const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a,               \
                const int* nd3b, const int* nd4a, const int* nd4b, const double* xlo, const double* xhi, const double* dx,    \
                const double* time, const double* velocities, const double* dparams) {                                        \
    (void)xhi;                                                                                                                \
    const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};                                                \
    tz_call(#set_name, kind, f, nullptr, nd, xlo, dx, time, velocities, dparams);                                            \
  }                                                                                                                           \
  void err_name(double* error, const double* soln, const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b,        \
                const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b, const double* xlo, const double* xhi,     \
                const double* dx, const double* time, const double* velocities, const double* dparams) {                      \
    (void)xhi;                                                                                                                \
    const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};                                                \
    tz_call(#err_name, kind, error, soln, nd, xlo, dx, time, velocities, dparams);                                           \
  }
LK_TZ_PAIR(settrigtzsource_, computetrigtzsourceerror_, 0)
LK_TZ_PAIR(setelectrontrigtzsource_, computeelectrontrigtzsourceerror_, 1)
LK_TZ_PAIR(settwoelectrontrigtzsource_, computetwoelectrontrigtzsourceerror_, 2)
LK_TZ_PAIR(settwoiontrigtzsource_, computetwoiontrigtzsourceerror_, 3)
#undef LK_TZ_PAIR

void appendkrook_(const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b,
                  const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b,
                  const int* n3a, const int* n3b, const int* n4a, const int* n4b, const double* dt, const int64_t* ic,
                  const double* nu, const double* u, double* rhs) {
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  lk_geom g;
  if (!geom_from("appendkrook_", nd, n, 0, &g)) return;
  const lk_inflow* inflow = ic ? (const lk_inflow*)(intptr_t)*ic : nullptr;
  if (check(lk_append_krook(rhs, u, &g, nu, *dt, inflow, nullptr))) check(lk_sync(nullptr));
}

void appendpitchanglecollision_(double* rhs, const double* f, const double* velocities, const double* ivx, const double* ivy,
                                const double* vth, const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b,
                                const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b, const int* n1a,
                                const int* n1b, const int* n2a, const int* n2b, const int* n3a, const int* n3b,
                                const int* n4a, const int* n4b, const double* xlo, const double* xhi, const double* dx,
                                const double* range_lo, const double* range_hi, const double* dparams, const int* iparams) {
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  lk_geom g;
  if (!geom_from("appendpitchanglecollision_", nd, n, iparams[1], &g)) return;
  if (iparams[2] != 0) return fail("appendpitchanglecollision_", "do_relativity = 1 is not supported");
  for (int k = 0; k < 4; ++k) g.dx[k] = dx[k];
  lk_pitch_angle p;
  for (int k = 0; k < 2; ++k) {
    p.range_lo[k] = range_lo[k];
    p.range_hi[k] = range_hi[k];
  }
  p.vfloor = dparams[0];       // PitchAngleCollisionOperator.H: VFLOOR, VTHERMAL_DT, NU
  p.vthermal_dt = dparams[1];
  p.nu_coef = dparams[2];
  p.conservative = (iparams[0] == 1) ? 1 : 0;
  if (check(lk_append_pitch_angle_collision(rhs, f, &g, velocities, ivx, ivy, vth, xlo + 2, xhi + 2, &p, nullptr)))
    check(lk_sync(nullptr));
}

void computepitchanglespeciesmoments_(double* rn, double* rgammax, double* rgammay, const double* u, const int* nd1a,
                                      const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b,
                                      const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a,
                                      const int* n2b, const int* n3a, const int* n3b, const int* n4a, const int* n4b,
                                      const double* velocities) {
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  lk_geom g;
  if (!geom_from("computepitchanglespeciesmoments_", nd, n, 0, &g)) return;
  int64_t launches = 0;
  if (lkcoll::moments(rn, rgammax, rgammay, u, &g, velocities, nullptr, &launches) != cudaSuccess) {
    check(LK_ERR_CUDA);
    return;
  }
  check(lk_sync(nullptr));
}

void computepitchanglespecieskec_(double* rkec, const double* rvx0, const double* rvy0, const double* u, const int* nd1a,
                                  const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b,
                                  const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a,
                                  const int* n2b, const int* n3a, const int* n3b, const int* n4a, const int* n4b,
                                  const double* velocities) {
  const int* const nd[8] = {nd1a, nd1b, nd2a, nd2b, nd3a, nd3b, nd4a, nd4b};
  const int* const n[8] = {n1a, n1b, n2a, n2b, n3a, n3b, n4a, n4b};
  lk_geom g;
  if (!geom_from("computepitchanglespecieskec_", nd, n, 0, &g)) return;
  int64_t launches = 0;
  if (lkcoll::kec(rkec, rvx0, rvy0, u, &g, velocities, nullptr, &launches) != cudaSuccess) {
    check(LK_ERR_CUDA);
    return;
  }
  check(lk_sync(nullptr));
}

static void reduced_2d(const char* who, int mode, double* a, double* b, const double* n, const double* gx, const double* gy,
                       const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b) {
  const int64_t n1d = (int64_t)*nd1b - *nd1a + 1, n2d = (int64_t)*nd2b - *nd2a + 1;
  if (n1d < 1 || n2d < 1) return fail(who, "empty data box");
  int64_t launches = 0;
  if (lkcoll::reduced(mode, a, b, n, gx, gy, n1d * n2d, nullptr, &launches) != cudaSuccess) {
    check(LK_ERR_CUDA);
    return;
  }
  check(lk_sync(nullptr));
}
void computepitchanglespeciesreducedfields_(double* vx, double* vy, const double* n, const double* gammax,
                                            const double* gammay, const int* nd1a, const int* nd1b, const int* nd2a,
                                            const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a,
                                            const int* nd4b) {
  (void)nd3a; (void)nd3b; (void)nd4a; (void)nd4b;
  reduced_2d("computepitchanglespeciesreducedfields_", 0, vx, vy, n, gammax, gammay, nd1a, nd1b, nd2a, nd2b);
}
void computepitchanglespeciesvthermal_(double* vthsq, const double* kec, const double* n, const int* nd1a, const int* nd1b,
                                       const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a,
                                       const int* nd4b) {
  (void)nd3a; (void)nd3b; (void)nd4a; (void)nd4b;
  reduced_2d("computepitchanglespeciesvthermal_", 1, vthsq, nullptr, n, kec, nullptr, nd1a, nd1b, nd2a, nd2b);
}

void neutralizecharge4d_(const int* md1a, const int* md1b, const int* md2a, const int* md2b, const int* n1a, const int* n1b,
                         const int* n2a, const int* n2b, double* rhs, const int* comm) {
  (void)comm;
  const int* const md[4] = {md1a, md1b, md2a, md2b};
  const int* const n[4] = {n1a, n1b, n2a, n2b};
  int n1, n2, ng;
  if (!geom2_from("neutralizecharge4d_", md, n, &n1, &n2, &ng)) return;
  if (check(lk_neutralize_charge(rhs, n1, n2, ng, nullptr))) check(lk_sync(nullptr));
}

void computeefieldfrompotential_(const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* n1a,
                                 const int* n1b, const int* n2a, const int* n2b, const int* solution_order,
                                 const int* em_vars_dim, const double* dx, double* emvars, const double* phi) {
  const int* const nd[4] = {nd1a, nd1b, nd2a, nd2b};
  const int* const n[4] = {n1a, n1b, n2a, n2b};
  int n1, n2, ng;
  if (!geom2_from("computeefieldfrompotential_", nd, n, &n1, &n2, &ng)) return;
  if (*em_vars_dim < 2) return fail("computeefieldfrompotential_", "em_vars_dim < 2");
  if (check(lk_efield_from_potential(emvars, phi, n1, n2, ng, *solution_order, dx[0], dx[1], nullptr))) check(lk_sync(nullptr));
}

void maxwellevalrhs_(const int* md1a, const int* md1b, const int* md2a, const int* md2b, const int* m1a, const int* m1b,
                     const int* m2a, const int* m2b, const double* xlo, const double* xhi, const double* dx,
                     const double* c, const double* avweak, const double* avstrong, const int* solution_order,
                     const double* supergrid_lo, const double* supergrid_hi, const double* emvars, const double* jx,
                     const double* jy, const double* jz, double* demvars) {
  const int* const md[4] = {md1a, md1b, md2a, md2b};
  const int* const m[4] = {m1a, m1b, m2a, m2b};
  int n1, n2, ng;
  if (!geom2_from("maxwellevalrhs_", md, m, &n1, &n2, &ng)) return;
  // the supergrid stretching (SGMetricFunction, MaxwellF.f:393-438) is not built: the layer must be empty
  for (int d = 0; d < 2; ++d)
    if (supergrid_lo[d] > xlo[d] || supergrid_hi[d] < xhi[d])
      return fail("maxwellevalrhs_", "a supergrid layer inside the domain is not supported");
  if (check(lk_maxwell_rhs(demvars, emvars, jx, jy, jz, n1, n2, ng, *solution_order, dx, *c, *avweak, *avstrong, nullptr)))
    check(lk_sync(nullptr));
}

extern int lk_maxwell_vz_rhs(double* dvz, const double* em, int n1, int n2, int ng, double charge_per_mass, void* stream);
void maxwellevalvzrhs_(const int* md1a, const int* md1b, const int* md2a, const int* md2b, const int* m1a, const int* m1b,
                       const int* m2a, const int* m2b, const double* charge_per_mass, const double* emvars, double* dvz) {
  const int* const md[4] = {md1a, md1b, md2a, md2b};
  const int* const m[4] = {m1a, m1b, m2a, m2b};
  int n1, n2, ng;
  if (!geom2_from("maxwellevalvzrhs_", md, m, &n1, &n2, &ng)) return;
  if (check(lk_maxwell_vz_rhs(dvz, emvars, n1, n2, ng, *charge_per_mass, nullptr))) check(lk_sync(nullptr));
}

void xpby2d_(double* x, const double* y, const double* b, const int* nd1a, const int* nd1b, const int* nd2a,
             const int* nd2b, const int* n1a, const int* n1b, const int* n2a, const int* n2b, const int* dim) {
  const int* const nd[4] = {nd1a, nd1b, nd2a, nd2b};
  const int* const n[4] = {n1a, n1b, n2a, n2b};
  int n1, n2, ng;
  if (!geom2_from("xpby2d_", nd, n, &n1, &n2, &ng)) return;
  if (check(lk_xpby2d(x, y, *b, n1, n2, ng, *dim, nullptr))) check(lk_sync(nullptr));
}

// MaxwellF.H:26-37, :77-89, :105-133.  Arrays on the device; boxes, nx / ny, flags and c by reference on the host.
void zeroghost2d_(double* u, const int* n1a, const int* n1b, const int* n2a, const int* n2b, const int* nd1a, const int* nd1b,
                  const int* nd2a, const int* nd2b, const int* dim) {
  const int* const nd[4] = {nd1a, nd1b, nd2a, nd2b};
  const int* const n[4] = {n1a, n1b, n2a, n2b};
  int n1, n2, ng;
  if (!geom2_from("zeroghost2d_", nd, n, &n1, &n2, &ng)) return;
  if (check(lk_zero_ghost_2d(u, n1, n2, ng, *dim, nullptr))) check(lk_sync(nullptr));
}
void maxwelladdantennasource_(const int* md1a, const int* md1b, const int* md2a, const int* md2b, const int* m1a, const int* m1b,
                              const int* m2a, const int* m2b, const double* xlo, const double* xhi, const double* dx,
                              const double* antenna_source, double* dEMvars) {
  (void)xlo; (void)xhi; (void)dx;
  const int* const nd[4] = {md1a, md1b, md2a, md2b};
  const int* const n[4] = {m1a, m1b, m2a, m2b};
  int n1, n2, ng;
  if (!geom2_from("maxwelladdantennasource_", nd, n, &n1, &n2, &ng)) return;
  if (check(lk_maxwell_add_antenna_source(dEMvars, antenna_source, n1, n2, ng, nullptr))) check(lk_sync(nullptr));
}
static bool bc_geom(const char* who, const int* const nd[4], const int* const n[4], const int* nx, const int* ny, const int* order,
                    int* n1, int* n2, int at[4]) {
  int ng;
  if (!geom2_from(who, nd, n, n1, n2, &ng)) return false;
  if ((*order == 4 ? 2 : 3) != ng) {
    fail(who, "ghost width does not match solution_order");
    return false;
  }
  at[0] = *n[0] == 0;
  at[1] = *n[1] == *nx - 1;
  at[2] = *n[2] == 0;
  at[3] = *n[3] == *ny - 1;
  return true;
}
void maxwellsetembcs_(const int* md1a, const int* md1b, const int* md2a, const int* md2b, const int* m1a, const int* m1b,
                      const int* m2a, const int* m2b, double* EMvars, const int* nx, const int* ny, const int* xPeriodic,
                      const int* yPeriodic, const int* solution_order, const double* c) {
  const int* const nd[4] = {md1a, md1b, md2a, md2b};
  const int* const n[4] = {m1a, m1b, m2a, m2b};
  int n1, n2, at[4];
  if (!bc_geom("maxwellsetembcs_", nd, n, nx, ny, solution_order, &n1, &n2, at)) return;
  if (check(lk_maxwell_set_em_bcs(EMvars, n1, n2, *solution_order, at, *xPeriodic, *yPeriodic, *c, nullptr))) check(lk_sync(nullptr));
}
void maxwellsetvzbcs_(const int* md1a, const int* md1b, const int* md2a, const int* md2b, const int* m1a, const int* m1b,
                      const int* m2a, const int* m2b, double* vz, const int* nx, const int* ny, const int* xPeriodic,
                      const int* yPeriodic, const int* solution_order) {
  const int* const nd[4] = {md1a, md1b, md2a, md2b};
  const int* const n[4] = {m1a, m1b, m2a, m2b};
  int n1, n2, at[4];
  if (!bc_geom("maxwellsetvzbcs_", nd, n, nx, ny, solution_order, &n1, &n2, at)) return;
  if (check(lk_maxwell_set_vz_bcs(vz, n1, n2, *solution_order, at, *xPeriodic, *yPeriodic, nullptr))) check(lk_sync(nullptr));
}

}  // extern "C"
