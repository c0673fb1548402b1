// lk_flux.cu -- flux-form diagnostics: the species kinetic-energy flux through the eight
// phase-space boundaries that KineticSpecies::accumulateSequencesCommon (KineticSpecies.C:2052-2097) adds to the
// time histories, and the routines it is made of:
//
//   * WENO43Avg4D / WENO65Avg4D + computeFlux4D (KineticSpeciesF.f:630-720, 797-910, 2359-2396) as called by
//     computeadvectionfluxes4D (:1838-1945) and computeaccelerationfluxes4D (:2249-2355): face fits and
//     flux = vel * face on the ROTATED face arrays of KineticSpecies.C:1569-1584, with the reference's index ranges
//     (fits over faces w .. nd-w of the direction and the whole data box across; products over 2 .. extent-3 of all
//     four rotated extents whatever the order; everything else untouched);
//   * accumfluxdiv4D (:985-1032), vy term divided by dvx as there (:1024);
//   * computekeflux (:2734-2893) and computekevelspaceflux (:2897-2990) from materialised flux arrays (the Fortran-ABI
//     entry points use these);
//   * the PRODUCT path, lk_ke_flux_boundaries: the same eight numbers without any flux array -- a boundary flux needs
//     one face fit per boundary cell, so each is a pass over one 3D slab of f (the reference materialises four full
//     4D face arrays, four flux arrays and reads one slab of each).
//
// Built with -fmad=false and the strict flavour of the fit (the reference's operation order): every face value and
// flux is bit for bit the reference's in both arithmetic modes.  Sums over a boundary are deterministic two-level
// trees, not the reference's sequential sums (tests: 1e-13 relative).  Not on the stage path: called at
// sequence_write_times only.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/loki_b200.h"
#define LK_STRICT 1
#include "lk_device.cuh"

namespace lkflux {

using lkstrict::DAccel;
using lkstrict::DGeo;
typedef long long i64;
constexpr int FLUX_BLOCKS = 148 * 4;

static DGeo make_geo(const lk_geom* g) {
  DGeo d;
  i64 s = 1;
  for (int k = 0; k < 4; ++k) {
    d.n[k] = g->n[k];
    d.nd[k] = g->n[k] + 2 * g->ng;
    d.dx[k] = g->dx[k];
    d.s[k] = s;
    s *= d.nd[k];
  }
  d.ng = g->ng;
  d.order = g->order;
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

// rotated extents e[k] = nd[(d+k)%4] (+1 for k = 0) and the cell stride of rotated position k
struct Rot {
  i64 e[4], cs[4];
};
static Rot make_rot(const DGeo& g, int d) {
  Rot r;
  for (int k = 0; k < 4; ++k) {
    r.e[k] = g.nd[(d + k) % 4] + (k == 0 ? 1 : 0);
    r.cs[k] = g.s[(d + k) % 4];
  }
  return r;
}

__device__ __forceinline__ double fit_at(const double* __restrict__ c, i64 s, int order, bool pos) {
  // c: the cell above the face
  if (order == 4) return lkstrict::weno43(c[-2 * s], c[-s], c[0], c[s], pos);
  return lkstrict::weno65(c[-3 * s], c[-2 * s], c[-s], c[0], c[s], c[2 * s], pos);
}

// faces j0 = w .. e0-1-w, all j1, j2, j3
__global__ void k_face_fit(Rot r, int order, const double* __restrict__ u, const double* __restrict__ vel,
                           double* __restrict__ face) {
  const int w = (order == 4) ? 2 : 3;
  const i64 m0 = r.e[0] - 2 * w;
  const i64 total = m0 * r.e[1] * r.e[2] * r.e[3];
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const i64 j0 = t % m0 + w;
    i64 q = t / m0;
    const i64 j1 = q % r.e[1];
    q /= r.e[1];
    const i64 j2 = q % r.e[2], j3 = q / r.e[2];
    const i64 fi = j0 + r.e[0] * (j1 + r.e[1] * (j2 + r.e[2] * j3));
    const double v = vel[fi];
    face[fi] = fit_at(u + j0 * r.cs[0] + j1 * r.cs[1] + j2 * r.cs[2] + j3 * r.cs[3], r.cs[0], order, v > 0.0);
  }
}
// computeFlux4D: faces 2 .. extent-3 of all four rotated extents
__global__ void k_face_flux(Rot r, const double* __restrict__ vel, const double* __restrict__ face,
                            double* __restrict__ flux) {
  const i64 m0 = r.e[0] - 4, m1 = r.e[1] - 4, m2 = r.e[2] - 4, m3 = r.e[3] - 4;
  const i64 total = m0 * m1 * m2 * m3;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const i64 j0 = t % m0 + 2;
    i64 q = t / m0;
    const i64 j1 = q % m1 + 2;
    q /= m1;
    const i64 j2 = q % m2 + 2, j3 = q / m2 + 2;
    const i64 fi = j0 + r.e[0] * (j1 + r.e[1] * (j2 + r.e[2] * j3));
    flux[fi] = vel[fi] * face[fi];
  }
}

__global__ void k_accum_flux_div(DGeo g, double* __restrict__ rhs, const double* __restrict__ f1,
                                 const double* __restrict__ f2, const double* __restrict__ f3,
                                 const double* __restrict__ f4) {
  const i64 total = (i64)g.n[0] * g.n[1] * g.n[2] * g.n[3];
  const i64 e1 = g.nd[0] + 1, e2 = g.nd[1] + 1, e3 = g.nd[2] + 1, e4 = g.nd[3] + 1;
  const double dx = g.dx[0], dy = g.dx[1], dvx = g.dx[2];
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const int i1 = (int)(t % g.n[0]) + g.ng;
    i64 q = t / g.n[0];
    const int i2 = (int)(q % g.n[1]) + g.ng;
    q /= g.n[1];
    const int i3 = (int)(q % g.n[2]) + g.ng, i4 = (int)(q / g.n[2]) + g.ng;
    const i64 a1 = i1 + e1 * (i2 + (i64)g.nd[1] * (i3 + (i64)g.nd[2] * i4));
    const i64 a2 = i2 + e2 * (i3 + (i64)g.nd[2] * (i4 + (i64)g.nd[3] * i1));
    const i64 a3 = i3 + e3 * (i4 + (i64)g.nd[3] * (i1 + (i64)g.nd[0] * i2));
    const i64 a4 = i4 + e4 * (i1 + (i64)g.nd[0] * (i2 + (i64)g.nd[1] * i3));
    // -(..)/dx - (..)/dy - (..)/dvx - (..)/dvx, left to right (KineticSpeciesF.f:1021-1024)
    const double temp = -(f1[a1 + 1] - f1[a1]) / dx - (f2[a2 + 1] - f2[a2]) / dy - (f3[a3 + 1] - f3[a3]) / dvx -
                        (f4[a4 + 1] - f4[a4]) / dvx;
    rhs[lkstrict::gidx(g, i1, i2, i3, i4)] = temp;
  }
}

// ---- boundary sums -------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double s, double* sh) {
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  double b = 0.0;
  if (threadIdx.x == 0)
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) b += sh[k];
  __syncthreads();
  return b;  // thread 0 only
}
__global__ void k_sum_finish(const double* __restrict__ part, int n, double mass, double ddir, double* __restrict__ out) {
  double s = 0.0;
  for (int k = 0; k < n; ++k) s += part[k];
  out[0] = s * mass * ddir;  // ke_flux*mass*ddir, left to right (KineticSpeciesF.f:2889)
}

// v2 of the boundary term and the three loop extents; (ia, ib, ic) = the interior indices of the three directions
// other than `dir`, fastest first in memory order
struct Slab {
  int dir, fidx;
  int m[3];  // interior extents of the other three directions (ascending direction number)
};
__device__ __forceinline__ double slab_v2(const DGeo& g, const Slab& sl, const double* __restrict__ velocities,
                                          const double* __restrict__ vxf, const double* __restrict__ vyf, int i3, int i4) {
  double vx, vy;
  if (sl.dir <= 1) {
    vx = velocities[i3 + (i64)g.nd[2] * i4];
    vy = velocities[i3 + (i64)g.nd[2] * (i4 + (i64)g.nd[3])];
  } else if (sl.dir == 2) {
    vx = vxf[sl.fidx + (i64)(g.nd[2] + 1) * i4];
    vy = vxf[sl.fidx + (i64)(g.nd[2] + 1) * (i4 + (i64)g.nd[3])];
  } else {
    vx = vyf[i3 + (i64)g.nd[2] * sl.fidx];
    vy = vyf[i3 + (i64)g.nd[2] * (sl.fidx + (i64)(g.nd[3] + 1))];
  }
  return vx * vx + vy * vy;
}
// decode a slab cell: the three free indices, with the face index in direction dir
__device__ __forceinline__ void slab_cell(const DGeo& g, const Slab& sl, i64 t, int idx[4]) {
  int o[3], k = 0;
  for (int d = 0; d < 4; ++d)
    if (d != sl.dir) o[k++] = d;
  idx[o[0]] = (int)(t % sl.m[0]) + g.ng;
  t /= sl.m[0];
  idx[o[1]] = (int)(t % sl.m[1]) + g.ng;
  idx[o[2]] = (int)(t / sl.m[1]) + g.ng;
  idx[sl.dir] = sl.fidx;
}
__device__ __forceinline__ i64 rot_index(const DGeo& g, int d, const int idx[4]) {
  // rotated face array of direction d: (idx[d], idx[d+1], idx[d+2], idx[d+3]) with extents (nd[d]+1, nd[d+1], ...)
  return idx[d] + (i64)(g.nd[d] + 1) * (idx[(d + 1) & 3] + (i64)g.nd[(d + 1) & 3] * (idx[(d + 2) & 3] + (i64)g.nd[(d + 2) & 3] * idx[(d + 3) & 3]));
}

// computekeflux from a materialised flux array of direction sl.dir
__global__ void k_ke_flux_arrays(DGeo g, Slab sl, const double* __restrict__ flux, const double* __restrict__ velocities,
                                 const double* __restrict__ vxf, const double* __restrict__ vyf, double* __restrict__ part) {
  __shared__ double sh[8];
  const i64 total = (i64)sl.m[0] * sl.m[1] * sl.m[2];
  double s = 0.0;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    int idx[4];
    slab_cell(g, sl, t, idx);
    s += 0.5 * flux[rot_index(g, sl.dir, idx)] * slab_v2(g, sl, velocities, vxf, vyf, idx[2], idx[3]);
  }
  const double b = block_sum(s, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = b;
}
// the same with the flux formed on the fly: vel at the face (velocity tables in x, y; the acceleration in vx, vy)
// times the face fit of f
__global__ void k_ke_flux_fused(DGeo g, DAccel a, Slab sl, const double* __restrict__ f,
                                const double* __restrict__ velocities, double* __restrict__ part) {
  __shared__ double sh[8];
  const i64 total = (i64)sl.m[0] * sl.m[1] * sl.m[2];
  double s = 0.0;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    int idx[4];
    slab_cell(g, sl, t, idx);
    double vel;
    if (sl.dir == 0) vel = velocities[idx[2] + (i64)g.nd[2] * idx[3]];                                   // initializeVelocity
    else if (sl.dir == 1) vel = velocities[idx[2] + (i64)g.nd[2] * (idx[3] + (i64)g.nd[3])];
    else if (sl.dir == 2) vel = lkstrict::accel_x(a, g, idx[0], idx[1], idx[2], idx[3]);
    else vel = lkstrict::accel_y(a, g, idx[0], idx[1], idx[2], idx[3]);
    const double face = fit_at(f + lkstrict::gidx(g, idx[0], idx[1], idx[2], idx[3]), g.s[sl.dir], g.order, vel > 0.0);
    s += 0.5 * (vel * face) * slab_v2(g, sl, velocities, a.vxf, a.vyf, idx[2], idx[3]);
  }
  const double b = block_sum(s, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = b;
}
// computekevelspaceflux: ke_flux(i1,i2) += 0.5*mass*flux*v2*ddir summed along the boundary line, sequentially in the
// reference's order (one thread per (i1,i2): bit for bit)
__global__ void k_ke_vel_space_flux(DGeo g, Slab sl, const double* __restrict__ flux, const double* __restrict__ vxf,
                                    const double* __restrict__ vyf, double mass, double ddir, double* __restrict__ ke) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.n[0] * g.n[1]) return;
  int idx[4];
  idx[0] = t % g.n[0] + g.ng;
  idx[1] = t / g.n[0] + g.ng;
  const int other = (sl.dir == 2) ? 3 : 2;
  idx[sl.dir] = sl.fidx;
  double acc = ke[idx[0] + (i64)g.nd[0] * idx[1]];
  for (int k = g.ng; k < g.ng + g.n[other]; ++k) {
    idx[other] = k;
    const double v2 = slab_v2(g, sl, nullptr, vxf, vyf, idx[2], idx[3]);
    acc = acc + 0.5 * mass * flux[rot_index(g, sl.dir, idx)] * v2 * ddir;
  }
  ke[idx[0] + (i64)g.nd[0] * idx[1]] = acc;
}

static Slab make_slab(const DGeo& g, int dir, int side) {
  Slab sl;
  sl.dir = dir;
  sl.fidx = side == 0 ? g.ng : g.ng + g.n[dir];
  int k = 0;
  for (int d = 0; d < 4; ++d)
    if (d != dir) sl.m[k++] = g.n[d];
  return sl;
}
static double slab_ddir(const DGeo& g, int dir) {
  // dx(2)*dx(3)*dx(4) etc., left to right (KineticSpeciesF.f:2780, 2808, 2836, 2864)
  double p = 1.0;
  bool first = true;
  for (int d = 0; d < 4; ++d)
    if (d != dir) {
      p = first ? g.dx[d] : p * g.dx[d];
      first = false;
    }
  return p;
}
static inline unsigned nblk(i64 n, int t, i64 cap) {
  i64 b = (n + t - 1) / t;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace lkflux

using namespace lkflux;

// scratch for the per-block partial sums, one per (device, stream)
#include <map>
#include <mutex>
#include <utility>
static std::mutex g_flux_mu;
static std::map<std::pair<int, void*>, double*> g_flux_part;
static double* flux_part(void* stream) {
  std::lock_guard<std::mutex> lk(g_flux_mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  double*& p = g_flux_part[std::make_pair(dev, stream)];
  if (!p && cudaMalloc(&p, sizeof(double) * FLUX_BLOCKS) != cudaSuccess) p = nullptr;
  return p;
}

extern "C" {

int lk_face_fluxes_4d(double* flux, double* face, const double* u, const lk_geom* g, const double* vel, int dir,
                      void* stream) {
  if (!flux || !face || !u || !g || !vel || dir < 0 || dir > 3 || (g->order != 4 && g->order != 6)) return LK_ERR_ARG;
  for (int k = 0; k < 4; ++k)
    if (g->n[k] < 1) return LK_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  DGeo d = make_geo(g);
  Rot r = make_rot(d, dir);
  const int w = (g->order == 4) ? 2 : 3;
  const i64 nfit = (r.e[0] - 2 * w) * r.e[1] * r.e[2] * r.e[3];
  const i64 nflux = (r.e[0] - 4) * (r.e[1] - 4) * (r.e[2] - 4) * (r.e[3] - 4);
  k_face_fit<<<nblk(nfit, 256, 148 * 32), 256, 0, st>>>(r, g->order, u, vel, face);
  k_face_flux<<<nblk(nflux, 256, 148 * 32), 256, 0, st>>>(r, vel, face, flux);
  return cudaGetLastError() == cudaSuccess ? LK_OK : LK_ERR_CUDA;
}

int lk_accum_flux_div_4d(double* rhs, const lk_geom* g, const double* flux1, const double* flux2, const double* flux3,
                         const double* flux4, void* stream) {
  if (!rhs || !g || !flux1 || !flux2 || !flux3 || !flux4) return LK_ERR_ARG;
  DGeo d = make_geo(g);
  const i64 total = (i64)g->n[0] * g->n[1] * g->n[2] * g->n[3];
  k_accum_flux_div<<<nblk(total, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(d, rhs, flux1, flux2, flux3, flux4);
  return cudaGetLastError() == cudaSuccess ? LK_OK : LK_ERR_CUDA;
}

int lk_ke_flux_from_fluxes(double* out_dev, const lk_geom* g, const double* flux, const double* velocities,
                           const double* vxface_velocities, const double* vyface_velocities, int dir, int side,
                           double mass, void* stream) {
  if (!out_dev || !g || !flux || dir < 0 || dir > 3 || side < 0 || side > 1) return LK_ERR_ARG;
  if ((dir <= 1 && !velocities) || (dir == 2 && !vxface_velocities) || (dir == 3 && !vyface_velocities)) return LK_ERR_ARG;
  double* part = flux_part(stream);
  if (!part) return LK_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  DGeo d = make_geo(g);
  Slab sl = make_slab(d, dir, side);
  const unsigned nb = nblk((i64)sl.m[0] * sl.m[1] * sl.m[2], 256, FLUX_BLOCKS);
  k_ke_flux_arrays<<<nb, 256, 0, st>>>(d, sl, flux, velocities, vxface_velocities, vyface_velocities, part);
  k_sum_finish<<<1, 1, 0, st>>>(part, (int)nb, mass, slab_ddir(d, dir), out_dev);
  return cudaGetLastError() == cudaSuccess ? LK_OK : LK_ERR_CUDA;
}

int lk_ke_vel_space_flux(double* ke_flux_xy, const lk_geom* g, const double* flux, const double* vxface_velocities,
                         const double* vyface_velocities, int dir, int side, double mass, void* stream) {
  if (!ke_flux_xy || !g || !flux || (dir != 2 && dir != 3) || side < 0 || side > 1) return LK_ERR_ARG;
  if ((dir == 2 && !vxface_velocities) || (dir == 3 && !vyface_velocities)) return LK_ERR_ARG;
  DGeo d = make_geo(g);
  Slab sl = make_slab(d, dir, side);
  const double ddir = (dir == 2) ? g->dx[3] : g->dx[2];   // KineticSpeciesF.f:2934, 2963
  const int n = g->n[0] * g->n[1];
  k_ke_vel_space_flux<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d, sl, flux, vxface_velocities, vyface_velocities, mass,
                                                                      ddir, ke_flux_xy);
  return cudaGetLastError() == cudaSuccess ? LK_OK : LK_ERR_CUDA;
}

int lk_ke_flux_boundaries(double* out8_dev, const double* f, const lk_geom* g, const double* velocities, const lk_accel* a,
                          double mass, const int at_boundary[8], void* stream) {
  if (!out8_dev || !f || !g || !velocities || !a || !at_boundary) return LK_ERR_ARG;
  if (!a->vxface_velocities || !a->vyface_velocities || (a->kind != 2 && !a->field)) return LK_ERR_ARG;
  double* part = flux_part(stream);
  if (!part) return LK_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  DGeo d = make_geo(g);
  DAccel da = make_accel(a);
  if (cudaMemsetAsync(out8_dev, 0, sizeof(double) * 8, st) != cudaSuccess) return LK_ERR_CUDA;
  for (int dir = 0; dir < 4; ++dir)
    for (int side = 0; side < 2; ++side) {
      if (!at_boundary[2 * dir + side]) continue;  // `dosum = 0`: this box does not touch that boundary
      Slab sl = make_slab(d, dir, side);
      const unsigned nb = nblk((i64)sl.m[0] * sl.m[1] * sl.m[2], 256, FLUX_BLOCKS);
      k_ke_flux_fused<<<nb, 256, 0, st>>>(d, da, sl, f, velocities, part);
      k_sum_finish<<<1, 1, 0, st>>>(part, (int)nb, mass, slab_ddir(d, dir), out8_dev + 2 * dir + side);
    }
  return cudaGetLastError() == cudaSuccess ? LK_OK : LK_ERR_CUDA;
}

}  // extern "C"
