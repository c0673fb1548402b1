// lk_diag.cu -- time-history diagnostics on the device: the species kinetic energies and
// momenta of computeke / computekemaxwell (KineticSpeciesF.f:2447-2559) and the field histories of
// Poisson::accumulateSequences (Poisson.C:796-860) / Maxwell::accumulateSequences (Maxwell.C:753-875).
// The reference sums cell by cell; here every sum is a deterministic two-level tree (warp shuffles, fixed
// block order), so results agree with the oracle to rounding (tests: 1e-13 relative), not bit for bit.
// One pass over f (8 B/cell), called at sequence_write_times only -- not on the stage path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/loki_b200.h"

namespace lkdiag {

typedef long long i64;
constexpr int KE_BLOCKS = 148 * 4;

__device__ __forceinline__ double warp_sum(double s) {
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}
__device__ __forceinline__ double warp_max(double s) {
  for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
  return s;
}

// part[b][0..4] = sum over the block's cells of {0.5 u vx^2, 0.5 u vy^2, u vx, u vy, 0.5 u vz^2}
__global__ void k_ke_partial(lk_geom g, const double* __restrict__ f, const double* __restrict__ vel,
                             const double* __restrict__ vz, double* __restrict__ part) {
  __shared__ double sh[5][8];
  const int ng = g.ng;
  const i64 n1d = g.n[0] + 2 * ng, n2d = g.n[1] + 2 * ng, n3d = g.n[2] + 2 * ng, n4d = g.n[3] + 2 * ng;
  const i64 total = (i64)g.n[0] * g.n[1] * g.n[2] * g.n[3];
  double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const int i1 = (int)(t % g.n[0]) + ng;
    i64 r = t / g.n[0];
    const int i2 = (int)(r % g.n[1]) + ng;
    r /= g.n[1];
    const int i3 = (int)(r % g.n[2]) + ng, i4 = (int)(r / g.n[2]) + ng;
    const double vx = __ldg(vel + i3 + n3d * i4), vy = __ldg(vel + i3 + n3d * (i4 + n4d));
    const double u = f[i1 + n1d * (i2 + n2d * (i3 + n3d * i4))];
    a[0] += 0.5 * u * (vx * vx);
    a[1] += 0.5 * u * (vy * vy);
    a[2] += u * vx;
    a[3] += u * vy;
    if (vz) {
      const double w = __ldg(vz + i1 + n1d * i2);
      a[4] += 0.5 * u * (w * w);
    }
  }
  for (int k = 0; k < 5; ++k) {
    const double s = warp_sum(a[k]);
    if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double b = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) b += sh[threadIdx.x][w];
    part[blockIdx.x * 5 + threadIdx.x] = b;
  }
}
// out = {ke, ke_x, ke_y, px, py}; with vz: ke = ke_x + ke_y + ke_z, px = py = 0 (computekemaxwell has none)
__global__ void k_ke_finish(const double* __restrict__ part, int nblocks, double scale, int maxwell, double* __restrict__ out) {
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int b = 0; b < nblocks; ++b)
    for (int k = 0; k < 5; ++k) s[k] += part[b * 5 + k];
  for (int k = 0; k < 5; ++k) s[k] *= scale;
  out[1] = s[0];
  out[2] = s[1];
  if (maxwell) {
    out[0] = s[0] + s[1] + s[4];
    out[3] = 0.0;
    out[4] = 0.0;
  } else {
    out[0] = s[0] + s[1];
    out[3] = s[2];
    out[4] = s[3];
  }
}

// one CTA: {max |E|, sum |E| * area, max |Ex|, max |Ey| [, max |Ez|], 0.5 sum |E|^2 * area} per field triple
__global__ void k_field_history(const double* __restrict__ em, int n1, int n2, int ng, int ncomp, double area,
                                double* __restrict__ out) {
  __shared__ double sh[32];
  const i64 n1d = n1 + 2 * ng, pl = n1d * (n2 + 2 * ng);
  const int total = n1 * n2;
  const int nh = (ncomp == 6) ? 2 : 1, nc = (ncomp == 6) ? 3 : 2;
  auto block = [&](double v, bool is_max) -> double {
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double b = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) b = is_max ? fmax(b, sh[w]) : b + sh[w];
    return b;
  };
  for (int h = 0; h < nh; ++h) {
    double s = 0.0, mx = 0.0, tot = 0.0, cm[3] = {0.0, 0.0, 0.0};
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
      const i64 o = (t % n1 + ng) + n1d * (t / n1 + ng) + pl * (3 * h);
      double tmp = 0.0;
      for (int c = 0; c < nc; ++c) {
        const double a = em[o + pl * c];
        tmp += a * a;
        cm[c] = fmax(cm[c], fabs(a));
      }
      const double loc = sqrt(tmp);
      s += 0.5 * tmp;
      mx = fmax(mx, loc);
      tot += loc;
    }
    const double bs = block(s, false), bm = block(mx, true), bt = block(tot, false);
    double bc[3];
    for (int c = 0; c < nc; ++c) bc[c] = block(cm[c], true);
    if (threadIdx.x == 0) {
      double* o = out + (nc + 3) * h;
      o[0] = bm;
      o[1] = bt * area;
      for (int c = 0; c < nc; ++c) o[2 + c] = bc[c];
      o[2 + nc] = bs * area;
    }
  }
}

int ke_scratch_doubles() { return KE_BLOCKS * 5; }
cudaError_t compute_ke(double* out5, const double* f, const lk_geom* g, double mass, const double* velocities,
                       const double* vz, double* scratch, cudaStream_t st, int64_t* launches) {
  const i64 total = (i64)g->n[0] * g->n[1] * g->n[2] * g->n[3];
  int blocks = (int)((total + 255) / 256);
  if (blocks > KE_BLOCKS) blocks = KE_BLOCKS;
  k_ke_partial<<<blocks, 256, 0, st>>>(*g, f, velocities, vz, scratch);
  // ke_x*mass*dx*dy*dvx*dvy, left to right (KineticSpeciesF.f:2492-2495)
  const double scale = mass * g->dx[0] * g->dx[1] * g->dx[2] * g->dx[3];
  k_ke_finish<<<1, 1, 0, st>>>(scratch, blocks, scale, vz != nullptr, out5);
  *launches += 2;
  return cudaGetLastError();
}
cudaError_t field_history(double* out, const double* em, int n1, int n2, int ng, int ncomp, double dx, double dy,
                          cudaStream_t st, int64_t* launches) {
  k_field_history<<<1, 1024, 0, st>>>(em, n1, n2, ng, ncomp, dx * dy, out);
  *launches += 1;
  return cudaGetLastError();
}

}  // namespace lkdiag
