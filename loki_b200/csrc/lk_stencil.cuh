// lk_stencil.cuh -- the fused Vlasov RHS stencil kernel (SURVEY 8a rows a3+a4/a5+a8+a10/a11).
//
// One CTA owns a 4D tile (T0 x T1 x T2 x T3 cells in x,y,vx,vy).  It stages the tile plus its STAR
// halo (the tile grown by ng along one axis at a time -- the WENO stencils are axis aligned, so the
// corners of the grown box are never read) in shared memory, then makes four LINE SWEEPS, one per
// direction.  In a sweep each thread walks one grid line of the tile with a sliding register window:
// one shared-memory load and ONE WENO fit per face, the left face of a cell being the right face of
// its predecessor exactly as in the reference's `uLeft = uRight` loops (KineticSpeciesF.f:1990-2005,
// 2137-2152).  A line of T cells costs T+1 fits, so the redundant work is (T+1)/T per direction
// instead of the 2x of a thread-per-cell kernel.  Partial sums live in a shared accumulator tile in
// the reference's order ((x + y) + vx) + vy; the last sweep adds its term and applies the RK stage
// update straight to global memory, so f is read once and rhs never touches HBM.
//
// The kernel is fp64-FMA bound on B200 (about 45 fp64 ops per order-4 fit, 4.6 fits per cell) rather
// than HBM bound; see DESIGN.md for both ceilings.
#pragma once
#include "lk_device.cuh"

namespace LK_NS {

template <int ORDER, int T0, int T1, int T2, int T3>
struct TileCfg {
  static constexpr int NG = (ORDER == 4) ? 2 : 3;
  static constexpr int W = 2 * NG;                       // window = stencil width of one fit
  static constexpr int PC = ((T0 + W) % 2 == 0) ? (T0 + W + 1) : (T0 + W);  // odd pitch: conflict-free x sweep
  static constexpr int PR = (T0 % 2 == 0) ? (T0 + 1) : T0;
  static constexpr int NC = T3 * T2 * T1 * PC;           // core + x halo
  static constexpr int NY = T3 * T2 * W * T0;            // y halo
  static constexpr int NV = T3 * W * T1 * T0;            // vx halo
  static constexpr int NW = W * T2 * T1 * T0;            // vy halo
  static constexpr int NR = T3 * T2 * T1 * PR;           // rhs accumulators
  static constexpr int SMEM_DOUBLES = NC + NY + NV + NW + NR;
  static constexpr size_t SMEM_BYTES = sizeof(double) * SMEM_DOUBLES;
};

template <int ORDER>
__device__ __forceinline__ double fit_window(const double (&w)[(ORDER == 4) ? 4 : 6], bool pos) {
  if constexpr (ORDER == 4) return weno43(w[0], w[1], w[2], w[3], pos);
  else return weno65(w[0], w[1], w[2], w[3], w[4], w[5], pos);
}

// FULL: flags == 3 (advection + acceleration, no accumulate) -- the production instantiation.
template <int ORDER, int T0, int T1, int T2, int T3, int NT, bool FULL>
__global__ void __launch_bounds__(NT, 2)
k_stencil_tiled(const DGeo g, const double* __restrict__ f, const double* __restrict__ vel, const DAccel a,
                const DUpd upd, double* __restrict__ rhs_out, const int flags, const int nt0, const int nt1,
                const int nt2) {
  using C = TileCfg<ORDER, T0, T1, T2, T3>;
  constexpr int NG = C::NG, W = C::W, PC = C::PC, PR = C::PR;
  extern __shared__ double smem[];
  double* sC = smem;
  double* sY = sC + C::NC;
  double* sV = sY + C::NY;
  double* sW = sV + C::NV;
  double* sR = sW + C::NW;

  const int tid = threadIdx.x;
  // tile origin (interior coordinates); x tiles fastest so that neighbouring CTAs share halos in L2
  int b = blockIdx.x;
  const int o0 = (b % nt0) * T0; b /= nt0;
  const int o1 = (b % nt1) * T1; b /= nt1;
  const int o2 = (b % nt2) * T2; b /= nt2;
  const int o3 = b * T3;
  const int ng = g.ng;  // == NG
  const i64 base = gidx(g, o0 + ng, o1 + ng, o2 + ng, o3 + ng);  // data index of tile cell (0,0,0,0)
  const double* fb = f + base;

  // ---------------- stage tile + star halo ----------------
  // core + x halo: element (k, b1, c, d), k in [0, T0+W) <-> x cell o0 + k - NG
  for (int e = tid; e < T3 * T2 * T1 * (T0 + W); e += NT) {
    const int k = e % (T0 + W);
    int r = e / (T0 + W);
    const int b1 = r % T1; r /= T1;
    const int c = r % T2;
    const int d = r / T2;
    const bool ok = (o0 + k - NG < g.n[0] + NG) && (o1 + b1 < g.n[1] + NG) && (o2 + c < g.n[2] + NG) && (o3 + d < g.n[3] + NG);
    double v = 0.0;
    if (ok) v = fb[(k - NG) + g.s[1] * b1 + g.s[2] * c + g.s[3] * d];
    sC[((d * T2 + c) * T1 + b1) * PC + k] = v;
  }
  // y halo: (a0, h, c, d), h in [0,W): h<NG -> y cell o1 - NG + h ; else o1 + T1 + (h-NG)
  for (int e = tid; e < C::NY; e += NT) {
    const int a0 = e % T0;
    int r = e / T0;
    const int h = r % W; r /= W;
    const int c = r % T2;
    const int d = r / T2;
    const int yb = (h < NG) ? (h - NG) : (T1 + h - NG);
    const bool ok = (o0 + a0 < g.n[0] + NG) && (o1 + yb < g.n[1] + NG) && (o2 + c < g.n[2] + NG) && (o3 + d < g.n[3] + NG);
    double v = 0.0;
    if (ok) v = fb[a0 + g.s[1] * yb + g.s[2] * c + g.s[3] * d];
    sY[e] = v;  // layout [d][c][h][a0]
  }
  // vx halo: (a0, b1, h, d)
  for (int e = tid; e < C::NV; e += NT) {
    const int a0 = e % T0;
    int r = e / T0;
    const int b1 = r % T1; r /= T1;
    const int h = r % W;
    const int d = r / W;
    const int cb = (h < NG) ? (h - NG) : (T2 + h - NG);
    const bool ok = (o0 + a0 < g.n[0] + NG) && (o1 + b1 < g.n[1] + NG) && (o2 + cb < g.n[2] + NG) && (o3 + d < g.n[3] + NG);
    double v = 0.0;
    if (ok) v = fb[a0 + g.s[1] * b1 + g.s[2] * cb + g.s[3] * d];
    sV[e] = v;  // layout [d][h][b1][a0]
  }
  // vy halo: (a0, b1, c, h)
  for (int e = tid; e < C::NW; e += NT) {
    const int a0 = e % T0;
    int r = e / T0;
    const int b1 = r % T1; r /= T1;
    const int c = r % T2;
    const int h = r / T2;
    const int db = (h < NG) ? (h - NG) : (T3 + h - NG);
    const bool ok = (o0 + a0 < g.n[0] + NG) && (o1 + b1 < g.n[1] + NG) && (o2 + c < g.n[2] + NG) && (o3 + db < g.n[3] + NG);
    double v = 0.0;
    if (ok) v = fb[a0 + g.s[1] * b1 + g.s[2] * c + g.s[3] * db];
    sW[e] = v;  // layout [h][c][b1][a0]
  }
  __syncthreads();

  const double rdx0 = 1.0 / g.dx[0], rdx1 = 1.0 / g.dx[1], rdx2 = 1.0 / g.dx[2], rdx3 = 1.0 / g.dx[3];
  const bool do_adv = FULL || (flags & 1);
  const bool do_acc = FULL || (flags & 2);

  // ---------------- x sweep: lines (b1, c, d) ----------------
  for (int l = tid; l < T1 * T2 * T3; l += NT) {
    const int b1 = l % T1;
    const int c = (l / T1) % T2;
    const int d = l / (T1 * T2);
    double* rrow = sR + ((d * T2 + c) * T1 + b1) * PR;
    if (do_adv) {
      const int i3 = min(o2 + c, g.n[2] - 1) + ng, i4 = min(o3 + d, g.n[3] - 1) + ng;
      const double vx = __ldg(vel + i3 + (i64)g.nd[2] * i4);
      const bool pos = vx > 0.0;
      const double* row = sC + ((d * T2 + c) * T1 + b1) * PC;
      double w[W];
#pragma unroll
      for (int k = 0; k < W; ++k) w[k] = row[k];
      double uL = fit_window<ORDER>(w, pos);
#pragma unroll
      for (int a0 = 0; a0 < T0; ++a0) {
#pragma unroll
        for (int k = 0; k < W - 1; ++k) w[k] = w[k + 1];
        w[W - 1] = row[a0 + W];
        const double uR = fit_window<ORDER>(w, pos);
        double init = 0.0;
        if (!FULL && (flags & 4)) {
          const bool ok = (o0 + a0 < g.n[0]) && (o1 + b1 < g.n[1]) && (o2 + c < g.n[2]) && (o3 + d < g.n[3]);
          if (ok) init = rhs_out[base + a0 + g.s[1] * b1 + g.s[2] * c + g.s[3] * d];
          rrow[a0] = init - flux_diff(vx, uR, uL, g.dx[0], rdx0);
        } else {
          rrow[a0] = -flux_diff(vx, uR, uL, g.dx[0], rdx0);
        }
        uL = uR;
      }
    } else {
#pragma unroll
      for (int a0 = 0; a0 < T0; ++a0) {
        double init = 0.0;
        const bool ok = (o0 + a0 < g.n[0]) && (o1 + b1 < g.n[1]) && (o2 + c < g.n[2]) && (o3 + d < g.n[3]);
        if ((flags & 4) && ok) init = rhs_out[base + a0 + g.s[1] * b1 + g.s[2] * c + g.s[3] * d];
        rrow[a0] = init;
      }
    }
  }
  __syncthreads();

  // ---------------- y sweep: lines (a0, c, d) ----------------
  if (do_adv) {
    for (int l = tid; l < T0 * T2 * T3; l += NT) {
      const int a0 = l % T0;
      const int c = (l / T0) % T2;
      const int d = l / (T0 * T2);
      const int i3 = min(o2 + c, g.n[2] - 1) + ng, i4 = min(o3 + d, g.n[3] - 1) + ng;
      const double vy = __ldg(vel + i3 + (i64)g.nd[2] * (i4 + (i64)g.nd[3]));
      const bool pos = vy > 0.0;
      const double* core = sC + ((d * T2 + c) * T1) * PC + NG + a0;  // + b1*PC
      const double* halo = sY + ((d * T2 + c) * W) * T0 + a0;        // + h*T0
      double* racc = sR + ((d * T2 + c) * T1) * PR + a0;             // + b1*PR
      // line position k in [0, T1+W): k<NG -> halo h=k ; k<NG+T1 -> core b1=k-NG ; else halo h=k-T1
      auto ld = [&](int k) -> double {
        if (k < NG) return halo[k * T0];
        if (k < NG + T1) return core[(k - NG) * PC];
        return halo[(k - T1) * T0];
      };
      double w[W];
#pragma unroll
      for (int k = 0; k < W; ++k) w[k] = ld(k);
      double uL = fit_window<ORDER>(w, pos);
#pragma unroll
      for (int b1 = 0; b1 < T1; ++b1) {
#pragma unroll
        for (int k = 0; k < W - 1; ++k) w[k] = w[k + 1];
        w[W - 1] = ld(b1 + W);
        const double uR = fit_window<ORDER>(w, pos);
        racc[b1 * PR] = racc[b1 * PR] - flux_diff(vy, uR, uL, g.dx[1], rdx1);
        uL = uR;
      }
    }
  }
  __syncthreads();

  // ---------------- vx sweep: lines (a0, b1, d) ----------------
  if (do_acc) {
    for (int l = tid; l < T0 * T1 * T3; l += NT) {
      const int a0 = l % T0;
      const int b1 = (l / T0) % T1;
      const int d = l / (T0 * T1);
      const int i1 = min(o0 + a0, g.n[0] - 1) + ng, i2 = min(o1 + b1, g.n[1] - 1) + ng;
      const int i4 = min(o3 + d, g.n[3] - 1) + ng;
      const double* core = sC + (d * T2 * T1 + b1) * PC + NG + a0;  // + c*T1*PC
      const double* halo = sV + (d * W * T1 + b1) * T0 + a0;        // + h*T1*T0
      double* racc = sR + (d * T2 * T1 + b1) * PR + a0;             // + c*T1*PR
      auto ld = [&](int k) -> double {
        if (k < NG) return halo[k * T1 * T0];
        if (k < NG + T2) return core[(k - NG) * T1 * PC];
        return halo[(k - T2) * T1 * T0];
      };
      // the face below the first cell was fitted by the cell below it with ITS coefficient, unless
      // that cell is outside the interior (KineticSpeciesF.f:2137-2141)
      const int i3first = o2 + ng;
      const double axl = accel_x(a, g, i1, i2, (o2 > 0) ? (i3first - 1) : i3first, i4);
      double w[W];
#pragma unroll
      for (int k = 0; k < W; ++k) w[k] = ld(k);
      double uL = fit_window<ORDER>(w, axl > 0.0);
#pragma unroll
      for (int c = 0; c < T2; ++c) {
#pragma unroll
        for (int k = 0; k < W - 1; ++k) w[k] = w[k + 1];
        w[W - 1] = ld(c + W);
        const double ax = accel_x(a, g, i1, i2, min(i3first + c, g.n[2] - 1 + ng), i4);
        const double uR = fit_window<ORDER>(w, ax > 0.0);
        racc[c * T1 * PR] = racc[c * T1 * PR] - flux_diff(ax, uR, uL, g.dx[2], rdx2);
        uL = uR;
      }
    }
  }
  __syncthreads();

  // ---------------- vy sweep + epilogue: lines (a0, b1, c) ----------------
  for (int l = tid; l < T0 * T1 * T2; l += NT) {
    const int a0 = l % T0;
    const int b1 = (l / T0) % T1;
    const int c = l / (T0 * T1);
    const bool ok3 = (o0 + a0 < g.n[0]) && (o1 + b1 < g.n[1]) && (o2 + c < g.n[2]);
    const i64 gofs = base + a0 + g.s[1] * b1 + g.s[2] * c;  // + d*s[3]
    const double* racc = sR + (c * T1 + b1) * PR + a0;      // + d*T2*T1*PR
    // prefetch the RK operands so their latency hides behind the fits
    double fo[T3], di[T3];
    if (upd.active) {
#pragma unroll
      for (int d = 0; d < T3; ++d) {
        const bool ok = ok3 && (o3 + d < g.n[3]);
        fo[d] = ok ? upd.f_old[gofs + g.s[3] * d] : 0.0;
        di[d] = (ok && upd.delta_in) ? upd.delta_in[gofs + g.s[3] * d] : 0.0;
      }
    }
    double res[T3];
    if (do_acc) {
      const int i1 = min(o0 + a0, g.n[0] - 1) + ng, i2 = min(o1 + b1, g.n[1] - 1) + ng;
      const int i3 = min(o2 + c, g.n[2] - 1) + ng;
      const double* core = sC + (c * T1 + b1) * PC + NG + a0;  // + d*T2*T1*PC
      const double* halo = sW + (c * T1 + b1) * T0 + a0;       // + h*T2*T1*T0
      auto ld = [&](int k) -> double {
        if (k < NG) return halo[k * T2 * T1 * T0];
        if (k < NG + T3) return core[(k - NG) * T2 * T1 * PC];
        return halo[(k - T3) * T2 * T1 * T0];
      };
      const int i4first = o3 + ng;
      const double ayl = accel_y(a, g, i1, i2, i3, (o3 > 0) ? (i4first - 1) : i4first);
      double w[W];
#pragma unroll
      for (int k = 0; k < W; ++k) w[k] = ld(k);
      double uL = fit_window<ORDER>(w, ayl > 0.0);
#pragma unroll
      for (int d = 0; d < T3; ++d) {
#pragma unroll
        for (int k = 0; k < W - 1; ++k) w[k] = w[k + 1];
        w[W - 1] = ld(d + W);
        const double ay = accel_y(a, g, i1, i2, i3, min(i4first + d, g.n[3] - 1 + ng));
        const double uR = fit_window<ORDER>(w, ay > 0.0);
        res[d] = racc[d * T2 * T1 * PR] - flux_diff(ay, uR, uL, g.dx[3], rdx3);
        uL = uR;
      }
    } else {
#pragma unroll
      for (int d = 0; d < T3; ++d) res[d] = racc[d * T2 * T1 * PR];
    }
#pragma unroll
    for (int d = 0; d < T3; ++d) {
      const bool ok = ok3 && (o3 + d < g.n[3]);
      if (!ok) continue;
      const i64 idx = gofs + g.s[3] * d;
      if (rhs_out) rhs_out[idx] = res[d];
      if (upd.active) {
        double dl = upd.w_delta * res[d];
        if (upd.delta_in) dl = di[d] + dl;
        if (upd.delta_out) upd.delta_out[idx] = dl;
        const double inc = upd.use_delta ? dl : res[d];
        upd.pred[idx] = fo[d] + upd.c_pred * inc;
      }
    }
  }
}

template <int ORDER, int T0, int T1, int T2, int T3, int NT>
static cudaError_t launch_tiled_cfg(const DGeo& g, const double* f, const double* vel, const DAccel& a, const DUpd& u,
                                    double* rhs_out, int flags, cudaStream_t st) {
  using C = TileCfg<ORDER, T0, T1, T2, T3>;
  const int nt0 = (g.n[0] + T0 - 1) / T0, nt1 = (g.n[1] + T1 - 1) / T1, nt2 = (g.n[2] + T2 - 1) / T2,
            nt3 = (g.n[3] + T3 - 1) / T3;
  const long long tiles = (long long)nt0 * nt1 * nt2 * nt3;
  if (tiles > 0x7fffffffLL) return cudaErrorNotSupported;
  cudaError_t e;
  if (flags == 3) {
    auto kern = k_stencil_tiled<ORDER, T0, T1, T2, T3, NT, true>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)tiles, NT, C::SMEM_BYTES, st>>>(g, f, vel, a, u, rhs_out, flags, nt0, nt1, nt2);
  } else {
    auto kern = k_stencil_tiled<ORDER, T0, T1, T2, T3, NT, false>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)tiles, NT, C::SMEM_BYTES, st>>>(g, f, vel, a, u, rhs_out, flags, nt0, nt1, nt2);
  }
  return cudaGetLastError();
}

static cudaError_t launch_stencil_tiled(const DGeo& g, const double* f, const double* vel, const DAccel& a,
                                        const DUpd& u, double* rhs_out, int flags, cudaStream_t st) {
  if (g.order == 4) return launch_tiled_cfg<4, 8, 8, 8, 4, 256>(g, f, vel, a, u, rhs_out, flags, st);
  return launch_tiled_cfg<6, 8, 8, 8, 4, 256>(g, f, vel, a, u, rhs_out, flags, st);
}

}  // namespace LK_NS
