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
#include <cuda.h>

#include "lk_device.cuh"

namespace LK_NS {

template <int ORDER, int T0, int T1, int T2, int T3>
struct TileCfg {
  static constexpr int NG = (ORDER == 4) ? 2 : 3;
  static constexpr int W = 2 * NG;                       // window = stencil width of one fit
  static constexpr int PC = T0 + W;                      // dense rows: the TMA box is written as is
  static constexpr int PR = (T0 % 2 == 0) ? (T0 + 1) : T0;  // odd pitch: conflict-free accumulators
  static constexpr int NC = T3 * T2 * T1 * PC;           // core + x halo      box (T0+W, T1, T2, T3)
  // TMA needs the innermost start coordinate 16-byte aligned: with ng = 3 the halo slabs start one
  // cell early (XO) and are two cells wider (HX), so their x origin is even.
  static constexpr int XO = NG % 2;
  static constexpr int HX = T0 + 2 * XO;
  static constexpr int NY = T3 * T2 * NG * HX;           // one y halo slab    box (HX, NG, T2, T3)
  static constexpr int NV = T3 * NG * T1 * HX;           // one vx halo slab   box (HX, T1, NG, T3)
  static constexpr int NW = NG * T2 * T1 * HX;           // one vy halo slab   box (HX, T1, T2, NG)
  static constexpr int NR = T3 * T2 * T1 * PR;           // rhs accumulators
  static constexpr int SMEM_DOUBLES = NC + 2 * NY + 2 * NV + 2 * NW + NR;
  static constexpr size_t SMEM_BYTES = sizeof(double) * SMEM_DOUBLES + 128 + 16;  // + alignment slack + mbarrier
  static constexpr unsigned TMA_BYTES = sizeof(double) * (NC + 2 * NY + 2 * NV + 2 * NW);
};

// the four tensor maps of one distribution array (one per box shape)
struct TileMaps {
  CUtensorMap core, yh, vh, wh;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, void* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int ORDER>
__device__ __forceinline__ double fit_window(const double (&w)[(ORDER == 4) ? 4 : 6], bool pos) {
  if constexpr (ORDER == 4) return weno43(w[0], w[1], w[2], w[3], pos);
  else return weno65(w[0], w[1], w[2], w[3], w[4], w[5], pos);
}

// FULL: flags == 3 (advection + acceleration, no accumulate) -- the production instantiation.
// TMA: the tile and its six halo slabs arrive by seven cp.async.bulk.tensor copies issued by one
// thread (out-of-box elements are zero-filled by the hardware); otherwise plain loads (fallback for
// rows whose byte pitch is not a multiple of 16).
template <int ORDER, int T0, int T1, int T2, int T3, int NT, bool FULL, bool TMA>
__global__ void __launch_bounds__(NT, 2)
k_stencil_tiled(const DGeo g, const double* __restrict__ f, const double* __restrict__ vel, const DAccel a,
                const DUpd upd, double* __restrict__ rhs_out, const int flags, const int nt0, const int nt1,
                const int nt2, const __grid_constant__ TileMaps maps) {
  using C = TileCfg<ORDER, T0, T1, T2, T3>;
  constexpr int NG = C::NG, W = C::W, PC = C::PC, PR = C::PR, XO = C::XO, HX = C::HX;
  extern __shared__ unsigned char smem_raw[];
  double* smem = (double*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  double* sC = smem;
  double* sYl = sC + C::NC;
  double* sYh = sYl + C::NY;
  double* sVl = sYh + C::NY;
  double* sVh = sVl + C::NV;
  double* sWl = sVh + C::NV;
  double* sWh = sWl + C::NW;
  double* sR = sWh + C::NW;
  unsigned long long* bar = (unsigned long long*)(sR + C::NR);

  const int tid = threadIdx.x;
  // tile origin (interior coordinates); x tiles fastest so that neighbouring CTAs share halos in L2
  int b = blockIdx.x;
  const int o0 = (b % nt0) * T0; b /= nt0;
  const int o1 = (b % nt1) * T1; b /= nt1;
  const int o2 = (b % nt2) * T2; b /= nt2;
  const int o3 = b * T3;
  const int ng = g.ng;  // == NG
  const i64 base = gidx(g, o0 + ng, o1 + ng, o2 + ng, o3 + ng);  // data index of tile cell (0,0,0,0)

  // pull this tile's rows of the RK operands towards L2 now; the epilogue reads them ~10 us later
  if (upd.active) {
    for (int row = tid; row < T1 * T2 * T3; row += NT) {
      const int b1 = row % T1, c = (row / T1) % T2, d = row / (T1 * T2);
      if ((o1 + b1 < g.n[1]) && (o2 + c < g.n[2]) && (o3 + d < g.n[3])) {
        const i64 o = base + g.s[1] * b1 + g.s[2] * c + g.s[3] * d;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(upd.f_old + o));
        if (upd.delta_in) asm volatile("prefetch.global.L2 [%0];" ::"l"(upd.delta_in + o));
      }
    }
  }

  // ---------------- stage tile + star halo ----------------
  if (TMA) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(C::TMA_BYTES) : "memory");
      // data-box coordinates of tile cell (0,0,0,0) are (o + ng)
      const int c0 = o0 + ng, c1 = o1 + ng, c2 = o2 + ng, c3 = o3 + ng;
      tma_load_4d(sC, &maps.core, bar, c0 - NG, c1, c2, c3);
      tma_load_4d(sYl, &maps.yh, bar, c0 - XO, c1 - NG, c2, c3);
      tma_load_4d(sYh, &maps.yh, bar, c0 - XO, c1 + T1, c2, c3);
      tma_load_4d(sVl, &maps.vh, bar, c0 - XO, c1, c2 - NG, c3);
      tma_load_4d(sVh, &maps.vh, bar, c0 - XO, c1, c2 + T2, c3);
      tma_load_4d(sWl, &maps.wh, bar, c0 - XO, c1, c2, c3 - NG);
      tma_load_4d(sWh, &maps.wh, bar, c0 - XO, c1, c2, c3 + T3);
    }
    __syncthreads();  // barrier initialised before anyone polls it
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
          : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    }
  } else {
    const double* fb = f + base;
    auto inbox = [&](int q0, int q1, int q2, int q3) {
      return (o0 + q0 < g.n[0] + NG) && (o1 + q1 < g.n[1] + NG) && (o2 + q2 < g.n[2] + NG) && (o3 + q3 < g.n[3] + NG);
    };
    for (int e = tid; e < C::NC; e += NT) {  // (k, b1, c, d): x cell o0 + k - NG
      const int k = e % PC;
      int r = e / PC;
      const int b1 = r % T1; r /= T1;
      const int c = r % T2;
      const int d = r / T2;
      sC[e] = inbox(k - NG, b1, c, d) ? fb[(k - NG) + g.s[1] * b1 + g.s[2] * c + g.s[3] * d] : 0.0;
    }
    for (int e = tid; e < 2 * C::NY; e += NT) {  // [side][d][c][h][a0]
      const int side = e / C::NY, q = e % C::NY;
      const int a0 = q % HX - XO;
      int r = q / HX;
      const int h = r % NG; r /= NG;
      const int c = r % T2;
      const int d = r / T2;
      const int yb = side ? (T1 + h) : (h - NG);
      (side ? sYh : sYl)[q] = inbox(a0, yb, c, d) ? fb[a0 + g.s[1] * yb + g.s[2] * c + g.s[3] * d] : 0.0;
    }
    for (int e = tid; e < 2 * C::NV; e += NT) {  // [side][d][h][b1][a0]
      const int side = e / C::NV, q = e % C::NV;
      const int a0 = q % HX - XO;
      int r = q / HX;
      const int b1 = r % T1; r /= T1;
      const int h = r % NG;
      const int d = r / NG;
      const int cb = side ? (T2 + h) : (h - NG);
      (side ? sVh : sVl)[q] = inbox(a0, b1, cb, d) ? fb[a0 + g.s[1] * b1 + g.s[2] * cb + g.s[3] * d] : 0.0;
    }
    for (int e = tid; e < 2 * C::NW; e += NT) {  // [side][h][c][b1][a0]
      const int side = e / C::NW, q = e % C::NW;
      const int a0 = q % HX - XO;
      int r = q / HX;
      const int b1 = r % T1; r /= T1;
      const int c = r % T2;
      const int h = r / T2;
      const int db = side ? (T3 + h) : (h - NG);
      (side ? sWh : sWl)[q] = inbox(a0, b1, c, db) ? fb[a0 + g.s[1] * b1 + g.s[2] * c + g.s[3] * db] : 0.0;
    }
    __syncthreads();
  }

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
          rrow[a0] = sub_flux(init, vx, uR, uL, g.dx[0], rdx0);
        } else {
          rrow[a0] = sub_flux(0.0, vx, uR, uL, g.dx[0], rdx0);
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
      const double* hlo = sYl + ((d * T2 + c) * NG) * HX + XO + a0;  // + h*HX
      const double* hhi = sYh + ((d * T2 + c) * NG) * HX + XO + a0;
      double* racc = sR + ((d * T2 + c) * T1) * PR + a0;             // + b1*PR
      // line position k in [0, T1+W): k<NG -> halo h=k ; k<NG+T1 -> core b1=k-NG ; else halo h=k-T1
      auto ld = [&](int k) -> double {
        if (k < NG) return hlo[k * HX];
        if (k < NG + T1) return core[(k - NG) * PC];
        return hhi[(k - NG - T1) * HX];
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
        racc[b1 * PR] = sub_flux(racc[b1 * PR], vy, uR, uL, g.dx[1], rdx1);
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
      const double* hlo = sVl + (d * NG * T1 + b1) * HX + XO + a0;  // + h*T1*HX
      const double* hhi = sVh + (d * NG * T1 + b1) * HX + XO + a0;
      double* racc = sR + (d * T2 * T1 + b1) * PR + a0;             // + c*T1*PR
      auto ld = [&](int k) -> double {
        if (k < NG) return hlo[k * T1 * HX];
        if (k < NG + T2) return core[(k - NG) * T1 * PC];
        return hhi[(k - NG - T2) * T1 * HX];
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
        racc[c * T1 * PR] = sub_flux(racc[c * T1 * PR], ax, uR, uL, g.dx[2], rdx2);
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
      const double* hlo = sWl + (c * T1 + b1) * HX + XO + a0;  // + h*T2*T1*HX
      const double* hhi = sWh + (c * T1 + b1) * HX + XO + a0;
      auto ld = [&](int k) -> double {
        if (k < NG) return hlo[k * T2 * T1 * HX];
        if (k < NG + T3) return core[(k - NG) * T2 * T1 * PC];
        return hhi[(k - NG - T3) * T2 * T1 * HX];
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
        res[d] = sub_flux(racc[d * T2 * T1 * PR], ay, uR, uL, g.dx[3], rdx3);
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
        const double dl = rk_delta(upd, res[d], di[d], upd.delta_in != nullptr);
        if (upd.delta_out) upd.delta_out[idx] = dl;
        upd.pred[idx] = rk_pred(upd, fo[d], upd.use_delta ? dl : res[d], idx);
      }
    }
  }
}

// ---- tensor maps: built on the host through the driver entry point, cached per (pointer, geometry) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    (void)cudaGetLastError();
  }
  return fn;
}
static bool encode_map(CUtensorMap* m, const DGeo& g, const double* f, int b0, int b1, int b2, int b3) {
  EncodeTiledFn enc = get_encoder();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)g.nd[0], (cuuint64_t)g.nd[1], (cuuint64_t)g.nd[2], (cuuint64_t)g.nd[3]};
  cuuint64_t strides[3] = {(cuuint64_t)g.s[1] * 8, (cuuint64_t)g.s[2] * 8, (cuuint64_t)g.s[3] * 8};
  cuuint32_t box[4] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2, (cuuint32_t)b3};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)f, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
struct MapKey {
  const double* f;
  int nd[4];
  int t[4];
  int ng;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
template <int ORDER, int T0, int T1, int T2, int T3>
static bool get_maps(const DGeo& g, const double* f, TileMaps* out) {
  using C = TileCfg<ORDER, T0, T1, T2, T3>;
  // TMA needs 16-byte aligned base and pitches
  if (((uintptr_t)f & 15) || (g.nd[0] & 1)) return false;
  static MapKey keys[16];
  static TileMaps vals[16];
  static int count = 0, next = 0;
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.f = f;
  for (int d = 0; d < 4; ++d) k.nd[d] = g.nd[d];
  k.t[0] = T0; k.t[1] = T1; k.t[2] = T2; k.t[3] = T3;
  k.ng = C::NG;
  for (int i = 0; i < count; ++i)
    if (keys[i] == k) { *out = vals[i]; return true; }
  TileMaps m;
  if (!encode_map(&m.core, g, f, T0 + C::W, T1, T2, T3)) return false;
  if (!encode_map(&m.yh, g, f, C::HX, C::NG, T2, T3)) return false;
  if (!encode_map(&m.vh, g, f, C::HX, T1, C::NG, T3)) return false;
  if (!encode_map(&m.wh, g, f, C::HX, T1, T2, C::NG)) return false;
  const int slot = (count < 16) ? count++ : (next++ % 16);
  keys[slot] = k;
  vals[slot] = m;
  *out = m;
  return true;
}

template <int ORDER, int T0, int T1, int T2, int T3, int NT, bool FULL, bool TMA>
static cudaError_t launch_one(const DGeo& g, const double* f, const double* vel, const DAccel& a, const DUpd& u,
                              double* rhs_out, int flags, long long tiles, int nt0, int nt1, int nt2,
                              const TileMaps& maps, cudaStream_t st) {
  using C = TileCfg<ORDER, T0, T1, T2, T3>;
  auto kern = k_stencil_tiled<ORDER, T0, T1, T2, T3, NT, FULL, TMA>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  kern<<<(unsigned)tiles, NT, C::SMEM_BYTES, st>>>(g, f, vel, a, u, rhs_out, flags, nt0, nt1, nt2, maps);
  return cudaGetLastError();
}

template <int ORDER, int T0, int T1, int T2, int T3, int NT>
static cudaError_t launch_tiled_cfg(const DGeo& g, const double* f, const double* vel, const DAccel& a, const DUpd& u,
                                    double* rhs_out, int flags, cudaStream_t st) {
  const int nt0 = (g.n[0] + T0 - 1) / T0, nt1 = (g.n[1] + T1 - 1) / T1, nt2 = (g.n[2] + T2 - 1) / T2,
            nt3 = (g.n[3] + T3 - 1) / T3;
  const long long tiles = (long long)nt0 * nt1 * nt2 * nt3;
  if (tiles > 0x7fffffffLL) return cudaErrorNotSupported;
  TileMaps maps;
  memset(&maps, 0, sizeof(maps));
  static int use_tma_env = -1;
  if (use_tma_env < 0) {
    const char* e = getenv("LK_NO_TMA");
    use_tma_env = (e && e[0] == '1') ? 0 : 1;
  }
  const bool tma = use_tma_env && get_maps<ORDER, T0, T1, T2, T3>(g, f, &maps);
  if (flags == 3) {
    if (tma) return launch_one<ORDER, T0, T1, T2, T3, NT, true, true>(g, f, vel, a, u, rhs_out, flags, tiles, nt0, nt1, nt2, maps, st);
    return launch_one<ORDER, T0, T1, T2, T3, NT, true, false>(g, f, vel, a, u, rhs_out, flags, tiles, nt0, nt1, nt2, maps, st);
  }
  if (tma) return launch_one<ORDER, T0, T1, T2, T3, NT, false, true>(g, f, vel, a, u, rhs_out, flags, tiles, nt0, nt1, nt2, maps, st);
  return launch_one<ORDER, T0, T1, T2, T3, NT, false, false>(g, f, vel, a, u, rhs_out, flags, tiles, nt0, nt1, nt2, maps, st);
}

static cudaError_t launch_stencil_tiled(const DGeo& g, const double* f, const double* vel, const DAccel& a,
                                        const DUpd& u, double* rhs_out, int flags, cudaStream_t st) {
  if (g.order == 4) return launch_tiled_cfg<4, 8, 8, 8, 4, 256>(g, f, vel, a, u, rhs_out, flags, st);
  return launch_tiled_cfg<6, 8, 8, 8, 4, 256>(g, f, vel, a, u, rhs_out, flags, st);
}

}  // namespace LK_NS
