// lk_march.cuh -- the fused Vlasov stage kernel (SURVEY 8a rows a3 + a4/a5 + a8 + a9 + a10/a11 + a12/a13).
//
// One CTA owns a 3D tile (T0 x T1 x T2 cells in x, y, vx) and MARCHES along vy.  Per vy plane it
//   1. sweeps the plane's lines in x, y and vx out of shared memory with a sliding register window
//      (one shared-memory read and one WENO fit per face; the left face of a cell is the right face of
//      its predecessor exactly as in the reference's `uLeft = uRight` loops, KineticSpeciesF.f:
//      1990-2005, 2137-2152), accumulating ((x + y) + vx) in the reference's order;
//   2. fits the vy face above the plane from the ring of planes it already holds (the face below is
//      kept in a register from the previous step), so f is read ONCE along vy: no vy halo, no
//      redundant vy fits;
//   3. applies the Runge-Kutta stage update (RK4Integrator.H:149-171, RK6Integrator.H:105-130)
//      straight to global memory and adds the new predictor to per-thread velocity moments (charge
//      density / currents of the NEXT stage input, ReductionSchedule.C:421-444).
// Planes travel HBM -> shared memory by TMA (cp.async.bulk.tensor.4d, out-of-box elements zero filled
// by the hardware) into a ring of 2*ng-1 slots, one plane ahead of the compute; only the plane being
// swept needs its y / vx star halos, which are single-buffered and re-armed as soon as their sweep is
// done.  rhs never touches HBM.
//
// The kernel is bounded by the fp64 pipe: ~120 fp64 instructions per order-4
// cell-update after the algebraic restatement of the fit (lk_device.cuh).
#pragma once
#include <cuda.h>

#include "lk_device.cuh"

namespace LK_NS {

template <int ORDER, int T0, int T1, int T2>
struct MarchCfg {
  static constexpr int NG = (ORDER == 4) ? 2 : 3;
  static constexpr int W = 2 * NG;          // stencil width of one fit
  static constexpr int NS = W - 1;          // ring slots (the oldest plane of a vy fit lives in registers)
  static constexpr int SX = 8;              // x-line segment walked by one thread
  static constexpr int PC = T0 + 2 * NG;    // row pitch of every staged box (TMA writes boxes densely)
  static constexpr int PA = T0 + 1;         // accumulator pitch (odd: conflict-free column access)
  static constexpr int NCORE = PC * T1 * T2;
  static constexpr int NYH = PC * NG * T2;  // one side
  static constexpr int NVH = PC * T1 * NG;  // one side
  // RK operand tiles (f_old, delta_in) staged by TMA: the box starts on a 16-byte boundary of the row
  // (x = o0 + OPX, OPX even) and is OPW wide; the tile's first cell sits OPO elements into it
  static constexpr int OPX = NG & ~1, OPO = NG & 1, OPW = T0 + 2 * OPO;
  static constexpr int NOP = OPW * T1 * T2;
  static constexpr int NACC = (PA * T1 * T2 > NOP) ? PA * T1 * T2 : NOP;  // accumulator, then f_old tile
  // sAcc doubles as the landing zone of the f_old tile (TMA, once the accumulator has been read into
  // registers); NACC more doubles hold the delta_in tile
  static constexpr int SMEM_DOUBLES = NS * NCORE + 2 * NYH + 2 * NVH + 2 * NACC;
  // + the (vx, vy) cell-centre velocities of the tile's T2 slices for the current and the next vy plane
  static constexpr size_t SMEM_BYTES = sizeof(double) * SMEM_DOUBLES + 8 * (NS + 4) + sizeof(double) * 4 * T2;
};

struct MarchMaps {
  CUtensorMap core, yh, vh;  // boxes of the array being differentiated
  CUtensorMap fo, di;        // dense tile boxes of the RK operands f_old and delta_in
};
// velocity moments of the predictor written by the epilogue: part[(m * nparts + p) * n0 * n1 + x + n0 * y],
// p = chunk * nt2 + (vx tile); m = 0: sum f, 1: sum vx f, 2: sum vy f
struct DMom {
  double* part;
  int nmom, nparts;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, void* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void mbar_expect(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait returns within a few cycles when the phase is not complete: a bare retry loop burns the issue slots the
// other warps of the SM sub-partition need (measured with ncu: 15 % of all executed instructions were polls), so
// the retry path sleeps 64 ns between polls
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LK_MDONE;\n\t"
      "LK_MRETRY:\n\t"
      "nanosleep.u32 64;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra LK_MRETRY;\n\t"
      "LK_MDONE:\n\t"
      "}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---- faces along a line: init with the first W-1 values, then one value in, one face out ----
// Faces carry the factor 1/FaceScale (production order 4 keeps 12*face and folds 1/12 into the flux
// coefficient).
template <int V>
struct IntTag {
  static constexpr int value = V;
};
template <int ORDER>
struct FaceScale {
#if LK_STRICT
  static constexpr double v = 1.0;
#else
  static constexpr double v = (ORDER == 4) ? (1.0 / 12.0) : (1.0 / 60.0);
#endif
};
template <int ORDER>
struct Walker {
  static constexpr int W = (ORDER == 4) ? 4 : 6;
  double w[W - 1];
  template <class LD>
  __device__ __forceinline__ void init(LD ld) {
#pragma unroll
    for (int k = 0; k < W - 1; ++k) w[k] = ld(k);
  }
  __device__ __forceinline__ double next(double un, bool pos) {
    double F;
    if constexpr (ORDER == 4) F = weno43(w[0], w[1], w[2], un, pos);
    else F = weno65(w[0], w[1], w[2], w[3], w[4], un, pos);
#pragma unroll
    for (int k = 0; k < W - 2; ++k) w[k] = w[k + 1];
    w[W - 2] = un;
    return F;
  }
};
#if !LK_STRICT
template <>
struct Walker<4> {
  double um, u0, up, d, c, pc;  // u_{j-1}, u_j, u_{j+1}, d_j, c_j, pc_j
  template <class LD>
  __device__ __forceinline__ void init(LD ld) {
    um = ld(0); u0 = ld(1); up = ld(2);
    const double dm = ADD(u0, -um);
    d = ADD(up, -u0);
    c = ADD(d, -dm);
    pc = w43_pc(c);
  }
  __device__ __forceinline__ double next(double un, bool pos) {
    const double dn = ADD(un, -up), cn = ADD(dn, -d), pcn = w43_pc(cn);
    const double F = w43_face12(um, u0, up, un, d, c, pc, cn, pcn, pos);
    um = u0; u0 = up; up = un; d = dn; c = cn; pc = pcn;
    return F;
  }
};
template <>
struct Walker<6> {
  double u0, u1, u2, u3, u4, E0, E1, E2, E3;  // um3..up1 and their forward differences
  template <class LD>
  __device__ __forceinline__ void init(LD ld) {
    u0 = ld(0); u1 = ld(1); u2 = ld(2); u3 = ld(3); u4 = ld(4);
    E0 = ADD(u1, -u0); E1 = ADD(u2, -u1); E2 = ADD(u3, -u2); E3 = ADD(u4, -u3);
  }
  __device__ __forceinline__ double next(double un, bool pos) {
    const double E4 = ADD(un, -u4);
    const double F = w65_face60(u0, u1, u2, u3, u4, un, E0, E1, E2, E3, E4, pos);
    u0 = u1; u1 = u2; u2 = u3; u3 = u4; u4 = un;
    E0 = E1; E1 = E2; E2 = E3; E3 = E4;
    return F;
  }
};
#endif
// the same face from its W values at once (vy direction, one-thread-per-cell kernel); identical bits
template <int ORDER>
__device__ __forceinline__ double fit_face(const double* w, bool pos) {
#if !LK_STRICT
  if constexpr (ORDER == 4) return weno43_face12(w[0], w[1], w[2], w[3], pos);
  else return weno65_face60(w[0], w[1], w[2], w[3], w[4], w[5], pos);
#endif
  if constexpr (ORDER == 4) return weno43(w[0], w[1], w[2], w[3], pos);
  else return weno65(w[0], w[1], w[2], w[3], w[4], w[5], pos);
}

// LEAN: the production instantiation -- advection + acceleration, no accumulate, acceleration constant
// along its own sweep line (Vlasov-Poisson without a constant B field: vel3 is independent of i3 and
// vel4 of i4, KineticSpeciesF.f:78-80, 98-100), no earlier RK6 stage results to add; everything else is
// decided at run time.
template <int ORDER, int T0, int T1, int T2, int NT, bool TMA, bool LEAN>
__global__ void __launch_bounds__(NT, (MarchCfg<ORDER, T0, T1, T2>::SMEM_BYTES <= 113 * 1024) ? 2 : 1)
k_stage_march(const DGeo g, const double* __restrict__ f, const double* __restrict__ vel, const DAccel a,
              const DUpd upd, double* __restrict__ rhs_out, const int flags, const int nt0, const int nt1,
              const int nt2, const int chunk_len, const DMom mom, const __grid_constant__ MarchMaps maps) {
  using C = MarchCfg<ORDER, T0, T1, T2>;
  constexpr int NG = C::NG, W = C::W, NS = C::NS, SX = C::SX, PC = C::PC, PA = C::PA;
  static_assert(T0 * T1 == NT, "the vy/epilogue phase maps one thread to one (x,y) column of the tile");
  static_assert(T0 % SX == 0 && (T0 % 2) == 0, "x segments");
#ifdef LK_EXP_NOSLICE
  constexpr bool WARP_SLICE = false;
#else
  constexpr bool WARP_SLICE = (T0 == 32) && (T1 * (T0 / SX) == 32) && (NT == 32 * T2);
#endif
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  double* sCore = smem;                       // NS slots
  double* sYh = sCore + NS * C::NCORE;        // [side][c][h][PC]
  double* sVh = sYh + 2 * C::NYH;             // [side][h][b1][PC]
  double* sAcc = sVh + 2 * C::NVH;            // [c][b1][PA]
  double* sDi = sAcc + C::NACC;               // [c][b1][PA]  delta_in of the plane being updated
  unsigned long long* bars = (unsigned long long*)(sDi + C::NACC);  // NS core barriers, y, v, f_old tile, delta_in tile
  double* sVel = (double*)(bars + NS + 4);                          // [plane parity][vx | vy][T2]

  const int tid = threadIdx.x;
  int b = blockIdx.x;
  // x tiles fastest, then y, then vx: the CTAs resident together cover whole (x,y) planes of a few vx
  // slices (measured 1 % faster than vx-before-y)
#if defined(LK_GY) && defined(LK_GV)
  // supertile order: the CTAs resident together cover a block of LK_GY x LK_GV (y, vx) tiles over all x, so
  // both the y and the vx star halos of a tile are mostly the core boxes of co-resident neighbours (L2 hits)
  const int o0 = (b % nt0) * T0; b /= nt0;
  int o1, o2;
  {
    const int nyv = nt1 * nt2;
    int r = b % nyv;
    b /= nyv;
    const int gv = r / (nt1 * LK_GV);
    r -= gv * nt1 * LK_GV;
    const int hv = min(LK_GV, nt2 - gv * LK_GV);
    const int gy = r / (LK_GY * hv);
    r -= gy * LK_GY * hv;
    const int hy = min(LK_GY, nt1 - gy * LK_GY);
    o1 = (gy * LK_GY + r % hy) * T1;
    o2 = (gv * LK_GV + r / hy) * T2;
  }
#elif defined(LK_EXP_ORDER_V)
  const int o0 = (b % nt0) * T0; b /= nt0;
  const int o2 = (b % nt2) * T2; b /= nt2;
  const int o1 = (b % nt1) * T1; b /= nt1;
#else
  const int o0 = (b % nt0) * T0; b /= nt0;
  const int o1 = (b % nt1) * T1; b /= nt1;
  const int o2 = (b % nt2) * T2; b /= nt2;
#endif
  const int chunk = b;
  // g.ng == NG by construction.  Order 4 (two CTAs per SM, 128 registers): a constant, and the plane stride
  // below pinned in registers -- the compiler otherwise re-reads both from the constant bank every plane and
  // the epilogue's address arithmetic waits on those loads (A/B on one box: +1.3 %).  The 255-register
  // order-6 instantiation is better off without (-1.2 %).
  const int ng = (ORDER == 4) ? NG : g.ng;
  const int q0 = chunk * chunk_len;                       // first interior vy plane of this CTA
  const int nq = min(chunk_len, g.n[3] - q0);
  if (nq <= 0) return;
  const int X0 = o0, Y0 = o1 + ng, V0 = o2 + ng;          // data-box coordinates of the staged boxes' origin (x grown by NG)
  const int pbase = q0 + ng - NG + 1;                     // data index of the plane in ring slot 0 at start

  const bool do_adv = LEAN || (flags & 1) != 0, do_acc = LEAN || (flags & 2) != 0, accumulate = !LEAN && (flags & 4) != 0;
  const double FS = FaceScale<ORDER>::v;
  const double rdx0 = (1.0 / g.dx[0]) * FS, rdx1 = (1.0 / g.dx[1]) * FS, rdx2 = (1.0 / g.dx[2]) * FS,
               rdx3 = (1.0 / g.dx[3]) * FS;

  // ---- staging helpers -------------------------------------------------------------------------
  auto coop_box = [&](double* dst, int x0, int y0, int v0, int p, int by, int bv) {  // fallback: plain loads
    for (int e = tid; e < PC * by * bv; e += NT) {
      const int k = e % PC, r = e / PC, j = r % by, m = r / by;
      const int x = x0 + k, y = y0 + j, v = v0 + m;
      const bool in = x >= 0 && x < g.nd[0] && y >= 0 && y < g.nd[1] && v >= 0 && v < g.nd[2] && p >= 0 && p < g.nd[3];
      dst[e] = in ? f[gidx(g, x, y, v, p)] : 0.0;
    }
  };
  auto stage_core = [&](int p, int slot) {
    double* dst = sCore + slot * C::NCORE;
    if (TMA) {
      if (tid == 0) {
        mbar_expect(&bars[slot], (unsigned)(C::NCORE * sizeof(double)));
        tma_load_4d(dst, &maps.core, &bars[slot], X0, Y0, V0, p);
      }
    } else {
      coop_box(dst, X0, Y0, V0, p, T1, T2);
    }
  };
  auto stage_yh = [&](int p) {
    if (TMA) {
      if (tid == 0) {
        mbar_expect(&bars[NS], (unsigned)(2 * C::NYH * sizeof(double)));
        tma_load_4d(sYh, &maps.yh, &bars[NS], X0, Y0 - NG, V0, p);
        tma_load_4d(sYh + C::NYH, &maps.yh, &bars[NS], X0, Y0 + T1, V0, p);
      }
    } else {
      coop_box(sYh, X0, Y0 - NG, V0, p, NG, T2);
      coop_box(sYh + C::NYH, X0, Y0 + T1, V0, p, NG, T2);
    }
  };
  auto stage_vh = [&](int p) {
    if (TMA) {
      if (tid == 0) {
        mbar_expect(&bars[NS + 1], (unsigned)(2 * C::NVH * sizeof(double)));
        tma_load_4d(sVh, &maps.vh, &bars[NS + 1], X0, Y0, V0 - NG, p);
        tma_load_4d(sVh + C::NVH, &maps.vh, &bars[NS + 1], X0, Y0, V0 + T2, p);
      }
    } else {
      coop_box(sVh, X0, Y0, V0 - NG, p, T1, NG);
      coop_box(sVh + C::NVH, X0, Y0, V0 + T2, p, T1, NG);
    }
  };
  unsigned ph_core = 0, ph_y = 0, ph_v = 0, ph_o = 0;  // phase parities (bit per core slot)
  auto wait_core = [&](int slot) {
    if (TMA) {
      mbar_wait(&bars[slot], (ph_core >> slot) & 1u);
      ph_core ^= 1u << slot;
    }
  };

  if (TMA) {
    if (tid == 0) {
      for (int k = 0; k < NS + 4; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[k])));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }

  // ---- this thread's (x,y) column for the vx sweep, the vy fit and the epilogue --------------------
  const int ea0 = tid % T0, eb1 = tid / T0;
  const bool col_ok = (o0 + ea0 < g.n[0]) && (o1 + eb1 < g.n[1]);
  const int ei1 = min(o0 + ea0, g.n[0] - 1) + ng, ei2 = min(o1 + eb1, g.n[1] - 1) + ng;
  const i64 col = (i64)(o0 + ea0 + ng) + g.s[1] * (o1 + eb1 + ng) + g.s[2] * (o2 + ng);  // + s2*c + s3*p
  const int ecell = eb1 * PC + NG + ea0;                                                  // + c*T1*PC within a slot
  const int ncv = col_ok ? min(T2, g.n[2] - o2) : 0;  // valid cells of this thread's column
  const bool simple_acc = LEAN || ((a.kind == 0) && (a.bz == 0.0));
  const i64 pxy = ei1 + (i64)g.nd[0] * ei2;

  // ---- prologue: fill the ring, fit the face below the first plane ------------------------------
  for (int k = 0; k < NS; ++k) stage_core(pbase + k, k);
  stage_yh(q0 + ng);
  stage_vh(q0 + ng);
  double uold[T2], Fprev[T2];
#pragma unroll
  for (int c = 0; c < T2; ++c) {
    const bool ok = col_ok && (o2 + c < g.n[2]);
    uold[c] = ok ? f[col + g.s[2] * c + g.s[3] * (pbase - 1)] : 0.0;
  }
  // velocities of the slices at plane p: one warp fetches them a plane ahead, every sweep reads shared memory
  auto stage_vel = [&](int p) {
    if (tid < 2 * T2) {
      const int c = tid % T2, comp = tid / T2;
      const int i3 = min(o2 + c, g.n[2] - 1) + ng;
      sVel[(p & 1) * 2 * T2 + tid] = __ldg(vel + i3 + (i64)g.nd[2] * (p + (i64)comp * g.nd[3]));
    }
  };
  if (vel) stage_vel(q0 + ng);
  if (!TMA) __syncthreads();
  for (int k = 0; k < NS; ++k) wait_core(k);
  if (do_acc) {
    const int i4b = (q0 > 0) ? (q0 + ng - 1) : (q0 + ng);  // the cell whose coefficient fitted this face (KineticSpeciesF.f:2137-2141)
#pragma unroll
    for (int c = 0; c < T2; ++c) {
      double w[W];
      w[0] = uold[c];
#pragma unroll
      for (int k = 0; k < NS; ++k) w[k + 1] = sCore[k * C::NCORE + ecell + c * T1 * PC];
      const int i3 = min(o2 + c, g.n[2] - 1) + ng;
      const double ayb = simple_acc ? __ldg(a.field + pxy + (i64)g.nd[0] * g.nd[1]) : accel_y(a, g, ei1, ei2, i3, i4b);
      Fprev[c] = fit_face<ORDER>(w, ayb > 0.0);
    }
  }
#pragma unroll
  for (int c = 0; c < T2; ++c) uold[c] = sCore[ecell + c * T1 * PC];  // plane pbase
  __syncthreads();
  stage_core(pbase + NS, 0);

  double m0 = 0.0, m1 = 0.0, m2 = 0.0;
  int ekind = 0;
  if (upd.active && !rhs_out && upd.n_prev == 0 && do_adv && do_acc && !accumulate) {
    if (!upd.delta_in && upd.delta_out && !upd.use_delta) ekind = 1;
    else if (upd.delta_in && upd.delta_out && !upd.use_delta) ekind = 2;
    else if (upd.delta_in && !upd.delta_out && upd.use_delta) ekind = 3;
  }
  if (upd.krook_nu) ekind = 0;  // the Krook term goes between the rhs and the update: the per-cell epilogue
  // E (times q/m) at this thread's (x,y): constant along the whole march when vel3/vel4 do not depend
  // on the velocity indices
  const double ax0 = (do_acc && simple_acc) ? __ldg(a.field + pxy) : 0.0;
  const double ay0 = (do_acc && simple_acc) ? __ldg(a.field + pxy + (i64)g.nd[0] * g.nd[1]) : 0.0;
  double* const racc = sAcc + eb1 * PA + ea0;  // + c*T1*PA: this thread's column of the accumulator
  // periodic ghost copies of the predictor written by the cell that owns the value (upd.wrap): element
  // offsets of this column's images, 0 = none.  The corner image keeps the whole data box consistent.
  int gxo = 0, gyo = 0;
  if (upd.active && col_ok) {
    const int x = o0 + ea0, y = o1 + eb1;
    if ((upd.wrap & 1) && g.n[0] >= 2 * ng) gxo = (x < ng) ? g.n[0] : ((x >= g.n[0] - ng) ? -g.n[0] : 0);
    if ((upd.wrap & 2) && g.n[1] >= 2 * ng) gyo = (y < ng) ? g.n[1] * (int)g.s[1] : ((y >= g.n[1] - ng) ? -g.n[1] * (int)g.s[1] : 0);
  }

  // ---- march -------------------------------------------------------------------------------------
  // EK: epilogue kind, decided once per kernel (uniform) so that the per-cell code carries no pointer
  // tests.  1..3 = the RK4 stage shapes (RK4Integrator.H:149-171): 1: delta = w*rhs (stage 1);
  // 2: delta += w*rhs (stages 2,3); 3: pred = f_old + c*(delta + w*rhs), delta not stored (stage 4);
  // 0: everything decided per cell (RK6, rhs_out, partial evaluations).
  // Every sweep loads its whole line into registers BEFORE the first fit and stores after the last: no
  // shared-memory store sits between the loads, so the W+T fits of a line are independent instruction
  // streams the scheduler can interleave (the fp64 dependent-issue latency is what bounds this kernel).
  // RK operand tiles land dense: [c][b1][T0]
  const double* const ofo = sAcc + eb1 * C::OPW + C::OPO + ea0;
  const double* const odi = sDi + eb1 * C::OPW + C::OPO + ea0;
  i64 s3 = g.s[3];
  if constexpr (ORDER == 4) asm volatile("" : "+l"(s3));  // opaque to the compiler: stays in registers
  auto march = [&](auto ek_tag, auto full_tag) {
    constexpr int EK = decltype(ek_tag)::value;
    constexpr bool FULL = decltype(full_tag)::value != 0;  // the tile lies inside the interior in x, y and vx
    for (int q = q0; q < q0 + nq; ++q) {
      const int p = q + ng;                       // data index of the plane being updated
      const int sc = (p - pbase) % NS;            // its ring slot
      const double* cur = sCore + sc * C::NCORE;
      const bool more = (q + 1 < q0 + nq);
      const double* const sv = sVel + (p & 1) * 2 * T2;  // [vx(c) | vy(c)] of this plane
#ifndef LK_EXP_LATEVEL
      // the next plane's velocities are fetched now and stored at the end of the plane: the load latency of
      // the sixteen fetching threads must not delay the barrier that closes the plane
      double vel_next = 0.0;
      if (more && vel && tid < 2 * T2) {
        const int c = tid % T2, comp = tid / T2;
        const int i3 = min(o2 + c, g.n[2] - 1) + ng;
        vel_next = __ldg(vel + i3 + (i64)g.nd[2] * ((p + 1) + (i64)comp * g.nd[3]));
      }
#endif
      if constexpr (EK >= 2 && TMA) {
        // the delta_in tile of this plane has its own buffer, free since the barrier that closed the previous
        // plane: fetch it now, a whole plane ahead of the epilogue (only the f_old tile has to wait for the
        // accumulator to be drained)
        if (tid == 0) {
          mbar_expect(&bars[NS + 3], (unsigned)(C::NOP * sizeof(double)));
          tma_load_4d(sDi, &maps.di, &bars[NS + 3], o0 + C::OPX, o1 + ng, o2 + ng, p);
        }
      }

#ifdef LK_EXP_L2PF  // measured: the explicit L2 prefetch of the next plane's RK operands costs 2 % (A/B on one box)
      if (upd.active && more) {
        constexpr int LPR = (T0 * 8 + 127) / 128;  // lines per row
        for (int e = tid; e < T1 * T2 * LPR; e += NT) {
          const int ln = e % LPR, row = e / LPR, b1 = row % T1, c = row / T1;
          if ((o1 + b1 < g.n[1]) && (o2 + c < g.n[2]) && (o0 + ln * 16 < g.n[0])) {
            const i64 o = (i64)(o0 + ng + ln * 16) + g.s[1] * (o1 + b1 + ng) + g.s[2] * (o2 + c + ng) + g.s[3] * (p + 1);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(upd.f_old + o));
            if (upd.delta_in) asm volatile("prefetch.global.L2 [%0];" ::"l"(upd.delta_in + o));
          }
        }
      }
#endif

      // ---------------- x sweep: rows (b1, c) in segments of SX cells ----------------
      {
        constexpr int ITEMS = T1 * T2 * (T0 / SX);
#pragma unroll
        for (int it = 0; it < (ITEMS + NT - 1) / NT; ++it) {
          const int l = tid + it * NT;
          if ((ITEMS % NT) != 0 && l >= ITEMS) break;
          // WARP_SLICE: warp w owns the vx slice c = w in both the x and the y sweep (its 32 lanes are the
          // slice's T1 rows x T0/SX segments here, its T0 columns there), so the two sweeps of a slice are
          // ordered by __syncwarp() and the CTA needs no barrier between them
          const int row = WARP_SLICE ? ((l >> 5) * T1 + (l & 31) % T1) : l % (T1 * T2);
          const int seg = WARP_SLICE ? ((l & 31) / T1) : l / (T1 * T2);
          const int b1 = row % T1, c = row / T1;
          double* arow = sAcc + row * PA + seg * SX;
          double init[SX];
#pragma unroll
          for (int k = 0; k < SX; ++k) init[k] = 0.0;
          if (accumulate) {
#pragma unroll
            for (int k = 0; k < SX; ++k) {
              const bool ok = (o0 + seg * SX + k < g.n[0]) && (o1 + b1 < g.n[1]) && (o2 + c < g.n[2]);
              if (ok) init[k] = rhs_out[(i64)(o0 + seg * SX + k + ng) + g.s[1] * (o1 + b1 + ng) + g.s[2] * (o2 + c + ng) + g.s[3] * p];
            }
          }
          if (do_adv) {
            const double vx = sv[c];
            const bool pos = vx > 0.0;
            const double2* r2 = reinterpret_cast<const double2*>(cur + row * PC + seg * SX);
            double v[SX + W];
#pragma unroll
            for (int k = 0; k < (SX + W) / 2; ++k) {
              const double2 t = r2[k];
              v[2 * k] = t.x;
              v[2 * k + 1] = t.y;
            }
            Walker<ORDER> wk;
            wk.init([&](int k) { return v[k]; });
            double uL = wk.next(v[W - 1], pos);
#pragma unroll
            for (int k = 0; k < SX; ++k) {
              const double uR = wk.next(v[k + W], pos);
              init[k] = sub_flux(init[k], vx, uR, uL, g.dx[0], rdx0);
              uL = uR;
            }
          }
#pragma unroll
          for (int k = 0; k < SX; ++k) arow[k] = init[k];
        }
      }
      if constexpr (WARP_SLICE) __syncwarp();
      else __syncthreads();

      // ---------------- y sweep: lines (a0, c) ----------------
      if (TMA) { mbar_wait(&bars[NS], ph_y); ph_y ^= 1u; }
      if (do_adv) {
        constexpr int ITEMS = T0 * T2;
#pragma unroll
        for (int it = 0; it < (ITEMS + NT - 1) / NT; ++it) {
          const int l = tid + it * NT;
          if ((ITEMS % NT) != 0 && l >= ITEMS) break;
          const int a0 = l % T0, c = l / T0;
          const double vy = sv[T2 + c];
          const bool pos = vy > 0.0;
          const double* core = cur + c * T1 * PC + NG + a0;        // + b1*PC
          const double* hlo = sYh + c * NG * PC + NG + a0;         // + h*PC
          const double* hhi = hlo + C::NYH;
          double* yacc = sAcc + c * T1 * PA + a0;                  // + b1*PA
          double v[T1 + W], acc[T1];
#pragma unroll
          for (int k = 0; k < NG; ++k) v[k] = hlo[k * PC];
#pragma unroll
          for (int k = 0; k < T1; ++k) v[NG + k] = core[k * PC];
#pragma unroll
          for (int k = 0; k < NG; ++k) v[NG + T1 + k] = hhi[k * PC];
#pragma unroll
          for (int b1 = 0; b1 < T1; ++b1) acc[b1] = yacc[b1 * PA];
          Walker<ORDER> wk;
          wk.init([&](int k) { return v[k]; });
          double uL = wk.next(v[W - 1], pos);
#pragma unroll
          for (int b1 = 0; b1 < T1; ++b1) {
            const double uR = wk.next(v[b1 + W], pos);
            acc[b1] = sub_flux(acc[b1], vy, uR, uL, g.dx[1], rdx1);
            uL = uR;
          }
#pragma unroll
          for (int b1 = 0; b1 < T1; ++b1) yacc[b1 * PA] = acc[b1];
        }
      }
      __syncthreads();
      if (more) stage_yh(p + 1);

      // ---------------- vx sweep: this thread's column (ea0, eb1), all c, kept in registers ----------
      if (TMA) { mbar_wait(&bars[NS + 1], ph_v); ph_v ^= 1u; }
      double res[T2];
#pragma unroll
      for (int c = 0; c < T2; ++c) res[c] = racc[c * T1 * PA];
      const i64 idx0 = col + s3 * p;
      const int s2 = (int)g.s[2];
      if constexpr (EK != 0 && TMA) {
        // the RK operands of the tile travel HBM -> shared memory by TMA while the vx and vy fits run:
        // f_old into the accumulator (every thread has just moved its column of it into registers),
        // delta_in next to it.  Out-of-range cells are zero filled and never stored.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes of sAcc before the async ones
        // split barrier: every warp announces that its columns of the accumulator are in registers and goes
        // on to the vx sweep; only warp 0 waits for the announcements before it lets the TMA overwrite sAcc
        // (measured on one box: +1 % with one resident CTA per SM -- order 6 --, -1.3 % with two, where the
        // second CTA already fills the wait and warp 0 only becomes a straggler)
        constexpr bool SPLIT_BARRIER = C::SMEM_BYTES > 113 * 1024;
        if (SPLIT_BARRIER && tid >= 32) {
          asm volatile("bar.arrive 1, %0;" ::"r"(NT) : "memory");
        } else {
          if constexpr (SPLIT_BARRIER) asm volatile("bar.sync 1, %0;" ::"r"(NT) : "memory");
          else __syncthreads();
          if (tid == 0) {
            mbar_expect(&bars[NS + 2], (unsigned)(C::NOP * sizeof(double)));
            tma_load_4d(sAcc, &maps.fo, &bars[NS + 2], o0 + C::OPX, o1 + ng, o2 + ng, p);
          }
        }
      }
      if (do_acc) {
        const double* core = cur + ecell;                          // + c*T1*PC
        const double* hlo = sVh + eb1 * PC + NG + ea0;             // + h*T1*PC
        const double* hhi = hlo + C::NVH;
        double v[T2 + W];
#pragma unroll
        for (int k = 0; k < NG; ++k) v[k] = hlo[k * T1 * PC];
#pragma unroll
        for (int k = 0; k < T2; ++k) v[NG + k] = core[k * T1 * PC];
#pragma unroll
        for (int k = 0; k < NG; ++k) v[NG + T2 + k] = hhi[k * T1 * PC];
        auto AX = [&](int i3) -> double { return simple_acc ? ax0 : accel_x(a, g, ei1, ei2, i3, p); };
        const int i3first = o2 + ng;
        // the face below the first cell was fitted by the cell below it with ITS coefficient, unless
        // that cell is outside the interior (KineticSpeciesF.f:2137-2141)
        const double axl = AX((o2 > 0) ? (i3first - 1) : i3first);
        Walker<ORDER> wk;
        wk.init([&](int k) { return v[k]; });
        double uL = wk.next(v[W - 1], axl > 0.0);
#pragma unroll
        for (int c = 0; c < T2; ++c) {
          const double ax = AX(min(i3first + c, g.n[2] - 1 + ng));
          const double uR = wk.next(v[c + W], ax > 0.0);
          res[c] = sub_flux(res[c], ax, uR, uL, g.dx[2], rdx2);
          uL = uR;
        }
      }

      // ---------------- vy face above this plane + epilogue ----------------
      const int s_new = (p + NG - pbase) % NS;
      wait_core(s_new);  // always: no TMA write may be outstanding when the CTA exits
      {
        double* do_p = upd.delta_out + idx0;
        double* pr_p = upd.pred + idx0;
        if (do_acc) {
          // ring slots of the planes p-NG+2 .. p+NG (w[1..W-1] of the vy fit), this thread's cell
          const double* wp[W];
#pragma unroll
          for (int k = 1; k < W; ++k) wp[k] = sCore + ((sc + k - (NG - 1) + NS) % NS) * C::NCORE + ecell;
#pragma unroll
          for (int c = 0; c < T2; ++c) {
            double w[W];
            w[0] = uold[c];
#pragma unroll
            for (int k = 1; k < W; ++k) w[k] = wp[k][c * T1 * PC];
            const double ay = simple_acc ? ay0 : accel_y(a, g, ei1, ei2, min(o2 + c, g.n[2] - 1) + ng, p);
            const double F = fit_face<ORDER>(w, ay > 0.0);
            res[c] = sub_flux(res[c], ay, F, Fprev[c], g.dx[3], rdx3);
            Fprev[c] = F;
            uold[c] = w[1];
          }
        }
        double psum = 0.0, pvx = 0.0, pvy = 0.0;
        if constexpr (EK != 0) {
          // branch-free: every cell computes, only the stores and the moment terms are predicated
          double fo[T2], di[T2];
          if constexpr (TMA) {
            mbar_wait(&bars[NS + 2], ph_o);
            if (EK >= 2) mbar_wait(&bars[NS + 3], ph_o);
            ph_o ^= 1u;
#pragma unroll
            for (int c = 0; c < T2; ++c) {
              fo[c] = ofo[c * T1 * C::OPW];
              di[c] = (EK >= 2) ? odi[c * T1 * C::OPW] : 0.0;
            }
          } else {
            const double* fo_p = upd.f_old + idx0;
            const double* di_p = upd.delta_in + idx0;
#pragma unroll
            for (int c = 0; c < T2; ++c) {
              const int oc = (FULL || c < ncv) ? c * s2 : 0;
              fo[c] = (FULL || ncv > 0) ? fo_p[oc] : 0.0;
              di[c] = (EK >= 2 && (FULL || ncv > 0)) ? di_p[oc] : 0.0;
            }
          }
#pragma unroll
          for (int c = 0; c < T2; ++c) {
            const bool live = FULL || c < ncv;
            const int oc = c * s2;
            double pr;
            if constexpr (EK == 1) {
              const double dl = rk_delta(upd, res[c], 0.0, false);
              if (live) do_p[oc] = dl;
              pr = rk_axpy(fo[c], upd.c_pred, res[c]);
            } else if constexpr (EK == 2) {
              const double dl = rk_delta(upd, res[c], di[c], true);
              if (live) do_p[oc] = dl;
              pr = rk_axpy(fo[c], upd.c_pred, res[c]);
            } else {
              pr = rk_axpy(fo[c], upd.c_pred, rk_delta(upd, res[c], di[c], true));
            }
            if (live) {
              pr_p[oc] = pr;
              if (gxo) pr_p[oc + gxo] = pr;
              if (gyo) {
                pr_p[oc + gyo] = pr;
                if (gxo) pr_p[oc + gyo + gxo] = pr;
              }
            }
            if (mom.nmom > 0) {
              const double prm = live ? pr : 0.0;
              const int cc = FULL ? c : min(c, max(ncv - 1, 0));
              psum = ADD(psum, prm);
              if (mom.nmom > 1) {
                pvx = FMA(sv[cc], prm, pvx);
                pvy = FMA(sv[T2 + cc], prm, pvy);
              }
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < T2; ++c) {
            if (c < ncv) {
              const i64 idx = idx0 + c * s2;
              if (upd.krook_nu && do_acc) res[c] = krook_term(upd, g, res[c], f[idx], o0 + ea0 + ng, o1 + eb1 + ng, o2 + c + ng, p);
              if (rhs_out) rhs_out[idx] = res[c];
              if (!upd.active) continue;
              const double dl = rk_delta(upd, res[c], upd.delta_in ? upd.delta_in[idx] : 0.0, upd.delta_in != nullptr);
              if (upd.delta_out) upd.delta_out[idx] = dl;
              const double pr = rk_pred(upd, upd.f_old[idx], upd.use_delta ? dl : res[c], idx);
              pr_p[c * s2] = pr;
              if (gxo) pr_p[c * s2 + gxo] = pr;
              if (gyo) {
                pr_p[c * s2 + gyo] = pr;
                if (gxo) pr_p[c * s2 + gyo + gxo] = pr;
              }
              if (mom.nmom > 0) {
                psum = ADD(psum, pr);
                if (mom.nmom > 1) {
                  pvx = FMA(sv[c], pr, pvx);
                  pvy = FMA(sv[T2 + c], pr, pvy);
                }
              }
            }
          }
        }
        if (mom.nmom > 0) {
          m0 = ADD(m0, psum);
          if (mom.nmom > 1) {
            m1 = ADD(m1, pvx);
            m2 = ADD(m2, pvy);
          }
        }
      }
#ifdef LK_EXP_LATEVEL
      if (more && vel) stage_vel(p + 1);
#else
      if (more && vel && tid < 2 * T2) sVel[((p + 1) & 1) * 2 * T2 + tid] = vel_next;
#endif
      __syncthreads();
      if (more) {
        stage_vh(p + 1);
        stage_core(p + NG + 1, (p + NG + 1 - pbase) % NS);
      }
    }
  };
  const bool full = (o0 + T0 <= g.n[0]) && (o1 + T1 <= g.n[1]) && (o2 + T2 <= g.n[2]);
  if (ekind == 0) march(IntTag<0>{}, IntTag<0>{});
  else if (full) {
    if (ekind == 1) march(IntTag<1>{}, IntTag<1>{});
    else if (ekind == 2) march(IntTag<2>{}, IntTag<1>{});
    else march(IntTag<3>{}, IntTag<1>{});
  } else {
    if (ekind == 1) march(IntTag<1>{}, IntTag<0>{});
    else if (ekind == 2) march(IntTag<2>{}, IntTag<0>{});
    else march(IntTag<3>{}, IntTag<0>{});
  }

  if (mom.nmom > 0 && col_ok) {
    const i64 nxy = (i64)g.n[0] * g.n[1];
    const i64 part = (i64)chunk * nt2 + (o2 / T2);
    const i64 o = (o0 + ea0) + (i64)g.n[0] * (o1 + eb1);
    mom.part[part * nxy + o] = m0;
    if (mom.nmom > 1) {
      mom.part[((i64)mom.nparts + part) * nxy + o] = m1;
      mom.part[((i64)2 * mom.nparts + part) * nxy + o] = m2;
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
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
struct MapKey {
  const double* f;
  int nd[4];
  int box[4];
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
static bool get_map(const DGeo& g, const double* f, int b0, int b1, int b2, CUtensorMap* out) {
  // TMA needs a 16-byte aligned base and pitches
  if (f == nullptr || ((uintptr_t)f & 15) || (g.nd[0] & 1)) return false;
  constexpr int CAP = 128;
  static MapKey keys[CAP];
  static CUtensorMap vals[CAP];
  static int count = 0, next = 0;
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.f = f;
  for (int d = 0; d < 4; ++d) k.nd[d] = g.nd[d];
  k.box[0] = b0; k.box[1] = b1; k.box[2] = b2; k.box[3] = 1;
  for (int i = 0; i < count; ++i)
    if (keys[i] == k) { *out = vals[i]; return true; }
  CUtensorMap m;
  if (!encode_map(&m, g, f, b0, b1, b2, 1)) return false;
  const int slot = (count < CAP) ? count++ : (next++ % CAP);
  keys[slot] = k;
  vals[slot] = m;
  *out = m;
  return true;
}
template <int ORDER, int T0, int T1, int T2>
static bool get_maps(const DGeo& g, const double* f, const DUpd& u, MarchMaps* out) {
  using C = MarchCfg<ORDER, T0, T1, T2>;
  if (!get_map(g, f, C::PC, T1, T2, &out->core)) return false;
  if (!get_map(g, f, C::PC, C::NG, T2, &out->yh)) return false;
  if (!get_map(g, f, C::PC, T1, C::NG, &out->vh)) return false;
  // the RK operand tiles (only read by the RK4-shaped epilogues; a missing operand keeps a valid dummy map)
  out->fo = out->core;
  out->di = out->core;
  if (u.active && u.f_old && !get_map(g, u.f_old, C::OPW, T1, T2, &out->fo)) return false;
  if (u.active && u.delta_in && !get_map(g, u.delta_in, C::OPW, T1, T2, &out->di)) return false;
  return true;
}

// number of moment partials per (x,y) the march kernel writes for this geometry, and its decomposition
template <int T2>
static void march_plan(const DGeo& g, int tiles_xyv, int* nchunk, int* chunk_len) {
  // enough CTAs for a few waves over 148 SMs x 2 resident CTAs; chunks no shorter than 8 planes
  int nc = 1;
  const int want = 148 * 2 * 2;
  if (tiles_xyv < want) nc = (want + tiles_xyv - 1) / tiles_xyv;
  int len = (g.n[3] + nc - 1) / nc;
  if (len < 8) len = (g.n[3] < 8) ? g.n[3] : 8;
  nc = (g.n[3] + len - 1) / len;
  *nchunk = nc;
  *chunk_len = len;
}

template <int ORDER, int T0, int T1, int T2, int NT>
static cudaError_t launch_march_cfg(const DGeo& g, const double* f, const double* vel, const DAccel& a, const DUpd& u,
                                    double* rhs_out, int flags, const DMom& mom, cudaStream_t st) {
  using C = MarchCfg<ORDER, T0, T1, T2>;
  const int nt0 = (g.n[0] + T0 - 1) / T0, nt1 = (g.n[1] + T1 - 1) / T1, nt2 = (g.n[2] + T2 - 1) / T2;
  int nchunk, chunk_len;
  march_plan<T2>(g, nt0 * nt1 * nt2, &nchunk, &chunk_len);
  const long long ctas = (long long)nt0 * nt1 * nt2 * nchunk;
  if (ctas > 0x7fffffffLL) return cudaErrorNotSupported;
  MarchMaps maps;
  memset(&maps, 0, sizeof(maps));
  static int use_tma_env = -1;
  if (use_tma_env < 0) {
    const char* e = getenv("LK_NO_TMA");
    use_tma_env = (e && e[0] == '1') ? 0 : 1;
  }
  const bool tma = use_tma_env && get_maps<ORDER, T0, T1, T2>(g, f, u, &maps);
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)ctas, NT, C::SMEM_BYTES, st>>>(g, f, vel, a, u, rhs_out, flags, nt0, nt1, nt2, chunk_len, mom, maps);
    return cudaGetLastError();
  };
  const bool lean = tma && flags == 3 && a.kind == 0 && a.bz == 0.0 && u.n_prev == 0 && !u.krook_nu;
  if (lean) return launch(k_stage_march<ORDER, T0, T1, T2, NT, true, true>);
  if (tma) return launch(k_stage_march<ORDER, T0, T1, T2, NT, true, false>);
  return launch(k_stage_march<ORDER, T0, T1, T2, NT, false, false>);
}

constexpr int MARCH_T2 = 8;
static int march_moment_parts(const DGeo& g) {
  const int nt0 = (g.n[0] + 31) / 32, nt1 = (g.n[1] + 7) / 8, nt2 = (g.n[2] + MARCH_T2 - 1) / MARCH_T2;
  int nchunk, chunk_len;
  march_plan<MARCH_T2>(g, nt0 * nt1 * nt2, &nchunk, &chunk_len);
  return nt2 * nchunk;
}
static cudaError_t launch_stage_march(const DGeo& g, const double* f, const double* vel, const DAccel& a, const DUpd& u,
                                      double* rhs_out, int flags, const DMom& mom, cudaStream_t st) {
  if (g.order == 4) return launch_march_cfg<4, 32, 8, MARCH_T2, 256>(g, f, vel, a, u, rhs_out, flags, mom, st);
  return launch_march_cfg<6, 32, 8, MARCH_T2, 256>(g, f, vel, a, u, rhs_out, flags, mom, st);
}

}  // namespace LK_NS
