// lk_pipe.cuh -- the production instantiation of the fused Vlasov stage kernel for aligned grids
// (n1 % 32 == 0, n2 % 8 == 0, n3 % 8 == 0), RK4-shaped stage updates, acceleration independent of the swept
// velocity index (non-relativistic Vlasov-Poisson without a constant B field: vel3 does not depend on i3,
// vel4 not on i4, KineticSpeciesF.f:78-80, 98-100).  Same tile, same march along vy, same per-cell arithmetic
// as k_stage_march (lk_march.cuh) -- a cell gets the same bits from either kernel -- but a software pipeline
// instead of three CTA-wide barriers per plane:
//
//   * warp w owns the vx slice c = w in the x and the y sweep and the tile row b1 = w in the vx sweep, the vy
//     fit and the epilogue; the only cross-warp hand-overs of a plane are
//        Q(p): "every slice's x+y accumulator is in shared memory"     (y sweep  -> column phase)
//        P(p): "every column's epilogue operands are in registers"     (epilogue -> next plane's x stores)
//     Both are mbarriers with one arrival per warp, and both are SPLIT: a warp arrives, keeps computing what
//     does not depend on the others (the vx fits between arrive-Q and wait-Q; the global stores of the
//     epilogue and the next plane's x fits between arrive-P and wait-P), and only then waits.
//   * TMA requests are issued by whichever warp arrives LAST at the hand-over that frees their landing zone
//     (a shared-memory counter next to the mbarrier), so no warp ever waits in order to issue a copy and no
//     fixed thread becomes the straggler: y halos of plane p+1 at Q(p); the core plane p+NG+1, the vx halos
//     and the delta_in tile of plane p+1 at P(p).
//   * the f_old tile of a plane lands in the accumulator it replaces, one warp's rows at a time: warp w reads
//     its columns of the accumulator (rows b1 = w of every slice, laid out contiguously per warp) and then
//     issues the TMA for exactly that region itself -- no barrier between draining and refilling.
//   * CTAs are rasterised in supertiles of gy x gv (y, vx) tiles over all x, so the y and vx star halos of a
//     tile are the core boxes of CTAs resident at the same time (L2 hits instead of DRAM re-reads).
//   * setAccelerationBCs4D (KineticSpeciesF.f:1036-1162) is folded into the tiles at the ends of the vx range and into
//     the first / last planes of the march (`bcfold`): the caller keeps the inflow sample in the velocity ghost layers,
//     the kernel overwrites the STAGED copy of an outflow column's ghosts with the extrapolation before it is read.
//   * a launch can be restricted to the tiles on the faces of the cut directions, or to the others (`bcfold` bits
//     4-7, lk_rk_update.tile_set): the two launches together are one, bit for bit.
//
// Reference arithmetic restated: KineticSpeciesF.f:723-790, 914-979 (fits), 1949-2245 (derivatives), 10-38
// (xpby4d); RK4Integrator.H:149-171; ReductionSchedule.C:421-444 (moments of the new predictor).
#pragma once
#include "lk_march.cuh"

#if !LK_STRICT
namespace LK_NS {

template <int ORDER>
struct PipeCfg {
  static constexpr int T0 = 32, T1 = 8, T2 = 8, NT = 256, NW = 8;
  static constexpr int NG = (ORDER == 4) ? 2 : 3;
  static constexpr int W = 2 * NG, NS = W - 1, SX = 8;
  static constexpr int PC = T0 + 2 * NG;
  static constexpr int NCORE = PC * T1 * T2, NYH = PC * NG * T2, NVH = PC * T1 * NG;
  // RK operand boxes start on a 16-byte boundary of the row (x = o0 + OPX) and are OPW wide; the tile's first
  // cell sits OPO elements into them
  static constexpr int OPX = NG & ~1, OPO = NG & 1, OPW = T0 + 2 * OPO;
  static constexpr int NOPW = OPW * T2;                 // f_old rows of one warp: [c][OPW]
  static constexpr int SB = (NOPW + T1 + 15) & ~15;     // one warp's accumulator region (128-byte multiple)
  static constexpr int NACC = SB * T1;
  static constexpr int NDI = OPW * T1 * T2;
  static constexpr int NBAR = NS + 2 + NW + 1 + 2;      // core slots, yh, vh, f_old per warp, delta_in, Q, P
  static constexpr int SMEM_DOUBLES = NS * NCORE + 2 * NYH + 2 * NVH + NACC + NDI;
  static constexpr size_t SMEM_BYTES = sizeof(double) * SMEM_DOUBLES + 8 * NBAR + sizeof(double) * 3 * 2 * T2 + 16;  // + counters, TMEM base
};

struct PipeMaps {
  CUtensorMap core, yh, vh;  // boxes of the array being differentiated
  CUtensorMap fo, di;        // f_old: one tile row of every slice (OPW, 1, T2); delta_in: the tile (OPW, T1, T2)
  CUtensorMap fot;           // f_old, the whole tile (OPW, T1, T2): L2 prefetch a plane ahead of the per-warp loads
};

// ---- shared memory through explicit 32-bit addresses: one base register plus immediates per access, no
// generic-address arithmetic in the plane loop (the kernel is bounded by instruction issue: one fp64
// instruction every other cycle per scheduler leaves one slot per fp64 instruction for everything else) ----
__device__ __forceinline__ double lds(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void lds2(unsigned a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ void sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
__device__ __forceinline__ void p_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void p_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p_tma(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void p_prefetch_l2(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// mbarrier wait: one try on the fast path.  try_wait returns within a few cycles when the phase is not complete,
// so a bare retry loop would burn the issue slots (and the scheduler priority) the other warps of the SM
// sub-partition need for their fp64 stream: the retry loop sleeps 64 ns between polls.  Watchdog: a hand-over
// that never completes -- a protocol bug -- traps after 2^22 polls (~0.3 s) instead of hanging the device.
#ifndef LK_PIPE_SLEEP_NS
#define LK_PIPE_SLEEP_NS 64
#endif
#ifndef LK_PIPE_WATCHDOG
#define LK_PIPE_WATCHDOG 1
#endif
#define LK_STR2(x) #x
#define LK_STR(x) LK_STR2(x)
__device__ __forceinline__ void p_wait(unsigned bar, unsigned parity) {
#if LK_PIPE_WATCHDOG
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .u32 n;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LK_DONE;\n\t"
      "mov.u32 n, 0;\n\t"
      "LK_RETRY:\n\t"
      "nanosleep.u32 " LK_STR(LK_PIPE_SLEEP_NS) ";\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LK_DONE;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 q, n, 4194304;\n\t"
      "@q bra LK_RETRY;\n\t"
      "trap;\n\t"
      "LK_DONE:\n\t"
      "}"
      ::"r"(bar), "r"(parity) : "memory");
#else
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LK_DONE;\n\t"
      "LK_RETRY:\n\t"
      "nanosleep.u32 " LK_STR(LK_PIPE_SLEEP_NS) ";\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra LK_RETRY;\n\t"
      "LK_DONE:\n\t"
      "}"
      ::"r"(bar), "r"(parity) : "memory");
#endif
}
// one arrival per warp at a hand-over; true (for the whole warp) in the warp that arrived last: it then owns the
// buffers the hand-over frees.  The count is relaxed; the last warp acquires the others' releases by
// observing the completed phase of the mbarrier itself before it lets the async proxy overwrite anything.
__device__ __forceinline__ bool p_handover(unsigned bar, unsigned cnt, unsigned parity, unsigned nw, int lane) {
  unsigned last = 0;
  if (lane == 0) {
    p_arrive(bar);
    unsigned old;
    asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(cnt) : "memory");
    last = ((old & (nw - 1)) == nw - 1) ? 1u : 0u;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (last) {
    p_wait(bar, parity);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  return last != 0;
}
// p + s*k elements as ONE integer instruction (IMAD.WIDE with the constant k*8 as immediate)
__device__ __forceinline__ double* ptr_off(double* p, int s, int k) {
  long long r;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(s), "r"(k * 8), "l"((long long)p));
  return (double*)r;
}
// global store through an explicit st.global (pointers that went through ptr_off lose their address space)
__device__ __forceinline__ void stg(double* p, double v) { asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v)); }
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
  return pred != 0;
}

#ifdef LK_PIPE_TRACE
// development aid (build_variant "trace"): per-warp clock stamps of 4 planes of 296 CTAs, read back by
// lk_debug_pipe_trace (tools/pipe_trace.py)
__device__ long long g_pipe_trace[296 * 8 * 4 * 8];
__device__ int g_pipe_smid[296];
#endif
// ---- tensor memory as a parking lot for loop-carried column state ----
// The vy fit of a column carries 2*T2 doubles from plane to plane (the oldest ring plane and the face below, per
// cell).  They are dead in every other phase of a plane, but as loop-carried registers they take 32 of the 128
// registers a thread has at two CTAs per SM -- exactly what the compiler needs to interleave independent fits in
// the x / y / vx sweeps (tools/fit_peak.cu: 89 % of the fp64 ceiling with that state live, 99 % without).  The SM's
// 256 KB of tensor memory is otherwise unused by this kernel and its 32x32b access shape is "lane i of warp w owns
// row 32 (w % 4) + i": a per-thread spill space.  One tcgen05.st after the vy fit, one tcgen05.ld before the next.
#ifndef LK_PIPE_TMEM
#define LK_PIPE_TMEM 0
#endif
// LK_PIPE_HWBAR = 1: the two hand-overs of a plane are hardware named barriers (bar.sync: a blocked warp costs no
// issue slot) instead of split mbarrier arrive / wait pairs whose retry path polls; warp 0 issues the copies
#ifndef LK_PIPE_HWBAR
#define LK_PIPE_HWBAR 0
#endif
__device__ __forceinline__ void tmem_st16(unsigned taddr, const double (&v)[8]) {
  unsigned r[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    r[2 * k] = (unsigned)__double2loint(v[k]);
    r[2 * k + 1] = (unsigned)__double2hiint(v[k]);
  }
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, double (&v)[8]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}

#ifndef LK_PIPE_FOLD
#define LK_PIPE_FOLD 1
#endif
#ifndef LK_PIPE_EDGE_FIRST
#define LK_PIPE_EDGE_FIRST 0
#endif

// EK: 1 = RK4 stage 1 (delta = w rhs), 2 = stages 2,3 (delta += w rhs), 3 = stage 4 (pred = f_old + c (delta + w rhs))
// NMOM: velocity moments of the new predictor left behind (0, 1: sum f, 3: + sum vx f, sum vy f)
template <int ORDER, int EK, int NMOM>
#ifndef LK_PIPE_MINB
#define LK_PIPE_MINB ((PipeCfg<ORDER>::SMEM_BYTES <= 113 * 1024) ? 2 : 1)
#endif
__global__ void __launch_bounds__(256, LK_PIPE_MINB)
k_stage_pipe(const DGeo g, const double* __restrict__ f, const double* __restrict__ vel,
             const double* __restrict__ afield, const DUpd upd, const int nt0, const int nt1, const int nt2,
             const int gy, const int gv, const int chunk_len, const DMom mom, const int bcfold,
             const __grid_constant__ PipeMaps maps) {
  using C = PipeCfg<ORDER>;
  constexpr int T0 = C::T0, T1 = C::T1, T2 = C::T2, NW = C::NW, NG = C::NG, W = C::W, NS = C::NS, SX = C::SX, PC = C::PC;
  constexpr int SB = C::SB, OPW = C::OPW, OPO = C::OPO;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // byte offsets of the regions
  constexpr unsigned O_CORE = 0;                                   // [slot][c][b1][PC]
  constexpr unsigned O_YH = O_CORE + 8u * NS * C::NCORE;           // [side][c][h][PC]
  constexpr unsigned O_VH = O_YH + 8u * 2 * C::NYH;                // [side][h][b1][PC]
  constexpr unsigned O_ACC = O_VH + 8u * 2 * C::NVH;               // warp b1: [c][32] at b1*(SB+1); f_old rows [c][OPW] at b1*SB
  constexpr unsigned O_DI = O_ACC + 8u * C::NACC;                  // [c][b1][OPW]
  constexpr unsigned O_BAR = O_DI + 8u * C::NDI;
  constexpr unsigned B_CORE = O_BAR, B_YH = O_BAR + 8u * NS, B_VH = B_YH + 8u, B_FO = B_VH + 8u, B_DI = B_FO + 8u * NW,
                     B_Q = B_DI + 8u, B_P = B_Q + 8u;
  constexpr unsigned O_VEL = O_BAR + 8u * C::NBAR;                 // [3][vx(c) | vy(c)]
  constexpr unsigned O_CNT = O_VEL + 8u * 3 * 2 * T2;              // [0]: Q arrivals, [1]: P arrivals
  unsigned sb = smem_u32(smem_raw);
  asm volatile("" : "+r"(sb));  // opaque: stays in a register instead of being re-derived (S2R + LEA) at every use

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- supertile rasterisation ----
  int b = blockIdx.x;
  const int o0 = (b % nt0) * T0;
  b /= nt0;
  int o1, o2;
  {
    const int nyv = nt1 * nt2;
    int r = b % nyv;
    b /= nyv;
    const int jv = r / (nt1 * gv);
    r -= jv * nt1 * gv;
    const int hv = min(gv, nt2 - jv * gv);
    const int jy = r / (gy * hv);
    r -= jy * gy * hv;
    const int hy = min(gy, nt1 - jy * gy);
    o1 = (jy * gy + r % hy) * T1;
    int iv = jv * gv + r / hy;
#if LK_PIPE_EDGE_FIRST
    // the two vx tiles that carry the folded velocity-boundary fill do a little more per plane: schedule them first
    // instead of leaving the top one for the last, partly filled wave
    if (nt2 > 2) iv = (iv == 0) ? 0 : ((iv == 1) ? nt2 - 1 : iv - 1);
#endif
    o2 = iv * T2;
  }
  // tile subsets (bcfold bits 4-5: 1 = only the tiles on a face of a cut direction, 2 = only the others; bits 6-7: the
  // cut directions, 1 x, 2 y): a rank whose configuration space is cut launches the face tiles first, so that the halo
  // exchange of the new predictor travels while the remaining tiles are still being computed
  if (bcfold & 0x30) {
    const int cutd = (bcfold >> 6) & 3;
    const bool face = ((cutd & 1) && (o0 == 0 || o0 + T0 == g.n[0])) || ((cutd & 2) && (o1 == 0 || o1 + T1 == g.n[1]));
    if (face != (((bcfold >> 4) & 3) == 1)) return;
  }
  const int chunk = b;
  const int q0 = chunk * chunk_len;              // first interior vy plane of this CTA
  const int nq = min(chunk_len, g.n[3] - q0);
  if (nq <= 0) return;
  const int X0 = o0, Y0 = o1 + NG, V0 = o2 + NG;  // data-box origin of the staged boxes (x grown by NG)
  const int pbase = q0 + 1;                       // data index of the plane in ring slot 0 at start

  const double FS = FaceScale<ORDER>::v;
  const double rdx0 = (1.0 / g.dx[0]) * FS, rdx1 = (1.0 / g.dx[1]) * FS, rdx2 = (1.0 / g.dx[2]) * FS,
               rdx3 = (1.0 / g.dx[3]) * FS;

  // the core plane pc lands in ring slot `slot`; with_di: the delta_in tile of plane pd rides on the same barrier
  // (both are requested at the same hand-over and the vy fit needs the core plane before the epilogue needs the
  // tile: one wait instead of two)
  auto issue_core = [&](int pc, int slot, bool with_di, int pd) {
    const unsigned bar = sb + B_CORE + 8u * slot;
    p_expect(bar, (unsigned)((C::NCORE + (with_di ? C::NDI : 0)) * sizeof(double)));
    p_tma(sb + O_CORE + 8u * C::NCORE * slot, &maps.core, bar, X0, Y0, V0, pc);
    if (with_di) p_tma(sb + O_DI, &maps.di, bar, o0 + C::OPX, o1 + NG, o2 + NG, pd);
  };
  // y and vx star halos of plane p: one barrier
  auto issue_halos = [&](int p) {
    p_expect(sb + B_YH, (unsigned)(2 * (C::NYH + C::NVH) * sizeof(double)));
    p_tma(sb + O_YH, &maps.yh, sb + B_YH, X0, Y0 - NG, V0, p);
    p_tma(sb + O_YH + 8u * C::NYH, &maps.yh, sb + B_YH, X0, Y0 + T1, V0, p);
    p_tma(sb + O_VH, &maps.vh, sb + B_YH, X0, Y0, V0 - NG, p);
    p_tma(sb + O_VH + 8u * C::NVH, &maps.vh, sb + B_YH, X0, Y0, V0 + T2, p);
  };

  if (tid == 0) {
    for (int k = 0; k < C::NBAR; ++k) {
      const unsigned n = (k >= NS + 2 + NW + 1) ? NW : 1;  // Q and P take one arrival per warp
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sb + O_BAR + 8u * k), "r"(n));
    }
    asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(sb + O_CNT), "r"(0u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#if LK_PIPE_TMEM
  static_assert(T2 == 8, "the parked state is two groups of 8 doubles per thread");
  // 64 columns: (uold, Fprev) = 32 columns for warps 0-3 (lane quarters 0..3), 32 more for warps 4-7 on the same lanes
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb + O_CNT + 8u), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
  __syncthreads();
#if LK_PIPE_TMEM
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  unsigned tm_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tm_base) : "r"(sb + O_CNT + 8u));
  const unsigned tm_u = tm_base + ((32u * (warp & 3)) << 16) + 32u * (warp >> 2);  // uold: 16 columns, Fprev: the next 16
#endif

  // ---- this thread's (x,y) column for the vx sweep, the vy fit and the epilogue: (lane, warp) ----
  const i64 col = (i64)(o0 + lane + NG) + g.s[1] * (o1 + warp + NG) + g.s[2] * (o2 + NG);  // + s2*c + s3*p
  const i64 pxy = (o0 + lane + NG) + (i64)g.nd[0] * (o1 + warp + NG);
  // per-thread byte offsets into the regions
  const unsigned t_col = 8u * (warp * PC + NG + lane);                     // core / vx-halo column cell, + c*T1*PC (h*T1*PC)
  const unsigned t_y = 8u * (warp * T1 * PC + NG + lane);                  // core y line of slice c = warp, + b1*PC
  const unsigned t_yh = 8u * (warp * NG * PC + NG + lane);                 // y halo of slice c = warp, + h*PC
  const int xb1 = lane & 7, xseg = lane >> 3;                              // x sweep: row b1, segment of this lane
  const unsigned t_x = 8u * ((warp * T1 + xb1) * PC + xseg * SX);          // core x segment, slice c = warp
  const unsigned t_xa = O_ACC + 8u * (xb1 * (SB + 1) + warp * 32 + xseg * SX);
  const unsigned t_ya = O_ACC + 8u * (warp * 32 + lane);                   // + b1*(SB+1)
  const unsigned t_ca = O_ACC + 8u * (warp * (SB + 1) + lane);             // + c*32
  const unsigned t_fo = O_ACC + 8u * (warp * SB + OPO + lane);             // + c*OPW
  const unsigned t_di = O_DI + 8u * (warp * OPW + OPO + lane);             // + c*T1*OPW

  // ---- prologue: fill the ring, fit the face below the first plane ----
  if (tid == 0) {
    for (int k = 0; k < NS; ++k) issue_core(pbase + k, k, false, 0);
    issue_halos(q0 + NG);
    p_prefetch_l2(&maps.fot, o0 + C::OPX, o1 + NG, o2 + NG, q0 + NG);
  }
  double uold[T2], Fprev[T2];  // with LK_PIPE_TMEM: the prologue's copy only (parked below, shadowed in the vy phase)
#pragma unroll
  for (int c = 0; c < T2; ++c) uold[c] = f[col + g.s[2] * c + g.s[3] * (pbase - 1)];
  // velocities of the slices: threads 0..2*T2-1 fetch them a plane ahead; running pointer, one plane per step
  const double* velp = vel + (o2 + (tid % T2) + NG) + (i64)g.nd[2] * ((q0 + NG) + (i64)((tid / T2) & 1) * g.nd[3]);
  if (tid < 2 * T2) sts(sb + O_VEL + 8u * tid, __ldg(velp));
  // E (times q/m) at this thread's (x,y): constant along the whole march
  const double ax0 = __ldg(afield + pxy);
  const double ay0 = __ldg(afield + pxy + (i64)g.nd[0] * g.nd[1]);
  const bool axpos = ax0 > 0.0, aypos = ay0 > 0.0;
  const double kax = MUL(ax0, rdx2), kay = MUL(ay0, rdx3);
  for (int k = 0; k < NS; ++k) p_wait(sb + B_CORE + 8u * k, 0);
  unsigned phc = (1u << NS) - 1;  // phase parity of every core slot's next completion
  // Folded velocity-boundary fill (setAccelerationBCs4D, KineticSpeciesF.f:1036-1162; bcfold bit 0: vx, bit 1: vy).
  // The caller keeps the INFLOW sample of the initial condition in f's velocity ghost layers (it depends on the
  // position only: lk_preset_inflow_ghosts_4d writes it once per array), so a ghost cell whose face has the
  // acceleration pointing inward is already right when TMA brings it in.  Where it points outward the ghost is the
  // extrapolation 3 u_-1 - 3 u_-2 + u_-3 marching outward (:1079-1096, :1124-1141): computed here by the thread that
  // owns the column and written over the staged copy in shared memory -- never to global memory, so the preset
  // survives.  Rolled loops: boundary tiles only, the march loop's code must stay small.
#if LK_PIPE_FOLD
  const bool fold_vx = (bcfold & 1) != 0, fold_vy = (bcfold & 2) != 0;
#else
  constexpr bool fold_vx = false, fold_vy = false;
#endif
  // the NG planes below the first interior vy plane: outflow if a_y <= 0 at the lower face (:1143)
  const bool low_ghosts = fold_vy && q0 == 0 && !aypos;
  if (low_ghosts) {
#pragma unroll 1
    for (int c = 0; c < T2; ++c) {
      const unsigned cell = sb + O_CORE + 8u * (c * T1 * PC) + t_col;   // + slot*NCORE: plane 1 + slot
      // interior planes NG, NG+1, NG+2 = ring slots NG-1, NG, NG+1 (the last one is not staged yet at order 4)
      double u0 = lds(cell + 8u * C::NCORE * (NG - 1)), u1 = lds(cell + 8u * C::NCORE * NG);
      double u2 = (NG + 1 < NS) ? lds(cell + 8u * C::NCORE * ((NG + 1 < NS) ? NG + 1 : 0))
                                : f[col + g.s[2] * c + g.s[3] * (NG + 2)];
#pragma unroll 1
      for (int ig = 1; ig <= NG; ++ig) {   // ghost plane NG - ig
        const double gv_ = bc_extrap(u0, u1, u2);
        u2 = u1; u1 = u0; u0 = gv_;
        // planes 1 .. NG-1 live in slots 0 .. NG-2; plane 0 (the first window's lowest) in this thread's accumulator cell
        sts((ig < NG) ? (cell + 8u * C::NCORE * (NG - 1 - ig)) : (sb + t_ca + 8u * (c * 32)), gv_);
      }
    }
  }
  // a last chunk shorter than NG planes already holds planes above the last interior one in its first window (the
  // march patches the newest plane of each later window, below): outflow if a_y >= 0 at the upper face (:1124)
  // one register of loop-invariant flags for the march: bit 0 / 1: this tile extrapolates its vx-low / vx-high ghosts
  // (bottom: inflow if a_x > 0 at the lower face, :1098; top: outflow if a_x >= 0 at the upper face, :1079), bit 2: its
  // vy-high ghosts
  const unsigned bcf = ((fold_vx && o2 == 0 && !axpos) ? 1u : 0u) | ((fold_vx && o2 + T2 == g.n[2] && ax0 >= 0.0) ? 2u : 0u) |
                       ((fold_vy && ay0 >= 0.0) ? 4u : 0u);
  if ((bcf & 4u) && pbase + NS - 1 > g.n[3] + NG - 1) {
#pragma unroll 1
    for (int c = 0; c < T2; ++c) {
      const unsigned cell = sb + O_CORE + 8u * (c * T1 * PC) + t_col;
#pragma unroll 1
      for (int k = 0; k < NS; ++k) {
        if (pbase + k <= g.n[3] + NG - 1) continue;
        auto below = [&](int j) -> double {   // plane pbase + j: interior planes when they are not in the ring (n4 >= 2 NG)
          return (j >= 0) ? lds(cell + 8u * C::NCORE * (j >= 0 ? j : 0)) : f[col + g.s[2] * c + g.s[3] * (pbase + j)];
        };
        sts(cell + 8u * C::NCORE * k, bc_extrap(below(k - 1), below(k - 2), below(k - 3)));
      }
    }
  }
  {
#pragma unroll
    for (int c = 0; c < T2; ++c) {
      double w[W];
      w[0] = low_ghosts ? lds(sb + t_ca + 8u * (c * 32)) : uold[c];
#pragma unroll
      for (int k = 0; k < NS; ++k) w[k + 1] = lds(sb + O_CORE + 8u * (k * C::NCORE + c * T1 * PC) + t_col);
      Fprev[c] = fit_face<ORDER>(w, aypos);
    }
  }
#pragma unroll
  for (int c = 0; c < T2; ++c) uold[c] = lds(sb + O_CORE + 8u * (c * T1 * PC) + t_col);  // plane pbase
#if LK_PIPE_TMEM
  tmem_st16(tm_u, uold);
  tmem_st16(tm_u + 16u, Fprev);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
  __syncthreads();
  if (tid == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    issue_core(pbase + NS, 0, EK >= 2, q0 + NG);  // plane p0+NG and the first delta_in tile: waited for by the first vy fit
  }

  double m0 = 0.0, m1 = 0.0, m2 = 0.0;
  // periodic ghost copies of the predictor written by the cell that owns the value (upd.wrap)
  int gxo = 0, gyo = 0;
  {
    const int x = o0 + lane, y = o1 + warp;
    if ((upd.wrap & 1) && g.n[0] >= 2 * NG) gxo = (x < NG) ? g.n[0] : ((x >= g.n[0] - NG) ? -g.n[0] : 0);
    if ((upd.wrap & 2) && g.n[1] >= 2 * NG) gyo = (y < NG) ? g.n[1] * (int)g.s[1] : ((y >= g.n[1] - NG) ? -g.n[1] * (int)g.s[1] : 0);
  }
  const i64 s3 = g.s[3];
  const int s2 = (int)g.s[2];
  double* pr_p = upd.pred + col + s3 * (q0 + NG);        // this column at the plane being updated; += s3 per plane
  double* do_p = (EK == 3) ? nullptr : upd.delta_out + col + s3 * (q0 + NG);
  const double w_delta = upd.w_delta, c_pred = upd.c_pred;

#ifdef LK_PIPE_TRACE
  int smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const bool trace_on = (blockIdx.x >= 1480 && blockIdx.x < 1480 + 296);
  if (trace_on && tid == 0) g_pipe_smid[blockIdx.x - 1480] = smid;
#endif
  int sc = NG - 1;       // ring slot of the plane being updated
  unsigned vb = 0;       // velocity buffer of the plane being updated (byte offset into O_VEL)
  unsigned pp = 0;       // parity of the per-plane barriers (yh, vh, f_old / delta_in, Q); P runs one behind
  for (int q = q0; q < q0 + nq; ++q) {
    const int p = q + NG;
    const unsigned cur = sb + O_CORE + 8u * C::NCORE * sc;
    const bool more = (q + 1 < q0 + nq);
    const unsigned sv = sb + O_VEL + vb;
    const unsigned vbn = (vb == 2u * 8u * 2 * T2) ? 0u : vb + 8u * 2 * T2;
    double vel_next = 0.0;
    if (more && tid < 2 * T2) {
      velp += g.nd[2];
      vel_next = __ldg(velp);
    }

#ifdef LK_PIPE_TRACE
#define LK_TR(ev) do { if (trace_on && lane == 0 && q >= q0 + 40 && q < q0 + 44) g_pipe_trace[((((int)blockIdx.x - 1480) * 8 + warp) * 4 + (q - q0 - 40)) * 8 + ev] = clock64(); } while (0)
#else
#define LK_TR(ev)
#endif
    LK_TR(0);
    // ---------------- A: x fits of slice c = warp, rows b1 = lane % 8, segments of SX cells ----------------
    double xr[SX];
    {
      const double vx = lds(sv + 8u * warp);
      const bool pos = vx > 0.0;
      const double kx = MUL(vx, rdx0);
      double v[SX + W];
#pragma unroll
      for (int k = 0; k < (SX + W) / 2; ++k) lds2(cur + t_x + 16u * k, v[2 * k], v[2 * k + 1]);
      Walker<ORDER> wk;
      wk.init([&](int k) { return v[k]; });
      double uL = wk.next(v[W - 1], pos);
#pragma unroll
      for (int k = 0; k < SX; ++k) {
        const double uR = wk.next(v[k + W], pos);
        xr[k] = FMA(-kx, ADD(uR, -uL), 0.0);
        uL = uR;
      }
    }
    // ---------------- B: the accumulator is free once every column of the previous plane has its operands ----
    LK_TR(1);
#if LK_PIPE_HWBAR
    if (q > q0) {
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (warp == 0 && elect_one()) {   // everything plane p-1 was staged in is free
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        int sf = sc + NG;             // the slot the previous plane's oldest ring plane lived in
        if (sf >= NS) sf -= NS;
        issue_core(p + NG, sf, EK >= 2, p);
        p_prefetch_l2(&maps.fot, o0 + C::OPX, o1 + NG, o2 + NG, p);
      }
    }
#else
    if (q > q0) p_wait(sb + B_P, pp ^ 1u);
#endif
    LK_TR(2);
    if (more && tid < 2 * T2) sts(sb + O_VEL + vbn + 8u * tid, vel_next);
#pragma unroll
    for (int k = 0; k < SX; ++k) sts(sb + t_xa + 8u * k, xr[k]);
    __syncwarp();

    // ---------------- C: y sweep of slice c = warp, lines a0 = lane ----------------
    p_wait(sb + B_YH, pp);
    {
      const double vy = lds(sv + 8u * (T2 + warp));
      const bool pos = vy > 0.0;
      const double ky = MUL(vy, rdx1);
      double v[T1 + W], acc[T1];
#pragma unroll
      for (int k = 0; k < NG; ++k) v[k] = lds(sb + O_YH + t_yh + 8u * (k * PC));
#pragma unroll
      for (int k = 0; k < T1; ++k) v[NG + k] = lds(cur + t_y + 8u * (k * PC));
#pragma unroll
      for (int k = 0; k < NG; ++k) v[NG + T1 + k] = lds(sb + O_YH + 8u * C::NYH + t_yh + 8u * (k * PC));
#pragma unroll
      for (int b1 = 0; b1 < T1; ++b1) acc[b1] = lds(sb + t_ya + 8u * (b1 * (SB + 1)));
      Walker<ORDER> wk;
      wk.init([&](int k) { return v[k]; });
      double uL = wk.next(v[W - 1], pos);
#pragma unroll
      for (int b1 = 0; b1 < T1; ++b1) {
        const double uR = wk.next(v[b1 + W], pos);
        acc[b1] = FMA(-ky, ADD(uR, -uL), acc[b1]);
        uL = uR;
      }
#pragma unroll
      for (int b1 = 0; b1 < T1; ++b1) sts(sb + t_ya + 8u * (b1 * (SB + 1)), acc[b1]);
    }
    // the vx line of this thread's column goes into registers BEFORE the hand-over, so that the hand-over frees
    // the vx halos as well as the y halos and both are re-armed a whole plane ahead of their use
    double res[T2];
    // folded velocity-boundary fill in vx (KineticSpeciesF.f:1072-1113): the tiles at the two ends of the vx range
    // overwrite the staged ghost cells of their outflow columns before the line is loaded
    if (bcf & 3u) {
      auto cell = [&](int k) -> unsigned {   // window position k = data index o2 + k along vx
        return (k < NG) ? (sb + O_VH + t_col + 8u * (k * T1 * PC))
                        : ((k < NG + T2) ? (cur + t_col + 8u * ((k - NG) * T1 * PC)) : (sb + O_VH + 8u * C::NVH + t_col + 8u * ((k - NG - T2) * T1 * PC)));
      };
#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        if (!(bcf & (1u << side))) continue;
        const int k0 = (side == 0) ? NG : NG + T2 - 1, step = (side == 0) ? -1 : 1;   // the boundary cell, outward step
#pragma unroll 1
        for (int ig = 1; ig <= NG; ++ig) {
          const int k = k0 + step * ig;
          sts(cell(k), bc_extrap(lds(cell(k - step)), lds(cell(k - 2 * step)), lds(cell(k - 3 * step))));
        }
      }
    }
    {
      double v[T2 + W];
#pragma unroll
      for (int k = 0; k < NG; ++k) v[k] = lds(sb + O_VH + t_col + 8u * (k * T1 * PC));
#pragma unroll
      for (int k = 0; k < T2; ++k) v[NG + k] = lds(cur + t_col + 8u * (k * T1 * PC));
#pragma unroll
      for (int k = 0; k < NG; ++k) v[NG + T2 + k] = lds(sb + O_VH + 8u * C::NVH + t_col + 8u * (k * T1 * PC));
      __syncwarp();
      LK_TR(3);
#if !LK_PIPE_HWBAR
      if (p_handover(sb + B_Q, sb + O_CNT, pp, NW, lane) && more) {
        if (elect_one()) issue_halos(p + 1);
      }
#endif

      // ---------------- D: vx fits of this thread's column (lane, warp), all c ----------------
      Walker<ORDER> wk;
      wk.init([&](int k) { return v[k]; });
      double uL = wk.next(v[W - 1], axpos);
#pragma unroll
      for (int c = 0; c < T2; ++c) {
        const double uR = wk.next(v[c + W], axpos);
        res[c] = ADD(uR, -uL);
        uL = uR;
      }
    }
    // ---------------- E: add the x+y accumulator, hand its rows over to the f_old tile ----------------
    LK_TR(4);
#if LK_PIPE_HWBAR
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (more && warp == 0 && elect_one()) {   // every warp has its vx line in registers and its y sweep stored
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue_halos(p + 1);
    }
#else
    p_wait(sb + B_Q, pp);
#endif
    LK_TR(5);
#pragma unroll
    for (int c = 0; c < T2; ++c) res[c] = FMA(-kax, res[c], lds(sb + t_ca + 8u * (c * 32)));
    __syncwarp();
    if (elect_one()) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      p_expect(sb + B_FO + 8u * warp, (unsigned)(C::NOPW * sizeof(double)));
      p_tma(sb + O_ACC + 8u * SB * warp, &maps.fo, sb + B_FO + 8u * warp, o0 + C::OPX, o1 + NG + warp, o2 + NG, p);
    }

    // ---------------- F: vy face above this plane ----------------
    {
      int sn = sc + NG;
      if (sn >= NS) sn -= NS;
      p_wait(sb + B_CORE + 8u * sn, (phc >> sn) & 1u);  // plane p+NG
      phc ^= 1u << sn;
#if LK_PIPE_TMEM
      double uold[T2], Fprev[T2];
      tmem_ld16(tm_u, uold);
      tmem_ld16(tm_u + 16u, Fprev);
#endif
      unsigned wp[W];
#pragma unroll
      for (int k = 1; k < W; ++k) {  // planes p-NG+2 .. p+NG
        int s = sc + k - (NG - 1);
        if (s < 0) s += NS;
        if (s >= NS) s -= NS;
        wp[k] = sb + O_CORE + 8u * C::NCORE * s + t_col;
      }
      // folded velocity-boundary fill above the last interior vy plane (KineticSpeciesF.f:1117-1140): the newest plane
      // of the window is a ghost plane for the last NG planes of the march -- an outflow column extrapolates it from
      // the three values below it and writes it over the ring's copy, where the next planes' windows read it
      if ((bcf & 4u) && p >= g.n[3]) {
#pragma unroll 1
        for (int c = 0; c < T2; ++c) {
          const unsigned o = 8u * (c * T1 * PC);
          // the window's lowest plane left the ring a plane ago (order 4 only needs it): an interior plane, from memory
          const double w3 = (W - 4 >= 1) ? lds(wp[(W - 4 >= 1) ? W - 4 : 1] + o) : __ldg(f + col + g.s[2] * c + g.s[3] * (p - NG + 1));
          sts(wp[W - 1] + o, bc_extrap(lds(wp[W - 2] + o), lds(wp[W - 3] + o), w3));
        }
      }
#pragma unroll
      for (int c = 0; c < T2; ++c) {
        double w[W];
        w[0] = uold[c];
#pragma unroll
        for (int k = 1; k < W; ++k) w[k] = lds(wp[k] + 8u * (c * T1 * PC));
        const double F = fit_face<ORDER>(w, aypos);
        res[c] = FMA(-kay, ADD(F, -Fprev[c]), res[c]);
        Fprev[c] = F;
        uold[c] = w[1];
      }
#if LK_PIPE_TMEM
      if (more) {
        tmem_st16(tm_u, uold);
        tmem_st16(tm_u + 16u, Fprev);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
#endif
    }

    // ---------------- G/H: RK stage update straight to global memory, moments of the new predictor ----------
    LK_TR(6);
    p_wait(sb + B_FO + 8u * warp, pp);
    LK_TR(7);
    {
      double pr[T2];
#pragma unroll
      for (int c = 0; c < T2; ++c) {
        const double fo = lds(sb + t_fo + 8u * (c * OPW));
        if constexpr (EK == 1) {
          stg(ptr_off(do_p, s2, c), MUL(w_delta, res[c]));
          pr[c] = FMA(c_pred, res[c], fo);
        } else if constexpr (EK == 2) {
          const double di = lds(sb + t_di + 8u * (c * T1 * OPW));
          stg(ptr_off(do_p, s2, c), FMA(w_delta, res[c], di));
          pr[c] = FMA(c_pred, res[c], fo);
        } else {
          const double di = lds(sb + t_di + 8u * (c * T1 * OPW));
          pr[c] = FMA(c_pred, FMA(w_delta, res[c], di), fo);
        }
      }
      // every operand of this plane is in registers: hand the plane's buffers over
      __syncwarp();
#if !LK_PIPE_HWBAR
      if (more) {
        if (p_handover(sb + B_P, sb + O_CNT + 4u, pp, NW, lane)) {
          if (elect_one()) {
            int sf = sc + NG + 1;  // the slot of the oldest plane of the ring (now in registers)
            if (sf >= NS) sf -= NS;
            if (sf >= NS) sf -= NS;
            issue_core(p + NG + 1, sf, EK >= 2, p + 1);
            p_prefetch_l2(&maps.fot, o0 + C::OPX, o1 + NG, o2 + NG, p + 1);
          }
        }
      }
#endif
#pragma unroll
      for (int c = 0; c < T2; ++c) stg(ptr_off(pr_p, s2, c), pr[c]);
      if (gxo) {
#pragma unroll
        for (int c = 0; c < T2; ++c) stg(ptr_off(pr_p + gxo, s2, c), pr[c]);
      }
      if (gyo) {
#pragma unroll
        for (int c = 0; c < T2; ++c) stg(ptr_off(pr_p + gyo, s2, c), pr[c]);
        if (gxo) {
#pragma unroll
          for (int c = 0; c < T2; ++c) stg(ptr_off(pr_p + gyo + gxo, s2, c), pr[c]);
        }
      }
      if constexpr (NMOM > 0) {
        double psum = 0.0, pvx = 0.0, pvy = 0.0;
#pragma unroll
        for (int c = 0; c < T2; ++c) {
          psum = ADD(psum, pr[c]);
          if constexpr (NMOM > 1) {
            pvx = FMA(lds(sv + 8u * c), pr[c], pvx);
            pvy = FMA(lds(sv + 8u * (T2 + c)), pr[c], pvy);
          }
        }
        m0 = ADD(m0, psum);
        if constexpr (NMOM > 1) {
          m1 = ADD(m1, pvx);
          m2 = ADD(m2, pvy);
        }
      }
    }
    pr_p += s3;
    if (EK != 3) do_p += s3;
    sc = (sc + 1 == NS) ? 0 : sc + 1;
    vb = vbn;
    pp ^= 1u;
  }

#if LK_PIPE_TMEM
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "r"(64u) : "memory");
#endif
  if constexpr (NMOM > 0) {
    const i64 nxy = (i64)g.n[0] * g.n[1];
    const i64 part = (i64)chunk * nt2 + (o2 / T2);
    const i64 o = (o0 + lane) + (i64)g.n[0] * (o1 + warp);
    mom.part[part * nxy + o] = m0;
    if constexpr (NMOM > 1) {
      mom.part[((i64)mom.nparts + part) * nxy + o] = m1;
      mom.part[((i64)2 * mom.nparts + part) * nxy + o] = m2;
    }
  }
}

// ---- launch ----
static bool pipe_eligible(const DGeo& g, const DAccel& a, const DUpd& u, const double* rhs_out, int flags) {
  using C4 = PipeCfg<4>;
  if (!u.active || rhs_out || flags != 3 || u.n_prev != 0 || u.krook_nu) return false;
  if (a.kind != 0 || a.bz != 0.0) return false;
  if ((g.n[0] % C4::T0) || (g.n[1] % C4::T1) || (g.n[2] % C4::T2)) return false;
  const bool k1 = !u.delta_in && u.delta_out && !u.use_delta;
  const bool k2 = u.delta_in && u.delta_out && !u.use_delta;
  const bool k3 = u.delta_in && !u.delta_out && u.use_delta;
  return k1 || k2 || k3;
}

template <int ORDER>
static cudaError_t launch_pipe(const DGeo& g, const double* f, const double* vel, const DAccel& a, const DUpd& u,
                               const DMom& mom, int bcfold, cudaStream_t st, bool* used) {
  using C = PipeCfg<ORDER>;
  *used = false;
  PipeMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (!get_map(g, f, C::PC, C::T1, C::T2, &maps.core)) return cudaSuccess;
  if (!get_map(g, f, C::PC, C::NG, C::T2, &maps.yh)) return cudaSuccess;
  if (!get_map(g, f, C::PC, C::T1, C::NG, &maps.vh)) return cudaSuccess;
  if (!get_map(g, u.f_old, C::OPW, 1, C::T2, &maps.fo)) return cudaSuccess;
  if (!get_map(g, u.f_old, C::OPW, C::T1, C::T2, &maps.fot)) return cudaSuccess;
  maps.di = maps.fo;
  if (u.delta_in && !get_map(g, u.delta_in, C::OPW, C::T1, C::T2, &maps.di)) return cudaSuccess;
  const int nt0 = g.n[0] / C::T0, nt1 = g.n[1] / C::T1, nt2 = g.n[2] / C::T2;
  int nchunk, chunk_len;
  march_plan<C::T2>(g, nt0 * nt1 * nt2, &nchunk, &chunk_len);
  const long long ctas = (long long)nt0 * nt1 * nt2 * nchunk;
  if (ctas > 0x7fffffffLL) return cudaSuccess;
  // supertile shape: about one wave of resident CTAs, as square as the tile counts allow
  static int env_gy = -1, env_gv = -1;
  if (env_gy < 0) {
    const char* e = getenv("LK_PIPE_GY");
    env_gy = e ? atoi(e) : 0;
    e = getenv("LK_PIPE_GV");
    env_gv = e ? atoi(e) : 0;
  }
  int gy = env_gy, gv = env_gv;
  if (gy <= 0 || gv <= 0) {
    const int resident = 148 * ((C::SMEM_BYTES <= 113 * 1024) ? 2 : 1);
    int per = resident / (nt0 > 0 ? nt0 : 1);
    if (per < 1) per = 1;
    gy = 1;
    while (gy * gy < per) ++gy;
    if (gy > nt1) gy = nt1;
    gv = (per + gy - 1) / gy;
    if (gv > nt2) gv = nt2;
    if (gv < 1) gv = 1;
  }
  const int ek = (!u.delta_in) ? 1 : (u.delta_out ? 2 : 3);
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)ctas, C::NT, C::SMEM_BYTES, st>>>(g, f, vel, a.field, u, nt0, nt1, nt2, gy, gv, chunk_len, mom, bcfold, maps);
    return cudaGetLastError();
  };
  *used = true;
  const int nm = mom.nmom;
  if (nm != 0 && nm != 1 && nm != 3) return cudaErrorInvalidValue;
#define LK_PIPE_CASE(E, M) if (ek == E && nm == M) return launch(k_stage_pipe<ORDER, E, M>);
  LK_PIPE_CASE(1, 0) LK_PIPE_CASE(1, 1) LK_PIPE_CASE(1, 3)
  LK_PIPE_CASE(2, 0) LK_PIPE_CASE(2, 1) LK_PIPE_CASE(2, 3)
  LK_PIPE_CASE(3, 0) LK_PIPE_CASE(3, 1) LK_PIPE_CASE(3, 3)
#undef LK_PIPE_CASE
  return cudaErrorInvalidValue;
}

}  // namespace LK_NS
#endif  // !LK_STRICT
