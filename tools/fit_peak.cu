// fit_peak.cu -- the issue ceiling of the production WENO fit stream itself: the sliding-window walker of
// lk_march.cuh (Walker<4>/Walker<6>, 9 faces per 8-cell line exactly as the stage kernel's sweeps execute them)
// run from registers only -- no shared memory, no barriers, no TMA -- at the stage kernel's occupancy
// (256 threads, 128 registers, 2 CTAs/SM for order 4; 255 registers, 1 CTA/SM for order 6).  What this reaches
// is what a perfectly overlapped stage kernel could reach; the gap between it and the fp64 paper ceiling is
// the in-order dependent-issue cost of the arithmetic, not a memory or synchronisation effect.
// Also: fp64 instructions with two / three live register sources (register-file read bandwidth).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tools/fit_peak tools/fit_peak.cu
#define LK_STRICT 0
#include <stdio.h>
#include <stdlib.h>

#include "../loki_b200/csrc/lk_march.cuh"

using namespace lkfast;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

template <int ORDER, int MINB>
__global__ void __launch_bounds__(256, MINB) k_fits(double* out, const double* in, int iters, double k) {
  constexpr int W = (ORDER == 4) ? 4 : 6, SX = 8;
  double v[SX + W], acc[SX];
#pragma unroll
  for (int j = 0; j < SX + W; ++j) v[j] = in[threadIdx.x + 256 * j];
#pragma unroll
  for (int j = 0; j < SX; ++j) acc[j] = 0.0;
  const bool pos = in[threadIdx.x] > 0.5;
  for (int it = 0; it < iters; ++it) {
    Walker<ORDER> wk;
    wk.init([&](int j) { return v[j]; });
    double uL = wk.next(v[W - 1], pos);
#pragma unroll
    for (int j = 0; j < SX; ++j) {
      const double uR = wk.next(v[j + W], pos);
      acc[j] = FMA(-k, ADD(uR, -uL), acc[j]);
      uL = uR;
    }
    // keep the line changing (one extra fp64 instruction per cell, counted below)
#pragma unroll
    for (int j = 0; j < SX; ++j) v[j + W / 2] = FMA(1e-9, acc[j], v[j + W / 2]);
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < SX; ++j) s += acc[j];
  if (s == 123.456) out[0] = s;
}

// the same stream under the stage kernel's register pressure: NB doubles of "column state" stay live across the
// sweep (as uold / Fprev / pointers do in k_stage_pipe), MAXR caps the registers like the kernel's occupancy does
template <int NT, int MINB, int NB>
__global__ void __launch_bounds__(NT, MINB) k_fits_pressure(double* out, const double* in, int iters, double k) {
  constexpr int W = 4, SX = 8;
  double v[SX + W], acc[SX], ballast[NB];
#pragma unroll
  for (int j = 0; j < SX + W; ++j) v[j] = in[(threadIdx.x & 255) + 256 * j];
#pragma unroll
  for (int j = 0; j < NB; ++j) ballast[j] = in[(threadIdx.x & 255) + 256 * (j % 16)] + j;
#pragma unroll
  for (int j = 0; j < SX; ++j) acc[j] = 0.0;
  const bool pos = in[threadIdx.x & 255] > 0.5;
  for (int it = 0; it < iters; ++it) {
    Walker<4> wk;
    wk.init([&](int j) { return v[j]; });
    double uL = wk.next(v[W - 1], pos);
#pragma unroll
    for (int j = 0; j < SX; ++j) {
      const double uR = wk.next(v[j + W], pos);
      acc[j] = FMA(-k, ADD(uR, -uL), acc[j]);
      uL = uR;
    }
#pragma unroll
    for (int j = 0; j < SX; ++j) v[j + W / 2] = FMA(1e-9, acc[j], v[j + W / 2]);
#pragma unroll
    for (int j = 0; j < NB; ++j) asm volatile("" : "+d"(ballast[j]));  // live in registers once per line
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < SX; ++j) s += acc[j];
#pragma unroll
  for (int j = 0; j < NB; ++j) s += ballast[j];
  if (s == 123.456) out[0] = s;
}

template <int OP>
__global__ void __launch_bounds__(256) k_live(double* out, const double* in, int iters) {
  double x[4], y[4], z[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    x[j] = in[threadIdx.x + 256 * j];
    y[j] = in[threadIdx.x + 256 * (j + 4)];
    z[j] = in[threadIdx.x + 256 * (j + 8)];
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (OP == 0) { if (r & 1) x[j] = ADD(x[j], y[j]); else y[j] = ADD(y[j], -x[j]); }        // DADD, two live sources
        else if (OP == 1) { if (r & 1) x[j] = MUL(x[j], y[j]); else y[j] = MUL(y[j], x[j]); }    // DMUL, two live sources
        else if (OP == 2) { if (r & 1) x[j] = FMA(x[j], y[j], 1e-9); else y[j] = FMA(y[j], x[j], 1e-9); }  // DFMA, two live + constant
        else { if (r % 3 == 0) x[j] = FMA(x[j], y[j], z[j]); else if (r % 3 == 1) y[j] = FMA(y[j], z[j], x[j]); else z[j] = FMA(z[j], x[j], y[j]); }
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) s += x[j] + y[j] + z[j];
  if (s == 123.456) out[0] = s;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double *in, *out;
  CK(cudaMalloc(&in, 256 * 16 * 8));
  CK(cudaMalloc(&out, 64));
  double h[256 * 16];
  for (int i = 0; i < 256 * 16; ++i) h[i] = 0.3 + 0.5 * ((i * 2654435761u) % 1000) / 1000.0;
  CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto timeit = [&](auto launch) {
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    return (double)best * 1e-3;
  };
  const int it = 4000;
  const double peak = sms * 64 * 1.965e9;  // fp64 lane-instructions / s at the maximum clock
  printf("fp64 paper ceiling %.2f T lane-instr/s\n", peak / 1e12);
  {
    const double li = (double)sms * 2 * 256.0 * 20000 * 8.0 * 4;
    const char* names[4] = {"DADD two live sources", "DMUL two live sources", "DFMA two live + constant", "DFMA three live sources"};
    double t[4];
    t[0] = timeit([&] { k_live<0><<<sms * 2, 256>>>(out, in, 20000); });
    t[1] = timeit([&] { k_live<1><<<sms * 2, 256>>>(out, in, 20000); });
    t[2] = timeit([&] { k_live<2><<<sms * 2, 256>>>(out, in, 20000); });
    t[3] = timeit([&] { k_live<3><<<sms * 2, 256>>>(out, in, 20000); });
    for (int i = 0; i < 4; ++i) printf("  %-26s %.2f T lane-instr/s (%.0f %% of paper)\n", names[i], li / t[i] / 1e12, 100.0 * li / t[i] / peak);
  }
  // fits: 9 per iteration per thread; fp64 instructions per iteration: 9 fits + 8 flux (2 each) + 8 perturbation
  auto report = [&](const char* name, double secs, int ctas, double instr_per_iter) {
    const double threads = (double)sms * ctas * 256.0;
    const double fits = threads * it * 9.0 / secs;
    printf("  %-44s %.1f G fits/s  = %.1f G cell-updates/s at 4.375 fits/cell; ~%.2f T fp64 lane-instr/s (%.0f %% of paper)\n", name,
           fits / 1e9, fits / 4.375 / 1e9, threads * it * instr_per_iter / secs / 1e12, 100.0 * threads * it * instr_per_iter / secs / peak);
  };
  const double i4 = 9 * 22.0 + 8 * 2 + 8, i6 = 9 * 47.0 + 8 * 2 + 8;
  report("order 4, 128 regs, 2 CTAs/SM (16 warps)", timeit([&] { k_fits<4, 2><<<sms * 2, 256>>>(out, in, it, 0.37); }), 2, i4);
  report("order 4, 255 regs, 1 CTA/SM (8 warps)", timeit([&] { k_fits<4, 1><<<sms * 1, 256>>>(out, in, it, 0.37); }), 1, i4);
  report("order 4, 128 regs, 4 CTAs queued per SM", timeit([&] { k_fits<4, 2><<<sms * 4, 256>>>(out, in, it, 0.37); }), 4, i4);
  {
    auto rep2 = [&](const char* name, double secs, double threads) {
      const double fits = threads * it * 9.0 / secs;
      printf("  %-60s %.1f G fits/s (%.0f %% of the unpressured stream)\n", name, fits / 1e9, 100.0 * fits / 735e9);
    };
    rep2("16 warps/SM, 128 regs, 24 doubles of live column state", timeit([&] { k_fits_pressure<256, 2, 24><<<sms * 2, 256>>>(out, in, it, 0.37); }), sms * 2 * 256.0);
    rep2("16 warps/SM, 128 regs, 32 doubles of live column state", timeit([&] { k_fits_pressure<256, 2, 32><<<sms * 2, 256>>>(out, in, it, 0.37); }), sms * 2 * 256.0);
    rep2("16 warps/SM, 128 regs, 16 doubles of live column state", timeit([&] { k_fits_pressure<256, 2, 16><<<sms * 2, 256>>>(out, in, it, 0.37); }), sms * 2 * 256.0);
    rep2("16 warps/SM, 128 regs,  8 doubles of live column state", timeit([&] { k_fits_pressure<256, 2, 8><<<sms * 2, 256>>>(out, in, it, 0.37); }), sms * 2 * 256.0);
    rep2("12 warps/SM (3 x 128 threads), 168 regs, 16 doubles live", timeit([&] { k_fits_pressure<128, 3, 16><<<sms * 3, 128>>>(out, in, it, 0.37); }), sms * 3 * 128.0);
    rep2("12 warps/SM (3 x 128 threads), 168 regs, 32 doubles live", timeit([&] { k_fits_pressure<128, 3, 32><<<sms * 3, 128>>>(out, in, it, 0.37); }), sms * 3 * 128.0);
    rep2(" 8 warps/SM, 255 regs, 32 doubles live", timeit([&] { k_fits_pressure<256, 1, 32><<<sms * 1, 256>>>(out, in, it, 0.37); }), sms * 1 * 256.0);
  }
  report("order 6, 255 regs, 1 CTA/SM (8 warps)", timeit([&] { k_fits<6, 1><<<sms * 1, 256>>>(out, in, it, 0.37); }), 1, i6);
  report("order 6, 128 regs, 2 CTAs/SM (16 warps)", timeit([&] { k_fits<6, 2><<<sms * 2, 256>>>(out, in, it, 0.37); }), 2, i6);
  return 0;
}
