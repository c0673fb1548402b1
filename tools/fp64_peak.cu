// fp64_peak.cu -- measures the fp64 (non-tensor) issue ceiling of this B200 that the fused Vlasov stage
// kernel is co-limited by (VERDICT r1 "the fp64 ceiling is derived, not measured"):
//   * DFMA / DADD / DMUL throughput at 16 and 32 warps per SM with 1..8 independent chains per thread,
//   * the dependent-issue latency of one DFMA chain (cycles),
//   * the SUSTAINED rate under the board power cap (>= 2 s back to back), first launch vs last second,
//   * the accuracy of the MUFU.RCP64H seed (rcp.approx.ftz.f64) that fast_rcp() refines.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
// run:   tools/fp64_peak [seconds]      (prints one JSON object on the last line)
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

// OP: 0 = DFMA, 1 = DADD, 2 = DMUL, 3 = the stage kernel's mix (DFMA : DADD : DMUL = 24 : 21 : 15)
template <int ILP, int OP>
__global__ void __launch_bounds__(256) k_chain(double* out, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = a + threadIdx.x * 1e-9 + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int k = 0; k < ILP; ++k) {
        if (OP == 0) x[k] = __fma_rn(x[k], a, b);
        else if (OP == 1) x[k] = __dadd_rn(x[k], b);
        else if (OP == 2) x[k] = __dmul_rn(x[k], a);
        else {
          if (r % 5 < 2) x[k] = __fma_rn(x[k], a, b);
          else if (r % 5 < 4) x[k] = __dadd_rn(x[k], b);
          else x[k] = __dmul_rn(x[k], a);
        }
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  if (s == 123.456) out[0] = s;
}

// register-file pressure: DFMA whose three sources are three different (non-constant) register pairs,
// and DFMA streams with NI independent integer instructions per DFMA (co-issue in the off cycle)
template <int ILP>
__global__ void __launch_bounds__(256) k_dfma3(double* out, const double* in, int iters) {
  double x[ILP], y[ILP], z[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) {
    x[k] = in[threadIdx.x + 256 * k];
    y[k] = in[threadIdx.x + 256 * (k + ILP)];
    z[k] = in[threadIdx.x + 256 * (k + 2 * ILP)];
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int k = 0; k < ILP; ++k) {
        // rotate roles so that no operand is loop invariant: every source is a live, changing register pair
        if (r % 3 == 0) x[k] = __fma_rn(x[k], y[k], z[k]);
        else if (r % 3 == 1) y[k] = __fma_rn(y[k], z[k], x[k]);
        else z[k] = __fma_rn(z[k], x[k], y[k]);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k] + y[k] + z[k];
  if (s == 123.456) out[0] = s;
}
template <int ILP, int NI>
__global__ void __launch_bounds__(256) k_mixint(double* out, double a, double b, int iters, int seed) {
  double x[ILP];
  int j[ILP * (NI > 0 ? NI : 1)];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = a + threadIdx.x * 1e-9 + k;
#pragma unroll
  for (int k = 0; k < ILP * (NI > 0 ? NI : 1); ++k) j[k] = seed + threadIdx.x + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int k = 0; k < ILP; ++k) {
        x[k] = __fma_rn(x[k], a, b);
#pragma unroll
        for (int n = 0; n < NI; ++n) j[k * NI + n] = (j[k * NI + n] ^ (j[k * NI + n] >> 3)) + seed;  // SHF + LOP3/IADD3
      }
    }
  }
  double s = 0.0;
  int t = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
#pragma unroll
  for (int k = 0; k < ILP * (NI > 0 ? NI : 1); ++k) t += j[k];
  if (s == 123.456 || t == 123456789) out[0] = s + t;
}

__global__ void k_latency(double* out, long long* cyc, double a, double b, int iters) {
  double x = a;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r) x = __fma_rn(x, a, b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) {
    cyc[0] = t1 - t0;
    out[0] = x;
  }
}

__global__ void k_rcp_err(double* maxerr, unsigned long long seed, int per_thread) {
  unsigned long long s = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  double worst = 0.0;
  for (int i = 0; i < per_thread; ++i) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    // mantissa random, exponent in [-60, 60]
    const double m = 1.0 + (double)(s >> 12) * (1.0 / 4503599627370496.0);
    const int e = (int)((s >> 3) % 121) - 60;
    const double x = ldexp(m, e);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double err = fabs(__fma_rn(-x, r, 1.0));
    worst = fmax(worst, err);
  }
  // max over the grid (positive doubles order like their bit patterns)
  atomicMax((unsigned long long*)maxerr, (unsigned long long)__double_as_longlong(worst));
}

static int g_sms = 0;
static double g_extra[3] = {0, 0, 0};
static double g_clock_ghz_max = 0;

template <int ILP, int OP>
static double rate(int ctas_per_sm, int iters, double* out, int reps = 3) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k_chain<ILP, OP><<<g_sms * ctas_per_sm, 256>>>(out, 1.0000001, 1e-9, iters);
  CK(cudaDeviceSynchronize());
  double best = 0;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    k_chain<ILP, OP><<<g_sms * ctas_per_sm, 256>>>(out, 1.0000001, 1e-9, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double lane_instr = (double)g_sms * ctas_per_sm * 256.0 * iters * 8.0 * ILP;
    const double t = lane_instr / (ms * 1e-3) / 1e12;
    if (t > best) best = t;
  }
  return best;  // T lane-instructions / s
}

int main(int argc, char** argv) {
  const double seconds = (argc > 1) ? atof(argv[1]) : 3.0;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  g_sms = p.multiProcessorCount;
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  g_clock_ghz_max = khz * 1e-6;
  double* out;
  long long* cyc;
  CK(cudaMalloc(&out, 64));
  CK(cudaMalloc(&cyc, 64));
  CK(cudaMemset(out, 0, 64));
  printf("device %s, %d SMs, max clock %.3f GHz; paper ceiling %d x 64 lanes x clock = %.2f T lane-instr/s\n", p.name,
         g_sms, g_clock_ghz_max, g_sms, g_sms * 64 * g_clock_ghz_max * 1e-3);

  // ---- latency of a dependent DFMA chain ----
  k_latency<<<1, 32>>>(out, cyc, 1.0000001, 1e-9, 4096);
  CK(cudaDeviceSynchronize());
  long long c;
  CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
  const double lat = (double)c / (4096.0 * 16.0);
  printf("DFMA dependent-issue latency: %.2f cycles\n", lat);

  // ---- throughput table (burst: each launch ~ 10-30 ms) ----
  const int it = 20000;
  printf("throughput, T lane-instr/s (burst):            ILP1    ILP2    ILP4    ILP8\n");
  double dfma16[4], dfma32[4], dfma8[4];
  dfma8[0] = rate<1, 0>(1, it, out); dfma8[1] = rate<2, 0>(1, it, out); dfma8[2] = rate<4, 0>(1, it, out); dfma8[3] = rate<8, 0>(1, it / 2, out);
  printf("  DFMA,  8 warps/SM (2 per scheduler):        %6.2f  %6.2f  %6.2f  %6.2f\n", dfma8[0], dfma8[1], dfma8[2], dfma8[3]);
  dfma16[0] = rate<1, 0>(2, it, out); dfma16[1] = rate<2, 0>(2, it, out); dfma16[2] = rate<4, 0>(2, it, out); dfma16[3] = rate<8, 0>(2, it / 2, out);
  printf("  DFMA, 16 warps/SM (4 per scheduler):        %6.2f  %6.2f  %6.2f  %6.2f\n", dfma16[0], dfma16[1], dfma16[2], dfma16[3]);
  dfma32[0] = rate<1, 0>(4, it, out); dfma32[1] = rate<2, 0>(4, it, out); dfma32[2] = rate<4, 0>(4, it / 2, out); dfma32[3] = rate<8, 0>(4, it / 4, out);
  printf("  DFMA, 32 warps/SM (8 per scheduler):        %6.2f  %6.2f  %6.2f  %6.2f\n", dfma32[0], dfma32[1], dfma32[2], dfma32[3]);
  const double dadd = rate<4, 1>(2, it, out), dmul = rate<4, 2>(2, it, out), mix = rate<4, 3>(2, it, out);
  printf("  16 warps/SM, ILP4: DADD %.2f  DMUL %.2f  stage mix (2 DFMA : 2 DADD : 1 DMUL) %.2f\n", dadd, dmul, mix);

  // ---- register-operand and co-issue sensitivity ----
  {
    double* in;
    CK(cudaMalloc(&in, 256 * 12 * 8));
    std::vector<double> h(256 * 12);
    for (size_t i = 0; i < h.size(); ++i) h[i] = 1.0 + 1e-9 * (double)(i % 97);
    CK(cudaMemcpy(in, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    cudaEvent_t a0, a1;
    CK(cudaEventCreate(&a0));
    CK(cudaEventCreate(&a1));
    auto timeit = [&](auto launch, double lane_instr) {
      launch();
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a0));
      launch();
      CK(cudaEventRecord(a1));
      CK(cudaEventSynchronize(a1));
      float ms;
      CK(cudaEventElapsedTime(&ms, a0, a1));
      return lane_instr / (ms * 1e-3) / 1e12;
    };
    const double li = (double)g_sms * 2 * 256.0 * it * 8.0 * 4;
    const double r3 = timeit([&] { k_dfma3<4><<<g_sms * 2, 256>>>(out, in, it); }, li);
    printf("DFMA with three distinct live register sources (16 warps/SM, ILP4): %.2f T lane-instr/s\n", r3);
    const double i0 = timeit([&] { k_mixint<4, 0><<<g_sms * 2, 256>>>(out, 1.0000001, 1e-9, it, 7); }, li);
    const double i1 = timeit([&] { k_mixint<4, 1><<<g_sms * 2, 256>>>(out, 1.0000001, 1e-9, it, 7); }, li);
    const double i2 = timeit([&] { k_mixint<4, 2><<<g_sms * 2, 256>>>(out, 1.0000001, 1e-9, it, 7); }, li);
    printf("DFMA rate with 0 / ~2 / ~4 independent integer instructions per DFMA: %.2f / %.2f / %.2f T lane-instr/s\n", i0, i1, i2);
    g_extra[0] = r3; g_extra[1] = i1; g_extra[2] = i2;
  }

  // ---- sustained under the power cap ----
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<float> ms;
  double elapsed = 0;
  const int its = 40000;
  const double lane_instr = (double)g_sms * 2 * 256.0 * its * 8.0 * 4;
  while (elapsed < seconds) {
    CK(cudaEventRecord(e0));
    k_chain<4, 0><<<g_sms * 2, 256>>>(out, 1.0000001, 1e-9, its);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float t;
    CK(cudaEventElapsedTime(&t, e0, e1));
    ms.push_back(t);
    elapsed += t * 1e-3;
  }
  double tail = 0, tail_t = 0;
  int ntail = 0;
  for (int i = (int)ms.size() - 1; i >= 0 && tail_t < 1000.0; --i) { tail_t += ms[i]; ++ntail; }
  tail = lane_instr * ntail / (tail_t * 1e-3) / 1e12;
  const double first = lane_instr / (ms[0] * 1e-3) / 1e12;
  printf("sustained DFMA (16 warps/SM, ILP4), %.1f s back to back: first launch %.2f, last second %.2f T lane-instr/s\n",
         elapsed, first, tail);
  // the mixed stream, sustained
  ms.clear();
  elapsed = 0;
  while (elapsed < seconds) {
    CK(cudaEventRecord(e0));
    k_chain<4, 3><<<g_sms * 2, 256>>>(out, 1.0000001, 1e-9, its);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float t;
    CK(cudaEventElapsedTime(&t, e0, e1));
    ms.push_back(t);
    elapsed += t * 1e-3;
  }
  tail_t = 0; ntail = 0;
  for (int i = (int)ms.size() - 1; i >= 0 && tail_t < 1000.0; --i) { tail_t += ms[i]; ++ntail; }
  const double tail_mix = lane_instr * ntail / (tail_t * 1e-3) / 1e12;
  printf("sustained stage mix, last second %.2f T lane-instr/s\n", tail_mix);

  // ---- accuracy of the reciprocal seed ----
  double* me;
  CK(cudaMalloc(&me, 8));
  CK(cudaMemset(me, 0, 8));
  k_rcp_err<<<g_sms * 8, 256>>>(me, 12345ull, 4096);
  CK(cudaDeviceSynchronize());
  double worst;
  CK(cudaMemcpy(&worst, me, 8, cudaMemcpyDeviceToHost));
  printf("rcp.approx.ftz.f64 seed: max |1 - x*r| = %.3e (2^%.2f); after one Newton step %.2e, after the cubic step %.2e\n",
         worst, log2(worst), worst * worst, worst * worst * worst);

  printf("{\"fp64_peak\": {\"unit\": \"T lane-instr/s\", \"paper_at_max_clock\": %.3f, \"burst\": %.3f, \"sustained\": %.3f, "
         "\"sustained_stage_mix\": %.3f, \"dfma_latency_cycles\": %.2f, \"sms\": %d, \"max_clock_ghz\": %.3f, "
         "\"rcp_seed_max_rel_err\": %.4e, \"dfma_three_register_sources\": %.3f, \"dfma_with_2_int_per_dfma\": %.3f, \"dfma_with_4_int_per_dfma\": %.3f, \"how\": \"tools/fp64_peak.cu: DFMA chains, 16 warps/SM x 4 chains/thread, CUDA events; "
         "sustained = last second of %.0f s back to back\"}}\n",
         g_sms * 64 * g_clock_ghz_max * 1e-3, dfma16[2] > dfma32[2] ? dfma16[2] : dfma32[2], tail, tail_mix, lat, g_sms,
         g_clock_ghz_max, worst, g_extra[0], g_extra[1], g_extra[2], seconds);
  return 0;
}
