// lk_fft.cu -- production Poisson solve for power-of-two grids: the r2c -> divide by the discrete symbol ->
// c2r algebra of LokiPoissonSolveFFT::solve (LokiPoissonSolveFFT.C:128-170) as three shared-memory FFT
// kernels, O(N log N) instead of the O(N^1.5) direct DFT of lk_kernels.cu (which stays the strict-mode
// path: it adds in the oracle's order).  On an 8-GPU box every rank solves the same global 512 x 1024
// problem per stage, so the solve has to stay a small fraction of a stage.
//
//   k_fft_x_fwd   one CTA per grid row b: rho(:,b) (real, contiguous) -> FFT along x -> F1[i][b]
//   k_fft_y_solve one CTA per x mode i:   F1[i][:] (contiguous) -> FFT along y -> / (sx[i] + sy[j]) unless
//                                         both vanish (LokiPoissonSolveFFT.C:150-158) -> inverse FFT -> F2[i][b]
//   k_fft_x_inv   one CTA per grid row b: F2[:][b] -> inverse FFT along x -> phi(:,b) = real part
// Symbols carry the factor nx*ny (FFTW is unnormalised, :97, :114), so the inverse passes are unnormalised too.
// Stockham autosort radix-2 passes in shared memory, twiddles from the plan's cos/sin tables.
// neutralizeCharge4D (PoissonF.f:41-61) is a deterministic two-level sum here.
#include <cuda_runtime.h>
#include <stdint.h>

namespace lkfft {

typedef long long i64;

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// in-place-style FFT of the N values in A (scratch B); returns the buffer that holds the result.
// sign = -1: X[k] = sum x[n] exp(-2 pi i nk/N); sign = +1: the unnormalised inverse.  tw[2m], tw[2m+1] =
// cos, sin(2 pi m / N).  All threads of the CTA must call it.
__device__ double2* fft_pow2(double2* A, double2* B, int N, const double* __restrict__ tw, double sign) {
  const int half = N >> 1;
  for (int p = 1; p < N; p <<= 1) {
    const int tstride = half / p;  // N / (2p)
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
      const int k = i & (p - 1);
      const double2 u0 = A[i];
      double2 u1 = A[i + half];
      const int m = k * tstride;
      const double c = tw[2 * m], s = sign * tw[2 * m + 1];
      u1 = make_double2(u1.x * c - u1.y * s, u1.x * s + u1.y * c);
      const int j = (i << 1) - k;
      B[j] = cadd(u0, u1);
      B[j + p] = csub(u0, u1);
    }
    __syncthreads();
    double2* t = A; A = B; B = t;
  }
  return A;
}

__global__ void k_fft_x_fwd(const double* __restrict__ rho, int nx, int ny, int ng, const double* __restrict__ cx,
                            double2* __restrict__ F1) {
  extern __shared__ double2 sm[];
  double2 *A = sm, *B = sm + nx;
  const int b = blockIdx.x;
  const i64 n1d = nx + 2 * ng;
  const double* row = rho + ng + n1d * (b + ng);
  for (int a = threadIdx.x; a < nx; a += blockDim.x) A[a] = make_double2(row[a], 0.0);
  __syncthreads();
  const double2* R = fft_pow2(A, B, nx, cx, -1.0);
  for (int i = threadIdx.x; i < nx; i += blockDim.x) F1[(i64)i * ny + b] = R[i];
}

__global__ void k_fft_y_solve(const double2* __restrict__ F1, int nx, int ny, const double* __restrict__ cy,
                              const double* __restrict__ sx, const double* __restrict__ sy, double2* __restrict__ F2) {
  extern __shared__ double2 sm[];
  double2 *A = sm, *B = sm + ny;
  const int i = blockIdx.x;
  const double2* src = F1 + (i64)i * ny;
  for (int b = threadIdx.x; b < ny; b += blockDim.x) A[b] = src[b];
  __syncthreads();
  double2* R = fft_pow2(A, B, ny, cy, -1.0);
  const double sxi = sx[i];
  for (int j = threadIdx.x; j < ny; j += blockDim.x) {
    const double syj = sy[(2 * j <= ny) ? j : ny - j];  // the symbol is even in the mode number
    if (sxi != 0.0 || syj != 0.0) {
      const double den = sxi + syj;
      R[j] = make_double2(R[j].x / den, R[j].y / den);
    }
  }
  __syncthreads();
  double2* other = (R == A) ? B : A;
  const double2* Q = fft_pow2(R, other, ny, cy, +1.0);
  double2* dst = F2 + (i64)i * ny;
  for (int b = threadIdx.x; b < ny; b += blockDim.x) dst[b] = Q[b];
}

__global__ void k_fft_x_inv(const double2* __restrict__ F2, int nx, int ny, int ng, const double* __restrict__ cx,
                            double* __restrict__ phi) {
  extern __shared__ double2 sm[];
  double2 *A = sm, *B = sm + nx;
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) A[i] = F2[(i64)i * ny + b];
  __syncthreads();
  const double2* R = fft_pow2(A, B, nx, cx, +1.0);
  const i64 n1d = nx + 2 * ng;
  double* row = phi + ng + n1d * (b + ng);
  for (int a = threadIdx.x; a < nx; a += blockDim.x) row[a] = R[a].x;
}

// ---- neutralizeCharge4D: rho -= mean(rho) over the interior, two-level fixed-order sum ----
constexpr int NEUT_BLOCKS = 128;
__global__ void k_neut_partial(const double* __restrict__ rho, int n1, int n2, int ng, double* __restrict__ part) {
  __shared__ double sh[32];
  const i64 n1d = n1 + 2 * ng;
  const int total = n1 * n2;
  double s = 0.0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
    s += rho[(t % n1 + ng) + n1d * (t / n1 + ng)];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) b += sh[k];
    part[blockIdx.x] = b;
  }
}
__global__ void k_neut_apply(double* __restrict__ rho, int n1, int n2, int ng, const double* __restrict__ part, int nparts) {
  __shared__ double mean;
  const i64 n1d = n1 + 2 * ng;
  const int total = n1 * n2;
  if (threadIdx.x == 0) {
    double b = 0.0;
    for (int k = 0; k < nparts; ++k) b += part[k];
    mean = b / total;
  }
  __syncthreads();
  const double m = mean;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const i64 o = (t % n1 + ng) + n1d * (t / n1 + ng);
    rho[o] = rho[o] - m;
  }
}

static bool pow2(int n) { return n >= 2 && (n & (n - 1)) == 0; }
bool supported(int nx, int ny) { return pow2(nx) && pow2(ny) && nx <= 4096 && ny <= 4096; }

// part: NEUT_BLOCKS doubles of scratch; F1, F2: nx*ny complex each
cudaError_t neutralize(double* rho, int n1, int n2, int ng, double* part, cudaStream_t st, int64_t* launches) {
  int blocks = (n1 * n2 + 255) / 256;
  if (blocks > NEUT_BLOCKS) blocks = NEUT_BLOCKS;
  k_neut_partial<<<blocks, 256, 0, st>>>(rho, n1, n2, ng, part);
  k_neut_apply<<<blocks, 256, 0, st>>>(rho, n1, n2, ng, part, blocks);
  *launches += 2;
  return cudaGetLastError();
}
int neutralize_scratch_doubles() { return NEUT_BLOCKS; }

cudaError_t poisson_fft(double* phi, const double* rho, int nx, int ny, int ng, const double* sx, const double* sy,
                        const double* cx, const double* cy, double* F1, double* F2, cudaStream_t st, int64_t* launches) {
  static bool attr_done_dev[64] = {false};   // the attribute belongs to the device's context
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  bool& attr_done = attr_done_dev[dev];
  if (!attr_done) {
    const int big = 2 * 4096 * (int)sizeof(double2);
    cudaError_t e = cudaFuncSetAttribute(k_fft_x_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fft_y_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fft_x_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  auto threads = [](int n) { int t = n / 2; return t < 32 ? 32 : (t > 512 ? 512 : t); };
  k_fft_x_fwd<<<ny, threads(nx), 2 * nx * sizeof(double2), st>>>(rho, nx, ny, ng, cx, (double2*)F1);
  k_fft_y_solve<<<nx, threads(ny), 2 * ny * sizeof(double2), st>>>((const double2*)F1, nx, ny, cy, sx, sy, (double2*)F2);
  k_fft_x_inv<<<ny, threads(nx), 2 * nx * sizeof(double2), st>>>((const double2*)F2, nx, ny, ng, cx, phi);
  *launches += 3;
  return cudaGetLastError();
}

}  // namespace lkfft
