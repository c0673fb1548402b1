// lk_launch.h -- host-callable launchers exported by each arithmetic build of lk_kernels.cu
// (namespace lkfast / lkstrict).  Included by lk_kernels.cu (definition) and lk_capi.cu (dispatch).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/loki_b200.h"

#define LK_DECLARE_LAUNCHERS(NS)                                                                       \
  namespace NS {                                                                                      \
  cudaError_t weno_fit(int order, const double* u, const double* vel, double* face, int64_t count,    \
                       cudaStream_t st);                                                              \
  cudaError_t xpby4d(double* x, const double* y, double b, const lk_geom* g, cudaStream_t st);        \
  cudaError_t max_accel(const lk_geom* g, const lk_accel* a, double* out2, double* scratch4,          \
                        cudaStream_t st);                                                             \
  cudaError_t set_phase_space_vel(double* vel3, double* vel4, const lk_geom* g, const lk_accel* a,    \
                                  double* out2, cudaStream_t st);                                     \
  cudaError_t set_accel_bcs(double* f, const lk_geom* g, const lk_accel* a, const lk_inflow* ic,      \
                            const int at[4], cudaStream_t st);                                        \
  cudaError_t periodic_fill_4d(double* f, const lk_geom* g, int px, int py, cudaStream_t st);         \
  cudaError_t halo_pack(double* buf, const double* f, const lk_geom* g, int dir, int side,            \
                        cudaStream_t st);                                                             \
  cudaError_t halo_unpack(double* f, const double* buf, const lk_geom* g, int dir, int side,          \
                          cudaStream_t st);                                                           \
  /* flags: bit0 advection terms, bit1 acceleration terms, bit2 accumulate into rhs_out */            \
  cudaError_t vlasov_rhs(double* rhs_out, const double* f, const lk_geom* g, const double* velocities,\
                         const lk_accel* a, const lk_rk_update* upd, int flags, int variant,          \
                         double* mom_part, int nmom, cudaStream_t st);                                \
  cudaError_t rk_stage_update(const double* rhs, const lk_geom* g, const lk_rk_update* upd, cudaStream_t st); \
  cudaError_t preset_inflow(double* f, const lk_geom* g, const lk_inflow* ic, cudaStream_t st);       \
  /* true when vlasov_rhs would take the pipelined kernel (which folds upd->accel_bcs into its boundary tiles) */ \
  bool stage_uses_pipe(const lk_geom* g, const lk_accel* a, const lk_rk_update* upd, double* rhs_out, int flags, int variant); \
  bool stage_folds_bcs(const lk_geom* g, const lk_accel* a, const lk_rk_update* upd, double* rhs_out, int flags, int variant); \
  int stage_moment_parts(const lk_geom* g);                                                           \
  cudaError_t moments_finish(double* d0, double* d1, double* d2, const double* part, int nparts,      \
                             int nmom, const lk_geom* g, double dv, double w, cudaStream_t st);       \
  cudaError_t ke_from_moment(double* out, const double* part1, int nparts, const lk_geom* g,          \
                             double charge, const double* ext, cudaStream_t st);                      \
  cudaError_t reduce_4d_to_2d(double* dst, const double* f, const lk_geom* g, double dv, double w,    \
                              double* scratch, int chunks, cudaStream_t st);                          \
  cudaError_t current_density(double* Jx, double* Jy, double* Jz, const double* f, const lk_geom* g,  \
                              const double* velocities, const double* vz, double dv, double w,        \
                              double* scratch, int chunks, cudaStream_t st);                          \
  cudaError_t ke_e_dot(double* out, const double* f, const lk_geom* g, double charge,                 \
                       const double* velocities, const double* ext, double* scratch, int nblocks,     \
                       cudaStream_t st);                                                              \
  cudaError_t neutralize(double* rho, int n1, int n2, int ng, cudaStream_t st);                       \
  cudaError_t poisson_dft(double* phi, const double* rho, int nx, int ny, int ng, const double* sx,   \
                          const double* sy, const double* cx, const double* cy, double* T, double* X, \
                          cudaStream_t st);                                                           \
  cudaError_t efield_from_phi(double* em, const double* phi, int n1, int n2, int ng, int order,       \
                              double dx, double dy, cudaStream_t st);                                 \
  cudaError_t periodic_fill_2d(double* u, int n1, int n2, int ng, int ncomp, int px, int py,          \
                               cudaStream_t st);                                                      \
  cudaError_t xpby2d(double* x, const double* y, double b, int n1, int n2, int ng, int ncomp,         \
                     cudaStream_t st);                                                                \
  cudaError_t form_accel(double* accel, const double* em, const double* ext, double norm, int n1,     \
                         int n2, int ng, cudaStream_t st);                                            \
  cudaError_t maxwell_rhs(double* rhs, const double* em, const double* Jx, const double* Jy,          \
                          const double* Jz, int n1, int n2, int ng, int order, double dx, double dy,  \
                          double c, double av_weak, double av_strong, cudaStream_t st);               \
  int64_t launches();                                                                                 \
  int64_t pipe_launches();                                                                            \
  }

LK_DECLARE_LAUNCHERS(lkfast)
LK_DECLARE_LAUNCHERS(lkstrict)
