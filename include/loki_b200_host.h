/*
 * loki_b200_host.h -- C ABI of the HOST-SIDE mirror of the reference's operator interface for the
 * Vlasov-Poisson hot path: what VPSystem / KineticSpecies / Poisson / RK4Integrator / RK6Integrator do
 * per time step (VPSystem.C:372-525, KineticSpecies.H:404-562, KineticSpecies.C:647-774,
 * EMSolverBase.C:270-371, RK4Integrator.H:66-171, RK6Integrator.H:69-133), with the distribution
 * functions resident in HBM.  The C++ classes behind it (loki_b200/csrc/lk_host.cu, namespace loki)
 * keep the reference's names and call order; they call nothing but the kernels of loki_b200.h.
 *
 * One lk_vp_system is one rank's share of the problem: configuration space (x,y) is block-decomposed
 * over the ranks, velocity space is whole on every rank.  The split-phase entry points let
 * the caller (one process per GPU, torch.distributed/NCCL for the plumbing) put the two exchanges
 * where the reference has its MPI calls: the all-gather of charge-density tiles
 * (ReductionSchedule.C:473-679) and the face halo exchange (ParallelArray.C:925-1113).
 * lk_vp_advance() runs the whole step when there is a single rank.
 */
#ifndef LOKI_B200_HOST_H
#define LOKI_B200_HOST_H
#include "loki_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* KineticSpecies construction parameters (KineticSpecies.C:120-232) */
typedef struct lk_species_desc {
  int nv[2];           /* Nvx, Nvy                                            */
  double vlo[2], vhi[2]; /* velocity-domain bounds                             */
  double mass, charge, bz_const;
  int has_driver;      /* 1: a ShapedRampedCosineDriver acts on this species  */
  double driver[16];   /* its parameter vector (ShapedRampedCosineDriver.H enum order:
                          xwidth,ywidth,shape,omega,E0,t0,trampup,thold,trampdown,xshape,lwidth,x0,alpha,tres,.,.) */
  double driver_phase; /* m_phase                                             */
  int driver_shape_type;
} lk_species_desc;

typedef struct lk_vp_desc {
  int nspecies;
  const lk_species_desc* species;
  int order;          /* spatial_solution_order 4|6   */
  int rk_order;       /* temporal_solution_order 4|6  */
  int nglobal[2];     /* global Nx, Ny                */
  double xlo[2], xhi[2];
  int tile_lo[2];     /* this rank's interior range in global cell indices: [tile_lo, tile_lo+tile_n) */
  int tile_n[2];
  int ntiles;         /* number of ranks (1 = periodic wrap done locally)  */
} lk_vp_desc;

typedef struct lk_vp_system lk_vp_system;

int lk_vp_create(lk_vp_system** sys, const lk_vp_desc* desc, void* stream);
void lk_vp_destroy(lk_vp_system* sys);
/* geometry of species s on this rank (filled from the descriptor) */
int lk_vp_species_geom(const lk_vp_system* sys, int s, lk_geom* g);

/* state I/O in the reference's restart layout: the rank's dataBox incl. ghosts, x fastest
 * (RestartWriter.C:543-560).  Synchronous. */
int lk_vp_set_state(lk_vp_system* sys, int s, const double* f_host);
int lk_vp_get_state(lk_vp_system* sys, int s, double* f_host);
/* device pointer of the current state / of the state the next evalRHS will read */
double* lk_vp_state_ptr(lk_vp_system* sys, int s);
double* lk_vp_eval_ptr(lk_vp_system* sys, int s);
/* factored initial condition used for inflow at the velocity boundaries
 * (PerturbedMaxwellianIC.C:267-289); fx: (n1d,n2d) of this rank, fv: (n3d,n4d); host pointers */
int lk_vp_set_inflow(lk_vp_system* sys, int s, const double* fx, const double* fv, double fnorm, double frac);
/* 1: refill ghosts in evalRHS exactly like the reference (default); cheap, kept for clarity */
int lk_vp_set_time(lk_vp_system* sys, double t);
double lk_vp_time(const lk_vp_system* sys);

/* VPSystem::stableDt (VPSystem.C:489-505) from the accelerations of the last evalRHS; local to this
 * rank: the caller takes the MIN over ranks (Loki_Utilities::getMinValue).  Synchronises. */
int lk_vp_stable_dt(lk_vp_system* sys, double* dt);
/* {axmax, aymax} of species s from the last evalRHS (m_lambda_max[V1], [V2]) */
int lk_vp_lambda_max(lk_vp_system* sys, int s, double out[2]);

/* whole step, single rank: copySolnData(old, state); integrator->advance (VPSystem.C:508-525) */
int lk_vp_advance(lk_vp_system* sys, double dt);

/* ---- split-phase step for ntiles > 1 (also valid for 1) ----
 * begin_step; then for stage = 0 .. nstages-1:
 *   stage_moments  -> this rank's charge-density tile (dense tile_n[0] x tile_n[1], net over species)
 *   [all-gather the tiles into the buffer lk_vp_rho_gather_ptr(): ntiles dense tiles in rank order]
 *   stage_field(tile table)    -> global neutralise + Poisson solve + E, on every rank redundantly
 *   [face halo exchange of lk_vp_eval_ptr(s) with lk_halo_pack/unpack, x then y]   (ntiles > 1)
 *   stage_finish   -> acceleration, velocity-boundary fill, fused RHS + RK update
 * end_step. */
int lk_vp_nstages(const lk_vp_system* sys);
int lk_vp_begin_step(lk_vp_system* sys, double dt);
int lk_vp_stage_moments(lk_vp_system* sys, int stage);
/* optional: let the caller own the two exchange buffers (e.g. torch tensors registered with NCCL):
 * tile = tile_n[0]*tile_n[1] doubles, gather = nglobal[0]*nglobal[1] doubles */
int lk_vp_set_comm_buffers(lk_vp_system* sys, double* rho_tile, double* rho_gather);
double* lk_vp_rho_tile_ptr(lk_vp_system* sys);
double* lk_vp_rho_gather_ptr(lk_vp_system* sys);
/* tiles: ntiles x {lo0, lo1, n0, n1} in rank order; tiles == NULL means "single rank" */
int lk_vp_stage_field(lk_vp_system* sys, int stage, const int* tiles);
int lk_vp_stage_finish(lk_vp_system* sys, int stage);
int lk_vp_end_step(lk_vp_system* sys);

/* reference-ordered, UNFUSED evaluation of one RHS (VPSystem::evalRHS) of the current state into
 * rhs_dev[s] (device, same layout); used by the parity tests.  Single rank only. */
int lk_vp_eval_rhs(lk_vp_system* sys, double** rhs_dev, double time);
/* fields of the last evalRHS: em_vars (n1d_g, n2d_g, 2) and neutralised rho (n1d_g, n2d_g); device */
const double* lk_vp_em_vars_ptr(const lk_vp_system* sys);
const double* lk_vp_rho_ptr(const lk_vp_system* sys);
/* integrated_ke_e_dot of species s (KineticSpecies.C:282-284); synchronises */
int lk_vp_ke_e_dot(lk_vp_system* sys, int s, double* value);

#ifdef __cplusplus
}
#endif
#endif
