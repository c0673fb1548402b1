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
/* The same without stalling the device: a state can be on its way to the host while the next one arrives
 * (full-duplex PCIe) -- the three rotating arrays of a species give the double buffer, nothing is copied twice.
 *   lk_vp_download_state(sys, s, pinned, stream)  queue D2H of the current state on `stream`, behind the system's stream
 *   lk_vp_upload_next(sys, s, pinned, stream)     queue H2D of the NEXT state on `stream` into the array the previous
 *                                                 state occupied (free once the step the system's stream holds is done)
 *   lk_vp_adopt_next(sys)                         the uploaded arrays become the states (the system's stream waits for
 *                                                 the uploads); a later step waits for a pending download before it
 *                                                 reuses the array that download reads
 * The host buffers must stay valid until their stream has been synchronised. */
int lk_vp_download_state(lk_vp_system* sys, int s, double* f_host_pinned, void* stream);
int lk_vp_upload_next(lk_vp_system* sys, int s, const double* f_host_pinned, void* stream);
int lk_vp_adopt_next(lk_vp_system* sys);
/* device pointer of the current state / of the state the next evalRHS will read */
double* lk_vp_state_ptr(lk_vp_system* sys, int s);
double* lk_vp_eval_ptr(lk_vp_system* sys, int s);
/* factored initial condition used for inflow at the velocity boundaries
 * (PerturbedMaxwellianIC.C:267-289); fx: (n1d,n2d) of this rank, fv: (n3d,n4d); host pointers */
int lk_vp_set_inflow(lk_vp_system* sys, int s, const double* fx, const double* fv, double fnorm, double frac);
/* the two factored forms of InterpenetratingStreamIC::getIC_At_Pt (InterpenetratingStreamIC.C:265-286):
 * kind 2 (two-sided) fx*fv + fx2*fv2, kind 4 (centred) fv*fx*fx2; fnorm is folded into fv / fv2 as the
 * reference's cache does (:240-252) */
int lk_vp_set_inflow2(lk_vp_system* sys, int s, int kind, const double* fx, const double* fv, const double* fx2,
                      const double* fv2);
/* Deck options beyond the benchmark decks.
 * lk_vp_set_boundary_options: a non-periodic x / y direction gets setadvectionbcs4d_ at its two physical boundaries
 * before every advection sweep (setPhysicalBCs, KineticSpecies.H:998-1031; inflow from the species' tables; the Poisson
 * solve stays periodic as in Poisson.C:147-152); use_new_bcs selects the "JB" variants of both boundary fills
 * (VPSystem.C:819-821, KineticSpecies.H:421-453).  Non-periodic directions need a single rank.
 * lk_vp_set_krook: nu (n1d,n2d) of this rank incl. ghosts, host pointer (KrookLayer::initialize, KrookLayer.C:54-160);
 * completeRHS then adds -nu/dt (f - f_IC) to the species' rhs (KineticSpecies.C:1049-1062; f_IC from the inflow
 * tables).  NULL removes the layer.  The term is applied inside the fused stage (generic kernel's per-cell epilogue,
 * lk_rk_update.krook_*); a Krook species does not take the pipelined instantiation. */
int lk_vp_set_boundary_options(lk_vp_system* sys, int nonperiodic_x, int nonperiodic_y, int use_new_bcs);
int lk_vp_set_krook(lk_vp_system* sys, int s, const double* nu_host);
/* lk_vp_set_pitch_angle: a "Pitch Angle Collision Operator" on species s (KineticSpecies.C:1036-1046): every stage
 * materialises the species' rhs, adds C(f) (lk_pitch_angle_fields + lk_append_pitch_angle_collision), then the Krook
 * layer, then does the Runge-Kutta update (lk_rk_stage_update); lk_vp_stable_dt takes the operator's real eigenvalue
 * into the step estimate (KineticSpecies.C:666-672).  NULL removes the operator.  Returns LK_ERR_ARG for a range the
 * reference aborts on. */
int lk_vp_set_pitch_angle(lk_vp_system* sys, int s, const lk_pitch_angle* p);
int lk_vp_set_time(lk_vp_system* sys, double t);
double lk_vp_time(const lk_vp_system* sys);

/* VPSystem::stableDt (VPSystem.C:489-505) from the accelerations of the last evalRHS.  With configuration
 * space cut over several ranks the caller first makes the velocity-space maxima global, component by
 * component, as KineticSpecies::computeDt does with its MPI_Allreduce(MAX) over m_lambda_max
 * (KineticSpecies.C:650-656): lk_vp_lambda_max on every rank -> MAX over ranks -> lk_vp_set_lambda_max ->
 * lk_vp_stable_dt (loki_b200/decomp.py::DistributedVP.stable_dt).  Synchronises. */
int lk_vp_stable_dt(lk_vp_system* sys, double* dt);
/* {axmax, aymax} of species s from the last evalRHS (m_lambda_max[V1], [V2]); local to this rank's tile */
int lk_vp_lambda_max(lk_vp_system* sys, int s, double out[2]);
/* overwrite them with the maxima over all ranks (valid until the next evalRHS) */
int lk_vp_set_lambda_max(lk_vp_system* sys, int s, const double in[2]);

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
/* periodic wrap of lk_vp_eval_ptr(s) inside this rank for a direction that is not cut (0: x, 1: y); a
 * no-op when the fused stage kernel has already written those ghost cells (lk_rk_update::wrap) */
int lk_vp_local_fill(lk_vp_system* sys, int s, int dir);
/* 1 when that call would launch anything (lets the caller skip the stream ordering around it) */
int lk_vp_local_fill_needed(lk_vp_system* sys, int s, int dir);
int lk_vp_stage_finish(lk_vp_system* sys, int stage);
/* the same for one species: lets the caller start the halo exchange of species s's new predictor
 * (lk_vp_eval_ptr(s) after this call) on another stream while the next species' stage kernel runs */
int lk_vp_stage_finish_species(lk_vp_system* sys, int stage, int s);
/* ... and in two launches, so that a species' own halo exchange overlaps its own stage kernel: part 1 runs the stage
 * with the kernel restricted to the CTA tiles on a face of a cut direction (everything a neighbour needs of the new
 * predictor; lk_vp_eval_ptr(s) is the new predictor from here on), part 2 launches the remaining tiles.  Between the two
 * the caller queues the exchange on another stream behind lk_vp_wait_faces.  When the stage cannot be
 * split (strict arithmetic, unaligned tiles, Krook species: not the pipelined kernel) part 1 is the whole stage and
 * part 2 is a no-op: the caller's sequence stays the same. */
int lk_vp_stage_finish_species_part(lk_vp_system* sys, int stage, int s, int part);
/* The face tiles of part 1 run on a stream of their own (high priority), the remaining tiles on the system's stream
 * without waiting for them: no idle tail between the two launches.  lk_vp_wait_faces makes `stream` (the caller's
 * exchange stream) wait for the face tiles of species s -- or, when the stage was not split, for what the system's
 * stream holds now.  Part 2 makes the system's stream wait for the face tiles as well. */
int lk_vp_wait_faces(lk_vp_system* sys, int s, void* stream);
int lk_vp_end_step(lk_vp_system* sys);

/* reference-ordered, UNFUSED evaluation of one RHS (VPSystem::evalRHS) of the current state into
 * rhs_dev[s] (device, same layout); used by the parity tests.  Single rank only. */
int lk_vp_eval_rhs(lk_vp_system* sys, double** rhs_dev, double time);
/* fields of the last evalRHS: em_vars (n1d_g, n2d_g, 2) and neutralised rho (n1d_g, n2d_g); device */
const double* lk_vp_em_vars_ptr(const lk_vp_system* sys);
const double* lk_vp_rho_ptr(const lk_vp_system* sys);
/* VPSystem::accumulateSequences (VPSystem.C:591-636) without probes, particles and flux histories:
 * out = {e_max, e_tot, ex_max, ey_max, e_sum_tot} of the field of the last evalRHS (Poisson.C:796-860), then
 * per species {ke, ke_x, ke_y, px, py} (computeke_) and the integrated driver work: 5 + 6*nspecies values,
 * *written receives the count.  Local to this rank: sums / maxima over ranks are the caller's
 * (Loki_Utilities::getSum / getMaxValue).  Synchronises. */
int lk_vp_time_history(lk_vp_system* sys, double* out, int capacity, int* written);
/* The probe histories of Poisson::accumulateSequences (Poisson.C:852-887): out[2k], out[2k+1] = Ex, Ey of the last
 * evalRHS at the cell floor(frac_x[k] * Nx), floor(frac_y[k] * Ny) (global cell indices; Simulation.C:393-412 reads
 * `number_of_probes` / `probe.N.location`, default one probe at (0.5, 0): the reference assigns m_probes[X1][0] twice).
 * A rank reports the probes inside its own tile, 0 for the others: add over ranks.  Synchronises. */
int lk_vp_probe_history(lk_vp_system* sys, int nprobes, const double* frac_x, const double* frac_y, double* out);
/* The flux histories of KineticSpecies::accumulateSequencesCommon (KineticSpecies.C:2052-2097): per species the
 * kinetic-energy flux through the eight phase-space boundaries, out[8 s + 2 dir + side] (dir 0..3 = x, y, vx, vy; side
 * 0 = low), 8 * nspecies values.  Uses the face accelerations of the last evalRHS, as the reference does; rewrites the
 * state's ghost cells (periodic wrap in the directions this rank is not cut in, velocity-boundary fill).  A rank only
 * sums the boundaries its tile touches: add over ranks (Loki_Utilities::getSum).  In a cut direction the caller
 * exchanges the halos of lk_vp_state_ptr first.  Synchronises. */
int lk_vp_flux_history(lk_vp_system* sys, double* out, int capacity, int* written);
/* integrated_ke_e_dot of species s (KineticSpecies.C:282-284); synchronises */
int lk_vp_ke_e_dot(lk_vp_system* sys, int s, double* value);
/* kinetic_species.N.tz.* (TZSourceFactory.C:22-56, KineticSpecies.C:186, :1077-1080): on = 0 none, 1 "TrigTZSource",
 * 2 "ElectronTrigTZSource", 3 "TwoSpecies_ElectronTrigTZSource", 4 "TwoSpecies_IonTrigTZSource" (tz.amp; the two-species
 * sources also tz.electron_mass and tz.ion_mass) adds the twilight-zone source (lk_set_trig_tz_source, kind = on - 1) to
 * species s's right-hand side in completeRHS, after the collision operator and the Krook layer.  Such a species takes the
 * three-pass stage (rhs materialised). */
int lk_vp_set_trig_tz(lk_vp_system* sys, int s, int on, double amp, double electron_mass, double ion_mass);
/* TrigTZSource::computeError (TrigTZSource.C:63-82): error_host (the state's extents, host) = state - f_exact(time) over
 * the whole data box; what a restart dump holds in place of the distribution (KineticSpecies.C:987-1004).  Synchronises. */
int lk_vp_trig_tz_error(lk_vp_system* sys, int s, double time, double* error_host);
/* VPSystem::updateGhosts (VPSystem.C:779-797) before a restart dump (Simulation.C:148-151): the x / y ghost layers of
 * every species' state on this rank -- boundary conditions of a non-periodic direction, periodic wrap of the
 * directions this rank is not cut in (a cut direction's halos are the caller's exchange) */
int lk_vp_update_ghosts(lk_vp_system* sys);
/* restore integrated_ke_e_dot of species s from a restart dump (KineticSpecies.C:925); synchronises */
int lk_vp_set_ke_e_dot(lk_vp_system* sys, int s, double value);
/* The driver histories of KineticSpecies::accumulateSequencesCommon (KineticSpecies.C:2099-2150): *ke_e_dot =
 * computekeedot_ of the current state against the external driver evaluated at `time` (this rank's part: add over
 * ranks), *envel = the driver's time envelope (ShapedRampedCosineDriver::evaluateTimeEnvelope); both 0 for a species
 * without a driver.  Overwrites the species' m_ext_efield as the reference does.  Synchronises. */
int lk_vp_driver_history(lk_vp_system* sys, int s, double time, double* ke_e_dot, double* envel);

/* ---------------------------------------------------------------------------------------------
 * Vlasov-Maxwell: the host mirror of VMSystem / VMState / Maxwell (VMSystem.C:407-581, Maxwell.C:299-353,
 * 562-623, Maxwell.H:199-204, 371-381) driven by RK4Integrator (RK4Integrator.H:66-171) or RK6Integrator
 * (RK6Integrator.H:69-133).  The RK state is
 * the distribution functions plus em_vars (n1d,n2d,6: Ex,Ey,Ez,Bx,By,Bz) and one transverse drift
 * velocity vz (n1d,n2d) per species; x and y periodic; no E-field drivers, antennae or particles.
 * Runs on one GPU (base.ntiles must be 1 and the tile must be the whole configuration space).
 * --------------------------------------------------------------------------------------------- */
typedef struct lk_vm_desc {
  lk_vp_desc base;      /* rk_order 4 (RK4Integrator) or 6 (RK6Integrator) */
  double light_speed;   /* Simulation::s_LIGHT_SPEED */
  double av_weak, av_strong; /* maxwell.avWeak / avStrong (MaxwellF.f:298-352) */
} lk_vm_desc;
typedef struct lk_vm_system lk_vm_system;

int lk_vm_create(lk_vm_system** sys, const lk_vm_desc* desc, void* stream);
void lk_vm_destroy(lk_vm_system* sys);
int lk_vm_species_geom(const lk_vm_system* sys, int s, lk_geom* g);
/* restart-layout state I/O (dataBox incl. ghosts, x fastest); synchronous */
int lk_vm_set_state(lk_vm_system* sys, int s, const double* f_host);
int lk_vm_get_state(lk_vm_system* sys, int s, double* f_host);
double* lk_vm_state_ptr(lk_vm_system* sys, int s);
int lk_vm_set_fields(lk_vm_system* sys, const double* em_host); /* (n1d,n2d,6) */
int lk_vm_get_fields(lk_vm_system* sys, double* em_host);
int lk_vm_set_vz(lk_vm_system* sys, int s, const double* vz_host); /* (n1d,n2d) */
int lk_vm_get_vz(lk_vm_system* sys, int s, double* vz_host);
const double* lk_vm_fields_ptr(lk_vm_system* sys);              /* device em_vars of the current state */
const double* lk_vm_current_ptr(lk_vm_system* sys, int comp);   /* device net Jx/Jy/Jz of the last evalRHS */
/* inflow values at the velocity boundaries: factored tables, or -- for initial conditions that do not
 * factor (PerturbedMaxwellianIC.C:176-246 with a flow-velocity wave, the emDamping deck) -- the
 * velocity-ghost layers of the cached IC array: ghost3 (n1d,n2d,2*ng,n4d), ghost4 (n1d,n2d,n3d,2*ng),
 * layers [0,ng) below the interior, [ng,2ng) above; host pointers */
int lk_vm_set_inflow(lk_vm_system* sys, int s, const double* fx, const double* fv, double fnorm, double frac);
int lk_vm_set_inflow_ghosts(lk_vm_system* sys, int s, const double* ghost3, const double* ghost4);
int lk_vm_set_time(lk_vm_system* sys, double t);
double lk_vm_time(const lk_vm_system* sys);
/* VMSystem::advance (copySolnData(old, state); RK4 over the whole VMState) */
int lk_vm_advance(lk_vm_system* sys, double dt);
/* VMSystem::stableDt: min over species of computeDt and Maxwell::computeDt = 1/(c(1/dx+1/dy)) */
int lk_vm_stable_dt(lk_vm_system* sys, double* dt);
int lk_vm_lambda_max(lk_vm_system* sys, int s, double out[2]);
/* Maxwell::accumulateSequences (Maxwell.C:753-875) field histories {e_max, e_tot, ex_max, ey_max, ez_max,
 * e_sum_tot, b_max, b_tot, bx_max, by_max, bz_max, b_sum_tot} of the current em_vars, then per species
 * {ke, ke_x, ke_y, 0, 0} (computekemaxwell_): 12 + 5*nspecies values, *written receives the count.
 * Synchronises. */
int lk_vm_time_history(lk_vm_system* sys, double* out, int capacity, int* written);
/* VMSystem::evalRHS of the current state in the reference's UNFUSED order (parity hook): rhs_dev[s] 4D,
 * rhs_em_dev (n1d,n2d,6), rhs_vz_dev[s] (n1d,n2d); device pointers, interior written */
int lk_vm_eval_rhs(lk_vm_system* sys, double** rhs_dev, double* rhs_em_dev, double** rhs_vz_dev, double time);

#ifdef __cplusplus
}
#endif
#endif
