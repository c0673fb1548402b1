/*
 * loki_oracle.h -- CPU ORACLE for the Vlasov right-hand-side hot path of LLNL/LOKI.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (loki_b200/) never
 * links, imports or calls anything in oracle/.
 *
 * It is a hand-written C restatement of the *live* Fortran-77 arithmetic of the reference (the
 * reference cannot be built here: no Fortran compiler, MPI, FFTW or HDF5).  Every function cites the
 * reference file:line it follows; loop order and operation order are those of the reference so the
 * result is the one a `gfortran -O2` (no FMA contraction, configure.in:219-232) build produces.
 * Compile with `gcc -O2 -ffp-contract=off`.
 *
 * PINNING: the restatement is pinned against the reference's own Fortran source, mechanically
 * transliterated to C by oracle/f77toc.py into oracle/_ref/ (see oracle/Makefile, tests/test_oracle_pin.py).
 * The reference ships no golden outputs (checkTests.C:423 reads baselines from outside the repo).
 *
 * Index conventions: all arrays are the reference's Fortran (column-major, first index fastest)
 * layouts; indices here are 0-based offsets into the *data box* (interior grown by ng ghosts), i.e.
 * reference index i (global, interior starting at n?a) maps to (i - nd?a).
 */
#ifndef LOKI_ORACLE_H
#define LOKI_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ok_geom {
  int n[4];      /* interior cells (Nx, Ny, Nvx, Nvy) of this box                        */
  int ng;        /* ghost width: 2 (order 4) or 3 (order 6), KineticSpecies.C:155-160    */
  int order;     /* spatial_solution_order: 4 or 6                                       */
  double dx[4];  /* cell sizes (dx, dy, dvx, dvy), ProblemDomain.C:20-27                 */
} ok_geom;

/* data-box extents and the linear offset of (i1,i2,i3,i4), ParallelArray.H:1121-1132 */
static inline int64_t ok_nd(const ok_geom* g, int d) { return (int64_t)g->n[d] + 2 * g->ng; }
static inline int64_t ok_idx(const ok_geom* g, int i1, int i2, int i3, int i4) {
  return (((int64_t)i4 * ok_nd(g, 2) + i3) * ok_nd(g, 1) + i2) * ok_nd(g, 0) + i1;
}

/* callback standing in for initialconditionatpoint_ (ICInterface.C:36-57); data-box indices */
typedef double (*ok_ic_fn)(void* ctx, int i1, int i2, int i3, int i4);
/* the IC classes' cached tables as a C point callback (what initialconditionatpoint_ evaluates, ICInterface.C:36-57):
 * kind 1 fnorm*fv*fx*frac (PerturbedMaxwellianIC.C:279-281), 2 fx*fv + fx2*fv2 and 4 fv*fx*fx2
 * (InterpenetratingStreamIC.C:275-281), 0 fv*fx, 3 the cached full array m_f (PerturbedMaxwellianIC.C:176-246) */
typedef struct ok_ic_tables {
  int kind, n1d, n2d, n3d, n4d;
  const double *fx, *fv, *fx2, *fv2, *full;
  double fnorm, frac;
} ok_ic_tables;
double ok_ic_from_tables(void* ctx, int i1, int i2, int i3, int i4);

/* ---- KineticSpeciesF.f ---- */
double ok_weno43_fit(double um2, double um1, double u0, double up1, double vel);
double ok_weno65_fit(double um3, double um2, double um1, double u0, double up1, double up2, double vel);
void ok_weno43_fit_v(const double* u4, const double* vel, double* face, int64_t count); /* batch of 4-tuples */
void ok_weno65_fit_v(const double* u6, const double* vel, double* face, int64_t count);

void ok_xpby4d(double* x, const double* y, double b, const ok_geom* g);

/* vel3: (i3,i4,i1,i2) ext (n3d+1,n4d,n1d,n2d); vel4: (i4,i1,i2,i3) ext (n4d+1,n1d,n2d,n3d)
 * vxface_vel: (n3d+1,n4d,2); vyface_vel: (n3d,n4d+1,2); accel: (n1d,n2d,2) */
void ok_set_phase_space_vel_4d(double* vel3, double* vel4, const ok_geom* g, const double* vxface_vel,
                               const double* vyface_vel, double normalization, double bz_const,
                               const double* accel, double* axmax, double* aymax);
/* em_vars: (n1d,n2d,6) Ex,Ey,Ez,Bx,By,Bz; vz: (n1d,n2d) */
void ok_set_phase_space_vel_maxwell_4d(double* vel3, double* vel4, const ok_geom* g,
                                       const double* vxface_vel, const double* vyface_vel,
                                       double normalization, double bz_const, const double* em_vars,
                                       const double* vz, double* axmax, double* aymax);

/* at_*: does this box touch the global lower/upper velocity boundary (ng3a+nghosts==n3a etc.) */
void ok_set_acceleration_bcs_4d(double* u, const ok_geom* g, const double* vel3, const double* vel4,
                                int at_lo3, int at_hi3, int at_lo4, int at_hi4, ok_ic_fn ic,
                                void* ic_ctx);

/* setAdvectionBCs4D (KineticSpeciesF.f:1166-1297): non-periodic x / y physical boundaries */
void ok_set_advection_bcs_4d(double* u, const ok_geom* g, const double* vel1, const double* vel2, int at_lo1,
                             int at_hi1, int at_lo2, int at_hi2, int x_periodic, int y_periodic, ok_ic_fn ic,
                             void* ic_ctx);
/* the "JB" variants (use_new_bcs): KineticSpeciesF.f:1301-1520, 1524-1733 */
void ok_set_acceleration_bcs_4d_jb(double* u, const ok_geom* g, const double* vel3, const double* vel4, int at_lo3,
                                   int at_hi3, int at_lo4, int at_hi4, ok_ic_fn ic, void* ic_ctx);
void ok_set_advection_bcs_4d_jb(double* u, const ok_geom* g, const double* vel1, const double* vel2, int at_lo1,
                                int at_hi1, int at_lo2, int at_hi2, int x_periodic, int y_periodic, ok_ic_fn ic,
                                void* ic_ctx);
/* vel1 (n1d+1,n2d,n3d,n4d), vel2 (n2d+1,n3d,n4d,n1d): face-velocity arrays as the reference holds them */
void ok_advection_derivatives_4d(double* rhs, const double* f, const ok_geom* g, const double* vel1,
                                 const double* vel2);
void ok_acceleration_derivatives_4d(double* rhs, const double* f, const ok_geom* g,
                                    const double* vel3, const double* vel4);

/* velocities: (n3d,n4d,2) cell-centre (vx,vy) */
void ok_compute_currents(const ok_geom* g, const double* velocities, const double* u, const double* vz,
                         double* Jx, double* Jy, double* Jz);
double ok_compute_ke_e_dot(const ok_geom* g, const double* u, double charge, const double* velocities,
                           const double* ext_efield, double ke_e_dot_in);

/* appendkrook (KineticSpeciesF.f:2995-3034): rhs -= nu(x,y)/dt * (u - IC) where nu != 0; nu: (n1d,n2d) */
/* TrigTZSource (TZSourceF.f:10-137): the manufactured-solution source added to rhs over the whole data box, and
 * error = soln - f_exact; lo = global index of array cell 0 in x and y */
void ok_set_trig_tz_source(double* f, const ok_geom* g, const int* lo, const double* xlo, const double* dx, double time,
                           const double* velocities, double amp);
void ok_compute_trig_tz_source_error(double* error, const double* soln, const ok_geom* g, const int* lo, const double* xlo,
                                     const double* dx, double time, const double* velocities, double amp);
/* ElectronTrigTZSource (ElectronTZSourceF.f:10-143): kx = ky = 4 */
void ok_set_electron_trig_tz_source(double* f, const ok_geom* g, const int* lo, const double* xlo, const double* dx, double time,
                                    const double* velocities, double amp);
void ok_compute_electron_trig_tz_source_error(double* error, const double* soln, const ok_geom* g, const int* lo,
                                              const double* xlo, const double* dx, double time, const double* velocities,
                                              double amp);
/* TwoSpecies_ElectronTrigTZSource (species 0) / TwoSpecies_IonTrigTZSource (species 1); dparams = {amp, me, mi} */
void ok_set_two_species_trig_tz_source(double* f, const ok_geom* g, const int* lo, const double* xlo, const double* dx,
                                       double time, const double* velocities, const double* dparams, int species);
void ok_compute_two_species_trig_tz_source_error(double* error, const double* soln, const ok_geom* g, const int* lo,
                                                 const double* xlo, const double* dx, double time, const double* velocities,
                                                 const double* dparams, int species);
void ok_append_krook(double* rhs, const double* u, const ok_geom* g, const double* nu, double dt, ok_ic_fn ic,
                     void* ic_ctx);
/* ---- PitchAngleCollisionOperatorF.f / PitchAngleCollisionOperator.C (loki_oracle_coll.c) ---- */
double ok_pitch_angle_collisionality(double vx, double vy, double vxgrid, double vygrid, const double* range_lo,
                                     const double* range_hi, double vxmin, double vxmax, double vymin, double vymax,
                                     double vfloor, double vthermal, double nu_coef, int order);
/* IVx, IVy, IVth (n1d,n2d): flow and thermal velocity of max(|u|, 1e-10) over the interior velocity cells */
void ok_pitch_angle_fields(double* IVx, double* IVy, double* IVth, const double* u, const ok_geom* g,
                           const double* velocities);
void ok_append_pitch_angle_collision(double* rhs, const double* f, const ok_geom* g, const double* velocities,
                                     const double* IVx, const double* IVy, const double* IVth, const double* vlo,
                                     const double* vhi, const double* range_lo, const double* range_hi, double vfloor,
                                     double nu_coef, int conservative);
double ok_pitch_angle_real_lam(const ok_geom* g, double nu_coef, double vthermal_dt, double vfloor);
/* time-history diagnostics: computeke / computekemaxwell (KineticSpeciesF.f:2447-2559) and the field
 * histories of Poisson / Maxwell ::accumulateSequences (Poisson.C:796-860, Maxwell.C:753-875) */
/* flux-form diagnostics (KineticSpeciesF.f:630-720, 797-910, 1838-1945, 2249-2396, 985-1032, 2734-2990); d = 0..3 */
void ok_face_fluxes_4d(double* flux, double* face, const double* u, const ok_geom* g, const double* vel, int d);
void ok_accum_flux_div_4d(double* rhs, const ok_geom* g, const double* flux1, const double* flux2, const double* flux3,
                          const double* flux4);
double ok_compute_ke_flux(const ok_geom* g, const double* flux1, const double* flux2, const double* flux3,
                          const double* flux4, const double* velocities, const double* vxface_velocities,
                          const double* vyface_velocities, int dir, int side, double mass);
void ok_compute_ke_vel_space_flux(double* ke_flux, const ok_geom* g, const double* flux3, const double* flux4,
                                  const double* vxface_velocities, const double* vyface_velocities, int dir, int side,
                                  double mass);
void ok_compute_ke(const ok_geom* g, const double* u, double mass, const double* velocities, double* out5);
void ok_compute_ke_maxwell(const ok_geom* g, const double* u, double mass, const double* velocities,
                           const double* vz_in, double* out3);
void ok_field_history(const double* em, int n1, int n2, int ng, int ncomp, const double* dx, double* out);

/* ---- ReductionSchedule.C: 4D -> 2D velocity moment on one rank ---- */
void ok_reduce_4d_to_2d(double* dst2d, const double* src4d, const ok_geom* g, double dv, double weight);

/* ---- ParallelArray: periodic wrap in x then y (single rank) ---- */
void ok_periodic_fill_4d(double* u, const ok_geom* g, int periodic_x, int periodic_y);
void ok_periodic_fill_2d(double* u, int n1, int n2, int ng, int ncomp, int periodic_x, int periodic_y);

/* ---- KineticSpecies.C:1656-1694, 1953-2048: velocity tables (non-relativistic) ---- */
void ok_build_velocity_tables(const ok_geom* g, const int lo34[2], double vxlo, double vylo,
                              double* velocities, double* vxface_vel, double* vyface_vel);
void ok_initialize_velocity(const ok_geom* g, const double* velocities, double* vel1, double* vel2);

/* ---- PoissonF.f / LokiPoissonSolveFFT.C ---- */
void ok_neutralize_charge(double* rho, int n1, int n2, int ng);
void ok_poisson_symbols(int nx, int ny, double Lx, double Ly, int order, double* sx, double* sy);
void ok_poisson_fft_solve(double* phi, const double* rho, int nx, int ny, int ng, const double* sx,
                          const double* sy);
void ok_efield_from_potential(double* em_vars, const double* phi, int n1, int n2, int ng, int order,
                              int em_vars_dim, const double* dx);

/* ---- MaxwellF.f ---- */
void ok_xpby2d(double* x, const double* y, double b, int n1, int n2, int ng, int ncomp);
void ok_maxwell_eval_rhs(double* rhs, const double* em, const double* Jx, const double* Jy,
                         const double* Jz, int n1, int n2, int ng, int order, const double* dx,
                         double light_speed, double av_weak, double av_strong);
/* MaxwellF.f:10-58 zeroghost2d, :359-389 maxwelladdantennasource, :473-657 maxwellsetembcs, :661-731 maxwellsetvzbcs;
 * at[4] = {x low, x high, y low, y high}: the box touches that physical boundary */
void ok_zero_ghost_2d(double* u, int n1, int n2, int ng, int dim);
void ok_maxwell_add_antenna_source(double* dem, const double* antenna, int n1, int n2, int ng);
void ok_maxwell_set_em_bcs(double* em, int n1, int n2, int order, const int* at, int x_periodic, int y_periodic, double c);
void ok_maxwell_set_vz_bcs(double* vz, int n1, int n2, int order, const int* at, int x_periodic, int y_periodic);
void ok_maxwell_eval_vz_rhs(double* rhs, const double* em, double charge_per_mass, int n1, int n2, int ng);

/* ---- composite: one single-rank VP RHS evaluation and one RK4 step (VPSystem.C:372-476,
 *      RK4Integrator.H:66-171), x/y periodic, FFT Poisson.  All species share one ok_geom except n/dx.
 */
typedef struct ok_species {
  ok_geom g;
  double mass, charge, bz_const;
  double vlo[2], vhi[2];      /* velocity-domain bounds                                */
  ok_ic_fn ic; void* ic_ctx;  /* inflow BC                                             */
  int has_driver;             /* ShapedRampedCosineDriver acting on this species       */
  double driver[16];          /* parameter vector in the reference's enum order        */
  double driver_phase;
  int driver_shape_type;
} ok_species;

/* ShapedRampedCosineDriverF.f:10-189: adds the driver field into em_vars(:,:,0) and ext_efield(:,:,0)
 * over the whole 2D data box (lo_index = global index of element 0) */
void ok_shaped_ramped_driver(double* em_vars, double* ext_efield, int n1d, int n2d, int lo1, int lo2,
                             const double* xlo, const double* dx, int sums_into, double t,
                             const double* e_ext_param, double phase, int shape_type);

typedef struct ok_vp_work ok_vp_work;
ok_vp_work* ok_vp_work_create(int nspecies, const ok_species* sp, const double* xlo, const double* xhi);
void ok_vp_work_destroy(ok_vp_work* w);
/* options beyond the benchmark decks: non-periodic x / y (KineticSpecies.H:998-1031; the Poisson solve stays periodic,
 * Poisson.C:147-152), use_new_bcs (VPSystem.C:819-821), a Krook layer nu(n1d,n2d) of species s (KineticSpecies.C:1049-1062) */
void ok_vp_set_options(ok_vp_work* w, int nonperiodic_x, int nonperiodic_y, int use_new_bcs);
void ok_vp_set_krook(ok_vp_work* w, int s, const double* nu);
/* a pitch-angle collision operator on species s (KineticSpecies.C:1036-1046, 666-672); p = {range_lo[2], range_hi[2],
 * vfloor, vthermal_dt, nuCoeff, conservative}; NULL removes it */
void ok_vp_set_pitch_angle(ok_vp_work* w, int s, const double* p);
/* twilight-zone source of species s in completeRHS: on = 0 none, 1 TrigTZSource, 2 ElectronTrigTZSource,
 * 3 / 4 TwoSpecies_Electron / IonTrigTZSource (these two with the masses) */
void ok_vp_set_trig_tz(ok_vp_work* w, int s, int on, double amp, double electron_mass, double ion_mass);
void ok_vp_set_dt(ok_vp_work* w, double dt);   /* the a_dt of a bare ok_vp_eval_rhs call (completeRHS) */
/* rhs[s], f[s]: 4D arrays incl. ghosts; f's ghosts are modified like the reference does.
 * ke_e_dot[s] receives rhs.m_integrated_ke_e_dot for driven species. */
void ok_vp_eval_rhs(ok_vp_work* w, double** rhs, double** f, double time, double* ke_e_dot, double* axmax,
                    double* aymax);
const double* ok_vp_em_vars(const ok_vp_work* w); /* (n1d,n2d,2) E field of the last evalRHS */
const double* ok_vp_rho(const ok_vp_work* w);     /* neutralised net charge density          */
/* RK4Integrator / RK6Integrator ::advance; ke[s] = m_integrated_ke_e_dot of the state (in: old, out: new) */
void ok_vp_rk4_step(ok_vp_work* w, double** f_new, double** f_old, double time, double dt, double* ke);
void ok_vp_rk6_step(ok_vp_work* w, double** f_new, double** f_old, double time, double dt, double* ke);
/* KineticSpecies::computeDt + VPSystem::stableDt from given axmax/aymax */
/* axmax/aymax of the most recent evalRHS (the last RK stage of the previous step): what
 * VPSystem::stableDt sees (KineticSpecies.C:771-772) */
/* the eight boundary kinetic-energy fluxes per species (KineticSpecies.C:2052-2097): out[8 s + 2 dir + side] */
void ok_vp_ke_flux_history(ok_vp_work* w, double** f, double* out);
void ok_vp_last_accel_max(const ok_vp_work* w, double* axmax, double* aymax);
double ok_vp_stable_dt(const ok_vp_work* w, const double* axmax, const double* aymax, int rk_order);

/* ---- composite: single-rank Vlasov-Maxwell (VMSystem.C:407-549, Maxwell.C:299-353, 562-623); x/y periodic,
 *      no drivers/antennae/particles.  em: (n1d,n2d,6) Ex,Ey,Ez,Bx,By,Bz; vz[s]: (n1d,n2d).  loki_oracle_vm.c */
typedef struct ok_vm_work ok_vm_work;
ok_vm_work* ok_vm_work_create(int nspecies, const ok_species* sp, const double* xlo, const double* xhi,
                              double light_speed, double av_weak, double av_strong);
void ok_vm_work_destroy(ok_vm_work* w);
void ok_vm_eval_rhs(ok_vm_work* w, double** rhs, double* rhs_em, double** rhs_vz, double** f, double* em,
                    double** vz, double time, double* axmax, double* aymax);
const double* ok_vm_net_current(const ok_vm_work* w, int comp); /* net Jx/Jy/Jz of the last evalRHS */
void ok_vm_rk4_step(ok_vm_work* w, double** f_new, double** f_old, double* em_new, double* em_old,
                    double** vz_new, double** vz_old, double time, double dt);
void ok_vm_rk6_step(ok_vm_work* w, double** f_new, double** f_old, double* em_new, double* em_old,
                    double** vz_new, double** vz_old, double time, double dt);
void ok_vm_last_accel_max(const ok_vm_work* w, double* axmax, double* aymax);
double ok_vm_stable_dt(const ok_vm_work* w, const double* axmax, const double* aymax, int rk_order);
/* SimpleEMICF.f:10-47 (field = 1: E, 4: B) and SimpleVELICF.f:10-40, one wave, whole data box */
void ok_simple_em_ic(double* em, int n1, int n2, int ng, const double* xlo, const double* dx, int field,
                     double xamp, double yamp, double zamp, double kx, double ky, double phi);
void ok_simple_vel_ic(double* vz, int n1, int n2, int ng, const double* xlo, const double* dx, double amp,
                      double kx, double ky, double phi);

/* unfused CPU timing leg used by bench.py (same passes as the reference does per RK4 stage) */
double ok_time_rk4_stage_reference_style(const ok_geom* g, int nthreads, int reps);

#ifdef __cplusplus
}
#endif
#endif
