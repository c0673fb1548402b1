/*
 * loki_b200.h -- C ABI of the B200-native Vlasov right-hand-side path for LLNL/LOKI.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry point
 * replaces one interface of the reference (cited file:line into the LOKI source tree).  The
 * reference's only FFI is its C++ -> Fortran-77 seam (KineticSpeciesF.H, PoissonF.H, MaxwellF.H:
 * extern "C", trailing-underscore symbols, every argument by reference, boxes expanded to 8 ints by
 * BOX4D_TO_FORT, tbox/Box.H:893-897).  Here the same operators take DEVICE pointers, a geometry
 * struct instead of 16 box integers, and return an int status (0 = ok) instead of aborting
 * (Loki_Defines.H:33-39); lk_last_error() gives the message.  INTEGRATION.md shows the stub a LOKI
 * maintainer adds to KineticSpecies.H / VPSystem.C to bind them.
 *
 * Arrays: fp64, Fortran order (first index contiguous), 4D distribution arrays hold the interior
 * grown by ng ghost cells in all four directions exactly like ParallelArray (ParallelArray.H:1121-1132,
 * ParallelArray.C:664-665) so restart dumps (RestartWriter.C:543-560) can be uploaded verbatim.
 * All indices into tables below are 0-based offsets into the data box.
 *
 * All functions are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream)
 * unless stated otherwise.  There is NO CPU fallback: without a CUDA device every compute entry
 * point fails with LK_ERR_CUDA.
 */
#ifndef LOKI_B200_H
#define LOKI_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LK_OK 0
#define LK_ERR_ARG 1
#define LK_ERR_CUDA 2
#define LK_ERR_UNSUPPORTED 3

/* Geometry of one species' local 4D box.  Replaces BOX4D_TO_FORT(dataBox()), BOX4D_TO_FORT(interiorBox())
 * and PROBLEMDOMAIN_TO_FORT (ProblemDomain.H:318-321). */
typedef struct lk_geom {
  int n[4];     /* interior cells (Nx, Ny, Nvx, Nvy) on this device                        */
  int ng;       /* ghost width: 2 for order 4, 3 for order 6 (KineticSpecies.C:155-160)    */
  int order;    /* spatial_solution_order, 4 or 6                                          */
  double dx[4]; /* cell sizes                                                               */
} lk_geom;

/* How the acceleration on velocity faces is formed; replaces the materialised vel3/vel4 arrays of
 * setphasespacevel4D / setphasespacevelmaxwell4D (KineticSpeciesF.f:42-114, 118-197): the kernels
 * evaluate the same expression on the fly from the 2D fields. */
typedef struct lk_accel {
  int kind;               /* 0: Vlasov-Poisson, field = accel(n1d,n2d,2) already scaled by q/m
                             (KineticSpecies.C:755); 1: Vlasov-Maxwell, field = em_vars(n1d,n2d,6)
                             Ex,Ey,Ez,Bx,By,Bz and vz(n1d,n2d);
                             2: the materialised arrays themselves (what the Fortran-ABI entry points of
                             loki_b200_f77.h receive): field = vel3(n3d+1,n4d,n1d,n2d), vz = vel4(n4d+1,n1d,
                             n2d,n3d), the other members unused; not for lk_max_accel / lk_set_phase_space_vel_4d */
  const double* field;    /* device */
  const double* vz;       /* device, kind 1 only */
  const double* vxface_velocities; /* device (n3d+1, n4d, 2)  KineticSpecies.C:2033-2039 */
  const double* vyface_velocities; /* device (n3d, n4d+1, 2)  KineticSpecies.C:2040-2046 */
  double normalization;   /* q/m (or q when relativistic), KineticSpecies.C:748-754 */
  double bz_const;
} lk_accel;

/* Inflow values for the velocity-boundary fill; replaces the initialconditionatpoint_ callback
 * (ICInterface.C:36-57) that setaccelerationbcs4d_ makes per ghost cell. */
typedef struct lk_inflow {
  int kind;          /* 0: zero inflow; 1: factored fnorm*fv(i3,i4)*fx(i1,i2)*frac
                        (PerturbedMaxwellianIC.C:267-289); 2: fx*fv + fx2*fv2
                        (InterpenetratingStreamIC.C:275-278, two-sided); 3: explicit ghost-layer tables;
                        4: fv*fx*fx2 (InterpenetratingStreamIC.C:279-281, centred) */
  const double* fx;  /* device (n1d,n2d)  */
  const double* fv;  /* device (n3d,n4d)  */
  const double* fx2; /* kind 2 */
  const double* fv2; /* kind 2 */
  double fnorm, frac;
  const double* ghost3; /* kind 3: (n1d,n2d,2*ng,n4d): layers [0,ng) below n3a (layer k = cell n3a-ng+k), [ng,2ng) above */
  const double* ghost4; /* kind 3: (n1d,n2d,n3d,2*ng) */
} lk_inflow;

/* One fused Runge-Kutta stage update applied to the freshly evaluated rhs.
 * RK4 (RK4Integrator.H:149-171):
 *   delta_out = (delta_in ? delta_in : 0) + w_delta * rhs           [addSolnData(m_delta, m_rhs, a_dt_eval)]
 *   pred      = f_old + c_pred * (use_delta ? delta_out : rhs)       [copySolnData + addSolnData]
 * RK6 (RK6Integrator.H:105-130): the stage result k_i is written through lk_vlasov_rhs's rhs_out and
 *   pred      = ((f_old + c_prev[0]*k_prev[0]) + ... + c_prev[n_prev-1]*k_prev[n_prev-1]) + c_pred * rhs
 * i.e. the next stage's predictor (or the end-of-step sum with the b weights), added in the reference's
 * order.  n_prev = 0 for RK4. */
struct lk_inflow;
typedef struct lk_rk_update {
  const double* f_old;
  const double* delta_in; /* NULL in stage 1 (delta starts from zero) */
  double* delta_out;      /* NULL in the last stage if the caller does not need it */
  double* pred;           /* must not alias f_eval (neighbours still read the old values) */
  double w_delta, c_pred;
  int use_delta;          /* 1 in RK4 stage 4 */
  int n_prev;             /* RK6: number of earlier stage results to add (0..7) */
  const double* k_prev[7];
  double c_prev[7];
  int wrap;               /* bit 0 / bit 1: also write pred's periodic ghost cells in x / y (the copies
                             communicatePeriodicBoundaries would make, ParallelArray.H:580-606), so that the
                             next stage needs no separate wrap pass; needs n >= 2*ng in that direction */
  const struct lk_inflow* accel_bcs; /* non-NULL: the stage also does setaccelerationbcs4d_ (KineticSpecies.H:421-453) for
                             the evaluated state f -- all four velocity boundaries on this device, this inflow description,
                             the stage's own acceleration */
  int inflow_preset;      /* with accel_bcs.  1: the caller guarantees that f's velocity ghost layers hold the inflow sample
                             (lk_preset_inflow_ghosts_4d).  The pipelined kernel then folds the fill into its boundary tiles:
                             ghosts facing an outflow are extrapolated in shared memory, f's ghost layers are neither
                             rewritten nor invalidated (lk_vlasov_stage_folds_bcs tells whether a given call does this).
                             0, or a call the pipelined kernel does not take: lk_set_acceleration_bcs_4d runs on f first and
                             WRITES f's velocity ghosts (f is then no longer preset) */
  const double* krook_nu; /* non-NULL: completeRHS's Krook layer (KineticSpecies.C:1049-1062, appendkrook_): where
                             nu(n1d,n2d) != 0, rhs -= nu / krook_dt * (f - f_IC) before the update, f_IC sampled from
                             krook_ic's tables (kinds 1, 2, 4).  Applied by the per-cell epilogue of the generic kernels
                             (the pipelined kernel is not taken); with rhs_out the stored rhs includes the term */
  double krook_dt;        /* the step's dt (the a_dt of completeRHS) */
  const struct lk_inflow* krook_ic;
  int tile_set;           /* 0: the whole box.  1: only the CTA tiles (32 x 8 cells in x, y) that touch a face of a
                             direction named in cut_dirs; 2: only the others.  The two launches together equal one launch
                             with 0, bit for bit; a rank of a decomposed run issues 1, starts the halo exchange of pred, then
                             issues 2.  Pipelined kernel only: lk_vlasov_stage_can_split tells, other calls fail with
                             LK_ERR_UNSUPPORTED */
  int cut_dirs;           /* with tile_set: bit 0 x, bit 1 y */
} lk_rk_update;

/* ---- library ---- */
int lk_version(void);
const char* lk_last_error(void);
/* 0 = production arithmetic (FMA contraction, one reciprocal per WENO fit; within 1e-12 of the
 * reference), 1 = strict: the reference's operation order with no contraction, bit-identical to a
 * gfortran -O2 build of KineticSpeciesF.f.  Returns the previous mode. */
int lk_set_strict(int strict);
int lk_get_strict(void);
int lk_device_count(void);
/* 0 = marching shared-memory kernel (default; aligned grids with an RK4-shaped update take its pipelined
 * instantiation, lk_pipe.cuh), 1 = one-thread-per-cell cross-check kernel, 2 = marching kernel, generic
 * instantiation only (cross-check of the pipelined one: the two give identical bits) */
int lk_set_rhs_variant(int variant);

/* ---- a1/a2: WENO43Fit4D / WENO65Fit4D (KineticSpeciesF.f:723-790, 914-979); test hook ----
 * u: count x order-sized stencils (um2,um1,u0,up1 | um3..up2), vel: count upwind velocities */
int lk_weno_fit(int order, const double* u, const double* vel, double* face, int64_t count, void* stream);

/* ---- a10: xpby4d_ (KineticSpeciesF.H:41-61, KineticSpeciesF.f:10-38): x += b*y on the interior ---- */
int lk_xpby4d(double* x, const double* y, double b, const lk_geom* g, void* stream);

/* ---- a4/a5: setphasespacevel4d_ / setphasespacevelmaxwell4d_ (KineticSpeciesF.H:191-250).
 * lk_max_accel returns {axmax, aymax} (device, 2 doubles) without materialising vel3/vel4;
 * lk_set_phase_space_vel_4d also writes the two rotated 4D arrays for callers that still want them. */
int lk_max_accel(const lk_geom* g, const lk_accel* a, double* axaymax_dev, void* stream);
int lk_set_phase_space_vel_4d(double* vel3, double* vel4, const lk_geom* g, const lk_accel* a,
                              double* axaymax_dev, void* stream);

/* ---- a6: setaccelerationbcs4d_ (KineticSpeciesF.H:97-127, KineticSpeciesF.f:1036-1162) ----
 * at_[0..3] = box touches global vx-low, vx-high, vy-low, vy-high boundary */
int lk_set_acceleration_bcs_4d(double* f, const lk_geom* g, const lk_accel* a, const lk_inflow* ic,
                               const int at[4], void* stream);

/* The inflow half of a6, once: every velocity ghost cell of f := the inflow sample of `ic` at that cell (what
 * setaccelerationbcs4d_ stores where the acceleration points inward, KineticSpeciesF.f:1087-1112, 1132-1159; it
 * depends on the position only).  An array prepared this way can be handed to lk_vlasov_stage with
 * lk_rk_update.inflow_preset = 1. */
int lk_preset_inflow_ghosts_4d(double* f, const lk_geom* g, const lk_inflow* ic, void* stream);

/* ---- a14: periodic wrap of x then y ghosts on one device (ParallelArray.H:580-606) ---- */
int lk_periodic_fill_4d(double* f, const lk_geom* g, int periodic_x, int periodic_y, void* stream);
/* a7: setadvectionbcs4d_ (KineticSpeciesF.f:1166-1297; wrapper KineticSpecies.H:998-1031): physical boundaries
 * of a NON-periodic x / y direction (outflow extrapolation or inflow from the IC tables, by the sign of the
 * face velocity).  at_boundary = {x lo, x hi, y lo, y hi}: does this box touch that global boundary.
 * velocities: the cell-centre table (n3d,n4d,2); inflow kinds 0, 1, 2, 4. */
int lk_set_advection_bcs_4d(double* f, const lk_geom* g, const double* velocities, const lk_inflow* ic,
                            const int at_boundary[4], int periodic_x, int periodic_y, void* stream);
/* a6': the "JB" variants selected by use_new_bcs (VPSystem.C:819-821): setaccelerationbcs4djb_ /
 * setadvectionbcs4djb_ (KineticSpeciesF.f:1301-1520, 1524-1733).  Inflow (lower side: face velocity > 0, upper
 * side: < 0) samples the IC tables, otherwise ghosts are extrapolated with the binomial formula of order
 * min(interior extent, solution_order).  at_boundary as above ({vx lo, vx hi, vy lo, vy hi} for the
 * acceleration flavour): the reference derives it from the cell coordinate (:1368-1371). */
int lk_set_acceleration_bcs_4d_jb(double* f, const lk_geom* g, const lk_accel* a, const lk_inflow* ic,
                                  const int at_boundary[4], void* stream);
int lk_set_advection_bcs_4d_jb(double* f, const lk_geom* g, const double* velocities, const lk_inflow* ic,
                               const int at_boundary[4], int periodic_x, int periodic_y, void* stream);
/* halo slabs for the (x,y)-decomposed multi-GPU exchange (ParallelArray.C:925-1113, faces only).
 * dir 0 = x, 1 = y; side 0 = low, 1 = high.  pack copies the ng interior layers next to that side
 * into a dense buffer; unpack writes a received buffer into the ghost layers on that side.
 * Buffer layout: x: (ng, n2, n3d, n4d); y: (n1d, ng, n3d, n4d). */
int64_t lk_halo_count(const lk_geom* g, int dir);
int lk_halo_pack(double* buf, const double* f, const lk_geom* g, int dir, int side, void* stream);
int lk_halo_unpack(double* f, const double* buf, const lk_geom* g, int dir, int side, void* stream);

/* ---- a3 + a8: computeadvectionderivatives4d_ + computeaccelerationderivatives4d_
 * (KineticSpeciesF.H:293-341, KineticSpeciesF.f:1949-2245).  velocities: device (n3d,n4d,2) cell-centre
 * table (KineticSpecies.C:2026-2032).  advection ASSIGNS rhs, acceleration ACCUMULATES. */
int lk_advection_derivatives_4d(double* rhs, const double* f, const lk_geom* g, const double* velocities,
                                void* stream);
int lk_acceleration_derivatives_4d(double* rhs, const double* f, const lk_geom* g, const lk_accel* a,
                                   void* stream);

/* ---- fused production path: a3 + a4/a5 + a8 (+ a10/a11 when upd != NULL) in ONE pass ----
 * rhs_out may be NULL when upd != NULL.  f must have valid x/y ghosts and velocity-boundary ghosts. */
int lk_vlasov_rhs(double* rhs_out, const double* f, const lk_geom* g, const double* velocities,
                  const lk_accel* a, const lk_rk_update* upd, void* stream);

/* Velocity moments of the stage's OUTPUT (the predictor = the next stage's input), accumulated by the
 * fused kernel's epilogue so that chargeDensity / currentDensity / computekeedot of the next evalRHS
 * (VPSystem.C:396, KineticSpecies.C:853-895, 1084-1093) cost no extra pass over f.
 * partial: device scratch of nmom * lk_stage_moment_parts(g) * n[0] * n[1] doubles; nmom = 1: sum f;
 * nmom = 3: sum f, sum vx f, sum vy f.  Sums are formed in a fixed order (deterministic), which is not the
 * reference's sequential order: production arithmetic only (strict mode uses lk_reduce_4d_to_2d). */
typedef struct lk_stage_moments {
  int nmom;
  double* partial;
  int64_t capacity; /* doubles */
} lk_stage_moments;
int lk_stage_moment_parts(const lk_geom* g);
int lk_vlasov_stage(double* rhs_out, const double* f, const lk_geom* g, const double* velocities,
                    const lk_accel* a, const lk_rk_update* upd, const lk_stage_moments* mom, void* stream);
/* The stage update alone, from a materialised rhs: delta / pred of `upd` exactly as the fused kernel's epilogue forms
 * them (RK4Integrator.H:149-171, RK6Integrator.H:96-131), on the interior.  For callers that modify the rhs between its
 * evaluation and the update (completeRHS's Krook layer, KineticSpecies.C:1049-1062).  upd->wrap is ignored. */
int lk_rk_stage_update(const double* rhs, const lk_geom* g, const lk_rk_update* upd, void* stream);
/* 1 when lk_vlasov_stage(rhs_out, ., g, ., a, upd, ...) honours upd->tile_set (it takes the pipelined kernel) */
int lk_vlasov_stage_can_split(const double* rhs_out, const lk_geom* g, const lk_accel* a, const lk_rk_update* upd);
/* 1 when lk_vlasov_stage(rhs_out, ., g, ., a, upd, ...) would fold upd->accel_bcs into the pipelined kernel (f's velocity
 * ghosts stay as they are), 0 when it would run lk_set_acceleration_bcs_4d on f first */
int lk_vlasov_stage_folds_bcs(const double* rhs_out, const lk_geom* g, const lk_accel* a, const lk_rk_update* upd);
/* dst_m(n1d,n2d) = (sum of partials of moment m) * dv * weight, ghosts zeroed (ReductionSchedule.C:86-89) */
int lk_moments_finish(double* dst0, double* dst1, double* dst2, const lk_stage_moments* mom, const lk_geom* g,
                      double dv, double weight, void* stream);
/* computekeedot_ from the vx moment: out = charge*dx*dy*dvx*dvy * sum_xy ext(x,y,0) * M1(x,y) */
int lk_ke_e_dot_from_moments(double* out_dev, const lk_stage_moments* mom, const lk_geom* g, double charge,
                             const double* ext_efield, void* stream);

/* ---- a12: ReductionSchedule 4D->2D (ReductionSchedule.C:69-113, 421-444): dst(n1d,n2d) ghosts zeroed,
 * dst = (sum_{i3,i4} f) * dv * weight ---- */
int lk_reduce_4d_to_2d(double* dst2d, const double* f, const lk_geom* g, double dv, double weight,
                       void* stream);
/* ---- a13: computecurrents_ + three reductions fused (KineticSpeciesF.f:2400-2443,
 * KineticSpecies.C:853-895): Jx,Jy,Jz (n1d,n2d) ---- */
int lk_current_density(double* Jx, double* Jy, double* Jz, const double* f, const lk_geom* g,
                       const double* velocities, const double* vz, double dv, double weight, void* stream);
/* ---- a9: computekeedot_ (KineticSpeciesF.f:2563-2602); out_dev: 1 double ---- */
int lk_ke_e_dot(double* out_dev, const double* f, const lk_geom* g, double charge, const double* velocities,
                const double* ext_efield, void* stream);

/* ---- a16: Poisson (PoissonF.f:10-123, LokiPoissonSolveFFT.C:31-170, EMSolverBase.C:270-371) ----
 * 2D arrays (n1d,n2d[,comp]).  lk_poisson_plan builds the symbol/twiddle tables on the device. */
typedef struct lk_poisson_plan lk_poisson_plan;
int lk_poisson_plan_create(lk_poisson_plan** plan, int nx, int ny, int ng, int order, double Lx, double Ly);
void lk_poisson_plan_destroy(lk_poisson_plan* plan);
/* rho is neutralised in place; phi gets interior + periodic ghosts; em_vars comps 0,1 = Ex,Ey with
 * periodic ghosts (the whole electricField sequence) */
int lk_electric_field(lk_poisson_plan* plan, double* rho, double* phi, double* em_vars, const double* dx,
                      void* stream);
int lk_periodic_fill_2d(double* u, int n1, int n2, int ng, int ncomp, int periodic_x, int periodic_y,
                        void* stream);
/* the two Fortran pieces of electricField on their own (Level 0): neutralizecharge4d_ (PoissonF.f:10-64) and
 * computeefieldfrompotential_ (PoissonF.f:68-123: Ex, Ey into comps 0, 1 of em_vars, interior only) */
int lk_neutralize_charge(double* rho, int n1, int n2, int ng, void* stream);
int lk_efield_from_potential(double* em_vars, const double* phi, int n1, int n2, int ng, int order, double dx,
                             double dy, void* stream);
/* x += b*y on the interior of a (n1d,n2d,ncomp) array: xpby2d_ (MaxwellF.f:62-93) */
int lk_xpby2d(double* x, const double* y, double b, int n1, int n2, int ng, int ncomp, void* stream);
/* computeAcceleration glue (KineticSpecies.C:697-774): accel = (em[0:2] + ext) * normalization */
int lk_form_accel(double* accel, const double* em_vars, const double* ext_efield, double normalization,
                  int n1, int n2, int ng, void* stream);

/* ---- a17: Maxwell (MaxwellF.f:97-355, 442-469) ---- */
int lk_maxwell_rhs(double* rhs, const double* em, const double* Jx, const double* Jy, const double* Jz,
                   int n1, int n2, int ng, int order, const double* dx, double light_speed, double av_weak,
                   double av_strong, void* stream);

/* The boundary routines of MaxwellF.f that Maxwell::fillGhostCells / evalRHS call when a direction is not periodic
 * (Maxwell.C:562-623), and the antenna source (Maxwell.C:299-353).  Arrays are device (n1d, n2d, ncomp) over the interior
 * n1 x n2 grown by ng; at[4] = {x low, x high, y low, y high}: this rank's box touches that physical boundary (the
 * Fortran's m1a .eq. 0, m1b .eq. nx-1, m2a .eq. 0, m2b .eq. ny-1).
 *   lk_zero_ghost_2d               zeroghost2d_ (MaxwellF.f:10-58): ghost layers of all ncomp components = 0
 *   lk_maxwell_add_antenna_source  maxwelladdantennasource_ (:359-389): dEMvars -= antenna_source, interior, 6 components
 *   lk_maxwell_set_em_bcs          maxwellsetembcs_ (:473-657): third-order extrapolation into the ghost layers of the
 *                                  six fields, then the incoming characteristic of (Ey,Bz), (-Ez,By) [x edges] and
 *                                  (-Ex,Bz), (Ez,Bx) [y edges] zeroed; interior rows / columns only, corners untouched
 *   lk_maxwell_set_vz_bcs          maxwellsetvzbcs_ (:661-731): even reflection of vz about the boundary cell, x edges
 *                                  over every row of the data box, then y edges over every column */
int lk_zero_ghost_2d(double* u, int n1, int n2, int ng, int ncomp, void* stream);
int lk_maxwell_add_antenna_source(double* dem, const double* antenna_source, int n1, int n2, int ng, void* stream);
int lk_maxwell_set_em_bcs(double* em, int n1, int n2, int order, const int at[4], int x_periodic, int y_periodic,
                          double light_speed, void* stream);
int lk_maxwell_set_vz_bcs(double* vz, int n1, int n2, int order, const int at[4], int x_periodic, int y_periodic,
                          void* stream);

/* maxwellevalvzrhs_ (MaxwellF.f:442-469): dvz = (q/m) Ez on the interior of a (n1d,n2d) array */
int lk_maxwell_vz_rhs(double* dvz, const double* em, int n1, int n2, int ng, double charge_per_mass, void* stream);

/* The reference's twilight-zone (manufactured-solution) forcings, called from KineticSpecies::completeRHS
 * (KineticSpecies.C:1077-1080: source added to the rhs) and putToRestart (:987-1004: error against the exact solution):
 *   kind 0  TrigTZSource                      settrigtzsource_ / computetrigtzsourceerror_ (TZSourceF.f:10-137), deck TrigTZ
 *   kind 1  ElectronTrigTZSource              setelectrontrigtzsource_ / ...error_ (ElectronTZSourceF.f:10-143), deck EPWTZ
 *   kind 2  TwoSpecies_ElectronTrigTZSource   settwoelectrontrigtzsource_ / ...error_ (TwoSpecies_ElectronTZSourceF.f), deck IAWTZ
 *   kind 3  TwoSpecies_IonTrigTZSource        settwoiontrigtzsource_ / ...error_ (TwoSpecies_IonTZSourceF.f), deck IAWTZ
 * params = the Fortran's dparams {amp, electron_mass, ion_mass} (the masses matter for kinds 2 / 3 only).
 * The sources' transcendental factors are separable and time-independent except the sines / cosines of t:
 * lk_trig_tz_tables builds them on the HOST with libm (the Fortran's argument expressions; lo = global index of array
 * cell 0 in x and y, xlo = the domain's lower corner) and stores lk_trig_tz_table_count doubles on the device; the two
 * kernels then evaluate the Fortran's expression tree on those operands, so the result carries the reference's bits.
 * Both run over the whole data box like the Fortran.  lk_trig_tz_tables synchronises. */
int lk_trig_tz_table_count(const lk_geom* g, int64_t* count);
int lk_trig_tz_tables(double* tables, const lk_geom* g, const int lo[2], const double xlo[2], const double* velocities,
                      int kind, const double* params, void* stream);
int lk_set_trig_tz_source(double* rhs, const lk_geom* g, const double* tables, const double* velocities, double time, int kind,
                          const double* params, void* stream);
int lk_compute_trig_tz_source_error(double* error, const double* soln, const lk_geom* g, const double* tables,
                                    const double* velocities, double time, int kind, const double* params, void* stream);
/* appendkrook_ (KineticSpeciesF.f:2995-3034; completeRHS, KineticSpecies.C:1049-1080): Krook-layer damping of an
 * UNFUSED rhs towards the initial condition, rhs -= nu(x,y)/dt * (u - IC) where nu != 0; nu: (n1d,n2d) device.
 * Level-0 only: the fused stage never materialises rhs, and no benchmark deck has a Krook layer. */
int lk_append_krook(double* rhs, const double* u, const lk_geom* g, const double* nu, double dt, const lk_inflow* ic,
                    void* stream);

/* ---- pitch-angle collision operator (PitchAngleCollisionOperator.C, PitchAngleCollisionOperatorF.f) ----
 * The operator's input keys (PitchAngleCollisionOperator::parseParameters, PitchAngleCollisionOperator.C:195-270). */
typedef struct lk_pitch_angle {
  double range_lo[2], range_hi[2]; /* collision_vel_range_lo / _hi (vx, vy)                                  */
  double vfloor;                   /* collision_vfloor                                                         */
  double vthermal_dt;              /* collision_vthermal_dt: only the time-step estimate uses it (computeRealLam) */
  double nu_coef;                  /* collision_nuCoeff                                                        */
  int conservative;                /* collision_conservative (default 1); 0 is defined for order 4 only        */
} lk_pitch_angle;
/* The reduced fields every evaluate() starts with (PitchAngleCollisionOperator.C:69-117:
 * computepitchanglespeciesmoments_, three velocity-space sums times dvx dvy, ...reducedfields_, ...kec_, one more sum,
 * ...vthermal_): flow velocity IVx, IVy and thermal speed IVth of max(|u|, 1e-10), each (n1d,n2d) device, over the
 * whole configuration data box.  Sums in the reference's order (bit for bit); the rank must hold all of velocity space. */
int lk_pitch_angle_fields(double* IVx, double* IVy, double* IVth, const double* u, const lk_geom* g,
                          const double* velocities, void* stream);
/* appendpitchanglecollision_ (PitchAngleCollisionOperatorF.f:1618-1702): rhs += C(f) on the interior cells; f's velocity
 * ghosts must be filled.  vlo / vhi: the velocity domain (xlo(3:4), xhi(3:4)).  Conservative order 4 / 6: the scheme of
 * the reference's generated code in operator form (equal to rounding, see lk_coll.cuh); non-conservative order 4: bit
 * for bit; non-conservative order 6 applies nothing, as in the reference.  Non-relativistic. */
int lk_append_pitch_angle_collision(double* rhs, const double* f, const lk_geom* g, const double* velocities,
                                    const double* IVx, const double* IVy, const double* IVth, const double* vlo,
                                    const double* vhi, const lk_pitch_angle* p, void* stream);
/* PitchAngleCollisionOperator::computeRealLam (PitchAngleCollisionOperator.C:137-144); host arithmetic, no device work */
double lk_pitch_angle_real_lam(const lk_geom* g, const lk_pitch_angle* p);
/* PitchAngleCollisionOperator::parseParameters' sanity checks (:253-269): LK_OK or LK_ERR_ARG with lk_last_error */
int lk_pitch_angle_check(const lk_geom* g, const double* vlo, const double* vhi, const lk_pitch_angle* p);

/* ---- time-history diagnostics (called at sequence_write_times, not on the stage path) ----
 * computeke_ / computekemaxwell_ (KineticSpeciesF.f:2447-2559): out5_dev = {ke, ke_x, ke_y, px, py}; with
 * vz != NULL the Maxwell flavour: ke includes 0.5 m vz(x,y)^2 f and px = py = 0.  Tree sums (deterministic). */
int lk_compute_ke(double* out5_dev, const double* f, const lk_geom* g, double mass, const double* velocities,
                  const double* vz, void* stream);
/* Poisson::accumulateSequences (Poisson.C:796-860), ncomp = 2: out_dev = {e_max, e_tot, ex_max, ey_max,
 * e_sum_tot}; Maxwell::accumulateSequences (Maxwell.C:753-875), ncomp = 6: {e_max, e_tot, ex_max, ey_max,
 * ez_max, e_sum_tot, b_max, b_tot, bx_max, by_max, bz_max, b_sum_tot} */
int lk_field_history(double* out_dev, const double* em_vars, int n1, int n2, int ng, int ncomp, const double* dx,
                     void* stream);

/* ---- device memory helpers for hosts without their own allocator (synchronous) ---- */
int lk_malloc(void** p, int64_t bytes);
int lk_free(void* p);
int lk_memcpy_h2d(void* dst, const void* src, int64_t bytes);
int lk_memcpy_d2h(void* dst, const void* src, int64_t bytes);
int lk_memset(void* p, int value, int64_t bytes);
int lk_sync(void* stream);
/* CUDA-event timing of every lk_vlasov_rhs launch on its own stream (the dominant kernel's live
 * duration for the roofline): enable(1) resets the counters; summary synchronises and returns the
 * number of timed launches and the sum of their durations in milliseconds. */
int lk_profile_enable(int on);
int lk_profile_summary(int64_t* launches, double* total_ms);
/* kernels launched by this library since load (bench.py's gpu_launches) */
int64_t lk_launch_count(void);
/* how many of those were the pipelined instantiation of the stage kernel (lk_pipe.cuh) */
int64_t lk_pipe_launch_count(void);

/* ---- f2: flux-form diagnostics (loki_b200/csrc/lk_flux.cu) ----
 * The kinetic-energy flux through the eight phase-space boundaries that KineticSpecies::accumulateSequencesCommon
 * (KineticSpecies.C:2052-2097) adds to the time histories.  Face values and fluxes carry the reference's bits in
 * both arithmetic modes; boundary sums are deterministic trees (not the reference's sequential order).
 *
 * lk_face_fluxes_4d: WENO43Avg4D / WENO65Avg4D + computeFlux4D (KineticSpeciesF.f:630-720, 797-910, 2359-2396) for
 * direction dir (0..3), as computeadvectionfluxes4d_ / computeaccelerationfluxes4d_ (:1838-1945, 2249-2355) call
 * them.  vel, face, flux: the rotated face arrays of KineticSpecies.C:1569-1584, extents (nd[dir]+1, nd[dir+1],
 * nd[dir+2], nd[dir+3]) with the direction numbers mod 4.  Entries outside the reference's loop ranges are left alone. */
int lk_face_fluxes_4d(double* flux, double* face, const double* u, const lk_geom* g, const double* vel, int dir,
                      void* stream);
/* accumfluxdiv4d_ (KineticSpeciesF.f:985-1032): rhs = -div(flux) on the interior (the vy term over dvx, as there) */
int lk_accum_flux_div_4d(double* rhs, const lk_geom* g, const double* flux1, const double* flux2, const double* flux3,
                         const double* flux4, void* stream);
/* computekeflux_ (:2734-2893) for a box that touches boundary (dir, side): out_dev[0] = mass * ddir * sum over the
 * boundary face of 0.5 * flux * v^2; flux = the flux array of direction dir */
int lk_ke_flux_from_fluxes(double* out_dev, const lk_geom* g, const double* flux, const double* velocities,
                           const double* vxface_velocities, const double* vyface_velocities, int dir, int side,
                           double mass, void* stream);
/* computekevelspaceflux_ (:2897-2990), dir 2 or 3: ke_flux_xy (n1d,n2d) += the flux through that velocity boundary */
int lk_ke_vel_space_flux(double* ke_flux_xy, const lk_geom* g, const double* flux, const double* vxface_velocities,
                         const double* vyface_velocities, int dir, int side, double mass, void* stream);
/* The product path: all eight boundary fluxes of f straight from f (one face fit per boundary cell, no face or flux
 * arrays).  f needs valid x / y ghosts and velocity-boundary ghosts (lk_set_acceleration_bcs_4d with the same
 * acceleration); out8_dev[2*dir+side]; at_boundary[2*dir+side] = this box touches that boundary of the domain
 * (`n1a .eq. ng1a` ..., :2781-2793), 0 leaves a 0. */
int lk_ke_flux_boundaries(double* out8_dev, const double* f, const lk_geom* g, const double* velocities, const lk_accel* a,
                          double mass, const int at_boundary[8], void* stream);

#ifdef __cplusplus
}
#endif
#endif
