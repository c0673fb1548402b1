/*
 * loki_oracle_vp.c -- CPU ORACLE (test infrastructure only; see loki_oracle.h).
 *
 * Single-rank restatement of the reference's stage sequencing: VPSystem::evalRHS (VPSystem.C:372-476),
 * EMSolverBase::electricField (EMSolverBase.C:270-371), KineticSpecies::computeAcceleration
 * (KineticSpecies.C:697-774), RK4Integrator (RK4Integrator.H:66-171).  x and y periodic, FFT Poisson.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "loki_oracle.h"

struct ok_vp_work {
  int ns;
  ok_species* sp;
  double Lx, Ly;
  /* per species */
  double **velocities, **vxface, **vyface, **vel1, **vel2, **vel3, **vel4, **accel, **rho_s;
  /* field */
  double *rho, *phi, *em, *sx, *sy;
  /* RK scratch */
  double **rhs, **delta;
  double *ke_rhs, *ke_delta;
};

static int64_t vol4(const ok_geom* g) { return ok_nd(g, 0) * ok_nd(g, 1) * ok_nd(g, 2) * ok_nd(g, 3); }

ok_vp_work* ok_vp_work_create(int ns, const ok_species* sp, double Lx, double Ly) {
  ok_vp_work* w = (ok_vp_work*)calloc(1, sizeof(*w));
  w->ns = ns;
  w->sp = (ok_species*)malloc(sizeof(ok_species) * ns);
  memcpy(w->sp, sp, sizeof(ok_species) * ns);
  w->Lx = Lx;
  w->Ly = Ly;
#define PP(name) w->name = (double**)calloc(ns, sizeof(double*))
  PP(velocities); PP(vxface); PP(vyface); PP(vel1); PP(vel2); PP(vel3); PP(vel4); PP(accel); PP(rho_s);
  PP(rhs); PP(delta);
#undef PP
  w->ke_rhs = (double*)calloc(ns, sizeof(double));
  w->ke_delta = (double*)calloc(ns, sizeof(double));
  const ok_geom* g0 = &sp[0].g;
  const int64_t n1d = ok_nd(g0, 0), n2d = ok_nd(g0, 1);
  for (int s = 0; s < ns; ++s) {
    const ok_geom* g = &sp[s].g;
    const int64_t n3d = ok_nd(g, 2), n4d = ok_nd(g, 3);
    w->velocities[s] = (double*)calloc(n3d * n4d * 2, sizeof(double));
    w->vxface[s] = (double*)calloc((n3d + 1) * n4d * 2, sizeof(double));
    w->vyface[s] = (double*)calloc(n3d * (n4d + 1) * 2, sizeof(double));
    w->vel1[s] = (double*)calloc((n1d + 1) * n2d * n3d * n4d, sizeof(double));
    w->vel2[s] = (double*)calloc((n2d + 1) * n3d * n4d * n1d, sizeof(double));
    w->vel3[s] = (double*)calloc((n3d + 1) * n4d * n1d * n2d, sizeof(double));
    w->vel4[s] = (double*)calloc((n4d + 1) * n1d * n2d * n3d, sizeof(double));
    w->accel[s] = (double*)calloc(n1d * n2d * 2, sizeof(double));
    w->rho_s[s] = (double*)calloc(n1d * n2d, sizeof(double));
    w->rhs[s] = (double*)calloc(vol4(g), sizeof(double));
    w->delta[s] = (double*)calloc(vol4(g), sizeof(double));
    int lo34[2] = {-g->ng, -g->ng};
    ok_build_velocity_tables(g, lo34, sp[s].vlo[0], sp[s].vlo[1], w->velocities[s], w->vxface[s], w->vyface[s]);
    ok_initialize_velocity(g, w->velocities[s], w->vel1[s], w->vel2[s]);
  }
  w->rho = (double*)calloc(n1d * n2d, sizeof(double));
  w->phi = (double*)calloc(n1d * n2d, sizeof(double));
  w->em = (double*)calloc(n1d * n2d * 2, sizeof(double));
  w->sx = (double*)calloc(g0->n[0], sizeof(double));
  w->sy = (double*)calloc(g0->n[1] / 2 + 1, sizeof(double));
  ok_poisson_symbols(g0->n[0], g0->n[1], Lx, Ly, g0->order, w->sx, w->sy);
  return w;
}

void ok_vp_work_destroy(ok_vp_work* w) {
  if (!w) return;
  for (int s = 0; s < w->ns; ++s) {
    free(w->velocities[s]); free(w->vxface[s]); free(w->vyface[s]); free(w->vel1[s]); free(w->vel2[s]);
    free(w->vel3[s]); free(w->vel4[s]); free(w->accel[s]); free(w->rho_s[s]); free(w->rhs[s]); free(w->delta[s]);
  }
  free(w->velocities); free(w->vxface); free(w->vyface); free(w->vel1); free(w->vel2); free(w->vel3);
  free(w->vel4); free(w->accel); free(w->rho_s); free(w->rhs); free(w->delta); free(w->ke_rhs);
  free(w->ke_delta); free(w->rho); free(w->phi); free(w->em); free(w->sx); free(w->sy); free(w->sp);
  free(w);
}

const double* ok_vp_em_vars(const ok_vp_work* w) { return w->em; }
const double* ok_vp_rho(const ok_vp_work* w) { return w->rho; }

/* VPSystem::evalRHS on one rank */
void ok_vp_eval_rhs(ok_vp_work* w, double** rhs, double** f, double* ke_e_dot, double* axmax, double* aymax) {
  const ok_geom* g0 = &w->sp[0].g;
  const int ng = g0->ng, n1 = g0->n[0], n2 = g0->n[1];
  const int64_t n1d = ok_nd(g0, 0), n2d = ok_nd(g0, 1), pl = n1d * n2d;
  /* 1. charge density of every species (VPSystem.C:395-397; KineticSpecies.C:1736-1753) */
  for (int s = 0; s < w->ns; ++s) {
    const ok_geom* g = &w->sp[s].g;
    ok_reduce_4d_to_2d(w->rho_s[s], f[s], g, g->dx[2] * g->dx[3], w->sp[s].charge);
  }
  /* 2. net charge + field solve (VPSystem.C:407-414; EMSolverBase.C:270-371) */
  for (int64_t k = 0; k < pl; ++k) w->rho[k] = 0.0;
  for (int s = 0; s < w->ns; ++s)
    for (int64_t k = 0; k < pl; ++k) w->rho[k] += w->rho_s[s][k];
  ok_neutralize_charge(w->rho, n1, n2, ng);
  ok_poisson_fft_solve(w->phi, w->rho, n1, n2, ng, w->sx, w->sy);
  ok_periodic_fill_2d(w->phi, n1, n2, ng, 1, 1, 1);
  for (int64_t k = 0; k < 2 * pl; ++k) w->em[k] = 0.0;
  ok_efield_from_potential(w->em, w->phi, n1, n2, ng, g0->order, 2, g0->dx);
  ok_periodic_fill_2d(w->em, n1, n2, ng, 2, 1, 1);
  for (int s = 0; s < w->ns; ++s) {
    const ok_species* sp = &w->sp[s];
    const ok_geom* g = &sp->g;
    /* 3. ghost fill + advection derivatives (KineticSpecies.H:404-412, 489-507) */
    ok_periodic_fill_4d(f[s], g, 1, 1);
    ok_advection_derivatives_4d(rhs[s], f[s], g, w->vel1[s], w->vel2[s]);
    /* 4. acceleration (KineticSpecies.C:697-774): expansion, drivers, *= q/m, face accelerations */
    for (int64_t k = 0; k < 2 * pl; ++k) w->accel[s][k] = w->em[k];
    if (sp->ext_efield)
      for (int64_t k = 0; k < 2 * pl; ++k) w->accel[s][k] += sp->ext_efield[k];
    double normalization = sp->charge / sp->mass;
    for (int64_t k = 0; k < 2 * pl; ++k) w->accel[s][k] *= normalization;
    ok_set_phase_space_vel_4d(w->vel3[s], w->vel4[s], g, w->vxface[s], w->vyface[s], normalization,
                              sp->bz_const, w->accel[s], &axmax[s], &aymax[s]);
    /* 5. v-boundary fill, acceleration derivatives, completeRHS */
    ok_set_acceleration_bcs_4d(f[s], g, w->vel3[s], w->vel4[s], 1, 1, 1, 1, sp->ic, sp->ic_ctx);
    ok_acceleration_derivatives_4d(rhs[s], f[s], g, w->vel3[s], w->vel4[s]);
    if (sp->ext_efield && ke_e_dot)
      ke_e_dot[s] = ok_compute_ke_e_dot(g, f[s], sp->charge, w->velocities[s], sp->ext_efield, 0.0);
  }
}

/* RK4Integrator::advance with stageAdvance (RK4Integrator.H:66-171).  The external driver field is
 * frozen for the step (tests that need a time-dependent driver update sp.ext_efield between calls of
 * the stage-level API instead). */
void ok_vp_rk4_step(ok_vp_work* w, double** f_new, double** f_old, double dt) {
  static const double THIRD = 1.0 / 3.0;
  double dtOn2 = 0.5 * dt, dtOn3 = THIRD * dt, dtOn6 = 0.5 * dtOn3;
  const double w_eval[4] = {dtOn6, dtOn3, dtOn3, dtOn6};
  const double w_upd[4] = {dtOn2, dtOn2, dt, 1.0};
  double* ax = (double*)calloc(w->ns, sizeof(double));
  double* ay = (double*)calloc(w->ns, sizeof(double));
  for (int s = 0; s < w->ns; ++s) {
    memset(w->delta[s], 0, sizeof(double) * vol4(&w->sp[s].g));
    w->ke_delta[s] = 0.0;
  }
  for (int stage = 1; stage <= 4; ++stage) {
    double** eval = (stage == 1) ? f_old : f_new;
    for (int s = 0; s < w->ns; ++s) {
      memset(w->rhs[s], 0, sizeof(double) * vol4(&w->sp[s].g));
      w->ke_rhs[s] = 0.0;
    }
    ok_vp_eval_rhs(w, w->rhs, eval, w->ke_rhs, ax, ay);
    for (int s = 0; s < w->ns; ++s) {
      const ok_geom* g = &w->sp[s].g;
      ok_xpby4d(w->delta[s], w->rhs[s], w_eval[stage - 1], g);
      memcpy(f_new[s], f_old[s], sizeof(double) * vol4(g)); /* copySolnData copies ghosts too */
      if (stage < 4)
        ok_xpby4d(f_new[s], w->rhs[s], w_upd[stage - 1], g);
      else
        ok_xpby4d(f_new[s], w->delta[s], w_upd[stage - 1], g);
    }
  }
  free(ax);
  free(ay);
}

/* ------------------------------------------------------------------------------------------
 * CPU timing leg: one RK4 stage done the way the reference does it: zero rhs, x/y sweep,
 * vel3/vel4 materialisation, v-BC, v sweeps, two xpby, one copy, one 4D->2D reduction -- on a periodic
 * box with a synthetic Maxwellian, single thread per call (bench.py forks one process per core, each
 * owning an independent sub-box like the reference's MPI ranks).
 * Returns seconds per stage (average over reps).
 * ------------------------------------------------------------------------------------------ */
static double ic_zero(void* c, int i1, int i2, int i3, int i4) { (void)c; (void)i1; (void)i2; (void)i3; (void)i4; return 0.0; }

double ok_time_rk4_stage_reference_style(const ok_geom* g, int nthreads, int reps) {
  (void)nthreads;
  const int64_t n1d = ok_nd(g, 0), n2d = ok_nd(g, 1), n3d = ok_nd(g, 2), n4d = ok_nd(g, 3);
  const int64_t vol = vol4(g);
  double* f = (double*)malloc(sizeof(double) * vol);
  double* fold = (double*)malloc(sizeof(double) * vol);
  double* rhs = (double*)malloc(sizeof(double) * vol);
  double* delta = (double*)calloc(vol, sizeof(double));
  double* velocities = (double*)calloc(n3d * n4d * 2, sizeof(double));
  double* vxface = (double*)calloc((n3d + 1) * n4d * 2, sizeof(double));
  double* vyface = (double*)calloc(n3d * (n4d + 1) * 2, sizeof(double));
  double* vel1 = (double*)calloc((n1d + 1) * n2d * n3d * n4d, sizeof(double));
  double* vel2 = (double*)calloc((n2d + 1) * n3d * n4d * n1d, sizeof(double));
  double* vel3 = (double*)calloc((n3d + 1) * n4d * n1d * n2d, sizeof(double));
  double* vel4 = (double*)calloc((n4d + 1) * n1d * n2d * n3d, sizeof(double));
  double* accel = (double*)calloc(n1d * n2d * 2, sizeof(double));
  double* rho = (double*)calloc(n1d * n2d, sizeof(double));
  int lo34[2] = {-g->ng, -g->ng};
  double vlo = -0.5 * g->n[2] * g->dx[2], vlo2 = -0.5 * g->n[3] * g->dx[3];
  ok_build_velocity_tables(g, lo34, vlo, vlo2, velocities, vxface, vyface);
  ok_initialize_velocity(g, velocities, vel1, vel2);
  for (int i4 = 0; i4 < n4d; ++i4)
    for (int i3 = 0; i3 < n3d; ++i3)
      for (int i2 = 0; i2 < n2d; ++i2)
        for (int i1 = 0; i1 < n1d; ++i1) {
          double vx = velocities[i3 + n3d * i4], vy = velocities[i3 + n3d * (i4 + n4d)];
          f[ok_idx(g, i1, i2, i3, i4)] =
              exp(-0.5 * (vx * vx + vy * vy)) * (1.0 + 0.1 * cos(0.3 * i1) * cos(0.2 * i2)) / 6.283185307179586;
        }
  memcpy(fold, f, sizeof(double) * vol);
  for (int64_t k = 0; k < n1d * n2d; ++k) { accel[k] = 0.01 * sin(0.1 * (double)k); accel[k + n1d * n2d] = 0.01 * cos(0.07 * (double)k); }
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int r = 0; r < reps; ++r) {
    double ax, ay;
    memset(rhs, 0, sizeof(double) * vol);                                       /* zeroSolnData */
    ok_reduce_4d_to_2d(rho, f, g, g->dx[2] * g->dx[3], -1.0);                    /* chargeDensity */
    ok_periodic_fill_4d(f, g, 1, 1);                                            /* fillAdvectionGhostCells */
    ok_advection_derivatives_4d(rhs, f, g, vel1, vel2);
    ok_set_phase_space_vel_4d(vel3, vel4, g, vxface, vyface, -1.0, 0.0, accel, &ax, &ay);
    ok_set_acceleration_bcs_4d(f, g, vel3, vel4, 1, 1, 1, 1, ic_zero, NULL);
    ok_acceleration_derivatives_4d(rhs, f, g, vel3, vel4);
    ok_xpby4d(delta, rhs, 1e-3, g);                                             /* addSolnData(delta) */
    memcpy(f, fold, sizeof(double) * vol);                                      /* copySolnData */
    ok_xpby4d(f, rhs, 1e-3, g);                                                 /* addSolnData(pred) */
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  double sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  free(f); free(fold); free(rhs); free(delta); free(velocities); free(vxface); free(vyface);
  free(vel1); free(vel2); free(vel3); free(vel4); free(accel); free(rho);
  return sec / reps;
}
