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
  double xlo[2], xhi[2];
  /* per species */
  double **velocities, **vxface, **vyface, **vel1, **vel2, **vel3, **vel4, **accel, **rho_s, **ext;
  /* field */
  double *rho, *phi, *em, *sx, *sy;
  /* RK scratch */
  double **rhs, **delta, **k[8];
  double *ke_rhs, *ke_delta, *ke_k[8];
  /* m_lambda_max[V1], [V2] as the last evalRHS left them (KineticSpecies.C:771-772): what stableDt sees
   * at the start of the next step (VPSystem.C:489-505) */
  double *last_ax, *last_ay;
  /* deck options beyond the five benchmark decks: non-periodic x / y (setPhysicalBCs,
   * KineticSpecies.H:998-1031), the "JB" boundary conditions (use_new_bcs, VPSystem.C:819-821) and a Krook layer per
   * species (KineticSpecies.C:1049-1062); cur_dt: the a_dt the integrators hand to completeRHS */
  int nonperiodic[2], use_new_bcs;
  double** krook_nu;
  double** coll;   /* pitch-angle operator parameters per species (8 doubles) or NULL */
  double** tz;     /* twilight-zone source per species {amp, me, mi, kind 1..4} or NULL (KineticSpecies.C:1077-1080) */
  double cur_dt;
};

static int64_t vol4(const ok_geom* g) { return ok_nd(g, 0) * ok_nd(g, 1) * ok_nd(g, 2) * ok_nd(g, 3); }

/* ------------------------------------------------------------------------------------------
 * ShapedRampedCosineDriver (ShapedRampedCosineDriverF.f:10-189)
 * ------------------------------------------------------------------------------------------ */
static double drv_ghat(double x, double t, const double* p, double phase, double pi, int shape_type) {
  double xwidth = p[0], omega = p[3], t0 = p[5], x_shape = p[9], lwidth = p[10], x0 = p[11], alpha = p[12], t_res = p[13];
  double ghat;
  if (shape_type == 0) {
    if (fabs(x - x0) < 0.5 * lwidth) {
      double sn = sin(pi * (x - x0) / lwidth);
      ghat = 1.0 - x_shape * (sn * sn);
    } else {
      ghat = 1.0 - x_shape;
    }
  } else {
    if (lwidth >= 0) {
      if (x <= x0) ghat = 1.0; else ghat = 1.0 - x_shape * (1.0 - exp(-(x - x0) / lwidth));
    } else {
      if (x <= x0) ghat = 1.0 - x_shape * (1.0 - exp(-(x - x0) / lwidth)); else ghat = 1.0;
    }
  }
  double tt = t - t0 - t_res;
  ghat = ghat * cos(pi * x / xwidth - omega * (t - t0) + phase - 0.5 * alpha * (tt * tt));
  return ghat;
}
static double drv_h(double y, const double* p, double pi) {
  double ywidth = p[1], shape = p[2];
  if (fabs(y) < 0.5 * ywidth) {
    double sn = sin(pi * y / ywidth);
    return 1.0 - shape * (sn * sn);
  }
  return 1.0 - shape;
}
static double drv_envelope(double t, double t0, double t_rampup, double t_hold, double t_rampdown, double E_0) {
  if ((t < t0) || (t >= t0 + t_rampup + t_hold + t_rampdown)) return 0.0;
  if (t < t0 + t_rampup) return E_0 * (0.5 + 0.5 * tanh(4.0 * (2.0 * (t - t0) / t_rampup - 1.0)));
  if (t < t0 + t_rampup + t_hold) return E_0 * (0.5 + 0.5 * tanh(4.0));
  return E_0 * (0.5 - 0.5 * tanh(4.0 * (2.0 * (t - t0 - t_rampup - t_hold) / t_rampdown - 1.0)));
}
void ok_shaped_ramped_driver(double* em_vars, double* ext_efield, int n1d, int n2d, int lo1, int lo2,
                             const double* xlo, const double* dx, int sums_into, double t,
                             const double* p, double phase, int shape_type) {
  const double one = 1.0, four = 4.0;
  const double pi = four * atan(one);
  const double E_0 = p[4], t0 = p[5], t_rampup = p[6], t_hold = p[7], t_rampdown = p[8];
  if ((t < (t0 + t_rampup + t_hold + t_rampdown)) && t >= t0) {
    double envel = drv_envelope(t, t0, t_rampup, t_hold, t_rampdown, E_0);
    for (int i2 = 0; i2 < n2d; ++i2)
      for (int i1 = 0; i1 < n1d; ++i1) {
        double xcoord = xlo[0] + dx[0] * (0.5 + (lo1 + i1));
        double ycoord = xlo[1] + dx[1] * (0.5 + (lo2 + i2));
        double g = drv_ghat(xcoord, t, p, phase, pi, shape_type);
        double h = drv_h(ycoord, p, pi);
        ext_efield[i1 + (int64_t)n1d * i2] = ext_efield[i1 + (int64_t)n1d * i2] + envel * h * g;
        if (sums_into == 2) em_vars[i1 + (int64_t)n1d * i2] = em_vars[i1 + (int64_t)n1d * i2] + envel * h * g;
      }
  }
}

ok_vp_work* ok_vp_work_create(int ns, const ok_species* sp, const double* xlo, const double* xhi) {
  ok_vp_work* w = (ok_vp_work*)calloc(1, sizeof(*w));
  w->ns = ns;
  w->sp = (ok_species*)malloc(sizeof(ok_species) * ns);
  memcpy(w->sp, sp, sizeof(ok_species) * ns);
  for (int k = 0; k < 2; ++k) { w->xlo[k] = xlo[k]; w->xhi[k] = xhi[k]; }
#define PP(name) w->name = (double**)calloc(ns, sizeof(double*))
  PP(velocities); PP(vxface); PP(vyface); PP(vel1); PP(vel2); PP(vel3); PP(vel4); PP(accel); PP(rho_s); PP(ext); PP(krook_nu); PP(coll); PP(tz);
  PP(rhs); PP(delta);
  for (int i = 0; i < 8; ++i) { PP(k[i]); w->ke_k[i] = (double*)calloc(ns, sizeof(double)); }
#undef PP
  w->ke_rhs = (double*)calloc(ns, sizeof(double));
  w->ke_delta = (double*)calloc(ns, sizeof(double));
  w->last_ax = (double*)calloc(ns, sizeof(double));
  w->last_ay = (double*)calloc(ns, sizeof(double));
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
    w->ext[s] = (double*)calloc(n1d * n2d * 2, sizeof(double));
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
  ok_poisson_symbols(g0->n[0], g0->n[1], xhi[0] - xlo[0], xhi[1] - xlo[1], g0->order, w->sx, w->sy);
  return w;
}

void ok_vp_last_accel_max(const ok_vp_work* w, double* axmax, double* aymax) {
  for (int s = 0; s < w->ns; ++s) {
    axmax[s] = w->last_ax[s];
    aymax[s] = w->last_ay[s];
  }
}

void ok_vp_work_destroy(ok_vp_work* w) {
  if (!w) return;
  for (int s = 0; s < w->ns; ++s) {
    free(w->velocities[s]); free(w->vxface[s]); free(w->vyface[s]); free(w->vel1[s]); free(w->vel2[s]);
    free(w->vel3[s]); free(w->vel4[s]); free(w->accel[s]); free(w->ext[s]); free(w->rho_s[s]); free(w->rhs[s]);
    free(w->delta[s]);
    for (int i = 0; i < 8; ++i) free(w->k[i][s]);
  }
  for (int i = 0; i < 8; ++i) { free(w->k[i]); free(w->ke_k[i]); }
  free(w->last_ax); free(w->last_ay);
  for (int s = 0; s < w->ns; ++s) { free(w->krook_nu[s]); free(w->coll[s]); free(w->tz[s]); }
  free(w->krook_nu);
  free(w->coll);
  free(w->tz);
  free(w->velocities); free(w->vxface); free(w->vyface); free(w->vel1); free(w->vel2); free(w->vel3);
  free(w->vel4); free(w->accel); free(w->ext); free(w->rho_s); free(w->rhs); free(w->delta); free(w->ke_rhs);
  free(w->ke_delta); free(w->rho); free(w->phi); free(w->em); free(w->sx); free(w->sy); free(w->sp);
  free(w);
}

const double* ok_vp_em_vars(const ok_vp_work* w) { return w->em; }
const double* ok_vp_rho(const ok_vp_work* w) { return w->rho; }

void ok_vp_set_options(ok_vp_work* w, int nonperiodic_x, int nonperiodic_y, int use_new_bcs) {
  w->nonperiodic[0] = nonperiodic_x;
  w->nonperiodic[1] = nonperiodic_y;
  w->use_new_bcs = use_new_bcs;
}
void ok_vp_set_dt(ok_vp_work* w, double dt) { w->cur_dt = dt; }
/* nu: (n1d,n2d) of species s (KrookLayer::initialize, KrookLayer.C:54-160), copied; NULL removes the layer */
void ok_vp_set_krook(ok_vp_work* w, int s, const double* nu) {
  const int64_t pl = ok_nd(&w->sp[s].g, 0) * ok_nd(&w->sp[s].g, 1);
  free(w->krook_nu[s]);
  w->krook_nu[s] = NULL;
  if (nu) {
    w->krook_nu[s] = (double*)malloc(sizeof(double) * pl);
    memcpy(w->krook_nu[s], nu, sizeof(double) * pl);
  }
}
void ok_vp_set_pitch_angle(ok_vp_work* w, int s, const double* p) {
  free(w->coll[s]);
  w->coll[s] = NULL;
  if (p) {
    w->coll[s] = (double*)malloc(sizeof(double) * 8);
    memcpy(w->coll[s], p, sizeof(double) * 8);
  }
}
void ok_vp_set_trig_tz(ok_vp_work* w, int s, int on, double amp, double electron_mass, double ion_mass) {
  free(w->tz[s]);
  w->tz[s] = NULL;
  if (on) {
    w->tz[s] = (double*)malloc(4 * sizeof(double));
    w->tz[s][0] = amp;
    w->tz[s][1] = electron_mass;
    w->tz[s][2] = ion_mass;
    w->tz[s][3] = (double)on;
  }
}
/* fillAdvectionGhostCells on one rank (KineticSpecies.H:404-412, 998-1031): physical boundary conditions of the
 * non-periodic directions, then the periodic wrap of the periodic ones */
static void vp_fill_advection_ghosts(ok_vp_work* w, int s, double* f) {
  const ok_species* sp = &w->sp[s];
  const int xper = !w->nonperiodic[0], yper = !w->nonperiodic[1];
  if (!xper || !yper) {
    if (w->use_new_bcs)
      ok_set_advection_bcs_4d_jb(f, &sp->g, w->vel1[s], w->vel2[s], 1, 1, 1, 1, xper, yper, sp->ic, sp->ic_ctx);
    else
      ok_set_advection_bcs_4d(f, &sp->g, w->vel1[s], w->vel2[s], 1, 1, 1, 1, xper, yper, sp->ic, sp->ic_ctx);
  }
  ok_periodic_fill_4d(f, &sp->g, xper, yper);
}
static void vp_set_acceleration_bcs(ok_vp_work* w, int s, double* f) {
  const ok_species* sp = &w->sp[s];
  if (w->use_new_bcs)
    ok_set_acceleration_bcs_4d_jb(f, &sp->g, w->vel3[s], w->vel4[s], 1, 1, 1, 1, sp->ic, sp->ic_ctx);
  else
    ok_set_acceleration_bcs_4d(f, &sp->g, w->vel3[s], w->vel4[s], 1, 1, 1, 1, sp->ic, sp->ic_ctx);
}

/* VPSystem::evalRHS on one rank */
void ok_vp_eval_rhs(ok_vp_work* w, double** rhs, double** f, double time, double* ke_e_dot, double* axmax, double* aymax) {
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
    vp_fill_advection_ghosts(w, s, f[s]);
    ok_advection_derivatives_4d(rhs[s], f[s], g, w->vel1[s], w->vel2[s]);
    /* 4. acceleration (KineticSpecies.C:697-774): expansion, drivers, *= q/m, face accelerations */
    for (int64_t k = 0; k < 2 * pl; ++k) w->accel[s][k] = w->em[k];
    if (sp->has_driver) {
      for (int64_t k = 0; k < 2 * pl; ++k) w->ext[s][k] = 0.0;
      ok_shaped_ramped_driver(w->accel[s], w->ext[s], (int)n1d, (int)n2d, -ng, -ng, w->xlo, g->dx, 2, time, sp->driver,
                              sp->driver_phase, sp->driver_shape_type);
    }
    double normalization = sp->charge / sp->mass;
    for (int64_t k = 0; k < 2 * pl; ++k) w->accel[s][k] *= normalization;
    ok_set_phase_space_vel_4d(w->vel3[s], w->vel4[s], g, w->vxface[s], w->vyface[s], normalization,
                              sp->bz_const, w->accel[s], &axmax[s], &aymax[s]);
    w->last_ax[s] = axmax[s];
    w->last_ay[s] = aymax[s];
    /* 5. v-boundary fill, acceleration derivatives, completeRHS */
    vp_set_acceleration_bcs(w, s, f[s]);
    ok_acceleration_derivatives_4d(rhs[s], f[s], g, w->vel3[s], w->vel4[s]);
    /* completeRHS (KineticSpecies.C:1024-1093): collision operators, Krook layer, then the driver's energy input rate */
    if (w->coll[s]) {
      const double* p = w->coll[s];
      double* iv = (double*)malloc(sizeof(double) * 3 * pl);
      ok_pitch_angle_fields(iv, iv + pl, iv + 2 * pl, f[s], g, w->velocities[s]);
      ok_append_pitch_angle_collision(rhs[s], f[s], g, w->velocities[s], iv, iv + pl, iv + 2 * pl, sp->vlo, sp->vhi, p, p + 2,
                                      p[4], p[6], (int)p[7]);
      free(iv);
    }
    if (w->krook_nu[s]) ok_append_krook(rhs[s], f[s], g, w->krook_nu[s], w->cur_dt, sp->ic, sp->ic_ctx);
    if (w->tz[s]) {
      /* the twilight-zone source (KineticSpecies.C:1077-1080) */
      const int lo[2] = {-g->ng, -g->ng};
      const int on = (int)w->tz[s][3];
      if (on >= 3)
        ok_set_two_species_trig_tz_source(rhs[s], g, lo, w->xlo, g->dx, time, w->velocities[s], w->tz[s], on - 3);
      else if (on == 2)
        ok_set_electron_trig_tz_source(rhs[s], g, lo, w->xlo, g->dx, time, w->velocities[s], w->tz[s][0]);
      else
        ok_set_trig_tz_source(rhs[s], g, lo, w->xlo, g->dx, time, w->velocities[s], w->tz[s][0]);
    }
    if (sp->has_driver && ke_e_dot)
      ke_e_dot[s] = ok_compute_ke_e_dot(g, f[s], sp->charge, w->velocities[s], w->ext[s], 0.0);
  }
}

/* KineticSpecies::accumulateSequencesCommon (KineticSpecies.C:2052-2097): the kinetic-energy flux of every species'
 * state f[s] through the eight phase-space boundaries, out[8 s + 2 dir + side].  As there: ghosts refreshed
 * (Simulation.C:119 updateGhosts, setPhysicalBCs: periodic here), advection fluxes, then the velocity-boundary fill
 * and the acceleration fluxes with the face accelerations the LAST evalRHS left in vel3 / vel4 (the last stage of the
 * step just taken), then computekeflux on a box that touches every boundary.  f's ghost cells are rewritten. */
void ok_vp_ke_flux_history(ok_vp_work* w, double** f, double* out) {
  for (int s = 0; s < w->ns; ++s) {
    const ok_species* sp = &w->sp[s];
    const ok_geom* g = &sp->g;
    const int64_t n1d = ok_nd(g, 0), n2d = ok_nd(g, 1), n3d = ok_nd(g, 2), n4d = ok_nd(g, 3);
    const int64_t len[4] = {(n1d + 1) * n2d * n3d * n4d, (n2d + 1) * n3d * n4d * n1d, (n3d + 1) * n4d * n1d * n2d,
                            (n4d + 1) * n1d * n2d * n3d};
    const double* vel[4] = {w->vel1[s], w->vel2[s], w->vel3[s], w->vel4[s]};
    double *face[4], *flux[4];
    for (int d = 0; d < 4; ++d) {
      face[d] = (double*)calloc(len[d], sizeof(double));
      flux[d] = (double*)calloc(len[d], sizeof(double));
    }
    vp_fill_advection_ghosts(w, s, f[s]);
    ok_face_fluxes_4d(flux[0], face[0], f[s], g, vel[0], 0);
    ok_face_fluxes_4d(flux[1], face[1], f[s], g, vel[1], 1);
    vp_set_acceleration_bcs(w, s, f[s]);
    ok_face_fluxes_4d(flux[2], face[2], f[s], g, vel[2], 2);
    ok_face_fluxes_4d(flux[3], face[3], f[s], g, vel[3], 3);
    for (int dir = 0; dir < 4; ++dir)
      for (int side = 0; side < 2; ++side)
        out[8 * s + 2 * dir + side] = ok_compute_ke_flux(g, flux[0], flux[1], flux[2], flux[3], w->velocities[s], w->vxface[s],
                                                         w->vyface[s], dir, side, sp->mass);
    for (int d = 0; d < 4; ++d) {
      free(face[d]);
      free(flux[d]);
    }
  }
}

/* RK4Integrator::advance with stageAdvance (RK4Integrator.H:66-171) */
void ok_vp_rk4_step(ok_vp_work* w, double** f_new, double** f_old, double time, double dt, double* ke) {
  static const double THIRD = 1.0 / 3.0;
  double dtOn2 = 0.5 * dt, dtOn3 = THIRD * dt, dtOn6 = 0.5 * dtOn3;
  const double w_eval[4] = {dtOn6, dtOn3, dtOn3, dtOn6};
  const double w_upd[4] = {dtOn2, dtOn2, dt, 1.0};
  const double t_stage[4] = {time, time + dtOn2, time + dtOn2, time + dt};
  w->cur_dt = dt;
  double* ax = (double*)calloc(w->ns, sizeof(double));
  double* ay = (double*)calloc(w->ns, sizeof(double));
  double* ke_old = (double*)calloc(w->ns, sizeof(double));
  for (int s = 0; s < w->ns; ++s) {
    memset(w->delta[s], 0, sizeof(double) * vol4(&w->sp[s].g));
    w->ke_delta[s] = 0.0;
    ke_old[s] = ke ? ke[s] : 0.0;
  }
  for (int stage = 1; stage <= 4; ++stage) {
    double** eval = (stage == 1) ? f_old : f_new;
    for (int s = 0; s < w->ns; ++s) {
      memset(w->rhs[s], 0, sizeof(double) * vol4(&w->sp[s].g));
      w->ke_rhs[s] = 0.0;
    }
    ok_vp_eval_rhs(w, w->rhs, eval, t_stage[stage - 1], w->ke_rhs, ax, ay);
    for (int s = 0; s < w->ns; ++s) {
      const ok_geom* g = &w->sp[s].g;
      ok_xpby4d(w->delta[s], w->rhs[s], w_eval[stage - 1], g);
      w->ke_delta[s] += w->ke_rhs[s] * w_eval[stage - 1];
      memcpy(f_new[s], f_old[s], sizeof(double) * vol4(g)); /* copySolnData copies ghosts too */
      double kn = ke_old[s];
      if (stage < 4) {
        ok_xpby4d(f_new[s], w->rhs[s], w_upd[stage - 1], g);
        kn += w->ke_rhs[s] * w_upd[stage - 1];
      } else {
        ok_xpby4d(f_new[s], w->delta[s], w_upd[stage - 1], g);
        kn += w->ke_delta[s] * w_upd[stage - 1];
      }
      if (ke) ke[s] = kn;
    }
  }
  free(ax); free(ay); free(ke_old);
}

/* RK6Integrator::advance (RK6Integrator.H:69-133) */
void ok_vp_rk6_step(ok_vp_work* w, double** f_new, double** f_old, double time, double dt, double* ke) {
  static const double A[8][8] =
      {{0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
       {1.0/9.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
       {1.0/24.0, 1.0/8.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0},
       {1.0/6.0, -1.0/2.0, 2.0/3.0, 0.0, 0.0, 0.0, 0.0, 0.0},
       {935.0/2536.0, -2781.0/2536.0, 309.0/317.0, 321.0/1268.0, 0.0, 0.0, 0.0, 0.0},
       {-12710.0/951.0, 8287.0/317.0, -40.0/317.0, -6335.0/317.0, 8.0, 0.0, 0.0, 0.0},
       {5840285.0/3104064.0, -7019.0/2536.0, -52213.0/86224.0, 1278709.0/517344.0, -433.0/2448.0, 33.0/1088.0, 0.0, 0.0},
       {-5101675.0/1767592.0, 112077.0/25994.0, 334875.0/441898.0, -973617.0/883796.0, -1421.0/1394.0, 333.0/5576.0, 36.0/41.0, 0.0}};
  static const double b[8] = {41.0/840.0, 0.0, 9.0/35.0, 9.0/280.0, 34.0/105.0, 9.0/280.0, 9.0/35.0, 41/840.0};
  static const double c[8] = {0.0, 1.0/9.0, 1.0/6.0, 1.0/3.0, 1.0/2.0, 2.0/3.0, 5.0/6.0, 1.0};
  w->cur_dt = dt;
  double* ax = (double*)calloc(w->ns, sizeof(double));
  double* ay = (double*)calloc(w->ns, sizeof(double));
  double* ke_old = (double*)calloc(w->ns, sizeof(double));
  for (int s = 0; s < w->ns; ++s) {
    ke_old[s] = ke ? ke[s] : 0.0;
    for (int i = 0; i < 8; ++i) {
      if (!w->k[i][s]) w->k[i][s] = (double*)malloc(sizeof(double) * vol4(&w->sp[s].g));
      memset(w->k[i][s], 0, sizeof(double) * vol4(&w->sp[s].g));
      w->ke_k[i][s] = 0.0;
    }
    memcpy(f_new[s], f_old[s], sizeof(double) * vol4(&w->sp[s].g));
  }
  ok_vp_eval_rhs(w, w->k[0], f_new, time + c[0] * dt, w->ke_k[0], ax, ay);
  for (int i = 1; i < 8; ++i) {
    for (int s = 0; s < w->ns; ++s) {
      const ok_geom* g = &w->sp[s].g;
      memcpy(f_new[s], f_old[s], sizeof(double) * vol4(g));
      for (int j = 0; j < i; ++j) ok_xpby4d(f_new[s], w->k[j][s], dt * A[i][j], g);
    }
    ok_vp_eval_rhs(w, w->k[i], f_new, time + c[i] * dt, w->ke_k[i], ax, ay);
  }
  for (int s = 0; s < w->ns; ++s) {
    const ok_geom* g = &w->sp[s].g;
    memcpy(f_new[s], f_old[s], sizeof(double) * vol4(g));
    double kn = ke_old[s];
    for (int i = 0; i < 8; ++i) {
      ok_xpby4d(f_new[s], w->k[i][s], dt * b[i], g);
      kn += w->ke_k[i][s] * (dt * b[i]);
    }
    if (ke) ke[s] = kn;
  }
  free(ax); free(ay); free(ke_old);
}

/* KineticSpecies::computeDt (KineticSpecies.C:647-694) and VPSystem::stableDt (VPSystem.C:489-505) */
double ok_vp_stable_dt(const ok_vp_work* w, const double* axmax, const double* aymax, int rk_order) {
  const double pi = 4.0 * atan(1.0);
  double dt_stable = 1.7976931348623157e308;
  for (int s = 0; s < w->ns; ++s) {
    const ok_geom* g = &w->sp[s].g;
    const ok_species* sp = &w->sp[s];
    double lam[4];
    /* KineticSpecies.C:1547-1555: note the upper bound is vhi + dv/2 as written in the reference */
    double vlo = sp->vlo[0] + 0.5 * (sp->vhi[0] - sp->vlo[0]) / g->n[2];
    double vhi = sp->vhi[0] + 0.5 * (sp->vhi[0] - sp->vlo[0]) / g->n[2];
    lam[0] = fmax(fabs(vlo), fabs(vhi));
    vlo = sp->vlo[1] + 0.5 * (sp->vhi[1] - sp->vlo[1]) / g->n[3];
    vhi = sp->vhi[1] + 0.5 * (sp->vhi[1] - sp->vlo[1]) / g->n[3];
    lam[1] = fmax(fabs(vlo), fabs(vhi));
    lam[2] = axmax[s];
    lam[3] = aymax[s];
    double imLam = 0.0, reLam = 0.0;
    for (int d = 0; d < 4; ++d) imLam += pi * lam[d] / g->dx[d];
    if (w->coll[s]) reLam = fabs(ok_pitch_angle_real_lam(g, w->coll[s][6], w->coll[s][5], w->coll[s][4]));  /* KineticSpecies.C:666-672 */
    double alpha = rk_order == 4 ? 2.6 : 4.95, beta = rk_order == 4 ? 2.6 : 3.168;
    double ddt = sqrt(1.0 / (reLam * reLam / (alpha * alpha) + imLam * imLam / (beta * beta)));
    if (ddt < dt_stable) dt_stable = ddt;
  }
  return dt_stable;
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
