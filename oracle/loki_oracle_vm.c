/*
 * loki_oracle_vm.c -- CPU ORACLE (test infrastructure only; see loki_oracle.h).
 *
 * Single-rank restatement of the reference's Vlasov-Maxwell stage sequencing:
 * VMSystem::evalRHS (VMSystem.C:407-549), KineticSpecies::currentDensity (KineticSpecies.C:853-895,
 * schedules :1895-1950), KineticSpecies::computeAcceleration, Maxwell flavour (KineticSpecies.C:777-850),
 * Maxwell::fillGhostCells / setPhysicalBCs (Maxwell.H:371-381, Maxwell.C:1239-1288; x and y periodic, so
 * the Fortran BC routines are no-ops and only communicatePeriodicBoundaries acts), Maxwell::evalRHS
 * (Maxwell.C:562-623), Maxwell::addData / copySolnData (Maxwell.C:299-353), Maxwell::computeDt
 * (Maxwell.H:199-204), VMSystem::stableDt (VMSystem.C:563-581), RK4Integrator (RK4Integrator.H:66-171),
 * and the two field initial conditions SimpleEMICF.f:10-47, SimpleVELICF.f:10-40.
 * No external E-field drivers, antennae or particles (none of the Maxwell decks in scope has them).
 *
 * Parity pinning: every Fortran kernel this file sequences is pinned bit-for-bit against the
 * transliterated reference Fortran (tests/test_oracle_pin.py); the sequencing itself restates C++ and
 * has no in-repo golden output ("parity unpinned" for the C++ call order, like loki_oracle_vp.c).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "loki_oracle.h"

struct ok_vm_work {
  int ns;
  ok_species* sp;
  double xlo[2], xhi[2];
  double light_speed, av_weak, av_strong;
  /* per species */
  double **velocities, **vxface, **vyface, **vel1, **vel2, **vel3, **vel4, **em_s, **vz_s;
  double **J4x, **J4y, **J4z, **Jx_s, **Jy_s, **Jz_s;
  /* net currents */
  double *Jx, *Jy, *Jz;
  /* RK scratch: rhs and delta of the whole VMState */
  double **rhs, **delta, *rhs_em, *delta_em, **rhs_vz, **delta_vz;
  double *last_ax, *last_ay;
};

static int64_t vol4(const ok_geom* g) { return ok_nd(g, 0) * ok_nd(g, 1) * ok_nd(g, 2) * ok_nd(g, 3); }

/* SimpleEMICF.f:10-47: u(:,:,field..field+2) += amp * cos(kx x + ky y + phi) over the whole data box;
 * field = 1 (E) or 4 (B), 1-based like the Fortran */
void ok_simple_em_ic(double* em, int n1, int n2, int ng, const double* xlo, const double* dx, int field,
                     double xamp, double yamp, double zamp, double kx, double ky, double phi) {
  const int64_t n1d = n1 + 2 * ng, n2d = n2 + 2 * ng, pl = n1d * n2d;
  for (int j = 0; j < n2d; ++j) {
    int i2 = j - ng;
    double x2 = xlo[1] + (i2 + 0.5) * dx[1];
    for (int i = 0; i < n1d; ++i) {
      int i1 = i - ng;
      double x1 = xlo[0] + (i1 + 0.5) * dx[0];
      double env = cos(kx * x1 + ky * x2 + phi);
      int64_t o = i + n1d * j;
      em[o + pl * (field - 1)] = em[o + pl * (field - 1)] + xamp * env;
      em[o + pl * (field)] = em[o + pl * (field)] + yamp * env;
      em[o + pl * (field + 1)] = em[o + pl * (field + 1)] + zamp * env;
    }
  }
}
/* SimpleVELICF.f:10-40 */
void ok_simple_vel_ic(double* vz, int n1, int n2, int ng, const double* xlo, const double* dx, double amp,
                      double kx, double ky, double phi) {
  const int64_t n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  for (int j = 0; j < n2d; ++j) {
    double x2 = xlo[1] + ((j - ng) + 0.5) * dx[1];
    for (int i = 0; i < n1d; ++i) {
      double x1 = xlo[0] + ((i - ng) + 0.5) * dx[0];
      vz[i + n1d * j] = vz[i + n1d * j] + amp * cos(kx * x1 + ky * x2 + phi);
    }
  }
}

ok_vm_work* ok_vm_work_create(int ns, const ok_species* sp, const double* xlo, const double* xhi,
                              double light_speed, double av_weak, double av_strong) {
  ok_vm_work* w = (ok_vm_work*)calloc(1, sizeof(*w));
  w->ns = ns;
  w->sp = (ok_species*)malloc(sizeof(ok_species) * ns);
  memcpy(w->sp, sp, sizeof(ok_species) * ns);
  for (int k = 0; k < 2; ++k) { w->xlo[k] = xlo[k]; w->xhi[k] = xhi[k]; }
  w->light_speed = light_speed; w->av_weak = av_weak; w->av_strong = av_strong;
#define PP(name) w->name = (double**)calloc(ns, sizeof(double*))
  PP(velocities); PP(vxface); PP(vyface); PP(vel1); PP(vel2); PP(vel3); PP(vel4); PP(em_s); PP(vz_s);
  PP(J4x); PP(J4y); PP(J4z); PP(Jx_s); PP(Jy_s); PP(Jz_s); PP(rhs); PP(delta); PP(rhs_vz); PP(delta_vz);
#undef PP
  w->last_ax = (double*)calloc(ns, sizeof(double));
  w->last_ay = (double*)calloc(ns, sizeof(double));
  const ok_geom* g0 = &sp[0].g;
  const int64_t n1d = ok_nd(g0, 0), n2d = ok_nd(g0, 1), pl = n1d * n2d;
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
    w->em_s[s] = (double*)calloc(pl * 6, sizeof(double));
    w->vz_s[s] = (double*)calloc(pl, sizeof(double));
    w->J4x[s] = (double*)calloc(vol4(g), sizeof(double));
    w->J4y[s] = (double*)calloc(vol4(g), sizeof(double));
    w->J4z[s] = (double*)calloc(vol4(g), sizeof(double));
    w->Jx_s[s] = (double*)calloc(pl, sizeof(double));
    w->Jy_s[s] = (double*)calloc(pl, sizeof(double));
    w->Jz_s[s] = (double*)calloc(pl, sizeof(double));
    w->rhs[s] = (double*)calloc(vol4(g), sizeof(double));
    w->delta[s] = (double*)calloc(vol4(g), sizeof(double));
    w->rhs_vz[s] = (double*)calloc(pl, sizeof(double));
    w->delta_vz[s] = (double*)calloc(pl, sizeof(double));
    int lo34[2] = {-g->ng, -g->ng};
    ok_build_velocity_tables(g, lo34, sp[s].vlo[0], sp[s].vlo[1], w->velocities[s], w->vxface[s], w->vyface[s]);
    ok_initialize_velocity(g, w->velocities[s], w->vel1[s], w->vel2[s]);
  }
  w->Jx = (double*)calloc(pl, sizeof(double));
  w->Jy = (double*)calloc(pl, sizeof(double));
  w->Jz = (double*)calloc(pl, sizeof(double));
  w->rhs_em = (double*)calloc(pl * 6, sizeof(double));
  w->delta_em = (double*)calloc(pl * 6, sizeof(double));
  return w;
}

void ok_vm_work_destroy(ok_vm_work* w) {
  if (!w) return;
  for (int s = 0; s < w->ns; ++s) {
    free(w->velocities[s]); free(w->vxface[s]); free(w->vyface[s]); free(w->vel1[s]); free(w->vel2[s]);
    free(w->vel3[s]); free(w->vel4[s]); free(w->em_s[s]); free(w->vz_s[s]); free(w->J4x[s]); free(w->J4y[s]);
    free(w->J4z[s]); free(w->Jx_s[s]); free(w->Jy_s[s]); free(w->Jz_s[s]); free(w->rhs[s]); free(w->delta[s]);
    free(w->rhs_vz[s]); free(w->delta_vz[s]);
  }
  free(w->velocities); free(w->vxface); free(w->vyface); free(w->vel1); free(w->vel2); free(w->vel3);
  free(w->vel4); free(w->em_s); free(w->vz_s); free(w->J4x); free(w->J4y); free(w->J4z); free(w->Jx_s);
  free(w->Jy_s); free(w->Jz_s); free(w->rhs); free(w->delta); free(w->rhs_vz); free(w->delta_vz);
  free(w->Jx); free(w->Jy); free(w->Jz); free(w->rhs_em); free(w->delta_em); free(w->last_ax); free(w->last_ay);
  free(w->sp);
  free(w);
}

const double* ok_vm_net_current(const ok_vm_work* w, int comp) { return comp == 0 ? w->Jx : (comp == 1 ? w->Jy : w->Jz); }

/* VMSystem::evalRHS on one rank.  f[s], em (n1d,n2d,6), vz[s] (n1d,n2d) are the state being evaluated
 * (their ghost cells are refreshed exactly like the reference does); rhs_* receive the derivative. */
void ok_vm_eval_rhs(ok_vm_work* w, double** rhs, double* rhs_em, double** rhs_vz, double** f, double* em,
                    double** vz, double time, double* axmax, double* aymax) {
  (void)time;
  const ok_geom* g0 = &w->sp[0].g;
  const int ng = g0->ng, n1 = g0->n[0], n2 = g0->n[1];
  const int64_t n1d = ok_nd(g0, 0), n2d = ok_nd(g0, 1), pl = n1d * n2d;
  /* 1. current density of every species (VMSystem.C:432-440): vz expansion, computecurrents,
   *    three velocity reductions with measure dvx*dvy and weight q */
  for (int s = 0; s < w->ns; ++s) {
    const ok_geom* g = &w->sp[s].g;
    memcpy(w->vz_s[s], vz[s], sizeof(double) * pl);          /* m_vz = 0; expansion (whole data box) */
    memset(w->J4x[s], 0, sizeof(double) * vol4(g));
    memset(w->J4y[s], 0, sizeof(double) * vol4(g));
    memset(w->J4z[s], 0, sizeof(double) * vol4(g));
    ok_compute_currents(g, w->velocities[s], f[s], w->vz_s[s], w->J4x[s], w->J4y[s], w->J4z[s]);
    ok_reduce_4d_to_2d(w->Jx_s[s], w->J4x[s], g, g->dx[2] * g->dx[3], w->sp[s].charge);
    ok_reduce_4d_to_2d(w->Jy_s[s], w->J4y[s], g, g->dx[2] * g->dx[3], w->sp[s].charge);
    ok_reduce_4d_to_2d(w->Jz_s[s], w->J4z[s], g, g->dx[2] * g->dx[3], w->sp[s].charge);
  }
  /* 2. Maxwell::fillGhostCells(false) (periodic: wrap em_vars and every vz), net currents */
  ok_periodic_fill_2d(em, n1, n2, ng, 6, 1, 1);
  for (int s = 0; s < w->ns; ++s) ok_periodic_fill_2d(vz[s], n1, n2, ng, 1, 1, 1);
  for (int64_t k = 0; k < pl; ++k) { w->Jx[k] = 0.0; w->Jy[k] = 0.0; w->Jz[k] = 0.0; }
  for (int s = 0; s < w->ns; ++s)
    for (int64_t k = 0; k < pl; ++k) {
      w->Jx[k] += w->Jx_s[s][k];
      w->Jy[k] += w->Jy_s[s][k];
      w->Jz[k] += w->Jz_s[s][k];
    }
  /* 3. advection (VMSystem.C:483-494) */
  for (int s = 0; s < w->ns; ++s) {
    const ok_geom* g = &w->sp[s].g;
    ok_periodic_fill_4d(f[s], g, 1, 1);
    ok_advection_derivatives_4d(rhs[s], f[s], g, w->vel1[s], w->vel2[s]);
  }
  /* 4. acceleration (KineticSpecies.C:777-850): EM expansion, Lorentz force on the v faces.  The
   *    species-local m_vz is the one expanded in step 1. */
  for (int s = 0; s < w->ns; ++s) {
    const ok_species* sp = &w->sp[s];
    const ok_geom* g = &sp->g;
    memcpy(w->em_s[s], em, sizeof(double) * pl * 6);
    double normalization = sp->charge / sp->mass;
    ok_set_phase_space_vel_maxwell_4d(w->vel3[s], w->vel4[s], g, w->vxface[s], w->vyface[s], normalization,
                                      sp->bz_const, w->em_s[s], w->vz_s[s], &axmax[s], &aymax[s]);
    w->last_ax[s] = axmax[s];
    w->last_ay[s] = aymax[s];
  }
  /* 5. v-boundary fill + acceleration derivatives (VMSystem.C:513-529); completeRHS adds nothing */
  for (int s = 0; s < w->ns; ++s) {
    const ok_species* sp = &w->sp[s];
    const ok_geom* g = &sp->g;
    ok_set_acceleration_bcs_4d(f[s], g, w->vel3[s], w->vel4[s], 1, 1, 1, 1, sp->ic, sp->ic_ctx);
    ok_acceleration_derivatives_4d(rhs[s], f[s], g, w->vel3[s], w->vel4[s]);
  }
  /* 6. Maxwell::evalRHS (Maxwell.C:562-623) */
  ok_maxwell_eval_rhs(rhs_em, em, w->Jx, w->Jy, w->Jz, n1, n2, ng, g0->order, g0->dx, w->light_speed, w->av_weak,
                      w->av_strong);
  for (int s = 0; s < w->ns; ++s)
    ok_maxwell_eval_vz_rhs(rhs_vz[s], em, w->sp[s].charge / w->sp[s].mass, n1, n2, ng);
}

/* RK4Integrator::advance over a VMState (kinetic species + em_vars + vz per species) */
void ok_vm_rk4_step(ok_vm_work* w, double** f_new, double** f_old, double* em_new, double* em_old, double** vz_new,
                    double** vz_old, double time, double dt) {
  static const double THIRD = 1.0 / 3.0;
  double dtOn2 = 0.5 * dt, dtOn3 = THIRD * dt, dtOn6 = 0.5 * dtOn3;
  const double w_eval[4] = {dtOn6, dtOn3, dtOn3, dtOn6};
  const double w_upd[4] = {dtOn2, dtOn2, dt, 1.0};
  const double t_stage[4] = {time, time + dtOn2, time + dtOn2, time + dt};
  const ok_geom* g0 = &w->sp[0].g;
  const int ng = g0->ng, n1 = g0->n[0], n2 = g0->n[1];
  const int64_t pl = ok_nd(g0, 0) * ok_nd(g0, 1);
  double* ax = (double*)calloc(w->ns, sizeof(double));
  double* ay = (double*)calloc(w->ns, sizeof(double));
  for (int s = 0; s < w->ns; ++s) {
    memset(w->delta[s], 0, sizeof(double) * vol4(&w->sp[s].g));
    memset(w->delta_vz[s], 0, sizeof(double) * pl);
  }
  memset(w->delta_em, 0, sizeof(double) * pl * 6);
  for (int stage = 1; stage <= 4; ++stage) {
    double** ef = (stage == 1) ? f_old : f_new;
    double* eem = (stage == 1) ? em_old : em_new;
    double** evz = (stage == 1) ? vz_old : vz_new;
    for (int s = 0; s < w->ns; ++s) {
      memset(w->rhs[s], 0, sizeof(double) * vol4(&w->sp[s].g));
      memset(w->rhs_vz[s], 0, sizeof(double) * pl);
    }
    memset(w->rhs_em, 0, sizeof(double) * pl * 6);
    ok_vm_eval_rhs(w, w->rhs, w->rhs_em, w->rhs_vz, ef, eem, evz, t_stage[stage - 1], ax, ay);
    const double we = w_eval[stage - 1], wu = w_upd[stage - 1];
    for (int s = 0; s < w->ns; ++s) {
      const ok_geom* g = &w->sp[s].g;
      ok_xpby4d(w->delta[s], w->rhs[s], we, g);
      ok_xpby2d(w->delta_vz[s], w->rhs_vz[s], we, n1, n2, ng, 1);
    }
    ok_xpby2d(w->delta_em, w->rhs_em, we, n1, n2, ng, 6);
    for (int s = 0; s < w->ns; ++s) {
      const ok_geom* g = &w->sp[s].g;
      memcpy(f_new[s], f_old[s], sizeof(double) * vol4(g));
      memcpy(vz_new[s], vz_old[s], sizeof(double) * pl);
      ok_xpby4d(f_new[s], (stage < 4) ? w->rhs[s] : w->delta[s], wu, g);
      ok_xpby2d(vz_new[s], (stage < 4) ? w->rhs_vz[s] : w->delta_vz[s], wu, n1, n2, ng, 1);
    }
    memcpy(em_new, em_old, sizeof(double) * pl * 6);
    ok_xpby2d(em_new, (stage < 4) ? w->rhs_em : w->delta_em, wu, n1, n2, ng, 6);
  }
  free(ax); free(ay);
}

/* RK6Integrator::advance (RK6Integrator.H:69-133) over the whole VMState (distributions, em_vars, vz) */
void ok_vm_rk6_step(ok_vm_work* w, double** f_new, double** f_old, double* em_new, double* em_old, double** vz_new,
                    double** vz_old, double time, double dt) {
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
  const ok_geom* g0 = &w->sp[0].g;
  const int ng = g0->ng, n1 = g0->n[0], n2 = g0->n[1], ns = w->ns;
  const int64_t pl = ok_nd(g0, 0) * ok_nd(g0, 1);
  double* ax = (double*)calloc(ns, sizeof(double));
  double* ay = (double*)calloc(ns, sizeof(double));
  double** kf[8];
  double** kvz[8];
  double* kem[8];
  for (int i = 0; i < 8; ++i) {
    kf[i] = (double**)calloc(ns, sizeof(double*));
    kvz[i] = (double**)calloc(ns, sizeof(double*));
    kem[i] = (double*)calloc(pl * 6, sizeof(double));
    for (int s = 0; s < ns; ++s) {
      kf[i][s] = (double*)calloc(vol4(&w->sp[s].g), sizeof(double));
      kvz[i][s] = (double*)calloc(pl, sizeof(double));
    }
  }
  for (int i = 0; i < 8; ++i) {
    /* the predictor of stage i: old + dt * sum_{j<i} a_ij k_j (stage 0: the old state itself) */
    for (int s = 0; s < ns; ++s) {
      const ok_geom* g = &w->sp[s].g;
      memcpy(f_new[s], f_old[s], sizeof(double) * vol4(g));
      memcpy(vz_new[s], vz_old[s], sizeof(double) * pl);
      for (int j = 0; j < i; ++j) {
        ok_xpby4d(f_new[s], kf[j][s], dt * A[i][j], g);
        ok_xpby2d(vz_new[s], kvz[j][s], dt * A[i][j], n1, n2, ng, 1);
      }
    }
    memcpy(em_new, em_old, sizeof(double) * pl * 6);
    for (int j = 0; j < i; ++j) ok_xpby2d(em_new, kem[j], dt * A[i][j], n1, n2, ng, 6);
    ok_vm_eval_rhs(w, kf[i], kem[i], kvz[i], f_new, em_new, vz_new, time + c[i] * dt, ax, ay);
  }
  for (int s = 0; s < ns; ++s) {
    const ok_geom* g = &w->sp[s].g;
    memcpy(f_new[s], f_old[s], sizeof(double) * vol4(g));
    memcpy(vz_new[s], vz_old[s], sizeof(double) * pl);
    for (int i = 0; i < 8; ++i) {
      ok_xpby4d(f_new[s], kf[i][s], dt * b[i], g);
      ok_xpby2d(vz_new[s], kvz[i][s], dt * b[i], n1, n2, ng, 1);
    }
  }
  memcpy(em_new, em_old, sizeof(double) * pl * 6);
  for (int i = 0; i < 8; ++i) ok_xpby2d(em_new, kem[i], dt * b[i], n1, n2, ng, 6);
  for (int i = 0; i < 8; ++i) {
    for (int s = 0; s < ns; ++s) { free(kf[i][s]); free(kvz[i][s]); }
    free(kf[i]); free(kvz[i]); free(kem[i]);
  }
  free(ax); free(ay);
}

void ok_vm_last_accel_max(const ok_vm_work* w, double* axmax, double* aymax) {
  for (int s = 0; s < w->ns; ++s) { axmax[s] = w->last_ax[s]; aymax[s] = w->last_ay[s]; }
}

/* VMSystem::stableDt (VMSystem.C:563-581): min of the species' computeDt and Maxwell::computeDt */
double ok_vm_stable_dt(const ok_vm_work* w, const double* axmax, const double* aymax, int rk_order) {
  const double pi = 4.0 * atan(1.0);
  double dt_stable = 1.7976931348623157e308;
  for (int s = 0; s < w->ns; ++s) {
    const ok_geom* g = &w->sp[s].g;
    const ok_species* sp = &w->sp[s];
    double lam[4];
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
    double alpha = rk_order == 4 ? 2.6 : 4.95, beta = rk_order == 4 ? 2.6 : 3.168;
    double ddt = sqrt(1.0 / (reLam * reLam / (alpha * alpha) + imLam * imLam / (beta * beta)));
    if (ddt < dt_stable) dt_stable = ddt;
  }
  const ok_geom* g0 = &w->sp[0].g;
  double dt_maxwell = 1.0 / (w->light_speed * (1.0 / g0->dx[0] + 1.0 / g0->dx[1]));
  if (dt_maxwell < dt_stable) dt_stable = dt_maxwell;
  return dt_stable;
}

/* ------------------------------------------------------------------------------------------
 * The boundary routines of MaxwellF.f that the periodic decks never reach (Maxwell.C:562-623 calls them when a
 * direction is not periodic) and the antenna source.  Arrays are (n1d, n2d, ncomp) over the interior n1 x n2 grown by ng;
 * at[4] = {x low, x high, y low, y high}: this box touches that physical boundary (the Fortran's m1a .eq. 0,
 * m1b .eq. nx-1, m2a .eq. 0, m2b .eq. ny-1).
 * ------------------------------------------------------------------------------------------ */
#define E3(a, i1, i2, c) (a)[(i1) + n1d * ((i2) + n2d * (int64_t)(c))]
/* zeroghost2d (MaxwellF.f:10-58) */
void ok_zero_ghost_2d(double* u, int n1, int n2, int ng, int dim) {
  const int64_t n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  for (int c = 0; c < dim; ++c)
    for (int i2 = 0; i2 < n2d; ++i2)
      for (int i1 = 0; i1 < n1d; ++i1)
        if (i1 < ng || i1 >= ng + n1 || i2 < ng || i2 >= ng + n2) E3(u, i1, i2, c) = 0.0;
}
/* maxwelladdantennasource (MaxwellF.f:359-389): dEMvars -= antenna_source on the interior, all six components */
void ok_maxwell_add_antenna_source(double* dem, const double* antenna, int n1, int n2, int ng) {
  const int64_t n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  for (int c = 0; c < 6; ++c)
    for (int i2 = ng; i2 < ng + n2; ++i2)
      for (int i1 = ng; i1 < ng + n1; ++i1) E3(dem, i1, i2, c) = E3(dem, i1, i2, c) - E3(antenna, i1, i2, c);
}
/* one ghost cell of maxwellsetembcs: the outgoing characteristic of the pair (s1 * comp a, comp b) is kept, the incoming
 * one zeroed (MaxwellF.f:519-541): low side w1 = 0, high side w2 = 0 */
static void em_characteristic(double* pa, double* pb, double s1, double c, int high) {
  double u1 = s1 * *pa, u2 = *pb;
  double w1 = +u1 / (2. * c) + u2 / 2.;
  double w2 = -u1 / (2. * c) + u2 / 2.;
  if (high) w2 = 0.0; else w1 = 0.0;
  u1 = c * (w1 - w2);
  u2 = w1 + w2;
  *pa = s1 * u1;
  *pb = u2;
}
/* maxwellsetembcs (MaxwellF.f:473-657) */
void ok_maxwell_set_em_bcs(double* em, int n1, int n2, int order, const int* at, int x_periodic, int y_periodic, double c) {
  const int ng = (order == 4) ? 2 : 3;
  const int64_t n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  if (x_periodic == 0) {
    for (int high = 0; high < 2; ++high) {
      if (!at[high]) continue;
      const int i1 = high ? ng + n1 - 1 : ng, d = high ? 1 : -1;
      for (int i2 = ng; i2 < ng + n2; ++i2)
        for (int i4 = 1; i4 <= ng; ++i4) {
          const int ig = i1 + d * i4;
          for (int k = 0; k < 6; ++k)
            E3(em, ig, i2, k) = +3.0 * E3(em, ig - d, i2, k) - 3.0 * E3(em, ig - 2 * d, i2, k) + 1.0 * E3(em, ig - 3 * d, i2, k);
          em_characteristic(&E3(em, ig, i2, 1), &E3(em, ig, i2, 5), 1.0, c, high);   /* Ey, Bz */
          em_characteristic(&E3(em, ig, i2, 2), &E3(em, ig, i2, 4), -1.0, c, high);  /* -Ez, By */
        }
    }
  }
  if (y_periodic == 0) {
    for (int high = 0; high < 2; ++high) {
      if (!at[2 + high]) continue;
      const int i2 = high ? ng + n2 - 1 : ng, d = high ? 1 : -1;
      for (int i1 = ng; i1 < ng + n1; ++i1)
        for (int i4 = 1; i4 <= ng; ++i4) {
          const int ig = i2 + d * i4;
          for (int k = 0; k < 6; ++k)
            E3(em, i1, ig, k) = +3.0 * E3(em, i1, ig - d, k) - 3.0 * E3(em, i1, ig - 2 * d, k) + 1.0 * E3(em, i1, ig - 3 * d, k);
          em_characteristic(&E3(em, i1, ig, 0), &E3(em, i1, ig, 5), -1.0, c, high);  /* -Ex, Bz */
          em_characteristic(&E3(em, i1, ig, 2), &E3(em, i1, ig, 3), 1.0, c, high);   /* Ez, Bx */
        }
    }
  }
}
/* maxwellsetvzbcs (MaxwellF.f:661-731): even reflection about the boundary cell, x edges over the whole y extent of the
 * data box first, then y edges over the whole x extent */
void ok_maxwell_set_vz_bcs(double* vz, int n1, int n2, int order, const int* at, int x_periodic, int y_periodic) {
  const int ng = (order == 4) ? 2 : 3;
  const int64_t n1d = n1 + 2 * ng, n2d = n2 + 2 * ng;
  if (x_periodic == 0) {
    if (at[0])
      for (int i2 = 0; i2 < n2d; ++i2)
        for (int i3 = 1; i3 <= ng; ++i3) E3(vz, ng - i3, i2, 0) = E3(vz, ng + i3, i2, 0);
    if (at[1])
      for (int i2 = 0; i2 < n2d; ++i2)
        for (int i3 = 1; i3 <= ng; ++i3) E3(vz, ng + n1 - 1 + i3, i2, 0) = E3(vz, ng + n1 - 1 - i3, i2, 0);
  }
  if (y_periodic == 0) {
    if (at[2])
      for (int i1 = 0; i1 < n1d; ++i1)
        for (int i3 = 1; i3 <= ng; ++i3) E3(vz, i1, ng - i3, 0) = E3(vz, i1, ng + i3, 0);
    if (at[3])
      for (int i1 = 0; i1 < n1d; ++i1)
        for (int i3 = 1; i3 <= ng; ++i3) E3(vz, i1, ng + n2 - 1 + i3, 0) = E3(vz, i1, ng + n2 - 1 - i3, 0);
  }
}
#undef E3
