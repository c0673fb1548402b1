/*
 * loki_b200_f77.h -- Level 0 of the drop-in boundary: the reference's own Fortran-77 symbols.
 *
 * LOKI's only FFI is its C++ -> Fortran seam: extern "C" prototypes whose names are macro-aliased to the
 * lower-case, trailing-underscore symbols gfortran emits (KineticSpeciesF.H:13-35, PoissonF.H:13-14,
 * MaxwellF.H:13-20), every argument passed by reference, boxes expanded to 8 (4D) or 4 (2D) integers by
 * BOX4D_TO_FORT / BOX2D_TO_FORT (tbox/Box.H:893-897).  libloki_b200.so exports the SAME symbols with the SAME
 * argument lists, so the inline wrappers of KineticSpecies.H:404-562, Poisson.C and Maxwell.C link against it
 * unchanged (relink, not edit).  Differences a caller must know:
 *
 *   * every ParallelArray argument is a DEVICE pointer (the first element of the array, as `*array.getData()`
 *     is in the reference); scalars, boxes and the domain metadata arrays of PROBLEMDOMAIN_TO_FORT
 *     (xlo, xhi, dx / deltax, supergrid_lo / hi; ProblemDomain.H:318-321) stay HOST references; scalar RESULTS
 *     (axmax, aymax, ke_e_dot ...) are written to the host reference before the call returns (the calls synchronise);
 *   * `ic` (KineticSpecies.H:439: the initial-condition object laundered through an int64) must hold the
 *     address of an lk_inflow (loki_b200.h) describing the same initial condition as device tables: device
 *     code cannot call back into ICInterface.C:36-57;
 *   * there is no status argument in the Fortran ABI: failures are reported through lk_f77_status() /
 *     lk_last_error() and leave the outputs untouched;
 *   * arithmetic: lk_set_strict(1) gives the bits of a gfortran -O2 build of the reference (the parity harness
 *     uses it); the default is the production arithmetic.
 *
 * tests/test_gpu_f77abi.py drives this library and the transliterated reference Fortran (oracle/_ref) through
 * the same argument lists (tests/ref_binding.py) and compares bits.
 */
#ifndef LOKI_B200_F77_H
#define LOKI_B200_F77_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status of the last Fortran-ABI call on this thread's library state (LK_OK, LK_ERR_*) */
int lk_f77_status(void);

/* KineticSpeciesF.H:41-61 (KineticSpeciesF.f:10-38) */
void xpby4d_(double* x, const double* y, const double* b, const int* nd1lo, const int* nd1hi, const int* nd2lo,
             const int* nd2hi, const int* nd3lo, const int* nd3hi, const int* nd4lo, const int* nd4hi, const int* n1lo,
             const int* n1hi, const int* n2lo, const int* n2hi, const int* n3lo, const int* n3hi, const int* n4lo,
             const int* n4hi);
/* KineticSpeciesF.H:191-222 (KineticSpeciesF.f:42-114) */
void setphasespacevel4d_(double* vel3, double* vel4, const int* nv1a, const int* nv1b, const int* nv2a, const int* nv2b,
                         const int* nv3a, const int* nv3b, const int* nv4a, const int* nv4b, const int* ni1a,
                         const int* ni1b, const int* ni2a, const int* ni2b, const int* ni3a, const int* ni3b,
                         const int* ni4a, const int* ni4b, const double* vxface_velocities,
                         const double* vyface_velocities, const double* normalization, const double* bz_const,
                         const double* accel, const int* na1a, const int* na1b, const int* na2a, const int* na2b,
                         double* axmax, double* aymax);
/* KineticSpeciesF.H:224-250 (KineticSpeciesF.f:118-197) */
void setphasespacevelmaxwell4d_(double* vel3, double* vel4, const int* nv1a, const int* nv1b, const int* nv2a,
                                const int* nv2b, const int* nv3a, const int* nv3b, const int* nv4a, const int* nv4b,
                                const int* ni1a, const int* ni1b, const int* ni2a, const int* ni2b, const int* ni3a,
                                const int* ni3b, const int* ni4a, const int* ni4b, const double* vxface_velocities,
                                const double* vyface_velocities, const double* normalization, const double* bz_const,
                                const double* em_vars, const double* vz, double* axmax, double* aymax);
/* KineticSpeciesF.H:97-127 (KineticSpeciesF.f:1036-1162); ng*: global box grown by the ghosts, nl*: data box */
void setaccelerationbcs4d_(double* u, const int* ng1a, const int* ng1b, const int* ng2a, const int* ng2b, const int* ng3a,
                           const int* ng3b, const int* ng4a, const int* ng4b, const int* nl1a, const int* nl1b,
                           const int* nl2a, const int* nl2b, const int* nl3a, const int* nl3b, const int* nl4a,
                           const int* nl4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b, const int* n3a,
                           const int* n3b, const int* n4a, const int* n4b, const int* solution_order, const double* vel3,
                           const double* vel4, const int64_t* ic);
/* KineticSpeciesF.H:63-95 (KineticSpeciesF.f:1166-1297) */
void setadvectionbcs4d_(double* u, const int* ng1a, const int* ng1b, const int* ng2a, const int* ng2b, const int* ng3a,
                        const int* ng3b, const int* ng4a, const int* ng4b, const int* nl1a, const int* nl1b,
                        const int* nl2a, const int* nl2b, const int* nl3a, const int* nl3b, const int* nl4a,
                        const int* nl4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b, const int* n3a,
                        const int* n3b, const int* n4a, const int* n4b, const int* solution_order, const double* vel1,
                        const double* vel2, const int* xperiodic, const int* yperiodic, const int64_t* ic);
/* KineticSpeciesF.H:293-316 (KineticSpeciesF.f:1949-2089): ASSIGNS rhs */
void computeadvectionderivatives4d_(double* rhs, const double* f, const int* nd1a, const int* nd1b, const int* nd2a,
                                    const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b,
                                    const int* n1a, const int* n1b, const int* n2a, const int* n2b, const int* n3a,
                                    const int* n3b, const int* n4a, const int* n4b, const double* vel1,
                                    const double* vel2, const double* deltax, const int* solution_order);
/* KineticSpeciesF.H:318-341 (KineticSpeciesF.f:2093-2245): ACCUMULATES into rhs */
void computeaccelerationderivatives4d_(double* rhs, const double* f, const int* nd1a, const int* nd1b, const int* nd2a,
                                       const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a,
                                       const int* nd4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b,
                                       const int* n3a, const int* n3b, const int* n4a, const int* n4b, const double* vel3,
                                       const double* vel4, const double* dx, const int* solution_order);
/* KineticSpeciesF.f:2400-2443 */
void computecurrents_(const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b,
                      const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b,
                      const int* n3a, const int* n3b, const int* n4a, const int* n4b, const double* velocities,
                      const double* u, const double* vz, double* jx, double* jy, double* jz);
/* KineticSpeciesF.f:2563-2602 */
void computekeedot_(const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b,
                    const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b,
                    const int* n3a, const int* n3b, const int* n4a, const int* n4b, const double* xlo, const double* xhi,
                    const double* dx, const double* u, const double* charge, const double* velocities,
                    const double* ext_efield, double* ke_e_dot);
/* KineticSpeciesF.H:273-291 (KineticSpeciesF.f:1838-1945): face fits and fluxes in x and y on the rotated face arrays */
void computeadvectionfluxes4d_(double* flux1, double* flux2, const int* nd1a, const int* nd1b, const int* nd2a,
                               const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b,
                               const double* vel1, const double* vel2, double* face1, double* face2, const double* u,
                               const double* dx, const int* solution_order);
/* KineticSpeciesF.H:343-361 (KineticSpeciesF.f:2249-2355): the same in vx and vy */
void computeaccelerationfluxes4d_(double* flux3, double* flux4, const int* nd1a, const int* nd1b, const int* nd2a,
                                  const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b,
                                  const double* vel3, const double* vel4, double* face3, double* face4, const double* u,
                                  const double* dx, const int* solution_order);
/* KineticSpeciesF.H:363-380 (KineticSpeciesF.f:985-1032) */
void accumfluxdiv4d_(double* rhs, const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a,
                     const int* nd3b, const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a,
                     const int* n2b, const int* n3a, const int* n3b, const int* n4a, const int* n4b, const double* fluxx1,
                     const double* fluxx2, const double* fluxx3, const double* fluxx4, const double* deltax);
/* KineticSpeciesF.H:554-591 (KineticSpeciesF.f:2734-2893); ng*: the domain box; ke_flux is a HOST scalar that must come in
 * as 0 for a box that touches the boundary (the reference's callers zero it); summed as a tree, not sequentially */
void computekeflux_(const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b,
                    const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b,
                    const int* n3a, const int* n3b, const int* n4a, const int* n4b, const int* ng1a, const int* ng1b,
                    const int* ng2a, const int* ng2b, const int* ng3a, const int* ng3b, const int* ng4a, const int* ng4b,
                    const double* dx, const double* face_flux1, const double* face_flux2, const double* face_flux3,
                    const double* face_flux4, const double* velocities, const double* vxface_velocities,
                    const double* vyface_velocities, const int* dir, const int* side, const double* mass, double* ke_flux);
/* KineticSpeciesF.H:593-625 (KineticSpeciesF.f:2897-2990); ke_flux: device (n1d,n2d), accumulated into */
void computekevelspaceflux_(const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a,
                            const int* nd3b, const int* nd4a, const int* nd4b, const int* n1a, const int* n1b,
                            const int* n2a, const int* n2b, const int* n3a, const int* n3b, const int* n4a, const int* n4b,
                            const int* ng1a, const int* ng1b, const int* ng2a, const int* ng2b, const int* ng3a,
                            const int* ng3b, const int* ng4a, const int* ng4b, const double* dx, const double* face_flux3,
                            const double* face_flux4, double* ke_flux, const double* mass, const double* vxface_velocities,
                            const double* vyface_velocities, const int* side, const int* dir);
/* TZSourceF.f:10-27, :79-97 (declared in TZSourceF.H; called from TrigTZSource.C:44-82) */
void settrigtzsource_(double* f, const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a,
                      const int* nd3b, const int* nd4a, const int* nd4b, const double* xlo, const double* xhi, const double* dx,
                      const double* time, const double* velocities, const double* dparams);
void computetrigtzsourceerror_(double* error, const double* soln, const int* nd1a, const int* nd1b, const int* nd2a,
                               const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b,
                               const double* xlo, const double* xhi, const double* dx, const double* time,
                               const double* velocities, const double* dparams);
/* ElectronTZSourceF.f:10-27, :79-97 (ElectronTZSourceF.H; called from ElectronTrigTZSource.C:44-82) */
void setelectrontrigtzsource_(double* f, const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a,
                              const int* nd3b, const int* nd4a, const int* nd4b, const double* xlo, const double* xhi,
                              const double* dx, const double* time, const double* velocities, const double* dparams);
void computeelectrontrigtzsourceerror_(double* error, const double* soln, const int* nd1a, const int* nd1b, const int* nd2a,
                                       const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b,
                                       const double* xlo, const double* xhi, const double* dx, const double* time,
                                       const double* velocities, const double* dparams);
/* TwoSpecies_ElectronTZSourceF.f:10-27, :165-183 and TwoSpecies_IonTZSourceF.f:10-27, :143-161 (called from
 * TwoSpecies_ElectronTrigTZSource.C / TwoSpecies_IonTrigTZSource.C); dparams = {amp, electron_mass, ion_mass} */
void settwoelectrontrigtzsource_(double* f, const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a,
                                 const int* nd3b, const int* nd4a, const int* nd4b, const double* xlo, const double* xhi,
                                 const double* dx, const double* time, const double* velocities, const double* dparams);
void computetwoelectrontrigtzsourceerror_(double* error, const double* soln, const int* nd1a, const int* nd1b, const int* nd2a,
                                          const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b,
                                          const double* xlo, const double* xhi, const double* dx, const double* time,
                                          const double* velocities, const double* dparams);
void settwoiontrigtzsource_(double* f, const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a,
                            const int* nd3b, const int* nd4a, const int* nd4b, const double* xlo, const double* xhi,
                            const double* dx, const double* time, const double* velocities, const double* dparams);
void computetwoiontrigtzsourceerror_(double* error, const double* soln, const int* nd1a, const int* nd1b, const int* nd2a,
                                     const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b,
                                     const double* xlo, const double* xhi, const double* dx, const double* time,
                                     const double* velocities, const double* dparams);
/* KineticSpeciesF.f:2995-3034 */
void appendkrook_(const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b,
                  const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a, const int* n2b,
                  const int* n3a, const int* n3b, const int* n4a, const int* n4b, const double* dt, const int64_t* ic,
                  const double* nu, const double* u, double* rhs);
/* ---- PitchAngleCollisionOperatorF.H:13-17, 24-134 (PitchAngleCollisionOperatorF.f:1618-1852) ----
 * Arrays on the device; xlo / xhi / dx / range_lo / range_hi / dparams / iparams are host references
 * (PROBLEMDOMAIN_TO_FORT and the operator's parameter vectors: dparams = {vfloor, vthermal_dt, nuCoeff},
 * iparams = {conservative, solution_order, do_relativity}; do_relativity must be 0). */
void appendpitchanglecollision_(double* rhs, const double* f, const double* velocities, const double* ivx, const double* ivy,
                                const double* vth, const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b,
                                const int* nd3a, const int* nd3b, const int* nd4a, const int* nd4b, const int* n1a,
                                const int* n1b, const int* n2a, const int* n2b, const int* n3a, const int* n3b,
                                const int* n4a, const int* n4b, const double* xlo, const double* xhi, const double* dx,
                                const double* range_lo, const double* range_hi, const double* dparams, const int* iparams);
/* :1706-1750: rn, rgammax, rgammay (n1d,n2d) = sums of max(|u|, 1e-10) {1, vx, vy} over the interior velocity cells */
void computepitchanglespeciesmoments_(double* rn, double* rgammax, double* rgammay, const double* u, const int* nd1a,
                                      const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b,
                                      const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a,
                                      const int* n2b, const int* n3a, const int* n3b, const int* n4a, const int* n4b,
                                      const double* velocities);
/* :1754-1781: vx = gammax / n, vy = gammay / n */
void computepitchanglespeciesreducedfields_(double* vx, double* vy, const double* n, const double* gammax,
                                            const double* gammay, const int* nd1a, const int* nd1b, const int* nd2a,
                                            const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a,
                                            const int* nd4b);
/* :1785-1824: rkec += sum of ((vx - rvx0)^2 + (vy - rvy0)^2) max(|u|, 1e-10) */
void computepitchanglespecieskec_(double* rkec, const double* rvx0, const double* rvy0, const double* u, const int* nd1a,
                                  const int* nd1b, const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b,
                                  const int* nd4a, const int* nd4b, const int* n1a, const int* n1b, const int* n2a,
                                  const int* n2b, const int* n3a, const int* n3b, const int* n4a, const int* n4b,
                                  const double* velocities);
/* :1828-1852: vthsq = sqrt(0.5 kec / n) */
void computepitchanglespeciesvthermal_(double* vthsq, const double* kec, const double* n, const int* nd1a, const int* nd1b,
                                       const int* nd2a, const int* nd2b, const int* nd3a, const int* nd3b, const int* nd4a,
                                       const int* nd4b);
/* PoissonF.H (PoissonF.f:10-64): rhs -= mean(rhs) over the interior; comm is ignored (one rank solves) */
void neutralizecharge4d_(const int* md1a, const int* md1b, const int* md2a, const int* md2b, const int* n1a, const int* n1b,
                         const int* n2a, const int* n2b, double* rhs, const int* comm);
/* PoissonF.f:68-123 */
void computeefieldfrompotential_(const int* nd1a, const int* nd1b, const int* nd2a, const int* nd2b, const int* n1a,
                                 const int* n1b, const int* n2a, const int* n2b, const int* solution_order,
                                 const int* em_vars_dim, const double* dx, double* emvars, const double* phi);
/* MaxwellF.f:97-355; the supergrid metric is not built: supergrid_lo/hi must leave the whole domain unstretched */
void maxwellevalrhs_(const int* md1a, const int* md1b, const int* md2a, const int* md2b, const int* m1a, const int* m1b,
                     const int* m2a, const int* m2b, const double* xlo, const double* xhi, const double* dx,
                     const double* c, const double* avweak, const double* avstrong, const int* solution_order,
                     const double* supergrid_lo, const double* supergrid_hi, const double* emvars, const double* jx,
                     const double* jy, const double* jz, double* demvars);
/* MaxwellF.f:442-469 */
void maxwellevalvzrhs_(const int* md1a, const int* md1b, const int* md2a, const int* md2b, const int* m1a, const int* m1b,
                       const int* m2a, const int* m2b, const double* charge_per_mass, const double* emvars, double* dvz);
/* MaxwellF.H:26-37 (MaxwellF.f:10-58), :77-89 (:359-389), :105-121 (:473-657), :123-133 (:661-731) */
void zeroghost2d_(double* u, const int* n1a, const int* n1b, const int* n2a, const int* n2b, const int* nd1a, const int* nd1b,
                  const int* nd2a, const int* nd2b, const int* dim);
void maxwelladdantennasource_(const int* md1a, const int* md1b, const int* md2a, const int* md2b, const int* m1a, const int* m1b,
                              const int* m2a, const int* m2b, const double* xlo, const double* xhi, const double* dx,
                              const double* antenna_source, double* dEMvars);
void maxwellsetembcs_(const int* md1a, const int* md1b, const int* md2a, const int* md2b, const int* m1a, const int* m1b,
                      const int* m2a, const int* m2b, double* EMvars, const int* nx, const int* ny, const int* xPeriodic,
                      const int* yPeriodic, const int* solution_order, const double* c);
void maxwellsetvzbcs_(const int* md1a, const int* md1b, const int* md2a, const int* md2b, const int* m1a, const int* m1b,
                      const int* m2a, const int* m2b, double* vz, const int* nx, const int* ny, const int* xPeriodic,
                      const int* yPeriodic, const int* solution_order);
/* MaxwellF.f:62-93 */
void xpby2d_(double* x, const double* y, const double* b, const int* nd1a, const int* nd1b, const int* nd2a,
             const int* nd2b, const int* n1a, const int* n1b, const int* n2a, const int* n2b, const int* dim);

#ifdef __cplusplus
}
#endif
#endif
