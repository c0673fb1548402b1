"""System-level parity: the C++ host mirror (lk_vp_*: VPSystem / KineticSpecies / RK integrators on the
device) against the oracle's single-rank restatement of VPSystem::evalRHS and the RK4/RK6 integrators,
on deck-shaped problems (test/planeEPW_fixedIons, test/planeIAW, test/planeIAW_6)."""
import ctypes as C

import numpy as np
import pytest

import decks
from util import cell_rel_err, rel_err, star_rel_err

pytestmark = pytest.mark.gpu


def _oracle(ok, deck):
    keep = []
    sp = deck.oracle_species(keep)
    xlo = (C.c_double * 2)(deck.xlim[0], deck.xlim[2])
    xhi = (C.c_double * 2)(deck.xlim[1], deck.xlim[3])
    w = ok.ok_vp_work_create(len(deck.species), sp, C.byref(xlo), C.byref(xhi))
    # deck options beyond the benchmark decks: open boundaries, the JB fills, Krook layers
    ok.ok_vp_set_options(w, int(not deck.periodic[0]), int(not deck.periodic[1]), int(deck.use_new_bcs))
    for s_, sp_ in enumerate(deck.species):
        nu = deck.krook_nu(sp_)
        if nu is not None:
            ok.ok_vp_set_krook(w, s_, np.ascontiguousarray(nu).ravel())
        co = getattr(sp_, "collision", None)
        if co:   # a pitch-angle collision operator: {range_lo[2], range_hi[2], vfloor, vthermal_dt, nuCoeff, conservative}
            p = np.array(list(co["range_lo"]) + list(co["range_hi"]) + [co["vfloor"], co["vthermal_dt"], co["nu_coef"],
                                                                        float(co.get("conservative", 1))])
            ok.ok_vp_set_pitch_angle(w, s_, p.ctypes.data)
        tz = getattr(sp_, "tz", None)
        if tz:   # a TrigTZSource: the manufactured-solution forcing in completeRHS
            ok.ok_vp_set_trig_tz(w, s_, int(tz.get("kind", 1)), float(tz["amp"]), float(tz.get("electron_mass", 1.0)),
                                 float(tz.get("ion_mass", 1.0)))
    return w, sp, keep


def _ptrs(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def _product(deck, states, tables):
    from loki_b200 import host
    H = host.lib()
    d = deck.product_desc()
    sys_ = C.c_void_p()
    assert H.lk_vp_create(C.byref(sys_), C.byref(d), None) == 0, H.lk_last_error()
    for s, f in enumerate(states):
        assert H.lk_vp_set_state(sys_, s, f.ctypes.data) == 0
        assert deck.set_inflow(H, sys_, s) == 0      # the inflow tables in the form the species' IC class has
    assert deck.apply_options(H, sys_) == 0
    return H, sys_


def _perturb(f, seed, amp=0.05):
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(f * (1.0 + amp * rng.uniform(-1, 1, size=f.shape)))


DECKS = [
    lambda: decks.plane_epw(n=(16, 8), nv=(32, 16)),
    lambda: decks.plane_iaw(n=(12, 10), nv=(16, 12)),
    lambda: decks.plane_iaw(n=(10, 10), nv=(16, 10), order=6, rk=6),   # planeIAW_6 deck grid verbatim
    # InterpenetratingStreams: three species, two-sided and centred erf half-plane ICs, order 6 / RK6
    lambda: decks.interpenetrating_streams(n=(16, 7), nv=(12, 10)),
]


@pytest.mark.parametrize("mk", DECKS)
def test_eval_rhs_strict_matches_reference_order(lk, ok, strict, mk):
    """VPSystem::evalRHS, unfused and in the reference's order: rho, E, rhs, lambda_max bit for bit"""
    import torch
    deck = mk()
    w, sp, keep = _oracle(ok, deck)
    states, tables = [], []
    for k, s in enumerate(deck.species):
        f, fx, fv, fnorm = deck.initial_state(s)
        states.append(_perturb(f, 10 + k))
        tables.append((fx, fv, fnorm))
    t = 0.37
    ns = len(states)
    f_o = [s.copy() for s in states]
    rhs_o = [np.zeros_like(s) for s in states]
    ke, ax, ay = np.zeros(ns), np.zeros(ns), np.zeros(ns)
    ok.ok_vp_eval_rhs(w, _ptrs(rhs_o), _ptrs(f_o), t, ke, ax, ay)
    H, sys_ = _product(deck, states, tables)
    rhs_d = [torch.zeros(s.shape, dtype=torch.float64, device="cuda") for s in states]
    ptrs = (C.c_void_p * ns)(*[r.data_ptr() for r in rhs_d])
    assert H.lk_vp_eval_rhs(sys_, ptrs, t) == 0, H.lk_last_error()
    ng = deck.ng
    n1d, n2d = deck.n[0] + 2 * ng, deck.n[1] + 2 * ng
    em_o = np.ctypeslib.as_array(ok.ok_vp_em_vars(w), shape=(2, n2d, n1d))
    rho_o = np.ctypeslib.as_array(ok.ok_vp_rho(w), shape=(n2d, n1d))
    em_d, rho_d = np.empty_like(em_o), np.empty_like(rho_o)
    torch.cuda.synchronize()
    assert lk.lk_memcpy_d2h(em_d.ctypes.data, H.lk_vp_em_vars_ptr(sys_), em_d.nbytes) == 0
    assert lk.lk_memcpy_d2h(rho_d.ctypes.data, H.lk_vp_rho_ptr(sys_), rho_d.nbytes) == 0
    assert np.array_equal(rho_d, rho_o)
    assert np.array_equal(em_d, em_o)
    for s in range(ns):
        assert np.array_equal(rhs_d[s].cpu().numpy(), rhs_o[s])
        out = np.empty_like(states[s])
        assert H.lk_vp_get_state(sys_, s, out.ctypes.data) == 0
        assert np.array_equal(out, f_o[s])          # ghost cells of the evaluated state, like the reference mutates them
        lam = (C.c_double * 2)()
        assert H.lk_vp_lambda_max(sys_, s, C.byref(lam)) == 0
        assert (lam[0], lam[1]) == (ax[s], ay[s])
    dt_o = ok.ok_vp_stable_dt(w, ax, ay, deck.rk)
    dt_d = C.c_double()
    assert H.lk_vp_stable_dt(sys_, C.byref(dt_d)) == 0
    assert dt_d.value == dt_o
    H.lk_vp_destroy(sys_)
    ok.ok_vp_work_destroy(w)


@pytest.mark.parametrize("mode", ["strict", "production"])
@pytest.mark.parametrize("mk", DECKS)
def test_one_step_matches_oracle(lk, ok, mk, mode):
    """one full RK4 / RK6 step, fused on the device, against the unfused oracle.  Strict arithmetic:
    the distribution is bit-identical.  Production arithmetic: checkTests.C:345-358 per-cell relative
    difference <= 1e-12 (north-star tolerance for the distribution after one step)."""
    deck = mk()
    old = lk.lk_set_strict(1 if mode == "strict" else 0)
    try:
        w, sp, keep = _oracle(ok, deck)
        states, tables = [], []
        for k, s in enumerate(deck.species):
            f, fx, fv, fnorm = deck.initial_state(s)
            states.append(_perturb(f, 20 + k, amp=0.02))
            tables.append((fx, fv, fnorm))
        ns = len(states)
        t0, dt = 0.25, 0.02
        f_old = [s.copy() for s in states]
        f_new = [np.zeros_like(s) for s in states]
        ke = np.zeros(ns)
        (ok.ok_vp_rk4_step if deck.rk == 4 else ok.ok_vp_rk6_step)(w, _ptrs(f_new), _ptrs(f_old), t0, dt, ke)
        H, sys_ = _product(deck, states, tables)
        assert H.lk_vp_set_time(sys_, t0) == 0
        assert H.lk_vp_advance(sys_, dt) == 0, H.lk_last_error()
        ng = deck.ng
        I = (slice(ng, -ng),) * 4
        for s in range(ns):
            out = np.empty_like(states[s])
            assert H.lk_vp_get_state(sys_, s, out.ctypes.data) == 0
            assert np.any(out[I] != states[s][I])
            if mode == "strict":
                assert np.array_equal(out[I], f_new[s][I])
            else:
                # rough data: f + dt*rhs cancels in the Maxwellian tails, so the error is measured against
                # the cell's stencil neighbourhood (the smooth-deck test below uses checkTests' metric)
                assert star_rel_err(out, f_new[s], np.maximum(np.abs(states[s]), np.abs(f_new[s])), ng) <= 1e-12
            if deck.species[s].driver:
                v = C.c_double()
                assert H.lk_vp_ke_e_dot(sys_, s, C.byref(v)) == 0
                assert abs(v.value - ke[s]) <= 1e-12 * abs(ke[s]) + 1e-300   # tree vs sequential sum
                assert ke[s] != 0.0
        assert abs(H.lk_vp_time(sys_) - (t0 + dt)) < 1e-15
        H.lk_vp_destroy(sys_)
        ok.ok_vp_work_destroy(w)
    finally:
        lk.lk_set_strict(old)


def _opt(deck, periodic=(True, True), jb=False, krook=None):
    deck.periodic, deck.use_new_bcs = periodic, jb
    if krook:
        for sp_ in deck.species:
            sp_.krook = dict(krook)
    return deck


OPTION_DECKS = {
    # completeRHS's Krook layer (KineticSpecies.C:1049-1062): damping towards the IC in the outer thirds of x
    "krook": lambda: _opt(decks.plane_epw(n=(16, 8), nv=(32, 16)), krook=dict(x1a=-3.0, x1b=3.0, coefficient=0.3)),
    "krook_rk6": lambda: _opt(decks.plane_iaw(n=(10, 10), nv=(16, 10), order=6, rk=6), krook=dict(x2a=-100.0, x1b=4.0)),
    # use_new_bcs (VPSystem.C:819-821): the JB velocity-boundary fill
    "jb": lambda: _opt(decks.plane_iaw(n=(12, 10), nv=(16, 12)), jb=True),
    # open boundaries in x, in y, in both with the JB advection fill (KineticSpecies.H:998-1031)
    "open_x": lambda: _opt(decks.plane_epw(n=(16, 8), nv=(32, 16)), periodic=(False, True)),
    "open_y": lambda: _opt(decks.plane_iaw(n=(12, 10), nv=(16, 12)), periodic=(True, False)),
    "open_xy_jb_krook": lambda: _opt(decks.plane_epw(n=(16, 8), nv=(32, 16)), periodic=(False, False), jb=True,
                                     krook=dict(x1a=-4.0, coefficient=0.5)),
}


@pytest.mark.parametrize("mode", ["strict", "production"])
@pytest.mark.parametrize("name", sorted(OPTION_DECKS))
def test_deck_options_steps_match_oracle(lk, ok, name, mode):
    """Krook layers, the JB boundary conditions and open x / y boundaries wired into the stage loop: three RK steps
    against the oracle (whose pieces are pinned to the reference Fortran).  Strict: the distribution bit for bit;
    production: within 1e-12 of the stencil neighbourhood per step taken."""
    deck = OPTION_DECKS[name]()
    old = lk.lk_set_strict(1 if mode == "strict" else 0)
    try:
        w, sp, keep = _oracle(ok, deck)
        states, tables = [], []
        for k, s in enumerate(deck.species):
            f, fx, fv, fnorm = deck.initial_state(s)
            states.append(_perturb(f, 70 + k, amp=0.02))
            tables.append((fx, fv, fnorm))
        ns = len(states)
        H, sys_ = _product(deck, states, tables)
        f_old = [s.copy() for s in states]
        f_new = [np.zeros_like(s) for s in states]
        ke = np.zeros(ns)
        t, dt = 0.25, 0.02
        ng = deck.ng
        I = (slice(ng, -ng),) * 4
        for step in range(3):
            (ok.ok_vp_rk4_step if deck.rk == 4 else ok.ok_vp_rk6_step)(w, _ptrs(f_new), _ptrs(f_old), t, dt, ke)
            assert H.lk_vp_set_time(sys_, t) == 0
            assert H.lk_vp_advance(sys_, dt) == 0, H.lk_last_error()
            t += dt
            f_old, f_new = f_new, f_old
            for s in range(ns):
                out = np.empty_like(states[s])
                assert H.lk_vp_get_state(sys_, s, out.ctypes.data) == 0
                if mode == "strict":
                    assert np.array_equal(out[I], f_old[s][I]), (step, s)
                else:
                    assert star_rel_err(out, f_old[s], np.maximum(np.abs(states[s]), np.abs(f_old[s])), ng) <= (step + 1) * 1e-12
        H.lk_vp_destroy(sys_)
        ok.ok_vp_work_destroy(w)
    finally:
        lk.lk_set_strict(old)


def test_deck_options_change_the_answer(lk, ok, fast):
    """the options are not no-ops: each deck above differs from its plain periodic twin after one step"""
    for name in sorted(OPTION_DECKS):
        outs = []
        for plain in (False, True):
            deck = OPTION_DECKS[name]()
            if plain:
                deck.periodic, deck.use_new_bcs = (True, True), False
                for sp_ in deck.species:
                    sp_.krook = None
            states = [_perturb(deck.initial_state(s)[0], 70 + k, amp=0.02) for k, s in enumerate(deck.species)]
            H, sys_ = _product(deck, states, None)
            assert H.lk_vp_set_time(sys_, 0.25) == 0 and H.lk_vp_advance(sys_, 0.02) == 0
            out = np.empty_like(states[0])
            assert H.lk_vp_get_state(sys_, 0, out.ctypes.data) == 0
            outs.append(out)
            H.lk_vp_destroy(sys_)
        ng = 3 if "rk6" in name else 2
        I = (slice(ng, -ng),) * 4
        assert not np.array_equal(outs[0][I], outs[1][I]), name


SMOOTH_DECKS = [
    # velocity grids as fine as the decks' own (dv <= 0.22 thermal speeds): the per-cell metric compares
    # Maxwellian-tail cells with neighbours that are orders of magnitude larger on a coarser grid
    lambda: decks.plane_epw(n=(16, 8), nv=(64, 64), A=0.05, ky1=1.0 / 78),
    lambda: decks.plane_iaw(n=(12, 10), nv=(64, 64), A=0.05, ky1=1.0 / 234),
    lambda: decks.plane_iaw(n=(10, 10), nv=(64, 64), order=6, rk=6, A=0.05, ky1=1.0 / 234),
]


@pytest.mark.parametrize("mk", SMOOTH_DECKS)
def test_one_step_deck_ic_checktests_metric(lk, ok, fast, mk):
    """the decks' own analytic initial condition (PerturbedMaxwellianIC with a 5 % spatial mode), one
    RK step, production arithmetic: per-cell relative difference of checkTests.C:345-358 <= 1e-12, the
    north-star tolerance for the distribution after one step"""
    deck = mk()
    w, sp, keep = _oracle(ok, deck)
    states, tables = [], []
    for s in deck.species:
        f, fx, fv, fnorm = deck.initial_state(s)
        states.append(f)
        tables.append((fx, fv, fnorm))
    ns = len(states)
    t0, dt = 0.0, 0.05
    f_old = [s.copy() for s in states]
    f_new = [np.zeros_like(s) for s in states]
    ke = np.zeros(ns)
    (ok.ok_vp_rk4_step if deck.rk == 4 else ok.ok_vp_rk6_step)(w, _ptrs(f_new), _ptrs(f_old), t0, dt, ke)
    H, sys_ = _product(deck, states, tables)
    assert H.lk_vp_set_time(sys_, t0) == 0
    assert H.lk_vp_advance(sys_, dt) == 0, H.lk_last_error()
    ng = deck.ng
    I = (slice(ng, -ng),) * 4
    for s in range(ns):
        out = np.empty_like(states[s])
        assert H.lk_vp_get_state(sys_, s, out.ctypes.data) == 0
        assert np.any(out[I] != states[s][I])
        # every cell down to 1e-20 of the peak (|v| < 9.5 thermal speeds); below that a Maxwellian-tail
        # cell sits next to neighbours hundreds of times larger and its own rounding unit is not the scale
        big = f_new[s][I] >= 1e-20 * f_new[s][I].max()
        assert big.sum() > 0.5 * big.size
        assert cell_rel_err(out[I][big], f_new[s][I][big]) <= 1e-12
        assert star_rel_err(out, f_new[s], f_new[s], ng) <= 1e-13
    H.lk_vp_destroy(sys_)
    ok.ok_vp_work_destroy(w)


def test_interpenetrating_streams_deck_grid(lk, ok, fast):
    """test/InterpenetratingStreams at the deck's own grid (128 x 7 x 24 x 16, electrons + He + C, order 6 /
    RK6), its own initial condition, one step with dt = cfl * stableDt after the seeding evalRHS.  The
    deck's velocity grid is coarse (dv = 0.63 / 0.94 thermal speeds), so a Maxwellian-tail cell sits next to
    neighbours up to 700 times larger and its own rounding unit is not the scale of f + dt*rhs there:
    checkTests' per-cell relative difference (checkTests.C:345-358) <= 1e-12 is asserted on the bulk
    (cells above 1e-6 of the peak), and the difference relative to the cell's stencil neighbourhood
    <= 1e-12 everywhere."""
    import torch
    deck = decks.interpenetrating_streams()
    w, sp, keep = _oracle(ok, deck)
    states = [deck.initial_state(s)[0] for s in deck.species]
    ns = len(states)
    H, sys_ = _product(deck, states, None)
    rhs_d = [torch.zeros(s.shape, dtype=torch.float64, device="cuda") for s in states]
    assert H.lk_vp_eval_rhs(sys_, (C.c_void_p * ns)(*[r.data_ptr() for r in rhs_d]), 0.0) == 0
    f_old = [s.copy() for s in states]
    f_new = [np.zeros_like(s) for s in states]
    ax, ay = np.zeros(ns), np.zeros(ns)
    rhs0 = [np.zeros_like(s) for s in states]      # kept alive: the oracle writes through raw pointers
    ok.ok_vp_eval_rhs(w, _ptrs(rhs0), _ptrs(f_old), 0.0, np.zeros(ns), ax, ay)
    dt_o = deck.cfl * ok.ok_vp_stable_dt(w, ax, ay, deck.rk)
    dt_d = C.c_double()
    assert H.lk_vp_stable_dt(sys_, C.byref(dt_d)) == 0
    assert abs(dt_d.value * deck.cfl - dt_o) <= 1e-12 * dt_o
    ok.ok_vp_rk6_step(w, _ptrs(f_new), _ptrs(f_old), 0.0, dt_o, np.zeros(ns))
    assert H.lk_vp_set_time(sys_, 0.0) == 0
    assert H.lk_vp_advance(sys_, dt_o) == 0, H.lk_last_error()
    ng = deck.ng
    I = (slice(ng, -ng),) * 4
    for s in range(ns):
        out = np.empty_like(states[s])
        assert H.lk_vp_get_state(sys_, s, out.ctypes.data) == 0
        assert np.any(out[I] != states[s][I])
        big = f_new[s][I] >= 1e-6 * f_new[s][I].max()
        assert big.sum() > 0.1 * big.size
        err_bulk = cell_rel_err(out[I][big], f_new[s][I][big])
        err_star = star_rel_err(out, f_new[s], f_new[s], ng)
        assert err_bulk <= 1e-12 and err_star <= 1e-12, (deck.species[s].name, err_bulk, err_star)
    H.lk_vp_destroy(sys_)
    ok.ok_vp_work_destroy(w)


def test_several_steps_production_vs_oracle(lk, ok, fast):
    """a short run with dt from stableDt each step: distribution stays within 1e-12 per cell and the
    field-energy trace within 1e-10 (north-star tolerance for time-history traces)"""
    deck = decks.plane_epw(n=(12, 6), nv=(24, 12))
    w, sp, keep = _oracle(ok, deck)
    s0 = deck.species[0]
    f, fx, fv, fnorm = deck.initial_state(s0)
    state = _perturb(f, 5, amp=0.01)
    H, sys_ = _product(deck, [state], [(fx, fv, fnorm)])
    ng = deck.ng
    n1d, n2d = deck.n[0] + 2 * ng, deck.n[1] + 2 * ng
    f_old, f_new = state.copy(), np.zeros_like(state)
    ke = np.zeros(1)
    t = 0.0
    # the reference seeds lambda_max with a throw-away evalRHS at init (VPSystem.C:227-230)
    import torch
    rhs_d = torch.zeros(state.shape, dtype=torch.float64, device="cuda")
    assert H.lk_vp_eval_rhs(sys_, (C.c_void_p * 1)(rhs_d.data_ptr()), t) == 0
    rhs0 = np.zeros_like(state)
    ax, ay = np.zeros(1), np.zeros(1)
    ok.ok_vp_eval_rhs(w, _ptrs([rhs0]), _ptrs([f_old]), t, np.zeros(1), ax, ay)
    en_o, en_d = [], []
    for step in range(4):
        dt_o = deck.cfl * ok.ok_vp_stable_dt(w, ax, ay, deck.rk)
        dt_d = C.c_double()
        assert H.lk_vp_stable_dt(sys_, C.byref(dt_d)) == 0
        assert abs(dt_d.value * deck.cfl - dt_o) <= 1e-12 * dt_o
        ok.ok_vp_rk4_step(w, _ptrs([f_new]), _ptrs([f_old]), t, dt_o, ke)
        assert H.lk_vp_set_time(sys_, t) == 0
        assert H.lk_vp_advance(sys_, dt_o) == 0
        t += dt_o
        f_old, f_new = f_new, f_old
        # the accelerations that bound the next dt are those of the LAST STAGE of this step
        # (KineticSpecies.C:771-772), on both sides
        ok.ok_vp_last_accel_max(w, ax, ay)
        lam = (C.c_double * 2)()
        assert H.lk_vp_lambda_max(sys_, 0, C.byref(lam)) == 0
        assert abs(lam[0] - ax[0]) <= 1e-10 * ax[0] and abs(lam[1] - ay[0]) <= 1e-10 * ay[0]
        em_o = np.ctypeslib.as_array(ok.ok_vp_em_vars(w), shape=(2, n2d, n1d))
        en_o.append(float(np.sum(em_o[:, ng:-ng, ng:-ng] ** 2)))
        em_d = np.empty_like(em_o)
        assert lk.lk_sync(None) == 0
        assert lk.lk_memcpy_d2h(em_d.ctypes.data, H.lk_vp_em_vars_ptr(sys_), em_d.nbytes) == 0
        en_d.append(float(np.sum(em_d[:, ng:-ng, ng:-ng] ** 2)))
        assert abs(en_d[-1] - en_o[-1]) <= 1e-10 * en_o[-1]      # field-energy trace (tstol-style tolerance)
        out = np.empty_like(state)
        assert H.lk_vp_get_state(sys_, 0, out.ctypes.data) == 0
        assert star_rel_err(out, f_old, f_old, ng) <= 1e-12
    H.lk_vp_destroy(sys_)
    ok.ok_vp_work_destroy(w)


def _select_dt(time, dt_stable_cfl, last_save, save_times, max_time):
    """Simulation::selectTimeStep (Simulation.C:464-485): dt honours the plot frequency"""
    next_time = (last_save + 1) * save_times
    final_time = min(next_time, max_time)
    remaining = final_time - time
    if dt_stable_cfl <= remaining:
        n = int(np.ceil(remaining / dt_stable_cfl))
        return remaining / n
    return remaining * (1.0 + 10 * np.finfo(float).eps)


def test_regression_run_traces_plane_iaw(lk, ok, fast):
    """a regression-style run of the planeIAW deck (two species, driven electrons) on a reduced grid: the
    reference's own time-step selection (cfl * stableDt snapped to save_times = 0.2) to t = 0.4, production
    arithmetic on the device against the oracle.  North-star tolerances: time-history traces (field energy,
    max |Ex|, the integrated driver work ke_e_dot) within 1e-10 at every step, the distribution within 1e-12
    relative to its stencil neighbourhood at the end."""
    import torch
    deck = decks.plane_iaw(n=(12, 8), nv=(20, 12), A=0.03, ky1=1.0 / 234)
    w, sp, keep = _oracle(ok, deck)
    states = [deck.initial_state(s)[0] for s in deck.species]
    ns = len(states)
    H, sys_ = _product(deck, states, None)
    ng = deck.ng
    n1d, n2d = deck.n[0] + 2 * ng, deck.n[1] + 2 * ng
    I2 = (slice(None), slice(ng, -ng), slice(ng, -ng))
    # seeding evalRHS (VPSystem.C:227-230)
    rhs_d = [torch.zeros(s.shape, dtype=torch.float64, device="cuda") for s in states]
    assert H.lk_vp_eval_rhs(sys_, (C.c_void_p * ns)(*[r.data_ptr() for r in rhs_d]), 0.0) == 0
    f_old = [s.copy() for s in states]
    f_new = [np.zeros_like(s) for s in states]
    rhs0 = [np.zeros_like(s) for s in states]
    ax, ay = np.zeros(ns), np.zeros(ns)
    ok.ok_vp_eval_rhs(w, _ptrs(rhs0), _ptrs(f_old), 0.0, np.zeros(ns), ax, ay)
    ke = np.zeros(ns)
    vel_tables = []
    for s_ in range(ns):
        g_ = sp[s_].g
        nd_ = g_.nd
        vt = np.zeros(nd_[2] * nd_[3] * 2)
        lo_ = (C.c_int * 2)(-g_.ng, -g_.ng)
        ok.ok_build_velocity_tables(C.byref(g_), C.byref(lo_), deck.species[s_].vlim[0], deck.species[s_].vlim[2], vt,
                                    np.zeros((nd_[2] + 1) * nd_[3] * 2), np.zeros(nd_[2] * (nd_[3] + 1) * 2))
        vel_tables.append(vt)
    t, last_save, save_times, t_final = 0.0, 0, 0.2, 0.4
    nsteps = 0
    while t < t_final:
        dt_o = _select_dt(t, deck.cfl * ok.ok_vp_stable_dt(w, ax, ay, deck.rk), last_save, save_times, t_final)
        dt_d = C.c_double()
        assert H.lk_vp_stable_dt(sys_, C.byref(dt_d)) == 0
        dt_dev = _select_dt(t, deck.cfl * dt_d.value, last_save, save_times, t_final)
        assert abs(dt_dev - dt_o) <= 1e-10 * dt_o          # same number of sub-steps to the next save time
        ok.ok_vp_rk4_step(w, _ptrs(f_new), _ptrs(f_old), t, dt_o, ke)
        assert H.lk_vp_set_time(sys_, t) == 0
        assert H.lk_vp_advance(sys_, dt_o) == 0
        t += dt_o
        nsteps += 1
        if t >= (last_save + 1) * save_times:
            last_save += 1
        f_old, f_new = f_new, f_old
        ok.ok_vp_last_accel_max(w, ax, ay)
        em_o = np.ctypeslib.as_array(ok.ok_vp_em_vars(w), shape=(2, n2d, n1d))
        em_d = np.empty_like(em_o)
        assert lk.lk_sync(None) == 0
        assert lk.lk_memcpy_d2h(em_d.ctypes.data, H.lk_vp_em_vars_ptr(sys_), em_d.nbytes) == 0
        # the field of the last RK stage of the step, as both sides hold it
        e_o, e_d = float(np.sum(em_o[I2] ** 2)), float(np.sum(em_d[I2] ** 2))
        assert abs(e_d - e_o) <= 1e-10 * e_o
        m_o, m_d = float(np.max(np.abs(em_o[0][I2[1:]]))), float(np.max(np.abs(em_d[0][I2[1:]])))
        assert abs(m_d - m_o) <= 1e-10 * m_o
        v = C.c_double()
        assert H.lk_vp_ke_e_dot(sys_, 0, C.byref(v)) == 0
        assert abs(v.value - ke[0]) <= 1e-10 * abs(ke[0]) + 1e-300
        # the whole time-history record of the step (VPSystem::accumulateSequences): field histories of the
        # field both sides hold, species kinetic energies and momenta of the new state
        hist = np.zeros(5 + 6 * ns)
        nh = C.c_int()
        assert H.lk_vp_time_history(sys_, hist.ctypes.data, hist.size, C.byref(nh)) == 0 and nh.value == hist.size
        fh = np.zeros(12)
        ok.ok_field_history(np.ascontiguousarray(em_o).ravel(), deck.n[0], deck.n[1], ng, 2, np.array(deck.dx + (1.0, 1.0)), fh)
        # e_max, e_tot, ex_max, e_sum_tot within 1e-10; ey_max is 1 % of ex_max in this deck and is made of
        # the white rounding noise of the net charge density (electron and ion densities of size 1 cancel to
        # 1e-5 and the device sums them as trees; a mode of wavelength L_y = 1470 turns a density rounding of
        # 1e-16 into a field of 1e-16 * L_y / 2 pi = 2e-14): bounded absolutely, in the deck's units
        for k_ in (0, 1, 2, 4):
            assert abs(hist[k_] - fh[k_]) <= 1e-10 * abs(fh[k_])
        assert abs(hist[3] - fh[3]) <= 1e-12
        for s_ in range(ns):
            sp_ = deck.species[s_]
            o5 = np.zeros(5)
            ok.ok_compute_ke(C.byref(sp[s_].g), f_old[s_].ravel(), sp_.mass, vel_tables[s_], o5)
            got = hist[5 + 6 * s_: 5 + 6 * s_ + 5]
            assert np.all(np.abs(got[:3] - o5[:3]) <= 1e-10 * np.abs(o5[:3]))                 # energies
            assert np.all(np.abs(got[3:] - o5[3:]) <= 1e-10 * (sp_.mass * np.sqrt(2 * o5[0] / sp_.mass)))  # momenta ~ 0
        assert abs(hist[5 + 5] - ke[0]) <= 1e-10 * abs(ke[0]) + 1e-300
    assert nsteps >= 4 and abs(t - t_final) < 1e-12
    for s in range(ns):
        out = np.empty_like(states[s])
        assert H.lk_vp_get_state(sys_, s, out.ctypes.data) == 0
        assert star_rel_err(out, f_old[s], f_old[s], ng) <= 1e-12
    H.lk_vp_destroy(sys_)
    ok.ok_vp_work_destroy(w)


RUN_DECK = """
# a small two-species deck in the reference's .pp syntax, written for this test
$pi = 3.1415926535897932384626;
$klde = 1/3;
$xa = -$pi/$klde;
$xb =  $pi/$klde;
domain_limits = $xa $xb -30. 30.
N = 12 6
periodic_dir = true true
cfl = 0.9
final_time = 0.3
save_times = 0.1
max_step = 100
number_of_species = 2
kinetic_species.1.name = "electron"
kinetic_species.1.velocity_limits = -7 7 -7 7
kinetic_species.1.Nv = 20 12
kinetic_species.1.mass = 1.0
kinetic_species.1.charge = -1.0
kinetic_species.1.ic.name = "Perturbed Maxwellian"
kinetic_species.1.ic.A = 0.02
kinetic_species.1.ic.kx1 = $klde
kinetic_species.1.ic.ky1 = 0.0
kinetic_species.1.num_external_drivers = 1
kinetic_species.1.external_driver.1.name = "Shaped Ramped Cosine Driver"
kinetic_species.1.external_driver.1.xwidth = 9.42477796076938
kinetic_species.1.external_driver.1.ywidth = 200
kinetic_species.1.external_driver.1.shape = 0.0
kinetic_species.1.external_driver.1.omega = 1.1
kinetic_species.1.external_driver.1.E_0 = 0.05
kinetic_species.1.external_driver.1.t_ramp = 1.0
kinetic_species.1.external_driver.1.t_off = 2.0
kinetic_species.1.external_driver.1.lwidth = 50
kinetic_species.2.name = "ion"
kinetic_species.2.velocity_limits = -0.4 0.4 -0.4 0.4
kinetic_species.2.Nv = 16 12
kinetic_species.2.mass = 100.0
kinetic_species.2.charge = 1.0
kinetic_species.2.ic.name = "Perturbed Maxwellian"
kinetic_species.2.ic.tx = 0.1
kinetic_species.2.ic.ty = 0.1
"""


def test_run_deck_file_end_to_end(lk, ok, fast, tmp_path):
    """deck text -> loki_b200.pp -> loki_b200.run.Runner (Simulation::advance loop on the device) against the
    same loop driven through the oracle: times, dt sequence, final distribution and kinetic energies"""
    from loki_b200 import pp, run
    path = tmp_path / "two_species.pp"
    path.write_text(RUN_DECK)
    deck = decks._wrap(pp.load(str(path)))
    r = run.Runner(deck)
    w, sp, keep = _oracle(ok, deck)
    ns = len(deck.species)
    f_old = [deck.initial_state(s)[0] for s in deck.species]
    f_new = [np.zeros_like(f) for f in f_old]
    rhs0 = [np.zeros_like(f) for f in f_old]
    ax, ay = np.zeros(ns), np.zeros(ns)
    ok.ok_vp_eval_rhs(w, _ptrs(rhs0), _ptrs(f_old), 0.0, np.zeros(ns), ax, ay)
    ke = np.zeros(ns)
    t, last_save = 0.0, 0
    while not r.done():
        dt_o = run.select_dt(t, deck.cfl * ok.ok_vp_stable_dt(w, ax, ay, deck.rk), last_save, 0.1, 0.3)
        dt_d = r.advance()
        assert abs(dt_d - dt_o) <= 1e-10 * dt_o
        ok.ok_vp_rk4_step(w, _ptrs(f_new), _ptrs(f_old), t, dt_o, ke)
        t += dt_o
        if t >= (last_save + 1) * 0.1:      # exact, as Simulation::advance (Simulation.C:323-326)
            last_save += 1
        f_old, f_new = f_new, f_old
        ok.ok_vp_last_accel_max(w, ax, ay)
        hist = r.history()
        for s_ in range(ns):
            o5 = np.zeros(5)
            vt = np.zeros(sp[s_].g.nd[2] * sp[s_].g.nd[3] * 2)
            lo_ = (C.c_int * 2)(-deck.ng, -deck.ng)
            nd_ = sp[s_].g.nd
            ok.ok_build_velocity_tables(C.byref(sp[s_].g), C.byref(lo_), deck.species[s_].vlim[0], deck.species[s_].vlim[2], vt,
                                        np.zeros((nd_[2] + 1) * nd_[3] * 2), np.zeros(nd_[2] * (nd_[3] + 1) * 2))
            ok.ok_compute_ke(C.byref(sp[s_].g), f_old[s_].ravel(), deck.species[s_].mass, vt, o5)
            assert abs(hist[5 + 6 * s_] - o5[0]) <= 1e-10 * o5[0]
    assert r.step >= 3 and abs(r.time - 0.3) < 1e-12 and abs(t - 0.3) < 1e-12
    ng = deck.ng
    for s_ in range(ns):
        assert star_rel_err(r.state(s_), f_old[s_], f_old[s_], ng) <= 1e-12
    r.close()
    ok.ok_vp_work_destroy(w)


def test_full_regression_run_plane_epw_reduced_grid(lk, ok, fast):
    """the planeEPW_fixedIons regression run in full length (final_time = 5, save_times = 1, cfl = 1, the
    reference's time-step selection) on a reduced grid, production arithmetic on the device against the oracle:
    the north-star criterion, 1e-10 on the time-history energy and field traces at EVERY step of the run, and
    the distribution within 1e-10 (checkTests' per-cell metric on the bulk) at the end"""
    from loki_b200 import run
    deck = decks.plane_epw(n=(16, 8), nv=(32, 16))
    deck.run = dict(final_time=5.0, save_times=1.0, max_step=1000000)
    r = run.Runner(deck)
    w, sp, keep = _oracle(ok, deck)
    f_old = [deck.initial_state(deck.species[0])[0]]
    f_new = [np.zeros_like(f_old[0])]
    rhs0 = [np.zeros_like(f_old[0])]
    ax, ay = np.zeros(1), np.zeros(1)
    ok.ok_vp_eval_rhs(w, _ptrs(rhs0), _ptrs(f_old), 0.0, np.zeros(1), ax, ay)
    ng = deck.ng
    n1d, n2d = deck.n[0] + 2 * ng, deck.n[1] + 2 * ng
    g_ = sp[0].g
    nd_ = g_.nd
    vt = np.zeros(nd_[2] * nd_[3] * 2)
    lo_ = (C.c_int * 2)(-ng, -ng)
    ok.ok_build_velocity_tables(C.byref(g_), C.byref(lo_), deck.species[0].vlim[0], deck.species[0].vlim[2], vt,
                                np.zeros((nd_[2] + 1) * nd_[3] * 2), np.zeros(nd_[2] * (nd_[3] + 1) * 2))
    ke = np.zeros(1)
    t, last_save = 0.0, 0
    dev_tr, ora_tr = [], []
    while not r.done():
        dt_o = run.select_dt(t, deck.cfl * ok.ok_vp_stable_dt(w, ax, ay, deck.rk), last_save, 1.0, 5.0)
        dt_d = r.advance()
        assert abs(dt_d - dt_o) <= 1e-10 * dt_o
        ok.ok_vp_rk4_step(w, _ptrs(f_new), _ptrs(f_old), t, dt_o, ke)
        t += dt_o
        if t >= (last_save + 1) * 1.0:
            last_save += 1
        f_old, f_new = f_new, f_old
        ok.ok_vp_last_accel_max(w, ax, ay)
        hist = r.history()
        em_o = np.ctypeslib.as_array(ok.ok_vp_em_vars(w), shape=(2, n2d, n1d))
        fh = np.zeros(12)
        ok.ok_field_history(np.ascontiguousarray(em_o).ravel(), deck.n[0], deck.n[1], ng, 2, np.array(deck.dx + (1.0, 1.0)), fh)
        o5 = np.zeros(5)
        ok.ok_compute_ke(C.byref(g_), f_old[0].ravel(), deck.species[0].mass, vt, o5)
        # field: e_max, e_tot, ex_max, e_sum_tot (the deck is uniform in y: Ey is rounding noise); species: ke,
        # ke_x, ke_y and the integrated driver work
        dev_tr.append([hist[0], hist[1], hist[2], hist[4], hist[5], hist[6], hist[7], hist[10]])
        ora_tr.append([fh[0], fh[1], fh[2], fh[4], o5[0], o5[1], o5[2], ke[0]])
    dev_tr, ora_tr = np.array(dev_tr), np.array(ora_tr)
    # The run starts spatially uniform: for the first steps the self-consistent field is the response to a
    # driver that is still 2e-4 of its amplitude, 1e-8 of the density, and carries the summation-order noise
    # of the charge density at the 1e-7 level (the reference has the same sensitivity to its rank count).
    # Each trace is therefore compared in the norm of the whole run: max_t |difference| / max_t |trace|.
    worst = np.max(np.abs(dev_tr - ora_tr), axis=0) / np.max(np.abs(ora_tr), axis=0)
    assert np.all(worst <= 1e-10), worst
    # and once the field has grown (second half of the run) sample by sample as well
    half = len(ora_tr) // 2
    late = np.abs(dev_tr[half:] - ora_tr[half:]) / np.abs(ora_tr[half:])
    assert np.all(late <= 1e-9), late.max(axis=0)
    worst = float(worst.max())
    assert r.step >= 30 and abs(r.time - 5.0) < 1e-9
    out = r.state(0)
    I = (slice(ng, -ng),) * 4
    big = f_old[0][I] >= 1e-6 * f_old[0][I].max()
    assert cell_rel_err(out[I][big], f_old[0][I][big]) <= 1e-10
    print("planeEPW regression run: %d steps, worst trace difference %.2e" % (r.step, worst))
    r.close()
    ok.ok_vp_work_destroy(w)


def _full_run_vs_oracle(ok, deck, final_time, save_times, flux_traces=False):
    """Simulation::advance's loop on the device (loki_b200.run.Runner) and through the oracle, side by side;
    returns (steps, device traces, oracle traces, device states, oracle states).  Traces per step: e_max,
    e_tot, ex_max, e_sum_tot, then ke of every species."""
    from loki_b200 import run
    deck.run = dict(final_time=final_time, save_times=save_times, max_step=1000000)
    r = run.Runner(deck)
    w, sp, keep = _oracle(ok, deck)
    ns = len(deck.species)
    f_old = [deck.initial_state(s)[0] for s in deck.species]
    f_new = [np.zeros_like(f) for f in f_old]
    rhs0 = [np.zeros_like(f) for f in f_old]
    ax, ay = np.zeros(ns), np.zeros(ns)
    ok.ok_vp_eval_rhs(w, _ptrs(rhs0), _ptrs(f_old), 0.0, np.zeros(ns), ax, ay)
    ng = deck.ng
    n1d, n2d = deck.n[0] + 2 * ng, deck.n[1] + 2 * ng
    vts = []
    for s_ in range(ns):
        nd_ = sp[s_].g.nd
        vt = np.zeros(nd_[2] * nd_[3] * 2)
        lo_ = (C.c_int * 2)(-ng, -ng)
        ok.ok_build_velocity_tables(C.byref(sp[s_].g), C.byref(lo_), deck.species[s_].vlim[0], deck.species[s_].vlim[2], vt,
                                    np.zeros((nd_[2] + 1) * nd_[3] * 2), np.zeros(nd_[2] * (nd_[3] + 1) * 2))
        vts.append(vt)
    ke = np.zeros(ns)
    step_fn = ok.ok_vp_rk4_step if deck.rk == 4 else ok.ok_vp_rk6_step
    t, last_save = 0.0, 0
    dev_tr, ora_tr = [], []
    while not r.done():
        dt_o = run.select_dt(t, deck.cfl * ok.ok_vp_stable_dt(w, ax, ay, deck.rk), last_save, save_times, final_time)
        dt_d = r.advance()
        assert abs(dt_d - dt_o) <= 1e-9 * dt_o
        step_fn(w, _ptrs(f_new), _ptrs(f_old), t, dt_o, ke)
        t += dt_o
        if t >= (last_save + 1) * save_times:
            last_save += 1
        f_old, f_new = f_new, f_old
        ok.ok_vp_last_accel_max(w, ax, ay)
        hist = r.history()
        em_o = np.ctypeslib.as_array(ok.ok_vp_em_vars(w), shape=(2, n2d, n1d))
        fh = np.zeros(12)
        ok.ok_field_history(np.ascontiguousarray(em_o).ravel(), deck.n[0], deck.n[1], ng, 2, np.array(deck.dx + (1.0, 1.0)), fh)
        d_row, o_row = [hist[0], hist[1], hist[2], hist[4]], [fh[0], fh[1], fh[2], fh[4]]
        for s_ in range(ns):
            o5 = np.zeros(5)
            ok.ok_compute_ke(C.byref(sp[s_].g), f_old[s_].ravel(), deck.species[s_].mass, vts[s_], o5)
            d_row.append(hist[5 + 6 * s_])
            o_row.append(o5[0])
        if flux_traces:
            # the eight boundary kinetic-energy fluxes per species (KineticSpecies.C:2052-2097), both sides with the
            # face accelerations of the step's last stage
            fo = np.zeros(8 * ns)
            ok.ok_vp_ke_flux_history(w, _ptrs(f_old), fo)
            d_row += list(r.flux_history())
            o_row += list(fo)
        dev_tr.append(d_row)
        ora_tr.append(o_row)
    assert abs(r.time - final_time) < 1e-9
    states = [r.state(s_) for s_ in range(ns)]
    steps = r.step
    r.close()
    ok.ok_vp_work_destroy(w)
    return steps, np.array(dev_tr), np.array(ora_tr), states, f_old


FULL_RUNS = [
    # (deck, final_time, save_times, minimum number of steps)
    (lambda: decks.plane_iaw(n=(16, 8), nv=(24, 16)), 5.0, 1.0, 30),                       # planeIAW, reduced grid
    (lambda: decks.plane_iaw(n=(10, 10), nv=(16, 10), order=6, rk=6), 5.0, 1.0, 10),        # planeIAW_6, the deck's grid
    (lambda: decks.interpenetrating_streams(n=(32, 7), nv=(16, 12)), 5.0, 1.0, 10),         # InterpenetratingStreams
]


def test_flux_histories_along_a_deck_run(lk, ok, fast):
    """the `*_flux` time histories along planeIAW (reduced grid, both species, final_time = 2): 16 more traces, each
    within 1e-10 of the oracle's in the norm of the run.  The x / y fluxes are odd in the normal velocity and cancel
    to a fraction of their terms; every trace is held relative to the largest flux trace of its species."""
    deck = decks.plane_iaw(n=(16, 8), nv=(24, 16))
    steps, dev_tr, ora_tr, got, want = _full_run_vs_oracle(ok, deck, 2.0, 1.0, flux_traces=True)
    ns = len(deck.species)
    base = dev_tr.shape[1] - 8 * ns
    assert steps >= 10 and base == 4 + ns
    for s_ in range(ns):
        d_, o_ = dev_tr[:, base + 8 * s_: base + 8 * s_ + 8], ora_tr[:, base + 8 * s_: base + 8 * s_ + 8]
        scale = np.max(np.abs(o_))
        assert scale > 0 and np.all(np.max(np.abs(o_), axis=0) > 0)
        assert np.max(np.abs(d_ - o_)) <= 1e-10 * scale, np.max(np.abs(d_ - o_), axis=0) / scale


@pytest.mark.parametrize("mk,final_time,save_times,min_steps", FULL_RUNS)
def test_full_regression_runs_vp(lk, ok, fast, mk, final_time, save_times, min_steps):
    """the Vlasov-Poisson regression decks run in full length (final_time = 5) with the reference's time-step
    selection, production arithmetic on the device against the oracle: every time-history trace within 1e-10
    in the norm of the run (north-star tolerance), the distribution within 1e-10 of its stencil neighbourhood
    at the end"""
    deck = mk()
    steps, dev_tr, ora_tr, got, want = _full_run_vs_oracle(ok, deck, final_time, save_times)
    assert steps >= min_steps
    worst = np.max(np.abs(dev_tr - ora_tr), axis=0) / np.max(np.abs(ora_tr), axis=0)
    assert np.all(worst <= 1e-10), worst
    for s_ in range(len(got)):
        assert star_rel_err(got[s_], want[s_], want[s_], deck.ng) <= 1e-10
    print("%s regression run: %d steps, worst trace difference %.2e" % (deck.name, steps, float(worst.max())))


def test_streaming_state_equals_blocking_state_io(lk, ok, fast):
    """lk_vp_download_state / lk_vp_upload_next / lk_vp_adopt_next (the double-buffered host streaming bench.py's e2e
    leg uses): every step takes its input from pinned host memory and returns its result there, uploads and downloads
    in flight together -- same bits as lk_vp_set_state + lk_vp_advance + lk_vp_get_state"""
    import torch
    deck = decks.plane_iaw(n=(32, 8), nv=(16, 8))
    states = [_perturb(deck.initial_state(s)[0], 90 + k, amp=0.02) for k, s in enumerate(deck.species)]
    ns = len(states)
    nsteps, dt = 4, 0.02
    # every step starts from a DIFFERENT host state (a scaled copy), as if a producer filled the buffer
    inputs = [[np.ascontiguousarray(f * (1.0 + 0.01 * k)) for f in states] for k in range(nsteps)]
    H, sys_ = _product(deck, states, None)
    want = []
    for k in range(nsteps):
        for s in range(ns):
            assert H.lk_vp_set_state(sys_, s, inputs[k][s].ctypes.data) == 0
        assert H.lk_vp_set_time(sys_, 0.1) == 0 and H.lk_vp_advance(sys_, dt) == 0
        outs = [np.empty_like(f) for f in states]
        for s in range(ns):
            assert H.lk_vp_get_state(sys_, s, outs[s].ctypes.data) == 0
        want.append(outs)
    H.lk_vp_destroy(sys_)
    H, sys_ = _product(deck, states, None)
    up, down = torch.cuda.Stream(), torch.cuda.Stream()
    pin_in = [[torch.from_numpy(a).pin_memory() for a in step] for step in inputs]
    pin_out = [[torch.empty(f.size, dtype=torch.float64).pin_memory() for f in states] for _ in range(nsteps)]
    for s in range(ns):
        assert H.lk_vp_upload_next(sys_, s, pin_in[0][s].data_ptr(), C.c_void_p(up.cuda_stream)) == 0
    assert H.lk_vp_adopt_next(sys_) == 0
    for k in range(nsteps):
        assert H.lk_vp_set_time(sys_, 0.1) == 0 and H.lk_vp_advance(sys_, dt) == 0, H.lk_last_error()
        for s in range(ns):
            assert H.lk_vp_download_state(sys_, s, pin_out[k][s].data_ptr(), C.c_void_p(down.cuda_stream)) == 0
        if k + 1 < nsteps:
            for s in range(ns):
                assert H.lk_vp_upload_next(sys_, s, pin_in[k + 1][s].data_ptr(), C.c_void_p(up.cuda_stream)) == 0
            # a step must not start over an upload that has not been adopted
            assert H.lk_vp_advance(sys_, dt) != 0
            assert H.lk_vp_adopt_next(sys_) == 0
    torch.cuda.synchronize()
    ng = deck.ng
    I = (slice(ng, -ng),) * 4
    for k in range(nsteps):
        for s in range(ns):
            got = pin_out[k][s].numpy().reshape(states[s].shape)
            assert np.array_equal(got[I], want[k][s][I]), (k, s)
    H.lk_vp_destroy(sys_)


@pytest.mark.parametrize("mode", ["strict", "production"])
def test_probe_histories(lk, ok, mode):
    """the probe time histories (Poisson.C:852-887): E of the last field solve at the cell floor(frac * N) of every
    probe, against the oracle's field -- strict: the same bits, production: 1e-8 of the field's maximum"""
    from loki_b200 import run
    deck = decks.plane_iaw(n=(16, 8), nv=(24, 16))
    deck.probes = [(0.5, 0.0), (0.26, 0.51), (0.999, 0.999), (0.0, 0.3)]
    deck.run = dict(final_time=1.0, save_times=1.0, max_step=1000000)
    old = lk.lk_set_strict(1 if mode == "strict" else 0)
    try:
        r = run.Runner(deck)
        w, sp, keep = _oracle(ok, deck)
        ns = len(deck.species)
        f_old = [deck.initial_state(s)[0] for s in deck.species]
        f_new = [np.zeros_like(f) for f in f_old]
        ng = deck.ng
        n1d, n2d = deck.n[0] + 2 * ng, deck.n[1] + 2 * ng
        ke = np.zeros(ns)
        t, dt = 0.0, 0.05
        for _ in range(3):
            capi_ok = r.H.lk_vp_set_time(r.sys, t) == 0 and r.H.lk_vp_advance(r.sys, dt) == 0
            assert capi_ok
            ok.ok_vp_rk4_step(w, _ptrs(f_new), _ptrs(f_old), t, dt, ke)
            f_old, f_new = f_new, f_old
            t += dt
        em_o = np.ctypeslib.as_array(ok.ok_vp_em_vars(w), shape=(2, n2d, n1d))
        got = r.probes()
        assert got.shape == (4, 2)
        scale = np.max(np.abs(em_o))
        for k, (fx, fy) in enumerate(deck.probes):
            ip, jp = int(np.floor(fx * deck.n[0])), int(np.floor(fy * deck.n[1]))
            want = em_o[:, jp + ng, ip + ng]
            if mode == "strict":
                assert got[k, 0] == want[0] and got[k, 1] == want[1], k
            else:
                # three steps into the run the field is still the response to a driver at 1e-3 of its amplitude: it carries
                # the summation-order noise of a charge density that cancels to 1e-10 of its terms (the short-run note of
                # tests/test_gpu_deck_grids.py); the regression-run tests hold the field traces to 1e-10
                assert np.max(np.abs(got[k] - want)) <= 1e-8 * scale, k
        assert np.any(got != 0.0)
        r.close()
        ok.ok_vp_work_destroy(w)
    finally:
        lk.lk_set_strict(old)
