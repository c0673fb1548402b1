"""System-level parity of the Vlasov-Maxwell path: the C++ host mirror (lk_vm_*: VMSystem / Maxwell /
RK4Integrator on the device) against the oracle's single-rank restatement of VMSystem::evalRHS
(VMSystem.C:407-549) and RK4Integrator over a VMState, on test/emDamping-shaped problems (Lorentz
acceleration, current densities, Maxwell RHS with avStrong hyper-dissipation, non-factorable inflow IC)."""
import ctypes as C

import numpy as np
import pytest

import decks

from util import cell_rel_err, star_rel_err

pytestmark = pytest.mark.gpu


def _ptrs(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def _oracle(ok, deck):
    keep = []
    sp = deck.oracle_species(keep)
    xlo = (C.c_double * 2)(deck.xlim[0], deck.xlim[2])
    xhi = (C.c_double * 2)(deck.xlim[1], deck.xlim[3])
    w = ok.ok_vm_work_create(len(deck.species), sp, C.byref(xlo), C.byref(xhi), deck.light_speed, deck.av_weak,
                             deck.av_strong)
    return w, sp, keep


def _product(deck, states, em, vz):
    from loki_b200 import host
    H = host.lib()
    d = deck.product_vm_desc()
    sys_ = C.c_void_p()
    assert H.lk_vm_create(C.byref(sys_), C.byref(d), None) == 0, H.lk_last_error()
    for s, f in enumerate(states):
        assert H.lk_vm_set_state(sys_, s, f.ctypes.data) == 0
        sp = deck.species[s]
        fic, fx, fv, fnorm = deck.initial_state(sp)
        if sp.factorable:
            assert H.lk_vm_set_inflow(sys_, s, fx.ctypes.data, fv.ctypes.data, fnorm, sp.frac) == 0
        else:
            g3, g4 = deck.inflow_ghost_tables(fic)
            assert H.lk_vm_set_inflow_ghosts(sys_, s, g3.ctypes.data, g4.ctypes.data) == 0
        assert H.lk_vm_set_vz(sys_, s, vz[s].ctypes.data) == 0
    assert H.lk_vm_set_fields(sys_, em.ctypes.data) == 0
    return H, sys_


def _setup(deck, seed, amp, em_scale=1.0, vz_amp=0.0):
    rng = np.random.default_rng(seed)
    states = []
    for sp in deck.species:
        f = deck.initial_state(sp)[0]
        states.append(np.ascontiguousarray(f * (1.0 + amp * rng.uniform(-1, 1, size=f.shape))))
    em, vz = deck.initial_fields()
    # exercise all six field components and the transverse drift, not only the deck's Ey/Bz wave
    em = np.ascontiguousarray(em * em_scale + vz_amp * 0.01 * rng.uniform(-1, 1, size=em.shape))
    vz = [np.ascontiguousarray(v + vz_amp * rng.uniform(-1, 1, size=v.shape)) for v in vz]
    return states, em, vz


DECKS = [
    lambda: decks.em_damping(n=(12, 5), nv=(16, 12)),
    lambda: decks.em_damping(n=(32, 5), nv=(24, 24)),       # the deck's own configuration-space grid
    lambda: decks.em_damping(n=(10, 8), nv=(16, 10), order=6),
]
# the sixth-order integrator over the whole VMState (RK6Integrator.H:69-133), with both spatial orders
RK6_DECKS = [
    lambda: decks.em_damping(n=(10, 8), nv=(16, 10), order=6, rk=6),
    lambda: decks.em_damping(n=(12, 5), nv=(16, 12), order=4, rk=6),
]


@pytest.mark.parametrize("mk", DECKS)
def test_vm_eval_rhs_strict_matches_reference_order(lk, ok, strict, mk):
    """VMSystem::evalRHS unfused and in the reference's order: currents, Maxwell RHS, dvz/dt, the Vlasov
    RHS with the Lorentz acceleration and lambda_max, all bit for bit"""
    import torch
    deck = mk()
    w, sp, keep = _oracle(ok, deck)
    states, em, vz = _setup(deck, 3, 0.05, em_scale=50.0, vz_amp=0.2)
    ns = len(states)
    f_o = [s.copy() for s in states]
    em_o, vz_o = em.copy(), [v.copy() for v in vz]
    rhs_o = [np.zeros_like(s) for s in states]
    rem_o, rvz_o = np.zeros_like(em), [np.zeros_like(v) for v in vz]
    ax, ay = np.zeros(ns), np.zeros(ns)
    ok.ok_vm_eval_rhs(w, _ptrs(rhs_o), rem_o, _ptrs(rvz_o), _ptrs(f_o), em_o, _ptrs(vz_o), 0.3, ax, ay)
    H, sys_ = _product(deck, states, em, vz)
    rhs_d = [torch.zeros(s.shape, dtype=torch.float64, device="cuda") for s in states]
    rem_d = torch.zeros(em.shape, dtype=torch.float64, device="cuda")
    rvz_d = [torch.zeros(v.shape, dtype=torch.float64, device="cuda") for v in vz]
    p4 = (C.c_void_p * ns)(*[r.data_ptr() for r in rhs_d])
    pz = (C.c_void_p * ns)(*[r.data_ptr() for r in rvz_d])
    assert H.lk_vm_eval_rhs(sys_, p4, rem_d.data_ptr(), pz, 0.3) == 0, H.lk_last_error()
    torch.cuda.synchronize()
    ng = deck.ng
    n1d, n2d = deck.n[0] + 2 * ng, deck.n[1] + 2 * ng
    I2 = (slice(ng, -ng), slice(ng, -ng))
    for comp in range(3):
        J_o = np.ctypeslib.as_array(ok.ok_vm_net_current(w, comp), shape=(n2d, n1d))
        J_d = np.empty_like(J_o)
        assert lk.lk_memcpy_d2h(J_d.ctypes.data, H.lk_vm_current_ptr(sys_, comp), J_d.nbytes) == 0
        assert np.any(J_o[I2] != 0.0)
        assert np.array_equal(J_d[I2], J_o[I2])
    assert np.array_equal(rem_d.cpu().numpy(), rem_o)
    assert np.any(rem_o != 0.0)
    em_back = np.empty_like(em)
    assert H.lk_vm_get_fields(sys_, em_back.ctypes.data) == 0
    assert np.array_equal(em_back, em_o)                      # ghosts refreshed by Maxwell::fillGhostCells
    for s in range(ns):
        assert np.array_equal(rvz_d[s].cpu().numpy(), rvz_o[s])
        assert np.array_equal(rhs_d[s].cpu().numpy(), rhs_o[s])
        lam = (C.c_double * 2)()
        assert H.lk_vm_lambda_max(sys_, s, C.byref(lam)) == 0
        assert (lam[0], lam[1]) == (ax[s], ay[s])
    dt_o = ok.ok_vm_stable_dt(w, ax, ay, 4)
    dt_d = C.c_double()
    assert H.lk_vm_stable_dt(sys_, C.byref(dt_d)) == 0
    assert dt_d.value == dt_o
    H.lk_vm_destroy(sys_)
    ok.ok_vm_work_destroy(w)


@pytest.mark.parametrize("mode", ["strict", "production"])
@pytest.mark.parametrize("mk", DECKS + RK6_DECKS)
def test_vm_one_step_matches_oracle(lk, ok, mk, mode):
    """one RK4 (or RK6) step of the whole VMState.  Strict arithmetic: distribution, fields and vz bit-identical.
    Production arithmetic (fused stage kernel, currents from its velocity moments): <= 1e-12 per cell on
    the distribution, 1e-12 relative on the fields."""
    deck = mk()
    old = lk.lk_set_strict(1 if mode == "strict" else 0)
    try:
        w, sp, keep = _oracle(ok, deck)
        states, em, vz = _setup(deck, 7, 0.02, em_scale=50.0, vz_amp=0.2)
        ns = len(states)
        t0, dt = 0.1, 0.01
        f_old = [s.copy() for s in states]
        f_new = [np.zeros_like(s) for s in states]
        em_old, em_new = em.copy(), np.zeros_like(em)
        vz_old, vz_new = [v.copy() for v in vz], [np.zeros_like(v) for v in vz]
        step = ok.ok_vm_rk6_step if deck.rk == 6 else ok.ok_vm_rk4_step
        step(w, _ptrs(f_new), _ptrs(f_old), em_new, em_old, _ptrs(vz_new), _ptrs(vz_old), t0, dt)
        H, sys_ = _product(deck, states, em, vz)
        assert H.lk_vm_set_time(sys_, t0) == 0
        assert H.lk_vm_advance(sys_, dt) == 0, H.lk_last_error()
        ng = deck.ng
        I = (slice(ng, -ng),) * 4
        I2 = (slice(None), slice(ng, -ng), slice(ng, -ng))
        em_d = np.empty_like(em)
        assert H.lk_vm_get_fields(sys_, em_d.ctypes.data) == 0
        assert np.any(em_d[I2] != em[I2])
        for s in range(ns):
            out = np.empty_like(states[s])
            assert H.lk_vm_get_state(sys_, s, out.ctypes.data) == 0
            vz_d = np.empty_like(vz[s])
            assert H.lk_vm_get_vz(sys_, s, vz_d.ctypes.data) == 0
            assert np.any(out[I] != states[s][I])
            if mode == "strict":
                assert np.array_equal(out[I], f_new[s][I])
                assert np.array_equal(vz_d[I2[1:]], vz_new[s][I2[1:]])
            else:
                assert star_rel_err(out, f_new[s], np.maximum(np.abs(states[s]), np.abs(f_new[s])), ng) <= 1e-12
                assert np.max(np.abs(vz_d[I2[1:]] - vz_new[s][I2[1:]])) <= 1e-12 * np.max(np.abs(vz_new[s]))
        if mode == "strict":
            assert np.array_equal(em_d[I2], em_new[I2])
        else:
            for c in range(6):
                scale = np.max(np.abs(em_new[c]))
                assert np.max(np.abs(em_d[c][I2[1:]] - em_new[c][I2[1:]])) <= 1e-12 * scale
        assert abs(H.lk_vm_time(sys_) - (t0 + dt)) < 1e-15
        H.lk_vm_destroy(sys_)
        ok.ok_vm_work_destroy(w)
    finally:
        lk.lk_set_strict(old)


def test_vm_deck_run_traces(lk, ok, fast):
    """the emDamping deck's own initial condition (unperturbed), several steps with dt = cfl*stableDt:
    distribution within 1e-12 per cell (checkTests.C:345-358 metric), field-energy trace within 1e-10"""
    deck = decks.em_damping(n=(32, 5), nv=(32, 32))
    w, sp, keep = _oracle(ok, deck)
    states, em, vz = _setup(deck, 0, 0.0)
    H, sys_ = _product(deck, states, em, vz)
    ng = deck.ng
    I = (slice(ng, -ng),) * 4
    I2 = (slice(None), slice(ng, -ng), slice(ng, -ng))
    f_old, f_new = [states[0].copy()], [np.zeros_like(states[0])]
    em_old, em_new = em.copy(), np.zeros_like(em)
    vz_old, vz_new = [vz[0].copy()], [np.zeros_like(vz[0])]
    # throw-away evalRHS that seeds lambda_max (VMSystem.C:262-265)
    import torch
    rhs_d = torch.zeros(states[0].shape, dtype=torch.float64, device="cuda")
    rem_d = torch.zeros(em.shape, dtype=torch.float64, device="cuda")
    rvz_d = torch.zeros(vz[0].shape, dtype=torch.float64, device="cuda")
    assert H.lk_vm_eval_rhs(sys_, (C.c_void_p * 1)(rhs_d.data_ptr()), rem_d.data_ptr(), (C.c_void_p * 1)(rvz_d.data_ptr()), 0.0) == 0
    ax, ay = np.zeros(1), np.zeros(1)
    rhs0, rvz0 = [np.zeros_like(states[0])], [np.zeros_like(vz[0])]   # kept alive: written through raw pointers
    ok.ok_vm_eval_rhs(w, _ptrs(rhs0), np.zeros_like(em), _ptrs(rvz0), _ptrs(f_old), em_old, _ptrs(vz_old), 0.0, ax, ay)
    g_ = sp[0].g
    nd_ = g_.nd
    vel_table = np.zeros(nd_[2] * nd_[3] * 2)
    lo_ = (C.c_int * 2)(-ng, -ng)
    ok.ok_build_velocity_tables(C.byref(g_), C.byref(lo_), deck.species[0].vlim[0], deck.species[0].vlim[2], vel_table,
                                np.zeros((nd_[2] + 1) * nd_[3] * 2), np.zeros(nd_[2] * (nd_[3] + 1) * 2))
    t = 0.0
    for step in range(4):
        dt_o = deck.cfl * ok.ok_vm_stable_dt(w, ax, ay, 4)
        dt_d = C.c_double()
        assert H.lk_vm_stable_dt(sys_, C.byref(dt_d)) == 0
        assert abs(dt_d.value * deck.cfl - dt_o) <= 1e-12 * dt_o
        ok.ok_vm_rk4_step(w, _ptrs(f_new), _ptrs(f_old), em_new, em_old, _ptrs(vz_new), _ptrs(vz_old), t, dt_o)
        assert H.lk_vm_set_time(sys_, t) == 0
        assert H.lk_vm_advance(sys_, dt_o) == 0
        t += dt_o
        f_old, f_new = f_new, f_old
        em_old, em_new = em_new, em_old
        vz_old, vz_new = vz_new, vz_old
        ok.ok_vm_last_accel_max(w, ax, ay)
        out = np.empty_like(states[0])
        assert H.lk_vm_get_state(sys_, 0, out.ctypes.data) == 0
        big = f_old[0][I] >= 1e-20 * f_old[0][I].max()
        assert cell_rel_err(out[I][big], f_old[0][I][big]) <= 1e-12
        em_d = np.empty_like(em)
        assert H.lk_vm_get_fields(sys_, em_d.ctypes.data) == 0
        e_o, e_d = float(np.sum(em_old[I2] ** 2)), float(np.sum(em_d[I2] ** 2))
        assert abs(e_d - e_o) <= 1e-10 * e_o
        # the time-history record (Maxwell::accumulateSequences + computekemaxwell) of the new state
        hist = np.zeros(12 + 5)
        nh = C.c_int()
        assert H.lk_vm_time_history(sys_, hist.ctypes.data, hist.size, C.byref(nh)) == 0 and nh.value == 17
        fh = np.zeros(12)
        ok.ok_field_history(np.ascontiguousarray(em_old).ravel(), deck.n[0], deck.n[1], ng, 6, np.array(deck.dx + (1.0, 1.0)), fh)
        # the wave lives in Ey and Bz: their maxima, the |E|, |B| maxima and the integrals to 1e-10; the
        # other components are rounding noise, bounded relative to the field they belong to
        for k_ in (0, 1, 3, 5, 6, 7, 10, 11):
            assert abs(hist[k_] - fh[k_]) <= 1e-10 * fh[k_], k_
        for k_, scale in ((2, fh[0]), (4, fh[0]), (8, fh[6]), (9, fh[6])):
            assert abs(hist[k_] - fh[k_]) <= 1e-10 * scale, k_
        o3 = np.zeros(3)
        ok.ok_compute_ke_maxwell(C.byref(sp[0].g), f_old[0].ravel(), deck.species[0].mass, vel_table, vz_old[0].ravel(), o3)
        assert np.all(np.abs(hist[12:15] - o3) <= 1e-10 * np.abs(o3)) and hist[15] == 0.0 and hist[16] == 0.0
    assert np.any(em_old[I2] != em[I2])
    H.lk_vm_destroy(sys_)
    ok.ok_vm_work_destroy(w)


def test_full_regression_run_em_damping(lk, ok, fast):
    _em_damping_full_run(ok, decks.em_damping(n=(32, 5), nv=(32, 32)))


def _em_damping_full_run(ok, deck):
    """the emDamping regression run in full length (final_time = 10, save_times = 0.2, cfl = 0.8; about 210 RK4
    steps, limited by Maxwell::computeDt) on the deck's own configuration-space grid with a 32 x 32 velocity
    grid, production arithmetic on the device against the oracle: the time-history traces of the wave (|E|,
    Ey, |B|, Bz maxima, field energies, species kinetic energy) within 1e-10 in the norm of the run and 1e-9
    sample by sample; the distribution within 1e-10 per cell on the bulk at the end"""
    from loki_b200 import run
    deck.run = dict(final_time=10.0, save_times=0.2, max_step=1000000)
    r = run.Runner(deck)
    w, sp, keep = _oracle(ok, deck)
    states, em, vz = _setup(deck, 0, 0.0)
    ng = deck.ng
    f_old, f_new = [states[0].copy()], [np.zeros_like(states[0])]
    em_old, em_new = em.copy(), np.zeros_like(em)
    vz_old, vz_new = [vz[0].copy()], [np.zeros_like(vz[0])]
    rhs0, rvz0 = [np.zeros_like(states[0])], [np.zeros_like(vz[0])]
    ax, ay = np.zeros(1), np.zeros(1)
    ok.ok_vm_eval_rhs(w, _ptrs(rhs0), np.zeros_like(em), _ptrs(rvz0), _ptrs(f_old), em_old, _ptrs(vz_old), 0.0, ax, ay)
    g_ = sp[0].g
    nd_ = g_.nd
    vt = np.zeros(nd_[2] * nd_[3] * 2)
    lo_ = (C.c_int * 2)(-ng, -ng)
    ok.ok_build_velocity_tables(C.byref(g_), C.byref(lo_), deck.species[0].vlim[0], deck.species[0].vlim[2], vt,
                                np.zeros((nd_[2] + 1) * nd_[3] * 2), np.zeros(nd_[2] * (nd_[3] + 1) * 2))
    t, last_save = 0.0, 0
    dev_tr, ora_tr = [], []
    while not r.done():
        dt_o = run.select_dt(t, deck.cfl * ok.ok_vm_stable_dt(w, ax, ay, 4), last_save, 0.2, 10.0)
        dt_d = r.advance()
        assert abs(dt_d - dt_o) <= 1e-10 * dt_o
        ok.ok_vm_rk4_step(w, _ptrs(f_new), _ptrs(f_old), em_new, em_old, _ptrs(vz_new), _ptrs(vz_old), t, dt_o)
        t += dt_o
        if t >= (last_save + 1) * 0.2:
            last_save += 1
        f_old, f_new = f_new, f_old
        em_old, em_new = em_new, em_old
        vz_old, vz_new = vz_new, vz_old
        ok.ok_vm_last_accel_max(w, ax, ay)
        hist = r.history()
        fh = np.zeros(12)
        ok.ok_field_history(np.ascontiguousarray(em_old).ravel(), deck.n[0], deck.n[1], ng, 6, np.array(deck.dx + (1.0, 1.0)), fh)
        o3 = np.zeros(3)
        ok.ok_compute_ke_maxwell(C.byref(g_), f_old[0].ravel(), deck.species[0].mass, vt, vz_old[0].ravel(), o3)
        idx = (0, 1, 3, 5, 6, 7, 10, 11)
        dev_tr.append([hist[k_] for k_ in idx] + list(hist[12:15]))
        ora_tr.append([fh[k_] for k_ in idx] + list(o3))
    assert r.step >= 150 and abs(r.time - 10.0) < 1e-9
    dev_tr, ora_tr = np.array(dev_tr), np.array(ora_tr)
    worst = np.max(np.abs(dev_tr - ora_tr), axis=0) / np.max(np.abs(ora_tr), axis=0)
    assert np.all(worst <= 1e-10), worst
    rel = np.abs(dev_tr - ora_tr) / np.abs(ora_tr)
    assert np.all(rel <= 1e-9), rel.max(axis=0)
    out = r.state(0)
    I = (slice(ng, -ng),) * 4
    big = f_old[0][I] >= 1e-6 * f_old[0][I].max()
    assert cell_rel_err(out[I][big], f_old[0][I][big]) <= 1e-10
    print("emDamping regression run: %d steps, worst trace difference %.2e" % (r.step, float(worst.max())))
    res = dict(steps=r.step, worst_trace=float(worst.max()), per_cell_max_all_cells=cell_rel_err(out[I], f_old[0][I]),
               per_cell_max_above_1e6_of_peak=cell_rel_err(out[I][big], f_old[0][I][big]),
               stencil_neighbourhood_max=star_rel_err(out, f_old[0], f_old[0], ng))
    r.close()
    ok.ok_vm_work_destroy(w)
    return res
