"""The pitch-angle collision operator (PitchAngleCollisionOperator.C, PitchAngleCollisionOperatorF.f) on
the device: the Level-0 kernels and the Fortran-ABI symbols against the oracle and the reference-derived golden
vectors, then the operator inside the stage loop of the host mirror (completeRHS, KineticSpecies.C:1036-1046) against
the oracle's RK steps, and the collisional time-step limit (KineticSpecies.C:666-672)."""
import ctypes as C
import os

import numpy as np
import pytest

import decks
import f77_cases
import ref_binding
from test_gpu_vp_system import _oracle, _perturb, _product, _ptrs
from test_oracle_pin import _pitch_case
from util import star_rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "f77abi_golden.npz")
CONS_TOL = 1e-13   # the conservative operators are the reference's generated code in operator form: equal to rounding


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("order", [4, 6])
def test_pitch_angle_kernels_match_oracle(lk, ok, order):
    """lk_pitch_angle_fields and lk_append_pitch_angle_collision through the C ABI: the reduced fields and the
    non-conservative operator carry the oracle's (= the reference Fortran's) bits, the conservative operator agrees to
    rounding; ghost cells of rhs are left alone"""
    import torch
    from loki_b200.capi import Geom, PitchAngle
    s, iv, xlo, xhi, rlo, rhi = _pitch_case(ok, order)
    g = Geom.make(s.n, order, s.dx)
    f_d, vel_d = _dev(torch, s.f), _dev(torch, s.velocities)
    iv_d = torch.zeros(iv.shape, dtype=torch.float64, device="cuda")
    pl = iv[0].size
    n0 = lk.lk_launch_count()
    assert lk.lk_pitch_angle_fields(iv_d.data_ptr(), iv_d.data_ptr() + 8 * pl, iv_d.data_ptr() + 16 * pl, f_d.data_ptr(), C.byref(g),
                                    vel_d.data_ptr(), None) == 0, lk.lk_last_error()
    torch.cuda.synchronize()
    assert lk.lk_launch_count() == n0 + 1
    assert np.array_equal(iv_d.cpu().numpy(), iv)
    vlo = (C.c_double * 2)(xlo[2], xlo[3])
    vhi = (C.c_double * 2)(xhi[2], xhi[3])
    rng = np.random.default_rng(11)
    ng = s.ng
    for cons in (0, 1):
        base = rng.uniform(-1, 1, size=s.f.shape) * 1e-3
        want = base.copy()
        ok.ok_append_pitch_angle_collision(want.ravel(), s.f.ravel(), C.byref(s.g), s.velocities, iv[0].ravel(), iv[1].ravel(),
                                           iv[2].ravel(), xlo[2:].copy(), xhi[2:].copy(), rlo, rhi, 0.05, 0.37, cons)
        r_d = _dev(torch, base)
        pa = PitchAngle.make(rlo, rhi, 0.05, 1.0, 0.37, cons)
        assert lk.lk_append_pitch_angle_collision(r_d.data_ptr(), f_d.data_ptr(), C.byref(g), vel_d.data_ptr(), iv_d.data_ptr(),
                                                  iv_d.data_ptr() + 8 * pl, iv_d.data_ptr() + 16 * pl, C.byref(vlo), C.byref(vhi),
                                                  C.byref(pa), None) == 0, lk.lk_last_error()
        torch.cuda.synchronize()
        got = r_d.cpu().numpy()
        ghosts = np.ones(s.f.shape, bool)
        ghosts[ng:-ng, ng:-ng, ng:-ng, ng:-ng] = False
        assert np.array_equal(got[ghosts], base[ghosts])
        if cons == 0:
            assert np.array_equal(got, want)
            assert np.any(got != base) == (order == 4)        # order 6 has no non-conservative form: nothing applied
        else:
            scale = np.abs(want - base).max()
            assert scale > 0.0 and np.abs(got - want).max() <= CONS_TOL * scale
    # arguments the reference aborts on (PitchAngleCollisionOperator.C:216-218, 253-269)
    bad = PitchAngle.make((-4.9, -0.9), rhi, 0.05, 1.0, 0.37, 1)
    assert lk.lk_pitch_angle_check(C.byref(g), C.byref(vlo), C.byref(vhi), C.byref(bad)) != 0
    assert b"too large" in lk.lk_last_error()
    good = PitchAngle.make(rlo, rhi, 0.05, 1.0, 0.37, 1)
    assert lk.lk_pitch_angle_check(C.byref(g), C.byref(vlo), C.byref(vhi), C.byref(good)) == 0
    lam = lk.lk_pitch_angle_real_lam(C.byref(g), C.byref(good))
    assert lam == ok.ok_pitch_angle_real_lam(C.byref(s.g), 0.37, 1.0, 0.05) and lam > 0.0


@pytest.mark.parametrize("order", [4, 6])
def test_fortran_abi_collision_routines_match_reference(lk, ok, order):
    """the five symbols of PitchAngleCollisionOperatorF.H replayed on device arrays through the argument lists that
    produced tests/golden/f77abi_golden.npz from the transliterated reference Fortran"""
    gold = np.load(GOLD)
    got = f77_cases.collision_cases(f77_cases.DeviceBackend(lk), ok, order)
    assert sorted(got) == ["coll_cons", "coll_noncons", "fields", "moments"]
    live = None
    if ref_binding.available():
        R = ref_binding.Ref()
        live = f77_cases.collision_cases(f77_cases.HostBackend(R.L, R.L.loki_ref_set_ic), ok, order)
    for name, val in got.items():
        for want in [gold["c%d_%s" % (order, name)]] + ([live[name]] if live else []):
            if name == "coll_cons":
                assert np.abs(val - want).max() <= CONS_TOL * np.abs(want).max(), name
            else:
                assert np.array_equal(val, want), "%s (order %d) differs from the reference Fortran's output" % (name, order)


COLL_DECKS = {
    # the reference's own deck shape (test/pitchAngleCollisions) on a grid the oracle steps in seconds; the spatial mode
    # and a 2 % roughness make the flow and thermal fields vary over (x, y)
    "cons4": lambda: decks.pitch_angle_collisions(n=(8, 5), nv=(32, 32), A=0.05),
    "noncons4": lambda: decks.pitch_angle_collisions(n=(8, 5), nv=(32, 32), A=0.05, conservative=0),
    "cons6_rk6": lambda: decks.pitch_angle_collisions(n=(8, 6), nv=(40, 40), order=6, rk=6, A=0.05),
}


@pytest.mark.parametrize("mode", ["strict", "production"])
@pytest.mark.parametrize("name", sorted(COLL_DECKS))
def test_collision_deck_steps_match_oracle(lk, ok, name, mode):
    """completeRHS with a collision operator inside the stage loop: two RK steps against the oracle.  Strict: the Vlasov
    part and the update carry the reference's bits and the operator the oracle's; production: 1e-12 of the stencil
    neighbourhood per step.  The step estimate includes the operator's real eigenvalue."""
    deck = COLL_DECKS[name]()
    old = lk.lk_set_strict(1 if mode == "strict" else 0)
    try:
        w, sp, keep = _oracle(ok, deck)
        s0 = deck.species[0]
        f, fx, fv, fnorm = deck.initial_state(s0)
        states = [_perturb(f, 91, amp=0.02)]
        H, sys_ = _product(deck, states, [(fx, fv, fnorm)])
        f_old, f_new = [states[0].copy()], [np.zeros_like(states[0])]
        ke = np.zeros(1)
        t, dt = 0.0, 0.01
        ng = deck.ng
        I = (slice(ng, -ng),) * 4
        for step in range(2):
            (ok.ok_vp_rk4_step if deck.rk == 4 else ok.ok_vp_rk6_step)(w, _ptrs(f_new), _ptrs(f_old), t, dt, ke)
            assert H.lk_vp_set_time(sys_, t) == 0
            assert H.lk_vp_advance(sys_, dt) == 0, H.lk_last_error()
            t += dt
            f_old, f_new = f_new, f_old
            out = np.empty_like(states[0])
            assert H.lk_vp_get_state(sys_, 0, out.ctypes.data) == 0
            if mode == "strict" and "noncons" in name:
                assert np.array_equal(out[I], f_old[0][I]), step
            elif mode == "strict":
                assert np.abs(out[I] - f_old[0][I]).max() <= 1e-14 * np.abs(f_old[0][I]).max(), step
            else:
                assert star_rel_err(out, f_old[0], np.maximum(np.abs(states[0]), np.abs(f_old[0])), ng) <= (step + 1) * 1e-12
        # stableDt after the step: the same accelerations on both sides, plus the collisional real eigenvalue
        ax, ay = np.zeros(1), np.zeros(1)
        lam = (C.c_double * 2)()
        assert H.lk_vp_lambda_max(sys_, 0, C.byref(lam)) == 0
        ax[0], ay[0] = lam[0], lam[1]
        dt_d = C.c_double()
        assert H.lk_vp_stable_dt(sys_, C.byref(dt_d)) == 0
        assert dt_d.value == ok.ok_vp_stable_dt(w, ax, ay, deck.rk)
        H.lk_vp_destroy(sys_)
        ok.ok_vp_work_destroy(w)
    finally:
        lk.lk_set_strict(old)


def test_collision_operator_changes_the_answer_and_can_be_removed(lk, ok, fast):
    """the operator is not a no-op, its step limit is tighter than the collisionless one, NULL removes it, and a range
    the reference aborts on is refused"""
    from loki_b200.capi import PitchAngle
    outs, dts = [], []
    for with_op in (True, False):
        deck = COLL_DECKS["cons4"]()
        states = [_perturb(deck.initial_state(deck.species[0])[0], 91, amp=0.02)]
        H, sys_ = _product(deck, states, None)
        if not with_op:
            assert H.lk_vp_set_pitch_angle(sys_, 0, None) == 0
        assert H.lk_vp_set_time(sys_, 0.0) == 0 and H.lk_vp_advance(sys_, 0.01) == 0
        out = np.empty_like(states[0])
        assert H.lk_vp_get_state(sys_, 0, out.ctypes.data) == 0
        outs.append(out)
        d = C.c_double()
        assert H.lk_vp_stable_dt(sys_, C.byref(d)) == 0
        dts.append(d.value)
        bad = PitchAngle.make((-6.9, -5.25), (5.25, 5.25), 0.01, 1.0, 0.1, 1)
        assert H.lk_vp_set_pitch_angle(sys_, 0, C.byref(bad)) != 0
        H.lk_vp_destroy(sys_)
    I = (slice(2, -2),) * 4
    assert not np.array_equal(outs[0][I], outs[1][I])
    assert dts[0] < dts[1]
