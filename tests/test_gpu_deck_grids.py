"""Parity at the regression decks' OWN grids (BASELINE.json configs[0], [1] physics, [3]):
planeEPW_fixedIons 32 x 32 x 128 x 32, planeIAW 2 x (32 x 32 x 64 x 32), emDamping 32 x 5 x 64 x 64, with the decks'
own initial conditions.  Against the oracle (pinned to the transliterated reference Fortran):

  * one RK4 step -- strict arithmetic: the distribution bit for bit; production arithmetic: BOTH metrics, the
    per-cell relative difference of checkTests.C:345-358 over every interior cell and the difference relative to
    the cell's stencil neighbourhood, each <= 1e-12 (the north-star tolerance for the distribution after one step);
  * the whole regression run (final_time = 5, the reference's time-step selection; the oracle's loop nests run on
    all host cores) -- every time-history trace within 1e-10 in the norm of the run (north-star tolerance for the
    traces), the distribution within 1e-10 per cell.  LOKI_SHORT_DECKS=1 stops at the first plot time instead
    (the fields are then still the response to a driver at 1e-3 of its amplitude and carry the summation-order
    noise of the charge density: the field traces are only held to 1e-8 there).

The figures are appended to gpurun_out/deck_parity.jsonl (BASELINE.md section 5 quotes them)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import decks
import test_gpu_vm_system as tvm
import test_gpu_vp_system as tvp
from util import cell_rel_err, star_rel_err

pytestmark = pytest.mark.gpu
FULL = os.environ.get("LOKI_SHORT_DECKS") != "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record(**kw):
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "deck_parity.jsonl"), "a") as fh:
            fh.write(json.dumps(kw) + "\n")
    except OSError:
        pass


VP_DECKS = {"planeEPW_fixedIons": lambda: decks.plane_epw(), "planeIAW": lambda: decks.plane_iaw()}


@pytest.mark.parametrize("mode", ["strict", "production"])
@pytest.mark.parametrize("name", sorted(VP_DECKS))
def test_vp_one_step_at_the_decks_own_grid(lk, ok, name, mode):
    deck = VP_DECKS[name]()
    assert deck.n == (32, 32) and [s.nv for s in deck.species][0] in ((128, 32), (64, 32))
    old = lk.lk_set_strict(1 if mode == "strict" else 0)
    try:
        w, sp, keep = tvp._oracle(ok, deck)
        states, tables = [], []
        for s in deck.species:
            f, fx, fv, fnorm = deck.initial_state(s)
            states.append(f)
            tables.append((fx, fv, fnorm))
        ns = len(states)
        t0, dt = 2.0, 0.05       # the driver is a few per cent up its ramp: E is not rounding noise
        f_old = [s.copy() for s in states]
        f_new = [np.zeros_like(s) for s in states]
        ke = np.zeros(ns)
        ok.ok_vp_rk4_step(w, tvp._ptrs(f_new), tvp._ptrs(f_old), t0, dt, ke)
        H, sys_ = tvp._product(deck, states, tables)
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
                per_cell = cell_rel_err(out[I], f_new[s][I])                       # checkTests.C:345-358, every cell
                star = star_rel_err(out, f_new[s], np.maximum(np.abs(states[s]), np.abs(f_new[s])), ng)
                _record(test="one_step", deck=name, species=deck.species[s].name, grid=list(deck.n) + list(deck.species[s].nv),
                        per_cell_max=per_cell, stencil_neighbourhood_max=star)
                assert star <= 1e-12 and per_cell <= 1e-12, (per_cell, star)
        H.lk_vp_destroy(sys_)
        ok.ok_vp_work_destroy(w)
    finally:
        lk.lk_set_strict(old)


@pytest.mark.parametrize("name", sorted(VP_DECKS))
def test_vp_regression_run_at_the_decks_own_grid(lk, ok, fast, name):
    deck = VP_DECKS[name]()
    final_time = 5.0 if FULL else 1.0
    steps, dev_tr, ora_tr, got, want = tvp._full_run_vs_oracle(ok, deck, final_time, 1.0)
    assert steps >= (40 if FULL else 8)
    worst = np.max(np.abs(dev_tr - ora_tr), axis=0) / np.max(np.abs(ora_tr), axis=0)
    ng = deck.ng
    I = (slice(ng, -ng),) * 4
    per_cell = [cell_rel_err(got[s][I], want[s][I]) for s in range(len(got))]     # every interior cell, reported
    # ... and over the cells that carry the distribution: after 80 steps the far Maxwellian tails (the ion grid
    # reaches 12 thermal speeds: f ~ 1e-35 of the peak) hold sums that cancel to rounding, where a per-cell
    # relative difference has no bound in either code; everything above 1e-12 of the peak is held to 1e-10
    bulk = [cell_rel_err(got[s][I][want[s][I] >= 1e-12 * want[s][I].max()], want[s][I][want[s][I] >= 1e-12 * want[s][I].max()])
            for s in range(len(got))]
    star = [star_rel_err(got[s], want[s], want[s], ng) for s in range(len(got))]
    _record(test="run", deck=name, final_time=final_time, steps=steps, worst_trace=float(worst.max()), per_cell_max_all_cells=per_cell,
            per_cell_max_above_1e12_of_peak=bulk, stencil_neighbourhood_max=star)
    assert np.all(worst <= (1e-10 if FULL else 1e-8)), worst
    assert max(star) <= 1e-10 and max(bulk) <= 1e-10 and max(per_cell) <= 1e-8, (per_cell, bulk, star)


@pytest.mark.parametrize("mode", ["strict", "production"])
def test_em_damping_one_step_at_the_decks_own_grid(lk, ok, mode):
    deck = decks.em_damping()
    assert deck.n == (32, 5) and deck.species[0].nv == (64, 64)
    old = lk.lk_set_strict(1 if mode == "strict" else 0)
    try:
        w, sp, keep = tvm._oracle(ok, deck)
        states, em, vz = tvm._setup(deck, 0, 0.0)          # the deck's own initial condition and fields
        t0, dt = 0.0, 0.01
        f_old, f_new = [states[0].copy()], [np.zeros_like(states[0])]
        em_old, em_new = em.copy(), np.zeros_like(em)
        vz_old, vz_new = [vz[0].copy()], [np.zeros_like(vz[0])]
        ok.ok_vm_rk4_step(w, tvm._ptrs(f_new), tvm._ptrs(f_old), em_new, em_old, tvm._ptrs(vz_new), tvm._ptrs(vz_old), t0, dt)
        H, sys_ = tvm._product(deck, states, em, vz)
        assert H.lk_vm_set_time(sys_, t0) == 0
        assert H.lk_vm_advance(sys_, dt) == 0, H.lk_last_error()
        ng = deck.ng
        I = (slice(ng, -ng),) * 4
        I2 = (slice(None), slice(ng, -ng), slice(ng, -ng))
        out = np.empty_like(states[0])
        assert H.lk_vm_get_state(sys_, 0, out.ctypes.data) == 0
        em_d = np.empty_like(em)
        assert H.lk_vm_get_fields(sys_, em_d.ctypes.data) == 0
        if mode == "strict":
            assert np.array_equal(out[I], f_new[0][I]) and np.array_equal(em_d[I2], em_new[I2])
        else:
            per_cell = cell_rel_err(out[I], f_new[0][I])
            star = star_rel_err(out, f_new[0], np.maximum(np.abs(states[0]), np.abs(f_new[0])), ng)
            # the deck's wave lives in Ey / Bz; the other components are rounding noise: E relative to |E|, B to |B|
            scale = [np.max(np.abs(em_new[0:3]))] * 3 + [np.max(np.abs(em_new[3:6]))] * 3
            field = max(float(np.max(np.abs(em_d[c][I2[1:]] - em_new[c][I2[1:]])) / scale[c]) for c in range(6))
            _record(test="one_step", deck="emDamping", species="electron", grid=[32, 5, 64, 64], per_cell_max=per_cell,
                    stencil_neighbourhood_max=star, field_max=field)
            assert star <= 1e-12 and per_cell <= 1e-12 and field <= 1e-12, (per_cell, star, field)
        H.lk_vm_destroy(sys_)
        ok.ok_vm_work_destroy(w)
    finally:
        lk.lk_set_strict(old)


def test_em_damping_regression_run_at_the_decks_own_grid(lk, ok, fast):
    """emDamping in full length (final_time = 10, about 300 RK4 steps) at 32 x 5 x 64 x 64"""
    res = tvm._em_damping_full_run(ok, decks.em_damping())
    _record(test="run", deck="emDamping", final_time=10.0, **res)
    assert res["worst_trace"] <= 1e-10 and res["stencil_neighbourhood_max"] <= 1e-10


@pytest.mark.parametrize("mode", ["strict", "production"])
def test_pitch_angle_collisions_one_step_at_the_decks_own_grid(lk, ok, mode):
    """test/pitchAngleCollisions at 32 x 5 x 64 x 64 with the deck's own initial condition and collision operator
    (conservative, order 4): one RK4 step.  Strict: the Vlasov part carries the reference's bits and the operator is
    the reference's scheme in operator form (equal to rounding): 1e-14 of the peak; production: both metrics <= 1e-12."""
    deck = decks.pitch_angle_collisions()
    assert deck.n == (32, 5) and deck.species[0].nv == (64, 64) and deck.species[0].collision["conservative"] == 1
    old = lk.lk_set_strict(1 if mode == "strict" else 0)
    try:
        w, sp, keep = tvp._oracle(ok, deck)
        f, fx, fv, fnorm = deck.initial_state(deck.species[0])
        t0, dt = 0.0, 0.003
        f_old, f_new = [f.copy()], [np.zeros_like(f)]
        ok.ok_vp_rk4_step(w, tvp._ptrs(f_new), tvp._ptrs(f_old), t0, dt, np.zeros(1))
        H, sys_ = tvp._product(deck, [f], [(fx, fv, fnorm)])
        assert H.lk_vp_set_time(sys_, t0) == 0
        assert H.lk_vp_advance(sys_, dt) == 0, H.lk_last_error()
        ng = deck.ng
        I = (slice(ng, -ng),) * 4
        out = np.empty_like(f)
        assert H.lk_vp_get_state(sys_, 0, out.ctypes.data) == 0
        assert np.any(out[I] != f[I])
        if mode == "strict":
            assert np.abs(out[I] - f_new[0][I]).max() <= 1e-14 * np.abs(f_new[0][I]).max()
        else:
            per_cell = cell_rel_err(out[I], f_new[0][I])
            star = star_rel_err(out, f_new[0], np.maximum(np.abs(f), np.abs(f_new[0])), ng)
            _record(test="one_step", deck="pitchAngleCollisions", species="electron", grid=[32, 5, 64, 64], per_cell_max=per_cell,
                    stencil_neighbourhood_max=star)
            assert star <= 1e-12 and per_cell <= 1e-12, (per_cell, star)
        H.lk_vp_destroy(sys_)
        ok.ok_vp_work_destroy(w)
    finally:
        lk.lk_set_strict(old)


def test_pitch_angle_collisions_run_at_the_decks_own_grid(lk, ok, fast):
    """the first steps of the deck through the runner (pp-style options -> lk_vp_set_pitch_angle, the collisional step
    limit in every stableDt) against the oracle: the same dt sequence, the kinetic-energy trace within 1e-10, the field
    traces within 1e-8 (the deck's perturbation is 1e-4 of the charge density, whose summation-order noise is 1e-14),
    the distribution within 1e-10.  The whole deck (final_time 5: 1 400 steps) is too long for the oracle."""
    deck = decks.pitch_angle_collisions()
    steps, dev_tr, ora_tr, got, want = tvp._full_run_vs_oracle(ok, deck, 0.03, 0.03)
    assert steps >= 8
    worst = np.max(np.abs(dev_tr - ora_tr), axis=0) / np.max(np.abs(ora_tr), axis=0)
    ng = deck.ng
    I = (slice(ng, -ng),) * 4
    per_cell = cell_rel_err(got[0][I], want[0][I])
    star = star_rel_err(got[0], want[0], want[0], ng)
    _record(test="run", deck="pitchAngleCollisions", final_time=0.03, steps=steps, worst_trace=float(worst.max()),
            ke_trace=float(worst[4]), per_cell_max_all_cells=per_cell, stencil_neighbourhood_max=star)
    assert np.all(worst[:4] <= 1e-8) and worst[4] <= 1e-10, worst
    assert star <= 1e-10 and per_cell <= 1e-8, (per_cell, star)
