"""Independent checks of the path's arithmetic that need no stored answer ("consequence for the new repo"):
invariants of the scheme evaluated on the oracle -- the GPU path is held to the oracle at 1e-12 / bit for bit elsewhere.

  * mass: with x and y periodic the advection and acceleration terms are differences of face fluxes, so every species'
    particle number changes only by what leaves through the velocity boundaries (e^-24 of the peak at 7 thermal speeds);
  * symmetry: a deck whose initial condition and driver are even in y keeps a state that is even in y."""
import ctypes as C

import numpy as np

import decks
from test_gpu_vp_system import _oracle, _ptrs


def _mass(deck, sp, f):
    ng = deck.ng
    n, dx = deck.geom_of(sp)
    return float(f[ng:-ng, ng:-ng, ng:-ng, ng:-ng].sum() * np.prod(dx))


def _widen(deck, factor):
    """the same deck with its velocity domains and cell counts scaled by `factor` (same cell sizes)"""
    for sp in deck.species:
        sp.vlim = tuple(v * factor for v in sp.vlim)
        sp.nv = (int(round(sp.nv[0] * factor)), int(round(sp.nv[1] * factor)))
    return deck


def test_mass_is_conserved_to_rounding_over_a_step(ok):
    """... and what is lost is what leaves through the velocity boundaries: 7e-9 with velocity grids that end at 7
    thermal speeds (the decks' extent; unit cells here), 1e-11 at 9, rounding (1e-15) at 11"""
    cases = ((lambda: decks.plane_iaw(n=(12, 10), nv=(14, 14)), ok.ok_vp_rk4_step),
             (lambda: decks.plane_iaw(n=(10, 10), nv=(14, 14), order=6, rk=6), ok.ok_vp_rk6_step))
    for mk, step in cases:
        loss = {}
        for vmax in (7, 9, 11):
            deck = _widen(mk(), vmax / 7.0)
            w, sp, keep = _oracle(ok, deck)
            f_old = [deck.initial_state(s)[0] for s in deck.species]
            f_new = [np.zeros_like(f) for f in f_old]
            before = [_mass(deck, s, f) for s, f in zip(deck.species, f_old)]
            step(w, _ptrs(f_new), _ptrs(f_old), 2.0, 0.05, np.zeros(len(f_old)))      # the driver is on: E is not zero
            after = [_mass(deck, s, f) for s, f in zip(deck.species, f_new)]
            assert all(np.any(fn != fo) for fn, fo in zip(f_new, f_old))
            loss[vmax] = max(abs(a - b) / abs(b) for a, b in zip(after, before))
            ok.ok_vp_work_destroy(w)
        assert loss[7] <= 1e-7 and loss[9] <= 1e-10 and loss[11] <= 1e-14, loss
        assert loss[7] > 100 * loss[9] > 1e4 * loss[11]     # the loss is the boundary's, not the scheme's


def test_a_state_even_in_y_stays_even_in_y(ok):
    deck = decks.plane_epw(n=(16, 8), nv=(32, 16))
    w, sp, keep = _oracle(ok, deck)
    s0 = deck.species[0]
    f0 = deck.initial_state(s0)[0]
    ng = deck.ng
    I = (slice(ng, -ng),) * 4
    assert np.array_equal(f0[I], f0[I][::-1, :, ::-1, :])            # even under (y, vy) -> (-y, -vy)
    f_old, f_new = [f0.copy()], [np.zeros_like(f0)]
    t = 2.0
    for _ in range(2):
        ok.ok_vp_rk4_step(w, _ptrs(f_new), _ptrs(f_old), t, 0.05, np.zeros(1))
        f_old, f_new = f_new, f_old
        t += 0.05
    g = f_old[0][I]
    assert np.any(g != f0[I])
    scale = np.abs(g).max()
    assert np.max(np.abs(g - g[::-1, :, ::-1, :])) <= 1e-13 * scale
    ok.ok_vp_work_destroy(w)
