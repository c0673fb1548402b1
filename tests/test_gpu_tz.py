"""The twilight-zone (manufactured-solution) forcing of the reference's TrigTZ regression deck (test/TrigTZ; TrigTZSource.C,
TZSourceF.f): f_exact = 1/(2 pi) exp(-v^2/2) (1 + amp cos x cos y sin t) solves the forced Vlasov-Poisson system, so a run
is checked against an analytic solution -- the one in-repo pin of the reference that does not need a baseline file
(SURVEY 8c).  Level 0 against the oracle (itself pinned bit for bit to the transliterated Fortran, test_oracle_pin.py),
then the deck at its own grid 16 x 16 x 64 x 64: one step against the oracle, and the error against f_exact along a run."""
import ctypes as C

import numpy as np
import pytest

import decks
import test_gpu_vp_system as tvp
from loki_b200 import capi, pp
from util import Setup, cell_rel_err, star_rel_err

pytestmark = pytest.mark.gpu

TRIG_TZ = """
# test/TrigTZ/TrigTZ.pp of the reference, restated (same numbers)
$pi = 3.1415926535897932384626;
$xa = -2*$pi;
$xb =  2*$pi;
domain_limits = $xa $xb $xa $xb
N = 16 16
periodic_dir = true true
cfl = 1.0
final_time = 5.
save_times = .1
sequence_write_times = .1
max_step = 1000000
number_of_species = 1
kinetic_species.1.name = "electron"
kinetic_species.1.velocity_limits = -7 7 -7 7
kinetic_species.1.Nv = 64 64
kinetic_species.1.mass = 1.0
kinetic_species.1.charge = -1.0
kinetic_species.1.tz.name = "TrigTZSource"
kinetic_species.1.tz.amp = 1
"""


def _deck(tmp_path):
    path = tmp_path / "TrigTZ.pp"
    path.write_text(TRIG_TZ)
    return decks._wrap(pp.load(str(path)))


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("order", [4, 6])
def test_trig_tz_kernels_equal_the_oracle_bits(lk, ok, order, kind):
    """lk_trig_tz_tables + lk_set_trig_tz_source / lk_compute_trig_tz_source_error on a box whose lower corner is not
    the domain's: bit for bit (the kernels take their sines, cosines and exponentials from host tables)"""
    import torch
    s = Setup(ok, (7, 6, 9, 8), order, bz=0.0)
    g = capi.Geom.make(s.n, order, s.dx)
    ng = s.ng
    lo = (C.c_int * 2)(3 - ng, -2 - ng)
    xlo = np.array([-2 * np.pi, -1.5])
    dx = np.array(s.dx)
    count = C.c_int64()
    assert lk.lk_trig_tz_table_count(C.byref(g), C.byref(count)) == 0
    n1d, n2d, n3d, n4d = s.nd
    assert count.value == 4 * n1d + 4 * n2d + n3d * n4d
    tab = torch.zeros(count.value, dtype=torch.float64, device="cuda")
    vel = torch.from_numpy(s.velocities).cuda()
    me, mi = 0.5, 25.0                                       # the two-species sources' masses (kinds 2, 3)

    def ok_set(arr, time, amp):
        if kind >= 2:
            ok.ok_set_two_species_trig_tz_source(arr, C.byref(s.g), lo, xlo, dx, time, s.velocities, np.array([amp, me, mi]), kind - 2)
        else:
            (ok.ok_set_electron_trig_tz_source if kind else ok.ok_set_trig_tz_source)(arr, C.byref(s.g), lo, xlo, dx, time, s.velocities, amp)

    def ok_err(arr, soln, time, amp):
        if kind >= 2:
            ok.ok_compute_two_species_trig_tz_source_error(arr, soln, C.byref(s.g), lo, xlo, dx, time, s.velocities,
                                                           np.array([amp, me, mi]), kind - 2)
        else:
            (ok.ok_compute_electron_trig_tz_source_error if kind else ok.ok_compute_trig_tz_source_error)(
                arr, soln, C.byref(s.g), lo, xlo, dx, time, s.velocities, amp)
    f = torch.from_numpy(s.f).cuda()
    for time, amp in ((0.0, 1.0), (0.37, 0.1), (2.5, 1.0)):
        par = (C.c_double * 3)(amp, me, mi)
        # the exponential table depends on the masses but not on the amplitude: built once per parameter set is enough
        assert lk.lk_trig_tz_tables(tab.data_ptr(), C.byref(g), C.byref(lo), C.byref((C.c_double * 2)(*xlo)), vel.data_ptr(), kind,
                                    C.byref(par), None) == 0
        base = np.random.default_rng(3).uniform(-1, 1, size=s.f.shape)
        want = base.copy()
        ok_set(want.ravel(), time, amp)
        d = torch.from_numpy(base).cuda()
        assert lk.lk_set_trig_tz_source(d.data_ptr(), C.byref(g), tab.data_ptr(), vel.data_ptr(), time, kind, C.byref(par), None) == 0
        got = d.cpu().numpy()
        assert np.array_equal(got, want) and not np.array_equal(got, base)
        e_want = np.zeros_like(base)
        ok_err(e_want.ravel(), s.f.ravel(), time, amp)
        e = torch.zeros_like(d)
        assert lk.lk_compute_trig_tz_source_error(e.data_ptr(), f.data_ptr(), C.byref(g), tab.data_ptr(), vel.data_ptr(), time, kind,
                                                  C.byref(par), None) == 0
        assert np.array_equal(e.cpu().numpy(), e_want)
    par = (C.c_double * 3)(1.0, me, mi)
    assert lk.lk_set_trig_tz_source(None, C.byref(g), tab.data_ptr(), vel.data_ptr(), 0.0, kind, C.byref(par), None) != 0
    assert lk.lk_set_trig_tz_source(tab.data_ptr(), C.byref(g), tab.data_ptr(), vel.data_ptr(), 0.0, 4, C.byref(par), None) != 0
    bad = (C.c_double * 3)(1.0, -1.0, -1.0)
    if kind >= 2:
        assert lk.lk_set_trig_tz_source(tab.data_ptr(), C.byref(g), tab.data_ptr(), vel.data_ptr(), 0.0, kind, C.byref(bad), None) != 0


@pytest.mark.parametrize("mode", ["strict", "production"])
def test_trig_tz_deck_one_step_at_its_own_grid(lk, ok, mode, tmp_path):
    deck = _deck(tmp_path)
    assert deck.n == (16, 16) and deck.species[0].nv == (64, 64) and deck.species[0].tz == dict(amp=1.0, kind=1)
    old = lk.lk_set_strict(1 if mode == "strict" else 0)
    try:
        w, sp, keep = tvp._oracle(ok, deck)
        s0 = deck.species[0]
        f, fx, fv, fnorm = deck.initial_state(s0)
        t0, dt = 0.4, 0.03          # sin t and cos t both matter
        f_old, f_new = [f.copy()], [np.zeros_like(f)]
        ok.ok_vp_rk4_step(w, tvp._ptrs(f_new), tvp._ptrs(f_old), t0, dt, np.zeros(1))
        H, sys_ = tvp._product(deck, [f], [(fx, fv, fnorm)])
        assert H.lk_vp_set_time(sys_, t0) == 0
        assert H.lk_vp_advance(sys_, dt) == 0, H.lk_last_error()
        out = np.empty_like(f)
        assert H.lk_vp_get_state(sys_, 0, out.ctypes.data) == 0
        ng = deck.ng
        I = (slice(ng, -ng),) * 4
        # the source moved the state: without it a spatially uniform Maxwellian would not change at all
        assert np.max(np.abs(out[I] - f[I])) > 1e-4
        if mode == "strict":
            assert np.array_equal(out[I], f_new[0][I])
        else:
            assert cell_rel_err(out[I], f_new[0][I]) <= 1e-12
            assert star_rel_err(out, f_new[0], np.maximum(np.abs(f), np.abs(f_new[0])), ng) <= 1e-12
        # the error array of a restart dump (TrigTZSource::computeError) against the oracle's
        err = np.empty_like(f)
        assert H.lk_vp_trig_tz_error(sys_, 0, t0 + dt, err.ctypes.data) == 0
        g = sp[0].g
        vt = np.zeros(g.nd[2] * g.nd[3] * 2)
        lo = (C.c_int * 2)(-ng, -ng)
        ok.ok_build_velocity_tables(C.byref(g), C.byref(lo), s0.vlim[0], s0.vlim[2], vt, np.zeros((g.nd[2] + 1) * g.nd[3] * 2),
                                    np.zeros(g.nd[2] * (g.nd[3] + 1) * 2))
        e_want = np.zeros_like(f)
        xlo = np.array([deck.xlim[0], deck.xlim[2]])
        ok.ok_compute_trig_tz_source_error(e_want.ravel(), out.ravel(), C.byref(g), lo, xlo, np.array(deck.geom_of(s0)[1]), t0 + dt, vt, 1.0)
        assert np.array_equal(err[I], e_want[I])
        H.lk_vp_destroy(sys_)
        ok.ok_vp_work_destroy(w)
    finally:
        lk.lk_set_strict(old)


def test_trig_tz_run_tracks_the_exact_solution(lk, fast, tmp_path):
    """the deck through the runner (Simulation::advance) to t = 0.2: the state stays within the scheme's truncation error
    of the manufactured solution (3.5e-5 at this grid, the oracle's figure; the unforced system would be off by 3e-2),
    and the error falls by the scheme's order when x and y are refined"""
    from loki_b200 import run
    errs = {}
    for n in (8, 16):
        path = tmp_path / ("TrigTZ_%d.pp" % n)
        path.write_text(TRIG_TZ.replace("N = 16 16", "N = %d %d" % (n, n)))
        deck = pp.load(str(path))
        deck.run["final_time"] = 0.2
        r = run.Runner(deck)
        while not r.done():
            r.advance()
        assert abs(r.time - 0.2) < 1e-12
        err = np.empty(r.shapes[0])
        capi.check(r.H.lk_vp_trig_tz_error(r.sys, 0, r.time, err.ctypes.data), "tz_error")
        ng = deck.ng
        errs[n] = float(np.abs(err[ng:-ng, ng:-ng, ng:-ng, ng:-ng]).max())
        r.close()
    assert 1e-6 < errs[16] < 5e-5, errs
    assert errs[8] / errs[16] > 6.0, errs          # fourth order in x, y (the velocity grid is not refined)


EPW_TZ = """
# test/EPWTZ/EPWTZ.pp of the reference, restated (same numbers): ElectronTrigTZSource, order 6 / RK6, 16^4 cells
$pi = 3.1415926535897932384626;
$xa = -2*$pi;
$xb =  2*$pi;
$ya = -1*$pi;
$yb =  1*$pi;
temporal_solution_order = 6
spatial_solution_order = 6
domain_limits = $xa $xb $ya $yb
N = 16 16
periodic_dir = true true
cfl = 1.0
final_time = 1
save_times = .2
sequence_write_times = .2
max_step = 1000000
number_of_species = 1
kinetic_species.1.name = "electron"
kinetic_species.1.velocity_limits = -7 7 -9 9
kinetic_species.1.Nv = 16 16
kinetic_species.1.mass = 1
kinetic_species.1.charge = -1.0
kinetic_species.1.tz.name = "ElectronTrigTZSource"
kinetic_species.1.tz.amp = 0.1
kinetic_species.1.ic.name = "Perturbed Maxwellian"
kinetic_species.1.ic.tx = 1.0
kinetic_species.1.ic.ty = 1.0
kinetic_species.1.ic.kx2 = 1.0
"""


@pytest.mark.parametrize("mode", ["strict", "production"])
def test_epw_tz_deck_one_step_at_its_own_grid(lk, ok, mode, tmp_path):
    """the reference's EPWTZ deck (ElectronTrigTZSource, kx = ky = 4, sixth order in space and time) against the oracle"""
    path = tmp_path / "EPWTZ.pp"
    path.write_text(EPW_TZ)
    deck = decks._wrap(pp.load(str(path)))
    assert deck.order == 6 and deck.rk == 6 and deck.species[0].tz == dict(amp=0.1, kind=2)
    old = lk.lk_set_strict(1 if mode == "strict" else 0)
    try:
        w, sp, keep = tvp._oracle(ok, deck)
        s0 = deck.species[0]
        f, fx, fv, fnorm = deck.initial_state(s0)
        t0, dt = 0.4, 0.03
        f_old, f_new = [f.copy()], [np.zeros_like(f)]
        ok.ok_vp_rk6_step(w, tvp._ptrs(f_new), tvp._ptrs(f_old), t0, dt, np.zeros(1))
        H, sys_ = tvp._product(deck, [f], [(fx, fv, fnorm)])
        assert H.lk_vp_set_time(sys_, t0) == 0
        assert H.lk_vp_advance(sys_, dt) == 0, H.lk_last_error()
        out = np.empty_like(f)
        assert H.lk_vp_get_state(sys_, 0, out.ctypes.data) == 0
        ng = deck.ng
        I = (slice(ng, -ng),) * 4
        assert np.max(np.abs(out[I] - f[I])) > 1e-5
        if mode == "strict":
            assert np.array_equal(out[I], f_new[0][I])
        else:
            assert star_rel_err(out, f_new[0], np.maximum(np.abs(f), np.abs(f_new[0])), ng) <= 1e-12
        H.lk_vp_destroy(sys_)
        ok.ok_vp_work_destroy(w)
    finally:
        lk.lk_set_strict(old)


IAW_TZ = """
# test/IAWTZ/IAWTZ.pp of the reference, restated (same numbers): electrons and ions of mass 10 with the two-species
# twilight-zone sources, order 6 / RK6, 16^4 cells per species
$pi = 3.1415926535897932384626;
$xa = -2*$pi;
$xb =  2*$pi;
$ya = -1*$pi;
$yb =  1*$pi;
$IMass = 10;
$EMass = 1 ;
$IAlpha = sqrt($IMass) ;
$EAlpha = sqrt($EMass) ;
$vthI = 1.0/$IAlpha ;
$vthE = 1.0/$EAlpha ;
$vExmax =  7*$vthE;
$vExmin = -7*$vthE;
$vEymax =  9*$vthE;
$vEymin = -9*$vthE;
$vIxmax =  7*$vthI;
$vIxmin = -7*$vthI;
$vIymax =  9*$vthI;
$vIymin = -9*$vthI;
temporal_solution_order = 6
spatial_solution_order = 6
domain_limits = $xa $xb $ya $yb
N = 16 16
periodic_dir = true true
cfl = 1.0
final_time = 1
save_times = .2
sequence_write_times = .2
max_step = 1000000
number_of_species = 2
kinetic_species.1.name = "electron"
kinetic_species.1.velocity_limits = $vExmin $vExmax $vEymin $vEymax
kinetic_species.1.Nv = 16 16
kinetic_species.1.mass = $EMass
kinetic_species.1.charge = -1.0
kinetic_species.1.tz.name = "TwoSpecies_ElectronTrigTZSource"
kinetic_species.1.tz.amp = 0.1
kinetic_species.1.tz.electron_mass = $EMass
kinetic_species.1.tz.ion_mass = $IMass
kinetic_species.1.ic.name = "Perturbed Maxwellian"
kinetic_species.1.ic.tx = 1.0
kinetic_species.1.ic.ty = 1.0
kinetic_species.2.name = "ion"
kinetic_species.2.velocity_limits = $vIxmin $vIxmax $vIymin $vIymax
kinetic_species.2.Nv = 16 16
kinetic_species.2.mass = $IMass
kinetic_species.2.charge = +1.0
kinetic_species.2.tz.name = "TwoSpecies_IonTrigTZSource"
kinetic_species.2.tz.amp = 0.1
kinetic_species.2.tz.electron_mass = $EMass
kinetic_species.2.tz.ion_mass = $IMass
kinetic_species.2.ic.name = "Perturbed Maxwellian"
kinetic_species.2.ic.tx = 1.0
kinetic_species.2.ic.ty = 1.0
"""


@pytest.mark.parametrize("mode", ["strict", "production"])
def test_iaw_tz_deck_one_step_at_its_own_grid(lk, ok, mode, tmp_path):
    """the reference's IAWTZ deck (two species coupled through the field, TwoSpecies_*TrigTZSource) against the oracle"""
    path = tmp_path / "IAWTZ.pp"
    path.write_text(IAW_TZ)
    deck = decks._wrap(pp.load(str(path)))
    assert deck.order == 6 and deck.rk == 6 and [sp.tz["kind"] for sp in deck.species] == [3, 4]
    assert deck.species[1].tz == dict(amp=0.1, kind=4, electron_mass=1.0, ion_mass=10.0)
    old = lk.lk_set_strict(1 if mode == "strict" else 0)
    try:
        w, sp, keep = tvp._oracle(ok, deck)
        states, tables = [], []
        for s_ in deck.species:
            f, fx, fv, fnorm = deck.initial_state(s_)
            states.append(f)
            tables.append((fx, fv, fnorm))
        t0, dt = 0.4, 0.02
        f_old = [s_.copy() for s_ in states]
        f_new = [np.zeros_like(s_) for s_ in states]
        ok.ok_vp_rk6_step(w, tvp._ptrs(f_new), tvp._ptrs(f_old), t0, dt, np.zeros(2))
        H, sys_ = tvp._product(deck, states, tables)
        assert H.lk_vp_set_time(sys_, t0) == 0
        assert H.lk_vp_advance(sys_, dt) == 0, H.lk_last_error()
        ng = deck.ng
        I = (slice(ng, -ng),) * 4
        for s_ in range(2):
            out = np.empty_like(states[s_])
            assert H.lk_vp_get_state(sys_, s_, out.ctypes.data) == 0
            assert np.max(np.abs(out[I] - states[s_][I])) > 1e-6
            if mode == "strict":
                assert np.array_equal(out[I], f_new[s_][I])
            else:
                assert star_rel_err(out, f_new[s_], np.maximum(np.abs(states[s_]), np.abs(f_new[s_])), ng) <= 1e-12
        H.lk_vp_destroy(sys_)
        ok.ok_vp_work_destroy(w)
    finally:
        lk.lk_set_strict(old)
