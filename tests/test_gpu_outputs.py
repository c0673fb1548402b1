"""A deck run that writes what Simulation writes: time-history, field and restart files from the
device state, and a run resumed from a restart dump continuing bit for bit (RestartManager::restore)."""
import ctypes as C
import os

import numpy as np
import pytest

from loki_b200 import capi, h5lite, outputs, pp, run
from test_gpu_vp_system import RUN_DECK
from util import star_rel_err

pytestmark = pytest.mark.gpu

DECK = RUN_DECK.replace("final_time = 0.3", "final_time = 0.4") + """
sequence_write_times = 0.05
restart.time_interval = 0.2
restart.write_directory = "two_species"
number_of_probes = 2
probe.1.location = 0.5 0.5
probe.2.location = 0.25 0.75
"""


def _runner(tmp_path, final_time=None):
    path = tmp_path / "two_species.pp"
    path.write_text(DECK)
    deck = pp.load(str(path))
    if final_time is not None:
        deck.run["final_time"] = final_time
    return run.Runner(deck), deck


def test_output_files_of_a_run(lk, fast, tmp_path):
    r, deck = _runner(tmp_path)
    base = str(tmp_path / "two_species")
    rs = deck.run["restart"]
    assert rs["time_interval"] == 0.2 and rs["write_directory"] == "two_species"
    r.open_outputs(base, rs.get("time_interval"))
    rows = [(0.0, r.sequence_record())]
    em = [r.em_vars()]
    times = [0.0]
    while not r.done():
        r.advance()
        if r.record:
            rows.append((r.time, r.sequence_record()))
        if r.save:
            em.append(r.em_vars())
            times.append(r.time)
    r.write_checkpoint_file()                                         # Simulation::finalize
    ns, npr = 2, 2
    names = outputs.poisson_time_history_names(npr, 0, ["electron", "ion"])
    # time histories: file <base>.time_hists_<k>.hdf is written at save k and holds every sequence so far
    nsave = len(times)
    assert nsave == 5 and abs(times[-1] - 0.4) < 1e-12
    last = h5lite.read("%s.time_hists_%d.hdf" % (base, nsave - 1))["root"]
    assert int(last["numProbes"].data[0]) == npr
    assert np.array_equal(last["sequence_times"].data, [t for t, _ in rows])
    # steps end on the save times (selectTimeStep), so the 0.05 cadence records once per step: t = 0, 0.1, ... 0.4
    assert len(rows) == 5
    for k, n in enumerate(names):
        assert np.array_equal(last[n].data, [row[k] for _, row in rows]), n
    first = h5lite.read(base + ".time_hists_0.hdf")["root"]
    assert first["E_max"].data.shape == (1,)
    # the histories carry physics: field energy grows from the driver, the electrons' driver work is recorded
    assert last["field_energy"].data[-1] > 0 and last["electron_driver_time_envel"].data[-1] > 0
    assert last["electron_integrated_ke_e_dot"].data[0] == 0.0 and last["electron_integrated_ke_e_dot"].data[-1] != 0.0
    assert np.all(last["ion_ke_e_dot"].data == 0.0)
    # probes: E at the probes' cells of the written field of the same time
    f4 = h5lite.read("%s.fields_%d.hdf" % (base, nsave - 1))["root"]
    ng = deck.ng
    ex = f4["time_slice_%d_EX" % (nsave - 1)].data
    for ip, (fx, fy) in enumerate(deck.probes):
        i, j = int(np.floor(fx * deck.n[0])), int(np.floor(fy * deck.n[1]))
        assert int(f4["ix_probe%d" % ip].data[0]) == i and int(f4["iy_probe%d" % ip].data[0]) == j
        assert last["Ex_probe%d" % ip].data[-1] == ex[j + ng, i + ng]
    # fields: one file per save (plot_times_per_file = 1), dataset = the configuration-space array with its ghosts
    f0 = h5lite.read(base + ".fields_0.hdf")["root"]
    assert int(f0["total_num_time_slices"].data[0]) == nsave
    for k in range(nsave):
        root = h5lite.read("%s.fields_%d.hdf" % (base, k))["root"]
        assert root["time_slice_%d_EX" % k].data.shape == (deck.n[1] + deck.order, deck.n[0] + deck.order)
        assert np.array_equal(root["time_slice_%d_EX" % k].data, em[k][0])
        assert np.array_equal(root["time_slice_%d_EY" % k].data, em[k][1])
        assert float(root["time_slice_%d_time" % k].data[0]) == times[k]
        assert np.array_equal(root["x"].data, deck.xlim[0] + (np.arange(deck.n[0]) + 0.5) * deck.dx[0])
    # restart dumps at t = 0, 0.2, 0.4 (time_interval 0.2)
    for idx, t in enumerate((0.0, 0.2, 0.4)):
        dump = outputs.read_vp_restart(os.path.join(base, "dist_%d.hdf" % idx))
        assert abs(dump["time"] - t) < 1e-12 and [s["name"] for s in dump["species"]] == ["electron", "ion"]
    assert not os.path.exists(os.path.join(base, "dist_3.hdf"))
    dump = outputs.read_vp_restart(os.path.join(base, "dist_2.hdf"))
    for s in range(ns):
        # the dump was taken after updateGhosts: interior bits are the state's, x / y ghosts the periodic images
        f = dump["species"][s]["distribution"]
        st = r.state(s)
        assert np.array_equal(f[ng:-ng, ng:-ng, ng:-ng, ng:-ng], st[ng:-ng, ng:-ng, ng:-ng, ng:-ng])
        assert np.array_equal(f[ng:-ng, ng:-ng, ng:-ng, :ng], f[ng:-ng, ng:-ng, ng:-ng, -2 * ng:-ng])
        assert np.array_equal(f[ng:-ng, ng:-ng, -ng:, ng:-ng], f[ng:-ng, ng:-ng, ng:2 * ng, ng:-ng])
    v = C.c_double()
    capi.check(r.H.lk_vp_ke_e_dot(r.sys, 0, C.byref(v)), "ke_e_dot")
    assert dump["species"][0]["integrated_e_dot_j"] == v.value and dump["species"][1]["integrated_e_dot_j"] is None
    r.close()


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_resumed_run_continues_bit_for_bit(lk, tmp_path, mode):
    """run to 0.4 in one go; run to 0.2, dump, resume in a new system from the dump, run to 0.4: same bits in strict
    arithmetic, 1e-13 of the stencil neighbourhood in production arithmetic"""
    old = lk.lk_set_strict(1 if mode == "strict" else 0)
    try:
        whole, _ = _runner(tmp_path)
        while not whole.done():
            whole.advance()
        want = [whole.state(s) for s in range(2)]
        v_want = C.c_double()
        capi.check(whole.H.lk_vp_ke_e_dot(whole.sys, 0, C.byref(v_want)), "ke_e_dot")
        t_want, steps_want = whole.time, whole.step
        whole.close()

        base = str(tmp_path / "two_species")
        first, deck = _runner(tmp_path, final_time=0.2)
        first.open_outputs(base, 0.2)
        while not first.done():
            first.advance()
        first.write_checkpoint_file()
        steps_first = first.step
        first.close()
        assert os.path.exists(os.path.join(base, "dist_1.hdf")) and not os.path.exists(os.path.join(base, "dist_2.hdf"))

        second, _ = _runner(tmp_path)
        assert second.restore(base) == 1
        assert abs(second.time - 0.2) < 1e-12 and second.last_save == 2
        while not second.done():
            second.advance()
        assert second.time == t_want and steps_first + second.step == steps_want
        v = C.c_double()
        capi.check(second.H.lk_vp_ke_e_dot(second.sys, 0, C.byref(v)), "ke_e_dot")
        for s in range(2):
            if mode == "strict":
                assert np.array_equal(second.state(s), want[s])
            else:
                # production arithmetic: the first charge density after a restore is a plain reduction, not the fused
                # stage kernel's partial sums, so the summation order (and the last bits) differ
                assert star_rel_err(second.state(s), want[s], want[s], deck.ng) <= 1e-13
        if mode == "strict":
            assert v.value == v_want.value
        else:
            assert abs(v.value - v_want.value) <= 1e-12 * abs(v_want.value)
        second.close()
    finally:
        lk.lk_set_strict(old)


def test_vlasov_maxwell_dump_and_resume(lk, strict, tmp_path):
    """emDamping physics on a small grid: field files carry the six fields and the species' vz, the dump carries group
    "Maxwell" (EMVars, vz0) with isMaxwell = 1 (Maxwell.C:942-964), and a system restored from it continues bit for bit"""
    from loki_b200 import decks as pdecks

    def mk(final_time):
        deck = pdecks.em_damping(n=(16, 5), nv=(16, 12))
        deck.run = dict(final_time=final_time, save_times=0.05, sequence_write_times=0.05, max_step=100, restart={})
        return run.Runner(deck), deck

    whole, _ = mk(0.1)
    while not whole.done():
        whole.advance()
    want_f, want_em, want_vz, t_want = whole.state(0), whole.em_vars(), whole.vz(0), whole.time
    whole.close()

    base = str(tmp_path / "emDamping")
    first, deck = mk(0.05)
    first.open_outputs(base, 0.05)
    while not first.done():
        first.advance()
    first.write_checkpoint_file()
    em_mid = first.em_vars()
    first.close()
    f1 = h5lite.read(base + ".fields_1.hdf")["root"]
    assert sorted(n for n in f1.names() if n.startswith("time_slice_1_")) == sorted(
        "time_slice_1_" + n for n in ["EX", "EY", "EZ", "BX", "BY", "BZ", "electron VZ", "time", "dt"])
    assert np.array_equal(f1["time_slice_1_EY"].data, em_mid[1]) and np.array_equal(f1["time_slice_1_BZ"].data, em_mid[5])
    hist = h5lite.read(base + ".time_hists_1.hdf")["root"]
    assert hist["E_max"].data.shape == (2,) and hist["electron_ke"].data[1] > 0 and hist["Bz_max"].data[0] > 0
    meta = h5lite.read(os.path.join(base, "dist_1.hdf"))["root"]
    assert int(meta["isMaxwell"].data) == 1 and meta["Maxwell"]["EMVars"]["distribInfo"].data.tolist()[:5] == [0, 0, -2, -2, -2]
    bulk = h5lite.read(os.path.join(base, "dist_1.hdf.g0"))["root"]
    assert bulk["Maxwell\\EMVars.p0"].data.shape == (6, 5 + 4, 16 + 4) and bulk["Maxwell\\vz0.p0"].data.shape == (9, 20)

    second, _ = mk(0.1)
    assert second.restore(base) == 1 and abs(second.time - 0.05) < 1e-12
    while not second.done():
        second.advance()
    assert second.time == t_want
    assert np.array_equal(second.state(0), want_f)
    assert np.array_equal(second.em_vars(), want_em) and np.array_equal(second.vz(0), want_vz)
    second.close()


def test_regression_chain_run_post_process_check(lk, tmp_path):
    """the reference's regression chain (test/*/Makefile.in): run -> vp4DPostProcess -> checkTests, here with the strict
    run as the baseline of the production run, under tolerance files in the reference's format"""
    from loki_b200 import post
    prefixes = {}
    for mode in ("strict", "fast"):
        old = lk.lk_set_strict(1 if mode == "strict" else 0)
        try:
            d = tmp_path / mode
            d.mkdir()
            r, deck = _runner(d)
            base = str(d / "two_species")
            r.open_outputs(base, 0.2)
            while not r.done():
                r.advance()
            r.write_checkpoint_file()
            r.close()
            meta = post.post_process(base)
            assert meta["species"] == ["electron", "ion"] and meta["nxy"] == tuple(deck.n)
            prefixes[mode] = base
        finally:
            lk.lk_set_strict(old)
    ts = h5lite.read(prefixes["fast"] + "_timeSeries.hdf")["root"]
    assert ts["series_time"].data.shape == (5,) and ts["electron_ke"].data.shape == (5,)
    fl = h5lite.read(prefixes["fast"] + "_fields.hdf")["root"]
    assert fl["EX"].data.shape == (5, 6, 12)
    # dist_tol lists the species to compare: the electrons.  (The cold ions of this deck reach 12 thermal speeds, f ~ 1e-35
    # of the peak there, where checkTests' per-cell relative difference is rounding noise over rounding noise in any two
    # runs that do not add in the same order; test_gpu_deck_grids.py treats the tails the same way.)
    (tmp_path / "dist_tol").write_text("electron\n1.0e-9\n")
    (tmp_path / "tstol").write_text("E_max\n1.0e-10\nfield_energy\n1.0e-10\nelectron_ke\n1.0e-10\nion_ke\n1.0e-10\n"
                                    "electron_integrated_ke_e_dot\n1.0e-10\nelectron_driver_time_envel\n1.0e-10\n")
    (tmp_path / "field_tol").write_text("EX\n1.0e-8\n")
    args = (prefixes["fast"], prefixes["strict"], 2, str(tmp_path / "dist_tol"), str(tmp_path / "tstol"), str(tmp_path / "field_tol"))
    assert post.check_tests(*args) == []
    (tmp_path / "dist_tol").write_text("electron\n1.0e-20\n")
    fails = post.check_tests(*args)
    assert len(fails) == 1 and "species electron" in fails[0]
