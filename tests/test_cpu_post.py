"""vp4DPostProcess / checkTests mirrors (loki_b200/post.py) on synthetic run output written through the same writer
classes a run uses: a 2 x 3 decomposition in (x, y) with uneven tiles is assembled back into the global array, field
files lose their ghosts and gain a slice axis, the last time-history file becomes the time series, and checkTests'
metric and tolerance files decide PASSED / FAILED."""
import os

import numpy as np
import pytest

from loki_b200 import h5lite, outputs, post


def _tiles(f, n, ng, nproc):
    """cut the global dataBox array f (n4d, n3d, n2d, n1d) the way ParallelArray does: rank = ix * py + iy"""
    info = outputs.distrib_info(0, nproc[0] * nproc[1] - 1, ng, n, [nproc[0], nproc[1], 1, 1])
    tiles = {}
    for rank in range(nproc[0] * nproc[1]):
        lo, hi = post._tile_box(np.array(info), rank)
        tiles[rank] = np.ascontiguousarray(f[:, :, lo[1]:hi[1] + 1 + 2 * ng, lo[0]:hi[0] + 1 + 2 * ng])
    return tiles, info


def _write_run(base, seed, n=(7, 8, 6, 5), ng=2, nproc=(2, 3), perturb=0.0):
    rng = np.random.default_rng(seed)
    shape = tuple(k + 2 * ng for k in reversed(n))
    x_lo, x_hi = [-1.0, -2.0, -7.0, -7.0], [1.0, 2.0, 7.0, 7.0]
    dx = [(x_hi[k] - x_lo[k]) / n[k] for k in range(4)]
    fs = {}
    for idx in range(2):
        items = []
        for name in ("electron", "ion"):
            f = rng.random(shape) * (1.0 + perturb)
            fs[(idx, name)] = f
            tiles, info = _tiles(f, list(n), ng, nproc)
            item = dict(sp=dict(name=name, mass=1.0, charge=-1.0), domain=(list(n), x_lo, x_hi, dx, (True, True)), tiles=tiles, info=info)
            if name == "electron":
                item["sp"]["driver_state"] = (0, 0.0, 0.0)
                item["integrated_e_dot_j"] = {r: 0.25 * (r + 1) for r in tiles}
            items.append(item)
        outputs.write_vp_restart(base, idx, items, ng, 0.5 * idx, 0.01, 0.9, 1.0, num_procs=nproc[0] * nproc[1], max_files=4)
    fw = outputs.FieldWriter(base, x_lo, dx, n, 4, 2)
    fields = []
    for k in range(3):
        ex, ey = rng.random((n[1] + 2 * ng, n[0] + 2 * ng)), rng.random((n[1] + 2 * ng, n[0] + 2 * ng))
        fields.append((ex, ey))
        fw.start_time_slice(0.25 * k, 0.01, ["EX", "EY"], 0, 1, ([0.5], [0.0]), n)
        for nm, a in (("EX", ex), ("EY", ey)):
            fw.write_field(nm, a, (-ng, -ng), (-ng, -ng), (n[0] + 2 * ng, n[1] + 2 * ng), ng)
        fw.end_time_slice()
    names = outputs.poisson_time_history_names(1, 0, ["electron", "ion"])
    seqs = [rng.random(6) * (1.0 + perturb) for _ in names]
    times = np.arange(6) * 0.1
    for k, upto in enumerate((1, 4, 6)):
        outputs.write_time_histories(base + ".time_hists", k, names, seqs, times, upto, 1, 0)
    return fs, fields, names, seqs, times


def test_post_process_assembles_a_decomposed_run(tmp_path):
    base = str(tmp_path / "deck")
    n, ng = (7, 8, 6, 5), 2
    fs, fields, names, seqs, times = _write_run(base, 1, n, ng)
    assert sorted(f for f in os.listdir(base) if f.startswith("dist_1")) == ["dist_1.hdf"] + ["dist_1.hdf.g%d" % k for k in range(4)]
    meta = post.post_process(base)
    assert meta["species"] == ["electron", "ion"] and meta["nxy"] == (7, 8) and meta["n_ghosts"] == 2 and not meta["is_maxwell"]
    for idx in range(2):
        m = h5lite.read("%s_dist_%d.hdf" % (base, idx))["root"]
        assert float(m["time"].data) == 0.5 * idx and int(m["species_list_size"].data) == 2
        assert m["electron"]["distribution"]["distribInfo"].data.tolist() == outputs.distrib_info(0, 0, ng, list(n), [1, 1, 1, 1])
        assert "Shaped Ramped Cosine Driver1_1" in m and "x_lo_krook" not in m["electron"]       # putToRestart_SkipKrook
        g = h5lite.read("%s_dist_%d.hdf.g0" % (base, idx))["root"]
        for name in ("electron", "ion"):
            assert np.array_equal(g[name + "\\distribution.p0"].data, fs[(idx, name)])
        assert float(g["electron\\integrated_e_dot_j.p0"].data[0]) == 0.25 * (1 + 2 + 3 + 4 + 5 + 6)
    ts = h5lite.read(base + "_timeSeries.hdf")["root"]
    assert np.array_equal(ts["series_time"].data, times) and "sequence_times" not in ts
    for k, nm in enumerate(names):
        assert np.array_equal(ts[nm].data, seqs[k])
    fl = h5lite.read(base + "_fields.hdf")["root"]
    assert fl["EX"].data.shape == (3, n[1], n[0]) and fl["time"].data.tolist() == [0.0, 0.25, 0.5]
    for k in range(3):
        assert np.array_equal(fl["EX"].data[k], fields[k][0][ng:-ng, ng:-ng])
        assert np.array_equal(fl["EY"].data[k], fields[k][1][ng:-ng, ng:-ng])
    assert fl["x"].data.shape == (n[0],)


def test_tile_boxes_follow_the_reference_split():
    info = np.array(outputs.distrib_info(0, 5, 2, [7, 8, 6, 5], [2, 3, 1, 1]))
    boxes = [post._tile_box(info, r) for r in range(6)]
    # x: 7 over 2 -> 4 + 3; y: 8 over 3 -> 3 + 3 + 2; rank = ix * 3 + iy (ParallelArray::getIndexRank)
    assert [b[0][:2] for b in boxes] == [[0, 0], [0, 3], [0, 6], [4, 0], [4, 3], [4, 6]]
    assert [b[1][:2] for b in boxes] == [[3, 2], [3, 5], [3, 7], [6, 2], [6, 5], [6, 7]]
    assert all(b[0][2:] == [0, 0] and b[1][2:] == [5, 4] for b in boxes)


def test_rel_diff_metric():
    """checkTests.C:345-358: |t - b| / |t|; |b| below where t == 0; 0 where both are 0"""
    assert post.rel_diff_max([1.0, 2.0], [1.0, 2.0]) == 0.0
    assert post.rel_diff_max([2.0], [1.0]) == 0.5
    assert post.rel_diff_max([0.0], [3.0]) == 1.0
    assert post.rel_diff_max([0.0, 0.0], [0.0, 0.0]) == 0.0
    assert post.rel_diff_max([-4.0, 1.0], [-5.0, 1.0]) == 0.25


def test_check_tests_passes_and_fails_like_the_reference(tmp_path):
    base_dir, test_dir = tmp_path / "BASELINES" / "deck", tmp_path / "run"
    os.makedirs(base_dir)
    os.makedirs(test_dir)
    b, t = str(base_dir / "deck"), str(test_dir / "deck")
    _write_run(b, 1)
    _write_run(t, 1, perturb=1e-9)
    post.post_process(b)
    post.post_process(t)
    (test_dir / "dist_tol").write_text("electron\n1.0e-8\nion\n1.0e-8\n")
    (test_dir / "tstol").write_text("E_max\n1.0e-8\nelectron_ke\n1.0e-8\n")
    (test_dir / "field_tol").write_text("EX\n1.0e-15\nEY\n1.0e-15\n")
    (test_dir / "input").write_text("dist_tol tstol field_tol deck deck 1 0\n")
    args = (t, b, 1, str(test_dir / "dist_tol"), str(test_dir / "tstol"), str(test_dir / "field_tol"))
    assert post.check_tests(*args) == []
    assert post.main(["check", str(test_dir / "input"), "--baselines", str(tmp_path / "BASELINES")]) == 0
    (test_dir / "dist_tol").write_text("electron\n1.0e-12\nion\n1.0e-8\n")
    (test_dir / "tstol").write_text("E_max\n1.0e-8\nelectron_ke\n1.0e-14\n")
    fails = post.check_tests(*args)
    assert len(fails) == 2 and "species electron" in fails[0] and "time series electron_ke" in fails[1]
    assert post.main(["check", str(test_dir / "input"), "--baselines", str(tmp_path / "BASELINES")]) == 1
    # the reference's own tolerance files parse (name / value line pairs)
    ref = "/root/reference/test/planeEPW_fixedIons/tstol"
    if os.path.exists(ref):
        tol = dict(post._tolerances(ref))
        assert tol["electron_ke"] == 1.0e-14 and tol["electron_driver_time_envel"] == 1.0e-10


def test_reader_classes(tmp_path):
    """FieldReader / TimeHistReader (FieldReader.C, TimeHistReader.C) over the raw and the post-processed files"""
    base = str(tmp_path / "deck")
    n, ng = (7, 8, 6, 5), 2
    fs, fields, names, seqs, times = _write_run(base, 3, n, ng)
    post.post_process(base)
    fr = outputs.FieldReader(base + ".fields_1.hdf")
    assert fr.read_num_time_slices_in_file() == 1 and fr.read_time("time_slice_2_time") == 0.5
    assert np.array_equal(fr.read_field("time_slice_2_EX"), fields[2][0].reshape(-1))
    assert outputs.FieldReader(base + ".fields_0.hdf").read_total_num_time_slices() == 3
    x, y = fr.read_coords()
    assert x.shape == (n[0],) and y.shape == (n[1],)
    th = outputs.TimeHistReader(base + ".time_hists_2.hdf")
    assert th.read_num_probes() == 1 and th.read_num_tracking_particles() == 0
    assert np.array_equal(th.read_time_history("electron_ke"), seqs[names.index("electron_ke")])
    assert np.array_equal(outputs.TimeHistReader(base + "_timeSeries.hdf").read_time_history("series_time"), times)
    assert np.array_equal(outputs.FieldReader(base + "_fields.hdf").read_field("EY").reshape(3, n[1], n[0])[1], fields[1][1][ng:-ng, ng:-ng])


@pytest.mark.parametrize("px,py", [(2, 3), (1, 4), (3, 1)])
def test_decomp_tiles_become_a_reference_dump(tmp_path, px, py):
    """a multi-GPU run's tiles (decomp.TileLayout: rank = ry * px + rx) written as the dump LOKI would write for the same
    decomposition (rank = ix * py + iy) and assembled back by the post processor: the split rule is the same one"""
    from loki_b200 import decomp
    n, ng = (7, 9, 4, 3), 2
    layout = decomp.TileLayout((n[0], n[1]), px, py)
    rng = np.random.default_rng(9)
    f = rng.random(tuple(k + 2 * ng for k in reversed(n)))
    tiles = {}
    for r in range(layout.world):
        lx, ly, nx, ny = layout.tile(r)
        tiles[r] = np.ascontiguousarray(f[:, :, ly:ly + ny + 2 * ng, lx:lx + nx + 2 * ng])
    ref_tiles, info = layout.restart_tiles(tiles, ng, (n[2], n[3]))
    assert sorted(ref_tiles) == list(range(layout.world))
    for r in range(layout.world):
        lo, hi = post._tile_box(np.array(info), layout.reference_rank(r))
        lx, ly, nx, ny = layout.tile(r)
        assert (lo[0], lo[1], hi[0] - lo[0] + 1, hi[1] - lo[1] + 1) == (lx, ly, nx, ny)
    x_lo, x_hi = [-1.0, -2.0, -7.0, -7.0], [1.0, 2.0, 7.0, 7.0]
    dx = [(x_hi[k] - x_lo[k]) / n[k] for k in range(4)]
    item = dict(sp=dict(name="electron", mass=1.0, charge=-1.0), domain=(list(n), x_lo, x_hi, dx, (True, True)), tiles=ref_tiles, info=info)
    base = str(tmp_path / "deck")
    outputs.write_vp_restart(base, 0, [item], ng, 0.0, 0.01, 0.9, 1.0, num_procs=layout.world, max_files=2)
    r = outputs.RestartReader(os.path.join(base, "dist_0.hdf"), 2)
    r.push_sub_dir("electron")
    assert np.array_equal(post.assemble_parallel_array(r, "distribution", list(n), ng), f)
