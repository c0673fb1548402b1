"""On-disk formats: the hand-laid HDF5 bytes (loki_b200/h5lite.py) and the writer classes that mirror
FieldWriter / TimeHistWriter / RestartWriter (loki_b200/outputs.py).

libhdf5 is absent from the image, so the pin is indirect: the reader is checked against a file libhdf5 itself wrote
(tests/golden/libhdf5_written_testhdf5_7.4_GLNX86.mat, scipy's MATLAB v7.3 fixture: a 512-byte user block, then a
version-0 superblock, old-style group, version-1 object headers), the writer's message bytes are compared with that
file's where the two describe the same thing, and everything the writer emits is read back by the pinned reader."""
import os
import struct

import numpy as np
import pytest

from loki_b200 import h5lite, outputs

HERE = os.path.dirname(os.path.abspath(__file__))
GENUINE = os.path.join(HERE, "golden", "libhdf5_written_testhdf5_7.4_GLNX86.mat")


def test_reader_reads_a_file_libhdf5_wrote():
    root = h5lite.read(GENUINE)
    assert root.names() == ["testdouble"]
    d = root["testdouble"]
    assert d.data.dtype == np.dtype("<f8") and d.data.shape == (9, 1)
    assert np.array_equal(d.data[:, 0], np.arange(9) * (np.pi / 4))        # MATLAB: testdouble = 0:pi/4:2*pi
    assert bytes(d.attrs["MATLAB_class"].tobytes()) == b"double"


def test_writer_messages_equal_libhdf5s_bytes():
    """the double datatype, the dataspace and the attribute of the genuine file, re-encoded by the writer"""
    buf = np.memmap(GENUINE, dtype="u1", mode="r")
    r = h5lite._Reader(memoryview(buf))
    (name, entry), = r.group_entries(r.root_entry["btree"], r.root_entry["heap"])
    msgs = {t: bytes(d) for t, _, d in r.messages(entry["oh"])}
    assert h5lite._pad8(h5lite._dtype_message("<f8")) == msgs[0x0003]
    assert h5lite._dspace_message((9, 1)) == msgs[0x0001]
    # superblock: same constants where the writer follows libhdf5's defaults
    assert (r.leaf_k, r.internal_k) == (h5lite.LEAF_K, h5lite.INTERNAL_K)


def _tree():
    root = h5lite.Group()
    g = outputs.ReaderWriterBase.create_group("root", root)
    rng = np.random.default_rng(5)
    for i in range(300):                                    # more than one symbol-table node, more than 2 * 16 of them
        g.put("time_slice_%d_EX" % i, rng.random((3, 5)))
    g.put("scalar int", np.array(-7, dtype=np.int32), scalar=True)
    g.put("scalar double", np.array(0.1), scalar=True)
    g.put("ivec", np.arange(-3, 3, dtype=np.int32))
    g.put("one", np.array([4], dtype=np.int32))
    g.put("name", np.frombuffer(b"electron\0", dtype="u1"))
    sub = outputs.ReaderWriterBase.create_group("species_list", g)
    sub.put("species.1", np.frombuffer(b"ion\0", dtype="u1"))
    g.put("root\\distribution.p0", rng.random((4, 6, 8, 10)))
    g.put("empty", np.zeros((0,)))
    return root


def test_write_read_round_trip(tmp_path):
    root = _tree()
    p = str(tmp_path / "t.hdf")
    h5lite.write(p, root)
    back = h5lite.read(p)

    def same(a, b):
        assert sorted(a.names()) == sorted(b.names())
        assert {k: v.tobytes() for k, v in a.attrs.items()} == {k: np.asarray(v).tobytes() for k, v in b.attrs.items()}
        for k, v in a.children.items():
            if isinstance(v, h5lite.Group):
                same(v, b[k])
            else:
                w = b[k].data
                assert w.dtype == v.data.dtype and w.shape == v.data.shape, k
                assert np.array_equal(w, v.data), k
    same(back, root)
    g = back["root"]
    assert g["scalar int"].data.shape == () and int(g["scalar int"].data) == -7
    assert g["one"].data.shape == (1,)
    assert g["ivec"].data.dtype == np.dtype(">i4")           # H5T_STD_I32BE
    assert bytes(g.attrs["className"]) == b"directory\0"     # ReaderWriterBase.C:118-160


def test_file_structure(tmp_path):
    """superblock fields, 8-byte alignment of every block, sorted symbol tables, B-tree keys"""
    p = str(tmp_path / "t.hdf")
    h5lite.write(p, _tree())
    b = open(p, "rb").read()
    assert b[:8] == h5lite.SIGNATURE
    assert b[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])
    leaf_k, internal_k = struct.unpack_from("<HH", b, 16)
    base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert (base, free, drv) == (0, h5lite.UNDEF, h5lite.UNDEF) and eof == len(b) and eof % 8 == 0
    assert leaf_k == 4 and internal_k >= 16
    r = h5lite._Reader(memoryview(b))
    root_group = r.load(r.root_entry)["root"]
    assert len(root_group.names()) == 308
    # walk the "root" group's B-tree by hand
    (name, e), = r.group_entries(r.root_entry["btree"], r.root_entry["heap"])
    assert name == "root" and e["cache"] == 1
    t = e["btree"]
    assert b[t:t + 4] == b"TREE" and t % 8 == 0
    ntype, level, used = struct.unpack_from("<BBH", b, t + 4)
    assert (ntype, level) == (0, 0) and used == -(-308 // 8) and used <= 2 * internal_k
    names = [n for n, _ in r.group_entries(e["btree"], e["heap"])]
    assert names == sorted(names, key=lambda s: s.encode())
    prev_key = struct.unpack_from("<Q", b, t + 24)[0]
    assert prev_key == 0
    for i in range(used):
        child, key = struct.unpack_from("<QQ", b, t + 24 + 8 + 16 * i)
        assert b[child:child + 4] == b"SNOD" and child % 8 == 0
        n = struct.unpack_from("<H", b, child + 6)[0]
        last = r.sym_entry(child + 8 + 40 * (n - 1))
        assert key == last["name_off"]                       # a key is the heap offset of the child's largest name


def test_existing_name_is_an_error():
    g = h5lite.Group()
    g.put("a", np.zeros(2))
    with pytest.raises(KeyError):
        g.put("a", np.zeros(2))


def test_time_hist_writer(tmp_path):
    """TimeHistWriter.C / EMSolverBase::writeTimeHistories: <base>_<n>.hdf, the first saved_seq entries of each sequence"""
    names = outputs.poisson_time_history_names(2, 0, ["electron", "ion"])
    assert len(names) == 5 + 2 * 2 + 16 * 2                 # Poisson::GLOBAL_TIME_HISTS, TIME_HISTS_PER_PROBE, per species
    assert names[:5] == ["E_max", "norm E", "Ex_max", "Ey_max", "field_energy"]
    assert names[5:9] == ["Ex_probe0", "Ey_probe0", "Ex_probe1", "Ey_probe1"]
    assert names[9] == "electron_ke" and names[9 + 13] == "electron_ke_e_dot" and names[-1] == "ion_driver_time_envel"
    seqs = [np.arange(8.0) + k for k in range(len(names))]
    times = np.arange(8.0) * 0.1
    f = outputs.write_time_histories(str(tmp_path / "run.time_hists"), 3, names, seqs, times, 5, 2, 0)
    assert f.endswith("run.time_hists_3.hdf")
    root = h5lite.read(f)["root"]
    assert np.array_equal(root["sequence_times"].data, times[:5])
    assert int(root["numProbes"].data[0]) == 2 and int(root["numTrackingParticles"].data[0]) == 0
    for k, n in enumerate(names):
        assert np.array_equal(root[n].data, seqs[k][:5])
    assert sorted(root.names()) == sorted(names + ["sequence_times", "numProbes", "numTrackingParticles"])


def test_field_writer_series_and_tiles(tmp_path):
    """FieldWriter.C: datasets of (Ny + order, Nx + order), ranks writing their localBox at lower + n_ghosts, a new
    file every time_slices_per_file slices, the running total in the first file"""
    nx, ny, ng, order = 12, 10, 2, 4
    base = str(tmp_path / "run")
    fw = outputs.FieldWriter(base, (-1.0, 2.0), (0.5, 0.25), (nx, ny), order, time_slices_per_file=2)
    rng = np.random.default_rng(1)
    fields = []
    for k in range(3):
        ex = rng.random((ny + 2 * ng, nx + 2 * ng))
        fields.append(ex)
        fw.start_time_slice(0.5 * k, 0.01, ["EX", "EY"], 0, 1, ([0.5], [0.0]), (nx, ny))
        # "EX" as one rank would write it: dataBox = localBox = the whole array with ghosts
        fw.write_field("EX", ex, (-ng, -ng), (-ng, -ng), (nx + 2 * ng, ny + 2 * ng), ng)
        # "EY" as a 2 x 1 decomposition would (ParallelArray::setupLocalDomain: a localBox keeps the domain-boundary ghosts)
        half = nx // 2
        left = ex[:, :half + 2 * ng]                         # dataBox [-ng, half + ng)
        right = ex[:, half:]                                 # dataBox [half - ng, nx + ng)
        fw.write_field("EY", left, (-ng, -ng), (-ng, -ng), (half + ng, ny + 2 * ng), ng)
        fw.write_field("EY", right, (half - ng, -ng), (half, -ng), (nx - half + ng, ny + 2 * ng), ng)
        fw.end_time_slice()
    f0 = h5lite.read(base + ".fields_0.hdf")["root"]
    f1 = h5lite.read(base + ".fields_1.hdf")["root"]
    assert int(f0["total_num_time_slices"].data[0]) == 3
    assert int(f0["num_time_slices_in_this_file"].data[0]) == 2 and int(f1["num_time_slices_in_this_file"].data[0]) == 1
    assert "total_num_time_slices" not in f1
    assert np.array_equal(f0["x"].data, -1.0 + (np.arange(nx) + 0.5) * 0.5)
    assert np.array_equal(f1["y"].data, 2.0 + (np.arange(ny) + 0.5) * 0.25)
    assert int(f0["numProbes"].data[0]) == 1 and int(f0["ix_probe0"].data[0]) == 6 and int(f0["iy_probe0"].data[0]) == 0
    for k, root in ((0, f0), (1, f0), (2, f1)):
        assert root["time_slice_%d_EX" % k].data.shape == (ny + order, nx + order)
        assert np.array_equal(root["time_slice_%d_EX" % k].data, fields[k])
        assert np.array_equal(root["time_slice_%d_EY" % k].data, fields[k])
        assert float(root["time_slice_%d_time" % k].data[0]) == 0.5 * k
        assert float(root["time_slice_%d_dt" % k].data[0]) == 0.01
    with pytest.raises(RuntimeError):
        fw.end_time_slice()


def test_distrib_info():
    """RestartWriter::writeParallelArray's integers against ParallelArray::setupLocalDomain (ParallelArray.C:641-653)"""
    assert outputs.distrib_info(0, 0, 2, [32, 32, 128, 32], [1, 1, 1, 1]) == [0, 0, -2, -2, -2, -2, 1, 1, 1, 1,
                                                                               32, 32, 128, 32, 1, 1, 1, 1]
    # 10 cells over 4 partitions: 3, 3, 2, 2 -> left partitions of 3 cells, two of them
    info = outputs.distrib_info(1, 8, 3, [10, 8, 6, 6], [4, 2, 1, 1])
    assert info[:2] == [1, 8] and info[2:6] == [-3] * 4 and info[6:10] == [4, 2, 1, 1]
    assert info[10:14] == [3, 4, 6, 6] and info[14:18] == [2, 2, 1, 1]


def _restart_items(rng, ng=2):
    n = [6, 4, 8, 6]
    x_lo, x_hi = [-1.0, -2.0, -7.0, -7.0], [1.0, 2.0, 7.0, 7.0]
    dx = [(x_hi[k] - x_lo[k]) / n[k] for k in range(4)]
    items = []
    for name, mass, charge, driven in (("electron", 1.0, -1.0, True), ("ion", 100.0, 1.0, False)):
        f = rng.random((n[3] + 2 * ng, n[2] + 2 * ng, n[1] + 2 * ng, n[0] + 2 * ng))
        sp = dict(name=name, mass=mass, charge=charge)
        item = dict(sp=sp, domain=(n, x_lo, x_hi, dx, (True, True)), tiles={0: f},
                    info=outputs.distrib_info(0, 0, ng, n, [1, 1, 1, 1]))
        if driven:
            sp["driver_state"] = (0, 0.25, 0.0)
            item["integrated_e_dot_j"] = {0: 1.5e-3}
        items.append(item)
    return items


def test_restart_dump_layout_and_round_trip(tmp_path):
    """RestartManager::write: dist_<n>.hdf with the items in registration order, the distribution in dist_<n>.hdf.g0"""
    items = _restart_items(np.random.default_rng(2))
    name = outputs.write_vp_restart(str(tmp_path / "planeEPW"), 3, items, 2, 1.25, 0.01, 0.9, 5.0)
    assert name.endswith("planeEPW/dist_3.hdf") and os.path.exists(name + ".g0")
    meta = h5lite.read(name)["root"]
    assert bytes(meta.attrs["className"]) == b"directory\0"
    for key, kind, val in (("species_list_size", ">i4", 2), ("isMaxwell", ">i4", 0), ("nGhost", ">i4", 2),
                           ("generating processes", ">i4", 1), ("major version", ">i4", 3), ("time", "<f8", 1.25),
                           ("time step", "<f8", 0.01), ("CFL", "<f8", 0.9), ("tf", "<f8", 5.0), ("bz_const", "<f8", 0.0)):
        d = meta[key].data
        assert d.shape == () and d.dtype == np.dtype(kind) and d == val, key     # H5S_SCALAR, RestartWriter.C:55-60
    assert bytes(meta["species_list"]["species.1"].data) == b"electron\0"         # length + 1 uchars, RestartWriter.C:414
    assert bytes(meta["species_list"]["species.2"].data) == b"ion\0"
    drv = meta["Shaped Ramped Cosine Driver1_1"]
    assert int(drv["num_phase_evals"].data) == 0 and float(drv["phase"].data) == 0.25
    el = meta["electron"]
    assert int(el["pdim"].data) == 4 and int(el["cdim"].data) == 2 and float(el["charge"].data) == -1.0
    assert el["N"].data.tolist() == [6, 4, 8, 6] and el["N"].data.dtype == np.dtype(">i4")
    assert el["x_lo"].data.tolist() == [-1.0, -2.0, -7.0, -7.0] and int(el["isPeriodic_1"].data) == 1
    assert el["distribution"]["distribInfo"].data.tolist() == items[0]["info"]
    assert el["x_lo_krook"].data.tolist() == [-1.0, -2.0] and int(el["krookHasLayer"].data) == 0
    assert el["x_hi_hi_external_dist_krook"].data.tolist() == [1.0, 2.0]
    bulk = h5lite.read(name + ".g0")["root"]
    assert sorted(bulk.names()) == ["electron\\distribution.p0", "electron\\integrated_e_dot_j.p0", "ion\\distribution.p0"]
    assert bulk["electron\\distribution.p0"].data.shape == (10, 12, 8, 10)        # (n4d, n3d, n2d, n1d), :536-552
    back = outputs.read_vp_restart(name)
    assert back["time"] == 1.25 and back["dt"] == 0.01 and back["n_ghosts"] == 2 and back["num_procs"] == 1
    for got, item in zip(back["species"], items):
        assert got["name"] == item["sp"]["name"] and got["mass"] == item["sp"]["mass"]
        assert np.array_equal(got["distribution"], item["tiles"][0])
    assert back["species"][0]["integrated_e_dot_j"] == 1.5e-3 and back["species"][1]["integrated_e_dot_j"] is None


def test_restart_bulk_files_of_many_ranks(tmp_path):
    """rank r writes into <base>.g<r * max_files / nprocs> as "<group>\\<name>.p<r>" (RestartWriter.C:27-29, :556-559)"""
    w = outputs.RestartWriter(str(tmp_path / "d" / "dist_0.hdf"), max_num_files=2, num_procs=4)
    w.push_sub_dir("electron")
    tiles = {r: np.full((2, 2, 3, 3), float(r)) for r in range(4)}
    tiles[3] = np.zeros((0, 2, 3, 3))                        # an empty dataBox is written as one 0.0 (:529-531)
    w.write_parallel_array("distribution", tiles, outputs.distrib_info(0, 3, 2, [4, 4, 2, 2], [2, 2, 1, 1]))
    w.pop_sub_dir()
    with pytest.raises(RuntimeError):
        w.pop_sub_dir()
    w.close()
    g0 = h5lite.read(str(tmp_path / "d" / "dist_0.hdf.g0"))["root"]
    g1 = h5lite.read(str(tmp_path / "d" / "dist_0.hdf.g1"))["root"]
    assert sorted(g0.names()) == ["electron\\distribution.p0", "electron\\distribution.p1"]
    assert sorted(g1.names()) == ["electron\\distribution.p2", "electron\\distribution.p3"]
    assert g1["electron\\distribution.p3"].data.shape == (1, 1, 1, 1)
    r = outputs.RestartReader(str(tmp_path / "d" / "dist_0.hdf"), 2)
    r.num_procs, r.max_num_files = 4, 2
    r.push_sub_dir("electron")
    info, t2 = r.read_parallel_array("distribution", 2)
    assert np.array_equal(t2, tiles[2]) and info[6:10].tolist() == [2, 2, 1, 1]


def test_deck_reader_keeps_the_restart_cadence():
    """RestartManager::parseParameters (RestartManager.C:150-192) through pp.py, on the reference's own deck"""
    from loki_b200 import pp
    deck_path = "/root/reference/test/planeEPW_fixedIons/planeEPW_fixedIons.pp"
    if not os.path.exists(deck_path):
        pytest.skip("the reference tree is not on this machine")
    deck = pp.load(deck_path)
    assert deck.run["restart"] == dict(time_interval=5.0, write_directory="planeEPW_fixedIons", start_from_restart=False)
    text = open(deck_path).read() + "\nrestart.step_interval = 3\n"
    with pytest.raises(ValueError, match="only one of steps or time"):
        pp.deck_from_params(pp.parse(text), name="t")


EXTERNAL_2D = """
# test/External2D/External2D.pp of the reference, restated (same numbers); the three "2D dist" files beside the deck
# are written by write_external2d_files from tests/golden/External2D_profiles.npz (the reference's rho_init_*.h5 data)
$mp_over_me = 1836.;
$m_he   = 4.*$mp_over_me;
$m_c    = 12.*$mp_over_me;
$vth_he = sqrt(1./$m_he);
$vth_c  = sqrt(1./$m_c);
$vmin_he = -7.5*$vth_he;
$vmax_he =  7.5*$vth_he;
$vmin_c  = -7.5*$vth_c;
$vmax_c  =  7.5*$vth_c;
$Nx = 128;
$Ny = 7;
$xa = -62.5;
$xb =  62.5;
$dx =  ($xb-$xa)/$Nx;
$ya = -0.5*$Ny*$dx;
$yb =  0.5*$Ny*$dx;
domain_limits = $xa $xb $ya $yb
N = $Nx $Ny
periodic_dir = true true
cfl = 0.95
final_time = 5.
save_times = 1.
sequence_write_times = .1
max_step = 1000000
spatial_solution_order = 6
temporal_solution_order = 6
number_of_species = 3
kinetic_species.1.name = "electron"
kinetic_species.1.velocity_limits = -7.5 7.5 -7.5 7.5
kinetic_species.1.Nv = 24 16
kinetic_species.1.mass = 1.0
kinetic_species.1.charge = -1.0
kinetic_species.1.ic.name = "External 2D"
kinetic_species.1.ic.file_name = "rho_init_e.h5"
kinetic_species.2.name = "He"
kinetic_species.2.velocity_limits = $vmin_he $vmax_he $vmin_he $vmax_he
kinetic_species.2.Nv = 24 16
kinetic_species.2.mass = $m_he
kinetic_species.2.charge = 2.
kinetic_species.2.ic.name = "External 2D"
kinetic_species.2.ic.file_name = "rho_init_He.h5"
kinetic_species.3.name = "C"
kinetic_species.3.velocity_limits = $vmin_c $vmax_c $vmin_c $vmax_c
kinetic_species.3.Nv = 24 16
kinetic_species.3.mass = $m_c
kinetic_species.3.charge = 6.
kinetic_species.3.ic.name = "External 2D"
kinetic_species.3.ic.file_name = "rho_init_C.h5"
number_of_probes = 0
"""


def write_external2d_files(directory):
    """the deck's three external files, written with h5lite from the committed profiles (tests/golden/make_golden.py)"""
    prof = np.load(os.path.join(HERE, "golden", "External2D_profiles.npz"))
    for n in ("e", "He", "C"):
        root = h5lite.Group()
        root.put("2D dist", prof[n])
        h5lite.write(os.path.join(str(directory), "rho_init_%s.h5" % n), root)
    return prof


def test_reader_reads_the_references_own_hdf5_files():
    """test/External2D/rho_init_*.h5: written by libhdf5 with compact new-style groups (Link messages in a version-1
    object header) -- a second, differently laid out pin of the reader, where the reference tree is present; the
    committed profiles are those files' data"""
    prof = np.load(os.path.join(HERE, "golden", "External2D_profiles.npz"))
    assert prof["e"].shape == (7 + 6, 128 + 6)                       # (Ny + 2 ng, Nx + 2 ng), order 6
    assert prof["e"].max() == 10.0 and prof["e"].min() == 0.05
    # the deck's plasma is neutral cell by cell: n_e = 2 n_He + 6 n_C
    assert np.max(np.abs(prof["e"] - 2.0 * prof["He"] - 6.0 * prof["C"])) <= 1e-14 * 10.0
    ref = "/root/reference/test/External2D"
    if not os.path.isdir(ref):
        pytest.skip("the reference tree is not on this machine")
    for n in ("e", "He", "C"):
        root = h5lite.read(os.path.join(ref, "rho_init_%s.h5" % n))
        assert root.names() == ["2D dist"]
        d = root["2D dist"].data
        assert d.dtype == np.dtype("<f8") and np.array_equal(d, prof[n])


def test_external_2d_deck_loads_like_the_reference_deck(tmp_path):
    """pp.py's "External 2D" initial condition (External2DIC.C): the spatial factor comes from the file, ghost cells of a
    periodic direction take their periodic image, f = frac fnorm fv fx; the restated deck equals the reference's own"""
    from loki_b200 import pp
    path = tmp_path / "External2D.pp"
    path.write_text(EXTERNAL_2D)
    prof = write_external2d_files(tmp_path)
    deck = pp.load(str(path))
    assert all(np.array_equal(sp.external, prof[n]) for sp, n in zip(deck.species, ("e", "He", "C")))
    assert deck.n == (128, 7) and deck.order == 6 and deck.rk == 6 and [s.name for s in deck.species] == ["electron", "He", "C"]
    ng = deck.ng
    total = 0.0
    for sp in deck.species:
        f, fx, fv, fnorm = deck.initial_state(sp)
        ext = sp.external
        assert np.array_equal(fx[ng:-ng, ng:-ng], ext[ng:-ng, ng:-ng])
        assert np.array_equal(fx[:, :ng], fx[:, -2 * ng:-ng]) and np.array_equal(fx[-ng:, :], fx[ng:2 * ng, :])   # periodic images
        n, dx = deck.geom_of(sp)
        total = total + sp.charge * f.sum(axis=(0, 1)) * dx[2] * dx[3]
    assert np.max(np.abs(total)) < 1e-12                      # neutral to rounding once integrated over velocity
    ref = "/root/reference/test/External2D/External2D.pp"
    if os.path.exists(ref):
        rdeck = pp.load(ref)
        assert rdeck.xlim == deck.xlim and rdeck.cfl == deck.cfl and rdeck.run["final_time"] == deck.run["final_time"]
        for a, b in zip(rdeck.species, deck.species):
            assert (a.name, a.nv, a.mass, a.charge, a.tx, a.ty) == (b.name, b.nv, b.mass, b.charge, b.tx, b.ty)
            assert a.vlim == b.vlim and np.array_equal(a.external, b.external)
    bad = tmp_path / "bad.pp"
    bad.write_text(EXTERNAL_2D.replace("N = $Nx $Ny", "N = 64 7"))
    with pytest.raises(ValueError, match="does not match configuration space"):
        d2 = pp.load(str(bad))
        d2.initial_state(d2.species[0])


def test_krook_layer_state_in_a_dump(tmp_path):
    """KrookLayer::putToDatabase (KrookLayer.C:249-282) from a deck's krook.* keys (parseParameters, :163-189)"""
    assert outputs.krook_state(None, [-1.0, -2.0], [1.0, 2.0]) is None
    k = outputs.krook_state(dict(x1a=-0.5, x2b=1.5, coefficient=2.0), [-1.0, -2.0, -7.0, -7.0], [1.0, 2.0, 7.0, 7.0])
    assert k == dict(x_lo=[-0.5, -2.0], x_hi=[1.0, 1.5], power=3.0, coefficient=2.0, has_lo=[1, 0], has_hi=[0, 1])
    items = _restart_items(np.random.default_rng(4))
    items[0]["krook"] = k
    name = outputs.write_vp_restart(str(tmp_path / "d"), 0, items, 2, 0.0, 0.01, 0.9, 1.0)
    el = h5lite.read(name)["root"]["electron"]
    assert el["x_lo_krook"].data.tolist() == [-0.5, -2.0] and el["x_hi_krook"].data.tolist() == [1.0, 1.5]
    assert float(el["krookCoeff"].data) == 2.0 and float(el["krookPower"].data) == 3.0
    assert [int(el["krookHasLayer" + t].data) for t in ("", "_lo_0", "_lo_1", "_hi_0", "_hi_1")] == [1, 1, 0, 0, 1]
    ion = h5lite.read(name)["root"]["ion"]
    assert int(ion["krookHasLayer"].data) == 0 and ion["x_lo_krook"].data.tolist() == [-1.0, -2.0]


def test_writer_structures_equal_libhdf5s_for_the_same_content(tmp_path):
    """the genuine libhdf5 file holds one dataset in the root group; the same content written by h5lite must come out as
    the same structures: superblock constants, root symbol-table entry, B-tree node, symbol-table node and local heap
    compared field by field (addresses aside), the heap's free list encoded by the same rule"""
    g = h5lite.read(GENUINE)
    root = h5lite.Group()
    root.put("testdouble", np.array(g["testdouble"].data))
    p = str(tmp_path / "same.hdf")
    h5lite.write(p, root)
    mine, theirs = open(p, "rb").read(), open(GENUINE, "rb").read()                # theirs: behind MATLAB's 512-byte user block

    def parts(b):
        r = h5lite._Reader(memoryview(b))
        e = r.root_entry
        sb = b.index(h5lite.SIGNATURE)
        out = dict(sb_versions=bytes(b[sb + 8:sb + 16]), ks=struct.unpack_from("<HH", b, sb + 16), root_cache=e["cache"])
        bt, hp = r.base + e["btree"], r.base + e["heap"]
        out["tree_head"] = bytes(b[bt:bt + 8])                                      # TREE, type 0, level 0, 1 entry
        out["tree_siblings"] = bytes(b[bt + 8:bt + 24])
        key0, child, key1 = struct.unpack_from("<QQQ", b, bt + 24)
        out["tree_keys"] = (key0, key1)
        sn = r.base + child
        out["snod_head"] = bytes(b[sn:sn + 8])                                      # SNOD, version 1, 1 symbol
        name_off, _, cache, resv = struct.unpack_from("<QQII", b, sn + 8)
        out["snod_entry"] = (name_off, cache, resv, bytes(b[sn + 24:sn + 40]))
        out["heap_head"] = bytes(b[hp:hp + 8])
        seg_size, free_head, seg = struct.unpack_from("<QQQ", b, hp + 8)
        seg += r.base
        out["heap_names"] = bytes(b[seg:seg + free_head])                           # "" then "testdouble", 8-byte padded
        nxt, fsize = struct.unpack_from("<QQ", b, seg + free_head)
        out["heap_free_rule"] = (nxt, fsize == seg_size - free_head)
        # the dataset's object header: version 1, reference count 1, 8-aligned message block
        oh = r.base + r.group_entries(e["btree"], e["heap"])[0][1]["oh"]
        ver, _, nmsg, refs, size = struct.unpack_from("<BBHII", b, oh)
        out["ohdr"] = (ver, refs, size % 8, oh % 8)
        msgs = {t: bytes(d) for t, _, d in r.messages(r.group_entries(e["btree"], e["heap"])[0][1]["oh"])}
        out["dtype_msg"], out["dspace_msg"] = msgs[0x0003][:20], msgs[0x0001]
        return out
    a, b = parts(mine), parts(theirs)
    assert a == b, {k: (a[k], b[k]) for k in a if a[k] != b[k]}


def test_every_regression_deck_of_the_reference_loads_or_is_rejected_loudly():
    """the reference's test/*/*.pp through pp.py: ten decks load (the systems the library runs), the two Rosenbluth decks
    are refused with a ValueError instead of running as something else"""
    import glob
    from loki_b200 import pp
    decks_ = sorted(glob.glob("/root/reference/test/*/*.pp"))
    if not decks_:
        pytest.skip("the reference tree is not on this machine")
    loaded, refused = [], []
    for f in decks_:
        try:
            pp.load(f)
            loaded.append(os.path.basename(f))
        except ValueError:
            refused.append(os.path.basename(f))
    assert loaded == ["EPWTZ.pp", "External2D.pp", "IAWTZ.pp", "InterpenetratingStreams.pp", "TrigTZ.pp", "emDamping.pp",
                      "pitchAngleCollisions.pp", "planeEPW_fixedIons.pp", "planeIAW.pp", "planeIAW_6.pp"]
    assert refused == ["rosenbluthCollisions_no_br.pp", "rosenbluthCollisions_w_br.pp"]


def test_flow_velocity_of_the_initial_maxwellian(tmp_path):
    """ic.vflowinitx / vflowinity (MaxwellianThermal.C:40-59): a constant drift of the Maxwellian, also for factorable ICs"""
    from loki_b200 import pp
    from test_gpu_vp_system import RUN_DECK
    text = RUN_DECK + "kinetic_species.1.ic.vflowinitx = 0.75\nkinetic_species.1.ic.vflowinity = -0.5\n"
    path = tmp_path / "flow.pp"
    path.write_text(text)
    deck = pp.load(str(path))
    sp = deck.species[0]
    assert (sp.vflowinitx, sp.vflowinity) == (0.75, -0.5) and deck.species[1].vflowinitx == 0.0
    f, fx, fv, fnorm = deck.initial_state(sp)
    ng = deck.ng
    n, dx = deck.geom_of(sp)
    x3 = sp.vlim[0] + (np.arange(-ng, sp.nv[0] + ng) + 0.5) * dx[2]
    x4 = sp.vlim[2] + (np.arange(-ng, sp.nv[1] + ng) + 0.5) * dx[3]
    want = np.exp(-0.5 * (((x3 - 0.75) ** 2)[None, :] / (sp.tx / sp.mass) + ((x4 + 0.5) ** 2)[:, None] / (sp.ty / sp.mass)))
    assert np.array_equal(fv, want)
    # the mean velocity of the state is the drift (up to the truncation of the Maxwellian at the velocity limits)
    I = (slice(ng, -ng),) * 4
    w = f[I].sum(axis=(2, 3))
    assert abs((w * x3[ng:-ng][None, :]).sum() / w.sum() - 0.75) < 1e-2       # midpoint rule on a coarse velocity grid
    assert abs((w * x4[ng:-ng][:, None]).sum() / w.sum() + 0.5) < 1e-2
    # the Rosenbluth decks of the reference carry a flow velocity too; they are refused for their collision operator now
    ref = "/root/reference/test/rosenbluthCollisions/rosenbluthCollisions_w_br.pp"
    if os.path.exists(ref):
        with pytest.raises(ValueError, match="collision operator"):
            pp.load(ref)
