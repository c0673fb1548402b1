"""The reference's regression chain after a run (test/*/Makefile.in: vlasovPoisson4D -> vp4DPostProcess -> checkTests):

  post_process(prefix)      vp4DPostProcess.C:393-679.  Reads what a run wrote -- <prefix>/dist_<n>.hdf(.g<k>),
                            <prefix>.fields_<k>.hdf, <prefix>.time_hists_<n>.hdf -- and writes the serial files the
                            reference's tools consume: <prefix>_dist_<n>.hdf + .g0 (every rank's tile assembled into one
                            global array), <prefix>_fields.hdf (one (slices, Ny, Nx) dataset per field, ghosts removed,
                            plus x, y, time), <prefix>_timeSeries.hdf (the last time-history file, `series_time`).
  check_tests(...)          checkTests.C:72-653.  Compares the post-processed files of a test run with a baseline's under
                            the tolerance files of the deck's directory (dist_tol, tstol, field_tol: name / tolerance line
                            pairs) with the reference's relative-difference metric (:345-358).  The reference hard-codes
                            its baseline directory; here it is an argument.

Host-side file handling only (h5lite.py / outputs.py)."""
import os

import numpy as np

from . import h5lite, outputs


def rel_diff_max(test, baseline):
    """the metric of checkTests.C:345-358 (and :115-134, :190-209): |test - baseline| / |test|, with |baseline| as the
    denominator where test == 0, and 0 where both are"""
    test = np.asarray(test, dtype=np.float64).ravel()
    baseline = np.asarray(baseline, dtype=np.float64).ravel()
    if test.size == 0:
        return 0.0
    den = np.where(test == 0.0, np.abs(baseline), np.abs(test))
    diff = np.abs(test - baseline)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(den == 0.0, 0.0, diff / den)
    return float(r.max())


def _tile_box(info, rank, dim=4):
    """the interior box (lower, upper inclusive) of `rank` from distribInfo: ParallelArray::getIndexRank (dimension 0
    varies slowest, ParallelArray.C:609-626) and setupLocalDomain's split (:641-661)"""
    proc_lo = int(info[0])
    nproc = [int(v) for v in info[2 + dim:2 + 2 * dim]]
    left_size = [int(v) for v in info[2 + 2 * dim:2 + 3 * dim]]
    num_left = [int(v) for v in info[2 + 3 * dim:2 + 4 * dim]]
    r = rank - proc_lo
    idx = []
    for d in range(dim):
        factor = 1
        for j in range(d + 1, dim):
            factor *= nproc[j]
        idx.append(r // factor)
        r -= idx[-1] * factor
    lower, upper = [], []
    for d in range(dim):
        if idx[d] < num_left[d]:
            lo = idx[d] * left_size[d]
            n = left_size[d]
        else:
            # the partitions right of the `num_left` larger ones hold one cell less
            lo = num_left[d] * left_size[d] + (idx[d] - num_left[d]) * (left_size[d] - 1)
            n = left_size[d] - 1
        lower.append(lo)
        upper.append(lo + n - 1)
    return lower, upper


def assemble_parallel_array(reader, name, num_cells, n_ghosts):
    """RestartReader::readParallelArray for the post processor (RestartReader.C:486-613): every generating rank's tile
    copied into the serial array of extents num_cells + 2 ghosts (returned C-ordered (n4, n3, n2, n1) like a dump)"""
    info = reader.groups[-1][name]["distribInfo"].data.astype(np.int64)
    dim = 4
    ng = n_ghosts
    out = np.zeros(tuple(int(n) + 2 * ng for n in reversed(num_cells)))
    for rank in range(int(info[0]), int(info[1]) + 1):
        tile = reader._bulk_root(rank)["%s\\%s.p%d" % (reader.group_names[-1], name, rank)].data
        lower, upper = _tile_box(info, rank, dim)
        ext = tuple(upper[d] - lower[d] + 1 + 2 * ng for d in reversed(range(dim)))
        if tile.shape != ext:
            if tile.size == 1:
                continue                                    # an empty dataBox, written as one 0.0 (RestartWriter.C:529-531)
            raise ValueError("tile of rank %d has extents %s, distribInfo says %s" % (rank, tile.shape, ext))
        # the tile's dataBox lands at its lower corner; tiles are visited in rank order, so a neighbour's interior
        # overwrites this tile's ghost copy of it and the domain-boundary ghosts come from the tiles that own them
        dst = tuple(slice(lower[d], lower[d] + ext[dim - 1 - d]) for d in reversed(range(dim)))
        out[dst] = tile
    for rank in range(int(info[0]), int(info[1]) + 1):      # second pass: interiors win over any neighbour's ghosts
        tile = reader._bulk_root(rank)["%s\\%s.p%d" % (reader.group_names[-1], name, rank)].data
        if tile.size == 1 and out.size != 1:
            continue
        lower, upper = _tile_box(info, rank, dim)
        src = tuple(slice(ng, ng + upper[d] - lower[d] + 1) for d in reversed(range(dim)))
        dst = tuple(slice(lower[d] + ng, upper[d] + 1 + ng) for d in reversed(range(dim)))
        out[dst] = tile[src]
    return out


def _restart_indices(prefix):
    idx = 0
    while os.path.exists(os.path.join(prefix, "dist_%d.hdf" % idx)):
        yield idx
        idx += 1


def _process_dist(prefix, idx):
    base = os.path.join(prefix, "dist_%d.hdf" % idx)
    nfiles = 0
    while os.path.exists("%s.g%d" % (base, nfiles)):
        nfiles += 1
    r = outputs.RestartReader(base, nfiles)
    r.max_num_files = min(nfiles, r.num_procs)
    t = r.read_double_value("time")
    ng = r.read_integer_value("nGhost")
    flux = r.read_integer_value("plot_ke_vel_bdy_flux")
    is_maxwell = r.read_integer_value("isMaxwell")
    ns = r.read_integer_value("species_list_size")
    r.push_sub_dir("species_list")
    names = [r.read_string("species.%d" % (s + 1)) for s in range(ns)]
    r.pop_sub_dir()
    w = outputs.RestartWriter("%s_dist_%d.hdf" % (prefix, idx), 1, 1)
    w.write_double_value("time", t)
    w.write_integer_value("nGhost", ng)
    w.write_integer_value("plot_ke_vel_bdy_flux", flux)
    w.write_integer_value("species_list_size", ns)
    w.push_sub_dir("species_list")
    for s, n in enumerate(names):
        w.write_string("species.%d" % (s + 1), n)
    w.pop_sub_dir()
    nxy = None
    for s, n in enumerate(names):
        # KineticSpecies(reader, ...) then putToRestart_SkipKrook (KineticSpecies.C:957-1020)
        drv = "%s%d_%d" % (outputs.DRIVER_CLASS_NAME, s + 1, 1)
        if drv in r.groups[-1]:
            r.push_sub_dir(drv)
            state = (r.read_integer_value("num_phase_evals"), r.read_double_value("phase"), r.read_double_value("phase_h"))
            r.pop_sub_dir()
            w.push_sub_dir(drv)
            w.write_integer_value("num_phase_evals", state[0])
            w.write_double_value("phase", state[1])
            w.write_double_value("phase_h", state[2])
            w.pop_sub_dir()
        r.push_sub_dir(n)
        num_cells = [int(v) for v in r.read_integer_array("N")]
        nxy = nxy or (num_cells[0], num_cells[1])
        f = assemble_parallel_array(r, "distribution", num_cells, ng)
        w.push_sub_dir(n)
        w.write_integer_value("pdim", r.read_integer_value("pdim"))
        w.write_integer_value("cdim", r.read_integer_value("cdim"))
        for key in ("mass", "charge", "bz_const"):
            w.write_double_value(key, r.read_double_value(key))
        outputs.put_problem_domain(w, num_cells, r.read_double_array("x_lo"), r.read_double_array("x_hi"),
                                   r.read_double_array("dx"), (r.read_integer_value("isPeriodic_0"), r.read_integer_value("isPeriodic_1")))
        w.write_parallel_array("distribution", {0: f}, outputs.distrib_info(0, 0, ng, num_cells, [1, 1, 1, 1]))
        # m_integrated_ke_e_dot: every generating rank's share (KineticSpecies.C:1011-1015), summed as getSum would
        parts = [r.read_bulk_double_value("integrated_e_dot_j", rank) for rank in range(r.num_procs)
                 if ("%s\\integrated_e_dot_j.p%d" % (n, rank)) in r._bulk_root(rank)]
        if parts:
            w.write_bulk_double_value("integrated_e_dot_j", {0: float(sum(parts))})
        w.pop_sub_dir()
        r.pop_sub_dir()
    w.close()
    return dict(n_ghosts=ng, is_maxwell=is_maxwell, plot_ke_vel_bdy_flux=flux, species=names, nxy=nxy)


def write_time_series(prefix, is_maxwell, species_names):
    """writeTimeHistories (vp4DPostProcess.C:100-210): the LAST time-history file holds every sequence so far"""
    idx = -1
    while os.path.exists("%s.time_hists_%d.hdf" % (prefix, idx + 1)):
        idx += 1
    if idx < 0:
        raise FileNotFoundError(prefix + ".time_hists_0.hdf")
    th_reader = outputs.TimeHistReader("%s.time_hists_%d.hdf" % (prefix, idx))
    root = th_reader.root
    nprobes = th_reader.read_num_probes() if "numProbes" in root else 0
    nparticles = th_reader.read_num_tracking_particles() if "numTrackingParticles" in root else 0
    if is_maxwell:
        # the Maxwell system writes the histories its device side computes (run.py); every dataset but the bookkeeping
        names = [n for n in root.names() if n not in ("sequence_times", "numProbes", "numTrackingParticles")]
    else:
        names = outputs.poisson_time_history_names(nprobes, nparticles, species_names)
    w = outputs.TimeHistWriter(prefix + "_timeSeries.hdf", th_reader.read_time_history("sequence_times"), for_post_proc=True)
    for n in names:
        w.write_time_history(n, th_reader.read_time_history(n))
    w.close()
    return w.name


def write_fields(prefix, nx, ny, n_ghosts, is_maxwell, species_names, plot_ke_vel_bdy_flux):
    """writeFields (vp4DPostProcess.C:215-385)"""
    total = outputs.FieldReader(prefix + ".fields_0.hdf").read_total_num_time_slices()
    names = (outputs.maxwell_plot_names if is_maxwell else outputs.poisson_plot_names)(bool(plot_ke_vel_bdy_flux), species_names)
    top, root = outputs.ReaderWriterBase.create_file_and_root()
    data = {n: np.zeros((total, ny, nx)) for n in names}
    times = []
    which = 0
    k = 0
    ng = n_ghosts
    while which < total and os.path.exists("%s.fields_%d.hdf" % (prefix, k)):
        fr = outputs.FieldReader("%s.fields_%d.hdf" % (prefix, k))
        if k == 0:
            x, y = fr.read_coords()
        for _ in range(fr.read_num_time_slices_in_file()):
            times.append(fr.read_time("time_slice_%d_time" % which))
            for n in names:
                plane = fr.read_field("time_slice_%d_%s" % (which, n)).reshape(ny + 2 * ng, nx + 2 * ng)
                data[n][which] = plane[ng:ng + ny, ng:ng + nx]
            which += 1
            if which == total:
                break
        k += 1
    for n in names:
        root.put(n, data[n] if total != 1 else data[n][0])   # createFieldDatasets: 2-D when there is one slice (:696-709)
    outputs.ReaderWriterBase.write_double_array("x", root, x)
    outputs.ReaderWriterBase.write_double_array("y", root, y)
    outputs.ReaderWriterBase.write_double_array("time", root, times)
    h5lite.write(prefix + "_fields.hdf", top)
    return prefix + "_fields.hdf"


def post_process(prefix, skip_dists=False):
    """vp4DPostProcess -prefix=<prefix> [-skip_dists]"""
    meta = None
    if skip_dists:
        r = outputs.RestartReader(os.path.join(prefix, "dist_0.hdf"))
        ns = r.read_integer_value("species_list_size")
        r.push_sub_dir("species_list")
        names = [r.read_string("species.%d" % (s + 1)) for s in range(ns)]
        r.pop_sub_dir()
        r.push_sub_dir(names[0])
        n = r.read_integer_array("N")
        r.pop_sub_dir()
        meta = dict(n_ghosts=r.read_integer_value("nGhost"), is_maxwell=r.read_integer_value("isMaxwell"),
                    plot_ke_vel_bdy_flux=r.read_integer_value("plot_ke_vel_bdy_flux"), species=names, nxy=(int(n[0]), int(n[1])))
    else:
        for idx in _restart_indices(prefix):
            meta = _process_dist(prefix, idx)
        if meta is None:
            raise FileNotFoundError(os.path.join(prefix, "dist_0.hdf"))
    write_time_series(prefix, meta["is_maxwell"], meta["species"])
    write_fields(prefix, meta["nxy"][0], meta["nxy"][1], meta["n_ghosts"], meta["is_maxwell"], meta["species"],
                 meta["plot_ke_vel_bdy_flux"])
    return meta


def _tolerances(path):
    """name / tolerance line pairs (checkTests.C:88-99)"""
    lines = [ln.rstrip("\n") for ln in open(path)]
    out = []
    for k in range(0, len(lines) - 1, 2):
        if lines[k] == "":
            break
        out.append((lines[k], float(lines[k + 1])))
    return out


def check_tests(test_prefix, baseline_prefix, file_index, dist_tol=None, ts_tol=None, field_tol=None):
    """checkTests.C main: the list of failure messages (empty = PASSED, as the Makefiles test `-s diffs`)"""
    fails = []
    tname, bname = "%s_dist_%d.hdf" % (test_prefix, file_index), "%s_dist_%d.hdf" % (baseline_prefix, file_index)
    t, b = h5lite.read(tname)["root"], h5lite.read(bname)["root"]
    ng = int(t["nGhost"].data)
    if ng != int(b["nGhost"].data):
        return ["FAILED: Baseline and test have different number of ghosts."]
    ns = int(t["species_list_size"].data)
    if ns != int(b["species_list_size"].data):
        return ["FAILED: Baseline and test have different number of species."]
    names = [bytes(t["species_list"]["species.%d" % (s + 1)].data).split(b"\0")[0].decode() for s in range(ns)]
    for n in names:
        if t[n]["N"].data.tolist() != b[n]["N"].data.tolist():
            return ["FAILED: Baseline and test have different computational domains."]
    if dist_tol:
        tb, bb = h5lite.read(tname + ".g0")["root"], h5lite.read(bname + ".g0")["root"]
        for n, tol in _tolerances(dist_tol):
            I = (slice(ng, -ng),) * 4
            d = rel_diff_max(tb[n + "\\distribution.p0"].data[I], bb[n + "\\distribution.p0"].data[I])
            if d > tol:
                fails.append("Maximum relative difference for species %s: %g exceeds tolerance %g" % (n, d, tol))
    if field_tol:
        tf, bf = h5lite.read(test_prefix + "_fields.hdf")["root"], h5lite.read(baseline_prefix + "_fields.hdf")["root"]
        for n, tol in _tolerances(field_tol):
            if tf[n].data.size != bf[n].data.size:
                fails.append("FAILED. Test and baseline field sizes differ.")
                continue
            d = rel_diff_max(tf[n].data, bf[n].data)
            if d > tol:
                fails.append("Maximum relative difference for field %s: %g exceeds tolerance %g" % (n, d, tol))
    if ts_tol:
        tt, bt = h5lite.read(test_prefix + "_timeSeries.hdf")["root"], h5lite.read(baseline_prefix + "_timeSeries.hdf")["root"]
        for n, tol in _tolerances(ts_tol):
            if tt[n].data.size != bt[n].data.size:
                fails.append("FAILED. Test and baseline timeseries sizes differ.")
                continue
            d = rel_diff_max(tt[n].data, bt[n].data)
            if d > tol:
                fails.append("Maximum relative difference for time series %s: %g exceeds tolerance %g" % (n, d, tol))
    return fails


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="vp4DPostProcess / checkTests on the files a loki_b200.run run wrote")
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("post")
    p.add_argument("-prefix", "--prefix", required=True)
    p.add_argument("-skip_dists", "--skip-dists", action="store_true")
    c = sub.add_parser("check")
    c.add_argument("input", help="the test directory's `input` file: dist_tol tstol field_tol dir prefix index coll")
    c.add_argument("--baselines", required=True, help="directory holding the baseline's post-processed files")
    a = ap.parse_args(argv)
    if a.cmd == "post":
        post_process(a.prefix, a.skip_dists)
        return 0
    dist_tol, ts_tol, field_tol, test_dir, prefix, index, _coll = open(a.input).read().split()[:7]
    here = os.path.dirname(os.path.abspath(a.input))
    fails = check_tests(os.path.join(here, prefix), os.path.join(a.baselines, test_dir, prefix), int(index),
                        os.path.join(here, dist_tol), os.path.join(here, ts_tol), os.path.join(here, field_tol))
    for f in fails:
        print(f)
    return 1 if fails else 0


if __name__ == "__main__":
    import sys
    sys.exit(main())
