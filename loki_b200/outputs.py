"""LOKI's on-disk formats: restart dumps, field plot files and time-history files.

Host-side mirror of the reference's writer classes, same names, same call order, same datasets:

  ReaderWriterBase        ReaderWriterBase.C:24-330      every file holds one group "root" carrying the uchar[10]
                                                         attribute className = "directory"; doubles are
                                                         H5T_NATIVE_DOUBLE, integers H5T_STD_I32BE
  TimeHistWriter          TimeHistWriter.C:15-139        <write_dir>.time_hists_<n>.hdf
  FieldWriter             FieldWriter.C:14-760           <write_dir>.fields_<k>.hdf, one (Ny + order, Nx + order) dataset
                                                         per field and time slice
  RestartWriter           RestartWriter.C:15-613         <write_dir>/dist_<n>.hdf (metadata) + .g<k> (bulk data)
  RestartReader           RestartReader.C                the inverse, for resuming a run
  FieldReader             FieldReader.C:15-104           what the post processor reads the field series with
  TimeHistReader          TimeHistReader.C:15-52         ... and the time histories

plus the putToRestart methods of Simulation / VPSystem / Poisson / KineticSpecies / ProblemDomain / KrookLayer /
ExternalDistKrookLayer / ShapedRampedCosineDriver that decide what goes into a dump.  The HDF5 bytes come from
h5lite.py (libhdf5 is absent from this image).  Nothing here touches the GPU: arrays arrive as numpy arrays in the
reference's dataBox layout, i.e. what lk_vp_get_state / lk_vp_em_vars_ptr hand out."""
import math
import os

import numpy as np

from . import h5lite

_DIRECTORY = np.frombuffer(b"directory\0", dtype="u1")


class ReaderWriterBase:
    """createFileAndRoot / createGroup / write*Value / write*Array of ReaderWriterBase.C"""

    @staticmethod
    def create_file_and_root():
        top = h5lite.Group()
        return top, ReaderWriterBase.create_group("root", top)

    @staticmethod
    def create_group(name, home):
        if name in home:
            raise KeyError("Can not create group.")            # H5Gcreate on an existing name
        g = home.group(name)
        g.attrs["className"] = _DIRECTORY                      # ReaderWriterBase.C:118-160
        return g

    @staticmethod
    def write_double_array(name, home, vals, shape=None):
        a = np.asarray(vals, dtype="<f8")
        home.put(name, a if shape is None else a.reshape(shape))

    @staticmethod
    def write_double_value(name, home, val):
        home.put(name, np.array([val], dtype="<f8"))           # a 1-element simple dataspace, ReaderWriterBase.C:247-250

    @staticmethod
    def write_integer_array(name, home, vals):
        home.put(name, np.asarray(vals, dtype=">i4"))

    @staticmethod
    def write_integer_value(name, home, val):
        home.put(name, np.array([val], dtype=">i4"))


class TimeHistWriter(ReaderWriterBase):
    """TimeHistWriter.C: `sequence_times` (or `series_time` for the post processor), numProbes,
    numTrackingParticles, then one 1-D dataset per time history"""

    def __init__(self, name, seq_times, num_seq_times=None, num_probes=None, num_tracking_particles=None,
                 for_post_proc=False):
        self.name = name
        self.top, self.root = self.create_file_and_root()
        n = len(seq_times) if num_seq_times is None else num_seq_times
        self.write_double_array("series_time" if for_post_proc else "sequence_times", self.root, np.asarray(seq_times)[:n])
        if num_probes is not None:
            self.write_integer_value("numProbes", self.root, num_probes)
            self.write_integer_value("numTrackingParticles", self.root, num_tracking_particles or 0)

    def write_time_history(self, hist_name, time_hist, num_seq_times=None):
        n = len(time_hist) if num_seq_times is None else num_seq_times
        self.write_double_array(hist_name, self.root, np.asarray(time_hist)[:n])

    def close(self):
        h5lite.write(self.name, self.top)


class FieldWriter(ReaderWriterBase):
    """FieldWriter.C: a series of files <base>.fields_<k>.hdf with `time_slices_per_file` slices each.  A field of a
    slice is the dataset time_slice_<n>_<field name> of extent (Ny + order, Nx + order): the configuration-space
    array with its ghost layers, each rank writing its localBox part at offset lower + n_ghosts (:225-300)."""

    def __init__(self, base_name, x_lo, dx, num_cells, solution_order, time_slices_per_file=1):
        self.base_name = base_name
        self.per_file = time_slices_per_file
        self.current_file_index = -1
        self.written_to_current = 0
        self.total_written = 0
        self.processing = False
        self.nx, self.ny = num_cells[0] + solution_order, num_cells[1] + solution_order
        self.x = x_lo[0] + (np.arange(num_cells[0]) + 0.5) * dx[0]
        self.y = x_lo[1] + (np.arange(num_cells[1]) + 0.5) * dx[1]
        self.files = {}     # file index -> (top, root): h5lite lays a file out in one pass, so open files stay in memory

    def file_name(self, index):
        return "%s.fields_%d.hdf" % (self.base_name, index)

    def start_time_slice(self, time, dt, field_names, num_tracking_particles=None, num_probes=None, probes=None,
                         num_cells=None):
        if self.processing:
            raise RuntimeError("Time slice already being processed.  Missing call to endTimeSlice.")
        self.processing = True
        new_file = self.written_to_current == 0 or self.written_to_current == self.per_file
        if new_file:
            self.current_file_index += 1
            self.written_to_current = 0
            top, root = self.create_file_and_root()
            self.files[self.current_file_index] = (top, root)
            self.write_double_array("x", root, self.x)      # the coordinates are written once to each field file
            self.write_double_array("y", root, self.y)
        top, root = self.files[self.current_file_index]
        self.root = root
        for name in field_names:                            # createFieldDatasets: created, not yet written (zeros)
            root.put("time_slice_%d_%s" % (self.total_written, name), np.zeros((self.ny, self.nx)))
        self.write_double_value("time_slice_%d_time" % self.total_written, root, time)
        self.write_double_value("time_slice_%d_dt" % self.total_written, root, dt)
        if new_file and num_probes is not None:
            self.write_integer_value("numTrackingParticles", root, num_tracking_particles or 0)
            self.write_integer_value("numProbes", root, num_probes)
            for ip in range(num_probes):
                self.write_integer_value("ix_probe%d" % ip, root, int(math.floor(probes[0][ip] * num_cells[0])))
                self.write_integer_value("iy_probe%d" % ip, root, int(math.floor(probes[1][ip] * num_cells[1])))

    def write_field(self, field_name, field, in_lower, out_lower, out_cells, n_ghosts):
        """`field`: this rank's (n2, n1) array whose first cell is global cell in_lower = (i, j) (dataBox); the part
        starting at out_lower of out_cells = (nx, ny) cells (localBox) is stored (FieldWriter.C:225-280)"""
        if not self.processing:
            raise RuntimeError("Time slice not being processed.  Missing call to startTimeSlice.")
        dset = self.root["time_slice_%d_%s" % (self.total_written, field_name)].data
        xs, ys = out_lower[0] - in_lower[0], out_lower[1] - in_lower[1]
        part = np.asarray(field)[ys:ys + out_cells[1], xs:xs + out_cells[0]]
        oy, ox = out_lower[1] + n_ghosts, out_lower[0] + n_ghosts
        dset[oy:oy + out_cells[1], ox:ox + out_cells[0]] = part

    def end_time_slice(self):
        if not self.processing:
            raise RuntimeError("Time slice not being processed.  Missing call to startTimeSlice.")
        self.written_to_current += 1
        self.total_written += 1
        self.processing = False
        # writeNumTimeSlices (:482-672): the running total lives in the first file of the series
        first = self.files[0][1]
        first.children.pop("total_num_time_slices", None)
        self.write_integer_value("total_num_time_slices", first, self.total_written)
        cur = self.files[self.current_file_index][1]
        cur.children.pop("num_time_slices_in_this_file", None)
        self.write_integer_value("num_time_slices_in_this_file", cur, self.written_to_current)
        for index in {0, self.current_file_index}:
            h5lite.write(self.file_name(index), self.files[index][0])
        for index in [k for k in self.files if k not in (0, self.current_file_index)]:
            del self.files[index]                           # complete files need not stay in memory


class TimeHistReader:
    """TimeHistReader.C:15-52: one time-history file (raw `.time_hists_<n>.hdf` or post-processed `_timeSeries.hdf`)"""

    def __init__(self, name):
        self.root = h5lite.read(name)["root"]

    def read_time_history(self, hist_name):
        return np.array(self.root[hist_name].data, dtype=np.float64)

    def read_num_probes(self):
        return int(self.root["numProbes"].data[0])

    def read_num_tracking_particles(self):
        return int(self.root["numTrackingParticles"].data[0])


class FieldReader:
    """FieldReader.C:15-104: one field file of the series (or the post-processed `_fields.hdf`)"""

    def __init__(self, name):
        self.root = h5lite.read(name)["root"]

    def read_time(self, name):
        return float(self.root[name].data.reshape(-1)[0])

    def read_coords(self):
        return np.array(self.root["x"].data), np.array(self.root["y"].data)

    def read_field(self, field_name):
        """the dataset flattened, as the reference hands it back (a vector of Ny * Nx values, x fastest)"""
        return np.array(self.root[field_name].data, dtype=np.float64).reshape(-1)

    def read_total_num_time_slices(self):
        return int(self.root["total_num_time_slices"].data[0])

    def read_num_time_slices_in_file(self):
        return int(self.root["num_time_slices_in_this_file"].data[0])


def distrib_info(proc_lo, proc_hi, n_ghosts, num_cells, dim_partitions):
    """the `distribInfo` integers of RestartWriter::writeParallelArray (:494-510) for an array of len(num_cells)
    dimensions split dim_partitions[d] ways (ParallelArray::setupLocalDomain, ParallelArray.C:641-653)"""
    dim = len(num_cells)
    left_size, num_left = [], []
    for d in range(dim):
        nloc, extra = divmod(num_cells[d], dim_partitions[d])
        if extra == 0:
            left_size.append(nloc)
            num_left.append(dim_partitions[d])
        else:
            left_size.append(nloc + 1)
            num_left.append(extra)
    return [proc_lo, proc_hi] + [-n_ghosts] * dim + list(dim_partitions) + left_size + num_left


class RestartWriter(ReaderWriterBase):
    """RestartWriter.C: metadata file <base> with nested groups of scalars, bulk-data files <base>.g<k> holding each
    rank's dataBox as the dataset "<group>\\<name>.p<rank>" (:520-575).  One process plays all ranks here: the
    caller hands write_parallel_array every rank's tile."""

    def __init__(self, base_name, max_num_files=16, num_procs=1):
        self.base_name = base_name
        self.num_procs = num_procs
        self.max_num_files = min(max_num_files, num_procs)     # RestartReaderWriterBase.C:33
        self.top, root = self.create_file_and_root()
        self.groups = [root]
        self.group_names = ["root"]
        self.bulk = {}
        for rank in range(num_procs):
            k = self.bulk_file_num(rank)
            if k not in self.bulk:
                self.bulk[k] = self.create_file_and_root()

    def bulk_file_num(self, rank):
        return rank * self.max_num_files // self.num_procs       # RestartWriter.C:27-29

    def push_sub_dir(self, name):
        self.groups.append(self.create_group(name, self.groups[-1]) if name not in self.groups[-1] else self.groups[-1][name])
        self.group_names.append(name)

    def pop_sub_dir(self):
        if len(self.groups) == 1:
            raise RuntimeError("Attempting to pop the root group.")
        self.groups.pop()
        self.group_names.pop()

    # scalars are H5S_SCALAR dataspaces here, unlike ReaderWriterBase's 1-element arrays (RestartWriter.C:55-60, :180)
    def write_integer_value(self, name, val):
        self.groups[-1].put(name, np.array(val, dtype=">i4"), scalar=True)

    def write_integer_array(self, name, vals):
        self.groups[-1].put(name, np.asarray(vals, dtype=">i4").reshape(-1))

    def write_double_value(self, name, val):
        self.groups[-1].put(name, np.array(val, dtype="<f8"), scalar=True)

    def write_double_array(self, name, vals):
        self.groups[-1].put(name, np.asarray(vals, dtype="<f8").reshape(-1))

    def write_string(self, name, val):
        self.groups[-1].put(name, np.frombuffer(val.encode() + b"\0", dtype="u1"))   # length + 1, RestartWriter.C:414

    def write_bulk_double_value(self, name, vals_by_rank):
        for rank, v in vals_by_rank.items():
            root = self.bulk[self.bulk_file_num(rank)][1]
            ReaderWriterBase.write_double_value("%s\\%s.p%d" % (self.group_names[-1], name, rank), root, v)

    def write_parallel_array(self, name, tiles_by_rank, info):
        """tiles_by_rank[r]: rank r's dataBox as a C-ordered array of extents (n_dim ... n_1), which is the Fortran
        array's memory order and the dataset's dims (:536-552); an empty tile is written as one 0.0 (:529-531)"""
        self.push_sub_dir(name)
        self.write_integer_array("distribInfo", info)
        self.pop_sub_dir()
        for rank, a in tiles_by_rank.items():
            root = self.bulk[self.bulk_file_num(rank)][1]
            a = np.asarray(a, dtype="<f8")
            if a.size == 0:
                a = np.zeros((1,) * a.ndim)
            root.put("%s\\%s.p%d" % (self.group_names[-1], name, rank), a)

    def close(self):
        os.makedirs(os.path.dirname(self.base_name) or ".", exist_ok=True)
        h5lite.write(self.base_name, self.top)
        for k, (top, _) in self.bulk.items():
            h5lite.write("%s.g%d" % (self.base_name, k), top)


class RestartReader:
    """RestartReader.C: the same navigation over a dump that RestartWriter (or LOKI itself) wrote"""

    def __init__(self, base_name, max_num_files=16):
        self.base_name = base_name
        self.groups = [h5lite.read(base_name)["root"]]
        self.group_names = ["root"]
        self.num_procs = int(self.groups[0]["generating processes"].data) if "generating processes" in self.groups[0] else 1
        self.max_num_files = min(max_num_files, self.num_procs)
        self.bulk = {}

    def push_sub_dir(self, name):
        self.groups.append(self.groups[-1][name])
        self.group_names.append(name)

    def pop_sub_dir(self):
        self.groups.pop()
        self.group_names.pop()

    def read_integer_value(self, name):
        return int(self.groups[-1][name].data.reshape(-1)[0])

    def read_double_value(self, name):
        return float(self.groups[-1][name].data.reshape(-1)[0])

    def read_integer_array(self, name):
        return self.groups[-1][name].data.astype(np.int64)

    def read_double_array(self, name):
        return np.array(self.groups[-1][name].data, dtype=np.float64)

    def read_string(self, name):
        return bytes(self.groups[-1][name].data).split(b"\0")[0].decode()

    def _bulk_root(self, rank):
        k = rank * self.max_num_files // self.num_procs
        if k not in self.bulk:
            self.bulk[k] = h5lite.read("%s.g%d" % (self.base_name, k))["root"]
        return self.bulk[k]

    def read_bulk_double_value(self, name, rank=0):
        return float(self._bulk_root(rank)["%s\\%s.p%d" % (self.group_names[-1], name, rank)].data.reshape(-1)[0])

    def read_parallel_array(self, name, rank=0):
        """(distribInfo, rank's dataBox)"""
        info = self.groups[-1][name]["distribInfo"].data.astype(np.int64)
        return info, self._bulk_root(rank)["%s\\%s.p%d" % (self.group_names[-1], name, rank)].data


# ------------------------------------------------------------------ what a run writes

DRIVER_CLASS_NAME = "Shaped Ramped Cosine Driver"      # ShapedRampedCosineDriver.C:20-21
LOKI_VERSION = (3, 0, 1)                               # Simulation.C:22-24


def poisson_time_history_names(num_probes, num_tracking_particles, species_names):
    """Poisson::buildTimeHistoryNames (Poisson.C:922-1014)"""
    names = ["E_max", "norm E", "Ex_max", "Ey_max", "field_energy"]
    for ip in range(num_probes):
        names += ["Ex_probe%d" % ip, "Ey_probe%d" % ip]
    for ip in range(num_tracking_particles):
        names += ["particle%d_%s" % (ip, c) for c in ("x", "y", "vx", "vy")]
    for sp in species_names:
        names += [sp + "_" + k for k in (
            "ke", "ke_x", "ke_y", "px", "py", "xlo_flux", "xhi_flux", "ylo_flux", "yhi_flux", "vxlo_flux", "vxhi_flux",
            "vylo_flux", "vyhi_flux", "ke_e_dot", "integrated_ke_e_dot", "driver_time_envel")]
    return names


def poisson_plot_names(plot_ke_vel_bdy_flux, species_names):
    """Poisson::buildPlotNames (Poisson.C:1048-1073)"""
    names = ["EX", "EY"]
    if plot_ke_vel_bdy_flux:
        for sp in species_names:
            names += ["%s ke flux %s" % (sp, k) for k in ("vx lo", "vx hi", "vy lo", "vy hi")]
    return names


def maxwell_plot_names(plot_ke_vel_bdy_flux, species_names):
    """Maxwell::buildPlotNames (Maxwell.C:1124-1157): the six fields, then per species its transverse drift velocity"""
    names = ["EX", "EY", "EZ", "BX", "BY", "BZ"]
    for sp in species_names:
        names.append("%s VZ" % sp)
        if plot_ke_vel_bdy_flux:
            names += ["%s ke flux %s" % (sp, k) for k in ("vx lo", "vx hi", "vy lo", "vy hi")]
    return names


MAXWELL_FIELD_HISTORIES = ["E_max", "norm E", "Ex_max", "Ey_max", "Ez_max", "E_tot", "B_max", "norm B", "Bx_max", "By_max",
                           "Bz_max", "B_tot"]                     # Maxwell::buildTimeHistoryNames (Maxwell.C:976-987)


def write_time_histories(file_name_base, saved_save, names, sequences, time_seq, saved_seq, num_probes,
                         num_tracking_particles=0):
    """EMSolverBase::writeTimeHistories (EMSolverBase.C:510-537): <base>_<saved_save>.hdf holding the first saved_seq
    entries of every sequence"""
    w = TimeHistWriter("%s_%d.hdf" % (file_name_base, saved_save), time_seq, saved_seq, num_probes, num_tracking_particles)
    for name, seq in zip(names, sequences):
        w.write_time_history(name, seq, saved_seq)
    w.close()
    return w.name


def put_problem_domain(w, num_cells, x_lo, x_hi, dx, periodic):
    """ProblemDomain::putToDatabase (ProblemDomain.C:174-199)"""
    w.write_integer_array("N", num_cells)
    w.write_double_array("x_lo", x_lo)
    w.write_double_array("x_hi", x_hi)
    w.write_double_array("dx", dx)
    w.write_integer_value("isPeriodic_0", int(periodic[0]))
    w.write_integer_value("isPeriodic_1", int(periodic[1]))


def put_krook_layer(w, x_lo, x_hi, krook=None):
    """KrookLayer::putToDatabase (KrookLayer.C:249-282); krook = None is a species without a layer: the constructor's
    defaults, the whole configuration-space domain and no layer in any direction (KrookLayer.C:24-38)"""
    k = krook or {}
    w.write_double_array("x_lo_krook", k.get("x_lo", x_lo[:2]))
    w.write_double_array("x_hi_krook", k.get("x_hi", x_hi[:2]))
    w.write_double_value("krookPower", k.get("power", 3.0))
    w.write_double_value("krookCoeff", k.get("coefficient", 1.0))
    has_lo, has_hi = k.get("has_lo", [0, 0]), k.get("has_hi", [0, 0])
    w.write_integer_value("krookHasLayer", int(any(has_lo) or any(has_hi)))
    w.write_integer_value("krookHasLayer_lo_0", int(has_lo[0]))
    w.write_integer_value("krookHasLayer_lo_1", int(has_lo[1]))
    w.write_integer_value("krookHasLayer_hi_0", int(has_hi[0]))
    w.write_integer_value("krookHasLayer_hi_1", int(has_hi[1]))


def krook_state(deck_krook, x_lo, x_hi):
    """KrookLayer::parseParameters (KrookLayer.C:163-189) on a deck's krook.* keys {x1a, x1b, x2a, x2b, power, coefficient}:
    the layer state put_krook_layer records; None for a species without the keys"""
    if not deck_krook:
        return None
    lo, hi = [float(x_lo[0]), float(x_lo[1])], [float(x_hi[0]), float(x_hi[1])]
    has_lo, has_hi = [0, 0], [0, 0]
    for d in range(2):
        if ("x%da" % (d + 1)) in deck_krook:
            lo[d], has_lo[d] = float(deck_krook["x%da" % (d + 1)]), 1
        if ("x%db" % (d + 1)) in deck_krook:
            hi[d], has_hi[d] = float(deck_krook["x%db" % (d + 1)]), 1
    return dict(x_lo=lo, x_hi=hi, power=float(deck_krook.get("power", 3.0)), coefficient=float(deck_krook.get("coefficient", 1.0)),
                has_lo=has_lo, has_hi=has_hi)


def put_external_dist_krook(w, x_lo, x_hi):
    """ExternalDistKrookLayer::putToDatabase (ExternalDistKrook.C:308-350) of a species without such a layer
    (constructor defaults, ExternalDistKrook.C:22-42)"""
    w.write_double_array("x_lo_lo_external_dist_krook", x_lo[:2])
    w.write_double_array("x_hi_lo_external_dist_krook", x_lo[:2])
    w.write_double_array("x_lo_hi_external_dist_krook", x_hi[:2])
    w.write_double_array("x_hi_hi_external_dist_krook", x_hi[:2])
    w.write_double_value("externalDistKrookCoeff", 1.0)
    w.write_integer_value("externalDistKrookHasLayer", 0)
    for name in ("lo_0", "lo_1", "hi_0", "hi_1"):
        w.write_integer_value("externalDistKrookHasLayer_" + name, 0)


def put_species(w, index, sp, domain, tiles, info, integrated_e_dot_j=None, krook=None):
    """KineticSpecies::putToRestart (KineticSpecies.C:943-1020).  sp: name / mass / charge / bz_const and, for a driven
    species, driver_state = (num_phase_evals, phase, phase_h); tiles: every rank's dataBox of the distribution."""
    if sp.get("driver_state") is not None:
        # ShapedRampedCosineDriver::putToDatabase (ShapedRampedCosineDriver.C:202-217): its own group beside the species'
        n, phase, phase_h = sp["driver_state"]
        w.push_sub_dir("%s%d_%d" % (DRIVER_CLASS_NAME, index + 1, 1))
        w.write_integer_value("num_phase_evals", n)
        w.write_double_value("phase", phase)
        w.write_double_value("phase_h", phase_h)
        w.pop_sub_dir()
    w.push_sub_dir(sp["name"])
    w.write_integer_value("pdim", 4)
    w.write_integer_value("cdim", 2)
    w.write_double_value("mass", sp["mass"])
    w.write_double_value("charge", sp["charge"])
    w.write_double_value("bz_const", sp.get("bz_const", 0.0))
    put_problem_domain(w, *domain)
    w.write_parallel_array("distribution", tiles, info)
    if integrated_e_dot_j is not None:
        w.write_bulk_double_value("integrated_e_dot_j", integrated_e_dot_j)
    put_krook_layer(w, domain[1], domain[2], krook)
    put_external_dist_krook(w, domain[1], domain[2])
    w.pop_sub_dir()


def write_vp_restart(write_dir, index, species, n_ghosts, time, dt, cfl, tf, bz_const=0.0, plot_ke_vel_bdy_flux=False,
                     num_procs=1, max_files=16):
    """RestartManager::write (RestartManager.C:78-117) for a Vlasov-Poisson run.  The items write in the order they
    registered with the manager: VPSystem (VPSystem.C:76, :733-760), each KineticSpecies (KineticSpecies.C:238), Poisson
    (Poisson.C:73, :909-919), Simulation (Simulation.C:243, :520-543).  `species`: list of dicts with the put_species
    arguments (sp, domain, tiles, info, integrated_e_dot_j, krook)."""
    name = os.path.join(write_dir, "dist_%d.hdf" % index)                # RestartManager.H:182-188
    w = RestartWriter(name, max_files, num_procs)
    w.write_integer_value("species_list_size", len(species))
    w.write_double_value("bz_const", bz_const)
    w.write_integer_value("plot_ke_vel_bdy_flux", int(plot_ke_vel_bdy_flux))
    w.push_sub_dir("species_list")
    for s, item in enumerate(species):
        w.write_string("species.%d" % (s + 1), item["sp"]["name"])
    w.pop_sub_dir()
    for s, item in enumerate(species):
        put_species(w, s, item["sp"], item["domain"], item["tiles"], item["info"], item.get("integrated_e_dot_j"),
                    item.get("krook"))
    w.write_integer_value("nGhost", n_ghosts)                              # EMSolverBase::putToRestartCommon
    w.write_integer_value("isMaxwell", 0)
    w.write_integer_value("generating processes", num_procs)
    for key, v in zip(("major version", "minor version", "patch level"), LOKI_VERSION):
        w.write_integer_value(key, v)
    w.write_double_value("time", time)
    w.write_double_value("time step", dt)
    w.write_double_value("CFL", cfl)
    w.write_double_value("tf", tf)
    w.close()
    return name


def write_vm_restart(write_dir, index, species, n_ghosts, em_vars, vz, field_info, time, dt, cfl, tf, bz_const=0.0,
                     plot_ke_vel_bdy_flux=False, max_files=16):
    """RestartManager::write for a Vlasov-Maxwell run on one process: VMSystem (VMSystem.C:757-785), each KineticSpecies,
    Maxwell (Maxwell.C:942-964: nGhost, isMaxwell = 1, group "Maxwell" with EMVars (6, n2d, n1d) and vz<k> (n2d, n1d)),
    Simulation.  field_info = (distribInfo of EMVars, distribInfo of a vz array)."""
    name = os.path.join(write_dir, "dist_%d.hdf" % index)
    w = RestartWriter(name, max_files, 1)
    w.write_integer_value("species_list_size", len(species))
    w.write_double_value("bz_const", bz_const)
    w.write_integer_value("plot_ke_vel_bdy_flux", int(plot_ke_vel_bdy_flux))
    w.push_sub_dir("species_list")
    for s, item in enumerate(species):
        w.write_string("species.%d" % (s + 1), item["sp"]["name"])
    w.pop_sub_dir()
    for s, item in enumerate(species):
        put_species(w, s, item["sp"], item["domain"], item["tiles"], item["info"], item.get("integrated_e_dot_j"),
                    item.get("krook"))
    w.write_integer_value("nGhost", n_ghosts)
    w.write_integer_value("isMaxwell", 1)
    w.push_sub_dir("Maxwell")
    w.write_parallel_array("EMVars", {0: em_vars}, field_info[0])
    for k, v in enumerate(vz):
        w.write_parallel_array("vz%d" % k, {0: v}, field_info[1])
    w.pop_sub_dir()
    w.write_integer_value("generating processes", 1)
    for key, v in zip(("major version", "minor version", "patch level"), LOKI_VERSION):
        w.write_integer_value(key, v)
    w.write_double_value("time", time)
    w.write_double_value("time step", dt)
    w.write_double_value("CFL", cfl)
    w.write_double_value("tf", tf)
    w.close()
    return name


def read_vp_restart(name, max_files=16):
    """the inverse of write_vp_restart for a one-process dump: time, dt and per species the distribution (dataBox) and
    the integrated driver work (VPSystem / KineticSpecies / Simulation::getFromRestart)"""
    r = RestartReader(name, max_files)
    out = dict(time=r.read_double_value("time"), dt=r.read_double_value("time step"), cfl=r.read_double_value("CFL"),
               tf=r.read_double_value("tf"), n_ghosts=r.read_integer_value("nGhost"),
               num_procs=r.read_integer_value("generating processes"), species=[])
    n = r.read_integer_value("species_list_size")
    r.push_sub_dir("species_list")
    names = [r.read_string("species.%d" % (s + 1)) for s in range(n)]
    r.pop_sub_dir()
    for s, sp_name in enumerate(names):
        r.push_sub_dir(sp_name)
        item = dict(name=sp_name, mass=r.read_double_value("mass"), charge=r.read_double_value("charge"),
                    N=r.read_integer_array("N"), x_lo=r.read_double_array("x_lo"), x_hi=r.read_double_array("x_hi"))
        item["info"], item["distribution"] = r.read_parallel_array("distribution", 0)
        try:
            item["integrated_e_dot_j"] = r.read_bulk_double_value("integrated_e_dot_j", 0)
        except KeyError:
            item["integrated_e_dot_j"] = None
        r.pop_sub_dir()
        out["species"].append(item)
    if r.read_integer_value("isMaxwell"):
        r.push_sub_dir("Maxwell")
        out["em_vars"] = r.read_parallel_array("EMVars", 0)[1]
        out["vz"] = [r.read_parallel_array("vz%d" % k, 0)[1] for k in range(n)]
        r.pop_sub_dir()
    return out
