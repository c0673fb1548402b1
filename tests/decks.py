"""Deck set-ups for the system-level parity tests: loki_b200.decks (the product-side deck mirror) plus
the oracle-side adapter (species descriptors with the IC as a point callback, the way the reference's
Fortran calls back into ICInterface.C:36-57).  Test infrastructure."""
import ctypes as C  # noqa: F401

from loki_b200 import decks as _d
from loki_b200.decks import Species, driver_params, PI  # noqa: F401
import numpy as np

from oracle_binding import OkGeom, OkSpecies, OkIcTables, IC_FN, load as _load_oracle


class Deck(_d.Deck):
    # ---- oracle side ----
    @staticmethod
    def _ok_ic_fn():
        return _load_oracle().ok_ic_from_tables

    def oracle_species(self, keep):
        arr = (OkSpecies * len(self.species))()
        for k, sp in enumerate(self.species):
            n, dx = self.geom_of(sp)
            arr[k].g = OkGeom.make(n, self.order, dx)
            arr[k].mass, arr[k].charge, arr[k].bz_const = sp.mass, sp.charge, sp.bz
            arr[k].vlo[0], arr[k].vlo[1] = sp.vlim[0], sp.vlim[2]
            arr[k].vhi[0], arr[k].vhi[1] = sp.vlim[1], sp.vlim[3]
            # the IC as the C point callback of the oracle over the IC classes' cached tables (same products in the
            # same order as the classes' getIC_At_Pt; a Python callback per ghost cell would dominate the run time)
            t = OkIcTables()
            nd = arr[k].g.nd
            t.n1d, t.n2d, t.n3d, t.n4d = nd
            tabs = []
            if sp.stream is not None:
                fx, fx2, fv, fv2 = self.stream_tables(sp)
                t.kind = {2: 2, 4: 4}.get(self.inflow_kind(sp), 0)
                tabs = [fx, fv, fx2, fv2]
                t.fx, t.fv = fx.ctypes.data, fv.ctypes.data
                if fx2 is not None:
                    t.fx2 = fx2.ctypes.data
                if fv2 is not None:
                    t.fv2 = fv2.ctypes.data
            else:
                fx, fv, fnorm = self.ic_tables(sp)
                if sp.factorable:
                    t.kind, t.fnorm, t.frac = 1, fnorm, sp.frac
                    tabs = [fx, fv]
                    t.fx, t.fv = fx.ctypes.data, fv.ctypes.data
                else:
                    fic = np.ascontiguousarray(self.initial_state_full(sp, fx, fnorm))  # the cached m_f (PerturbedMaxwellianIC.C:176-246)
                    t.kind = 3
                    tabs = [fic]
                    t.full = fic.ctypes.data
            keep.append((t, tabs))
            arr[k].ic = C.cast(self._ok_ic_fn(), IC_FN)
            arr[k].ic_ctx = C.addressof(t)
            arr[k].has_driver = 1 if sp.driver else 0
            if sp.driver:
                for j in range(16):
                    arr[k].driver[j] = sp.driver[j]
            arr[k].driver_phase = 0.0
            arr[k].driver_shape_type = 0
        return arr


class VMDeck(_d.VMDeck, Deck):
    pass


def _wrap(d):
    d.__class__ = VMDeck if isinstance(d, _d.VMDeck) else Deck
    return d


def interpenetrating_streams(*a, **k):
    return _wrap(_d.interpenetrating_streams(*a, **k))


def em_damping(*a, **k):
    return _wrap(_d.em_damping(*a, **k))


def plane_epw(*a, **k):
    return _wrap(_d.plane_epw(*a, **k))


def plane_iaw(*a, **k):
    return _wrap(_d.plane_iaw(*a, **k))


def pitch_angle_collisions(*a, **k):
    return _wrap(_d.pitch_angle_collisions(*a, **k))
