"""Deck set-ups for the system-level parity tests: loki_b200.decks (the product-side deck mirror) plus
the oracle-side adapter (species descriptors with the IC as a point callback, the way the reference's
Fortran calls back into ICInterface.C:36-57).  Test infrastructure."""
import ctypes as C  # noqa: F401

from loki_b200 import decks as _d
from loki_b200.decks import Species, driver_params, PI  # noqa: F401
from oracle_binding import OkGeom, OkSpecies, IC_FN


class Deck(_d.Deck):
    # ---- oracle side ----
    def oracle_species(self, keep):
        arr = (OkSpecies * len(self.species))()
        for k, sp in enumerate(self.species):
            n, dx = self.geom_of(sp)
            arr[k].g = OkGeom.make(n, self.order, dx)
            arr[k].mass, arr[k].charge, arr[k].bz_const = sp.mass, sp.charge, sp.bz
            arr[k].vlo[0], arr[k].vlo[1] = sp.vlim[0], sp.vlim[2]
            arr[k].vhi[0], arr[k].vhi[1] = sp.vlim[1], sp.vlim[3]
            if sp.stream is not None:
                fx, fx2, fv, fv2 = self.stream_tables(sp)
                kind = self.inflow_kind(sp)
                if kind == 2:      # InterpenetratingStreamIC.C:275-278
                    def cb(ctx, i1, i2, i3, i4, fx=fx, fv=fv, fx2=fx2, fv2=fv2):
                        return fx[i2, i1] * fv[i4, i3] + fx2[i2, i1] * fv2[i4, i3]
                elif kind == 4:    # :279-281
                    def cb(ctx, i1, i2, i3, i4, fx=fx, fv=fv, fx2=fx2):
                        return fv[i4, i3] * fx[i2, i1] * fx2[i2, i1]
                else:
                    def cb(ctx, i1, i2, i3, i4, fx=fx, fv=fv):
                        return fv[i4, i3] * fx[i2, i1]
                fx = fv = None
            else:
                fx, fv, fnorm = self.ic_tables(sp)
            frac = sp.frac
            if sp.stream is not None:
                pass
            elif sp.factorable:
                def cb(ctx, i1, i2, i3, i4, fx=fx, fv=fv, fnorm=fnorm, frac=frac):
                    return fnorm * fv[i4, i3] * fx[i2, i1] * frac
            else:
                fic = self.initial_state_full(sp, fx, fnorm)  # the cached m_f (PerturbedMaxwellianIC.C:176-246)

                def cb(ctx, i1, i2, i3, i4, fic=fic):
                    return fic[i4, i3, i2, i1]
            fn = IC_FN(cb)
            keep.append(fn)
            arr[k].ic = fn
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
