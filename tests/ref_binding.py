"""ctypes binding of oracle/_ref/libloki_ref.so: the reference's own Fortran kernels, transliterated to C by
oracle/f77toc.py (no Fortran compiler in the image) and called with the reference's own Fortran calling
convention (every argument by reference, 8 box integers).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

from oracle_binding import IC_FN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libloki_ref.so")


def available():
    return os.path.exists(SO)


class Ref:
    def __init__(self):
        self.L = C.CDLL(SO)
        self.L.loki_ref_set_ic.argtypes = [IC_FN, C.c_void_p, C.POINTER(C.c_int * 4)]

    @staticmethod
    def _i(v):
        return C.byref(C.c_int(int(v)))

    @staticmethod
    def _d(v):
        return C.byref(C.c_double(float(v)))

    @staticmethod
    def _p(a):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)

    def boxes(self, s, lo=None):
        """BOX4D_TO_FORT(dataBox), BOX4D_TO_FORT(interiorBox) with a global offset `lo` (interior lower)"""
        lo = lo or (0, 0, 0, 0)
        ng = s.ng
        data, inter = [], []
        for k in range(4):
            data += [lo[k] - ng, lo[k] + s.n[k] - 1 + ng]
            inter += [lo[k], lo[k] + s.n[k] - 1]
        return [self._i(v) for v in data], [self._i(v) for v in inter], data, inter

    def weno43(self, u, vel):
        face = C.c_double()
        a = [C.c_double(x) for x in u] + [face, C.c_double(vel)]
        self.L.weno43fit4d_(*[C.byref(x) for x in a])
        return face.value

    def weno65(self, u, vel):
        face = C.c_double()
        a = [C.c_double(x) for x in u] + [face, C.c_double(vel)]
        self.L.weno65fit4d_(*[C.byref(x) for x in a])
        return face.value
