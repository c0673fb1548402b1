"""The per-cell arithmetic of the pitch-angle collision kernels (loki_b200/csrc/lk_coll.cuh), compiled for the host by
tests/hostcheck/coll_hostcheck.cpp and compared with the oracle: a check of the device functions' logic in the
container without a GPU.  The GPU tests (test_gpu_coll.py) run the real kernels through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_oracle_pin import _pitch_case

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "coll_hostcheck.cpp")
HDR = os.path.join(HERE, "..", "loki_b200", "csrc", "lk_coll.cuh")
SO = os.path.join(HERE, "hostcheck", "libcoll_hostcheck.so")


@pytest.fixture(scope="module")
def hc():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", SO, SRC])
    return C.CDLL(SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("order", [4, 6])
def test_device_functions_on_the_host_equal_the_oracle(ok, hc, order):
    s, iv, xlo, xhi, rlo, rhi = _pitch_case(ok, order)
    n = (C.c_int * 4)(*s.n)
    dx = np.array(s.dx)
    iv2 = np.zeros_like(iv)
    hc.hc_fields(_p(iv2[0]), _p(iv2[1]), _p(iv2[2]), _p(s.f), n, s.ng, _p(dx), _p(s.velocities))
    assert np.array_equal(iv, iv2)
    vlo, vhi = xlo[2:].copy(), xhi[2:].copy()
    for cons in (0, 1):
        r1 = np.zeros_like(s.f)
        r2 = np.zeros_like(s.f)
        ok.ok_append_pitch_angle_collision(r1.ravel(), s.f.ravel(), C.byref(s.g), s.velocities, iv[0].ravel(), iv[1].ravel(),
                                           iv[2].ravel(), vlo, vhi, rlo, rhi, 0.05, 0.37, cons)
        hc.hc_append(_p(r2), _p(s.f), n, s.ng, order, _p(dx), _p(s.velocities), _p(iv[0]), _p(iv[1]), _p(iv[2]), _p(vlo), _p(vhi),
                     _p(rlo), _p(rhi), C.c_double(0.05), C.c_double(0.37), cons)
        assert np.any(r1 != 0.0) == (cons == 1 or order == 4)
        assert np.array_equal(r1, r2), np.abs(r1 - r2).max()      # the oracle's operation order: the same bits
