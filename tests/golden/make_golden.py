"""Generates tests/golden/kinetic_f77_golden.npz and tests/golden/f77abi_golden.npz from oracle/_ref/libloki_ref.so, i.e. from the
reference's own Fortran kernels (KineticSpeciesF.f) transliterated by oracle/f77toc.py and compiled with
gcc -O2 -ffp-contract=off.  Run in the build container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py
The inputs are seeded (tests/util.py:Setup) so the oracle and the GPU tests can regenerate them."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_binding  # noqa: E402
import ref_binding  # noqa: E402
from util import Setup  # noqa: E402


def main():
    ok = oracle_binding.load()       # only used to build the seeded inputs (velocity tables)
    R = ref_binding.Ref()
    out = {}
    rng = np.random.default_rng(2024)
    for order in (4, 6):
        count = 2000
        u = rng.uniform(-1, 1, size=(count, order))
        u[::3] = np.exp(-(rng.uniform(0, 6, size=(len(u[::3]), 1)) + 0.05 * np.arange(order)[None, :]) ** 2) * 0.16
        u[::5] = 0.25
        u[1::11] = np.where(np.arange(order)[None, :] < 2, 1.0, 0.0)
        vel = rng.uniform(-1, 1, size=count)
        vel[::7] = 0.0
        face = np.array([(R.weno43 if order == 4 else R.weno65)(u[k], vel[k]) for k in range(count)])
        out["weno%d_u" % order], out["weno%d_vel" % order], out["weno%d_face" % order] = u, vel, face
        n = (9, 6, 10, 7) if order == 4 else (7, 7, 8, 9)
        seed = 77 + order
        s = Setup(ok, n, order, bz=0.3, seed=seed)
        db, ib, data, inter = R.boxes(s)
        n1d, n2d, n3d, n4d = s.nd
        vel3 = np.zeros((n3d + 1) * n4d * n1d * n2d)
        vel4 = np.zeros((n4d + 1) * n1d * n2d * n3d)
        ax, ay = C.c_double(), C.c_double()
        a2 = [R._i(data[0]), R._i(data[1]), R._i(data[2]), R._i(data[3])]
        R.L.setphasespacevel4d_(R._p(vel3), R._p(vel4), *db, *ib, R._p(s.vxface), R._p(s.vyface), R._d(s.norm),
                                R._d(s.bz), R._p(s.accel), *a2, C.byref(ax), C.byref(ay))
        rhs = np.zeros_like(s.f)
        dxs = np.array(s.dx)
        R.L.computeadvectionderivatives4d_(R._p(rhs), R._p(s.f), *db, *ib, R._p(s.vel1), R._p(s.vel2), R._p(dxs), R._i(order))
        R.L.computeaccelerationderivatives4d_(R._p(rhs), R._p(s.f), *db, *ib, R._p(vel3), R._p(vel4), R._p(dxs), R._i(order))
        bc = s.f.copy()
        cb = s.ic_callback(0.7, 0.9)
        lower = (C.c_int * 4)(*[data[2 * k] for k in range(4)])
        R.L.loki_ref_set_ic(cb, None, C.byref(lower))
        R.L.setaccelerationbcs4d_(R._p(bc), *db, *db, *ib, R._i(order), R._p(vel3), R._p(vel4), C.byref(C.c_int64(0)))
        out["rhs%d_n" % order] = np.array(n)
        out["rhs%d_seed" % order] = np.array(seed)
        out["rhs%d_f" % order] = s.f
        out["rhs%d_rhs" % order] = rhs
        out["rhs%d_bc" % order] = bc
        out["rhs%d_amax" % order] = np.array([ax.value, ay.value])
        xlo = np.zeros(4)
        r5 = [C.c_double(0.0) for _ in range(5)]
        R.L.computeke_(*db, *ib, R._p(xlo), R._p(xlo), R._p(dxs), R._p(s.f), R._d(1.7), R._p(s.velocities), *[C.byref(v) for v in r5])
        out["rhs%d_ke" % order] = np.array([v.value for v in r5])
        r3 = [C.c_double(0.0) for _ in range(3)]
        R.L.computekemaxwell_(*db, *ib, R._p(xlo), R._p(xlo), R._p(dxs), R._p(s.f), R._d(1.7), R._p(s.velocities), R._p(s.vz),
                              *[C.byref(v) for v in r3])
        out["rhs%d_kem" % order] = np.array([v.value for v in r3])
    path = os.path.join(HERE, "kinetic_f77_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    # every routine of the Fortran ABI the CUDA library re-exports (include/loki_b200_f77.h), through the argument
    # lists of tests/f77_cases.py: the GPU tests replay the same calls on device arrays and compare bits
    from f77_cases import HostBackend, kinetic_cases, field_cases, collision_cases
    B = HostBackend(R.L, R.L.loki_ref_set_ic)
    out2 = {}
    for order in (4, 6):
        for k, v in kinetic_cases(B, ok, order).items():
            out2["k%d_%s" % (order, k)] = v
        for k, v in field_cases(B, order).items():
            out2["f%d_%s" % (order, k)] = v
        for k, v in collision_cases(B, ok, order).items():
            out2["c%d_%s" % (order, k)] = v
    path = os.path.join(HERE, "f77abi_golden.npz")
    np.savez_compressed(path, **out2)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()


def external2d_profiles():
    """tests/golden/External2D_profiles.npz: the "2D dist" datasets of the reference's test/External2D/rho_init_*.h5 (read
    with loki_b200/h5lite.py, which test_cpu_outputs.py pins against those very files when the reference tree is present);
    the tests write them back out as HDF5 files for the "External 2D" initial condition"""
    from loki_b200 import h5lite
    ref = "/root/reference/test/External2D"
    out = {n: np.array(h5lite.read(os.path.join(ref, "rho_init_%s.h5" % n))["2D dist"].data) for n in ("e", "He", "C")}
    path = os.path.join(HERE, "External2D_profiles.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__" and os.path.isdir("/root/reference/test/External2D"):
    external2d_profiles()
