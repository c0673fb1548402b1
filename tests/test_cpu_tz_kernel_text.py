"""The two-species twilight-zone kernels without a GPU: the expression text of k_trig_tz (loki_b200/csrc/lk_bcs.cu: T1..T5, U1..U5,
`wave`, the exact solution) is read out of the source and evaluated here in IEEE doubles on libm-built table values --
which is what the device does, the file being compiled with -fmad=false -- and compared with the oracle bit for bit.
The GPU tests (test_gpu_tz.py) run the real kernels; this guards the source text in the container."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from util import Setup

SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "loki_b200", "csrc", "lk_bcs.cu")


def tz_powi2(x):
    return x * x


def tz_powi4(x):
    t = x * x
    return t * t


@pytest.mark.parametrize("species", [0, 1])
def test_two_species_kernel_expressions_equal_the_oracle(ok, species):
    src = open(SRC).read()

    def grab(name):
        return " ".join(re.search(r"const double %s = (.*?);" % name, src, re.S).group(1).split())
    waves = re.findall(r"const double wave = (.*?);", src)
    assert len(waves) == 2
    names = ("T1", "T2", "T3", "T4", "T5") if species == 0 else ("U1", "U2", "U3", "U4", "U5")
    exprs = [compile(grab(k), k, "eval") for k in names]
    wave_expr = compile(waves[species], "wave", "eval")
    s = Setup(ok, (5, 4, 6, 5), 4, bz=0.0)
    n1d, n2d, n3d, n4d = s.nd
    lo = (C.c_int * 2)(1, -4)
    xlo, dx = np.array([-2 * np.pi, -1.5]), np.array(s.dx)
    vel = s.velocities.reshape(2, n4d, n3d)
    for time, dparams in ((0.37, [0.1, 1.0, 4.0]), (2.5, [0.7, 0.5, 25.0])):
        dp = np.array(dparams)
        base = np.zeros(s.f.shape)
        want = base.copy()
        ok.ok_set_two_species_trig_tz_source(want.ravel(), C.byref(s.g), lo, xlo, dx, time, s.velocities, dp, species)
        ewant = np.zeros_like(base)
        ok.ok_compute_two_species_trig_tz_source_error(ewant.ravel(), base.ravel(), C.byref(s.g), lo, xlo, dx, time, s.velocities, dp, species)
        env = dict(tz_powi2=tz_powi2, tz_powi4=tz_powi4)
        env["a"], env["m"] = float(dp[0]), float(dp[1] if species == 0 else dp[2])
        alpha = math.sqrt(env["m"])
        env["pi"] = 4.0 * math.atan(1.0)
        env.update(kxi=2.0, kyi=4.0, kti=1.0, kxe=4.0, kye=2.0, kte=1.0, kI2=20.0, kE2=20.0)
        env.update(ste=math.sin(time), cte=math.cos(time), sti=math.sin(time), cti=math.cos(time))
        env["al2"], env["al4"] = tz_powi2(alpha), tz_powi4(alpha)
        bad = 0
        for i4 in range(n4d):
            for i3 in range(n3d):
                vx, vy = float(vel[0, i4, i3]), float(vel[1, i4, i3])
                env.update(vx=vx, vy=vy, e=math.exp(-((alpha * alpha) * (vx * vx + vy * vy) / 0.2e1)))
                for i2 in range(n2d):
                    y = xlo[1] + ((lo[1] + i2) + 0.5) * dx[1]
                    env.update(sye=math.sin(2.0 * y), cye=math.cos(2.0 * y), syi=math.sin(4.0 * y), cyi=math.cos(4.0 * y))
                    for i1 in range(n1d):
                        x = xlo[0] + ((lo[0] + i1) + 0.5) * dx[0]
                        env.update(sxe=math.sin(4.0 * x), cxe=math.cos(4.0 * x), sxi=math.sin(2.0 * x), cxi=math.cos(2.0 * x))
                        env["wave"] = eval(wave_expr, env)
                        t = [eval(c, env) for c in exprs]
                        h = ((((t[0] - t[1]) - t[2]) - t[3]) - t[4]) if species == 0 else ((((t[0] - t[1]) - t[2]) + t[3]) + t[4])
                        fexact = (((env["al2"] / env["pi"]) * env["e"]) * env["wave"]) / 0.2e1
                        bad += int(h != want[i4, i3, i2, i1]) + int((0.0 - fexact) != ewant[i4, i3, i2, i1])
        assert bad == 0
    # the combination lines of the kernel are the ones evaluated above
    assert "((((T1 - T2) - T3) - T4) - T5)" in src and "((((U1 - U2) - U3) + U4) + U5)" in src
    assert "((((al2 / pi) * e) * wave) / 0.2e1)" in src
