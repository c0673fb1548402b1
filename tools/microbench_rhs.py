"""Micro-benchmark of the fused Vlasov stage kernel alone (development tool; bench.py is the contract).
usage: python tools/microbench_rhs.py [nx ny nvx nvy] [--order 4] [--variant 0] [--reps 5] [--strict]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import loki_b200 as lkm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("n", nargs="*", type=int, default=[256, 256, 128, 128])
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--strict", action="store_true")
    ap.add_argument("--mode", default="stage", choices=["stage", "rhs"])
    ap.add_argument("--split", type=int, default=0, help="two launches: the tiles on the faces of these cut directions (1 x, 2 y, 3 both), then the rest")
    ap.add_argument("--fold", action="store_true", help="hand the inflow over (lk_rk_update.accel_bcs): velocity-boundary fill inside the stage")
    a = ap.parse_args()
    lk = lkm.load()
    lk.lk_set_strict(int(a.strict))
    lk.lk_set_rhs_variant(a.variant)
    n = a.n
    ng = 2 if a.order == 4 else 3
    nd = [k + 2 * ng for k in n]
    g = lkm.Geom.make(n, a.order, (0.07, 0.07, 0.1, 0.1))
    dev = torch.device("cuda:0")
    vol = nd[0] * nd[1] * nd[2] * nd[3]
    cells = n[0] * n[1] * n[2] * n[3]
    # smooth Maxwellian with a spatial perturbation, built on the device
    v3 = (torch.arange(nd[2], device=dev, dtype=torch.float64) - ng + 0.5) * 0.1 - 0.05 * n[2]
    v4 = (torch.arange(nd[3], device=dev, dtype=torch.float64) - ng + 0.5) * 0.1 - 0.05 * n[3]
    x = torch.arange(nd[0], device=dev, dtype=torch.float64)
    y = torch.arange(nd[1], device=dev, dtype=torch.float64)
    fv = torch.exp(-0.5 * (v4[:, None] ** 2 + v3[None, :] ** 2)) / (2 * np.pi)
    fx = 1.0 + 0.1 * torch.cos(0.05 * x)[None, :] * torch.cos(0.03 * y)[:, None]
    f = (fv[:, :, None, None] * fx[None, None, :, :]).contiguous()
    assert f.numel() == vol
    f_old = f.clone()
    delta = torch.zeros_like(f)
    pred = torch.zeros_like(f)
    vel = torch.stack([v3[None, :].expand(nd[3], nd[2]), v4[:, None].expand(nd[3], nd[2])]).contiguous()
    vxf = torch.zeros(2, nd[3], nd[2] + 1, device=dev, dtype=torch.float64)
    vxf[1] = v4[:, None]
    vyf = torch.zeros(2, nd[3] + 1, nd[2], device=dev, dtype=torch.float64)
    vyf[0] = v3[None, :]
    accel = 0.01 * torch.randn(2, nd[1], nd[0], device=dev, dtype=torch.float64)
    A = lkm.Accel()
    A.kind, A.field, A.vz = 0, accel.data_ptr(), None
    A.vxface_velocities, A.vyface_velocities = vxf.data_ptr(), vyf.data_ptr()
    A.normalization, A.bz_const = -1.0, 0.0
    U = lkm.RkUpdate()
    U.f_old, U.delta_in, U.delta_out, U.pred = f_old.data_ptr(), delta.data_ptr(), delta.data_ptr(), pred.data_ptr()
    U.w_delta, U.c_pred, U.use_delta = 1e-3, 5e-4, 0
    st = torch.cuda.current_stream().cuda_stream
    if a.fold:
        I = lkm.Inflow()
        fxc, fvc = fx.contiguous(), fv.contiguous()
        I.kind, I.fx, I.fv, I.fnorm, I.frac = 1, fxc.data_ptr(), fvc.data_ptr(), 1.0, 1.0
        U.accel_bcs, U.inflow_preset = C.addressof(I), 1
        assert lk.lk_preset_inflow_ghosts_4d(f.data_ptr(), C.byref(g), C.byref(I), st) == 0

    def run():
        if a.split:
            for ts in (1, 2):
                U.tile_set, U.cut_dirs = ts, a.split
                s = lk.lk_vlasov_stage(None, f.data_ptr(), C.byref(g), vel.data_ptr(), C.byref(A), C.byref(U), None, st)
                assert s == 0, lk.lk_last_error()
            return
        if a.mode == "stage" and a.fold:
            s = lk.lk_vlasov_stage(None, f.data_ptr(), C.byref(g), vel.data_ptr(), C.byref(A), C.byref(U), None, st)
        elif a.mode == "stage":
            s = lk.lk_vlasov_rhs(None, f.data_ptr(), C.byref(g), vel.data_ptr(), C.byref(A), C.byref(U), st)
        else:
            s = lk.lk_vlasov_rhs(pred.data_ptr(), f.data_ptr(), C.byref(g), vel.data_ptr(), C.byref(A), None, st)
        assert s == 0, lk.lk_last_error()

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.reps)]
    for e0, e1 in evs:
        e0.record()
        run()
        e1.record()
    torch.cuda.synchronize()
    times = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    ms = sum(times) / len(times)
    bpc = 40 if a.mode == "stage" else 16
    print("n=%s order=%d variant=%d strict=%d mode=%s fold=%d split=%d: %.3f ms/launch, %.2f Gcell/s, %.1f GB/s algorithmic (%d B/cell)  [min %.3f median %.3f ms, %s]" % (
        n, a.order, a.variant, int(a.strict), a.mode, int(a.fold), a.split, ms, cells / ms / 1e6, cells * bpc / ms / 1e6, bpc, times[0],
        times[len(times) // 2], os.path.basename(os.environ.get("LOKI_B200_LIB", "libloki_b200.so"))))


if __name__ == "__main__":
    main()
