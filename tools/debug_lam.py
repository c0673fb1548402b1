import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import loki_b200, oracle_binding, decks, torch
from loki_b200 import host
deck = decks.plane_epw(n=(12, 6), nv=(24, 12))
L = loki_b200.load(); H = host.lib(); ok = oracle_binding.load(); L.lk_set_strict(0)
keep = []; sp = deck.oracle_species(keep)
xlo = (C.c_double * 2)(deck.xlim[0], deck.xlim[2]); xhi = (C.c_double * 2)(deck.xlim[1], deck.xlim[3])
w = ok.ok_vp_work_create(1, sp, C.byref(xlo), C.byref(xhi))
f, fx, fv, fnorm = deck.initial_state(deck.species[0])
rng = np.random.default_rng(5); state = np.ascontiguousarray(f * (1.0 + 0.01 * rng.uniform(-1, 1, size=f.shape)))
P = lambda a: (C.c_void_p * len(a))(*[x.ctypes.data for x in a])
d = deck.product_desc(); sys_ = C.c_void_p(); assert H.lk_vp_create(C.byref(sys_), C.byref(d), None) == 0
H.lk_vp_set_state(sys_, 0, state.ctypes.data); H.lk_vp_set_inflow(sys_, 0, fx.ctypes.data, fv.ctypes.data, fnorm, 1.0)
rhs_d = torch.zeros(state.shape, dtype=torch.float64, device="cuda")
H.lk_vp_eval_rhs(sys_, (C.c_void_p * 1)(rhs_d.data_ptr()), 0.0)
f_old, f_new = state.copy(), np.zeros_like(state); rhs0 = np.zeros_like(state); ax, ay = np.zeros(1), np.zeros(1); ke = np.zeros(1)
ok.ok_vp_eval_rhs(w, P([rhs0]), P([f_old]), 0.0, np.zeros(1), ax, ay)
lam = (C.c_double * 2)(); H.lk_vp_lambda_max(sys_, 0, C.byref(lam)); print("init lam dev", lam[0], lam[1], "oracle", ax[0], ay[0])
t = 0.0
for step in range(3):
    dt = deck.cfl * ok.ok_vp_stable_dt(w, ax, ay, deck.rk)
    dtd = C.c_double(); H.lk_vp_stable_dt(sys_, C.byref(dtd)); print("step", step, "dt oracle", dt, "dev", dtd.value)
    ok.ok_vp_rk4_step(w, P([f_new]), P([f_old]), t, dt, ke)
    H.lk_vp_set_time(sys_, t); H.lk_vp_advance(sys_, dt); t += dt
    f_old, f_new = f_new, f_old
    tmp = f_old.copy(); ok.ok_vp_eval_rhs(w, P([rhs0]), P([tmp]), t, np.zeros(1), ax, ay)
    H.lk_vp_lambda_max(sys_, 0, C.byref(lam)); print("  lam dev (stage 4)", lam[0], lam[1], " oracle eval(new)", ax[0], ay[0])
