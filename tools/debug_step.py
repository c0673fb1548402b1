"""debug aid: one production step of a deck vs the oracle, worst cell per species"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import loki_b200, oracle_binding, decks
from loki_b200 import host
from util import star_rel_err
which = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mk = [lambda: decks.plane_epw(n=(16, 8), nv=(32, 16)), lambda: decks.plane_iaw(n=(12, 10), nv=(16, 12)),
      lambda: decks.plane_iaw(n=(10, 10), nv=(16, 10), order=6, rk=6)][which]
deck = mk()
L = loki_b200.load(); H = host.lib(); ok = oracle_binding.load()
L.lk_set_strict(0)
keep = []
sp = deck.oracle_species(keep)
xlo = (C.c_double * 2)(deck.xlim[0], deck.xlim[2]); xhi = (C.c_double * 2)(deck.xlim[1], deck.xlim[3])
w = ok.ok_vp_work_create(len(deck.species), sp, C.byref(xlo), C.byref(xhi))
states, tables = [], []
for k, s in enumerate(deck.species):
    f, fx, fv, fnorm = deck.initial_state(s)
    rng = np.random.default_rng(20 + k)
    states.append(np.ascontiguousarray(f * (1.0 + 0.02 * rng.uniform(-1, 1, size=f.shape)))); tables.append((fx, fv, fnorm))
P = lambda a: (C.c_void_p * len(a))(*[x.ctypes.data for x in a])
f_old = [s.copy() for s in states]; f_new = [np.zeros_like(s) for s in states]; ke = np.zeros(len(states))
(ok.ok_vp_rk4_step if deck.rk == 4 else ok.ok_vp_rk6_step)(w, P(f_new), P(f_old), 0.25, 0.02, ke)
d = deck.product_desc(); sys_ = C.c_void_p()
assert H.lk_vp_create(C.byref(sys_), C.byref(d), None) == 0
for k, f in enumerate(states):
    H.lk_vp_set_state(sys_, k, f.ctypes.data); fx, fv, fnorm = tables[k]
    H.lk_vp_set_inflow(sys_, k, fx.ctypes.data, fv.ctypes.data, fnorm, deck.species[k].frac)
H.lk_vp_set_time(sys_, 0.25)
assert H.lk_vp_advance(sys_, 0.02) == 0
ng = deck.ng; I = (slice(ng, -ng),) * 4
for k in range(len(states)):
    out = np.empty_like(states[k]); H.lk_vp_get_state(sys_, k, out.ctypes.data)
    a = np.abs(states[k]); m = a.copy()
    for ax in range(4):
        for j in range(1, ng + 1):
            m = np.maximum(m, np.roll(a, j, axis=ax)); m = np.maximum(m, np.roll(a, -j, axis=ax))
    r = np.abs(out[I] - f_new[k][I]) / m[I]
    idx = np.unravel_index(np.argmax(r), r.shape)
    print("species", k, "worst", r.max(), "at (i4,i3,i2,i1)", idx, "out", out[I][idx], "ref", f_new[k][I][idx], "old", states[k][I][idx], "starmax", m[I][idx],
          "n_bad(>1e-13)", int((r > 1e-13).sum()), "of", r.size)
    v = C.c_double(); H.lk_vp_ke_e_dot(sys_, k, C.byref(v)); print("   ke", v.value, ke[k])
