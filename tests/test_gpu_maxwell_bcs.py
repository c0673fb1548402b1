"""The boundary routines of MaxwellF.f (zeroghost2d, maxwelladdantennasource, maxwellsetembcs, maxwellsetvzbcs) as Level-0
kernels against the oracle, itself pinned bit for bit to the transliterated Fortran (test_oracle_pin.py): a box in every
position of a 3 x 3 decomposition, every periodicity, both orders.  Bit for bit."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", [4, 6])
def test_maxwell_boundary_kernels_equal_the_oracle_bits(lk, ok, order):
    import torch
    ng = 2 if order == 4 else 3
    n1, n2 = 9, 7
    n1d, n2d = n1 + 2 * ng, n2 + 2 * ng
    rng = np.random.default_rng(60 + order)
    nx, ny = 3 * n1, 3 * n2
    changed = 0
    for bx in range(3):
        for by in range(3):
            lo1, lo2 = bx * n1, by * n2
            at = (C.c_int * 4)(int(lo1 == 0), int(lo1 + n1 == nx), int(lo2 == 0), int(lo2 + n2 == ny))
            for xper, yper in ((0, 0), (1, 0), (0, 1), (1, 1)):
                em = rng.uniform(-1, 1, size=(6, n2d, n1d))
                want = em.copy()
                ok.ok_maxwell_set_em_bcs(want.ravel(), n1, n2, order, at, xper, yper, 22.36)
                d = torch.from_numpy(em).cuda()
                assert lk.lk_maxwell_set_em_bcs(d.data_ptr(), n1, n2, order, C.byref(at), xper, yper, 22.36, None) == 0
                got = d.cpu().numpy()
                assert np.array_equal(got, want)
                changed += int(not np.array_equal(got, em))
                vz = rng.uniform(-1, 1, size=(n2d, n1d))
                vwant = vz.copy()
                ok.ok_maxwell_set_vz_bcs(vwant.ravel(), n1, n2, order, at, xper, yper)
                dv = torch.from_numpy(vz).cuda()
                assert lk.lk_maxwell_set_vz_bcs(dv.data_ptr(), n1, n2, order, C.byref(at), xper, yper, None) == 0
                assert np.array_equal(dv.cpu().numpy(), vwant)
    assert changed == 20                 # edge boxes under the periodicities that leave their boundary open, corner boxes under three
    u = rng.uniform(-1, 1, size=(6, n2d, n1d))
    want = u.copy()
    ok.ok_zero_ghost_2d(want.ravel(), n1, n2, ng, 6)
    d = torch.from_numpy(u).cuda()
    assert lk.lk_zero_ghost_2d(d.data_ptr(), n1, n2, ng, 6, None) == 0
    assert np.array_equal(d.cpu().numpy(), want)
    src = rng.uniform(-1, 1, size=(6, n2d, n1d))
    want = u.copy()
    ok.ok_maxwell_add_antenna_source(want.ravel(), src.ravel(), n1, n2, ng)
    d, ds = torch.from_numpy(u).cuda(), torch.from_numpy(src).cuda()
    assert lk.lk_maxwell_add_antenna_source(d.data_ptr(), ds.data_ptr(), n1, n2, ng, None) == 0
    assert np.array_equal(d.cpu().numpy(), want)
    at = (C.c_int * 4)(1, 1, 1, 1)
    assert lk.lk_maxwell_set_em_bcs(d.data_ptr(), n1, n2, 5, C.byref(at), 0, 0, 22.36, None) != 0
    assert lk.lk_maxwell_set_em_bcs(d.data_ptr(), n1, n2, order, C.byref(at), 0, 0, 0.0, None) != 0
