"""Flux-form diagnostics (loki_b200/csrc/lk_flux.cu) against the oracle, which is pinned to the
transliterated reference Fortran (tests/test_oracle_pin.py): face fits, fluxes, their divergence and the velocity-space
flux field bit for bit in BOTH arithmetic modes; the eight boundary kinetic-energy fluxes (tree sums) to 1e-12 of the
largest -- from materialised flux arrays (computekeflux_) and from the product path that never builds them."""
import ctypes as C

import numpy as np
import pytest

from util import Setup, Dev

pytestmark = pytest.mark.gpu
CASES = [((7, 6, 9, 8), 4), ((6, 7, 8, 7), 6), ((12, 5, 10, 16), 4)]


def chk(lk, status, what):
    assert status == 0, "%s: %s" % (what, lk.lk_last_error().decode())


def _oracle_fluxes(ok, s, u, vels):
    face = [np.zeros_like(v) for v in vels]
    flux = [np.zeros_like(v) for v in vels]
    for d in range(4):
        ok.ok_face_fluxes_4d(flux[d], face[d], u.ravel(), C.byref(s.g), vels[d], d)
    return face, flux


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,order", CASES)
def test_face_fluxes_divergence_and_velocity_space_flux_bit_exact(lk, ok, n, order, mode):
    import torch
    s = Setup(ok, n, order, bz=0.3)
    d = Dev(lk, s)
    vel3, vel4, _, _ = s.vel34(ok)
    vels = [s.vel1, s.vel2, vel3, vel4]
    face, flux = _oracle_fluxes(ok, s, s.f, vels)
    old = lk.lk_set_strict(mode)
    try:
        dflux = []
        for k in range(4):
            dv, dface, dfl = d.t(vels[k]), torch.zeros(vels[k].size, dtype=torch.float64, device="cuda"), torch.zeros(
                vels[k].size, dtype=torch.float64, device="cuda")
            chk(lk, lk.lk_face_fluxes_4d(dfl.data_ptr(), dface.data_ptr(), d.f.data_ptr(), C.byref(d.g), dv.data_ptr(), k, None), "flux")
            assert np.array_equal(dface.cpu().numpy(), face[k]) and np.array_equal(dfl.cpu().numpy(), flux[k]), k
            assert np.any(flux[k] != 0)
            dflux.append(dfl)
        want = np.zeros_like(s.f)
        ok.ok_accum_flux_div_4d(want.ravel(), C.byref(s.g), *flux)
        got = torch.zeros_like(d.f)
        chk(lk, lk.lk_accum_flux_div_4d(got.data_ptr(), C.byref(d.g), *[x.data_ptr() for x in dflux], None), "div")
        assert np.array_equal(got.cpu().numpy(), want) and np.any(want != 0)
        n1d, n2d = s.nd[0], s.nd[1]
        for dr in (2, 3):
            for side in (0, 1):
                q = np.random.default_rng(dr * 2 + side).uniform(-1, 1, size=(n2d, n1d))
                dq = d.t(q)
                ok.ok_compute_ke_vel_space_flux(q.ravel(), C.byref(s.g), flux[2], flux[3], s.vxface, s.vyface, dr, side, 1.7)
                chk(lk, lk.lk_ke_vel_space_flux(dq.data_ptr(), C.byref(d.g), dflux[dr].data_ptr(), d.vxface.data_ptr(),
                                                d.vyface.data_ptr(), dr, side, 1.7, None), "kev")
                assert np.array_equal(dq.cpu().numpy(), q)
    finally:
        lk.lk_set_strict(old)


@pytest.mark.parametrize("maxwell", [False, True])
@pytest.mark.parametrize("n,order", CASES)
def test_boundary_kinetic_energy_fluxes(lk, ok, n, order, maxwell):
    """computekeflux_ on all eight boundaries: from flux arrays, and lk_ke_flux_boundaries straight from f with the
    acceleration evaluated on the fly (Vlasov-Poisson and Vlasov-Maxwell forms)"""
    import torch
    s = Setup(ok, n, order, bz=0.3)
    d = Dev(lk, s, maxwell=maxwell)
    vel3, vel4, _, _ = s.vel34(ok, maxwell)
    vels = [s.vel1, s.vel2, vel3, vel4]
    face, flux = _oracle_fluxes(ok, s, s.f, vels)
    want = np.array([ok.ok_compute_ke_flux(C.byref(s.g), *flux, s.velocities, s.vxface, s.vyface, dr, side, 1.7)
                     for dr in range(4) for side in range(2)])
    assert np.all(want != 0)
    out = torch.zeros(8, dtype=torch.float64, device="cuda")
    for dr in range(4):
        dfl = d.t(flux[dr])
        for side in range(2):
            chk(lk, lk.lk_ke_flux_from_fluxes(out.data_ptr() + 8 * (2 * dr + side), C.byref(d.g), dfl.data_ptr(), d.velocities.data_ptr(),
                                              d.vxface.data_ptr(), d.vyface.data_ptr(), dr, side, 1.7, None), "kef")
    got = out.cpu().numpy()
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want)), (got, want)
    at = (C.c_int * 8)(*([1] * 8))
    out2 = torch.full((8,), 7.0, dtype=torch.float64, device="cuda")
    chk(lk, lk.lk_ke_flux_boundaries(out2.data_ptr(), d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), 1.7,
                                     C.byref(at), None), "fused")
    got2 = out2.cpu().numpy()
    assert np.max(np.abs(got2 - want)) <= 1e-12 * np.max(np.abs(want)), (got2, want)
    # a box that does not touch a boundary contributes nothing to it
    at = (C.c_int * 8)(1, 0, 0, 1, 1, 1, 0, 0)
    chk(lk, lk.lk_ke_flux_boundaries(out2.data_ptr(), d.f.data_ptr(), C.byref(d.g), d.velocities.data_ptr(), C.byref(d.accel), 1.7,
                                     C.byref(at), None), "fused")
    got3 = out2.cpu().numpy()
    for k in range(8):
        assert got3[k] == (got2[k] if at[k] else 0.0)
