"""The three initial conditions of PerturbedMaxwellianIC ("Perturbed Maxwellian", "Landau damping", "Maxwellian with
noise") as loki_b200/decks.py builds them, against a cell-by-cell restatement of PerturbedMaxwellianIC::cache and
getIC_At_Pt (PerturbedMaxwellianIC.C:95-281) and MaxwellianThermal (MaxwellianThermal.C:16-60) written with scalar
Python floats in the reference's operation order; and the deck reader's handling of their keys (parseParameters,
PerturbedMaxwellianIC.C:344-412; the constructor's defaults, :44-61).  Host-side set-up only: no GPU, no oracle."""
import math

import numpy as np
import pytest

from loki_b200 import decks, pp

cos = lambda v: float(np.cos(np.float64(v)))     # numpy's cos / exp on both sides: the test is about the ORDER of operations
exp = lambda v: float(np.exp(np.float64(v)))


def _restated_ic(deck, sp):
    """m_f(i1, i2, i3, i4) over the whole data box, as getIC_At_Pt returns it"""
    ng = deck.ng
    nx, ny = deck.n
    (_, _, nvx, nvy), dx = deck.geom_of(sp)
    xlo, xhi, ylo, yhi = deck.xlim
    Lx = dx[0] * nx
    pi = 4.0 * math.atan(1.0)
    fnorm = sp.mass / (2.0 * math.pi * math.sqrt(sp.tx * sp.ty))          # MaxwellianThermal.C:27-28
    a, b, c, phi = sp.A, sp.B, sp.Cc, sp.spatial_phase

    def thermal(sf, x3, x4):                                             # MaxwellianThermal::thermalFactor
        d3 = x3 - (sp.vx0 * sf + sp.vflowinitx)
        d4 = x4 - (sp.vy0 * sf + sp.vflowinity)
        return exp(-0.5 * ((d3 * d3) / (sp.tx / sp.mass) + (d4 * d4) / (sp.ty / sp.mass)))

    out = np.zeros((nvy + 2 * ng, nvx + 2 * ng, ny + 2 * ng, nx + 2 * ng))
    for j2 in range(ny + 2 * ng):
        x2 = ylo + ((j2 - ng) + 0.5) * dx[1]
        if deck.periodic[1]:
            if x2 < ylo:
                x2 += ny * dx[1]
            elif x2 > yhi:
                x2 -= ny * dx[1]
        for j1 in range(nx + 2 * ng):
            x1 = xlo + ((j1 - ng) + 0.5) * dx[0]
            if deck.periodic[0]:
                if x1 < xlo:
                    x1 += nx * dx[0]
                elif x1 > xhi:
                    x1 -= nx * dx[0]
            if sp.ic_option == 1:
                fx = 1.0 + a * cos(sp.kx1 * x1 + phi) * cos(sp.ky1 * x2 + phi) + b * cos(sp.kx2 * x1 + phi) + c * cos(sp.ky2 * x2 + phi)
            elif sp.ic_option == 2:
                fx = 1.0 + a * cos(sp.kx1 * x1 + sp.ky1 * x2 + phi)
            else:
                fx = 1.0
                for k in range(1, len(sp.noise_amp) + 1):
                    fx += sp.noise_amp[k - 1] * cos(2.0 * pi * k * (x1 + sp.noise_phase[k - 1]) / Lx + phi)
            sf = cos(sp.x_wave_number * x1 + sp.y_wave_number * x2 + sp.flow_phase) if not sp.factorable else 0.0
            for j4 in range(nvy + 2 * ng):
                x4 = sp.vlim[2] + ((j4 - ng) + 0.5) * dx[3]
                for j3 in range(nvx + 2 * ng):
                    x3 = sp.vlim[0] + ((j3 - ng) + 0.5) * dx[2]
                    fv = thermal(sf, x3, x4)
                    if sp.ic_option == 3:
                        out[j4, j3, j2, j1] = fx * fnorm * fv * sp.frac          # :240, :276-278
                    else:
                        out[j4, j3, j2, j1] = fnorm * fv * fx * sp.frac          # :226, :232, :279-281
    return out


def _species(option, **kw):
    sp = decks.Species("electron", (6, 5), (-6.0, 6.0, -5.0, 5.0), 1.5, -1.0, tx=1.25, ty=0.75, A=0.125, B=0.0625, Cc=0.03125,
                       kx1=1.0 / 3, ky1=1.0 / 7, kx2=2.0 / 3, ky2=3.0 / 7, frac=0.7, **kw)
    sp.ic_option, sp.spatial_phase = option, 0.3
    sp.vflowinitx, sp.vflowinity = 0.11, -0.07
    if option == 3:
        sp.noise_amp, sp.noise_phase = (0.01, 0.003, 0.0007), (0.1, 0.7, 1.9)
    return sp


@pytest.mark.parametrize("option", [1, 2, 3])
@pytest.mark.parametrize("drift", [False, True])
@pytest.mark.parametrize("periodic", [(True, True), (False, True)])
def test_initial_condition_variants_bit_for_bit(option, drift, periodic):
    kw = dict(vx0=0.2, vy0=-0.1, x_wave_number=0.4, y_wave_number=0.2, flow_phase=0.9) if drift else {}
    sp = _species(option, **kw)
    deck = decks.Deck("ic", (7, 5), (-9.0, 9.5, -6.0, 6.5), [sp], order=4)
    deck.periodic = periodic
    f, fx, fv, fnorm = deck.initial_state(sp)
    ref = _restated_ic(deck, sp)
    assert f.shape == ref.shape and np.array_equal(f, ref)
    if not drift:
        # what the device's inflow fill and the Krook term form from the tables, fnorm * fv * fx * frac in that order
        # (lk_device.cuh, inflow kind 1), is the same number -- for option 3 through the pre-multiplied fx and fnorm = 1
        j4, j3, j2, j1 = 1, 8, 3, 9
        assert fnorm * fv[j4, j3] * fx[j2, j1] * sp.frac == ref[j4, j3, j2, j1]
        assert (fnorm == 1.0) == (option == 3)
    if periodic == (False, True) and option != 3:
        # the x ghosts of a non-periodic direction are NOT the periodic image (PerturbedMaxwellianIC.C:135-143)
        ng = deck.ng
        assert not np.array_equal(fx[:, :ng], fx[:, -2 * ng:-ng])
        assert np.array_equal(fx[:ng, :], fx[-2 * ng:-ng, :])


BASE = """
domain_limits = -10. 10. -30. 30.
N = 12 6
periodic_dir = true true
number_of_species = 1
kinetic_species.1.name = "electron"
kinetic_species.1.velocity_limits = -7 7 -7 7
kinetic_species.1.Nv = 20 12
kinetic_species.1.mass = 1.0
kinetic_species.1.charge = -1.0
"""


def _load(extra):
    return pp.deck_from_params(pp.parse(BASE + extra))


def test_pp_reads_the_three_variants_and_the_reference_defaults():
    d = _load('kinetic_species.1.ic.name = "Perturbed Maxwellian"\nkinetic_species.1.ic.A = 0.1\n')
    sp = d.species[0]
    # a wave number the deck leaves out is 0.5, not 0 (PerturbedMaxwellianIC.C:54-60)
    assert (sp.kx1, sp.ky1, sp.kx2, sp.ky2) == (0.5, 0.5, 0.5, 0.5) and sp.ic_option == 1 and sp.spatial_phase == 0.0
    assert (sp.tx, sp.ty, sp.frac) == (1.0, 1.0, 1.0)
    d = _load("kinetic_species.1.ic.A = 0.1\n")                               # no name: the factory's default (ICFactory.C:26-27)
    assert d.species[0].ic_option == 1
    d = _load('kinetic_species.1.ic.name = "Landau damping"\nkinetic_species.1.ic.A = 0.01\nkinetic_species.1.ic.kx1 = 0.5\n'
              "kinetic_species.1.ic.ky1 = 0.0\nkinetic_species.1.ic.spatial_phase = 0.25\n")
    sp = d.species[0]
    assert sp.ic_option == 2 and sp.spatial_phase == 0.25 and sp.ky1 == 0.0
    f, fx, fv, fnorm = d.initial_state(sp)
    assert np.array_equal(f, _restated_ic(d, sp))
    assert np.ptp(fx, axis=0).max() == 0.0 and np.ptp(fx, axis=1).min() > 0.0       # a wave along x only
    d = _load('kinetic_species.1.ic.name = "Maxwellian with noise"\nkinetic_species.1.ic.number_of_noisy_modes = 2\n'
              "kinetic_species.1.ic.noise_amplitudes = 0.01 0.002\nkinetic_species.1.ic.noise_phases = 0.5 1.5\n")
    sp = d.species[0]
    assert sp.ic_option == 3 and sp.noise_amp == (0.01, 0.002) and sp.noise_phase == (0.5, 1.5)
    f, fx, fv, fnorm = d.initial_state(sp)
    assert np.array_equal(f, _restated_ic(d, sp))
    ng = d.ng
    assert np.array_equal(fx[:, :ng], fx[:, -2 * ng:-ng]) and np.ptp(fx, axis=0).max() == 0.0     # periodic image; no y dependence


@pytest.mark.parametrize("extra, message", [
    ('kinetic_species.1.ic.name = "Maxwellian with noise"\nkinetic_species.1.ic.number_of_noisy_modes = 3\n'
     "kinetic_species.1.ic.noise_amplitudes = 0.01 0.002\nkinetic_species.1.ic.noise_phases = 0.5 1.5 2.5\n", "noise_amplitudes"),
    ('kinetic_species.1.ic.name = "Landau damping"\nkinetic_species.1.ic.maxwellian_thermal = false\n', "Juttner"),
    ('kinetic_species.1.ic.name = "Perturbed Maxwellian"\nkinetic_species.1.ic.alpha = 1.0\n', "deprecated"),
    ('kinetic_species.1.ic.name = "Perturbed Maxwellian"\nkinetic_species.1.ic.phi = 0.1\n', "does not implement"),
    # a drifting Maxwellian in a Vlasov-Poisson system: the host mirror has no cached-distribution inflow for it
    ('kinetic_species.1.ic.name = "Perturbed Maxwellian"\nkinetic_species.1.ic.vx0 = 0.1\n', "non-factorable"),
    ('kinetic_species.1.ic.name = "Bump on tail"\n', "unsupported initial condition"),
])
def test_pp_refuses_what_it_cannot_reproduce(extra, message):
    with pytest.raises((ValueError, KeyError)) as e:
        _load(extra)
    assert message in str(e.value)


def test_vp_inflow_tables_refuse_a_drifting_maxwellian():
    sp = _species(1, vx0=0.2)
    deck = decks.Deck("ic", (7, 5), (-9.0, 9.5, -6.0, 6.5), [sp], order=4)
    with pytest.raises(ValueError):
        deck.set_inflow(None, None, 0)


DRIVER = ('kinetic_species.1.num_external_drivers = 1\n'
          'kinetic_species.1.external_driver.1.name = "Shaped Ramped Cosine Driver"\n')


def test_pp_driver_defaults_and_old_syntax():
    """ShapedRampedCosineDriver's constructor defaults (ShapedRampedCosineDriver.C:44-57) and the old k / L and
    t_ramp / t_off syntax (:241-291, 305-317)"""
    d = _load(DRIVER)
    drv = d.species[0].driver
    #        xwidth ywidth shape omega E_0  t0  rampup hold rampdown x_shape lwidth x0  alpha t_res
    assert drv[:14] == [0.5, 0.5, 1.0, 1.0, 0.01, 0.0, 10.0, 0.0, 10.0, 0.0, 0.5, 0.0, 0.0, 0.0]
    assert d.species[0].driver_phase == 0.0 and d.species[0].driver_shape_type == 0
    p = "kinetic_species.1.external_driver.1."
    d = _load(DRIVER + p + "kx = 0.25\n" + p + "Lx = 3.0\n" + p + "Ly = 5.0\n" + p + "kl = 4.0\n" + p + "t_ramp = 2.0\n")
    drv = d.species[0].driver
    assert drv[0] == 3.0 / (2.0 * 0.25) and drv[1] == 5.0 / (2.0 * 1.0) and drv[10] == 1.0 / (2.0 * 4.0)
    assert drv[6:9] == [2.0, 0.0, 10.0]                # t_off left out: the ramp-down keeps its default
    for bad in (p + "kx = 0.25\n" + p + "xwidth = 2.0\n", p + "Ly = 1\n" + p + "ywidth = 2.0\n",
                p + "phase = 0.1\n", p + "fwhm = 0.1\n" + p + "phase_decay_time_steps = 10\n"):
        with pytest.raises(ValueError):
            _load(DRIVER + bad)


STREAM = ('kinetic_species.1.ic.name = "Interpenetrating Stream"\nkinetic_species.1.ic.tl = 1.0\nkinetic_species.1.ic.tt = 1.0\n'
          "kinetic_species.1.ic.theta = 0.0\nkinetic_species.1.ic.d = 5.0\nkinetic_species.1.ic.beta = 0.768\n"
          "kinetic_species.1.ic.frac = 1.0\n")


def test_pp_stream_ic_required_keys():
    """parseParametersHalfPlane (InterpenetratingStreamIC.C:489-537): six required keys, frac2 with two_sided, not both
    slab forms, vl0 / vt0 retired"""
    d = _load(STREAM)
    assert d.species[0].stream["floor"] == 0.0 and "two_sided" not in d.species[0].stream
    for key, msg in (("tl", "Longitudinal"), ("tt", "Transverse"), ("theta", "Drift direction"), ("d ", "Distance"),
                     ("beta", "sharpness"), ("frac", "relative weight")):
        text = "".join(l + "\n" for l in STREAM.splitlines() if not l.startswith("kinetic_species.1.ic." + key.strip() + " "))
        with pytest.raises(ValueError) as e:
            _load(text)
        assert msg in str(e.value), (key, str(e.value))
    for extra, msg in (("kinetic_species.1.ic.two_sided = true\n", "Two sided"),
                       ("kinetic_species.1.ic.two_sided = true\nkinetic_species.1.ic.frac2 = 1\nkinetic_species.1.ic.centered = true\n", "Only one"),
                       ("kinetic_species.1.ic.vl0 = 0.1\n", "no longer used"),
                       ("kinetic_species.1.ic.maxwellian_thermal = false\n", "Juttner"),
                       ('kinetic_species.1.ic.syntax = "box"\n', "half-plane")):
        with pytest.raises(ValueError) as e:
            _load(STREAM + extra)
        assert msg in str(e.value), (extra, str(e.value))
