"""Reader for the reference's `.pp` input decks: Perl-style constant lines
`$name = expression;` followed by `key = value ...` parameter lines in which `$name` is replaced by the
constant's value.  The reference pipes the deck through an embedded Perl interpreter and interpolates the
constants as strings, i.e. with 15 significant digits (LokiParser.C:94-130); this reader evaluates the same
arithmetic in Python (+ - * / ** sqrt and parentheses are all the decks use) and rounds the same way, so a
deck parsed here gives the numbers Loki sees.  Host-side set-up only.

    params = pp.parse(open("planeIAW.pp").read())
    deck = pp.deck_from_params(params)          # loki_b200.decks.Deck / VMDeck
"""
import math
import re

from . import decks as _d

_CONST = re.compile(r"^\s*\$([A-Za-z_]\w*)\s*=\s*(.+?);\s*(#.*)?$")
_SAFE = {"sqrt": math.sqrt, "exp": math.exp, "log": math.log, "sin": math.sin, "cos": math.cos, "atan2": math.atan2,
         "abs": abs, "__builtins__": {}}


def _perl_number_string(v):
    """how Perl prints a number inside a string: %.15g"""
    return "%.15g" % v


def parse(text):
    """-> dict key -> list of tokens (strings; quotes stripped)"""
    consts, params = {}, {}
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line.startswith("#"):
            continue
        m = _CONST.match(line)
        if m:
            expr = re.sub(r"\$([A-Za-z_]\w*)", lambda mm: repr(consts[mm.group(1)]), m.group(2))
            consts[m.group(1)] = float(eval(expr, _SAFE, {}))
            continue
        if "=" not in line:
            continue
        key, val = line.split("=", 1)
        val = val.split("#", 1)[0].strip()
        val = re.sub(r"\$([A-Za-z_]\w*)", lambda mm: _perl_number_string(consts[mm.group(1)]), val)
        toks = re.findall(r'"[^"]*"|\S+', val)
        params[key.strip()] = [t.strip('"') for t in toks]
    return params


def _f(params, key, default=None, idx=0):
    if key not in params:
        if default is None:
            raise KeyError(key)
        return default
    return float(params[key][idx])


def _s(params, key, default=None):
    if key not in params:
        return default
    return " ".join(params[key])


def _species(params, k):
    pre = "kinetic_species.%d." % k
    nv = (int(params[pre + "Nv"][0]), int(params[pre + "Nv"][1]))
    vlim = tuple(float(t) for t in params[pre + "velocity_limits"][:4])
    mass, charge = _f(params, pre + "mass"), _f(params, pre + "charge")
    name = _s(params, pre + "name", "species%d" % k)
    icn = _s(params, pre + "ic.name")
    g = lambda key, dflt=0.0: _f(params, pre + "ic." + key, dflt)
    driver = None
    if int(_f(params, pre + "num_external_drivers", 0.0)) > 0:
        dp = pre + "external_driver.1."
        if _s(params, dp + "name") != "Shaped Ramped Cosine Driver":
            raise ValueError("unsupported driver %r" % _s(params, dp + "name"))
        q = lambda key, dflt=0.0: _f(params, dp + key, dflt)
        # ShapedRampedCosineDriver.C:255-320: old t_ramp / t_off syntax maps to (t_rampup, t_hold = 0, t_rampdown)
        if (dp + "t_ramp") in params:
            tr, th, td = q("t_ramp"), 0.0, q("t_off")
        else:
            tr, th, td = q("t_rampup"), q("t_hold"), q("t_rampdown")
        driver = [0.0] * 16
        driver[0], driver[1], driver[2], driver[3], driver[4], driver[5] = q("xwidth"), q("ywidth"), q("shape"), q("omega"), q("E_0"), q("t0")
        driver[6], driver[7], driver[8] = tr, th, td
        driver[9], driver[10], driver[11], driver[12], driver[13] = q("x_shape"), q("lwidth"), q("x0"), q("alpha"), q("t_res")
    if icn == "Perturbed Maxwellian":
        return _d.Species(name, nv, vlim, mass, charge, tx=g("tx", 1.0), ty=g("ty", 1.0), A=g("A"), B=g("B"), Cc=g("C"),
                          kx1=g("kx1"), ky1=g("ky1"), kx2=g("kx2"), ky2=g("ky2"), frac=g("frac", 1.0), driver=driver,
                          vx0=g("vx0"), vy0=g("vy0"), x_wave_number=g("x_wave_number"), y_wave_number=g("y_wave_number"),
                          flow_phase=g("phase"))
    if icn == "Interpenetrating Stream":
        if _s(params, pre + "ic.syntax", "half plane") != "half plane":
            raise ValueError("only the half-plane syntax of the Interpenetrating Stream IC is supported")
        st = dict(tl=g("tl"), tt=g("tt"), theta=g("theta"), d=g("d"), beta=g("beta"), floor=g("floor", 0.0), frac=g("frac"))
        if _s(params, pre + "ic.two_sided", "false") == "true":
            st["two_sided"] = True
            st["frac2"] = g("frac2")
        if _s(params, pre + "ic.centered", "false") == "true":
            st["centered"] = True
        return _d.Species(name, nv, vlim, mass, charge, stream=st, driver=driver)
    raise ValueError("unsupported initial condition %r" % icn)


def deck_from_params(params, name="deck"):
    n = (int(params["N"][0]), int(params["N"][1]))
    xlim = tuple(float(t) for t in params["domain_limits"][:4])
    if params.get("periodic_dir", ["true", "true"])[:2] != ["true", "true"]:
        raise ValueError("the host mirror is periodic in x and y")
    order = int(_f(params, "spatial_solution_order", 4.0))
    rk = int(_f(params, "temporal_solution_order", 4.0))
    cfl = _f(params, "cfl", 1.0)
    ns = int(_f(params, "number_of_species"))
    species = [_species(params, k) for k in range(1, ns + 1)]
    if _s(params, "sys_type", "poisson") == "maxwell":
        em_ics, vel_ics = [], []
        k = 1
        while ("maxwell.em_ic.%d.name" % k) in params:
            pre = "maxwell.em_ic.%d." % k
            em_ics.append(dict(field=_s(params, pre + "field"), xamp=_f(params, pre + "xamp", 0.0), yamp=_f(params, pre + "yamp", 0.0),
                               zamp=_f(params, pre + "zamp", 0.0), kx=_f(params, pre + "x_wave_number", 0.0),
                               ky=_f(params, pre + "y_wave_number", 0.0), phase=_f(params, pre + "phase", 0.0)))
            k += 1
        k = 1
        while ("maxwell.vel_ic.%d.name" % k) in params:
            pre = "maxwell.vel_ic.%d." % k
            vel_ics.append(dict(amp=_f(params, pre + "amp", 0.0), kx=_f(params, pre + "x_wave_number", 0.0),
                                ky=_f(params, pre + "y_wave_number", 0.0), phase=_f(params, pre + "phase", 0.0)))
            k += 1
        if rk != 4:
            raise ValueError("the Vlasov-Maxwell host mirror integrates with RK4")
        deck = _d.VMDeck(name, n, xlim, species, _f(params, "light_speed"), _f(params, "maxwell.avWeak", 0.0),
                         _f(params, "maxwell.avStrong", 0.0), em_ics, vel_ics, order=order, cfl=cfl)
    else:
        deck = _d.Deck(name, n, xlim, species, order=order, rk=rk, cfl=cfl)
    # time-step controls of Simulation (Simulation.C:415-440); not part of the deck's physics, kept aside
    deck.run = dict(final_time=_f(params, "final_time", 1.0), save_times=_f(params, "save_times", 1.0),
                    max_step=int(_f(params, "max_step", 1000000.0)))
    return deck


def load(path):
    import os
    return deck_from_params(parse(open(path).read()), name=os.path.splitext(os.path.basename(path))[0])
