"""Reader for the reference's `.pp` input decks: Perl-style constant lines
`$name = expression;` followed by `key = value ...` parameter lines in which `$name` is replaced by the
constant's value.  The reference pipes the deck through an embedded Perl interpreter and interpolates the
constants as strings, i.e. with 15 significant digits (LokiParser.C:94-130); this reader evaluates the same
arithmetic in Python (+ - * / ** sqrt and parentheses are all the decks use) and rounds the same way, so a
deck parsed here gives the numbers Loki sees.  Host-side set-up only.

    params = pp.parse(open("planeIAW.pp").read())
    deck = pp.deck_from_params(params)          # loki_b200.decks.Deck / VMDeck
"""
import ast
import math
import operator
import re

import numpy as np

from . import decks as _d

_CONST = re.compile(r"^\s*\$([A-Za-z_]\w*)\s*=\s*(.+?);\s*(#.*)?$")
_FUNCS = {"sqrt": math.sqrt, "exp": math.exp, "log": math.log, "sin": math.sin, "cos": math.cos, "atan2": math.atan2,
          "abs": abs}
_BINOPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
           ast.Pow: operator.pow}


def _arith(expr):
    """numbers, + - * / **, parentheses and the functions above -- nothing else (deck text is untrusted input;
    no eval)"""
    def ev(n):
        if isinstance(n, ast.Expression):
            return ev(n.body)
        if isinstance(n, ast.Constant) and isinstance(n.value, (int, float)) and not isinstance(n.value, bool):
            return float(n.value)
        if isinstance(n, ast.BinOp) and type(n.op) in _BINOPS:
            return _BINOPS[type(n.op)](ev(n.left), ev(n.right))
        if isinstance(n, ast.UnaryOp) and isinstance(n.op, (ast.UAdd, ast.USub)):
            v = ev(n.operand)
            return -v if isinstance(n.op, ast.USub) else v
        if isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and n.func.id in _FUNCS and not n.keywords:
            return float(_FUNCS[n.func.id](*[ev(a) for a in n.args]))
        raise ValueError("unsupported expression in a deck constant: %r" % expr)
    return ev(ast.parse(expr.strip(), mode="eval"))


class Params(dict):
    """key -> tokens, remembering which keys the deck builder consumed: whatever it did not consume is either
    known not to change the physics of the path (output, restart, probes ...) or an error"""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.used = set()

    def __getitem__(self, key):
        self.used.add(key)
        return super().__getitem__(key)

    def __contains__(self, key):
        if super().__contains__(key):
            self.used.add(key)
            return True
        return False

    def get(self, key, default=None):
        if super().__contains__(key):
            return self[key]
        return default


# keys that never reach the Vlasov RHS path: output cadence and files, restart, probes, verbosity
_IGNORABLE = re.compile(
    r"^(verbosity|start_from_restart|restart\..*|number_of_probes|probe\.\d+\..*|save_data|save_coll_data|plot_ke_fluxes|"
    r"plot_times_per_file|kinetic_species\.\d+\.name|poisson\.solver_type|poisson\.discretization|"
    r"kinetic_species\.\d+\.num_tracking_particles|kinetic_species\.\d+\.num_noise_source_particles)$")


def _perl_number_string(v):
    """how Perl prints a number inside a string: %.15g"""
    return "%.15g" % v


def parse(text):
    """-> dict key -> list of tokens (strings; quotes stripped)"""
    consts, params = {}, Params()
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line.startswith("#"):
            continue
        m = _CONST.match(line)
        if m:
            expr = re.sub(r"\$([A-Za-z_]\w*)", lambda mm: repr(consts[mm.group(1)]), m.group(2))
            consts[m.group(1)] = _arith(expr)
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
    icn = _s(params, pre + "ic.name", "Perturbed Maxwellian")      # the factory's default (ICFactory.C:26-27)
    g = lambda key, dflt=0.0: _f(params, pre + "ic." + key, dflt)
    driver, driver_phase, driver_shape_type = None, 0.0, 0
    ndrv = int(_f(params, pre + "num_external_drivers", 0.0))
    if ndrv > 1:
        raise ValueError("species %d: %d external drivers; the host mirror applies one" % (k, ndrv))
    if ndrv > 0:
        dp = pre + "external_driver.1."
        if _s(params, dp + "name") != "Shaped Ramped Cosine Driver":
            raise ValueError("unsupported driver %r" % _s(params, dp + "name"))
        q = lambda key, dflt=0.0: _f(params, dp + key, dflt)
        # the constructor's defaults (ShapedRampedCosineDriver.C:44-57): widths 0.5, shape 1, omega 1, E_0 0.01, ramps 10
        # old t_ramp / t_off syntax maps to (t_rampup, t_hold = 0, t_rampdown) (:305-317)
        if (dp + "t_ramp") in params:
            tr, th, td = q("t_ramp", 10.0), 0.0, q("t_off", 10.0)
        else:
            tr, th, td = q("t_rampup", 10.0), q("t_hold", 0.0), q("t_rampdown", 10.0)

        def width(new, k_key, l_key):
            # old wave-number / length syntax (:241-291): width = L / (2 k), each defaulting to 1; mixing the two aborts
            if (dp + k_key) in params or (dp + l_key) in params:
                if (dp + new) in params:
                    raise ValueError("Mixed old and new syntax for %s/%s" % (k_key, l_key))
                return q(l_key, 1.0) / (2.0 * q(k_key, 1.0))
            return q(new, 0.5)
        driver = [0.0] * 16
        driver[0], driver[1] = width("xwidth", "kx", "Lx"), width("ywidth", "ky", "Ly")
        driver[2], driver[3], driver[4], driver[5] = q("shape", 1.0), q("omega", 1.0), q("E_0", 0.01), q("t0", 0.0)
        driver[6], driver[7], driver[8] = tr, th, td
        driver[9], driver[10], driver[11], driver[12], driver[13] = q("x_shape"), width("lwidth", "kl", "Ll"), q("x0"), q("alpha"), q("t_res")
        # the driver's phase is restart state, not deck input: 0 at start (ShapedRampedCosineDriver.C:30), restored from a dump
        # and random-walked only by a noisy driver (:181-186, 365-370).  parseParameters never reads a "phase" key, so the
        # reference would ignore one; refused here rather than silently dropped or silently applied.  Noisy drivers
        # (phase_decay_time_steps / fwhm, :325-351) are not implemented: their keys stay unread and are refused below
        if (dp + "phase") in params:
            raise ValueError("external_driver.1.phase is not a deck key of the Shaped Ramped Cosine Driver (the reference "
                             "never reads it; the phase is restart state)")
        # shape_type (:292-303): 0 sinusoidal / 1 exponential envelope
        st = _s(params, dp + "shape_type", "sin2")
        if st not in ("sin2", "exp"):
            raise ValueError("unknown driver shape_type %r" % st)   # LOKI_ABORT("Unknown shape type")
        driver_shape_type = 0 if st == "sin2" else 1
    # physics this mirror does not implement must not be dropped silently
    # collision operators (KineticSpecies.C:206-215, 1426-1430; CollisionOperatorFactory.C:30-55): the pitch-angle
    # operator is implemented (one per species); the Rosenbluth operators are not
    collision = None
    ncoll = int(_f(params, pre + "num_collision_operators", 0.0))
    if ncoll > 1:
        raise ValueError("species %d: more than one collision operator is not supported" % k)
    if ncoll == 1:
        cp = pre + "collision_operator.1."
        cname = _s(params, cp + "name")
        if cname != "Pitch Angle Collision Operator":
            raise ValueError("species %d: collision operator %r is not supported (only the pitch-angle operator)" % (k, cname))
        # PitchAngleCollisionOperator::parseParameters (PitchAngleCollisionOperator.C:190-270): all but
        # collision_conservative are required
        for key in ("collision_vfloor", "collision_vthermal_dt", "collision_nuCoeff"):
            if (cp + key) not in params:
                raise ValueError("Must supply " + key)
        for key in ("collision_vel_range_lo", "collision_vel_range_hi"):
            if (cp + key) not in params:
                raise ValueError("Must supply %s." % key)
            if len(params[cp + key]) != 2:
                raise ValueError("%s must have 2 entries." % key)
        collision = dict(range_lo=tuple(float(t) for t in params[cp + "collision_vel_range_lo"]),
                         range_hi=tuple(float(t) for t in params[cp + "collision_vel_range_hi"]),
                         vfloor=_f(params, cp + "collision_vfloor"), vthermal_dt=_f(params, cp + "collision_vthermal_dt"),
                         nu_coef=_f(params, cp + "collision_nuCoeff"),
                         conservative=int(_f(params, cp + "collision_conservative", 1.0)))
    if any(key.startswith(pre + "external_dist_krook.") for key in list(params.keys())):
        raise ValueError("species %d: external-distribution Krook layers are not supported" % k)
    # KrookLayer::parseParameters (KrookLayer.C:163-190): x1a / x1b / x2a / x2b switch the layer on; `power` is read
    # but unused by KrookLayer::initialize (the ramp is the order's polynomial)
    krook = {}
    for key in ("x1a", "x1b", "x2a", "x2b", "coefficient", "power"):
        if (pre + "krook." + key) in params:
            krook[key] = _f(params, pre + "krook." + key)
    if krook.get("power", 3.0) <= 0.0 or krook.get("coefficient", 1.0) < 0.0:
        raise ValueError("species %d: Krook layer needs a positive power and a non-negative coefficient" % k)   # KrookLayer.C:195-201
    if not any(e in krook for e in ("x1a", "x1b", "x2a", "x2b")):
        krook = None
    # TZSourceFactory::create (TZSourceFactory.C:22-56): the twilight-zone (manufactured-solution) forcings of the
    # reference's TrigTZ, EPWTZ and IAWTZ decks
    tz = None
    if (pre + "tz.name") in params:
        tzname = _s(params, pre + "tz.name")
        kinds = {"TrigTZSource": 1, "ElectronTrigTZSource": 2, "TwoSpecies_ElectronTrigTZSource": 3,
                 "TwoSpecies_IonTrigTZSource": 4}
        if tzname not in kinds:
            raise ValueError("species %d: twilight-zone source %r is not supported" % (k, tzname))
        if (pre + "tz.amp") not in params:
            raise ValueError("Must supply amp")                                   # TrigTZSource.C:28-31
        tz = dict(amp=_f(params, pre + "tz.amp"), kind=kinds[tzname])
        if kinds[tzname] >= 3:
            # TwoSpecies_ElectronTrigTZSource.C:28-41: amp, electron_mass and ion_mass are all required
            for key in ("electron_mass", "ion_mass"):
                if (pre + "tz." + key) not in params:
                    raise ValueError("Must supply " + key.replace("_", " "))
                tz[key] = _f(params, pre + "tz." + key)
    elif any(key.startswith(pre + "tz.") for key in list(params.keys())):
        raise ValueError("species %d: tz.* keys without tz.name" % k)
    _pm = {"Perturbed Maxwellian": 1, "Landau damping": 2, "Maxwellian with noise": 3}   # PerturbedMaxwellianIC.C:22-33, 351-362
    if icn in _pm:
        if _s(params, pre + "ic.maxwellian_thermal", "true") != "true":
            raise ValueError("species %d: the Juttner thermal factor is not supported" % k)
        if (pre + "ic.alpha") in params or (pre + "ic.beta") in params:
            raise ValueError("Parameters alpha and beta are deprecated in favor of tx and ty.")   # :388-390
        # the constructor's defaults (PerturbedMaxwellianIC.C:44-61): the four wave numbers default to 0.5, not 0
        sp = _d.Species(name, nv, vlim, mass, charge, tx=g("tx", 1.0), ty=g("ty", 1.0), A=g("A"), B=g("B"), Cc=g("C"),
                          kx1=g("kx1", 0.5), ky1=g("ky1", 0.5), kx2=g("kx2", 0.5), ky2=g("ky2", 0.5), frac=g("frac", 1.0),
                          driver=driver, vx0=g("vx0"), vy0=g("vy0"), x_wave_number=g("x_wave_number"),
                          y_wave_number=g("y_wave_number"), flow_phase=g("phase"))
        sp.ic_option, sp.spatial_phase = _pm[icn], g("spatial_phase")                           # :410-411
        if (pre + "ic.number_of_noisy_modes") in params:                                        # :380-386, both arrays required
            nn = int(_f(params, pre + "ic.number_of_noisy_modes"))
            for key in ("noise_amplitudes", "noise_phases"):
                if len(params[pre + "ic." + key]) < nn:
                    raise ValueError("species %d: ic.%s needs %d values" % (k, key, nn))
            sp.noise_amp = tuple(float(t) for t in params[pre + "ic.noise_amplitudes"][:nn])
            sp.noise_phase = tuple(float(t) for t in params[pre + "ic.noise_phases"][:nn])
        sp.driver_phase, sp.driver_shape_type = driver_phase, driver_shape_type
        sp.vflowinitx, sp.vflowinity = g("vflowinitx"), g("vflowinity")      # MaxwellianThermal.C:48-49
        sp.krook = krook
        sp.collision = collision
        sp.tz = tz
        return sp
    if icn == "External 2D":
        # External2DIC::parseParameters (External2DIC.C:318-353) and its constructor's read of "2D dist" (:74-126)
        if (pre + "ic.file_name") not in params:
            raise ValueError("Must supply name of external 2D distribution file.")
        if _s(params, pre + "ic.maxwellian_thermal", "true") != "true":
            raise ValueError("species %d: the Juttner thermal factor is not supported" % k)
        if g("vx0") != 0.0 or g("vy0") != 0.0:
            raise ValueError("species %d: a non-factorable External 2D initial condition (vx0 / vy0) is not supported" % k)
        import os
        from . import h5lite
        fname = _s(params, pre + "ic.file_name")
        path = fname if os.path.isabs(fname) else os.path.join(getattr(params, "base_dir", None) or ".", fname)
        root = h5lite.read(path)
        if "2D dist" not in root:
            raise ValueError('Can not open dataset "2D dist".')
        sp = _d.Species(name, nv, vlim, mass, charge, tx=g("tx", 1.0), ty=g("ty", 1.0), driver=driver)
        sp.external = np.array(root["2D dist"].data, dtype=np.float64)
        sp.external_frac = g("frac", 1.0)
        sp.vflowinitx, sp.vflowinity = g("vflowinitx"), g("vflowinity")
        sp.driver_phase, sp.driver_shape_type = driver_phase, driver_shape_type
        sp.krook = krook
        sp.collision = collision
        sp.tz = tz
        return sp
    if icn == "Interpenetrating Stream":
        if _s(params, pre + "ic.syntax", "half plane") != "half plane":
            raise ValueError("only the half-plane syntax of the Interpenetrating Stream IC is supported")
        if _s(params, pre + "ic.maxwellian_thermal", "true") != "true":
            raise ValueError("species %d: the Juttner thermal factor is not supported" % k)
        # parseParametersHalfPlane (InterpenetratingStreamIC.C:489-537): these six are required, each with its own message
        required = (("tl", "Longitudinal temperature is required."), ("tt", "Transverse temperature is required."),
                    ("theta", "Drift direction is required."), ("d", "Distance from origin is required."),
                    ("beta", "Transition sharpness parameter is required."),
                    ("frac", "Species relative weight parameter is required."))
        for key, msg in required:
            if (pre + "ic." + key) not in params:
                raise ValueError(msg)
        for key in ("vl0", "vt0"):                                                              # :507-512
            if (pre + "ic." + key) in params:
                raise ValueError("InterpenetratingStream %s input no longer used in favor of vflowinit%s." % (key, "xy"[key == "vt0"]))
        st = dict(tl=g("tl"), tt=g("tt"), theta=g("theta"), d=g("d"), beta=g("beta"), floor=g("floor", 0.0), frac=g("frac"))
        two_sided = _s(params, pre + "ic.two_sided", "false") == "true"
        centered = _s(params, pre + "ic.centered", "false") == "true"
        if two_sided and centered:
            raise ValueError("Only one of a two sided or centered slab may be specified.")        # :531-533
        if two_sided:
            if (pre + "ic.frac2") not in params:
                raise ValueError("Two sided species relative weight parameter is required.")     # :534-536
            st["two_sided"] = True
            st["frac2"] = g("frac2")
        if centered:
            st["centered"] = True
        sp = _d.Species(name, nv, vlim, mass, charge, stream=st, driver=driver)
        sp.driver_phase, sp.driver_shape_type = driver_phase, driver_shape_type
        sp.krook = krook
        sp.collision = collision
        sp.tz = tz
        return sp
    raise ValueError("unsupported initial condition %r" % icn)


def deck_from_params(params, name="deck"):
    n = (int(params["N"][0]), int(params["N"][1]))
    xlim = tuple(float(t) for t in params["domain_limits"][:4])
    periodic = tuple(t == "true" for t in params.get("periodic_dir", ["true", "true"])[:2])
    order = int(_f(params, "spatial_solution_order", 4.0))
    rk = int(_f(params, "temporal_solution_order", 4.0))
    cfl = _f(params, "cfl", 0.9)                      # Simulation.C:174
    if _s(params, "do_relativity", "false") == "true":
        raise ValueError("do_relativity = true is not supported (non-relativistic velocity tables)")
    use_new_bcs = _s(params, "use_new_bcs", "false") == "true"            # VPSystem.C:819-821
    if _s(params, "do_new_algorithm", "true") != "true":
        raise ValueError("do_new_algorithm = false (flux form) is not the path this library accelerates")
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
        if periodic != (True, True) or use_new_bcs or any(getattr(sp, "krook", None) or getattr(sp, "collision", None) for sp in species):
            raise ValueError("the Vlasov-Maxwell host mirror is periodic, with the standard boundary fill, no Krook layers "
                             "and no collision operators")
        deck = _d.VMDeck(name, n, xlim, species, _f(params, "light_speed"), _f(params, "maxwell.avWeak", 0.0),
                         _f(params, "maxwell.avStrong", 0.0), em_ics, vel_ics, order=order, cfl=cfl, rk=rk)
    else:
        deck = _d.Deck(name, n, xlim, species, order=order, rk=rk, cfl=cfl)
        deck.periodic, deck.use_new_bcs = periodic, use_new_bcs
        for sp in species:
            if sp.stream is None and not sp.factorable:
                raise ValueError("species %r: a non-factorable initial condition (ic.vx0 / ic.vy0 != 0) is not supported in "
                                 "a Vlasov-Poisson system (its velocity ghosts need the cached distribution)" % sp.name)
    # time-step controls of Simulation (Simulation.C:415-440); not part of the deck's physics, kept aside
    # Simulation's defaults (Simulation.C:166-181): final_time 1, save_times 1, sequence_write_times 1, max_step 0
    deck.run = dict(final_time=_f(params, "final_time", 1.0), save_times=_f(params, "save_times", 1.0),
                    sequence_write_times=_f(params, "sequence_write_times", 1.0), max_step=int(_f(params, "max_step", 0.0)))
    # restart cadence and paths (RestartManager.C:150-192); start_from_restart (Simulation.C:211-215)
    rs = {}
    for key, conv in (("time_interval", float), ("step_interval", int), ("max_files_for_write", int)):
        if ("restart." + key) in params:
            v = params["restart." + key]
            rs["max_files" if key == "max_files_for_write" else key] = conv(float(v[0] if isinstance(v, (list, tuple)) else v))
    for key in ("write_directory", "read_directory"):
        if ("restart." + key) in params:
            v = params["restart." + key]
            rs[key] = str(v[0] if isinstance(v, (list, tuple)) else v).strip('"')
    if "time_interval" in rs and "step_interval" in rs:
        raise ValueError("Must set restart frequency for only one of steps or time, not both.")
    if "start_from_restart" in params:
        v = params["start_from_restart"]
        rs["start_from_restart"] = str(v[0] if isinstance(v, (list, tuple)) else v).strip('"') == "true"
    deck.run["restart"] = rs
    # probes (Simulation.C:393-412): fractions of the domain; without number_of_probes one probe at (0.5, 0) -- the
    # reference assigns m_probes[X1][0] twice and leaves the y fraction at 0
    if "number_of_probes" in params:
        npr = int(_f(params, "number_of_probes"))
        deck.probes = [tuple(float(t) for t in params["probe.%d.location" % (k + 1)][:2]) for k in range(npr)]
    else:
        deck.probes = [(0.5, 0.0)]
    if isinstance(params, Params):
        left = sorted(key for key in params.keys() if key not in params.used and not _IGNORABLE.match(key))
        if left:
            raise ValueError("deck keys this reader does not implement (they would change the run): " + ", ".join(left))
    return deck


def load(path):
    import os
    params = parse(open(path).read())
    params.base_dir = os.path.dirname(os.path.abspath(path))      # external files are named relative to the deck
    return deck_from_params(params, name=os.path.splitext(os.path.basename(path))[0])
