#!/usr/bin/env python
"""Generate julia/obm_fill.jl — the parameter-fill constructors of the reference-side binding: one keyword constructor
call per C parameter struct, every member assigned from the OceanBioME.jl object field it comes from.

How the table is obtained (so that it cannot drift from what the parity tests exercise): the Python host mirror
(oceanbiome.jl_b200/*.py) keeps the reference's struct and field names.  Every numeric field of a mirror object is set to
a unique sentinel value, the mirror's own `c_params()` (the function every GPU parity test goes through) is run, and each
member of the resulting C struct is traced back to the field whose sentinel it carries.  Members that are not plain
copies (enumerations, flags, the two day lengths) come from the MANUAL table below.  tests/test_abi.py then checks the
result against the reference's own struct definitions (every numeric reference field consumed, by the right name).

usage: python scripts/gen_julia_fill.py > julia/obm_fill.jl"""
import ctypes as C
import dataclasses
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oceanbiome_b200 as ob  # noqa: E402
from oceanbiome_b200 import _lib, pisces  # noqa: E402


class Sentinels:
    """Replace every float reachable from an object by a unique value and remember where it lives (as a Julia path)."""

    def __init__(self):
        self.where = {}
        self.count = 0

    def next(self, jpath):
        self.count += 1
        v = 1.0 + self.count * 2.0 ** -16
        self.where[v] = jpath
        return v

    def mark(self, obj, jpath, skip=()):
        fields = [f.name for f in dataclasses.fields(obj)] if dataclasses.is_dataclass(obj) else [k for k in vars(obj) if not k.startswith("_")]
        for name in fields:
            if name in skip:
                continue
            v = getattr(obj, name)
            here = f"{jpath}.{name}"
            if isinstance(v, bool) or v is None:
                continue
            if isinstance(v, (int, float)):
                if isinstance(v, float):
                    setattr(obj, name, self.next((jpath, name, None)))
            elif isinstance(v, dict) and v and all(isinstance(x, float) for x in v.values()):
                setattr(obj, name, {k: self.next((here, k, None)) for k in v})
            elif isinstance(v, (tuple, list)) and v and all(isinstance(x, float) for x in v):
                setattr(obj, name, type(v)(self.next((jpath, name, n + 1)) for n in range(len(v))))
            elif dataclasses.is_dataclass(v):
                self.mark(v, here)


def jexpr(entry):
    parent, name, index = entry
    if index is None:
        return f"fieldor({parent}, :{name})"
    return f"tupleor({parent}, :{name}, {index})"


def trace(struct, sent, manual, prefix=""):
    """→ list of (member, julia expression) for a ctypes struct filled by the mirror's c_params()."""
    out = []
    for fname, ftype in struct._fields_:
        key = prefix + fname
        v = getattr(struct, fname)
        if fname.startswith("_pad"):
            out.append((fname, "Int32(0)"))
        elif key in manual:
            out.append((fname, manual[key]))
        elif isinstance(v, C.Structure):
            inner = trace(v, sent, manual, key + ".")
            out.append((fname, (type(v).__name__, inner)))
        elif isinstance(v, C.Array):
            items = []
            for n, x in enumerate(v):
                if x in sent.where:
                    items.append(jexpr(sent.where[x]))
                elif x == 0:
                    items.append("0.0")
                else:
                    raise SystemExit(f"{key}[{n}] = {x!r}: not a sentinel — add a MANUAL entry")
            if all(i == "0.0" for i in items):  # not set by this variant
                out.append((fname, f"ntuple(_ -> 0.0, {len(items)})"))
            else:
                out.append((fname, "(" + ", ".join(items) + ("," if len(items) == 1 else "") + ")"))
        elif ftype is C.c_double:
            if v in sent.where:
                out.append((fname, jexpr(sent.where[v])))
            elif v == 0.0:
                out.append((fname, None))  # not set by this variant
            else:
                raise SystemExit(f"{key} = {v!r}: not a sentinel — add a MANUAL entry")
        else:
            raise SystemExit(f"{key}: integer member without a MANUAL entry")
    return out


def merge(a, b):
    """Union of two traces of the same struct (different component choices fill different members)."""
    out = []
    for (na, ea), (nb, eb) in zip(a, b):
        assert na == nb
        unset = lambda e: e is None or (isinstance(e, str) and e.startswith("ntuple(_ -> 0.0"))  # noqa: E731
        if unset(ea):
            out.append((na, eb))
        elif unset(eb) or ea == eb:
            out.append((na, ea))
        elif isinstance(ea, tuple) and isinstance(eb, tuple):
            out.append((na, (ea[0], merge(ea[1], eb[1]))))
        else:
            raise SystemExit(f"{na}: {ea} vs {eb} — two sources for one member, add a MANUAL entry")
    return out


def jname(cname):
    return "".join(p.capitalize() for p in cname.split("_"))


def emit(name, args, members, doc, indent="    "):
    print(f"# {doc}")
    print(f"{name}({args}) = {name}(;")
    emit_members(members, indent)
    print(")\n")


def emit_members(members, indent):
    for fname, e in members:
        if e is None:
            e = "0.0"
        if isinstance(e, tuple):
            print(f"{indent}{fname} = {jname(e[0])}(;")
            emit_members(e[1], indent + "    ")
            print(f"{indent}),")
        else:
            print(f"{indent}{fname} = {e},")


def main():
    print("# obm_fill.jl — GENERATED by scripts/gen_julia_fill.py from the Python host mirror's own c_params(); do not edit.")
    print("# One keyword constructor per C parameter struct of include/obm_b200.h; every member names the OceanBioME.jl field it")
    print("# is filled from (reference struct definitions cited per function; tests/test_abi.py checks field names and coverage).")
    print("# `fieldor(x, :f)` = Float64(x.f), or 0.0 when the component has no such field / is `nothing` / the field is `nothing`;")
    print("# `tupleor(x, :f, n)` = Float64(x.f[n]) under the same rule (helpers in OceanBioMEB200.jl).\n")
    g = ob.RectilinearGrid(size=(2, 2, 2), extent=(1.0, 1.0, 1.0), device="cpu")

    # ---- NutrientsPlanktonDetritus (NutrientsPlanktonDetritus.jl:27-33; plankton.jl:19-58; nutrients.jl:17-19,42-44;
    #      detritus.jl:25-37,71-82,264-272; oxygen.jl:14-17) -------------------------------------------------------------
    manual = {"nutrients": "nutrient_kind(bgc.nutrients)", "detritus": "detritus_kind(bgc.detritus)",
              "carbonate_replicates": "carbonate_replicates(bgc.carbonate_system)", "oxygen": "Int32(!isnothing(bgc.oxygen))",
              "light_limitation": "light_limitation_kind(bgc.plankton.light_limitation)",
              "phytoplankton_mortality_formulation": "formulation_kind(bgc.plankton.phytoplankton_mortality_formulation)",
              "grazing_concentration_formulation": "formulation_kind(bgc.plankton.grazing_concentration_formulation)",
              "has_temperature_coefficient": "Int32(!isnothing(bgc.plankton.temperature_coefficient))"}
    merged = None
    for nut, det in ((ob.NitrateAmmoniaIron, ob.TwoParticleAndDissolved), (ob.Nutrient, ob.Detritus),
                     (ob.NitrateAmmonia, ob.VariableRedfieldDetritus)):
        s = Sentinels()
        m = ob.NutrientsPlanktonDetritus(nut(), ob.PhytoZoo(temperature_coefficient=1.5), det(), ob.CarbonateSystem(), ob.Oxygen())
        for comp in ("nutrients", "plankton", "detritus", "oxygen"):
            s.mark(getattr(m, comp), f"bgc.{comp}")
        t = trace(m.c_params(), s, manual)
        merged = t if merged is None else merge(merged, t)
    emit("ObmNpdParams", "bgc::NutrientsPlanktonDetritus", merged,
         "NutrientsPlanktonDetritus{NUT, PLA, DET, CAR, OXY} → obm_npd_params")

    # ---- PISCES (PISCES.jl:53-92 and the component structs cited in oceanbiome.jl_b200/pisces.py) ------------------------
    s = Sentinels()
    bgc = ob.PISCES(g)
    u = bgc.underlying_biogeochemistry
    for comp in ("phytoplankton", "zooplankton", "dissolved_organic_matter", "particulate_organic_matter", "nitrogen", "iron",
                 "oxygen", "latitude"):
        s.mark(getattr(u, comp), f"bgc.{comp}")
    for name in ("first_anoxia_threshold", "second_anoxia_threshold", "nitrogen_redfield_ratio", "phosphate_redfield_ratio",
                 "mixed_layer_shear", "background_shear", "silicate_climatology"):
        setattr(u, name, s.next(("bgc", name, None)))
    manual = {"day_length_growth": "Float64(bgc.day_length(prescribed_latitude(bgc), clock.time))  # growth_rate.jl:30 — the reference's (φ, t) order (ModelLatitude: per row, obm_pisces_tendencies_rows)",
              "day_length_chlorophyll": "Float64(bgc.day_length(clock.time, prescribed_latitude(bgc)))  # growth_rate.jl:143 — (t, φ)"}
    for cls in ("nano", "diatoms"):
        manual[f"{cls}.growth_rate_kind"] = f"growth_rate_kind(bgc.phytoplankton.{cls}.growth_rate)"
        manual[f"{cls}.silicate_limited"] = f"Int32(bgc.phytoplankton.{cls}.nutrient_limitation.silicate_limited)"
    saved = u.day_length
    u.day_length = lambda a, b: 0.0  # the two day lengths are MANUAL entries
    t = trace(u.c_params(0.0), s, manual)
    u.day_length = saved
    emit("ObmPiscesParams", "bgc::PISCES, clock", t,
         "PISCES{…} → obm_pisces_params (the two day lengths depend on the clock: refill them every stage)")

    # ---- light (2band.jl:35-45; multi_band.jl:26-37) ----------------------------------------------------------------------------
    s = Sentinels()
    tb = ob.TwoBandPhotosyntheticallyActiveRadiation(grid=g)
    s.mark(tb, "par", skip=("grid", "field", "surface_PAR", "discrete_form", "parameters"))
    emit("ObmTwobandParams", "par::TwoBandPhotosyntheticallyActiveRadiation", trace(tb.c_params(), s, {}),
         "TwoBandPhotosyntheticallyActiveRadiation → obm_twoband_params")
    s = Sentinels()
    mb = ob.MultiBandPhotosyntheticallyActiveRadiation(grid=g)
    for name in ("water_attenuation_coefficient", "chlorophyll_exponent", "chlorophyll_attenuation_coefficient", "surface_PAR_division"):
        setattr(mb, name, [s.next(("par", name, n + 1)) for n in range(len(mb.bands))])
    t = trace(mb.c_params(), s, {"nbands": "Int32(length(par.fields))"})
    # the C arrays hold OBM_MAX_BANDS entries; bands beyond length(par.fields) are 0
    nb = len(mb.bands)
    cp = mb.c_params()
    t = [(n, "(" + ", ".join(f"bandor(par.{n}, {q + 1})" for q in range(len(getattr(cp, n)))) + ")")
         if isinstance(getattr(cp, n), C.Array) else (n, e) for n, e in t]
    assert nb == 3
    emit("ObmMultibandParams", "par::MultiBandPhotosyntheticallyActiveRadiation", t,
         "MultiBandPhotosyntheticallyActiveRadiation → obm_multiband_params (`bandor(v, n)` = Float64(v[n]) for n ≤ length(v), else 0.0)")

    # ---- sediments (simple_multi_G.jl:15-38; instant_remineralisation.jl:13-19) ---------------------------------------------------
    merged = None
    variants = ((ob.SimpleMultiGSediment, {}), (ob.SimpleMultiGSediment, {"sinking_carbon": ("sPOC", "bPOC"), "sinking_nitrogen": ("sPON", "bPON")}),
                (ob.InstantRemineralisationSediment, {}))
    manual = {"model": "sediment_kind(sed.biogeochemistry)", "timestepper": "timestepper_kind(sed.timestepper)",
              "advection": "advection_kind(advection)", "carbon": "Int32(has_carbon(sed.biogeochemistry))",
              "nsinking_nitrogen": "Int32(length(sinking_nitrogen(sed.biogeochemistry)))",
              "nsinking_carbon": "Int32(length(sinking_carbon(sed.biogeochemistry)))"}
    for ctor, kw in variants:
        s = Sentinels()
        sed = ctor(g, **kw)
        b = sed.biogeochemistry
        for name in [f.name for f in dataclasses.fields(b)] if dataclasses.is_dataclass(b) else list(vars(b)):
            v = getattr(b, name)
            if isinstance(v, float):
                setattr(b, name, s.next(("sed.biogeochemistry", name, None)))
            elif isinstance(v, (tuple, list)) and v and all(isinstance(x, float) for x in v):
                setattr(b, name, tuple(s.next(("sed.biogeochemistry", name, n + 1)) for n in range(len(v))))
        t = trace(sed.c_params(), s, manual)
        merged = t if merged is None else merge(merged, t)
    emit("ObmSedimentParams", "sed::BiogeochemicalSediment, advection", merged,
         "BiogeochemicalSediment{<:SimpleMultiG | <:InstantRemineralisation} → obm_sediment_params (`advection` = the model's tracer advection scheme)")


if __name__ == "__main__":
    main()
