"""The C-ABI shared library loads on a machine without a GPU, exports every symbol that
include/obm_b200.h declares, agrees with the ctypes mirror on struct sizes, and rejects bad
arguments with the documented error codes before launching anything (no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from oceanbiome_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "obm_b200.h"), encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(obm_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libobm_b200.so does not export {n}"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes PROTOTYPES and the header disagree"
    assert lib.obm_version() == 100


def test_struct_sizes_match_ctypes_mirror():
    lib = _lib.load()
    for name, cls in _lib.STRUCTS.items():
        assert lib.obm_sizeof(name.encode()) == C.sizeof(cls), name
    assert lib.obm_sizeof(b"no_such_struct") == -3


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    g = _lib.obm_grid(4, 4, 4, 3, 3, 3, 0, 0, 0, 0, None, None)
    p = _lib.obm_npd_params()
    assert lib.obm_npd_tendencies(C.byref(g), C.byref(p), None, None, None, 0, None) == -1  # OBM_ENULL
    assert b"NULL" in lib.obm_last_error()
    p.nutrients = 7
    assert lib.obm_npd_tracer_names(C.byref(p), None) == -3  # OBM_EENUM
    p.nutrients, p.carbonate_replicates = 1, 99
    assert lib.obm_npd_tracer_names(C.byref(p), None) == -4  # OBM_ENOTIMPL
    tb = _lib.obm_twoband_params()
    bad = _lib.obm_grid(0, 4, 4, 3, 3, 3, 0, 0, 0, 0, None, None)
    dummy = C.c_void_p(8)
    assert lib.obm_par_twoband(C.byref(bad), C.byref(tb), dummy, None, 1.0, dummy, None) == -2  # OBM_ESIZE
    assert lib.obm_par_twoband(C.byref(g), C.byref(tb), dummy, None, 1.0, dummy, None) == -1  # zc/zf NULL
    assert lib.obm_carbon_chemistry(-1, None, dummy, dummy, dummy, dummy, None, None, None, None, 0, dummy, None) == -2
    assert lib.obm_carbon_chemistry(4, None, dummy, dummy, dummy, dummy, None, None, None, None, 42, dummy, None) == -3
    assert lib.obm_carbon_chemistry(0, None, None, None, None, None, None, None, None, None, 0, None, None) == 0
    assert lib.obm_inventory_workspace_bytes(5) == 5 * 148 * 5 * 8  # groups × (148 SMs × 5 resident blocks) partial sums of 8 bytes
    # the ensemble entry point refuses a bad sweep description before it touches the device
    p = _lib.obm_npd_params()
    ens = lambda nvary, which, values: lib.obm_npd_tendencies_ensemble(  # noqa: E731
        C.byref(g), C.byref(p), nvary, which, values, dummy, dummy, dummy, 0, None)
    assert ens(_lib.OBM_NPD_MAX_VARIED + 1, (C.c_int32 * 17)(), dummy) == -2 and b"nvary" in lib.obm_last_error()
    assert ens(-1, None, None) == -2
    assert ens(1, None, dummy) == -1 and ens(1, (C.c_int32 * 1)(0), None) == -1
    assert ens(1, (C.c_int32 * 1)(35), dummy) == -3 and b"not a parameter index" in lib.obm_last_error()
    assert ens(1, (C.c_int32 * 1)(-1), dummy) == -3


def test_tracer_names_from_the_library_match_reference_order():
    lib = _lib.load()
    p = _lib.obm_npd_params()
    p.nutrients, p.detritus, p.carbonate_replicates, p.oxygen = _lib.NUT_NITRATE_AMMONIA, _lib.DET_TWO_PARTICLE, 2, 1
    names = ((C.c_char * 16) * _lib.OBM_NPD_MAX_TRACERS)()
    n = lib.obm_npd_tracer_names(C.byref(p), names)
    got = [bytes(names[i]).split(b"\0")[0].decode() for i in range(n)]
    assert got == ["NO₃", "NH₄", "P", "Z", "sPOM", "bPOM", "DOM", "DIC1", "DIC2", "Alk1", "Alk2", "O₂"]


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load(str(tmp_path / "nope.so"))


def test_generated_julia_structs_are_in_sync():
    """julia/obm_structs.jl is generated from the ctypes mirrors; regenerate it when a struct changes."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "gen_julia_structs.py")], capture_output=True,
                         text=True, check=True).stdout
    assert out == open(os.path.join(root, "julia", "obm_structs.jl")).read()


def test_npd_parameter_index_space_matches_the_library():
    """`NutrientsPlanktonDetritus.parameter_index` (from the ctypes struct) ≡ `obm_npd_param_index` (from the C struct)."""
    import ctypes as C
    import oceanbiome_b200 as ob
    lib = _lib.load()
    doubles = [n for n, t in _lib.obm_npd_params._fields_ if t is C.c_double]
    assert len(doubles) == 35
    for n in doubles:
        assert lib.obm_npd_param_index(n.encode()) == ob.NutrientsPlanktonDetritus.parameter_index(n)
    assert lib.obm_npd_param_index(b"nutrients") == -3  # structural members cannot vary per member
    assert lib.obm_npd_param_index(b"bogus") == -3


def test_julia_ccall_signatures_have_the_prototypes_arity():
    """julia/OceanBioMEB200.jl cannot run here (no Julia in the image): at least every `ccall((:name, libobm), Cint,
    (types…), …)` in it — commented sketches included — names an exported symbol and lists as many argument types as the
    C prototype has parameters."""
    root = os.path.join(os.path.dirname(__file__), "..")
    src = open(os.path.join(root, "julia", "OceanBioMEB200.jl"), encoding="utf-8").read()
    src += open(os.path.join(root, "INTEGRATION.md"), encoding="utf-8").read()  # the binding sketches shown to maintainers
    src = "\n".join(line.lstrip().lstrip("#") for line in src.splitlines())  # the commented call sketches count too
    calls = re.findall(r"ccall\(\(:(\w+),\s*libobm\),\s*(\w+),\s*\(([^()]*)\)", src, flags=re.S)
    assert len(calls) >= 8
    for name, ret, types in calls:
        assert name in _lib.PROTOTYPES, name
        n = len([t for t in types.split(",") if t.strip()])
        assert n == len(_lib.PROTOTYPES[name][1]), f"{name}: Julia lists {n} argument types, the C prototype has {len(_lib.PROTOTYPES[name][1])}"
        assert ret == {C.c_int: "Cint", C.c_char_p: "Cstring", C.c_double: "Cdouble"}.get(_lib.PROTOTYPES[name][0], ret), name
        # scalars where C has scalars, pointers where C has pointers
        for pos, (jt, ct) in enumerate(zip([t.strip() for t in types.split(",") if t.strip()], _lib.PROTOTYPES[name][1])):
            scalar = {C.c_double: "Cdouble", C.c_int: "Cint", C.c_int64: "Int64", C.c_longlong: "Int64"}.get(ct)
            if scalar is not None:
                assert jt == scalar, f"{name} argument {pos}: Julia {jt}, C {ct.__name__}"
            else:
                assert any(k in jt for k in ("Ptr", "Ref", "F64", "Cstring")), f"{name} argument {pos}: Julia {jt} for a C pointer"


# ---- the reference-side binding against the reference's own struct definitions ----------------------------------------
REFERENCE = "/root/reference"  # present in the build container only; these checks are skipped elsewhere

# Julia path in julia/obm_fill.jl → (reference file, struct name, … alternatives the component may be)
PISCES_DIR = "src/Models/AdvectedPopulations/PISCES/"
NPD_DIR = "src/Models/AdvectedPopulations/NutrientsPlanktonDetritus/"
FILL_SOURCES = {
    "ObmNpdParams": {
        "bgc.plankton": [(NPD_DIR + "plankton.jl", "PhytoZoo")],
        "bgc.nutrients": [(NPD_DIR + "nutrients.jl", "NitrateAmmonia"), (NPD_DIR + "nutrients.jl", "NitrateAmmoniaIron")],
        "bgc.detritus": [(NPD_DIR + "detritus.jl", "TwoParticleAndDissolved"), (NPD_DIR + "detritus.jl", "VariableRedfieldDetritus"),
                         (NPD_DIR + "detritus.jl", "Detritus")],
        "bgc.oxygen": [(NPD_DIR + "oxygen.jl", "Oxygen")]},
    "ObmPiscesParams": {
        "bgc": [(PISCES_DIR + "PISCES.jl", "PISCES")],
        "bgc.phytoplankton": [(PISCES_DIR + "phytoplankton/nano_and_diatoms.jl", "NanoAndDiatoms")],
        "bgc.phytoplankton.nano": [(PISCES_DIR + "phytoplankton/mixed_mondo.jl", "MixedMondo")],
        "bgc.phytoplankton.diatoms": [(PISCES_DIR + "phytoplankton/mixed_mondo.jl", "MixedMondo")],
        "bgc.phytoplankton.nano.growth_rate": [(PISCES_DIR + "phytoplankton/growth_rate.jl", "GrowthRespirationLimitedProduction"),
                                               (PISCES_DIR + "phytoplankton/growth_rate.jl", "NutrientLimitedProduction")],
        "bgc.phytoplankton.diatoms.growth_rate": [(PISCES_DIR + "phytoplankton/growth_rate.jl", "GrowthRespirationLimitedProduction"),
                                                  (PISCES_DIR + "phytoplankton/growth_rate.jl", "NutrientLimitedProduction")],
        "bgc.phytoplankton.nano.nutrient_limitation": [(PISCES_DIR + "phytoplankton/nutrient_limitation.jl", "NitrogenIronPhosphateSilicateLimitation")],
        "bgc.phytoplankton.diatoms.nutrient_limitation": [(PISCES_DIR + "phytoplankton/nutrient_limitation.jl", "NitrogenIronPhosphateSilicateLimitation")],
        "bgc.zooplankton": [(PISCES_DIR + "zooplankton/micro_and_meso.jl", "MicroAndMeso")],
        "bgc.zooplankton.micro": [(PISCES_DIR + "zooplankton/food_quality_dependant.jl", "QualityDependantZooplankton")],
        "bgc.zooplankton.meso": [(PISCES_DIR + "zooplankton/food_quality_dependant.jl", "QualityDependantZooplankton")],
        "bgc.dissolved_organic_matter": [(PISCES_DIR + "dissolved_organic_matter/dissolved_organic_carbon.jl", "DissolvedOrganicCarbon")],
        "bgc.particulate_organic_matter": [(PISCES_DIR + "particulate_organic_matter/two_size_class.jl", "TwoCompartmentCarbonIronParticles")],
        "bgc.nitrogen": [(PISCES_DIR + "nitrogen/nitrate_ammonia.jl", "NitrateAmmonia")],
        "bgc.iron": [(PISCES_DIR + "iron/simple_iron.jl", "SimpleIron")],
        "bgc.oxygen": [(PISCES_DIR + "oxygen.jl", "Oxygen")],
        "bgc.latitude": [(PISCES_DIR + "common.jl", "PrescribedLatitude")]},
    "ObmTwobandParams": {"par": [("src/Light/2band.jl", "TwoBandPhotosyntheticallyActiveRadiation")]},
    "ObmMultibandParams": {"par": [("src/Light/multi_band.jl", "MultiBandPhotosyntheticallyActiveRadiation")]},
    "ObmSedimentParams": {"sed.biogeochemistry": [("src/Models/Sediments/simple_multi_G.jl", "SimpleMultiG"),
                                                  ("src/Models/Sediments/instant_remineralisation.jl", "InstantRemineralisation")]},
}
# reference fields that are NOT parameters of a kernel, with where they go instead
NOT_PARAMETERS = {
    "PhytoZoo": {"phytoplankton_sinking_velocity", "zooplankton_sinking_velocity"},      # → biogeochemical_drift_velocity (w fields)
    "TwoParticleAndDissolved": {"small_particle_sinking_velocity", "large_particle_sinking_velocity"},
    "VariableRedfieldDetritus": {"small_particle_sinking_velocity", "large_particle_sinking_velocity"},
    "Detritus": {"sinking_speeds"},
    "MixedMondo": {"growth_rate", "nutrient_limitation"},                                 # sub-structs, consumed field by field
    "NanoAndDiatoms": {"nano", "diatoms"},
    "MicroAndMeso": {"micro", "meso"},
    "QualityDependantZooplankton": {"food_preferences"},                                   # NamedTuple, consumed key by key
    "DissolvedOrganicCarbon": {"bacteria_concentration_depth_exponent"},                   # unused by the reference's DOC methods (the zooplankton's is)
    "PISCES": {"phytoplankton", "zooplankton", "dissolved_organic_matter", "particulate_organic_matter", "nitrogen", "iron",
               "silicate", "oxygen", "phosphate", "inorganic_carbon", "latitude", "day_length",   # components; day length: two host-evaluated calls
               "mixed_layer_depth", "euphotic_depth", "mean_mixed_layer_vertical_diffusivity", "mean_mixed_layer_light",
               "carbon_chemistry", "calcite_saturation", "sinking_velocities"},           # fields → ObmPiscesFields / obm_calcite_saturation
    "TwoBandPhotosyntheticallyActiveRadiation": {"field", "surface_PAR"},                  # arrays / callables → kernel arguments
    "MultiBandPhotosyntheticallyActiveRadiation": {"total", "fields", "field_names", "surface_PAR"},
    "SimpleMultiG": {"sinking_nitrogen", "sinking_carbon"},                                # tracer-name tuples → counts + pointer tables
    "InstantRemineralisation": {"sinking_tracers", "remineralisation_reciever"},
}


def reference_struct_fields(path, name):
    """Field names of `struct name{…}` in a reference source file (up to its inner constructor or `end`)."""
    src = open(os.path.join(REFERENCE, path), encoding="utf-8").read()
    m = re.search(r"^\s*(?:@kwdef\s+)?struct\s+" + name + r"\b[^\n]*\n(.*?)^\s*end\b", src, flags=re.S | re.M)
    assert m, (path, name)
    fields = []
    for line in m.group(1).splitlines():
        line = line.split("#")[0]
        if re.match(r"\s*(function\b|" + name + r"\b)", line):
            break
        f = re.match(r"\s*([A-Za-z_][\w′]*)\s*::", line)
        if f:
            fields.append(f.group(1))
    return fields


def fill_functions():
    src = open(os.path.join(ROOT, "julia", "obm_fill.jl"), encoding="utf-8").read()
    out = {}
    for m in re.finditer(r"^(Obm\w+)\(([^)]*)\) = \1\(;\n(.*?)^\)\n", src, flags=re.S | re.M):
        out[m.group(1)] = m.group(3)
    return out


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree is only in the build container")
def test_julia_fill_constructors_consume_every_reference_field_once():
    """julia/obm_fill.jl against the reference's own struct definitions: every field a fill constructor reads exists in
    the reference struct it is read from; every field of those structs is consumed exactly once per place it occurs (or is
    listed in NOT_PARAMETERS with where it goes instead); every member of the C struct is assigned."""
    fills = fill_functions()
    assert set(FILL_SOURCES) <= set(fills), sorted(fills)
    for fn, sources in FILL_SOURCES.items():
        body = fills[fn]
        reads = re.findall(r"(?:fieldor|tupleor)\(([\w.]+), :([\w′]+)(?:, \d+)?\)", body) + \
            re.findall(r"bandor\((par)\.(\w+), \d+\)", body)
        by_parent = {}
        for parent, field in reads:
            by_parent.setdefault(parent, []).append(field)
        # fields consumed through a hand-written expression (an enumeration from a type, a flag, the day-length calls)
        manual = {}
        for parent in sources:
            for field in re.findall(r"(?<![\w.])" + re.escape(parent) + r"\.([A-Za-z_]\w*)(?![\w.])", body):
                manual.setdefault(parent, set()).add(field)
        for parent, fields in by_parent.items():
            if parent.endswith(".food_preferences"):
                assert sorted(set(fields)) == ["D", "P", "POC", "Z"], (fn, parent, fields)  # defaults.jl:4,12 NamedTuple keys
                continue
            assert parent in sources, f"{fn}: reads from {parent}, which has no reference struct listed"
            known = set()
            for path, name in sources[parent]:
                known |= set(reference_struct_fields(path, name))
            assert set(fields) <= known, f"{fn}: {parent} has no field(s) {sorted(set(fields) - known)} in {sources[parent]}"
        for parent, structs in sources.items():
            got = by_parent.get(parent, [])
            for path, name in structs:
                want = [f for f in reference_struct_fields(path, name) if f not in NOT_PARAMETERS.get(name, ())]
                assert want, (path, name)
                for f in want:
                    n = got.count(f)
                    if n == 0 and f in manual.get(parent, ()):
                        continue
                    # a scalar field is read once; a tuple field once per element
                    assert n >= 1, f"{fn}: reference field {name}.{f} ({path}) is never consumed from {parent}"
                    if n > 1:
                        idx = re.findall(r"(?:tupleor\(" + re.escape(parent) + r", :" + f + r", (\d+)\)|bandor\(par\." + f + r", (\d+)\))", body)
                        assert len(idx) == n and len(set(idx)) == n, f"{fn}: {name}.{f} consumed {n} times"
    # every member of every C struct with a fill constructor is assigned, by its own name, exactly once
    for fn, body in fills.items():
        cname = "obm_" + re.sub(r"(?<!^)([A-Z])", r"_\1", fn[3:]).lower()
        cls = _lib.STRUCTS[cname]
        top = re.findall(r"^    (\w+) = ", body, flags=re.M)
        assert top == [n for n, _ in cls._fields_], (fn, top)


def test_generated_julia_fill_constructors_are_in_sync():
    """julia/obm_fill.jl is generated from the Python host mirror's own c_params(); regenerate it when a mirror changes."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_julia_fill.py")], capture_output=True, text=True,
                         check=True).stdout
    assert out == open(os.path.join(ROOT, "julia", "obm_fill.jl"), encoding="utf-8").read()
