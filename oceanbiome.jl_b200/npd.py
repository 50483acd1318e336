"""Nutrients–Plankton–Detritus family: `NPZD`, `LOBSTER`, `NutrientsPlanktonDetritus` and their
components — the host-side mirror of
src/Models/AdvectedPopulations/NutrientsPlanktonDetritus/ (same names, keyword arguments and
defaults).  The arithmetic lives in csrc/npd_tendencies.cu behind `obm_npd_tendencies`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import _lib
from .grids import RectilinearGrid, current_stream_ptr, require_cuda
from .light import TwoBandPhotosyntheticallyActiveRadiation, default_surface_PAR

day = 86400.0


# ---- formulations (plankton.jl:83-84, 216-217) -----------------------------------------------
class Linear:
    pass


class Quadratic:
    pass


class MondoLightLimitation:
    pass


class AnalyticalLightLimitation:
    pass


# ---- nutrients (nutrients.jl) ---------------------------------------------------------------------
@dataclass
class NitrateAmmonia:
    nitrification_rate: float = 5.8e-7  # 1/s


@dataclass
class NitrateAmmoniaIron:
    nitrification_rate: float = 5.8e-7


@dataclass
class Nutrient:
    pass


# ---- plankton (plankton.jl:19-58) --------------------------------------------------------------
@dataclass
class PhytoZoo:
    nitrate_half_saturation: float = 0.7
    ammonia_half_saturation: float = 0.001
    iron_half_saturation: float = 2e-4
    nitrate_ammonia_inhibition: float = 3.0
    light_half_saturation: float = 33.0
    phytoplankton_maximum_growth_rate: float = 2.42e-5
    iron_ratio: float = 4.6375e-5
    phytoplankton_exudation_fraction: float = 0.05
    ammonia_fraction_of_exudate: float = 0.75
    light_limitation: object = field(default_factory=MondoLightLimitation)
    temperature_coefficient: Optional[float] = None
    phytoplankton_mortality_rate: float = 5.8e-7
    zooplankton_mortality_rate: float = 2.31e-6
    zooplankton_excretion_rate: float = 5.8e-7
    phytoplankton_mortality_formulation: object = field(default_factory=Quadratic)
    phytoplankton_solid_waste_fraction: float = 1.0
    excretion_inorganic_fraction: float = 0.5
    preference_for_phytoplankton: float = 0.5
    maximum_grazing_rate: float = 9.26e-6
    grazing_half_saturation: float = 1.0
    zooplankton_assimilation_fraction: float = 0.7
    grazing_concentration_formulation: object = field(default_factory=Quadratic)
    zooplankton_calcite_dissolution: float = 0.3
    redfield_ratio: float = 6.56
    carbon_calcite_ratio: float = 0.1
    zooplankton_gut_calcite_dissolution: float = 0.3
    phytoplankton_chlorophyll_ratio: float = 1.31
    # sinking speeds are consumed by the host model's advection (biogeochemical_drift_velocity),
    # not by the tendency kernels (plankton.jl:56-76)
    phytoplankton_sinking_speed: float = 0.0
    zooplankton_sinking_speed: float = 0.0
    open_bottom: bool = True  # `PhytoZoo(grid; …, open_bottom = true)` plankton.jl:65-76: False tapers w to 0 at the bottom face


# ---- detritus (detritus.jl) --------------------------------------------------------------------------
@dataclass
class TwoParticleAndDissolved:
    remineralisation_inorganic_fraction: float = 0.0
    small_remineralisation_rate: float = 5.88e-7
    large_remineralisation_rate: float = 5.88e-7
    dissolved_remineralisation_rate: float = 3.86e-7
    small_solid_waste_fraction: float = 0.5
    redfield_ratio: float = 6.56
    small_particle_sinking_speed: float = 3.47e-5  # m/s (w = -speed)
    large_particle_sinking_speed: float = 200 / day
    open_bottom: bool = True  # `TwoParticleAndDissolved(grid; …, open_bottom = true)` detritus.jl:107-123


@dataclass
class VariableRedfieldDetritus:
    remineralisation_inorganic_fraction: float = 0.0
    small_remineralisation_rate: float = 5.88e-7
    large_remineralisation_rate: float = 5.88e-7
    dissolved_remineralisation_rate: float = 3.86e-7
    small_solid_waste_fraction: float = 0.5
    small_particle_sinking_speed: float = 3.47e-5
    large_particle_sinking_speed: float = 200 / day
    open_bottom: bool = True


@dataclass
class Detritus:
    remineralisation_rate: float = 0.1213 / day
    small_particle_fraction: float = 0.5
    redfield_ratio: float = 6.56
    sinking_speed: float = 2.7489 / day
    open_bottom: bool = True  # `Detritus(grid; sinking_speed, open_bottom = true)` detritus.jl:274-283


# ---- carbonate system / oxygen (carbonate_system.jl:39-46, oxygen.jl:14-17) ------------
@dataclass
class CarbonateSystem:
    replicates: int = 1


@dataclass
class Oxygen:
    respiration_oxygen_nitrogen_ratio: float = 10.75
    nitrification_oxygen_nitrogen_ratio: float = 2.0


class NutrientsPlanktonDetritus:
    """The underlying biogeochemistry `NutrientsPlanktonDetritus{NUT, PLA, DET, CAR, OXY}`
    (NutrientsPlanktonDetritus.jl:27-33)."""

    def __init__(self, nutrients=None, plankton=None, detritus=None, carbonate_system=None, oxygen=None):
        if nutrients is None or plankton is None:
            raise ValueError("nutrients and plankton components are required")
        if isinstance(carbonate_system, int):
            carbonate_system = CarbonateSystem(carbonate_system)
        self.nutrients, self.plankton, self.detritus = nutrients, plankton, detritus
        self.carbonate_system, self.oxygen = carbonate_system, oxygen
        self.parameter_ensemble = None

    # -- parameter-sweep ensembles ---------------------------------------------------
    @staticmethod
    def parameter_index(name: str) -> int:
        """Position of `name` among the double members of `obm_npd_params` (= `obm_npd_param_index`)."""
        doubles = [n for n, t in _lib.obm_npd_params._fields_ if t is C.c_double]
        if name not in doubles:
            raise KeyError(f"obm_npd_params has no double member '{name}'; choose from {doubles}")
        return doubles.index(name)

    def set_parameter_ensemble(self, device=None, **values):
        """One value per ensemble member for each named parameter (`obm_npd_params` member names); a member is a
        horizontal column of the grid (box i of `BoxModelGrid(n)`).  The reference builds and runs one model per
        parameter vector (examples/data_assimilation.jl:26-52); here all members share every launch.  Structural
        choices (components, formulations, whether T is a tracer) are common to the ensemble.  No arguments: back
        to one parameter set."""
        if not values:
            self.parameter_ensemble = None
            return self
        if len(values) > _lib.OBM_NPD_MAX_VARIED:
            raise ValueError(f"at most {_lib.OBM_NPD_MAX_VARIED} varied parameters")
        which = [self.parameter_index(n) for n in values]
        cols = [torch.as_tensor(v, dtype=torch.float64).reshape(-1) for v in values.values()]
        if len({c.numel() for c in cols}) != 1:
            raise ValueError("every varied parameter needs one value per member")
        table = torch.stack(cols).contiguous()
        if device is not None:
            table = table.to(device)
        self.parameter_ensemble = ((C.c_int32 * len(which))(*which), table, tuple(values))
        return self

    # -- C parameter block -------------------------------------------------------------------------
    def c_params(self) -> _lib.obm_npd_params:
        p = _lib.obm_npd_params()
        pl, nu, de = self.plankton, self.nutrients, self.detritus
        p.nutrients = {Nutrient: _lib.NUT_NUTRIENT, NitrateAmmonia: _lib.NUT_NITRATE_AMMONIA,
                       NitrateAmmoniaIron: _lib.NUT_NITRATE_AMMONIA_IRON}[type(nu)]
        p.detritus = {type(None): _lib.DET_NONE, Detritus: _lib.DET_DETRITUS,
                      TwoParticleAndDissolved: _lib.DET_TWO_PARTICLE,
                      VariableRedfieldDetritus: _lib.DET_VARIABLE_REDFIELD}[type(de)]
        p.carbonate_replicates = self.carbonate_system.replicates if self.carbonate_system else 0
        p.oxygen = 1 if self.oxygen else 0
        p.light_limitation = _lib.LIGHT_MONDO if isinstance(pl.light_limitation, MondoLightLimitation) else _lib.LIGHT_ANALYTICAL
        p.phytoplankton_mortality_formulation = _lib.LINEAR if isinstance(pl.phytoplankton_mortality_formulation, Linear) else _lib.QUADRATIC
        p.grazing_concentration_formulation = _lib.LINEAR if isinstance(pl.grazing_concentration_formulation, Linear) else _lib.QUADRATIC
        p.has_temperature_coefficient = 0 if pl.temperature_coefficient is None else 1
        for name in ("nitrate_half_saturation", "ammonia_half_saturation", "iron_half_saturation",
                     "nitrate_ammonia_inhibition", "light_half_saturation", "phytoplankton_maximum_growth_rate",
                     "iron_ratio", "phytoplankton_exudation_fraction", "ammonia_fraction_of_exudate",
                     "phytoplankton_mortality_rate", "zooplankton_mortality_rate", "zooplankton_excretion_rate",
                     "phytoplankton_solid_waste_fraction", "excretion_inorganic_fraction",
                     "preference_for_phytoplankton", "maximum_grazing_rate", "grazing_half_saturation",
                     "zooplankton_assimilation_fraction", "zooplankton_calcite_dissolution", "redfield_ratio",
                     "carbon_calcite_ratio", "zooplankton_gut_calcite_dissolution",
                     "phytoplankton_chlorophyll_ratio"):
            setattr(p, name, float(getattr(pl, name)))
        p.temperature_coefficient = float(pl.temperature_coefficient or 0.0)
        p.nitrification_rate = float(getattr(nu, "nitrification_rate", 0.0))
        if isinstance(de, (TwoParticleAndDissolved, VariableRedfieldDetritus)):
            for name in ("remineralisation_inorganic_fraction", "small_remineralisation_rate",
                         "large_remineralisation_rate", "dissolved_remineralisation_rate",
                         "small_solid_waste_fraction"):
                setattr(p, name, float(getattr(de, name)))
            p.detritus_redfield_ratio = float(getattr(de, "redfield_ratio", 0.0))
        elif isinstance(de, Detritus):
            p.remineralisation_rate = float(de.remineralisation_rate)
            p.small_particle_fraction = float(de.small_particle_fraction)
            p.detritus_redfield_ratio = float(de.redfield_ratio)
        if self.oxygen:
            p.respiration_oxygen_nitrogen_ratio = float(self.oxygen.respiration_oxygen_nitrogen_ratio)
            p.nitrification_oxygen_nitrogen_ratio = float(self.oxygen.nitrification_oxygen_nitrogen_ratio)
        return p

    # -- plugin surface ---------------------------------------------------------------------------
    def required_biogeochemical_tracers(self):
        """NutrientsPlanktonDetritus.jl:69-74 — derived host-side, cross-checked against the
        library's obm_npd_tracer_names in the tests."""
        t = []
        t += {Nutrient: ["N"], NitrateAmmonia: ["NO₃", "NH₄"], NitrateAmmoniaIron: ["NO₃", "NH₄", "Fe"]}[type(self.nutrients)]
        t += ["P", "Z"] + ([] if self.plankton.temperature_coefficient is None else ["T"])
        t += {type(None): [], Detritus: ["D"], TwoParticleAndDissolved: ["sPOM", "bPOM", "DOM"],
              VariableRedfieldDetritus: ["sPOC", "bPOC", "DOC", "sPON", "bPON", "DON"]}[type(self.detritus)]
        if self.carbonate_system:
            N = self.carbonate_system.replicates
            t += ["DIC", "Alk"] if N == 1 else [f"DIC{n}" for n in range(1, N + 1)] + [f"Alk{n}" for n in range(1, N + 1)]
        if self.oxygen:
            t += ["O₂"]
        return tuple(t)

    def required_biogeochemical_auxiliary_fields(self):
        return ("PAR",)  # plankton.jl:81

    def biogeochemical_auxiliary_fields(self):
        return {}

    def biogeochemical_drift_velocity(self, name):
        """w of the sinking tracers (plankton.jl:60-63, detritus.jl:135-141,290-291); None = no sinking."""
        pl, de = self.plankton, self.detritus
        if name == "P" and pl.phytoplankton_sinking_speed:
            return -pl.phytoplankton_sinking_speed
        if name == "Z" and pl.zooplankton_sinking_speed:
            return -pl.zooplankton_sinking_speed
        if isinstance(de, (TwoParticleAndDissolved, VariableRedfieldDetritus)):
            if name in ("sPOM", "sPON", "sPOC"):
                return -de.small_particle_sinking_speed
            if name in ("bPOM", "bPON", "bPOC"):
                return -de.large_particle_sinking_speed
        if isinstance(de, Detritus) and name == "D":
            return -de.sinking_speed
        return None

    def drift_velocity_open_bottom(self, name) -> bool:
        """Whether the drift velocity of `name` keeps its full speed at the bottom face (`open_bottom = true`, the default)
        or is tapered to zero there, w·(1 − e^{(1−k)/2}) (`setup_velocity_fields`, sinking_velocity_fields.jl:15-17)."""
        owner = self.plankton if name in ("P", "Z") else self.detritus
        return bool(getattr(owner, "open_bottom", True))

    def conserved_tracers(self, labeled=False):
        """coupling_utils.jl:1-52 — nitrogen group, plus the carbon group (with scale factors)
        when a CarbonateSystem is present."""
        tracers = ["P", "Z"]
        tracers += ["N"] if isinstance(self.nutrients, Nutrient) else ["NO₃", "NH₄"]
        if isinstance(self.detritus, VariableRedfieldDetritus):
            tracers += ["sPON", "bPON", "DON"]
        elif isinstance(self.detritus, TwoParticleAndDissolved):
            tracers += ["sPOM", "bPOM", "DOM"]
        else:
            tracers += ["D"]
        nitrogen = tuple(tracers)
        if not self.carbonate_system:
            return {"nitrogen": nitrogen} if labeled else nitrogen
        R, rho = self.plankton.redfield_ratio, self.plankton.carbon_calcite_ratio
        ct, cs = ["P", "Z", "DIC"], [(1 + rho) * R, R, 1]
        if isinstance(self.detritus, VariableRedfieldDetritus):
            ct += ["sPOC", "bPOC", "DOC"]
            cs += [1, 1, 1]
        elif isinstance(self.detritus, TwoParticleAndDissolved):
            ct += ["sPOM", "bPOM", "DOM"]
            cs += [R, R, R]
        else:
            ct += ["D"]
            cs += [R]
        carbon = {"tracers": tuple(ct), "scalefactors": tuple(cs)}
        return {"nitrogen": nitrogen, "carbon": carbon} if labeled else (nitrogen, carbon)

    def chlorophyll(self, model):
        """coupling_utils.jl:54 — (chl_a, chl_b, scale) for the multi-band light model."""
        return model.tracers["P"], None, self.plankton.phytoplankton_chlorophyll_ratio

    def update_biogeochemical_state(self, model, stream=None):
        return None  # no method in the reference: Oceananigans' no-op fallback

    # -- the fused tendency pass ---------------------------------------------------------------------
    def compute_tendencies(self, grid: RectilinearGrid, tracers: dict, auxiliary_fields: dict, G: dict,
                           accumulate: bool = True, stream: Optional[int] = None, time: float = 0.0):
        """All per-tracer callables `bgc(i, j, k, grid, Val(name), clock, fields, aux)` for every
        cell in one launch; G[name] (+)= tendency."""
        names = self.required_biogeochemical_tracers()
        PAR = auxiliary_fields["PAR"]
        require_cuda(PAR, *[tracers[n] for n in names])
        lib = _lib.load()
        cg = grid.c_grid()
        p = self.c_params()
        tptr = _lib.pointer_table([tracers[n].ptr for n in names])
        gptr = _lib.pointer_table([G[n].ptr if (n in G and G[n] is not None and n != "T") else None for n in names])
        s = stream if stream is not None else current_stream_ptr(grid.device)
        if self.parameter_ensemble is None:
            rc = lib.obm_npd_tendencies(C.byref(cg), C.byref(p), tptr, PAR.ptr, gptr, 1 if accumulate else 0, s)
            _lib.check(rc, "obm_npd_tendencies")
            return
        which, table, _ = self.parameter_ensemble
        members = grid.Nx * grid.Ny
        if table.shape[1] != members:
            raise ValueError(f"parameter ensemble has {table.shape[1]} members, the grid has {members} columns")
        if table.device != PAR.data.device:
            table = table.to(PAR.data.device)
            self.parameter_ensemble = (which, table, self.parameter_ensemble[2])
        rc = lib.obm_npd_tendencies_ensemble(C.byref(cg), C.byref(p), len(which), which, table.data_ptr(), tptr, PAR.ptr,
                                             gptr, 1 if accumulate else 0, s)
        _lib.check(rc, "obm_npd_tendencies_ensemble")

    def compute_tendencies_and_substep(self, grid: RectilinearGrid, tracers: dict, auxiliary_fields: dict, Gm: dict,
                                       dt: float, gamma: float, zeta: Optional[float], G: Optional[dict] = None,
                                       accumulate: bool = False, store_Gn: bool = False, stream: Optional[int] = None):
        """f-2: `compute_tendencies!` + `rk3_substep!` + `cache_previous_tendencies!` (src/BoxModel/timesteppers.jl:30-93)
        of every tracer in ONE launch — U += Δt(γG + ζG⁻), G⁻ ← G with G evaluated from the cell's tracers in the same
        thread (plus the forcing found in G[name] when `accumulate`).  A tracer without an entry in `Gm` is not stepped.
        Bit-identical to `compute_tendencies` followed by `obm_rk3_substep`."""
        names = self.required_biogeochemical_tracers()
        PAR = auxiliary_fields["PAR"]
        require_cuda(PAR, *[tracers[n] for n in names])
        lib = _lib.load()
        cg = grid.c_grid()
        p = self.c_params()
        tptr = _lib.pointer_table([tracers[n].ptr for n in names])
        mptr = _lib.pointer_table([Gm[n].ptr if (n in Gm and Gm[n] is not None and n != "T") else None for n in names])
        gptr = None
        if G is not None:
            gptr = _lib.pointer_table([G[n].ptr if (n in G and G[n] is not None and n != "T") else None for n in names])
        elif accumulate or store_Gn:
            raise ValueError("accumulate / store_Gn need the Gⁿ fields")
        s = stream if stream is not None else current_stream_ptr(grid.device)
        nvary, which, values = 0, None, None
        if self.parameter_ensemble is not None:
            which, table, _ = self.parameter_ensemble
            members = grid.Nx * grid.Ny
            if table.shape[1] != members:
                raise ValueError(f"parameter ensemble has {table.shape[1]} members, the grid has {members} columns")
            if table.device != PAR.data.device:
                table = table.to(PAR.data.device)
                self.parameter_ensemble = (which, table, self.parameter_ensemble[2])
            nvary, values = len(which), table.data_ptr()
        rc = lib.obm_npd_tendencies_substep(C.byref(cg), C.byref(p), nvary, which, values, tptr, PAR.ptr, gptr,
                                            1 if accumulate else 0, 1 if store_Gn else 0, mptr, float(dt), float(gamma),
                                            0.0 if zeta is None else float(zeta), int(zeta is not None), s)
        _lib.check(rc, "obm_npd_tendencies_substep")

    def run_boxes(self, grid: RectilinearGrid, tracers: dict, auxiliary_fields: dict, Gm: dict, dt: float, stages, steps: int,
                  PAR_table: torch.Tensor, T_table: Optional[torch.Tensor] = None, output_every: int = 0,
                  snapshots: Optional[dict] = None, stream: Optional[int] = None):
        """f-3: the whole run of a box-model ensemble in ONE launch (`obm_npd_box_run`): `steps` time steps of
        `stages` = ((γ, ζ or None), …), each stage exactly `compute_tendencies_and_substep`, with the prescribed PAR (and
        temperature) read from device tables of one row per stage — (rows, 1) shared by every box or (rows, n) — and
        `snapshots[name]` (n_outputs, n) filled every `output_every` steps.  Every tracer must have its G⁻ in `Gm`."""
        names = self.required_biogeochemical_tracers()
        PAR = auxiliary_fields["PAR"]
        require_cuda(PAR, *[tracers[n] for n in names])
        rows = steps * len(stages)
        for tab, what in ((PAR_table, "PAR"), (T_table, "T")):
            if tab is not None and (tab.dim() != 2 or tab.shape[0] != rows or tab.shape[1] not in (1, grid.Nx)
                                    or tab.dtype != torch.float64 or not tab.is_contiguous()):
                raise ValueError(f"{what} table must be a contiguous float64 tensor ({rows}, 1) or ({rows}, {grid.Nx})")
        lib = _lib.load()
        cg = grid.c_grid()
        p = self.c_params()
        tptr = _lib.pointer_table([tracers[n].ptr for n in names])
        mptr = _lib.pointer_table([Gm[n].ptr if (n in Gm and Gm[n] is not None and n != "T") else None for n in names])
        sptr = None
        if snapshots:
            sptr = _lib.pointer_table([snapshots[n].data_ptr() if n in snapshots else None for n in names])
        nvary, which, values = 0, None, None
        if self.parameter_ensemble is not None:
            which, table, _ = self.parameter_ensemble
            if table.shape[1] != grid.Nx * grid.Ny:
                raise ValueError(f"parameter ensemble has {table.shape[1]} members, the grid has {grid.Nx * grid.Ny} columns")
            if table.device != PAR.data.device:
                table = table.to(PAR.data.device)
                self.parameter_ensemble = (which, table, self.parameter_ensemble[2])
            nvary, values = len(which), table.data_ptr()
        gam = (C.c_double * len(stages))(*[float(g) for g, _ in stages])
        zet = (C.c_double * len(stages))(*[float("nan") if z is None else float(z) for _, z in stages])
        s = stream if stream is not None else current_stream_ptr(grid.device)
        rc = lib.obm_npd_box_run(C.byref(cg), C.byref(p), nvary, which, values, tptr, mptr, PAR.ptr, PAR_table.data_ptr(),
                                 int(PAR_table.shape[1] != 1), T_table.data_ptr() if T_table is not None else None,
                                 int(T_table is not None and T_table.shape[1] != 1), int(steps), len(stages), gam, zet,
                                 float(dt), int(output_every), sptr, s)
        _lib.check(rc, "obm_npd_box_run")

    def __call__(self, name: str, *, PAR, device="cuda", **tracers):
        """The per-tracer form `bgc(Val(name), x, y, z, t, tracers..., PAR)` of the plugin API
        (docs/src/model_implementation.md:34-75; the built-in models' discrete form
        `bgc(i, j, k, grid, Val(name), clock, fields, aux)`, NutrientsPlanktonDetritus.jl:88): the tendency of ONE tracer
        at the given state — scalars, or arrays of states evaluated side by side.  The fused kernel runs on a row of
        boxes and the named tendency is returned (a float for scalar input, else a device tensor), so any single
        tendency can be inspected without assembling a model."""
        from .grids import CenterField
        names = self.required_biogeochemical_tracers()
        if name not in names:
            raise KeyError(f"{name} is not a tracer of this model {names}")
        vals = {n: torch.as_tensor(tracers.get(n, 0.0), dtype=torch.float64).reshape(-1) for n in names}
        unknown = set(tracers) - set(names)
        if unknown:
            raise KeyError(f"unknown tracers {sorted(unknown)}; this model carries {names}")
        vals["PAR"] = torch.as_tensor(PAR, dtype=torch.float64).reshape(-1)
        n = max(v.numel() for v in vals.values())
        grid = RectilinearGrid(size=(n,), x=(0.0, float(n)), topology=("Periodic", "Flat", "Flat"), halo=(0,), device=device)
        f = {k: CenterField(grid, k).set(v.expand(n).reshape(1, 1, n)) for k, v in vals.items()}
        G = {k: CenterField(grid, "G" + k) for k in names}
        saved = self.parameter_ensemble  # a parameter sweep applies when one state per member is given
        if saved is not None and saved[1].shape[1] != n:
            self.parameter_ensemble = None
        try:
            self.compute_tendencies(grid, f, {"PAR": f["PAR"]}, G, accumulate=False)
        finally:
            self.parameter_ensemble = saved
        out = G[name].interior.reshape(-1)
        return out.item() if n == 1 and all(v.numel() == 1 for v in vals.values()) and not torch.is_tensor(PAR) else out

    def summary(self):
        kind = "NPZD" if isinstance(self.nutrients, Nutrient) else "LOBSTER"
        return f"{kind} model ({', '.join(':' + t for t in self.required_biogeochemical_tracers())})"


def _assemble(grid, underlying, light_attenuation, sediment, scale_negatives, invalid_fill_value, particles, modifiers):
    """NutrientsPlanktonDetritus(grid; …) — NutrientsPlanktonDetritus.jl:35-66."""
    from .biogeochemistry import Biogeochemistry
    from .negative_tracers import ScaleNegativeTracers

    if scale_negatives:
        scaler = ScaleNegativeTracers.from_biogeochemistry(underlying, grid, invalid_fill_value=invalid_fill_value)
        if modifiers is None:
            modifiers = scaler
        elif isinstance(modifiers, tuple):
            modifiers = (*modifiers, *(scaler if isinstance(scaler, tuple) else (scaler,)))
        else:
            modifiers = (modifiers, *(scaler if isinstance(scaler, tuple) else (scaler,)))
    return Biogeochemistry(underlying, light_attenuation=light_attenuation, sediment=sediment, particles=particles,
                           modifiers=modifiers)


_DEFAULT = object()


def LOBSTER(grid, nutrients=None, plankton=None, detritus=_DEFAULT, carbonate_system=None, oxygen=None,
            surface_photosynthetically_active_radiation=default_surface_PAR, light_attenuation=_DEFAULT,
            sediment=None, scale_negatives=False, invalid_fill_value=float("nan"), particles=None, modifiers=None,
            parameter_ensemble=None):
    """`LOBSTER(grid; …)` — constructors.jl:65-96.  `parameter_ensemble = {name: one value per column}` (not in the
    reference) makes every horizontal column its own parameter set, see `set_parameter_ensemble`."""
    nutrients = nutrients if nutrients is not None else NitrateAmmonia()
    plankton = plankton if plankton is not None else PhytoZoo()
    detritus = TwoParticleAndDissolved() if detritus is _DEFAULT else detritus
    if light_attenuation is _DEFAULT:
        light_attenuation = TwoBandPhotosyntheticallyActiveRadiation(
            grid=grid, surface_PAR=surface_photosynthetically_active_radiation)
    underlying = NutrientsPlanktonDetritus(nutrients, plankton, detritus, carbonate_system, oxygen)
    if parameter_ensemble:
        underlying.set_parameter_ensemble(device=grid.device, **parameter_ensemble)
    return _assemble(grid, underlying, light_attenuation, sediment, scale_negatives, invalid_fill_value, particles,
                     modifiers)


def NPZD(grid, nutrients=None, plankton=None, detritus=_DEFAULT, carbonate_system=None, oxygen=None,
         surface_photosynthetically_active_radiation=default_surface_PAR, light_attenuation=_DEFAULT,
         sediment=None, scale_negatives=False, invalid_fill_value=float("nan"), particles=None, modifiers=None,
         parameter_ensemble=None):
    """`NPZD(grid; …)` — constructors.jl:177-227 (Kuhn et al. 2015 parameters); `parameter_ensemble` as in `LOBSTER`."""
    nutrients = nutrients if nutrients is not None else Nutrient()
    if plankton is None:
        plankton = PhytoZoo(
            nitrate_half_saturation=2.3868,
            phytoplankton_maximum_growth_rate=0.6989 / day,
            phytoplankton_exudation_fraction=0.0,
            temperature_coefficient=1.88,
            phytoplankton_mortality_formulation=Linear(),
            phytoplankton_mortality_rate=(0.066 + 0.0101) / day,
            preference_for_phytoplankton=1.0,
            grazing_concentration_formulation=Quadratic(),
            grazing_half_saturation=0.5573,
            zooplankton_mortality_rate=0.3395 / day,
            zooplankton_excretion_rate=0.0102 / day,
            zooplankton_assimilation_fraction=0.9116,
            phytoplankton_sinking_speed=0.2551 / day,
            excretion_inorganic_fraction=1.0,
            phytoplankton_solid_waste_fraction=0.0101 / (0.066 + 0.0101),
            maximum_grazing_rate=2.1522 / day,
            light_limitation=AnalyticalLightLimitation(),
            light_half_saturation=(0.6989 / day) / (0.1953 / day))
    detritus = Detritus() if detritus is _DEFAULT else detritus
    if light_attenuation is _DEFAULT:
        light_attenuation = TwoBandPhotosyntheticallyActiveRadiation(
            grid=grid, surface_PAR=surface_photosynthetically_active_radiation)
    underlying = NutrientsPlanktonDetritus(nutrients, plankton, detritus, carbonate_system, oxygen)
    if parameter_ensemble:
        underlying.set_parameter_ensemble(device=grid.device, **parameter_ensemble)
    return _assemble(grid, underlying, light_attenuation, sediment, scale_negatives, invalid_fill_value, particles,
                     modifiers)
