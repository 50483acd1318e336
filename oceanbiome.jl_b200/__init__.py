"""oceanbiome.jl_b200 — B200-native biogeochemical hot path behind OceanBioME.jl's plugin surface.

Host-side mirror (Python) of the reference's operator interface for this path; all arithmetic is in
hand-written sm_100a CUDA kernels (csrc/) reached through the C ABI of include/obm_b200.h.
There is no CPU fallback: compute entry points raise if the CUDA library or a CUDA device is missing.
"""
from . import _lib
from ._lib import ObmError, load as load_library
from .grids import CenterField, ConstantField, Field, Field2D, LatitudeLongitudeGrid, RectilinearGrid, ZFaceField
from .light import (MultiBandPhotosyntheticallyActiveRadiation, PrescribedPhotosyntheticallyActiveRadiation,
                    TwoBandPhotosyntheticallyActiveRadiation, compute_euphotic_depth, compute_mixed_layer_mean,
                    default_surface_PAR)
from .carbon_chemistry import CarbonChemistry
from .negative_tracers import ScaleNegativeTracers, ZeroNegativeTracers
from .npd import (LOBSTER, NPZD, AnalyticalLightLimitation, CarbonateSystem, Detritus, Linear, MondoLightLimitation,
                  NitrateAmmonia, NitrateAmmoniaIron, Nutrient, NutrientsPlanktonDetritus, Oxygen, PhytoZoo, Quadratic,
                  TwoParticleAndDissolved, VariableRedfieldDetritus)
from . import pisces
from .pisces import PISCES, CBMDayLength, DepthDependantSinkingSpeed, ModelLatitude, PrescribedLatitude
from .sediments import (BiogeochemicalSediment, InstantRemineralisation, InstantRemineralisationSediment, SimpleMultiG,
                        SimpleMultiGSediment, calculate_bottom_indices)
from .gas_exchange import (CarbonDioxideConcentration, CarbonDioxideGasExchangeBoundaryCondition,
                           CarbonDioxidePolynomialSchmidtNumber, GasExchange, GasExchangeBoundaryCondition,
                           OxygenConcentration, OxygenGasExchangeBoundaryCondition, OxygenPolynomialSchmidtNumber,
                           PartiallySolubleGas, PolynomialParameterisation, SchmidtScaledTransferVelocity)
from .particles import BiogeochemicalParticles, LinearOptimalTemperatureRange, SugarKelp, SugarKelpParticles
from .biogeochemistry import Biogeochemistry, BiogeochemicalModel, Clock
from .box_model import BoxModel, BoxModelGrid, SpeedyOutput, load_output

__version__ = "0.1.0"
