"""redback_b200: B200-native batched light-curve + Gaussian-likelihood engine behind redback's model /
likelihood API.  The CUDA library is the only compute path; importing the package does
not need a GPU, calling it does."""
from . import ejecta_relations  # noqa: F401
from . import (_capi, afterglow_models, bands, engine, extinction_models, kilonova_models, likelihoods,  # noqa: F401
               magnetar_driven_ejecta_models, phase_models, simulate, supernova_models, tde_models)
from .likelihoods import (GaussianLikelihood, GaussianLikelihoodQuadratureNoise,  # noqa: F401
                          GaussianLikelihoodWithFractionalNoise, GaussianLikelihoodWithSystematicNoise,
                          GaussianLikelihoodWithUpperLimits, MixtureGaussianLikelihood)

__version__ = "0.1.0"

# redback.model_library.all_models_dict equivalent (model_library.py:20-34) for the fused models
all_models_dict = {
    name: getattr(supernova_models, name)
    for name in ("arnett_bolometric", "arnett", "basic_magnetar_powered_bolometric", "basic_magnetar_powered",
                 "slsn_bolometric", "slsn", "type_1a", "type_1c", "magnetar_nickel", "csm_interaction_bolometric",
                 "csm_interaction", "general_magnetar_slsn_bolometric", "general_magnetar_slsn",
                 "exponential_powerlaw_bolometric", "sn_exponential_powerlaw", "sn_fallback", "sn_nickel_fallback",
                 "homologous_expansion_supernova_model_bolometric", "homologous_expansion_supernova",
                 "thin_shell_supernova_model_bolometric", "thin_shell_supernova",
                 "general_magnetar_driven_supernova_bolometric", "general_magnetar_driven_supernova")
}
all_models_dict.update({name: getattr(kilonova_models, name)
                        for name in ("one_component_kilonova_model", "metzger_kilonova_model",
                                     "two_component_kilonova_model", "three_component_kilonova_model",
                                     "one_component_ejecta_relation", "one_component_ejecta_relation_projection",
                                     "one_component_nsbh_ejecta_relation", "two_component_bns_ejecta_relation",
                                     "two_component_nsbh_ejecta_relation")})
all_models_dict.update({fn.__name__: fn for fn in afterglow_models.FUSED})
all_models_dict.update({fn.__name__: fn for fn in magnetar_driven_ejecta_models.FUSED})
all_models_dict.update({fn.__name__: fn for fn in tde_models.FUSED})
all_models_dict["t0_base_model"] = phase_models.t0_base_model
all_models_dict.update({fn.__name__: fn for fn in extinction_models.MODELS + phase_models.EXTINCTION_MODELS})
