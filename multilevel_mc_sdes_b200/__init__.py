"""multilevel_mc_sdes_b200 - B200-native per-level coupled sampler behind MLMC_RSCAM's
`Option.looper` / `Option.mlmc` contract.  See DESIGN.md; C ABI in include/mlmcb.h."""
from .options import (Option, GBM_Option, Merton_Option, Euro_GBM, Asian_GBM, Lookback_GBM, Digital_GBM,
                      Euro_Merton, Asian_Merton, Lookback_Merton, Digital_Merton, Amer_GBM, sqrt_boundary,
                      JumpDiffusion_Option, Diffusion_Option, MyOption, mlmc_driver)
from .giles import giles_analysis
from . import _lib

__all__ = ["Option", "GBM_Option", "Merton_Option", "Euro_GBM", "Asian_GBM", "Lookback_GBM", "Digital_GBM",
           "Euro_Merton", "Asian_Merton", "Lookback_Merton", "Digital_Merton", "Amer_GBM", "sqrt_boundary",
           "JumpDiffusion_Option", "Diffusion_Option", "MyOption", "mlmc_driver", "giles_analysis", "_lib"]
