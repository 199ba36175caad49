"""tinyedm_b200 — B200-native (sm_100a) implementation of the tinyedm denoiser hot path.

Same public names as the reference package (`/root/reference/src/tinyedm/__init__.py:1-9`) for the parts that
are in scope: the module surfaces are unchanged, the arithmetic underneath is hand-written CUDA
reached through the C ABI in include/tinyedm_b200.h. There is no CPU path and no fallback.
"""
from .edm import EDM, Diffuser
from .graphs import GraphedTrainStep
from .metric import WeightedMeanSquaredError, fused_edm_loss
from .networks import Conv2d, Denoiser, DenoiserWrapper, Embedding, Linear, UncertaintyNet
from .ops import to_uint8_images
from .optim import FusedAdamEMA, sigma_rel_to_gamma
from .solvers import DeterministicSolver
from .utils import deinstantiate, instantiate, load_reference_checkpoint, swap_tensors

__all__ = ["EDM", "Diffuser", "DeterministicSolver", "WeightedMeanSquaredError", "Denoiser", "Linear", "Conv2d",
           "Embedding", "DenoiserWrapper", "UncertaintyNet", "FusedAdamEMA", "sigma_rel_to_gamma", "fused_edm_loss", "GraphedTrainStep", "deinstantiate", "instantiate", "swap_tensors", "load_reference_checkpoint", "to_uint8_images"]
