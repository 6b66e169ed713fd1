"""ORACLE shim (test infrastructure): re-exports oracle.sd15 under the module paths the reference imports."""
from oracle.sd15 import (DDIMScheduler, DDPMScheduler, DPMSolverMultistepScheduler,
                         DPMSolverMultistepInverseScheduler, StableDiffusionPipeline, AutoencoderKL,
                         UNet2DConditionModel)
from . import pipelines, models, schedulers, configuration_utils  # noqa: F401
__version__ = "0.21.1+oracle-shim"
