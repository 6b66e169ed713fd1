"""ORACLE shim (test infrastructure): re-exports oracle.sd15 under the module paths the reference imports."""
from . import pipeline_stable_diffusion, pipeline_stable_diffusion_pix2pix_zero  # noqa: F401
