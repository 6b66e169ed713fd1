"""ORACLE shim (test infrastructure): re-exports oracle.sd15 under the module paths the reference imports."""
from . import unet_2d_condition, attention, attention_processor, resnet  # noqa: F401
