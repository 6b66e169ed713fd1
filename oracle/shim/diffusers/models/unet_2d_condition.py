"""ORACLE shim (test infrastructure): re-exports oracle.sd15 under the module paths the reference imports."""
from oracle.sd15 import UNet2DConditionOutput, UNet2DConditionModel  # noqa: F401
