"""ORACLE shim (test infrastructure): re-exports oracle.sd15 under the module paths the reference imports."""
from . import display  # noqa: F401
