"""Import-path compatibility: `from lic360_operator.Imp2mask import ...` (reference lic360_operator/Imp2mask.py)."""
from ._modules import Imp2mask  # noqa: F401
