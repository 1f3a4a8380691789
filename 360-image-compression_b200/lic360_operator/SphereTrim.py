"""Import-path compatibility: `from lic360_operator.SphereTrim import ...` (reference lic360_operator/SphereTrim.py)."""
from ._modules import SphereTrim  # noqa: F401
