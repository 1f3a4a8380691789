"""Import-path compatibility: `from lic360_operator.SpherePad import ...` (reference lic360_operator/SpherePad.py)."""
from ._modules import SpherePad  # noqa: F401
