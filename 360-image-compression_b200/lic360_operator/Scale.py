"""Import-path compatibility: `from lic360_operator.Scale import ...` (reference lic360_operator/Scale.py)."""
from ._modules import Scale  # noqa: F401
