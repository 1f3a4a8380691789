"""Import-path compatibility: `from lic360_operator.ContextShift import ...` (reference lic360_operator/ContextShift.py)."""
from ._modules import ContextShift  # noqa: F401
