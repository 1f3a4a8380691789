"""Import-path compatibility: `from lic360_operator.ImpMap import ...` (reference lic360_operator/ImpMap.py)."""
from ._modules import ImpMap  # noqa: F401
