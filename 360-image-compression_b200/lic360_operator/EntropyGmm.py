"""Import-path compatibility: `from lic360_operator.EntropyGmm import ...` (reference lic360_operator/EntropyGmm.py)."""
from ._modules import EntropyGmm  # noqa: F401
