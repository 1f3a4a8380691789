"""Import-path compatibility: `from lic360_operator.QUANT import ...` (reference lic360_operator/QUANT.py)."""
from ._modules import QUANT  # noqa: F401
