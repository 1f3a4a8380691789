"""Import-path compatibility: `from lic360_operator.Dquant import ...` (reference lic360_operator/Dquant.py)."""
from ._modules import Dquant  # noqa: F401
