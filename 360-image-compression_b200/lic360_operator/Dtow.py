"""Import-path compatibility: `from lic360_operator.Dtow import ...` (reference lic360_operator/Dtow.py)."""
from ._modules import Dtow  # noqa: F401
