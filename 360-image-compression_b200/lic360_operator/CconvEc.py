"""Import-path compatibility: `from lic360_operator.CconvEc import ...` (reference lic360_operator/CconvEc.py)."""
from ._modules import CconvEc, CconvEcBatch  # noqa: F401
