"""Import-path compatibility: `from lic360_operator.CconvDc import ...` (reference lic360_operator/CconvDc.py)."""
from ._modules import CconvDc, CconvDcBatch  # noqa: F401
