"""Import-path compatibility: `from lic360_operator.SphereLatScaleNet import ...` (reference lic360_operator/SphereLatScaleNet.py)."""
from ._modules import SphereLatScaleNet  # noqa: F401
