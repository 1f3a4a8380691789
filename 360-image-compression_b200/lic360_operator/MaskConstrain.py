"""Import-path compatibility: `from lic360_operator.MaskConstrain import ...` (reference lic360_operator/MaskConstrain.py)."""
from ._modules import MaskConv2  # noqa: F401
