"""Import-path compatibility: `from lic360_operator.SphereCutEdge import ...` (reference lic360_operator/SphereCutEdge.py)."""
from ._modules import SphereCutEdge  # noqa: F401
