"""Import-path compatibility: `from lic360_operator.TileAdd import ...` (reference lic360_operator/TileAdd.py)."""
from ._modules import TileAdd  # noqa: F401
