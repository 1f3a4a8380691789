"""Import-path compatibility: `from lic360_operator.TileInput import ...` (reference lic360_operator/TileInput.py)."""
from ._modules import TileInput  # noqa: F401
