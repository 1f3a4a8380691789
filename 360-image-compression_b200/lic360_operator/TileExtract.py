"""Import-path compatibility: `from lic360_operator.TileExtract import ...` (reference lic360_operator/TileExtract.py)."""
from ._modules import TileExtract, TileExtractBatch  # noqa: F401
