"""Import-path compatibility: `from lic360_operator.EntropyTable import ...` (reference lic360_operator/EntropyTable.py)."""
from ._modules import EntropyTable  # noqa: F401
