"""Import-path compatibility: `from lic360_operator.EntropyGmmTable import ...` (reference lic360_operator/EntropyGmmTable.py)."""
from ._modules import EntropyGmmTable, EntropyBatchGmmTable  # noqa: F401
