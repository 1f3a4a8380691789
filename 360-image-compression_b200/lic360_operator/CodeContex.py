"""Import-path compatibility: `from lic360_operator.CodeContex import ...` (reference lic360_operator/CodeContex.py)."""
from ._modules import CodeContex  # noqa: F401
