"""Import-path compatibility: `from lic360_operator.ContextReshape import ...` (reference lic360_operator/ContextReshape.py)."""
from ._modules import ContextReshape  # noqa: F401
