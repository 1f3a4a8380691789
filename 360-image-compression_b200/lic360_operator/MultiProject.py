from ._modules import MultiProject  # noqa: F401  (import-path compatibility with the reference layout)
