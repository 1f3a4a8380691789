"""`sphere_operator` alias: the reference's train/ scripts import this name (train/model_zoo.py:3-12,
train/EntropyNet2.py:2-5) although the shipped package is `lic360_operator`."""
from lic360_operator import *  # noqa: F401,F403
from lic360_operator import __getattr__ as _lazy


class _Missing(object):
    """SphereMap / BinaryQuant are imported by train/model_zoo.py:12 but exist nowhere in the reference."""

    def __init__(self, *a, **k):
        raise NotImplementedError("this class does not exist in the reference repository either")


SphereMap = _Missing
BinaryQuant = _Missing


def __getattr__(name):
    return _lazy(name)
