"""`lic360_operator` -- drop-in for the reference package of the same name (lic360_operator/__init__.py:2-29).

The context-model entropy path is implemented here on top of the B200-native `lic360` mirror.
Pure-PyTorch utilities that are outside that path (GDN, SSIM, DropGrad, ModuleSaver, Logger -- SURVEY.md s2.1 "OUT OF
SCOPE ... reused as-is") are resolved lazily from a reference checkout pointed to by $LIC360_REFERENCE_ROOT (default
/root/reference, then the copy staged under baseline/_ref by oracle/Makefile.ref); they are never copied into this
repository's history.
"""
import importlib.util
import os
import sys

from .BaseOpModule import BaseOpModule
from ._modules import (CconvDc, CconvDcBatch, CconvEc, CconvEcBatch, CodeContex, ContextReshape, ContextShift, Dquant,
                       Dtow, EntropyBatchGmmTable, EntropyGmm, EntropyGmmTable, EntropyTable, Imp2mask, ImpMap,
                       MaskConv2, MultiProject, QUANT, Scale, SphereCutEdge, SphereLatScaleNet, SpherePad, SphereTrim,
                       TileAdd, TileExtract, TileExtractBatch, TileInput)

_PASSTHROUGH = {  # attribute -> reference module that defines it
    'GDN': 'GDN', 'SSIM': 'pytorch_ssim', 'DropGrad': 'DropGrad', 'ModuleSaver': 'ModuleSaver', 'Logger': 'Logger',
}


def _load_reference_module(modname):
    here = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    roots = [os.environ.get('LIC360_REFERENCE_ROOT', '/root/reference'), os.path.join(here, 'baseline', '_ref')]
    path = os.path.join(roots[0], 'lic360_operator', modname + '.py')
    for root in roots:
        cand = os.path.join(root, 'lic360_operator', modname + '.py')
        if os.path.exists(cand):
            path = cand
            break
    if not os.path.exists(path):
        raise ImportError("lic360_operator.%s is outside the B200 hot path and is taken from a reference checkout; "
                          "set LIC360_REFERENCE_ROOT (looked for %s)" % (modname, path))
    full = 'lic360_operator.' + modname
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def __getattr__(name):
    if name in _PASSTHROUGH:
        return getattr(_load_reference_module(_PASSTHROUGH[name]), name)
    raise AttributeError("module 'lic360_operator' has no attribute %r" % name)
