"""The module layer keeps the reference's signatures: every path class of `lic360_operator` is compared with the fixture that
tests/golden/make_operator_signatures.py recorded from the reference sources (argument names, order, defaults of __init__ and
forward).  No op is constructed: runs without a GPU."""
import ast
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SIGS = json.load(open(os.path.join(HERE, "golden", "operator_signatures.json")))


@pytest.fixture(scope="module")
def ops(lib_built):
    import lic360_operator
    return lic360_operator


def _mine(fn):
    out = []
    for name, p in inspect.signature(fn).parameters.items():
        assert p.kind == p.POSITIONAL_OR_KEYWORD, (fn, name)
        out.append([name, None if p.default is p.empty else p.default])
    return out


def _same_default(mine, ref_src):
    if ref_src is None or mine is None:
        return ref_src is None and mine is None
    ref = ast.literal_eval(ref_src)
    return type(ref) is type(mine) and ref == mine or (isinstance(ref, float) and float(mine) == ref)


@pytest.mark.parametrize("name", sorted(SIGS))
def test_signature(ops, name):
    cls = getattr(ops, name)
    for meth in ("__init__", "forward"):
        ref = SIGS[name][meth]
        mine = _mine(getattr(cls, meth))
        assert [a for a, _ in mine] == [a for a, _ in ref], (name, meth, SIGS[name]["file"], SIGS[name]["line"])
        for (a, dm), (_, dr) in zip(mine, ref):
            assert _same_default(dm, dr), (name, meth, a, dm, dr)


def test_exports_cover_reference(ops):
    """lic360_operator/__init__.py:2-29 -- every name the reference package exports resolves here (path classes natively,
    the pure-torch utilities by pass-through)."""
    for name in SIGS:
        assert inspect.isclass(getattr(ops, name))
    for name in ("GDN", "SSIM", "DropGrad", "ModuleSaver", "Logger"):
        assert name in ops._PASSTHROUGH
    assert "MultiProject" in SIGS  # native since round 2 (lic360.ProjectsOp)


def test_sphere_operator_alias(lib_built):
    import sphere_operator  # train/model_zoo.py:3-12
    import lic360_operator
    assert sphere_operator.CconvDc is lic360_operator.CconvDc


def test_reference_python_imports_on_the_mirror(lib_built):
    """No GPU needed: the reference's own `lic360_operator` package, `test/model_zoo.py` and `test/lic360_demo.py` import on top of this
    repo's `lic360` mirror and their codec-form modules construct (tools/run_reference_scripts.py does the rest on the GPU box)."""
    import subprocess
    import sys
    root = os.path.dirname(HERE)
    have = any(os.path.exists(os.path.join(r, "test", "lic360_demo.py")) for r in ("/root/reference", os.path.join(root, "baseline", "_ref")))
    if not have:
        pytest.skip("reference Python not available (make -f oracle/Makefile.ref pyref)")
    code = r'''
import sys
sys.path.insert(0, %r)
import run_reference_scripts as h
ref = h.bind_backend("b200")
import lic360, lic360_operator, lic360_demo as demo
assert "360-image-compression_b200" in lic360.__file__ and lic360_operator.__file__.startswith(ref) and demo.__file__.startswith(ref)
from lic360_operator import CconvEcBatch, CconvDcBatch, TileAdd, TileExtractBatch, TileInput, EntropyBatchGmmTable, MultiProject
m = CconvEcBatch(48, 1, 4, 5, 3, False, True, device=0)
assert sorted(m.state_dict()) == ["bias", "relu", "weight"] and tuple(m.weight.shape) == (3, 192, 48, 5, 5)
CconvDcBatch(48, 4, 4, 5, 3, True, True, device=0); TileAdd(48, device=0); TileExtractBatch(48, True, device=0)
TileInput(48, -3.5, 1, 3, device=0); EntropyBatchGmmTable(8, 3.5, 3, 65536, device=0); MultiProject(171, 256, 0.5, False, 0)
print("ok")
''' % os.path.join(root, "tools")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert p.returncode == 0 and "ok" in p.stdout, p.stderr[-2000:]
