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
