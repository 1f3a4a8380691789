"""The C-ABI library loads and exports every symbol include/lic360_b200.h declares (no compute calls)."""
import os
import re
import subprocess


def _header_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "include", "lic360_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lic360_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(lib_built):
    syms = _header_symbols()
    assert len(syms) > 45
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_built]).decode()
    exported = set(re.findall(r" T (lic360_[a-z0-9_]+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, "declared in the header but not exported: %s" % missing


def test_python_binding_covers_header(lib_built):
    from lic360 import _lib
    syms = _header_symbols()
    assert sorted(_lib.SIGNATURES) == syms
    assert _lib.LIB.lic360_version() == 100
    assert _lib.LIB.lic360_launch_count() == 0


def test_library_is_sm100a_only(lib_built):
    out = subprocess.check_output(["cuobjdump", "-lelf", lib_built]).decode()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_oracle_in_product():
    """the product path must not import/link the oracle (tier rule 3)"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "360-image-compression_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, fn)).read()
                for pat in ("liboracle", "import oracle", "from oracle", "oracle.oracle", "orc_"):
                    assert pat not in src, "%s references the oracle (%s)" % (os.path.join(dp, fn), pat)
