"""Records the constructor / forward signatures of the reference's operator modules (lic360_operator/*.py) by parsing the
sources with `ast` (nothing is imported or executed):  python tests/golden/make_operator_signatures.py
-> tests/golden/operator_signatures.json.  Run in the build container, where /root/reference exists."""
import ast
import json
import os

REF = os.environ.get("LIC360_REFERENCE_ROOT", "/root/reference")
PKG = os.path.join(REF, "lic360_operator")
WANT = ["CconvDc", "CconvDcBatch", "CconvEc", "CconvEcBatch", "CodeContex", "ContextReshape", "ContextShift", "Dquant", "Dtow",
        "EntropyBatchGmmTable", "EntropyGmm", "EntropyGmmTable", "EntropyTable", "Imp2mask", "ImpMap", "MaskConv2", "QUANT", "Scale",
        "SphereCutEdge", "SphereLatScaleNet", "SpherePad", "SphereTrim", "TileAdd", "TileExtract", "TileExtractBatch", "TileInput",
        "MultiProject"]


def sig(fn):
    a = fn.args
    names = [x.arg for x in a.args]
    defaults = [ast.unparse(d) for d in a.defaults]
    pad = [None] * (len(names) - len(defaults))
    return [[n, d] for n, d in zip(names, pad + defaults)]


out = {}
for f in sorted(os.listdir(PKG)):
    if not f.endswith(".py"):
        continue
    tree = ast.parse(open(os.path.join(PKG, f)).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in WANT:
            ent = {"file": "lic360_operator/" + f, "line": node.lineno}
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in ("__init__", "forward"):
                    ent[item.name] = sig(item)
            out[node.name] = ent
missing = [w for w in WANT if w not in out]
assert not missing, missing
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "operator_signatures.json")
json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
print("wrote", dst, len(out), "classes")
