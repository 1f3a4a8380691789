"""GPU: the reference's own Python -- lic360_operator/*.py, test/model_zoo.py, test/lic360_demo.py, UNCHANGED -- runs on top of this
repo's `lic360` mirror (north_star: "the lic360_operator Python modules keep their signatures so test/lic360_demo.py ... run
unchanged"), end to end on one synthetic 512x1024 ERP image with seeded random-init model-idx-3 weights (BASELINE.json configs[1]):
encoding(), decoding() and decoding_and_test() of lic360_demo.py:339-449, i.e. what --enc / --dec / --test execute.

The same run is repeated with the unmodified reference CUDA extension as the backend (tools/run_reference_scripts.py, one
subprocess per backend: both are called `lic360` by the scripts), and the two are compared:
  * bpp: stream sizes within max(2 bytes, 0.1 %) (the context conv is float-tier, tests/test_gpu_parity_baseline.py);
  * the decoded image is IDENTICAL (entropy coding is lossless and the transforms run the same cuDNN / bit-exact sphere ops);
  * viewport PSNR within 1e-3 dB and SSIM within 1e-5 (MultiProject is float-tier).
And the product path is pinned against the scripts: FusedCodec emits byte for byte the two files the reference's Python loops
wrote through the mirror, and decodes them.
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "run_reference_scripts.py")


def _have_reference_python():
    return any(os.path.exists(os.path.join(r, "test", "lic360_demo.py"))
               for r in (os.environ.get("LIC360_REFERENCE_ROOT") or "/nonexistent", "/root/reference", os.path.join(ROOT, "baseline", "_ref")))


def _run(backend, workdir):
    p = subprocess.run([sys.executable, TOOL, "--backend", backend, "--workdir", workdir], capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0, "backend %s failed:\n%s\n%s" % (backend, p.stdout[-3000:], p.stderr[-6000:])
    return json.load(open(os.path.join(workdir, backend + ".json")))


@pytest.fixture(scope="module")
def runs(tmp_path_factory, ref_ext):
    if not _have_reference_python():
        pytest.skip("the reference's Python is not staged (make -f oracle/Makefile.ref pyref)")
    wd = str(tmp_path_factory.mktemp("refscripts"))
    out = {"b200": _run("b200", wd)}
    if ref_ext is not None:
        out["reference"] = _run("reference", wd)
    dst = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(dst):
        json.dump(out, open(os.path.join(dst, "r2_reference_scripts.json"), "w"), indent=1)
    return out


def test_demo_script_runs_on_the_mirror(runs):
    r = runs["b200"]
    assert "360-image-compression_b200" in r["lic360_module"]
    assert r["code_bytes"] > 1000 and r["imp_bytes"] > 10
    assert r["printed_test"] is not None and r["printed_test"]["psnr_db"] > 5 and 0 < r["printed_test"]["ssim"] <= 1
    assert r["decoded_shape"] == [512, 1024, 3]
    assert abs(r["printed_encode_bpp"] - (r["code_bytes"] + r["imp_bytes"]) * 8 / (512 * 1024)) < 1e-3   # lic360_demo.py:365-366
    assert abs(r["printed_test"]["bpp"] - r["code_bytes"] * 8 / (512 * 1024)) < 1e-3                        # :442 ignores the _imp file


def test_fused_codec_matches_the_script_files(runs):
    f = runs["b200"]["fused_codec"]
    assert f["native_launches"] > 0 and f["kept_symbols"] > 0
    assert f["imp_identical_to_script_file"] and f["code_identical_to_script_file"], f
    assert f["decodes_script_files_exactly"], f


def test_bpp_psnr_ssim_match_the_reference_extension(runs):
    if "reference" not in runs:
        pytest.skip("oracle/_ref/lic360_ref*.so not present")
    a, b = runs["b200"], runs["reference"]
    assert abs(a["code_bytes"] - b["code_bytes"]) <= max(2, 0.001 * b["code_bytes"]), (a["code_bytes"], b["code_bytes"])
    assert abs(a["imp_bytes"] - b["imp_bytes"]) <= max(2, 0.002 * b["imp_bytes"]), (a["imp_bytes"], b["imp_bytes"])
    assert a["decoded_png_sha256"] == b["decoded_png_sha256"], "decoded images differ"
    ma, mb = a["viewport_metrics_vs_decoded_png"], b["viewport_metrics_vs_decoded_png"]
    assert abs(ma["psnr_db"] - mb["psnr_db"]) <= 1e-3 and abs(ma["ssim"] - mb["ssim"]) <= 1e-5, (ma, mb)
    assert abs(a["printed_test"]["psnr_db"] - b["printed_test"]["psnr_db"]) <= 0.011   # printed with 2 decimals
    assert abs(a["printed_test"]["ssim"] - b["printed_test"]["ssim"]) <= 1.1e-4        # printed with 4 decimals
