#!/usr/bin/env python
"""Runs the reference's OWN Python -- lic360_operator/*.py, test/model_zoo.py and test/lic360_demo.py, unchanged -- on top of a
chosen native backend, end to end on one synthetic 512x1024 ERP image with seeded random-init model-idx-3 weights
(BASELINE.json configs[1], SURVEY.md s8d "Config 2"):

    python tools/run_reference_scripts.py --backend b200      --workdir D   # this repo's `lic360` mirror (C-ABI, sm_100a kernels)
    python tools/run_reference_scripts.py --backend reference --workdir D   # the unmodified reference extension (oracle/_ref)

Both runs share D/weights (made once, by whichever run comes first) and D/erp.png, call lic360_demo.encoding / decoding /
decoding_and_test (the functions behind --enc / --dec / --test, lic360_demo.py:339-449) with the hard-wired weight directories
redirected, and write D/<backend>.json: stream sizes and hashes, the bpp / PSNR / SSIM lines the script printed, and a hash of the
decoded image.  With --backend b200 the two bitstreams are additionally produced by the product path (FusedCodec) from the same
latent and compared byte for byte with the files the reference's Python loops wrote.

The reference sources are read from $LIC360_REFERENCE_ROOT (default /root/reference) or from the copy staged under baseline/_ref by
oracle/Makefile.ref (the GPU box has no /root/reference).  Test infrastructure: nothing under 360-image-compression_b200/ imports it.
"""
import argparse
import contextlib
import hashlib
import io
import json
import os
import re
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "360-image-compression_b200")


def reference_root():
    for cand in (os.environ.get("LIC360_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.exists(os.path.join(cand, "test", "lic360_demo.py")):
            return cand
    raise SystemExit("reference Python not found (run `make -f oracle/Makefile.ref pyref` in the build container)")


def bind_backend(backend):
    """make `import lic360` resolve to the chosen native backend and `import lic360_operator` to the REFERENCE's package"""
    if "tkinter" not in sys.modules:  # Dquant.py:1 imports tkinter.messagebox.NO (unused); the image has no tkinter
        tk = types.ModuleType("tkinter")
        mb = types.ModuleType("tkinter.messagebox")
        mb.NO = "no"
        tk.messagebox = mb
        sys.modules["tkinter"], sys.modules["tkinter.messagebox"] = tk, mb
    if backend == "b200":
        sys.path.insert(0, PKG)
        import lic360  # noqa: F401  this repo's mirror
        sys.path.remove(PKG)  # ... but NOT this repo's lic360_operator: the reference's own modules are what is under test
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        import lic360_ref
        sys.modules["lic360"] = lic360_ref
    ref = reference_root()
    sys.path.insert(0, os.path.join(ref, "test"))
    sys.path.insert(0, ref)
    import lic360_operator
    assert os.path.dirname(os.path.abspath(lic360_operator.__file__)).startswith(os.path.abspath(ref)), lic360_operator.__file__
    return ref


def synthetic_erp(path, seed=2024):
    """seeded low-pass noise + latitude gradient, 512x1024 RGB, written as PNG (SURVEY.md s8d Config 2)"""
    import cv2
    import numpy as np
    r = np.random.default_rng(seed)
    img = r.random((512 // 8, 1024 // 8, 3)).astype(np.float32)
    img = cv2.resize(img, (1024, 512), interpolation=cv2.INTER_CUBIC)
    lat = np.linspace(0.15, 0.85, 512, dtype=np.float32)[:, None, None]
    img = np.clip(0.6 * img + 0.4 * lat + 0.02 * r.standard_normal((512, 1024, 3)).astype(np.float32), 0, 1)
    cv2.imwrite(path, (img * 255).astype(np.uint8))


def make_weights(wdir, prex, seed=2024):
    """seeded random-init CMP_FULL (test/model_zoo.py:304-319: analysis / synthesis transforms, quantiser, importance map, EntropyNet2,
    EntropyNet3) saved under the file names lic360_demo.py:341-343 expects"""
    import torch
    import model_zoo
    args = argparse.Namespace(channels=192, code_channels=192, quant_levels=8, gpu_id=0, rt=1.0, la=0.0001, lb=0.0001,
                              scale_const=0.618, scale_weight=0.618, init=False)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        net = model_zoo.CMP_FULL(args).to("cuda:0")
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    os.makedirs(wdir, exist_ok=True)
    torch.save(sd, os.path.join(wdir, prex + "_v0_best_0.pt"))
    torch.save(sd, os.path.join(wdir, prex + "_imp_best_0.pt"))


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["b200", "reference"], required=True)
    ap.add_argument("--workdir", required=True)
    ap.add_argument("--model-idx", type=int, default=3)
    args = ap.parse_args()
    os.makedirs(args.workdir, exist_ok=True)
    ref = bind_backend(args.backend)
    import numpy as np
    import torch
    import lic360
    import lic360_demo as demo  # the reference script, imported as a module: its functions are what --enc/--dec/--test call
    assert os.path.abspath(demo.__file__).startswith(os.path.abspath(ref))
    prex = demo.model_ssim_list[args.model_idx]
    wdir = os.path.join(args.workdir, "weights")
    if not os.path.exists(os.path.join(wdir, prex + "_imp_best_0.pt")):
        make_weights(wdir, prex)
    demo.mse_model_dir = demo.ssim_model_dir = wdir  # the script hard-wires E:/360_dataset/... (lic360_demo.py:18-19)
    png = os.path.join(args.workdir, "erp.png")
    if not os.path.exists(png):
        synthetic_erp(png)
    tag = args.backend
    code = os.path.join(args.workdir, "%s_code" % tag)
    out_png = os.path.join(args.workdir, "%s_dec.png" % tag)
    log = io.StringIO()
    with contextlib.redirect_stdout(log):
        demo.encoding([png], [code], args.model_idx, False, 0)        # --enc --ssim
        demo.decoding([code], [out_png], args.model_idx, False, 0)     # --dec --ssim
        demo.decoding_and_test([code], [png], args.model_idx, False, 0)  # --test --ssim
    text = log.getvalue()
    sys.stderr.write(text)
    m_enc = re.search(r"bitrate: ([0-9.]+)bpp", text)
    m_test = re.search(r"Bitrate:([0-9.]+)bpp, PSNR:([0-9.]+)dB, SSIM:([0-9.]+)", text)
    import cv2
    dec = cv2.imread(out_png)
    rep = {"backend": tag, "model_idx": args.model_idx, "lic360_module": getattr(lic360, "__file__", str(lic360)),
           "code_bytes": os.path.getsize(code), "imp_bytes": os.path.getsize(code + "_imp"),
           "code_sha256": sha(code), "imp_sha256": sha(code + "_imp"),
           "printed_encode_bpp": float(m_enc.group(1)) if m_enc else None,
           "printed_test": {"bpp": float(m_test.group(1)), "psnr_db": float(m_test.group(2)), "ssim": float(m_test.group(3))} if m_test else None,
           "decoded_png_sha256": sha(out_png), "decoded_mean": float(dec.mean()), "decoded_shape": list(dec.shape)}

    # ---- more digits than the script prints: viewport MSE / PSNR / SSIM recomputed with the same calls (lic360_demo.py:424-441)
    import math
    from lic360_operator import MultiProject, SSIM
    pr1, pr2 = MultiProject(171, int(171 * 1.5), 0.5, False, 0).to("cuda:0"), MultiProject(171, int(171 * 1.5), 0.5, False, 0).to("cuda:0")
    x = pr1(demo.img2tensor(demo.check_img(cv2.imread(png)), "cuda:0"))
    y = pr2(demo.img2tensor(dec, "cuda:0"))
    mse = torch.mean((x - y) ** 2).item()
    rep["viewport_metrics_vs_decoded_png"] = {"mse": mse, "psnr_db": 10 * math.log10(1. / mse), "ssim": SSIM(11, 3).to("cuda:0")(x, y).item()}

    if args.backend == "b200":
        # ---- the product path on the same latent: FusedCodec bytes == the files the reference's Python loops wrote
        sys.path.insert(0, PKG)
        import lic360_pipeline as pl
        params = torch.load(os.path.join(wdir, prex + "_v0_best_0.pt"), map_location="cuda:0")
        with contextlib.redirect_stdout(io.StringIO()):
            enc_net = demo.CMP_Encoder(gpu_id=0).to("cuda:0")
            enc_net.load_state_dict({k: params[k] for k in enc_net.state_dict()})
            cparams = {
                "code": {k: v.contiguous() for k, v in demo.cast_entropy_parameter(params, demo.EntEncoderFast(ngroup=48).to("cuda:0").state_dict()).items()},
                "imp": {k: v.contiguous() for k, v in demo.cast_imp_entropy_parameter(params, demo.ImpEntEncoderFast().to("cuda:0").state_dict()).items()},
            }
        with torch.no_grad():
            qy_up, mask_up, imap_quant = enc_net(demo.img2tensor(demo.check_img(cv2.imread(png)), "cuda:0"))
        fused = pl.FusedCodec(cparams, H=64, W=128)
        bi, bc = fused.encode(qy_up.contiguous().clone(), mask_up.contiguous().clone(), imap_quant.contiguous().clone())
        fcode, fmask = fused.decode(open(code + "_imp", "rb").read(), open(code, "rb").read())
        rep["fused_codec"] = {"imp_bytes": len(bi), "code_bytes": len(bc),
                              "imp_identical_to_script_file": bi == open(code + "_imp", "rb").read(),
                              "code_identical_to_script_file": bc == open(code, "rb").read(),
                              "decodes_script_files_exactly": bool(torch.equal(fcode, qy_up * mask_up) and torch.equal(fmask, mask_up)),
                              "kept_symbols": int(mask_up.sum().item()), "native_launches": int(lic360.launch_count())}
    with open(os.path.join(args.workdir, tag + ".json"), "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
