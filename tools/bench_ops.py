"""Achieved HBM GB/s of the streaming kernels of the path (SURVEY.md s8d algorithmic bytes / CUDA-event time), at the
sizes of the config-5 training batch (32 images of 512x1024 -> latents (32,192,32,64) / (32,48,64,128)) so that every
tensor is far larger than L2 is NOT assumed: a 256 MB buffer is written between timed iterations instead (L2 flush).
Writes gpurun_out/ops_roofline.json.  usage: bench_ops.py [batch]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "360-image-compression_b200")):
    sys.path.insert(0, p)
import torch
import lic360

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
peak = 6464.6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(1)
rnd = lambda *s: torch.rand(*s, device=dev, generator=g)
results = []


def timed(name, fn, nbytes, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    t = ms[len(ms) // 2]
    gbs = nbytes / (t * 1e-3) / 1e9
    results.append({"op": name, "algorithmic_bytes": int(nbytes), "median_us": t * 1e3, "achieved_gbs": gbs, "frac_of_measured_hbm": gbs / peak})
    print("%-34s %10.1f us %9.1f GB/s  %5.1f%% of %.0f" % (name, t * 1e3, gbs, 100 * gbs / peak, peak))


lat = rnd(B, 192, 32, 64)                 # analysis-transform output (sigmoid range)
nel = lat.numel()
# ---- GMM -> CDF tables over all symbols of the batch (72 B / symbol)
S = B * 48 * 64 * 128
St = 4096 * 4096 // 4  # rows of one table call: data (3, 3, 2048, 2048) = 3 planes x rows x 3 floats, out rows x 9
planes = torch.cat([rnd(St, 3) * 2 - 1, rnd(St, 3) + 0.3, rnd(St, 3) * 6 - 3]).view(3, 3, 2048, 2048).contiguous()
gt = lic360.EntropyGmmTableOp(8, 3.5, 3, 65536, 1e-6, 0, False)
cnt = torch.tensor([St], dtype=torch.int32)
timed("EntropyGmmTable.forward_batch", lambda: gt.forward_batch(planes, cnt), 72 * St)
# ---- EntropyTable (396 B / symbol)
Si = 1024 * 1024
logits = (rnd(Si, 49) * 8 - 4).view(1, 49, 1024, 1024).contiguous()  # flat memory = rows x 49
et = lic360.EntropyTableOp(49, 65536, 0, False)
timed("EntropyTable.forward", lambda: et.forward(logits, torch.tensor([Si], dtype=torch.int32)), 396 * Si)
# ---- EntropyGmm training loss fwd / bwd (84 B / symbol each)
w3 = torch.softmax(rnd(S, 3), 1).contiguous(); d3 = (rnd(S, 3) + 0.5).contiguous()
lab = (torch.randint(0, 8, (S, 1), device=dev, generator=g).float() - 3.5).contiguous(); m3 = (lab + rnd(S, 3) - 0.5).contiguous()
eg = lic360.EntropyGmmOp(3, -1, 0, False)
timed("EntropyGmm.forward", lambda: eg.forward(w3, d3, m3, lab), 84 * S)
top = rnd(S)
timed("EntropyGmm.backward", lambda: eg.backward(top), 84 * S)
# ---- quantisation / masking (12 B / element; Imp2mask 4 B)
qw = torch.zeros(192, 8, device=dev); qw[:, 0] = 0.05; qw[:, 1:] = -2.0
qc = torch.zeros(192, 8, device=dev)
qo = lic360.QuantOp(192, 8, 0.9, 100, 2, 0.1, 0, False)
timed("Quant.forward", lambda: qo.forward(lat, qw, qc, False), 12 * nel)
qi = torch.randint(0, 8, lat.shape, device=dev, generator=g).float()
msk = (rnd(*lat.shape) > 0.3).float()
do = lic360.DquantOp(192, 8, 0, False)
timed("Dquant.forward", lambda: do.forward(qi, msk, qw), 12 * nel)
imp = rnd(B, 1, 32, 64)
im = lic360.ImpMapOp(48, 1.0, 1e-4, 0.5, 1.0, 1.0, 3, 2, 0, False)
timed("ImpMap.forward", lambda: im.forward(lat, imp), 12 * nel + 4 * imp.numel())
lv = torch.floor(imp * 48)
i2 = lic360.Imp2maskOp(48, 192, 0, False)
timed("Imp2mask.forward", lambda: i2.forward(lv), 4 * nel)
# ---- layout ops (8 B / element)
y144 = rnd(B, 144, 64, 128)
cr = lic360.ContextReshapeOp(48, 0, False)
timed("ContextReshape.forward", lambda: cr.forward(y144), 8 * y144.numel())
dw = lic360.DtowOp(2, True, 0, False)
timed("Dtow.forward (d2w)", lambda: dw.forward(lat), 8 * nel)
sc = lic360.ScaleOp(-1.0, 2.0 / 47, 0, False)
timed("Scale.forward", lambda: sc.forward(lat), 8 * nel)
# ---- sphere ops on a transform-sized tensor (B,192,128,256)
xt = rnd(max(1, B // 4), 192, 128, 256)
sp = lic360.SpherePadOp(2, False, 0, False)
out_el = xt.shape[0] * 192 * 132 * 260
timed("SpherePad.forward (alloc)", lambda: sp.forward(xt), 8 * out_el)
xp = sp.forward(xt)[0].clone()
border = out_el - xt.numel()
spi = lic360.SpherePadOp(2, True, 0, False)
timed("SpherePad.forward (in place)", lambda: spi.forward(xp), 8 * border)
st = lic360.SphereTrimOp(2, 0, False)
timed("SphereTrim.forward", lambda: st.forward(xp), 4 * border)
ce = lic360.SphereCutEdgeOp(2, 0, False)
timed("SphereCutEdge.forward", lambda: ce.forward(xp), 8 * xt.numel())
ls = lic360.SphereLatScaleOp(32, 0, False)
wl = rnd(B, 1, 32)
imap = rnd(B, 1, 32, 64)
big = rnd(B * 64, 1, 32, 64)
wl2 = rnd(B * 64, 1, 32)
timed("SphereLatScale.forward", lambda: ls.forward(big, wl2), 8 * big.numel())
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"batch": B, "peak_gbs": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)", "l2": "256 MB buffer written between iterations",
           "ops": results}, open(os.path.join(ROOT, "gpurun_out", "ops_roofline.json"), "w"), indent=1)
