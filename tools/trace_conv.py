"""Per-SM timeline of the tcgen05 convolution launches of one 720p forward (fav_debug_set_trace; diagnostics).
Prints, per conv launch: units per CTA, setup, time to the first MMA, MMA issue span, cycles the issuing warp waited for
patch stages / weight chunks, epilogue lag and tail -- all in SM cycles (clock64), averaged over the CTAs with the most units."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fast-artistic-videos_b200"))
import numpy as np
import torch
from fav_b200 import _lib, models_video, synth

W64, UNITS = 64, 6
arch = synth.PAPER_ARCH if os.environ.get("FAV_ABL_ARCH") == "paper" else synth.DEFAULT_ARCH
H, Wd = (int(v) for v in os.environ.get("FAV_TRACE_SIZE", "720x1280").split("x"))
net = models_video.synthetic_model("candy", arch)
x = torch.randn(1, 7, H, Wd, device="cuda") * 50
for _ in range(2):
    net.forward(x)
buf = torch.zeros(148 * W64 * 40, dtype=torch.int64, device="cuda")
_lib.check(_lib.lib.fav_debug_set_trace(_lib.dptr(buf), buf.numel() * 8))
net.forward(x)
torch.cuda.synchronize()
used = int(_lib.lib.fav_debug_trace_words())
_lib.check(_lib.lib.fav_debug_set_trace(None, 0))
t = buf.cpu().numpy()[:used]
names = [p["name"] for p in net.profile(x) if p["kind"] == "conv"]
pos, li = 0, 0
out = []
while pos < used:
    # a launch occupies grid*64 words; grid = 148 unless fewer tiles (not the case at these sizes)
    blk = t[pos:pos + 148 * W64].reshape(-1, W64)
    pos += 148 * W64
    nu = blk[:, 4]
    gt0 = blk[:, 0].astype(np.float64)
    c0, c_setup, c_exit = blk[:, 1], blk[:, 2], blk[:, 3]
    busy = nu == nu.max()
    b = blk[busy]
    rec = dict(layer=names[li] if li < len(names) else f"conv{li}", units_max=int(nu.max()), units_min=int(nu.min()),
               ctas_at_max=int(busy.sum()), start_skew_us=float((gt0.max() - gt0.min()) / 1e3),
               total=float((b[:, 3] - b[:, 1]).mean()), setup=float((b[:, 2] - b[:, 1]).mean()))
    u_rec = []
    for u in range(min(int(nu.max()), UNITS)):
        o = 8 + 8 * u
        u_rec.append(dict(start=float((b[:, o] - b[:, 1]).mean()), first_patch=float((b[:, o + 1] - b[:, o]).mean()),
                          issue_span=float((b[:, o + 2] - b[:, o]).mean()), wait_patch=float(b[:, o + 3].mean()),
                          wait_weights=float(b[:, o + 4].mean()), acc_done=float((b[:, o + 5] - b[:, 1]).mean()),
                          epi=float((b[:, o + 6] - b[:, o + 5]).mean())))
    rec["units"] = u_rec
    last = 8 + 8 * (min(int(nu.max()), UNITS) - 1)
    rec["tail_after_last_epilogue"] = float((b[:, 3] - b[:, last + 6]).mean())
    out.append(rec)
    li += 1
for r in out:
    print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items() if k != "units"}))
    for i, u in enumerate(r["units"]):
        print("    unit", i, {k: int(v) for k, v in u.items()})
