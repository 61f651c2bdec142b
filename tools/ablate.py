"""Timing ablations of conv_tc_kernel (FAV_DBG bits; diagnostics only)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fast-artistic-videos_b200"))
    import torch
    from fav_b200 import models_video
    net = models_video.synthetic_model("candy")
    x = torch.randn(1, 7, 720, 1280, device="cuda") * 50
    for _ in range(3): prof = net.profile(x)
    torch.cuda.synchronize(); print(json.dumps({p["name"]: round(p["ms"] * 1e3, 1) for p in prof if p["kind"] == "conv"}))
else:
    for dbg, nofold in ((0, ""), (8, "")):
        env = dict(os.environ, FAV_DBG=str(dbg))
        cbg = nofold
        if nofold: env["FAV_RF_" + nofold.split("=")[0]] = nofold.split("=")[1]
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print("FAV_DBG=%2d RF=%s" % (dbg, cbg), "\n".join(out.stdout.strip().splitlines()[-24:]) if out.stdout.strip() else out.stderr[-300:], flush=True)
