"""Timing ablations of the conv / apply kernels (env knobs: FAV_DBG bits, FAV_NO_NL, FAV_NO_KSPLIT, FAV_NO_FOLD; diagnostics only)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fast-artistic-videos_b200"))
    import torch
    from fav_b200 import models_video
    from fav_b200 import synth
    net = models_video.synthetic_model("candy", synth.PAPER_ARCH if os.environ.get("FAV_ABL_ARCH") == "paper" else synth.DEFAULT_ARCH)
    x = torch.randn(1, 7, 720, 1280, device="cuda") * 50
    for _ in range(3): prof = net.profile(x)
    torch.cuda.synchronize()
    keep = os.environ.get("FAV_ABL_ONLY")
    print(json.dumps({p["name"]: round(p["ms"] * 1e3, 1) for p in prof if not keep or p["name"] in keep.split("+")}), "total", round(sum(p["ms"] for p in prof) * 1e3, 1))
else:
    variants = [dict(), dict(FAV_NO_NL="1")] if len(sys.argv) < 2 else [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[1:]]
    for v in variants:
        out = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **v), capture_output=True, text=True)
        print(v, "\n".join(out.stdout.strip().splitlines()[-2:]) if out.stdout.strip() else out.stderr[-400:], flush=True)
