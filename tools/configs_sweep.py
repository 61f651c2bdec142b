"""BASELINE.json config shapes on one GPU: frames/s of run_next_image (device-resident) and a tcgen05-vs-CUDA-core
cross-check of the whole net at each size (the fp64 oracle is too slow beyond 720p)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fast-artistic-videos_b200"))
import torch
from fav_b200 import models_video, synth

dev = torch.device("cuda")
out = []
for name, arch, (H, W) in [("720p default", synth.DEFAULT_ARCH, (720, 1280)), ("1080p default", synth.DEFAULT_ARCH, (1080, 1920)),
                           ("VR face 2048^2 paper arch", synth.PAPER_ARCH, (2048, 2048)), ("4K default", synth.DEFAULT_ARCH, (2160, 3840))]:
    net = models_video.synthetic_model("candy", arch)
    g = torch.Generator(device="cuda").manual_seed(1)
    content = torch.rand((3, H, W), device=dev, generator=g)
    prev = torch.rand((3, H, W), device=dev, generator=g)
    flow = (torch.rand((2, H, W), device=dev, generator=g) - 0.5) * 8
    cert = (torch.rand((H, W), device=dev, generator=g) > 0.1).float()
    o = net.run_next_image(content, prev, flow, cert)
    for _ in range(4):  # eager first call, graph capture per destination buffer afterwards
        o = net.run_next_image(content, o, flow, cert)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        o = net.run_next_image(content, o, flow, cert)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    x7 = torch.randn((1, 7, H, W), device=dev, generator=g) * 40
    a = net.forward(x7)
    net.set_conv_impl("simt"); b = net.forward(x7); net.set_conv_impl("tcgen05")
    diff = float((a - b).abs().max()) / 255.0
    rec = dict(config=name, H=H, W=W, ms_per_frame=round(ms, 3), fps=round(1e3 / ms, 1), tc_vs_simt_maxabs_01=diff,
               finite=bool(torch.isfinite(o).all()), mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 2))
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del net
    torch.cuda.empty_cache()
