"""One 720p frame of the whole path (fused temporal stage + the 27 net kernels, eager launches) plus one launch of every
other kernel of the library, between cudaProfilerStart/Stop -- the target of the ncu passes whose summaries live in profiles/:
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python tools/ncu_frame.py
  ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r02_frame python tools/ncu_frame.py
FAV_NO_GRAPH=1 is set here so that the net's kernels are individual launches."""
import os, sys
os.environ["FAV_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fast-artistic-videos_b200"))
import numpy as np
import torch
from fav_b200 import _lib, consistencyChecker, models_video, synth, utils

H, W = 720, 1280
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
arch = synth.PAPER_ARCH if os.environ.get("FAV_ABL_ARCH") == "paper" else synth.DEFAULT_ARCH
net = models_video.synthetic_model("candy", arch)
frames = [T(synth.make_frame(H, W, i)) for i in (1, 2, 3)]
bw = [T(synth.make_backward_flow(H, W, i)) for i in (2, 3)]
fw = [T(synth.make_forward_flow(H, W, i)) for i in (2, 3)]
lua = [torch.stack([b[1], b[0]]).contiguous() for b in bw]
prev = net.run_image(frames[0])
for k in range(3):
    prev = net.run_next_image_flows(frames[1 + k % 2], prev, lua[k % 2], fw[k % 2], None, 7)
img255 = (frames[0] * 255).round()
out7 = torch.empty((7, H, W), device="cuda")
face = T(synth.make_frame(512, 512, 5))
torch.cuda.synchronize()
torch.cuda.profiler.start()
prev = net.run_next_image_flows(frames[2], prev, lua[1], fw[1], None, 7)       # the frame: temporal_stage + net
_, cert = consistencyChecker.check(bw[0], fw[0], want_cert=True)                  # consistency_kernel
certm = utils.min_filter(cert, 7)                                                # min_filter_kernel
_lib.check(_lib.lib.fav_temporal_input(_lib.dptr(frames[1]), _lib.dptr(prev), _lib.dptr(lua[0]), _lib.dptr(certm), None, None,
                                       _lib.dptr(out7), H, W, 0, _lib.stream_ptr()))  # temporal_input_kernel
w = utils.warp_image(prev, lua[0])                                               # warp_vec4_kernel
consistencyChecker.check(bw[0], fw[0], img255)                                    # 4-argument mode: corners, IIR, normalize, avg
utils.median_filter(face, 3)                                                     # median_kernel
utils.temporal_loss(prev, frames[1], lua[0], certm)                              # temporal_mse_kernel
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(prev.sum()))
