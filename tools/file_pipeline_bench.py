"""f-2 measurement: a 300-frame 1280x720 clip on tmpfs (PPM frames + .flo + PGM certainty) through the pipelined file driver
(python -m fav_b200.video) -> PNGs; reports frames/s files -> PNG beside the synchronous driver on a shorter prefix."""
import os, shutil, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fast-artistic-videos_b200"))
import numpy as np
from fav_b200 import synth, video, models_video

N = int(os.environ.get("FAV_CLIP_FRAMES", "300"))
H, W = 720, 1280
d = "/dev/shm/fav_clip"
shutil.rmtree(d, ignore_errors=True); os.makedirs(d)
POOL = 6  # distinct synthetic frames / flows, linked under the per-frame names
for k in range(POOL):
    synth.write_ppm(f"{d}/src_frame_{k}.ppm", synth.make_frame(H, W, k + 1))
    synth.write_flo(f"{d}/src_flow_{k}.flo", synth.make_backward_flow(H, W, k + 2))
    rel = (np.random.default_rng(k).uniform(size=(H, W)) > 0.05).astype(np.uint8) * 255
    open(f"{d}/src_rel_{k}.pgm", "wb").write(b"P5\n%d %d\n255\n" % (W, H) + rel.tobytes())
for i in range(1, N + 1):
    os.link(f"{d}/src_frame_{i % POOL}.ppm", f"{d}/frame_{i:04d}.ppm")
    if i > 1:
        os.link(f"{d}/src_flow_{i % POOL}.flo", f"{d}/backward_{i}_{i - 1}.flo")
        os.link(f"{d}/src_rel_{i % POOL}.pgm", f"{d}/reliable_{i}_{i - 1}.pgm")
net = models_video.synthetic_model("candy")
opt = video.build_parser().parse_args(["-input_pattern", f"{d}/frame_%04d.ppm", "-flow_pattern", f"{d}/backward_[%d]_{{%d}}.flo",
                                       "-occlusions_pattern", f"{d}/reliable_[%d]_{{%d}}.pgm", "-output_prefix", f"{d}/out", "-num_frames", str(N)])
out = {}
os.environ["FAV_PIPE_STATS"] = "1"
for nd, ne, lvl in ((8, 24, 1), (16, 40, 1), (16, 40, 0), (16, 40, 6)):
    video.run_native(opt, model_vid=net, n_decode=nd, n_encode=ne, png_level=lvl)
    r = video.run_native(opt, model_vid=net, n_decode=nd, n_encode=ne, png_level=lvl)
    out[f"native_dec{nd}_enc{ne}_z{lvl}"] = r["frames"] / r["seconds"]
for nd, ne in ((16, 64),):
    video.run_pipelined(opt, depth=max(8, ne // 2), n_decode=nd, n_encode=ne, model_vid=net)  # warm (page cache, plans, graphs)
    r = video.run_pipelined(opt, depth=max(8, ne // 2), n_decode=nd, n_encode=ne, model_vid=net)
    out[f"pipelined_dec{nd}_enc{ne}"] = r["frames"] / r["seconds"]
# synchronous driver on a prefix
import torch
from fav_b200 import core
ns = 40
opt_s = video.build_parser().parse_args(["-input_pattern", f"{d}/frame_%04d.ppm", "-flow_pattern", f"{d}/backward_[%d]_{{%d}}.flo",
                                         "-occlusions_pattern", f"{d}/reliable_[%d]_{{%d}}.pgm", "-output_prefix", f"{d}/sync", "-num_frames", str(ns), "-pipeline", "0"])
drv = video.Driver(opt_s)
t0 = time.perf_counter()
core.run_fast_neural_video(opt_s, drv.func_load_image, drv.func_load_cert, None, drv.func_make_last_frame_warped, drv.func_is_single_image,
                           drv.func_save_image, model_vid=net)
torch.cuda.synchronize()
out["synchronous_driver"] = ns / (time.perf_counter() - t0)
out["frames"] = N; out["cpus"] = os.cpu_count(); out["cpus_affinity"] = len(os.sched_getaffinity(0))
try:
    out["cgroup_cpu_max"] = open("/sys/fs/cgroup/cpu.max").read().strip()
except OSError:
    out["cgroup_cpu_max"] = None
print(json.dumps(out))
shutil.rmtree(d, ignore_errors=True)
