"""Golden vectors of the REFERENCE's own CUDA warp kernel (run on a GPU box: `python tests/golden/make_warp_golden.py`).

oracle/_ref/libref_warp.so = stnbdhw/BilinearSamplerBDHW.cu:48-109 compiled for sm_100a (oracle/Makefile refwarp).
Writes gpurun_out/warp_ref.npz = outputs on the seeded inputs of WARP_CASES; the file is then committed as
tests/golden/warp_ref.npz and pins oracle/fav_oracle.c:orc_warp_bdhw on the CPU (tests/test_oracle.py).
Inputs are regenerated from seeds at test time; only outputs are stored."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fast-artistic-videos_b200"))

from fav_b200 import synth  # noqa: E402

WARP_CASES = [  # (name, C, Hin, Win, Hout, Wout, flow kind)
    ("real_64x96", 3, 64, 96, 64, 96, "real"), ("stress_64x96", 3, 64, 96, 64, 96, "stress"),
    ("real_97x75", 3, 97, 75, 97, 75, "real"), ("stress_33x129", 1, 33, 129, 33, 129, "stress"),
    ("resize_40x56_to_52x44", 5, 40, 56, 52, 44, "rand"), ("sentinel_32x48", 1, 32, 48, 32, 48, "sentinel")]


def warp_inputs(case):
    name, C, Hin, Win, Ho, Wo, kind = case
    rng = np.random.default_rng(sum(map(ord, name)))
    if C == 3:
        img = synth.make_frame(Hin, Win, 1) * 1.2 - 0.1
    else:
        img = rng.uniform(-0.2, 1.2, size=(C, Hin, Win)).astype(np.float32)
    if kind == "real":
        flow = synth.checker_to_lua(synth.make_backward_flow(Ho, Wo, 2))
    elif kind == "stress":
        flow = synth.stress_flow(Ho, Wo)
    elif kind == "sentinel":  # vr_helper.lua:10 maps: 99999 outside the strip, a real offset inside
        flow = np.full((2, Ho, Wo), 99999.0, np.float32)
        flow[:, :, Wo // 3: Wo // 2] = rng.uniform(-5, 5, size=(2, Ho, Wo // 2 - Wo // 3)).astype(np.float32)
    else:
        flow = rng.uniform(-9, 9, size=(2, Ho, Wo)).astype(np.float32)
    return np.ascontiguousarray(img, np.float32), np.ascontiguousarray(flow, np.float32)


def main():
    import torch

    from oracle import refwarp

    out = {}
    for case in WARP_CASES:
        img, flow = warp_inputs(case)
        o = refwarp.warp(torch.from_numpy(img).cuda()[None], torch.from_numpy(flow).cuda()[None])[0]
        out[case[0]] = o.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "warp_ref.npz"), **out)
    print("wrote gpurun_out/warp_ref.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
