"""fast_artistic_video.lua -- the plain-video driver: CLI flags (:21-67), filename patterns (:70-77), callbacks
(:93-172) and main (:174-189).  `python -m fav_b200.video -input_pattern frame_%04d.ppm -flow_pattern ... `.

File formats: frames PPM/PNG/JPG (PIL), certainty PGM (the consistencyChecker output), flow Middlebury .flo.
PNG quantisation of image.save lives in the un-vendored torch `image` rock (unpinned): clamp to [0,1], x255, round.
"""
from __future__ import annotations

import argparse
import os
import re
import sys

import numpy as np
import torch

from . import _lib, core, flowFileLoader, utils


def build_parser():
    p = argparse.ArgumentParser(prefix_chars="-", description="Stylize a video with a trained feedforward model")
    a = p.add_argument
    a("-model_img", default="self"); a("-model_vid", default="synthetic:candy")
    a("-num_frames", type=int, default=9999); a("-continue_with", type=int, default=1)
    a("-input_pattern", default=""); a("-output_prefix", default="out")
    a("-flow_pattern", default=""); a("-occlusions_pattern", default="")
    a("-invert_occlusion", action="store_true"); a("-occlusions_min_filter", type=int, default=7)
    a("-fill_occlusions", default="vgg-mean"); a("-fix_occlusions", action="store_true")
    a("-median_filter", type=int, default=3)  # declared but unused by the plain driver (fast_artistic_video.lua:39)
    a("-scale_factor", type=float, default=1); a("-backward", action="store_true")
    a("-create_inconsistent", action="store_true")
    a("-gpu", type=int, default=0); a("-backend", default="cuda"); a("-use_cudnn", type=int, default=1)
    a("-cudnn_benchmark", type=int, default=0); a("-evaluate", action="store_true")
    a("-evaluation_file", default="evaluation.txt"); a("-flow_pattern_eval", default=""); a("-occlusions_pattern_eval", default="")
    a("-invert_occlusion_eval", action="store_true"); a("-fix_occlusions_eval", action="store_true"); a("-backward_eval", action="store_true")
    a("-model_img_arch", default="")
    a("-arch", default=core.synth.DEFAULT_ARCH)
    return p


def getFormatedFlowFileName(pattern: str, fromIndex: int, toIndex: int) -> str:
    """fast_artistic_video.lua:70-77: {fmt} is formatted with the from-index, [fmt] with the to-index."""
    s = re.sub(r"\{(.*?)\}", lambda m: m.group(1) % fromIndex, pattern)
    return re.sub(r"\[(.*?)\]", lambda m: m.group(1) % toIndex, s)


def load_image(path: str, channels: int) -> torch.Tensor:
    """image.load(path, channels): float [0,1], CxHxW."""
    from PIL import Image

    im = Image.open(path).convert("RGB" if channels == 3 else "L")
    a = np.asarray(im, dtype=np.float32) / 255.0
    return torch.from_numpy(a.transpose(2, 0, 1).copy() if channels == 3 else a[None].copy())


def save_image(path: str, img: torch.Tensor) -> None:
    from PIL import Image

    a = (img.detach().clamp(0, 1) * 255.0 + 0.5).floor().clamp(0, 255).byte().cpu().numpy().transpose(1, 2, 0)
    Image.fromarray(a).save(path)


class Driver:
    def __init__(self, opt):
        self.opt = opt
        self.last_frame_stylized = None  # fast_artistic_video.lua:89
        self.prev_last_frame_stylized = None
        self.last_frame = None

    def func_load_image(self, opt, i, dtype):  # :93-97
        path = opt.input_pattern % i
        if not utils.file_exists(path):
            return None
        self.last_frame = load_image(path, 3)
        return self.last_frame

    def fix_occlusions(self, flow, occluded):  # :79-86
        tmp = utils.warp_image(torch.ones_like(occluded), flow)
        occluded.mul_(((tmp - 0.5).sign()).clamp(min=0))

    def func_load_cert(self, opt, i, dtype):  # :99-112
        flowFileName = getFormatedFlowFileName(opt.flow_pattern, i - 1, i)
        certFileName = getFormatedFlowFileName(opt.occlusions_pattern, i - 1, i)
        utils.wait_for_file(certFileName)
        cert = load_image(certFileName, 1).cuda()
        if opt.invert_occlusion:
            cert = 1.0 - cert
        if opt.fix_occlusions:
            flow = torch.from_numpy(flowFileLoader.load(flowFileName)).cuda()
            self.fix_occlusions(flow, cert)
        return cert

    def func_make_last_frame_warped(self, opt, i, dtype, cert_mask=None):  # :153-158
        flowFileName = getFormatedFlowFileName(opt.flow_pattern, i - 1, i)
        flow = torch.from_numpy(flowFileLoader.load(flowFileName)).cuda()
        return core.FusedWarp(self.last_frame_stylized, flow), None

    def func_load_flow_cert_eval(self, opt, i, dtype):  # :114-126
        flow = torch.from_numpy(flowFileLoader.load(getFormatedFlowFileName(opt.flow_pattern_eval, i - 1, i))).cuda()
        cert = load_image(getFormatedFlowFileName(opt.occlusions_pattern_eval, i - 1, i), 1).cuda()
        if opt.invert_occlusion_eval:
            cert = 1.0 - cert
        if opt.fix_occlusions_eval:
            self.fix_occlusions(flow, cert)
        return flow, cert

    def func_eval(self, opt, i, func_percept_loss, dtype):  # :128-151
        nan = float("nan")  # style / content loss: the VGG-16 loss network is not available (core.py)
        if i > 1:
            flow_eval, cert_eval = self.func_load_flow_cert_eval(opt, i, dtype)
            if opt.backward_eval:
                t = utils.temporal_loss(self.last_frame_stylized, self.prev_last_frame_stylized, flow_eval, cert_eval)
            else:
                t = utils.temporal_loss(self.prev_last_frame_stylized, self.last_frame_stylized, flow_eval, cert_eval)
            return [nan, nan, t], 3
        return [nan, nan, 0.0], 3

    def func_save_image(self, opt, i, img, dtype=None):  # :160-170
        out_path = "%s-%05d.png" % (opt.output_prefix, i)
        print("Writing output image to " + out_path)
        d = os.path.dirname(out_path)
        if d and not os.path.isdir(d):
            os.makedirs(d)
        save_image(out_path, img)
        self.prev_last_frame_stylized = self.last_frame_stylized
        self.last_frame_stylized = img.clone()  # fp32, unclamped: NOT the PNG

    @staticmethod
    def func_is_single_image(i, opt):  # :172
        return i == 1 or opt.create_inconsistent


def main(argv=None):
    opt = build_parser().parse_args(argv)
    if opt.input_pattern == "":
        raise SystemExit("Must give -input_pattern")  # :177-179
    if not opt.create_inconsistent and (opt.flow_pattern == "" or opt.occlusions_pattern == ""):
        raise SystemExit("Must give -flow_pattern and -occlusions_pattern")  # :180-182
    d = Driver(opt)
    core.run_fast_neural_video(opt, d.func_load_image, d.func_load_cert, d.func_eval, d.func_make_last_frame_warped,
                               d.func_is_single_image, d.func_save_image)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main(sys.argv[1:])
