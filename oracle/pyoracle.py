"""ctypes binding of oracle/liboracle.so (fav_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
REF_CHECKER = os.path.join(_HERE, "_ref", "consistencyChecker")


def build(force: bool = False) -> None:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "fav_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    if os.path.isdir("/root/reference/consistencyChecker") and not os.path.exists(REF_CHECKER):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        _LIB.orc_avg.restype = C.c_float
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64x4(t):
    return (C.c_int64 * 4)(*t)


def warp_bdhw(img: np.ndarray, grid: np.ndarray, threads: int = 1) -> np.ndarray:
    """nn.BilinearSamplerBDHW forward (BilinearSamplerBDHW.cu:48-109). img BxCxHxW, grid Bx2xHoxWo."""
    img, grid = _f32(img), _f32(grid)
    squeeze = img.ndim == 3  # BilinearSamplerBDHW.lua:59-65
    if squeeze:
        img, grid = img[None], grid[None]
    B, Cc, Hin, Win = img.shape
    _, two, Ho, Wo = grid.shape
    assert two == 2 and grid.shape[0] == B
    out = np.empty((B, Cc, Ho, Wo), np.float32)
    es = lambda a: _i64x4([s // 4 for s in a.strides])
    lib().orc_warp_bdhw(_fp(img), es(img), _fp(grid), es(grid), _fp(out), es(out), B, Cc, Hin, Win, Ho, Wo,
                        int(threads))
    return out[0] if squeeze else out


def image_warp_pad(img, flow, pad=0.0):
    img, flow = _f32(img), _f32(flow)
    Cc, H, W = img.shape
    out = np.empty_like(img)
    lib().orc_image_warp_pad(_fp(img), _fp(flow), _fp(out), Cc, H, W, C.c_float(pad))
    return out


def vgg_preprocess(img):
    img = _f32(img)
    out = np.empty_like(img)
    lib().orc_vgg_preprocess(_fp(img), _fp(out), C.c_int64(img.shape[-1] * img.shape[-2]))
    return out


def vgg_deprocess(img):
    img = _f32(img)
    out = np.empty_like(img)
    lib().orc_vgg_deprocess(_fp(img), _fp(out), C.c_int64(img.shape[-1] * img.shape[-2]))
    return out


def min_filter(cert, r=7):
    cert = _f32(cert)
    H, W = cert.shape[-2:]
    out = np.empty_like(cert)
    lib().orc_min_filter(_fp(cert), _fp(out), H, W, int(r))
    return out


def temporal_input(content, prev, flow, cert, fill=None, flow_mask=None, warp_mode=0):
    content, prev, flow, cert = _f32(content), _f32(prev), _f32(flow), _f32(cert)
    _, H, W = content.shape
    out = np.empty((7, H, W), np.float32)
    fill = None if fill is None else _f32(fill)
    flow_mask = None if flow_mask is None else _f32(flow_mask)
    lib().orc_temporal_input(_fp(content), _fp(prev), _fp(flow), _fp(cert),
                             _fp(fill) if fill is not None else None,
                             _fp(flow_mask) if flow_mask is not None else None, _fp(out), H, W, int(warp_mode))
    return out


def first_frame_input(content, fill=None):
    content = _f32(content)
    _, H, W = content.shape
    out = np.empty((7, H, W), np.float32)
    fill = None if fill is None else _f32(fill)
    lib().orc_first_frame_input(_fp(content), _fp(fill) if fill is not None else None, _fp(out), H, W)
    return out


def flo_read(path: str, layout: int = 0) -> np.ndarray:
    W, H = C.c_int(), C.c_int()
    if lib().orc_flo_header(path.encode(), C.byref(W), C.byref(H)) != 0:
        raise IOError(path)
    out = np.empty((2, H.value, W.value), np.float32)
    if lib().orc_flo_read(path.encode(), _fp(out), int(layout)) != 0:
        raise IOError(path)
    return out


def compute_corners(image_planes, rho=3.0, normalize=True):
    image_planes = _f32(image_planes)
    Z, H, W = image_planes.shape
    out = np.empty((H, W), np.float32)
    lib().orc_compute_corners(_fp(image_planes), Z, W, H, C.c_float(rho), _fp(out), int(normalize))
    return out


def consistency(flow1_uv, flow2_uv, image_planes=None) -> np.ndarray:
    """consistencyChecker main() on in-memory data -> u8 HxW in {0,255}."""
    f1, f2 = _f32(flow1_uv), _f32(flow2_uv)
    _, H, W = f1.shape
    out = np.empty((H, W), np.uint8)
    if image_planes is not None:
        im = _f32(image_planes)
        lib().orc_consistency_main(_fp(f1), _fp(f2), _fp(im), im.shape[0], W, H, out.ctypes.data_as(C.c_void_p))
    else:
        lib().orc_consistency_main(_fp(f1), _fp(f2), None, 0, W, H, out.ctypes.data_as(C.c_void_p))
    return out


def run_ref_checker(flow1_path, flow2_path, out_pgm, image_ppm=None) -> None:
    """Run the reference's own compiled binary (oracle/_ref/consistencyChecker)."""
    args = [REF_CHECKER, flow1_path, flow2_path, out_pgm] + ([image_ppm] if image_ppm else [])
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)


def num_threads() -> int:
    return int(lib().orc_num_threads())
