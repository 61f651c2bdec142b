"""fast_artistic_video/utils.lua: warp_image (:141-149), min_filter (:161-169), wait_for_file (:74-80)."""
from __future__ import annotations

import os
import time

import torch

from . import _lib
from .stn import BilinearSamplerBDHW


def warp_image(img: torch.Tensor, map: torch.Tensor, dtype: str = "torch.CudaTensor") -> torch.Tensor:
    """utils.warp_image(img, map, dtype) (utils.lua:141-149).

    dtype 'torch.CudaTensor' -> the reference's CUDA semantics (per-corner zero fill, BilinearSamplerBDHW.cu);
    dtype 'torch.FloatTensor' -> the reference's CPU semantics (image.warp(...,'bilinear',true,'pad',0)).
    Both run on the GPU here; only the border rule differs.
    """
    mode = _lib.BORDER_PER_TAP if dtype == "torch.CudaTensor" else _lib.BORDER_PAD_PIXEL
    return BilinearSamplerBDHW(mode).forward((img.contiguous(), map.contiguous()))


def min_filter(batch: torch.Tensor, r: int, dtype=None) -> torch.Tensor:
    """utils.min_filter(batch, r, dtype) (utils.lua:161-169): 1 - maxpool_{r x r, s1, pad r//2}(1 - x)."""
    x = batch.contiguous()
    assert x.dtype == torch.float32 and x.dim() >= 2
    H, W = x.shape[-2:]
    n = x.numel() // (H * W)
    out = torch.empty_like(x)
    _lib.check(_lib.lib.fav_min_filter(_lib.dptr(x), _lib.dptr(out), n, H, W, int(r), _lib.stream_ptr()))
    return out


def temporal_loss(prev_stylized: torch.Tensor, stylized: torch.Tensor, flow: torch.Tensor, cert: torch.Tensor,
                  dtype: str = "torch.CudaTensor") -> float:
    """The temporal term of -evaluate (fast_artistic_video.lua:128-151):
    nn.MSECriterion(cmul(warp_image(prev_stylized, flow), cert), cmul(stylized, cert)) as ONE fused kernel."""
    p, c, f, m = prev_stylized.contiguous(), stylized.contiguous(), flow.contiguous(), cert.contiguous().reshape(-1)
    H, W = c.shape[-2:]
    acc = torch.zeros((1,), dtype=torch.float64, device=c.device)
    mode = _lib.BORDER_PER_TAP if dtype == "torch.CudaTensor" else _lib.BORDER_PAD_PIXEL
    _lib.check(_lib.lib.fav_temporal_mse(_lib.dptr(p), _lib.dptr(c), _lib.dptr(f), _lib.dptr(m), H, W, mode, _lib.dptr(acc),
                                         _lib.stream_ptr()))
    return float(acc.item()) / (3.0 * H * W)


def bytes_to_planes(rgb_hwc: torch.Tensor, flo_uv: torch.Tensor = None, cert8: torch.Tensor = None, invert_occlusion: bool = False):
    """Device-side image.load / flowFile.load / func_load_cert: uint8 [H,W,3] -> float [3,H,W] = byte / 255, float [H,W,2] (u,v)
    pairs -> [2,H,W] = (dy,dx), uint8 [H,W] -> float [1,H,W] = byte / 255 (1 - that with invert_occlusion).  One kernel."""
    H, W = rgb_hwc.shape[:2]
    assert rgb_hwc.dtype == torch.uint8 and rgb_hwc.is_contiguous()
    content = torch.empty((3, H, W), dtype=torch.float32, device=rgb_hwc.device)
    flow = torch.empty((2, H, W), dtype=torch.float32, device=rgb_hwc.device) if flo_uv is not None else None
    cert = torch.empty((1, H, W), dtype=torch.float32, device=rgb_hwc.device) if cert8 is not None else None
    _lib.check(_lib.lib.fav_bytes_to_planes(_lib.dptr(rgb_hwc), _lib.dptr(flo_uv.contiguous() if flo_uv is not None else None),
                                            _lib.dptr(cert8.contiguous() if cert8 is not None else None), 1 if invert_occlusion else 0,
                                            _lib.dptr(content), _lib.dptr(flow), _lib.dptr(cert), H, W, _lib.stream_ptr()))
    return content, flow, cert


def planes_to_png_rows(img: torch.Tensor) -> torch.Tensor:
    """Device-side image.save quantisation: float [3,H,W] -> uint8 [H, 1+3W] Sub-filtered PNG scanlines (filter byte 1)."""
    H, W = img.shape[-2:]
    rows = torch.empty((H, 1 + 3 * W), dtype=torch.uint8, device=img.device)
    _lib.check(_lib.lib.fav_planes_to_png_rows(_lib.dptr(img.contiguous()), _lib.dptr(rows), H, W, _lib.stream_ptr()))
    return rows


def bind_host_to_gpu_numa(device_index: int = 0) -> dict:
    """Restrict this process (threads created later inherit it) to the CPUs of the NUMA node the GPU hangs off, BEFORE pinned
    buffers are allocated: first-touch then places them in that node's memory and host<->device copies do not cross the
    socket interconnect.  Returns what was found / done; a box without NUMA information is left alone."""
    info = dict(numa_node=None, cpus_before=None, cpus_after=None)
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cur = os.sched_getaffinity(0)
        info["cpus_before"] = len(cur)
        both = cur & cpus
        if both:
            os.sched_setaffinity(0, both)
        info["cpus_after"] = len(os.sched_getaffinity(0))
    except (OSError, ValueError, AttributeError, RuntimeError):
        pass
    return info


def file_exists(name: str) -> bool:  # utils.lua:68-71
    return os.path.isfile(name)


def wait_for_file(path: str, poll_s: float = 1.0) -> None:
    """utils.wait_for_file (utils.lua:74-80): block until the producer (flow / occlusion job) wrote the file."""
    if not file_exists(path):
        print(f'Waiting for file "{path}"')
        while not file_exists(path):
            time.sleep(poll_s)
        time.sleep(poll_s)


def median_filter(img: torch.Tensor, r: int) -> torch.Tensor:
    """utils.median_filter(img, r) (utils.lua:151-159): r x r windows (unfold), lower median, VALID region
    (H-r+1) x (W-r+1) -- on the GPU (the reference casts to a CPU FloatTensor for it)."""
    x = img.contiguous()
    assert x.dim() == 3 and x.dtype == torch.float32
    C, H, W = x.shape
    out = torch.empty((C, H - r + 1, W - r + 1), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib.fav_median_filter(_lib.dptr(x), _lib.dptr(out), C, H, W, int(r), _lib.stream_ptr()))
    return out
