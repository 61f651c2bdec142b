"""consistencyChecker (consistencyChecker/consistencyChecker.cpp) on the GPU.

`check(flow1, flow2, image=None)` is the in-memory op; `main(argv)` keeps the reference's argv contract
`consistencyChecker <flow1.flo> <flow2.flo> <out.pgm> [<image.ppm>]` (consistencyChecker.cpp:136-172), which
makeOptFlow_*.sh:59-60 and video_dataset/make_occlusions.sh:31-36 invoke.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np
import torch

from . import _lib, flowFileLoader


def compute_corners(image_planes: torch.Tensor, rho: float = 3.0):
    """computeCorners + normalize(0,1) + avg (consistencyChecker.cpp:39-78,158-159). image [Z,H,W] in 0..255."""
    x = image_planes.contiguous()
    Z, H, W = x.shape
    ws = torch.empty(int(_lib.lib.fav_compute_corners_workspace(Z, W, H)), dtype=torch.uint8, device=x.device)
    corners = torch.empty((H, W), dtype=torch.float32, device=x.device)
    avg = torch.empty((1,), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib.fav_compute_corners(_lib.dptr(x), Z, W, H, C.c_float(rho), _lib.dptr(corners), _lib.dptr(avg),
                                            _lib.dptr(ws), _lib.stream_ptr()))
    return corners, avg


def check(flow1_uv: torch.Tensor, flow2_uv: torch.Tensor, image_planes: torch.Tensor = None, want_cert=False):
    """checkConsistency (:80-134): flows [2,H,W] with plane 0 = u, 1 = v.  Returns u8 [H,W] in {0,255}
    (and the fp32 certainty = byte/255 when want_cert)."""
    f1, f2 = flow1_uv.contiguous(), flow2_uv.contiguous()
    assert f1.shape == f2.shape and f1.size(0) == 2  # consistencyChecker.cpp:144-145
    _, H, W = f1.shape
    rel = torch.empty((H, W), dtype=torch.uint8, device=f1.device)
    cert = torch.empty((H, W), dtype=torch.float32, device=f1.device) if want_cert else None
    structure, avg = (None, 0.0)
    if image_planes is not None:
        structure, avg_t = compute_corners(image_planes)
        avg = float(avg_t.item())
    _lib.check(_lib.lib.fav_consistency_check(_lib.dptr(f1), _lib.dptr(f2), _lib.dptr(structure), C.c_float(avg),
                                              _lib.dptr(rel), _lib.dptr(cert), W, H, _lib.stream_ptr()))
    return (rel, cert) if want_cert else rel


def read_ppm_planes(path: str) -> np.ndarray:
    """CTensor::readFromPPM (CTensor.h:888-936): binary P6 (or P5) -> [Z,H,W] float planes 0..255."""
    with open(path, "rb") as f:
        data = f.read()
    pos = data.index(b"P")
    magic = data[pos:pos + 2]
    z = {b"P5": 1, b"P6": 3}[magic]
    pos += 2
    vals = []
    while len(vals) < 3:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        vals.append(int(data[pos:end]))
        pos = end
    pos += 1
    W, H, _ = vals
    a = np.frombuffer(data, np.uint8, count=W * H * z, offset=pos).reshape(H, W, z)
    return np.ascontiguousarray(np.transpose(a, (2, 0, 1))).astype(np.float32)


def main(argv=None) -> int:
    argv = sys.argv if argv is None else argv
    assert len(argv) >= 4  # consistencyChecker.cpp:137
    dev = torch.device("cuda")
    f1 = torch.from_numpy(flowFileLoader.load(argv[1], layout=1)).to(dev)
    f2 = torch.from_numpy(flowFileLoader.load(argv[2], layout=1)).to(dev)
    img = torch.from_numpy(read_ppm_planes(argv[4])).to(dev) if len(argv) >= 5 else None
    rel = check(f1, f2, img).cpu().numpy()
    H, W = rel.shape
    with open(argv[3], "wb") as f:  # CMatrix::writeToPGM (CMatrix.h:1059-1072)
        f.write(b"P5\n%d %d\n255\n" % (W, H))
        f.write(rel.tobytes())
    sys.stdout.write(argv[3])  # :166
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
