"""flowFileLoader.lua:17-37 -- `flowFile.load(path)` -> 2xHxW with channel 0 = dy (v), channel 1 = dx (u)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def load(fileName: str, layout: int = 0, out=None) -> np.ndarray:
    """out: optional preallocated (e.g. pinned) float32 array / tensor buffer of at least 2*H*W elements."""
    W, H = C.c_int(), C.c_int()
    _lib.check(_lib.lib.fav_flo_read_header(fileName.encode(), C.byref(W), C.byref(H)))
    if out is None:
        out = np.empty((2, H.value, W.value), np.float32)
    # the capacity travels with the call: a file rewritten since the header read cannot overflow the buffer
    _lib.check(_lib.lib.fav_flo_read(fileName.encode(), out.ctypes.data_as(C.c_void_p), out.size, int(layout)))
    return out
