"""flowFileLoader.lua:17-37 -- `flowFile.load(path)` -> 2xHxW with channel 0 = dy (v), channel 1 = dx (u)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def load(fileName: str, layout: int = 0) -> np.ndarray:
    W, H = C.c_int(), C.c_int()
    _lib.check(_lib.lib.fav_flo_read_header(fileName.encode(), C.byref(W), C.byref(H)))
    out = np.empty((2, H.value, W.value), np.float32)
    _lib.check(_lib.lib.fav_flo_read(fileName.encode(), out.ctypes.data_as(C.c_void_p), int(layout)))
    return out
