"""fast_artistic_video/preprocess.lua: preprocess.vgg.preprocess (:57-62) / deprocess (:66-71)."""
from __future__ import annotations

import torch

from . import _lib


def _check_input(img):  # preprocess.lua:7-10
    assert img.dim() == 4, "img must be N x C x H x W"
    assert img.size(1) == 3, "img must have three channels"


class vgg:
    @staticmethod
    def preprocess(img: torch.Tensor) -> torch.Tensor:
        _check_input(img)
        x = img.contiguous()
        out = torch.empty_like(x)
        _lib.check(_lib.lib.fav_vgg_preprocess(_lib.dptr(x), _lib.dptr(out), x.size(0), x.size(2), x.size(3),
                                               _lib.stream_ptr()))
        return out

    @staticmethod
    def deprocess(img: torch.Tensor) -> torch.Tensor:
        _check_input(img)
        x = img.contiguous()
        out = torch.empty_like(x)
        _lib.check(_lib.lib.fav_vgg_deprocess(_lib.dptr(x), _lib.dptr(out), x.size(0), x.size(2), x.size(3),
                                              _lib.stream_ptr()))
        return out
