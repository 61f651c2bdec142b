"""`require 'stn'` -> nn.BilinearSamplerBDHW  (stnbdhw/init.lua:1-11, stnbdhw/BilinearSamplerBDHW.lua:1-118).

Same operator surface as the reference's Lua module: `BilinearSamplerBDHW():forward({inputImages, grids})`.
The native call it wraps is fav_bilinear_sampler_bdhw_update_output (include/fav.h), which replaces
cunn_BilinearSamplerBDHW_updateOutput (stnbdhw/BilinearSamplerBDHW.cu:111-152).
"""
from __future__ import annotations

import torch

from . import _lib


class BilinearSamplerBDHW:
    """nn.BilinearSamplerBDHW (BilinearSamplerBDHW.lua:21-82).  Grids hold pixel offsets: channel 0 = dy, 1 = dx
    (BilinearSamplerBDHW.cu:72-73; the docstring in the .lua about normalised coordinates is stale)."""

    def __init__(self, border_mode: int = _lib.BORDER_PER_TAP):
        self.output = None
        self.gradInput = {}
        self.border_mode = border_mode

    # BilinearSamplerBDHW.lua:26-42
    @staticmethod
    def check(inputImages, grids, gradOutput=None):
        assert inputImages.is_contiguous(), "Input images have to be contiguous"
        assert grids.is_contiguous(), "Grids have to be contiguous"
        assert inputImages.dim() == 4
        assert grids.dim() == 4
        assert inputImages.size(0) == grids.size(0)  # batch
        assert grids.size(1) == 2  # coordinates
        if gradOutput is not None:
            assert grids.size(0) == gradOutput.size(0)
            assert grids.size(2) == gradOutput.size(2)
            assert grids.size(3) == gradOutput.size(3)

    def updateOutput(self, input):
        _inputImages, _grids = input
        if _inputImages.dim() == 3:  # addOuterDim, :44-65
            inputImages, grids = _inputImages.unsqueeze(0), _grids.unsqueeze(0)
        else:
            inputImages, grids = _inputImages, _grids
        self.check(inputImages, grids)
        assert inputImages.dtype == torch.float32 and grids.dtype == torch.float32
        out = torch.empty((inputImages.size(0), inputImages.size(1), grids.size(2), grids.size(3)),
                          dtype=torch.float32, device=inputImages.device)  # self.output:resize(...) :71
        _lib.check(_lib.lib.fav_bilinear_sampler_bdhw_update_output(
            _lib.dptr(inputImages), _lib.i64x4(inputImages.shape), _lib.i64x4(inputImages.stride()),
            _lib.dptr(grids), _lib.i64x4(grids.shape), _lib.i64x4(grids.stride()),
            _lib.dptr(out), _lib.i64x4(out.stride()), self.border_mode, _lib.stream_ptr()))
        self.output = out[0] if _inputImages.dim() == 3 else out  # :77-79
        return self.output

    forward = updateOutput

    def updateGradInput(self, _input, _gradOutput):
        # BilinearSamplerBDHW.cu:171-176: "Not implemented" -> THError
        _lib.check(_lib.lib.fav_bilinear_sampler_bdhw_update_grad_input())

    backward = updateGradInput
