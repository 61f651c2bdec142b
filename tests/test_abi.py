"""The C-ABI shared library loads and exports every symbol include/fav.h declares; without a GPU every compute
entry point fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "fav.h")).read()
    return sorted(set(re.findall(r"FAV_API\s+[\w\s\*]+?\b(fav_\w+)\s*\(", txt)))


def test_header_symbols_exported():
    from fav_b200 import _lib

    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in fav.h but not exported by libfav_b200.so"
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == syms


def test_every_declaration_cites_the_reference():
    txt = open(os.path.join(ROOT, "include", "fav.h")).read()
    assert txt.count(".cu:") + txt.count(".lua:") + txt.count(".cpp:") >= 15


def test_grad_entry_points_raise_like_the_reference():
    from fav_b200 import _lib

    assert _lib.lib.fav_bilinear_sampler_bdhw_update_grad_input() == _lib.FAV_ERR_NOT_IMPLEMENTED
    assert "Not implemented" in _lib.last_error()  # BilinearSamplerBDHW.cu:173
    assert _lib.lib.fav_bilinear_sampler_bdhw_update_grad_input_only_grid() == _lib.FAV_ERR_NOT_IMPLEMENTED


def test_argument_validation_without_compute():
    import ctypes as C

    from fav_b200 import _lib

    z = _lib.i64x4([1, 3, 4, 4])
    bad = _lib.i64x4([1, 3, 4, 4])  # grids:size(2) must be 2
    fake = C.c_void_p(16)
    st = _lib.lib.fav_bilinear_sampler_bdhw_update_output(fake, z, z, fake, bad, z, fake, z, 0, None)
    assert st == _lib.FAV_ERR_INVALID and "size(2)" in _lib.last_error()
    assert _lib.lib.fav_min_filter(fake, fake, 1, 4, 4, 4, None) == _lib.FAV_ERR_INVALID  # even r


def test_arch_parser_and_param_table():
    from fav_b200 import models_video, synth

    net = models_video.StyleNet(synth.DEFAULT_ARCH)
    shapes = net.param_shapes()
    w = synth.make_weights(synth.DEFAULT_ARCH, "candy")
    assert set(shapes) == set(w)
    for k, v in w.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    with pytest.raises(Exception) as e:
        models_video.StyleNet("c9s1-32,X5,c9s1-3")
    assert "not supported" in str(e.value)


def test_no_cpu_fallback_without_device():
    import torch

    from fav_b200 import _lib, models_video, synth

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    net = models_video.StyleNet()
    with pytest.raises(_lib.FavError) as e:
        net.load_state(synth.make_weights())
    assert e.value.status == _lib.FAV_ERR_NO_DEVICE
    import ctypes as C

    fake = C.c_void_p(16)
    assert _lib.lib.fav_vgg_preprocess(fake, fake, 1, 4, 4, None) == _lib.FAV_ERR_NO_DEVICE
    assert "no CPU fallback" in _lib.last_error()


def test_flo_reader_matches_oracle(tmp_path):
    from fav_b200 import flowFileLoader, synth
    from oracle import pyoracle

    uv = synth.make_backward_flow(37, 53, 2)
    p = str(tmp_path / "a.flo")
    synth.write_flo(p, uv)
    lua = flowFileLoader.load(p)  # [dy, dx]  (flowFileLoader.lua:31-32)
    assert np.array_equal(lua, pyoracle.flo_read(p, 0)) and np.array_equal(lua[0], uv[1]) and np.array_equal(lua[1], uv[0])
    chk = flowFileLoader.load(p, layout=1)
    assert np.array_equal(chk, pyoracle.flo_read(p, 1)) and np.array_equal(chk, uv)
    open(str(tmp_path / "trunc.flo"), "wb").write(open(p, "rb").read()[:100])
    with pytest.raises(Exception):
        flowFileLoader.load(str(tmp_path / "trunc.flo"))
    with pytest.raises(Exception):
        flowFileLoader.load(str(tmp_path / "missing.flo"))
