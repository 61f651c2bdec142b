"""VR (cube-map) path: host-side map makers vs the literal Lua-loop restatement (CPU), GPU median / fused border blend /
driver (gpu)."""
import os

import numpy as np
import pytest

from fav_b200 import synth, vr_helper
from oracle import vr_oracle


@pytest.mark.parametrize("size,overlap", [(64, 20), (96, 32), (100, 20)])
def test_perspective_maps_match_literal_restatement(size, overlap):
    assert np.array_equal(vr_helper.make_perspective_warp_map_left(size, overlap, size), vr_oracle.perspective_left(size, overlap, size))
    assert np.array_equal(vr_helper.make_perspective_warp_map_right(size, overlap, size), vr_oracle.perspective_right(size, overlap, size))
    assert np.array_equal(vr_helper.make_perspective_warp_map_top(size, overlap, size), vr_oracle.perspective_top(size, overlap, size))
    assert np.array_equal(vr_helper.make_perspective_warp_map_bottom(size, overlap, size), vr_oracle.perspective_bottom(size, overlap, size))
    m = vr_helper.make_perspective_warp_map_left(size, overlap, size)
    assert (m[:, :, : size - overlap] == 99999).all() and (m[:, :, size - overlap:] != 99999).all()  # sentinel outside the strip


def test_equirect_map_matches_literal_restatement():
    a = vr_helper.make_cube_to_equirectangular_map(62, 62, 19, 19, 48, 40)
    b = vr_oracle.cube_to_equirect(62, 62, 19, 19, 48, 40)
    assert np.array_equal(a, b)


def test_gradient_masks():
    assert np.allclose(vr_helper.make_gradient_mask_h_inc(4), [0.2, 0.4, 0.6, 0.8])
    assert np.allclose(vr_helper.make_gradient_mask_w_dec(4), [0.8, 0.6, 0.4, 0.2])


# ---------------------------------------------------------------------------------------------------------------
def T(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("r", [3, 5])
def test_median_filter_gpu_bit_exact(r):
    from fav_b200 import utils

    img = np.random.default_rng(0).uniform(size=(3, 37, 53)).astype(np.float32)
    img[:, 5:9, 5:9] = 0.5  # ties
    assert np.array_equal(utils.median_filter(T(img), r).cpu().numpy(), vr_oracle.median_filter(img, r))


@pytest.mark.gpu
def test_fused_border_blend_bit_exact():
    import ctypes as C

    import torch

    from fav_b200 import _lib

    S, ov = 64, 20
    rng = np.random.default_rng(1)
    base = rng.uniform(size=(3, S, S)).astype(np.float32)
    sides = [rng.uniform(size=(3, S, S)).astype(np.float32) for _ in range(4)]
    maps = [vr_helper.make_perspective_warp_map_right(S, ov, S).astype(np.float32),
            vr_helper.make_perspective_warp_map_left(S, ov, S).astype(np.float32),
            vr_helper.make_perspective_warp_map_bottom(S, ov, S).astype(np.float32),
            vr_helper.make_perspective_warp_map_top(S, ov, S).astype(np.float32)]
    rots = [0, 1, 2, 3]
    div = np.clip(rng.integers(0, 3, size=(S, S)), 1, None).astype(np.float32)
    mask = rng.uniform(size=(S, S)).astype(np.float32)
    ref = vr_oracle.blend_sides(base, sides, maps, rots, div, mask)
    tb, ts, tm, td, tk = T(base), [T(s) for s in sides], [T(m) for m in maps], T(div), T(mask)
    out = torch.empty_like(tb)
    _lib.check(_lib.lib.fav_vr_blend_sides(_lib.dptr(tb), (C.c_void_p * 4)(*[t.data_ptr() for t in ts]),
                                           (C.c_void_p * 4)(*[t.data_ptr() for t in tm]), (C.c_int * 4)(*rots),
                                           _lib.dptr(td), _lib.dptr(tk), _lib.dptr(out), S, _lib.stream_ptr()))
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.gpu
def test_vr_driver_two_frames(tmp_path):
    """2 VR frames x 6 faces through the cube-map driver (fast_artistic_video_vr.lua): border priors, flow-warped prior
    blend, fused re-blend, median, cube map + equirectangular output."""
    import torch

    from fav_b200 import models_video, vr

    S, ov = 64, 20
    d = str(tmp_path)
    for f in (1, 2):
        for face in range(1, 7):
            synth.write_ppm(f"{d}/in_{f:03d}_{face}.ppm", synth.make_frame(S, S, 10 * f + face))
    for face in range(1, 7):
        synth.write_flo(f"{d}/bw_2_1_{face}.flo", synth.make_backward_flow(S, S, face + 1))
        rel = (np.random.default_rng(face).uniform(size=(S, S)) > 0.1).astype(np.uint8) * 255
        with open(f"{d}/rel_2_1_{face}.pgm", "wb") as f:
            f.write(b"P5\n%d %d\n255\n" % (S, S) + rel.tobytes())
    net = models_video.synthetic_model("mosaic", synth.PAPER_ARCH)
    drv = vr.main(["-input_pattern", f"{d}/in_%03d_%d.ppm", "-flow_pattern", f"{d}/bw_[%d]_{{%d}}_%d.flo",
                   "-occlusions_pattern", f"{d}/rel_[%d]_{{%d}}_%d.pgm", "-output_prefix", f"{d}/out",
                   "-overlap_pixel_w", str(ov), "-overlap_pixel_h", str(ov), "-num_frames", "2", "-out_equi", "-out_cubemap",
                   "-out_equi_w", "96", "-out_equi_h", "48"], model_vid=net)
    assert sorted(drv.outputs) == [1, 2]
    for k in (1, 2):
        assert tuple(drv.outputs[k]["equi"].shape) == (3, 48, 96)
        assert tuple(drv.outputs[k]["cubemap"].shape) == (3, S - ov, 6 * (S - ov))
        assert torch.isfinite(drv.outputs[k]["cubemap"]).all()
    assert os.path.exists(f"{d}/out-00002_equi.png") and os.path.exists(f"{d}/out-00002_cubemap.png")
    # the re-blended faces differ from the raw stylized faces only inside the border strips
    f1 = drv.prev_last_segments[1]
    assert float((f1 - drv.last_segments[1])[:, ov:-ov, ov:-ov].abs().max()) == 0.0
