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
                                           _lib.dptr(td), _lib.dptr(tk), None, _lib.dptr(out), S, _lib.stream_ptr()))
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
        assert tuple(drv.outputs[k]["cubemap"].shape) == (3, S - ov + 2, 6 * (S - ov + 2))  # reference crop (:548-553)
        assert torch.isfinite(drv.outputs[k]["cubemap"]).all()
    assert os.path.exists(f"{d}/out-00002_equi.png") and os.path.exists(f"{d}/out-00002_cubemap.png")
    # the re-blended faces differ from the raw stylized faces only inside the border strips
    f1 = drv.prev_last_segments[1]
    assert float((f1 - drv.last_segments[1])[:, ov:-ov, ov:-ov].abs().max()) == 0.0


# ---- the cube-map DRIVER vs its statement-by-statement restatement (oracle/vr_oracle.py: VRRef, run_vr_clip) --------
def _vr_clip_files(d, S, n_frames, seed0=0):
    """n_frames VR frames x 6 faces as 8-bit PPMs + per-face backward flow / occlusion PGM for frames >= 2.
    Returns the same data as the arrays the reference would hold after image.load / flowFile.load."""
    frames, flows, certs = [], [], []
    for f in range(1, n_frames + 1):
        fr, fl, ce = {}, {}, {}
        for face in range(1, 7):
            img = synth.make_frame(S, S, 10 * f + face + seed0)
            synth.write_ppm(f"{d}/in_{f:03d}_{face}.ppm", img)
            fr[face] = np.clip(np.rint(img * 255.0), 0, 255).astype(np.uint8).astype(np.float32) / 255.0
            if f >= 2:
                bw = synth.make_backward_flow(S, S, face + f)
                synth.write_flo(f"{d}/bw_{f}_{f - 1}_{face}.flo", bw)
                fl[face] = synth.checker_to_lua(bw)
                rel = (np.random.default_rng(100 * f + face).uniform(size=(S, S)) > 0.1).astype(np.uint8) * 255
                with open(f"{d}/rel_{f}_{f - 1}_{face}.pgm", "wb") as fh:
                    fh.write(b"P5\n%d %d\n255\n" % (S, S) + rel.tobytes())
                ce[face] = (rel.astype(np.float32) / 255.0)[None]
        frames.append(fr); flows.append(fl); certs.append(ce)
    return frames, flows, certs


def _vr_argv(d, ov, n, extra=()):
    return ["-input_pattern", f"{d}/in_%03d_%d.ppm", "-flow_pattern", f"{d}/bw_[%d]_{{%d}}_%d.flo",
            "-occlusions_pattern", f"{d}/rel_[%d]_{{%d}}_%d.pgm", "-output_prefix", f"{d}/out", "-overlap_pixel_w", str(ov),
            "-overlap_pixel_h", str(ov), "-num_frames", str(n), "-out_equi", "-out_cubemap", "-out_equi_w", "96",
            "-out_equi_h", "48", *extra]


@pytest.mark.gpu
@pytest.mark.parametrize("smooth", [False, True])
def test_vr_driver_logic_bit_exact_vs_restatement(tmp_path, smooth):
    """Driver logic in isolation: the SAME trivial 'network' (exact fp32 elementwise ops) is injected on both sides, so the
    border priors, mask algebra, flow-warped prior blend (:239-302), fused re-blend (:454-509), median, the reference's
    cube-map crop (:548-553) and the equirectangular warp (:543) must agree BIT FOR BIT over 3 VR frames x 6 faces."""
    import torch

    from fav_b200 import vr

    class FakeGpuNet:
        def run_image(self, img, fill=None):
            return (img * 0.5).contiguous()

        def run_next_image(self, img, prev, flow, cert, fill=None, flow_mask=None, border_mode=0):
            assert float(flow.abs().max()) == 0.0  # the VR callbacks hand over an already blended prior
            m = cert if flow_mask is None else torch.minimum(cert, flow_mask)
            return (img * 0.25 + (prev * m) * 0.75).contiguous()

    class FakeRefNet:
        def run_image(self, img):
            return (img * np.float32(0.5)).astype(np.float32)

        def run_next_image_prior(self, img, prior, cert, fmask):
            m = cert if fmask is None else np.minimum(cert, fmask)
            return ((img * np.float32(0.25)).astype(np.float32) + ((prior * m[None]).astype(np.float32) * np.float32(0.75))
                    .astype(np.float32)).astype(np.float32)

    S, ov, n = 64, 20, 3
    d = str(tmp_path)
    frames, flows, certs = _vr_clip_files(d, S, n)
    # -smooth_certainty indexes a nil gradMask for face 6 unless every face 6 is a single image (:245,297,304-310)
    drv = vr.main(_vr_argv(d, ov, n, ("-smooth_certainty", "-create_inconsistent") if smooth else ()), model_vid=FakeGpuNet())
    ref = vr_oracle.VRRef(S, S, ov, ov, 3, 96, 48, smooth_certainty=smooth)
    outs = vr_oracle.run_vr_clip(ref, FakeRefNet(), frames, flows, certs, create_inconsistent=smooth)
    assert sorted(drv.outputs) == [1, 2, 3] and len(outs) == 3
    for k in range(3):
        g = drv.outputs[k + 1]
        assert tuple(g["cubemap"].shape) == outs[k]["cubemap"].shape == (3, S - ov + 2, 6 * (S - ov + 2))  # 6(h+2) x (h+2)
        assert np.array_equal(g["cubemap"].cpu().numpy(), outs[k]["cubemap"]), k
        assert np.array_equal(g["equi"].cpu().numpy(), outs[k]["equi"]), k
    for face in range(1, 7):
        assert np.array_equal(drv.prev_last_segments[face].cpu().numpy(), ref.prev_last_segments[face]), face


@pytest.mark.gpu
def test_vr_driver_with_real_net_vs_fp64_oracle(tmp_path):
    """cfg 4 in miniature: 2 VR frames x 6 faces, paper arch / mosaic weights, GPU path vs the driver restatement running the
    fp64 net oracle; the error of a face feeds the priors of later faces, so the bound is the north-star 1e-3 (logged)."""
    import torch

    from fav_b200 import models_video, vr
    from oracle import net_oracle, pyoracle

    class OraNet:
        def __init__(self):
            self.o = net_oracle.NetOracle(arch=synth.PAPER_ARCH, style="mosaic", dtype=torch.float64)

        def run_image(self, img):
            with torch.no_grad():
                return self.o.run_image(img).astype(np.float32)

        def run_next_image_prior(self, img, prior, cert, fmask):
            H, W = img.shape[-2:]
            x7 = pyoracle.temporal_input(img, prior, np.zeros((2, H, W), np.float32), cert, None, fmask)
            with torch.no_grad():
                return self.o.deprocess(self.o.forward(torch.from_numpy(x7)[None]))[0].numpy().astype(np.float32)

    S, ov, n = 64, 20, 2
    d = str(tmp_path)
    frames, flows, certs = _vr_clip_files(d, S, n, seed0=3)
    net = models_video.synthetic_model("mosaic", synth.PAPER_ARCH)
    drv = vr.main(_vr_argv(d, ov, n), model_vid=net)
    ref = vr_oracle.VRRef(S, S, ov, ov, 3, 96, 48)
    outs = vr_oracle.run_vr_clip(ref, OraNet(), frames, flows, certs)
    worst = 0.0
    for k in range(n):
        for key in ("cubemap", "equi"):
            worst = max(worst, float(np.abs(drv.outputs[k + 1][key].cpu().numpy() - outs[k][key]).max()))
    print(f"VR 2x6 faces: max-abs vs fp64 driver oracle {worst:.3e}")
    assert worst < 1e-3 and worst < 2e-4, worst


def test_vr_continue_with_is_rejected_like_the_reference_would_fail():
    from fav_b200 import _lib, vr

    with pytest.raises(_lib.FavError) as e:
        vr.main(["-input_pattern", "x_%d_%d.ppm", "-create_inconsistent", "-continue_with", "2"])
    assert e.value.status == _lib.FAV_ERR_UNSUPPORTED
