"""Torch7 .t7 checkpoint reader (fav_b200/t7.py): round trip through the writer, arch recovery from the module tree."""
import numpy as np
import pytest

from fav_b200 import synth, t7


@pytest.mark.parametrize("arch", [synth.DEFAULT_ARCH, synth.PAPER_ARCH])
def test_t7_round_trip_recovers_arch_and_weights(tmp_path, arch):
    w = synth.make_weights(arch, "scream")
    p = str(tmp_path / "checkpoint-scream-video.t7")
    t7.write_checkpoint(p, arch, w, tanh_constant=150.0, reflect_pad=40)
    got_arch, state, tanh_c, pad = t7.load_checkpoint(p)
    assert got_arch == arch and tanh_c == 150.0 and pad == 40
    assert set(state) == set(w)
    for k in w:
        assert state[k].dtype == np.float32 and np.array_equal(state[k], w[k]), k
    raw = t7.load(p)
    assert raw["iter"] == 60000 and raw["opt"]["padding_type"] == "reflect-start"
    assert raw["model"].torch_type == "nn.Sequential"


def test_t7_back_references_and_strided_tensors(tmp_path):
    w = t7._Writer()
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    shared = {"x": 1.5, "ok": True, "name": "candy"}
    w.obj({"a": a, "t1": shared, "n": None})
    p = tmp_path / "x.t7"
    p.write_bytes(bytes(w.b))
    r = t7.load(str(p))
    assert np.array_equal(r["a"], a) and r["t1"] == shared and r["n"] is None


def test_t7_hand_assembled_fixture():
    """tests/golden/tiny_video_model.t7 was assembled byte by byte by tests/golden/make_t7_fixture.py (no fav_b200 import)
    with what real checkpoints hold and our own writer never emits: all parameters as views into ONE flat storage at their
    own storageOffset (getParameters()), empty / Double tensors, number keys, a back-referenced table."""
    import os

    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_video_model.t7")
    geo = {}
    arch, state, tanh_c, pad = t7.load_checkpoint(p, geo)
    assert arch == "c3s1-4,R4,c3s1-3" and tanh_c == 150.0 and pad == 4
    assert geo["padding_type"] == "reflect-start" and geo["has_reflect_pad"]
    assert geo["conv_pads"] == [("l0", 1, 1), ("l1.c1", 0, 0), ("l1.c2", 0, 0), ("l2", 1, 1)]
    shapes = [("l0.weight", (4, 7, 3, 3)), ("l0.bias", (4,)), ("l0.n.weight", (4,)), ("l0.n.bias", (4,)),
              ("l1.c1.weight", (4, 4, 3, 3)), ("l1.c1.bias", (4,)), ("l1.n1.weight", (4,)), ("l1.n1.bias", (4,)),
              ("l1.c2.weight", (4, 4, 3, 3)), ("l1.c2.bias", (4,)), ("l1.n2.weight", (4,)), ("l1.n2.bias", (4,)),
              ("l2.weight", (3, 4, 3, 3)), ("l2.bias", (3,))]
    assert set(state) == {n for n, _ in shapes}
    off = 0
    for name, shp in shapes:  # value = 0.001 * flat index - 0.2 (closed form, independent of the generator)
        n = int(np.prod(shp))
        want = (0.001 * np.arange(off, off + n, dtype=np.float64) - 0.2).astype(np.float32).reshape(shp)
        assert state[name].shape == shp and np.array_equal(state[name], want), name
        off += n
    ck = t7.load(p)
    assert ck["opt"] is ck["opt_again"] and ck["train_loss_history"] == {1: 12.5, 2: 11.25} and ck["iter"] == 60000
    conv0 = ck["model"]["modules"][2]
    assert conv0.torch_type == "nn.SpatialConvolution" and conv0["gradBias"].size == 0 and conv0["output"].dtype == np.float64


def test_t7_geometry_is_validated(tmp_path):
    """a checkpoint whose paddings differ from what its arch tokens imply must be rejected, not run with another
    geometry; padding_type 'zero' checkpoints are recognised from their residual blocks."""
    arch = "c9s1-8,d16,R16,u8,c9s1-3"
    w = synth.make_weights(arch, "candy")
    p = str(tmp_path / "zero.t7")
    t7.write_checkpoint(p, arch, w, padding_type="zero")
    geo = {}
    got_arch, state, _, pad = t7.load_checkpoint(p, geo)
    assert got_arch == arch and pad == 0 and geo["padding_type"] == "zero" and not geo["has_reflect_pad"]
    # conv padding that disagrees with the token (a 'none'-padded first conv)
    ck = t7.load(p)
    ck["model"]["modules"][1]["padW"] = 0.0
    with pytest.raises(t7.GeometryError):
        t7.model_to_state(ck["model"])
    # residual blocks of padding_type 'reflect' carry their own padding modules: rejected
    p2 = str(tmp_path / "rs.t7")
    t7.write_checkpoint(p2, arch, w, reflect_pad=4)
    ck = t7.load(p2)
    res = next(m for m in ck["model"]["modules"].values() if m.torch_type == "nn.Sequential")
    block = res["modules"][1]["modules"][1]
    block["modules"] = {1: t7.T7Object("nn.SpatialReflectionPadding", {"pad_l": 1, "pad_r": 1, "pad_t": 1, "pad_b": 1}),
                        **{k + 1: v for k, v in block["modules"].items()}}
    with pytest.raises(t7.GeometryError):
        t7.model_to_state(ck["model"])


def test_load_model_rejects_mismatched_reflection_pad(tmp_path):
    """core.load_model: the checkpoint's leading SpatialReflectionPadding must be the one its arch implies
    (train_video.lua:319-324), otherwise FAV_ERR_UNSUPPORTED -- never a silent run with another geometry."""
    from fav_b200 import _lib, core

    arch = "c9s1-8,d16,R16,u8,c9s1-3"
    w = synth.make_weights(arch, "candy")
    p = str(tmp_path / "a.t7")
    t7.write_checkpoint(p, arch, w, reflect_pad=7)
    with pytest.raises(_lib.FavError) as e:
        core.load_model(p)
    assert e.value.status == _lib.FAV_ERR_UNSUPPORTED and "SpatialReflectionPadding(7)" in e.value.message
