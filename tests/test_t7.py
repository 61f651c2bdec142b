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
