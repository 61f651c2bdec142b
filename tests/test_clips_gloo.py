"""N>1 host logic on CPU: world_size-2 gloo run of the clip sharding (assign / scatter / process / gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys

    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from fav_b200 import clips

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 5
        shapes = [(2 + c, 3, 8, 12) for c in range(n)]  # ragged clip lengths
        data = [torch.arange(torch.Size(s).numel(), dtype=torch.float32).reshape(s) + 1000 * c
                for c, s in enumerate(shapes)] if rank == 0 else None
        mine = clips.scatter_clips(data, shapes)
        assert sorted(mine) == clips.assign_clips(n, world)[rank]
        # "stylize": a recurrent per-clip loop that depends only on the clip itself (stand-in for the GPU frame loop)
        outs = {}
        for c, t in mine.items():
            prev = torch.zeros_like(t[0])
            frames = []
            for i in range(t.shape[0]):
                prev = 0.5 * prev + t[i]
                frames.append(prev)
            outs[c] = torch.stack(frames)
        res = clips.gather_clips(outs, shapes)
        if rank == 0:
            ok = True
            for c, s in enumerate(shapes):
                prev = torch.zeros(s[1:])
                for i in range(s[0]):
                    prev = 0.5 * prev + data[c][i]
                    ok &= bool(torch.equal(res[c][i], prev))
            q.put(ok)
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


def _stream_worker(rank, world, port, q):
    import sys

    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from fav_b200 import clips

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_clips, n_frames, chunk, H, W = 3, 7, 3, 4, 6
        g = torch.Generator().manual_seed(5)
        frames = torch.rand((n_clips, n_frames, 3, H, W), generator=g)
        flows = torch.rand((n_clips, n_frames, 2, H, W), generator=g)
        state, got = {}, {}

        def load_chunk(c, f0, f1):
            return [frames[c, f0:f1].clone(), flows[c, f0:f1].clone()]

        def process_chunk(c, f0, inputs):  # recurrent stand-in for the GPU frame loop: needs frame order and its own state
            fr, fl = inputs
            outs = []
            for i in range(fr.shape[0]):
                prev = state.get(c, torch.zeros((3, H, W)))
                prev = 0.5 * prev + fr[i] + fl[i].sum(0, keepdim=True)
                state[c] = prev
                outs.append(prev)
            return torch.stack(outs)

        def store_chunk(c, f0, out):
            got[(c, f0)] = out.clone()

        st = clips.stream_clips(n_clips, n_frames, chunk, [(3, H, W), (2, H, W)], (3, H, W), load_chunk, process_chunk, store_chunk)
        if rank == 0:
            ok = True
            for c in range(n_clips):
                prev = torch.zeros((3, H, W))
                for i in range(n_frames):
                    prev = 0.5 * prev + frames[c, i] + flows[c, i].sum(0, keepdim=True)
                    f0 = (i // chunk) * chunk
                    ok &= bool(torch.equal(got[(c, f0)][i - f0], prev))
            ok &= st["bytes_in"] > 0 and len(got) == n_clips * 3
            q.put(ok)
        else:
            assert not got and st["bytes_out"] > 0
        # byte payloads (bench.py --config cfg3 --payload bytes): uint8 frames travel, float flows, uint8 results come back
        frames8 = (frames * 255).to(torch.uint8)
        state.clear(); got.clear()

        def load8(c, f0, f1):
            return [frames8[c, f0:f1].clone(), flows[c, f0:f1].clone()]

        def process8(c, f0, inputs):
            fr, fl = inputs
            assert fr.dtype == torch.uint8 and fl.dtype == torch.float32
            outs = []
            for i in range(fr.shape[0]):
                prev = state.get(c, torch.zeros((3, H, W)))
                prev = 0.5 * prev + fr[i].float() / 255.0 + fl[i].sum(0, keepdim=True)
                state[c] = prev
                outs.append((prev.clamp(0, 1) * 255.0 + 0.5).floor().to(torch.uint8))
            return torch.stack(outs)

        st = clips.stream_clips(n_clips, n_frames, chunk, [(3, H, W), (2, H, W)], (3, H, W), load8, process8, store_chunk,
                                in_dtypes=[torch.uint8, torch.float32], out_dtype=torch.uint8)
        if rank == 0:
            ok = True
            for c in range(n_clips):
                prev = torch.zeros((3, H, W))
                for i in range(n_frames):
                    prev = 0.5 * prev + frames8[c, i].float() / 255.0 + flows[c, i].sum(0, keepdim=True)
                    f0 = (i // chunk) * chunk
                    want = (prev.clamp(0, 1) * 255.0 + 0.5).floor().to(torch.uint8)
                    ok &= got[(c, f0)].dtype == torch.uint8 and bool(torch.equal(got[(c, f0)][i - f0], want))
            q.put(ok)
    finally:
        dist.destroy_process_group()


def test_streaming_data_plane_world2_gloo():
    """BASELINE.json config 3's data plane (rank 0 feeds chunks of frames to the owning ranks, outputs come back) on CPU."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stream_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True   # fp32 payloads
    assert q.get(timeout=5) is True   # byte payloads


def test_assign_clips_partitions():
    from fav_b200 import clips

    for n in (0, 1, 7, 8, 9):
        for w in (1, 2, 8):
            parts = clips.assign_clips(n, w)
            assert sorted(c for p in parts for c in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_scatter_process_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
