"""CPU tests of the N > 1 path: frame-parallel sharding + the single gather, world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lidar_rt_b200.sweep import frames_of_rank, gather_frames, render_sweep


def test_frames_of_rank_is_a_partition():
    for n in (0, 1, 7, 200):
        for world in (1, 2, 3, 8):
            parts = [frames_of_rank(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        frames_of_rank(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rendered = []

        def render_frame(f):                       # stands in for Tracer.forward: an (H, W, 9) buffer per frame
            rendered.append(f)
            return torch.full((4, 6, 9), float(f)) + torch.arange(9, dtype=torch.float32) * 0.01

        out = render_sweep(render_frame, n_frames, rank, world, gather=True)
        ok = out.shape == (n_frames, 4, 6, 9)
        for f in range(n_frames):
            ok = ok and torch.allclose(out[f], torch.full((4, 6, 9), float(f)) + torch.arange(9, dtype=torch.float32) * 0.01)
        ok = ok and rendered == frames_of_rank(n_frames, rank, world)
        local = render_sweep(render_frame, n_frames, rank, world, gather=False)
        ok = ok and local.shape[0] == len(frames_of_rank(n_frames, rank, world))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 8])
def test_frame_parallel_sweep_world2_gloo(n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


def test_gather_frames_single_rank_is_identity():
    x = torch.randn(3, 2, 2, 9)
    assert gather_frames(x, 3, 0, 1) is x
