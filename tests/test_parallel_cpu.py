"""CPU: the N>1 path (clip sharding + motion gather) with world_size-2 gloo process groups."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from artalk_b200 import parallel


def test_shard_bounds_partition():
    for n in (0, 1, 5, 64, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))         # contiguous, in clip order
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeEngine:
    """Stands in for ARTAvatarInferEngine on CPU: 'motion' encodes the clip index so ordering errors are visible."""

    def inference_batch(self, audio, style_motion=None, clip_length=None):
        idx = audio[:, 0]
        out = idx.view(-1, 1, 1).expand(-1, 7, 106).clone()
        if style_motion is not None:
            out = out + 0.25 * style_motion[:, :1, :1]
        return out


def _worker(rank, world, port, n_clips, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    parallel.init_process_group("gloo")
    try:
        make_audio = lambda lo, hi: torch.arange(lo, hi, dtype=torch.float32).view(-1, 1).repeat(1, 16)
        make_style = lambda lo, hi: torch.ones(hi - lo, 50, 106)
        out = parallel.sharded_inference(_FakeEngine(), make_audio, make_style, n_clips)
        lo, hi = parallel.shard_bounds(n_clips, world, rank)
        raw = parallel.gather_motion(torch.full((hi - lo, 2, 3), float(rank)), n_clips)
        mg = parallel.MotionGather(None)                         # CPU group: synchronous fallback of the side-stream gather
        h = mg.start(torch.full((hi - lo, 2, 3), float(rank)), n_clips)
        assert torch.equal(mg.wait(h), raw) and mg.mean_ms() == 0.0 and not mg._pending
        q.put((rank, out[:, 0, 0].tolist(), tuple(out.shape), raw[:, 0, 0].tolist()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [6, 5, 1])
def test_sharded_gather_world2_gloo(n_clips):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [i + 0.25 for i in range(n_clips)]
    for rank, vals, shape, raw in res:
        assert shape == (n_clips, 7, 106)
        assert vals == expect                                     # every rank holds all clips in clip order
        owner = [r for r in range(world) for _ in range(*parallel.shard_bounds(n_clips, world, r))]
        assert raw == [float(o) for o in owner]


def test_single_process_passthrough():
    x = torch.randn(3, 4, 106)
    assert parallel.gather_motion(x, 3) is x
    with pytest.raises(ValueError):
        parallel.gather_motion(x, 4)
