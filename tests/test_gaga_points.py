"""the FLAME side of GAGAvatar.build_forward_batch (app/GAGAvatar/models.py:98-128): scale 5.0,
avatar shape code, jaw-only pose, forehead EMA across frames. The oracle restates the reference's frame-by-frame loop; the
CUDA builder decodes a clip in one call (EMA as a device scan) and must agree with it, batched or frame by frame."""
import pytest
import torch

from artalk_b200 import synthetic
from oracle.artalk_oracle import gaga_t_points, flame_vertices

IDX = [3, 17, 256, 1024, 4999, 5022, 2048]            # stand-in for the reference's forehead vertex list (models.py:326)


def _inputs(n=12):
    g = torch.Generator().manual_seed(11)
    motion = 0.3 * torch.randn(n, 106, generator=g)
    shape = 0.5 * torch.randn(1, 300, generator=g)
    return motion, shape


def test_oracle_ema_recurrence_and_untouched_vertices():
    asset = synthetic.make_flame_asset(0)
    motion, shape = _inputs(6)
    pts = gaga_t_points(asset, shape, motion, IDX)
    raw = torch.cat([flame_vertices(asset, shape, motion[f:f + 1, :100],
                                    torch.cat([torch.zeros(1, 3), motion[f:f + 1, 103:]], -1), scale=5.0) for f in range(6)])
    rest = [v for v in range(5023) if v not in IDX]
    assert torch.equal(pts[:, rest], raw[:, rest])                      # only the forehead vertices are filtered
    assert torch.equal(pts[0, IDX], raw[0, IDX])                        # first frame initialises the state unblended
    u = raw[0, IDX]
    for f in range(1, 6):
        u = 0.98 * u + 0.02 * raw[f, IDX]
        assert torch.allclose(pts[f, IDX], u, atol=1e-6)
    assert (raw[:, IDX] - pts[:, IDX]).abs().max() > 1e-3               # the filter does something on these inputs


@pytest.mark.gpu
@pytest.mark.parametrize("fprec", ["fp32", "tc"])
def test_cuda_builder_matches_oracle_batched_and_streamed(fprec):
    from artalk_b200.flame import FLAMEModel
    from artalk_b200.gaga import GagaPointBuilder
    asset = synthetic.make_flame_asset(0)
    motion, shape = _inputs(12)
    ref = gaga_t_points(asset, shape, motion, IDX)
    fm = FLAMEModel(n_shape=300, n_exp=100, scale=5.0, no_lmks=True, asset=asset, device="cuda:0", precision=fprec)
    b = GagaPointBuilder(fm, shape, IDX)
    whole = b.t_points(motion).cpu()
    assert whole.shape == (12, 5023, 3)
    assert (whole - ref).abs().max().item() < 1e-3
    b.reset()
    parts = torch.cat([b.t_points(motion[:1]), b.t_points(motion[1:5]), b.t_points(motion[5:])]).cpu()   # state carried across calls
    assert (parts - whole).abs().max().item() < 1e-5      # N = 1 decodes with a per-frame shape row (other summation order)
    with pytest.raises(ValueError):
        GagaPointBuilder(fm, shape, [6000])
