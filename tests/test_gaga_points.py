"""the FLAME side of GAGAvatar.build_forward_batch (app/GAGAvatar/models.py:98-128): scale 5.0,
avatar shape code, jaw-only pose, forehead EMA across frames over the reference's 68 forehead vertices, plus the vertex
normals a mesh rasteriser needs. The oracle restates the reference's frame-by-frame loop and is pinned against the live
``build_forward_batch`` (tests/golden/gaga.npz, oracle/make_golden.py::run_gaga); the CUDA builder decodes a clip in one call
(EMA as a device scan) and must agree with it, batched or frame by frame."""
import numpy as np
import pytest
import torch

from artalk_b200 import synthetic
from artalk_b200.gaga import FOREHEAD_INDICES
from oracle.artalk_oracle import gaga_t_points, flame_vertices, vertex_normals
from oracle.cases import gaga_inputs
import golden_util as gu

IDX = list(FOREHEAD_INDICES)


def test_forehead_table_is_the_reference_list():
    g = gu.load("gaga")
    assert len(IDX) == 68 and len(set(IDX)) == 68 and max(IDX) < 5023
    assert IDX == g["forehead_indices"].tolist()                        # app/GAGAvatar/models.py:326-331, read from the live module


def test_oracle_matches_live_build_forward_batch():
    """Pin of the restatement: t_points of 12 consecutive frames from the unmodified reference (forehead vertices in full,
    every 8th vertex otherwise)."""
    g = gu.load("gaga")
    asset = synthetic.make_flame_asset(0)
    motion, shape = gaga_inputs()
    pts = gaga_t_points(asset, shape, motion, IDX)
    np.testing.assert_allclose(pts[:, IDX].numpy(), g["forehead"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(pts[:, ::8].numpy(), g["strided"], atol=1e-5, rtol=0)


def test_oracle_ema_recurrence_and_untouched_vertices():
    asset = synthetic.make_flame_asset(0)
    motion, shape = gaga_inputs()
    motion = motion[:6]
    pts = gaga_t_points(asset, shape, motion, IDX)
    raw = torch.cat([flame_vertices(asset, shape, motion[f:f + 1, :100],
                                    torch.cat([torch.zeros(1, 3), motion[f:f + 1, 103:]], -1), scale=5.0) for f in range(6)])
    rest = [v for v in range(5023) if v not in set(IDX)]
    assert torch.equal(pts[:, rest], raw[:, rest])                      # only the forehead vertices are filtered
    assert torch.equal(pts[0, IDX], raw[0, IDX])                        # first frame initialises the state unblended
    u = raw[0, IDX]
    for f in range(1, 6):
        u = 0.98 * u + 0.02 * raw[f, IDX]
        assert torch.allclose(pts[f, IDX], u, atol=1e-6)
    assert (raw[:, IDX] - pts[:, IDX]).abs().max() > 1e-3               # the filter does something on these inputs


def test_vertex_adjacency_reproduces_the_scatter_formula():
    """The CSR gather the CUDA kernel walks == pytorch3d's corner-wise index_add (restated in the oracle)."""
    from artalk_b200.flame import vertex_adjacency
    asset = synthetic.make_flame_asset(0)
    faces = asset["flame_model"]["f"]
    off, pairs = vertex_adjacency(faces, 5023)
    assert int(off[-1]) == 3 * faces.shape[0] and off.dtype == torch.int32 and pairs.shape == (3 * faces.shape[0], 2)
    verts = torch.randn(2, 5023, 3, generator=torch.Generator().manual_seed(0))
    vsel = torch.repeat_interleave(torch.arange(5023), (off[1:] - off[:-1]).long())
    a = verts[:, pairs[:, 0].long()] - verts[:, vsel]
    b = verts[:, pairs[:, 1].long()] - verts[:, vsel]
    n = torch.zeros_like(verts).index_add_(1, vsel, torch.cross(a, b, dim=-1))
    n = torch.nn.functional.normalize(n, eps=1e-6, dim=-1)
    assert (n - vertex_normals(verts, faces)).abs().max().item() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("fprec", ["fp32", "tc"])
def test_cuda_builder_matches_oracle_batched_and_streamed(fprec):
    from artalk_b200.flame import FLAMEModel
    from artalk_b200.gaga import GagaPointBuilder
    asset = synthetic.make_flame_asset(0)
    motion, shape = gaga_inputs()
    ref = gaga_t_points(asset, shape, motion, IDX)
    g = gu.load("gaga")
    fm = FLAMEModel(n_shape=300, n_exp=100, scale=5.0, no_lmks=True, asset=asset, device="cuda:0", precision=fprec)
    b = GagaPointBuilder(fm, shape)                                      # default = the reference's forehead list
    whole = b.t_points(motion).cpu()
    assert whole.shape == (12, 5023, 3)
    assert (whole - ref).abs().max().item() < 1e-3
    np.testing.assert_allclose(whole[:, IDX].numpy(), g["forehead"], atol=1e-3, rtol=0)       # live reference
    b.reset()
    parts = torch.cat([b.t_points(motion[:1]), b.t_points(motion[1:5]), b.t_points(motion[5:])]).cpu()   # state carried across calls
    assert (parts - whole).abs().max().item() < 1e-5      # N = 1 decodes with a per-frame shape row (other summation order)
    with pytest.raises(ValueError):
        GagaPointBuilder(fm, shape, [6000])


@pytest.mark.gpu
def test_vertex_normals_match_oracle():
    from artalk_b200.flame import FLAMEModel
    asset = synthetic.make_flame_asset(0)
    fm = FLAMEModel(n_shape=300, n_exp=100, scale=1.0, no_lmks=True, asset=asset, device="cuda:0")
    g = torch.Generator().manual_seed(3)
    motion = 0.3 * torch.randn(37, 106, generator=g)
    verts = fm(shape_params=torch.zeros(1, 300).expand(37, -1).to("cuda:0"), expression_params=motion[:, :100].to("cuda:0"),
               pose_params=motion[:, 100:].to("cuda:0"))
    n = fm.vertex_normals(verts).cpu()
    ref = vertex_normals(verts.cpu(), fm.get_faces().cpu())
    assert n.shape == (37, 5023, 3)
    assert (n - ref).abs().max().item() < 2e-5
    # a strided view of a larger buffer (frame stride != V*3) and a big batch spanning several waves of CTAs
    big = torch.randn(700, 5023, 3, generator=g).to("cuda:0")
    nb = fm.vertex_normals(big).cpu()
    rb = vertex_normals(big[::97].cpu(), fm.get_faces().cpu())
    assert (nb[::97] - rb).abs().max().item() < 2e-5
    with pytest.raises(ValueError):
        fm.vertex_normals(big[:, :100])
