"""GPU: size-independent properties at BASELINE.json's full sizes (the oracle is too slow there): determinism,
batched == per-clip, chunk-count arithmetic, output post-ops, graph replay == eager launches, FLAME at mesh-path scale."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from artalk_b200 import config, synthetic  # noqa: E402
from artalk_b200.engine import ARTAvatarInferEngine  # noqa: E402
import golden_util as gu  # noqa: E402

DEV = "cuda:0"


@pytest.fixture(scope="module")
def engine():
    eng = ARTAvatarInferEngine(load_gaga=False, clip_length=750, device=DEV, precision="bf16", state_dict=gu.state_dict("FULL"),
                               config=config.FULL.to_reference_json(), flame_asset=synthetic.make_flame_asset(0),
                               wav2vec=config.FULL.wav2vec, make_output_dir=False)
    yield eng
    eng.ARTalk.close()


def test_config1_batch64_10s_properties(engine):
    """BASELINE configs[1]: 64 clips x 10 s."""
    a, s = synthetic.make_audio(64, 160000), synthetic.make_style_motion(64)
    out1 = engine.inference_batch(a, s)
    out2 = engine.inference_batch(a, s)                       # second call replays the CUDA graphs
    assert out1.shape == (64, 250, 106) and torch.isfinite(out1).all()
    assert torch.equal(out1, out2)                            # deterministic; graph replay == eager warm-up
    assert float(out1[..., 104:].abs().max()) == 0.0          # inference.py:56
    for b in (0, 37, 63):                                     # batched == per-clip loop (reference semantics, B = 1)
        engine.set_style_motion(s[b])
        one = engine.inference(a[b])
        assert torch.equal(one, out1[b])
    engine.style_motion = None
    # eager launches give the same bits as the graph
    engine.ARTalk.enable_graphs(False)
    try:
        out3 = engine.inference_batch(a, s)
    finally:
        engine.ARTalk.enable_graphs(True)
    assert torch.equal(out1, out3)


def test_config2_30s_clip_length_and_chunks(engine):
    """BASELINE configs[2] shape per clip: 30 s = 750 frames = 8 chunks (last one ragged); clip_length truncates only."""
    a = synthetic.make_audio(8, 480000)
    full = engine.inference_batch(a, None, clip_length=10000)
    assert full.shape == (8, 750, 106)
    cut = engine.inference_batch(a, None, clip_length=100)
    assert cut.shape == (8, 100, 106)
    # smoothing of frame t only sees frames t-4..t+4: the first 96 frames of the truncated output are unaffected
    assert torch.equal(cut[:, :96], full[:, :96])
    # a longer clip's first chunk equals the same audio run alone (state flows forward only)
    first = engine.ARTalk.inference({"audio": a[:2, :64000], "style_motion": None})
    both = engine.ARTalk.inference({"audio": a[:2], "style_motion": None})
    assert torch.equal(first, both[:, :100])


def test_fix_pose_and_null_style(engine):
    a = synthetic.make_audio(2, 64000)
    eng2_fix, base_fix = engine.fix_pose, engine.inference_batch(a, None)
    engine.fix_pose = True
    try:
        fixed = engine.inference_batch(a, None)
    finally:
        engine.fix_pose = eng2_fix
    assert float(fixed[..., 100:103].abs().max()) == 0.0       # inference.py:53-54
    keep = [i for i in range(106) if not (100 <= i < 103)]
    assert torch.equal(fixed[..., keep], base_fix[..., keep])


def test_flame_mesh_path_scale(engine):
    """750-frame clip through the mesh branch (inference.py:62-69): tensor-core kernel == fp32 kernel within 1e-5."""
    from artalk_b200.flame import FLAMEModel
    motion = 0.3 * torch.randn(750, 106, generator=torch.Generator().manual_seed(5))
    v_tc = engine.mesh_vertices(motion.to(DEV))
    fm32 = FLAMEModel(n_shape=300, n_exp=100, scale=1.0, no_lmks=True, asset=synthetic.make_flame_asset(0), device=DEV,
                      precision="fp32")
    v_32 = engine.ARTalk.basic_vae.get_flame_verts(fm32, torch.zeros(1, 300, device=DEV).expand(750, -1), motion.to(DEV),
                                                   with_global=True)
    assert v_tc.shape == (750, 5023, 3)
    assert float((v_tc - v_32).abs().max()) < 1e-5
