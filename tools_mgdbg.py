"""Developer tool: per-phase times inside the opt-in whole-stack kernel (skinny.cu::ar_small_kernel).
  ARTALK_MG_DEBUG=1 python tools_mgdbg.py        -> the library prints CTA 0's work/wait ns per phase after every launch
(graphs off; 64 clips x one 4 s chunk, FULL config, bf16). Output kept in profiles/r1i_whole_stack_phases.md."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from artalk_b200 import config, synthetic, _lib
from artalk_b200.engine import ARTAvatarInferEngine

cfg = config.FULL
eng = ARTAvatarInferEngine(load_gaga=False, device="cuda:0", precision="bf16", state_dict=synthetic.make_state_dict(cfg, 0),
                           config=cfg.to_reference_json(), flame_asset=synthetic.make_flame_asset(0), wav2vec=cfg.wav2vec,
                           make_output_dir=False)
_lib.check(_lib.lib().artalk_set_option(b"ar_small", 1))
eng.ARTalk.enable_graphs(False)
B = 64
audio = synthetic.make_audio(B, cfg.chunk_samples).to("cuda:0")
style = synthetic.make_style_motion(B).to("cuda:0")
for _ in range(2):
    eng.inference_batch(audio, style)
torch.cuda.synchronize()
