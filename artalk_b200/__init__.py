"""B200-native (sm_100a) implementation of ARTalk's audio->motion hot path behind the reference's call surface."""
from .config import ModelConfig, Wav2VecConfig, FULL, TINY  # noqa: F401

__all__ = ["ModelConfig", "Wav2VecConfig", "FULL", "TINY", "ARTAvatarInferEngine", "BitwiseARModel", "FLAMEModel"]


def __getattr__(name):          # lazy: importing the config / synthetic helpers must not require CUDA
    if name == "ARTAvatarInferEngine":
        from .engine import ARTAvatarInferEngine
        return ARTAvatarInferEngine
    if name == "BitwiseARModel":
        from .model import BitwiseARModel
        return BitwiseARModel
    if name == "FLAMEModel":
        from .flame import FLAMEModel
        return FLAMEModel
    raise AttributeError(name)
