"""FLAME side of ``GAGAvatar.build_forward_batch`` (app/GAGAvatar/models.py:98-128), the second consumer of the motion codes:
FLAME decode with ``scale=5.0`` and the tracked avatar's shape code, jaw-only pose (``[0,0,0, motion[103:106]]``, zero eye
pose), then the forehead vertices follow an EMA across frames. The reference runs this one frame at a time inside its
render loop (inference.py:78-84); here a whole clip is decoded in one FLAME launch and the EMA is a device scan whose state
carries across calls, so frame-by-frame and batched calls agree. The Gaussian rasteriser, the camera transform
(``transform_emoca_to_p3d``, pytorch3d) and the image branches of the batch are rendering and stay out of scope.
"""
from __future__ import annotations

import torch

from . import _lib
from .flame import FLAMEModel


#: the reference's forehead vertex list (app/GAGAvatar/models.py:326-331): FLAME topology indices, part of the wire format
#: of the tracked-avatar pipeline (the EMA is applied to exactly these 68 vertices)
FOREHEAD_INDICES = (
    2168, 2165, 3068, 2199, 2196, 3720, 2091, 2088, 3524, 625, 628, 3871, 705, 708, 2030, 667, 670,
    3708, 3706, 3729, 3721, 3773, 3789, 3735, 3732, 3786, 3876, 3878, 3913, 3899, 3872, 3874, 3864, 3865,
    3158, 3157, 336, 335, 3153, 3705, 2177, 2176, 3540, 671, 672, 3863, 2134, 16, 17, 2138, 2139,
    2567, 2566, 337, 338, 3154, 3712, 2178, 2179, 3495, 674, 673, 3868, 2135, 27, 18, 1429, 1430,
)


class GagaPointBuilder:
    def __init__(self, flame_model: FLAMEModel, shapecode: torch.Tensor, forehead_indices=FOREHEAD_INDICES, keep: float = 0.98):
        """``flame_model`` built with ``scale=5.0`` (models.py:20), ``shapecode`` (1,300) of the tracked avatar,
        ``forehead_indices`` the reference's vertex list (models.py:326-331, the default)."""
        if shapecode.dim() != 2 or shapecode.shape[0] != 1:
            raise ValueError("shapecode must be (1, n_shape)")
        self.flame = flame_model
        self.device = flame_model.device
        self.shapecode = shapecode.to(self.device, torch.float32)
        self.idx = torch.as_tensor(list(forehead_indices), dtype=torch.int32, device=self.device)
        if self.idx.numel() and (int(self.idx.min()) < 0 or int(self.idx.max()) >= flame_model.n_verts):
            raise ValueError("forehead index out of range")
        self.keep = float(keep)
        self.state = torch.zeros(self.idx.numel(), 3, device=self.device)
        self.has_state = False

    def reset(self):
        self.has_state = False

    @torch.no_grad()
    def t_points(self, motion_code: torch.Tensor) -> torch.Tensor:
        """(N,106) motion codes of consecutive frames -> ``feature_batch['t_points']`` for each frame, (N,5023,3)."""
        m = motion_code.to(self.device, torch.float32)
        if m.dim() != 2 or m.shape[1] != 106:
            raise ValueError("motion_code must be (N, 106)")
        N = m.shape[0]
        if N == 0:
            return torch.zeros(0, self.flame.n_verts, 3, device=self.device)
        exp_code = m[:, :100]
        pose_code = torch.cat([m.new_zeros(N, 3), m[:, 103:]], dim=-1)                  # models.py:115
        pts = self.flame(shape_params=self.shapecode.expand(N, -1), pose_params=pose_code, expression_params=exp_code,
                         eye_pose_params=m.new_zeros(N, 6)).float()
        if self.idx.numel():
            _lib.call(self.device, _lib.lib().artalk_ema_scan, pts.data_ptr(), pts.stride(0), self.idx.data_ptr(), self.idx.numel(), N,
                                                  self.state.data_ptr(), int(self.has_state), self.keep, _lib.stream_ptr(self.device))
        self.has_state = True
        return pts
