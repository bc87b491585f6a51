"""Drop-in for ``app.flame_model.FLAMEModel`` on the mesh-decode path (``no_lmks=True``): same constructor
arguments, ``forward`` signature and ``get_faces``; vertices come from the fused sm_100a blend+LBS kernels behind
``artalk_flame_vertices`` (include/artalk_b200.h). Reference: app/flame_model/FLAME.py:16-69,117-149 and
app/flame_model/lbs.py:142-383. Landmark outputs (FLAME.py:150-204) are outside the path and not provided."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib


class FLAMEModel:
    N_JOINTS = 5
    MAX_FRAMES = 65536

    def __init__(self, n_shape, n_exp, scale=1.0, no_lmks=False, lmks_type="lmks70", *, asset=None, asset_path=None,
                 device="cuda", precision="tc"):
        """precision: "tc" = blend on the tensor cores with split-bf16 operands (hi+lo, ~fp32 accuracy);
        "fp32" = CUDA-core fp32 kernel."""
        if not no_lmks:
            raise NotImplementedError("landmark outputs (FLAME.py:150-204) are outside the audio->motion->mesh path; "
                                      "construct with no_lmks=True as inference.py:29 does")
        self.scale, self.no_lmks, self.lmks_type = float(scale), no_lmks, lmks_type
        self.n_shape, self.n_exp = int(n_shape), int(n_exp)
        self.device = _lib.require_cuda(device)
        if asset is None:
            path = asset_path or os.path.join(os.getcwd(), "assets", "FLAME_with_eye.pt")
            asset = torch.load(path, map_location="cpu", weights_only=True)       # raises if missing, like the reference
        fm = asset["flame_model"]
        f32 = torch.float32
        self.faces_tensor = fm["f"].to(self.device)
        v_template = fm["v_template"].to(f32)
        sdirs = fm["shapedirs"].to(f32)
        sdirs = torch.cat([sdirs[:, :, :n_shape], sdirs[:, :, 300:300 + n_exp]], 2)          # FLAME.py:36
        V = v_template.shape[0]
        posedirs = fm["posedirs"].to(f32).reshape(-1, fm["posedirs"].shape[-1]).T           # (36, V*3)  FLAME.py:37-38
        if posedirs.shape[0] != 36:
            raise ValueError("expected 36 pose-corrective bases, got %d" % posedirs.shape[0])
        jreg = fm["J_regressor"].to(f32)
        parents = fm["kintree_table"][0].clone().to(torch.int64)
        parents[0] = -1
        if jreg.shape[0] != self.N_JOINTS or parents.numel() != self.N_JOINTS:
            raise ValueError("FLAME asset must have 5 joints")
        nb = sdirs.shape[2]
        dirs = torch.cat([sdirs.permute(2, 0, 1).reshape(nb, V * 3), posedirs], 0)         # basis-major
        self._bufs = {
            "v_template": v_template.reshape(-1).contiguous().to(self.device),
            "dirs": dirs.contiguous().to(self.device),
            "j_template": (jreg @ v_template).reshape(-1).contiguous().to(self.device),
            "j_dirs": torch.einsum("jv,vkl->jkl", jreg, sdirs).reshape(15, nb).contiguous().to(self.device),   # [15][nb]
            "lbs_weights": fm["weights"].to(f32).contiguous().to(self.device),
        }
        self.parents = parents
        if precision not in ("tc", "fp32"):
            raise ValueError("precision must be 'tc' or 'fp32'")
        self.precision = precision
        if precision == "tc":
            def split(d):                       # d: (n_l, V*3) fp32 -> [V*3][3*KS] bf16 = [hi | hi | lo]
                n_l = d.shape[0]
                ks = (n_l + 63) // 64 * 64
                t = torch.zeros(V * 3, ks, dtype=torch.float32)
                t[:, :n_l] = d.t()
                hi = t.to(torch.bfloat16)
                lo = (t - hi.float()).to(torch.bfloat16)
                return torch.cat([hi, hi, lo], dim=1).contiguous().to(self.device), ks
            self._bufs["bsplit_full"], ks_full = split(dirs)
            self._bufs["bsplit_expr"], ks_expr = split(dirs[n_shape:])
        m = _lib.FlameModelC()
        m.n_verts, m.n_shape, m.n_exp = V, n_shape, nb - n_shape
        for k, t in self._bufs.items():
            setattr(m, k, t.data_ptr())
        if precision == "tc":
            m.ks_full, m.ks_expr = ks_full, ks_expr
        for i in range(5):
            m.parents[i] = int(parents[i])
        m.scale = self.scale
        self._c = m
        self.n_verts = V
        self._ws = None

    # nn.Module look-alikes used by callers of the reference class
    def to(self, device):
        if torch.device(device) != self.device and torch.device(device).index not in (None, self.device.index):
            raise _lib.ArtalkError("FLAMEModel buffers live on %s; rebuild it for %s" % (self.device, device))
        return self

    def eval(self):
        return self

    def get_faces(self):
        return self.faces_tensor.long()

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    def forward(self, shape_params=None, expression_params=None, pose_params=None, eye_pose_params=None, verts_sclae=None, out=None):
        """shape (N,n_shape), expression (N,n_exp), pose (N,6) [global rot, jaw] or (N,3) [jaw] -> (N,V,3)*scale."""
        lib = _lib.lib()
        N = shape_params.shape[0]
        dev = self.device
        if eye_pose_params is not None and bool((eye_pose_params != 0).any()):
            raise NotImplementedError("non-zero eye pose is not on the path (inference.py never passes it)")
        if expression_params is None:
            expression_params = torch.zeros(N, self.n_exp, device=dev)
        if pose_params is None:
            pose_params = torch.zeros(N, 6, device=dev)
        if pose_params.shape[-1] == 3:
            pose_params = torch.cat([torch.zeros(N, 3, device=pose_params.device), pose_params], dim=-1)
        shape_params = shape_params.to(dev, torch.float32)
        shared_shape = N > 1 and shape_params.stride(0) == 0          # expand()ed single row (inference.py:64)
        if not shared_shape:
            shape_params = shape_params.contiguous()
        elif shape_params.stride(1) != 1:
            shape_params = shape_params[:1].contiguous().expand(N, -1)
        expr = expression_params.to(dev, torch.float32).contiguous()
        pose = pose_params.to(dev, torch.float32).contiguous()
        if shape_params.shape[1] != self._c.n_shape or expr.shape[1] != self._c.n_exp or pose.shape[1] != 6:
            raise ValueError("bad FLAME parameter widths: shape %s expr %s pose %s" %
                             (tuple(shape_params.shape), tuple(expr.shape), tuple(pose.shape)))
        verts = torch.empty(N, self.n_verts, 3, dtype=torch.float32, device=dev) if out is None else out
        if tuple(verts.shape) != (N, self.n_verts, 3) or verts.dtype != torch.float32 or not verts.is_contiguous():
            raise ValueError("out must be a contiguous (N, %d, 3) fp32 tensor" % self.n_verts)
        # long clips / big batches go through in launches of at most MAX_FRAMES frames (bounded workspace: 1.7 KB per frame)
        step = self.MAX_FRAMES
        need = lib.artalk_flame_workspace_floats(C.byref(self._c), min(N, step))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.float32, device=dev)
        for i0 in range(0, N, step):
            n = min(step, N - i0)
            sp = shape_params if shared_shape else shape_params[i0:i0 + n]
            _lib.call(dev, lib.artalk_flame_vertices,
                      C.byref(self._c), sp.data_ptr(), 0 if shared_shape else shape_params.stride(0),
                      expr[i0:i0 + n].data_ptr(), expr.stride(0), pose[i0:i0 + n].data_ptr(), pose.stride(0), 0, self._ws.data_ptr(),
                      verts[i0:i0 + n].data_ptr(), n, _lib.stream_ptr(dev))
        return verts
