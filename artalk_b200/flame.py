"""Drop-in for ``app.flame_model.FLAMEModel`` on the mesh-decode path (``no_lmks=True``): same constructor
arguments, ``forward`` signature and ``get_faces``; vertices come from the fused sm_100a blend+LBS kernels behind
``artalk_flame_vertices`` (include/artalk_b200.h). Reference: app/flame_model/FLAME.py:16-69,117-149 and
app/flame_model/lbs.py:142-383. Landmark outputs (FLAME.py:150-204) are outside the path and not provided."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib


def vertex_adjacency(faces: torch.Tensor, n_verts: int):
    """CSR vertex -> incident faces for ``artalk_vertex_normals``: offsets (V+1,) int32 and pairs (E,2) int32 holding, for
    every incidence of vertex v in a face (i0,i1,i2), the (next, prev) vertices in the face's cyclic order, so that
    cross(P_next - P_v, P_prev - P_v) is the expression pytorch3d accumulates at that corner. Incidences of a vertex are kept
    in face order."""
    f = faces.detach().to("cpu", torch.int64)
    v = torch.cat([f[:, 0], f[:, 1], f[:, 2]])
    nxt = torch.cat([f[:, 1], f[:, 2], f[:, 0]])
    prv = torch.cat([f[:, 2], f[:, 0], f[:, 1]])
    order = torch.sort(v * (3 * f.shape[0]) + torch.arange(v.numel()) % f.shape[0] * 3 + torch.arange(v.numel()) // f.shape[0]).indices
    counts = torch.bincount(v, minlength=n_verts)
    offsets = torch.zeros(n_verts + 1, dtype=torch.int32)
    offsets[1:] = torch.cumsum(counts, 0).to(torch.int32)
    pairs = torch.stack([nxt[order], prv[order]], dim=1).to(torch.int32).contiguous()
    return offsets, pairs


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
        self._adj = None

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

    def vertex_normals(self, verts: torch.Tensor, out=None) -> torch.Tensor:
        """(N,V,3) decoded vertices -> (N,V,3) unit vertex normals over ``get_faces()`` (area-weighted, pytorch3d
        ``Meshes.verts_normals`` semantics): what a mesh rasteriser fed with the path's output needs."""
        if verts.dim() != 3 or verts.shape[1] != self.n_verts or verts.shape[2] != 3:
            raise ValueError("verts must be (N, %d, 3)" % self.n_verts)
        v = verts.to(self.device, torch.float32).contiguous()
        if self._adj is None:
            off, pairs = vertex_adjacency(self.faces_tensor, self.n_verts)
            self._adj = (off.to(self.device), pairs.to(self.device))
        normals = torch.empty_like(v) if out is None else out
        _lib.call(self.device, _lib.lib().artalk_vertex_normals, v.data_ptr(), v.stride(0), self.n_verts, self._adj[0].data_ptr(),
                  self._adj[1].data_ptr(), normals.data_ptr(), v.shape[0], _lib.stream_ptr(self.device))
        return normals

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
