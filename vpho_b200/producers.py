"""Host-side mirror of the modules that produce the hot path's inputs over `vpho_heads_*`.

Replaces, in `vpho_net.forward` (lib/model/VPHO.py:129-178), the calls

    pd_hm_hand = self.head_hm_hand(hf_hr); pd_hm_obj = self.head_hm_obj(of_or_rect)            # head_inplane.py:99-104
    ... align_hm_to_bbox_rectangle / flip_tensor_by_mask_index / F.interpolate ...              # VPHO.py:132-148
    encoding_hand, enc_hand_ls = self.encoder_hand(cat(hf_hr_rect, pd_hm_hand_rs))              # encoding.py:58-73
    encoding_obj, enc_obj_ls = self.encoder_obj(cat(of_or_rect, pd_hm_obj_ori_rs))
    pd_mano_pose, pd_mano_shape = self.head_mano(encoding_hand)                                 # head_mano.py:61-76
    enc_phy_hand, _, _ = self.cross_hand(...); _, enc_phy_obj, _ = self.cross_obj(...)          # cross_module.py:119-137
    pd_phy_dt = self.head_physics(enc_phy_hand, enc_phy_obj)                                    # physics.py:700-721

by one C call.  Weights are taken as the reference's state dict (the `rest` part that `vpho_b200.checkpoint` returns, or
`vpho_net.state_dict()` itself); all dimensions come from the tensors' shapes.  There is no CPU path: the library is the
sm_100a binary and `capi.lib()` raises without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import capi

PREFIXES = ("head_hm_hand.", "head_hm_obj.", "encoder_hand.", "encoder_obj.", "head_mano.", "cross_hand.", "cross_obj.",
            "head_physics.")


class FeatureHeads:
    """`FeatureHeads(state)(hf_hr=, of_or_rect=, hf_hr_rect=, data=)` -> dict with the reference's names:
    `hand_heatmap`, `obj_heatmap` (pd_hm_hand / pd_hm_obj), `encoding_hand`, `encoding_obj`, `mano_pose`, `mano_shape`
    (pd_mano_pose / pd_mano_shape), `force_local`, `scale`, `weight`, `CoM` (pd_phy_dt)."""

    def __init__(self, state: Dict[str, object], lib: Optional[capi.Library] = None):
        self.lib = lib or capi.lib()
        keep = {}
        for k, v in state.items():
            for w in ("module.", "_orig_mod."):
                while k.startswith(w):
                    k = k[len(w):]
            if not k.startswith(PREFIXES) or k.endswith("num_batches_tracked") or k.startswith("head_mano.mano_layer."):
                continue
            a = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            keep[k] = np.ascontiguousarray(a, dtype=np.float32)
        if not keep:
            raise capi.VphoError("FeatureHeads: the state dict holds none of the head_hm_* / encoder_* / head_mano / cross_* / "
                                 "head_physics tensors")
        names = sorted(keep)
        table = (capi.NamedTensor * len(names))()
        self._keepalive = [k.encode() for k in names]
        for i, k in enumerate(names):
            a = keep[k]
            if a.ndim > 4:
                raise capi.VphoError(f"{k}: {a.ndim}-d tensor")
            table[i].name = self._keepalive[i]
            table[i].data = a.ctypes.data
            table[i].ndim = a.ndim
            for d in range(a.ndim):
                table[i].shape[d] = a.shape[d]
        h = C.c_void_p()
        self.lib.check(self.lib.c.vpho_heads_create(table, len(names), C.byref(h)), "vpho_heads_create (missing key or "
                       "inconsistent shape in the state dict)")
        self.handle = h
        dims = (C.c_int32 * 8)()
        self.lib.check(self.lib.c.vpho_heads_dims(h, dims), "vpho_heads_dims")
        self.C, self.Jh, self.Jo, self.enc_dim, self.d_model, self.n_force, self.heat_hid, self.enc_hid = list(dims)
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.c.vpho_heads_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def forward(self, hf_hr: torch.Tensor, of_or_rect: torch.Tensor, hf_hr_rect: torch.Tensor, data: Dict[str, torch.Tensor],
                debug: bool = False, strict_fp32: bool = False, check_overflow: bool = False) -> Dict[str, torch.Tensor]:
        """`data`: bbox_hand, bbox_hand_rect, bbox_obj, bbox_obj_rect (bs, 4), is_right (bs,) bool, gravity (bs, 1, 3) or (bs, 3)
        -- the reference's batch keys (VPHO.py:113-160).
        strict_fp32: the FP32 SIMT kernels (VPHO_HEADS_STRICT_FP32) instead of the tcgen05 path -- cross-checks, and the only
        path of the emulator build.  check_overflow: synchronise and raise if an activation left the FP16 range of the
        tensor-core operand planes (|v| >= 60000; never the case for BatchNorm-ed features of a trained network)."""
        dev = hf_hr.device
        bs, Cc, roi, roi2 = hf_hr.shape
        if Cc != self.C or roi != roi2 or of_or_rect.shape != hf_hr.shape or hf_hr_rect.shape != hf_hr.shape:
            raise capi.VphoError(f"FeatureHeads: feature maps must be (bs, {self.C}, roi, roi), got {tuple(hf_hr.shape)}")
        f32 = dict(dtype=torch.float32, device=dev)

        def f(t, shape):
            t = torch.as_tensor(t, device=dev).to(torch.float32).reshape(shape).contiguous()
            return t
        feats = [t.to(torch.float32).contiguous() for t in (hf_hr, of_or_rect, hf_hr_rect)]
        boxes = [f(data[k], (bs, 4)) for k in ("bbox_hand", "bbox_hand_rect", "bbox_obj", "bbox_obj_rect")]
        is_right = torch.as_tensor(data["is_right"], device=dev).to(torch.bool).reshape(bs).contiguous()
        gravity = f(data["gravity"], (bs, 3))
        F, hm = self.n_force, 2 * roi
        out = {"hand_heatmap": torch.empty((bs, self.Jh, hm, hm), **f32), "obj_heatmap": torch.empty((bs, self.Jo, hm, hm), **f32),
               "encoding_hand": torch.empty((bs, self.enc_dim), **f32), "encoding_obj": torch.empty((bs, self.enc_dim), **f32),
               "mano_pose": torch.empty((bs, 48), **f32), "mano_shape": torch.empty((bs, 10), **f32),
               "force_local": torch.empty((bs, F, 3), **f32), "scale": torch.empty((bs, F), **f32),
               "weight": torch.empty((bs, F, 8), **f32), "CoM": torch.empty((bs, F, 3), **f32)}
        if debug:
            out["enc_phy_hand"] = torch.empty((bs, F, self.d_model), **f32)
            out["enc_phy_obj"] = torch.empty((bs, F, self.d_model), **f32)
        need = self.lib.c.vpho_heads_workspace_bytes(self.handle, bs, roi)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        P = capi.ptr
        a = capi.HeadsArgs(bs=bs, roi_size=roi, hf_hr=P(feats[0]), of_or_rect=P(feats[1]), hf_hr_rect=P(feats[2]),
                           bbox_hand=P(boxes[0]), bbox_hand_rect=P(boxes[1]), bbox_obj=P(boxes[2]), bbox_obj_rect=P(boxes[3]),
                           is_right=P(is_right), gravity=P(gravity), hand_heatmap=P(out["hand_heatmap"]),
                           obj_heatmap=P(out["obj_heatmap"]), encoding_hand=P(out["encoding_hand"]),
                           encoding_obj=P(out["encoding_obj"]), mano_pose=P(out["mano_pose"]), mano_shape=P(out["mano_shape"]),
                           force_local=P(out["force_local"]), force_scale=P(out["scale"]), force_weight=P(out["weight"]),
                           CoM=P(out["CoM"]), enc_phy_hand=P(out.get("enc_phy_hand")), enc_phy_obj=P(out.get("enc_phy_obj")),
                           flags=1 if strict_fp32 else 0)
        self.lib.check(self.lib.c.vpho_heads_forward(self.handle, C.byref(a), P(self._ws), self._ws.numel(), capi.stream_of(hf_hr)),
                       "vpho_heads_forward")
        if check_overflow and not strict_fp32:
            flag = C.c_int32(0)
            self.lib.check(self.lib.c.vpho_heads_overflow(self.handle, C.byref(flag), capi.stream_of(hf_hr)), "vpho_heads_overflow")
            if flag.value:
                raise capi.VphoError("FeatureHeads: activation outside the FP16 range of the tensor-core planes; use strict_fp32=True")
        return out

    __call__ = forward
