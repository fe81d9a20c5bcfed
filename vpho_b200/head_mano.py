"""Host-side mirror of `HeadMano.get_hand_verts` (lib/model/head_mano.py:78-87) over the CUDA MANO layer."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import capi


class HeadMano:
    """Drop-in for the evaluation-time part of the reference's `HeadMano`: `get_hand_verts(pose=, shape=)`.

    `model` holds the MANO tensors with manopth's layouts (v_template, shapedirs, posedirs, J_regressor, weights);
    with the licensed `MANO_RIGHT.pkl` at hand, pass its fields; `vpho_b200.synthetic.make_mano_model` builds a
    stand-in with identical shapes.
    """

    def __init__(self, model: Dict[str, np.ndarray], lib: Optional[capi.Library] = None):
        self.lib = lib or capi.lib()
        arrs = [np.ascontiguousarray(model[k], dtype=np.float32)
                for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "weights")]
        assert arrs[0].shape == (778, 3) and arrs[1].shape == (778, 3, 10) and arrs[2].shape == (778, 3, 135)
        assert arrs[3].shape == (16, 778) and arrs[4].shape == (778, 16)
        h = C.c_void_p()
        self.lib.check(self.lib.c.vpho_mano_create(*[capi.host_ptr(a) for a in arrs], C.byref(h)), "vpho_mano_create")
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.c.vpho_mano_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def get_hand_verts(self, **kwargs):
        """pose (n,48) axis-angle, shape (n,10) -> verts (n,778,3), joints (n,21,3), metres, wrist-centred."""
        pose = kwargs["pose"].contiguous().float()
        shape = kwargs["shape"].contiguous().float()
        n = pose.shape[0]
        assert pose.shape == (n, 48) and shape.shape == (n, 10)
        need_verts = kwargs.get("need_verts", True)
        verts = torch.empty((n, 778, 3), dtype=torch.float32, device=pose.device) if need_verts else None
        joints = torch.empty((n, 21, 3), dtype=torch.float32, device=pose.device)
        # strict_fp32=True: the FP32 SIMT kernel instead of the tcgen05 blend (VPHO_MANO_STRICT_FP32; cross-checks only)
        flags = 1 if kwargs.get("strict_fp32", False) else 0
        st = self.lib.c.vpho_mano_forward_ex(self.handle, capi.ptr(pose), capi.ptr(shape), n, capi.ptr(verts),
                                             capi.ptr(joints), flags, capi.stream_of(pose))
        self.lib.check(st, "vpho_mano_forward_ex")
        return verts, joints

    __call__ = get_hand_verts
