"""Host-side mirror of the aggregation modes the predict branch does not use:

    HandAggregator.__call__(mode=...)      lib/model/aggregation.py:63-80
        'heatmap'                           :82-113        'heatmap_cascade' / 'heatmap_cascade_n_level'   :115-178, 469-535
        '2D_pt_pose' / '2D_pt_joint'        :286-377       'average_all' :379-424        'random' :426-467
    ObjectAggregator.__call__(mode=...)    :632-644
        'heatmap'                           :646-659       'heatmap_cascade' with is_force_selection=False   :661-722
        '2D_pt_pose'                        :1001-1052     'average_all' :1054-1082      'random' :1084-1112

Same kwargs, same returned keys.  Every number is produced by the sm_100a library through the C ABI (`vpho_mano_forward`,
`vpho_joint_scores`, `vpho_hand_level`, `vpho_quat_average_all`, `vpho_obj_select`, `vpho_object_points`); torch only
allocates, gathers rows by the returned indices and concatenates.  ('physics' / the force-selection branch of the object
cascade are the hot path itself: `HOI_Aggregator`, vpho_b200/aggregation.py.)
Top-k ties resolve to (value descending, index ascending), the rule the oracle pins the reference to.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch

from . import capi
from .aggregation import Assets, HeadObject
from .head_mano import HeadMano

# lib/utils/hand_fn.py:240-263
MANO_PARAMS_LEVEL = {0: [0, 1, 2],
                     1: [39, 40, 41, 3, 4, 5, 12, 13, 14, 30, 31, 32, 21, 22, 23],
                     2: [42, 43, 44, 6, 7, 8, 15, 16, 17, 33, 34, 35, 24, 25, 26],
                     3: [45, 46, 47, 9, 10, 11, 18, 19, 20, 36, 37, 38, 27, 28, 29]}
MANO_JOINT_LEVEL = {0: [0], 1: [1, 5, 9, 13, 17], 2: [2, 6, 10, 14, 18], 3: [3, 7, 11, 15, 19], 4: [4, 8, 12, 16, 20]}


def _i32(v) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(list(v), dtype=np.int32))


class HandAggregator:
    """`HandAggregator` (lib/model/aggregation.py:59-626) minus 'physics' (hot path, see HOI_Aggregator)."""

    def __init__(self, head_mano: HeadMano):
        self.head_mano = head_mano
        self.lib = head_mano.lib

    # -- device primitives ------------------------------------------------------------------------------------------------
    def joint_scores(self, joint, root_joint, cam, bbox, heatmap, want_dist2d: bool = False):
        """heat [bs][n][21] (and the 2D-point score when asked for) of every candidate's joints."""
        bs, n = joint.shape[0], joint.shape[1]
        dev = joint.device
        heat = torch.empty((bs, n, 21), dtype=torch.float32, device=dev)
        dist = torch.empty((bs, n, 21), dtype=torch.float32, device=dev) if want_dist2d else None
        ws = torch.empty((max(bs, 1) * 21 * 2,), dtype=torch.float32, device=dev) if want_dist2d else None
        hold = [t.contiguous().float() for t in (joint, root_joint, cam, bbox, heatmap)]
        assert hold[4].shape == (bs, 21, 64, 64), "heat-maps are (bs, 21, 64, 64)"
        P = capi.ptr
        a = capi.JointScoresArgs(bs, n, 21, P(hold[0]), P(hold[1]), P(hold[2]), P(hold[3]), P(hold[4]), P(heat), P(dist))
        st = self.lib.c.vpho_joint_scores(C.byref(a), P(ws), 0 if ws is None else ws.numel() * 4, capi.stream_of(heat))
        self.lib.check(st, "vpho_joint_scores")
        self._hold_js = hold
        return heat, dist

    def level(self, score, pose, K: int, fuse_index, observe_index, independent: bool, is_weight: bool, fuse_kind: int = 0,
              joint=None, write_back: bool = True):
        """One `select_topk_hand_by_observed_heatmap_and_fuse_by_index` on precomputed joint scores.  `pose` (bs, n, 48) is
        updated in place at `fuse_index` like the reference's view assignment (:235-236)."""
        bs, n = score.shape[0], score.shape[1]
        dev = score.device
        fi, oi = _i32(fuse_index), _i32(observe_index)
        lists = len(fi) // 3 if independent else 1
        val = torch.empty((bs, K, lists), dtype=torch.float32, device=dev)
        topk = torch.empty((bs, K, lists), dtype=torch.int32, device=dev)
        fused = torch.empty((bs, lists, 3) if fuse_kind == 1 else (bs, len(fi)), dtype=torch.float32, device=dev)
        if pose is not None:
            assert pose.is_contiguous() and pose.dtype == torch.float32 and tuple(pose.shape) == (bs, n, 48)
        P = capi.ptr
        a = capi.HandLevelArgs(bs, n, int(K), 21, P(score), P(pose), P(joint), capi.host_ptr(oi), len(oi), capi.host_ptr(fi), len(fi),
                               int(independent), int(is_weight), int(fuse_kind), int(write_back), P(val), P(topk), P(fused))
        st = self.lib.c.vpho_hand_level(C.byref(a), capi.stream_of(score))
        self.lib.check(st, "vpho_hand_level")
        return val, topk.long(), fused

    def _mano(self, pose, shape, bs):
        vert, joint = self.head_mano.get_hand_verts(pose=pose.reshape(-1, 48), shape=shape.reshape(-1, 10))
        return vert.reshape(bs, -1, 778, 3), joint.reshape(bs, -1, 21, 3)

    def _finish(self, fused_pose, shape_all, bs):
        shape = shape_all.reshape(bs, -1, 10)[:, 0].contiguous()
        fused_pose = fused_pose.contiguous()
        fv, fj = self.head_mano.get_hand_verts(pose=fused_pose, shape=shape)
        return torch.cat((fused_pose, shape), dim=-1), fv.reshape(bs, 778, 3), fj.reshape(bs, 21, 3)

    def select_topk_hand_by_observed_heatmap_and_fuse_by_index(self, **kw) -> Dict:
        """lib/model/aggregation.py:180-284.  `pose` is modified in place when it is a contiguous float32 tensor, as the
        reference's view assignment does."""
        bs = kw["heatmap"].shape[0]
        pose_in = kw["pose"]
        pose = pose_in.reshape(bs, -1, 48)
        if not (pose.is_contiguous() and pose.dtype == torch.float32):
            pose = pose.contiguous().float()
        shape = kw["shape"].contiguous().float()
        vert, joint = self._mano(pose, shape, bs)
        heat, _ = self.joint_scores(joint, kw["root_joint"], kw["cam_intrinsic"], kw["bbox"], kw["heatmap"])
        K, fi = int(kw["k"]), list(kw["fuse_index"])
        indep = bool(kw["is_independent"])
        pose_before = pose.clone()
        val, topk, fused = self.level(heat, pose, K, fi, kw["observe_index"], indep, bool(kw["is_weight"]))
        bidx = torch.arange(bs, device=pose.device)
        if not indep:
            val, topk = val[:, :, 0], topk[:, :, 0]
            b2 = bidx[:, None].repeat(1, K)
            topk_aa = pose_before[b2, topk][:, :, fi].reshape(bs, K, -1, 3)
            topk_vert, topk_joint = vert[b2, topk], joint[b2, topk]
        else:
            nj = len(fi) // 3
            b3 = bidx[:, None, None].repeat(1, K, nj)
            j3 = (torch.tensor(fi, dtype=torch.long, device=pose.device).reshape(-1, 3)[:, 0] // 3)[None, None].repeat(bs, K, 1)
            topk_aa = pose_before.reshape(bs, -1, 16, 3)[b3, topk, j3]
            topk_vert, topk_joint = vert[b3, topk], joint[b3, topk]
        return {"val": val, "topk": topk, "fused_idx_pose": fused, "topk_idx_pose_aa": topk_aa, "fused_pose": pose,
                "topk_vert": topk_vert, "topk_joint": topk_joint, "vert": vert, "joint": joint}

    # -- modes ----------------------------------------------------------------------------------------------------------------
    def __call__(self, **kwargs):
        mode = kwargs["mode"]
        if mode == "heatmap":
            return self.select_by_heatmap(**kwargs)
        if mode == "heatmap_cascade":
            return self.select_by_heatmap_cascade_n_level(**dict(kwargs, n_level=4))
        if "2D_pt" in mode:
            return self.select_by_2D_pt(**kwargs)
        if mode == "average_all":
            return self.average_all(**kwargs)
        if mode == "random":
            return self.random(**kwargs)
        if mode == "heatmap_cascade_n_level":
            return self.select_by_heatmap_cascade_n_level(**kwargs)
        raise NotImplementedError(mode)

    def select_by_heatmap(self, **kw) -> Dict:
        """:82-113"""
        bs = kw["root_joint"].shape[0]
        fd = self.select_topk_hand_by_observed_heatmap_and_fuse_by_index(
            pose=kw["pose"], shape=kw["shape"], root_joint=kw["root_joint"], cam_intrinsic=kw["cam_intrinsic"],
            heatmap=kw["heatmap"], bbox=kw["bbox"], k=kw["k"], fuse_index=list(range(48)), observe_index=list(range(21)),
            is_independent=False, is_weight=kw["is_weight"])
        mano, fv, fj = self._finish(fd["fused_pose"][:, 0], kw["shape"], bs)
        return {"topk": fd["topk"], "diff_topk_vert": fd["topk_vert"], "diff_topk_joint": fd["topk_joint"], "agg_hand_mano": mano,
                "agg_vert": fv, "agg_joint": fj, "fused_data_ls": [fd], "diff_vert": fd["topk_vert"], "diff_joint": fd["topk_joint"]}

    def select_by_heatmap_cascade_n_level(self, **kw) -> Dict:
        """:469-535 (n_level = 4 is select_by_heatmap_cascade, :115-178)"""
        n_level = int(kw.get("n_level", 2))
        bs = kw["root_joint"].shape[0]
        pose = kw["pose"].clone().float().reshape(bs, -1, 48)
        shape = kw["shape"].clone().float().reshape(bs, -1, 10)
        num_candidate = pose.shape[1]
        if kw["use_regression_as_candidate"]:
            extra = torch.zeros_like(pose) + kw["pose_regression"][:, None].float()
            pose = torch.cat((pose, extra), dim=1)
            shape = shape.repeat(1, 2, 1)
        pose = pose.contiguous()
        fds: List[Dict] = []
        for lv in range(min(n_level, 4)):
            fuse_idx = MANO_PARAMS_LEVEL[lv]
            observe = [j for l in range(lv + 1, 5) for j in MANO_JOINT_LEVEL[l]]
            if kw["use_regression_as_candidate"] and lv == 0:
                pose[:, num_candidate:, fuse_idx] = pose[:, :num_candidate, fuse_idx]
            fd = self.select_topk_hand_by_observed_heatmap_and_fuse_by_index(
                pose=pose, shape=shape, root_joint=kw["root_joint"], cam_intrinsic=kw["cam_intrinsic"], heatmap=kw["heatmap"],
                bbox=kw["bbox"], k=kw["k"], fuse_index=fuse_idx, observe_index=observe, is_independent=lv != 0,
                is_weight=kw["is_weight"])
            fd = dict(fd, fused_pose=fd["fused_pose"].clone())
            fds.append(fd)
        mano, fv, fj = self._finish(fds[-1]["fused_pose"][:, 0], kw["shape"], bs)
        return {"topk": fds[-1]["topk"], "diff_topk_vert": fds[-1]["topk_vert"], "diff_topk_joint": fds[-1]["topk_joint"],
                "agg_hand_mano": mano, "agg_vert": fv, "agg_joint": fj, "diff_vert": fds[0]["vert"], "diff_joint": fds[0]["joint"],
                "fused_data_ls": fds}

    def select_by_2D_pt(self, **kw) -> Dict:
        """:286-377"""
        bs = kw["heatmap"].shape[0]
        K = int(kw["k"])
        pose = kw["pose"].reshape(bs, -1, 48).contiguous().float()
        vert, joint = self._mano(pose, kw["shape"].contiguous().float(), bs)
        _, score = self.joint_scores(joint, kw["root_joint"], kw["cam_intrinsic"], kw["bbox"], kw["heatmap"], want_dist2d=True)
        bidx = torch.arange(bs, device=pose.device)
        if "pose" in kw["mode"]:
            _, topk, fused = self.level(score, pose, K, list(range(48)), list(range(21)), False, False, write_back=False)
            topk = topk[:, :, 0]
            b2 = bidx[:, None].repeat(1, K)
            mano, fv, fj = self._finish(fused, kw["shape"], bs)
            tv, tj = vert[b2, topk], joint[b2, topk]
            return {"topk": topk, "diff_topk_vert": tv, "diff_topk_joint": tj, "agg_hand_mano": mano, "agg_vert": fv,
                    "agg_joint": fj, "diff_vert": tv, "diff_joint": tj}
        _, topk, fused_joint = self.level(score, None, K, list(range(63)), list(range(21)), True, False, fuse_kind=1,
                                          joint=joint.contiguous(), write_back=False)
        b3 = bidx[:, None, None].repeat(1, K, 21)
        j3 = torch.arange(21, device=pose.device)[None, None].repeat(bs, K, 1)
        tj = joint[b3, topk, j3]
        tv = torch.zeros(bs, K, 778, 3, device=pose.device)
        return {"topk": topk, "diff_topk_vert": tv, "diff_topk_joint": tj, "agg_hand_mano": torch.zeros(bs, 58, device=pose.device),
                "agg_vert": torch.zeros(bs, 778, 3, device=pose.device), "agg_joint": fused_joint, "diff_vert": tv, "diff_joint": tj}

    def average_all(self, **kw) -> Dict:
        """:379-424"""
        bs = kw["heatmap"].shape[0]
        pose = kw["pose"].reshape(bs, -1, 48).contiguous().float()
        vert, joint = self._mano(pose, kw["shape"].contiguous().float(), bs)
        fused = torch.empty((bs, 16, 3), dtype=torch.float32, device=pose.device)
        st = self.lib.c.vpho_quat_average_all(capi.ptr(pose), bs, pose.shape[1], 16, capi.ptr(fused), capi.stream_of(pose))
        self.lib.check(st, "vpho_quat_average_all")
        mano, fv, fj = self._finish(fused.reshape(bs, 48), kw["shape"], bs)
        return {"topk": None, "diff_topk_vert": vert, "diff_topk_joint": joint, "agg_hand_mano": mano, "agg_vert": fv,
                "agg_joint": fj, "diff_vert": vert, "diff_joint": joint}

    def random(self, **kw) -> Dict:
        """:426-467 (the reference's "random" pick is candidate 0: the samples are i.i.d.)"""
        bs = kw["heatmap"].shape[0]
        pose = kw["pose"].reshape(bs, -1, 48).contiguous().float()
        vert, joint = self._mano(pose, kw["shape"].contiguous().float(), bs)
        mano, fv, fj = self._finish(pose[:, 0], kw["shape"], bs)
        return {"topk": None, "diff_topk_vert": vert, "diff_topk_joint": joint, "agg_hand_mano": mano, "agg_vert": fv,
                "agg_joint": fj, "diff_vert": vert, "diff_joint": joint}


class ObjectAggregator:
    """Every mode of `ObjectAggregator.__call__` (lib/model/aggregation.py:628-781, 1001-1113); 'heatmap_cascade' with force
    selection is the predict branch itself (HOI_Aggregator)."""

    def __init__(self, assets: Assets):
        self.assets = assets
        self.lib = assets.lib
        self.obj_layer = HeadObject(assets)

    def select(self, pose6d, K: int, is_weight: bool, kw: Dict, topk_in=None, score_kind: int = 0):
        """select_topk_object_by_heatmap + fuse_topk (:729-781) -> topk (bs, K) int64, weight (bs, K), fused (bs, 9) f64.
        `topk_in`: winners of an earlier selection, fused (plain mean) on THIS pose set."""
        bs, n = pose6d.shape[0], pose6d.shape[1]
        dev = pose6d.device
        hold = dict(pose=pose6d.contiguous().double(), root=kw["root_joint"].contiguous().float(),
                    cam=kw["cam_intrinsic"].contiguous().float(), bbox=kw["bbox"].contiguous().float(),
                    hm=kw["heatmap"].contiguous().float(), is_right=kw["is_right"].to(torch.uint8).contiguous(),
                    obj_id=self.assets.ids(kw["obj_name"], dev),
                    topk_in=None if topk_in is None else topk_in.to(torch.int32).contiguous())
        assert hold["hm"].shape == (bs, 27, 64, 64), "object heat-maps are (bs, 27, 64, 64)"
        topk = torch.empty((bs, K), dtype=torch.int32, device=dev)
        weight = torch.empty((bs, K), dtype=torch.float32, device=dev)
        fused = torch.empty((bs, 9), dtype=torch.float64, device=dev)
        ws = torch.empty((max(bs * n, 1) + bs * 27 * 2,), dtype=torch.float32, device=dev)
        P = capi.ptr
        a = capi.ObjSelectArgs(bs, n, int(K), P(hold["pose"]), P(hold["root"]), P(hold["cam"]), P(hold["bbox"]), P(hold["hm"]),
                               P(hold["is_right"]), P(hold["obj_id"], torch.int32), int(is_weight), int(score_kind),
                               P(hold["topk_in"]), P(topk), P(weight), P(fused))
        st = self.lib.c.vpho_obj_select(self.assets.handle, C.byref(a), P(ws), ws.numel() * 4, capi.stream_of(hold["pose"]))
        self.lib.check(st, "vpho_obj_select")
        self._hold = hold
        return topk.long(), weight, fused

    def _verts(self, pose6d_fused, kw):
        p = pose6d_fused.clone()
        p[..., 6:] = p[..., 6:] + kw["root_joint"]
        v = self.obj_layer(p, kw["obj_name"], data_name="verts")
        return self.obj_layer.flip_pt3d(v, kw["is_right"])

    def __call__(self, **kwargs):
        if kwargs["mode"] == "heatmap":
            return self.select_by_heatmap(**kwargs)
        if kwargs["mode"] == "heatmap_cascade":
            return self.select_by_heatmap_cascade(**kwargs)
        if "2D_pt" in kwargs["mode"]:
            return self.select_by_2D_pt(**kwargs)
        if kwargs["mode"] == "average_all":
            return self.average_all(**kwargs)
        if kwargs["mode"] == "random":
            return self.random(**kwargs)
        raise NotImplementedError(kwargs["mode"])

    def _result(self, fused, kw) -> Dict:
        fused = fused.float()
        return {"agg_6d": fused, "candidate_6d": kw["pose6d"], "agg_obj_vert": self._verts(fused, kw)}

    def select_by_2D_pt(self, **kw):
        """:1001-1052 ('2D_pt_pose'; like the reference, any other 2D_pt mode returns None)"""
        if "pose" not in kw["mode"]:
            return None
        _, _, fused = self.select(kw["pose6d"], kw["k"], False, kw, score_kind=1)
        return self._result(fused, kw)

    def average_all(self, **kw) -> Dict:
        """:1054-1082: despite the name, the plain mean of the FIRST k candidates"""
        bs, K = kw["pose6d"].shape[0], int(kw["k"])
        idx = torch.arange(K, device=kw["pose6d"].device)[None].repeat(bs, 1)
        _, _, fused = self.select(kw["pose6d"], K, False, kw, topk_in=idx)
        return self._result(fused, kw)

    def random(self, **kw) -> Dict:
        """:1084-1112: candidate 0 (through fuse_topk: its rotation goes 6D -> quaternion -> 6D)"""
        bs = kw["pose6d"].shape[0]
        idx = torch.zeros(bs, 1, dtype=torch.long, device=kw["pose6d"].device)
        _, _, fused = self.select(kw["pose6d"], 1, False, kw, topk_in=idx)
        return self._result(fused, kw)

    def select_by_heatmap(self, **kw) -> Dict:
        """:646-659 (fuse_topk without weights)"""
        _, _, fused = self.select(kw["pose6d"], kw["k"], False, kw)
        fused = fused.float()
        return {"agg_6d": fused, "candidate_6d": kw["pose6d"], "agg_obj_vert": self._verts(fused, kw)}

    def select_by_heatmap_cascade(self, **kw) -> Dict:
        """:661-722 with is_force_selection=False (the force-selection branch is HOI_Aggregator's object path)"""
        if kw.get("is_force_selection", False):
            raise NotImplementedError("is_force_selection=True is the predict branch: use HOI_Aggregator")
        ori = kw["pose6d"].clone().double()
        w = bool(kw["is_weight"])
        _, _, f1 = self.select(ori, kw["k"], w, kw)
        trans1 = f1[:, 6:]
        p = ori.clone()
        p[..., 6:] = ori[..., 6:] * 0 + trans1[:, None]
        _, _, f2 = self.select(p, kw["k"], w, kw)
        rot1 = f2[:, :6]
        p = ori.clone()
        p[..., :6] = ori[..., :6] * 0 + rot1[:, None]
        topk_trans2, _, _ = self.select(p, kw["k"], False, kw)
        p = ori.clone()
        p[..., 6:] = ori[..., 6:] * 0 + trans1[:, None]
        _, _, f4 = self.select(p, kw["k"], False, kw)          # rot2: selected and fused on this pose set
        # the reference fuses topk_trans2 on the LAST pose set too (kwargs['pose6d'] is not restored, :713-714)
        _, _, f3 = self.select(p, kw["k"], False, kw, topk_in=topk_trans2)
        fused = torch.cat([f4[:, :6], f3[:, 6:]], dim=-1).float()
        return {"agg_6d": fused, "pose6d_candidate": p, "agg_obj_vert": self._verts(fused, kw)}
