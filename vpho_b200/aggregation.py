"""Host-side mirrors of the reference's scoring / aggregation interfaces over the CUDA kernels (csrc/aggregate.cu).

  * `Assets`          force-anchor tables + per-object point tables on the device
                      (lib/utils/physics_fn.py:121-140, lib/utils/hand_fn.py:427-433, lib/model/head_object.py:9-34)
  * `HeadObject`      <- `HeadObject.forward` / `flip_pt3d` (lib/model/head_object.py:36-67)
  * `HeadPhysics`     <- `from_local_to_global` (lib/model/physics.py:500 -> :362-371), evaluation only
  * `HOI_Aggregator`  <- `HOI_Aggregator.__call__(**kwargs) -> dict` (lib/model/aggregation.py:1167-1353)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import capi
from .head_mano import HeadMano

PHY_TOPK = 5   # hard-coded in the reference (aggregation.py:1246)


class Assets:
    def __init__(self, anchors: Dict[str, np.ndarray], objects: Dict[str, object], lib: Optional[capi.Library] = None):
        self.lib = lib or capi.lib()
        self.names: List[str] = list(objects["names"])
        face = np.ascontiguousarray(anchors["face_vertex_idx"], dtype=np.int32).reshape(32, 3)
        aw = np.ascontiguousarray(anchors["anchor_weight"], dtype=np.float32).reshape(32, 2)
        v2j = np.ascontiguousarray(anchors["vert2joint"], dtype=np.float32).reshape(21, 778)
        kpt = np.ascontiguousarray(objects["kpt3d"], dtype=np.float32)
        verts = np.ascontiguousarray(objects["verts_sampled"], dtype=np.float32)
        com = np.ascontiguousarray(objects["CoM"], dtype=np.float32)
        self.n_obj, self.n_pts = verts.shape[0], verts.shape[1]
        assert kpt.shape == (self.n_obj, 27, 3) and com.shape == (self.n_obj, 3)
        h = C.c_void_p()
        self.lib.check(self.lib.c.vpho_assets_create(capi.host_ptr(face), capi.host_ptr(aw), capi.host_ptr(v2j), self.n_obj,
                                                     self.n_pts, capi.host_ptr(kpt), capi.host_ptr(verts),
                                                     capi.host_ptr(com), C.byref(h)), "vpho_assets_create")
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.c.vpho_assets_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def ids(self, names, device) -> torch.Tensor:
        """Object ids as the contiguous int32 device tensor the kernels read (`const int32_t*`).  `names`: a sequence of
        object names, or integer ids as a list / numpy array / tensor of any integer dtype.  Ids that live on the host are
        range-checked here; ids already on the device are converted without a host sync and clamped by the kernels."""
        if isinstance(names, torch.Tensor):
            if names.dtype in (torch.float16, torch.float32, torch.float64, torch.bool):
                raise capi.VphoError(f"object ids must be integers, got {names.dtype}")
            if not names.is_cuda and names.numel() and (int(names.min()) < 0 or int(names.max()) >= self.n_obj):
                raise capi.VphoError(f"object id out of range [0, {self.n_obj})")
            return names.to(device=device, dtype=torch.int32).contiguous()
        if isinstance(names, np.ndarray):
            return self.ids(torch.from_numpy(np.ascontiguousarray(names)), device)
        if len(names) and not isinstance(names[0], str):
            return self.ids(torch.as_tensor(list(names)), device)
        return torch.tensor([self.names.index(n) for n in names], dtype=torch.int32, device=device)


class HeadObject:
    _WHICH = {"keypoint": 0, "verts": 1, "CoM": 2}

    def __init__(self, assets: Assets):
        self.assets = assets

    def __call__(self, pose: torch.Tensor, name, data_name: str = "keypoint", is_right: Optional[torch.Tensor] = None):
        """pose (bs, ..., 9) -> points (bs, ..., V, 3).  `is_right` (optional) fuses `flip_pt3d` into the same launch."""
        lib, a = self.assets.lib, self.assets
        lead = pose.shape[:-1]
        bs = lead[0]
        p = pose.reshape(bs, -1, 9).contiguous().float()
        Cn = p.shape[1]
        ids = a.ids(name, p.device)
        which = self._WHICH[data_name]
        V = (27, a.n_pts, 1)[which]
        out = torch.empty((bs, Cn, V, 3), dtype=torch.float32, device=p.device)
        ir = None if is_right is None else is_right.to(torch.uint8).contiguous()
        lib.check(lib.c.vpho_object_points(a.handle, capi.ptr(p), capi.ptr(ids), capi.ptr(ir), bs, Cn, which,
                                           0 if ir is None else 1, capi.ptr(out), capi.stream_of(p)), "vpho_object_points")
        return out.reshape(*lead, V, 3)

    forward = __call__

    @staticmethod
    def flip_pt3d(pt3d: torch.Tensor, is_right: torch.Tensor):
        idx = torch.arange(pt3d.shape[0], device=pt3d.device)[~is_right]
        pt3d[idx, ..., 0] = pt3d[idx, ..., 0] * -1
        return pt3d


class HeadPhysics:
    def __init__(self, assets: Assets):
        self.assets = assets

    def from_local_to_global(self, force_local: torch.Tensor, hand_vert: torch.Tensor):
        """force_local (..., 32, 3), hand_vert (..., 778, 3) camera frame -> force_point, force_global (..., 32, 3)."""
        lib, a = self.assets.lib, self.assets
        lead = hand_vert.shape[:-2]
        v = hand_vert.reshape(-1, 778, 3).contiguous().float()
        fl = force_local.reshape(-1, 32, 3).contiguous().float()
        n = v.shape[0]
        assert fl.shape[0] == n
        fp = torch.empty((n, 32, 3), dtype=torch.float32, device=v.device)
        fg = torch.empty_like(fp)
        lib.check(lib.c.vpho_force_anchors(a.handle, capi.ptr(v), capi.ptr(fl), n, 1, capi.ptr(fp), capi.ptr(fg),
                                           capi.stream_of(v)), "vpho_force_anchors")
        return fp.reshape(*lead, 32, 3), fg.reshape(*lead, 32, 3)


class HOI_Aggregator:
    """Drop-in for `HOI_Aggregator` (lib/model/aggregation.py:1160-1353): same kwargs, same returned keys/dtypes."""

    def __init__(self, head_mano: HeadMano, assets: Assets, debug: bool = False):
        self.head_mano = head_mano
        self.assets = assets
        self.lib = assets.lib
        self.debug = debug
        self._ws = None
        self.last_debug: dict = {}

    def __call__(self, **kw):
        lib, a = self.lib, self.assets
        f32 = lambda t: t.contiguous().float()   # noqa: E731
        dev = kw["hand_heatmap"].device
        bs = kw["root_joint"].shape[0]
        Kh, Ko = int(kw["hand_topk"]), int(kw["obj_topk"])
        obj_pose = kw["obj_pose6d"].contiguous().double()
        S = obj_pose.shape[1]
        hand_pose_diff = f32(kw["hand_pose_diff"]).reshape(-1, 48)
        assert hand_pose_diff.shape[0] == bs * S
        hold = dict(
            cam=f32(kw["cam_intrinsic"]), rjf=f32(kw["root_joint_flip"]), rj=f32(kw["root_joint"]),
            is_right=kw["is_right"].to(torch.uint8).contiguous(), is_grasped=kw["is_grasped"].to(torch.uint8).contiguous(),
            force_local=f32(kw["force_local"]), pose_diff=hand_pose_diff, pose_reg=f32(kw["hand_pose_regression"]),
            shape=f32(kw["hand_shape"]).reshape(-1, 10), hm_h=f32(kw["hand_heatmap"]), bb_h=f32(kw["hand_bbox"]),
            obj_pose=obj_pose, hm_o=f32(kw["obj_heatmap"]), bb_o=f32(kw["obj_bbox"]),
            obj_id=a.ids(kw["obj_name"], dev))
        assert hold["hm_h"].shape == (bs, 21, 64, 64) and hold["hm_o"].shape == (bs, 27, 64, 64)
        kk, nc, omax = Ko * Ko, Kh + 1, max(S, Ko * Ko)
        out = dict(
            obj_agg_6d=torch.empty((bs, 9), dtype=torch.float64, device=dev),
            pose6d_candidate=torch.empty((bs, kk, 9), dtype=torch.float64, device=dev),
            agg_obj_vert=torch.empty((bs, a.n_pts, 3), dtype=torch.float32, device=dev),
            hand_agg_mano=torch.empty((bs, 58), dtype=torch.float32, device=dev),
            hand_agg_vert=torch.empty((bs, 778, 3), dtype=torch.float32, device=dev),
            hand_agg_joint=torch.empty((bs, 21, 3), dtype=torch.float32, device=dev))
        dbg = {}
        if self.debug:
            z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)   # noqa: E731
            dbg = dict(hand_score=z(4, bs, 2 * S, 5), hand_topk=z(4, bs, 5, Kh, dt=torch.int32), cascade_pose=z(bs, 48),
                       obj_score=z(4, bs * omax), obj_topk=z(4, bs, max(Ko, PHY_TOPK), dt=torch.int32), finger_score=z(bs, 5, nc),
                       finger_topk=z(bs, 5, PHY_TOPK, dt=torch.int32), force_point=z(bs, 32, 3), force_global=z(bs, 32, 3))
        nbytes = int(lib.c.vpho_hoi_workspace_bytes(bs, S, Kh, Ko, a.n_pts))
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        P = capi.ptr
        args = capi.HoiArgs(
            bs, S, Kh, Ko, PHY_TOPK, P(hold["cam"]), P(hold["rjf"]), P(hold["rj"]), P(hold["is_right"]),
            P(hold["is_grasped"]), P(hold["force_local"]), P(hold["pose_diff"]), P(hold["pose_reg"]), P(hold["shape"]),
            P(hold["hm_h"]), P(hold["bb_h"]), P(hold["obj_pose"]), P(hold["hm_o"]), P(hold["bb_o"]), P(hold["obj_id"], torch.int32),
            P(out["obj_agg_6d"]), P(out["pose6d_candidate"]), P(out["agg_obj_vert"]), P(out["hand_agg_mano"]),
            P(out["hand_agg_vert"]), P(out["hand_agg_joint"]),
            P(dbg.get("hand_score")), P(dbg.get("hand_topk")), P(dbg.get("cascade_pose")), P(dbg.get("obj_score")),
            P(dbg.get("obj_topk")), P(dbg.get("finger_score")), P(dbg.get("finger_topk")), P(dbg.get("force_point")),
            P(dbg.get("force_global")))
        st = lib.c.vpho_hoi_aggregate(self.head_mano.handle, a.handle, C.byref(args), capi.ptr(self._ws),
                                      self._ws.numel(), capi.stream_of(hold["hm_h"]))
        lib.check(st, "vpho_hoi_aggregate")
        self._hold = hold   # inputs stay alive until the next call (stream-ordered kernels may still read them)
        if self.debug:
            dbg["obj_score"] = dbg["obj_score"].reshape(4, bs * omax)
            self.last_debug = dbg
        return out


def anchor_contact(force_point: torch.Tensor, force_global: torch.Tensor, obj_points: torch.Tensor,
                   lib: Optional[capi.Library] = None):
    """Scoring part of `HandAggregator.select_by_physics` (lib/model/aggregation.py:553-590) for posed hands
    force_point / force_global (G, C, 32, 3) against obj_points (G, P, 3): -> nearest anchor distances (G, C, 32) and
    per-finger physics scores (G, C, 5)."""
    lib = lib or capi.lib()
    G, Cn = force_point.shape[0], force_point.shape[1]
    fp, fg = force_point.contiguous().float(), force_global.contiguous().float()
    ov = obj_points.contiguous().float()
    dist = torch.empty((G, Cn, 32), dtype=torch.float32, device=fp.device)
    score = torch.empty((G, Cn, 5), dtype=torch.float32, device=fp.device)
    lib.check(lib.c.vpho_anchor_contact(capi.ptr(fp), capi.ptr(fg), capi.ptr(ov), G * Cn, Cn, ov.shape[1], capi.ptr(dist),
                                        capi.ptr(score), capi.stream_of(fp)), "vpho_anchor_contact")
    return dist, score


def vertex_contact(verts: torch.Tensor, obj_points: torch.Tensor, lib: Optional[capi.Library] = None):
    """Dense stress variant: nearest object-point distance of every MANO vertex.  verts (G, C, 778, 3), obj_points
    (G, P, 3) -> (G, C, 778)."""
    lib = lib or capi.lib()
    G, Cn = verts.shape[0], verts.shape[1]
    v, ov = verts.contiguous().float(), obj_points.contiguous().float()
    dist = torch.empty((G, Cn, 778), dtype=torch.float32, device=v.device)
    lib.check(lib.c.vpho_vertex_contact(capi.ptr(v), capi.ptr(ov), G * Cn, Cn, ov.shape[1], capi.ptr(dist),
                                        capi.stream_of(v)), "vpho_vertex_contact")
    return dist


def cone_anchor_table(friction_coeff: float = 0.8) -> torch.Tensor:
    """The 8 friction-cone anchors of `HeadPhysics` (lib/model/physics.py:692-698) with xy scaled by the friction
    coefficient as `get_local_force` does (:549-550); a constant table, built with the same torch ops."""
    a = torch.arange(0, 2 * torch.pi, 2 * torch.pi / 8)[:8]
    anchor = torch.stack([torch.cos(a), torch.sin(a), torch.ones_like(a)], dim=-1) / 8
    anchor[:, :2] *= friction_coeff
    return anchor.contiguous()


def force_eval(assets: Assets, vert3d: torch.Tensor, scale: torch.Tensor, weight: torch.Tensor,
               contact_mask: Optional[torch.Tensor], force_contact: Optional[torch.Tensor], gravity: torch.Tensor,
               com: torch.Tensor, return_forces: bool = False):
    """Forward math of one `ForceOptimizer.optimize_batch` iteration (lib/engine/force_optimization.py:141-171) per posed
    hand: vert3d (n,778,3), scale (n,32), weight (n,32,8), contact_mask (n,32) bool, force_contact (n,32), gravity / com
    (n,1,3) or (n,3) -> terms (n,4) [+ force_local, force_point, force_global (n,32,3)]."""
    lib = assets.lib
    v = vert3d.reshape(-1, 778, 3).contiguous().float()
    n, dev = v.shape[0], v.device
    sc, w = scale.contiguous().float(), weight.contiguous().float()
    m = None if contact_mask is None else contact_mask.to(torch.uint8).contiguous()
    fc = None if force_contact is None else force_contact.contiguous().float()
    g, c = gravity.reshape(n, 3).contiguous().float(), com.reshape(n, 3).contiguous().float()
    cone = cone_anchor_table().to(dev)
    terms = torch.empty((n, 4), dtype=torch.float32, device=dev)
    outs = [torch.empty((n, 32, 3), dtype=torch.float32, device=dev) if return_forces else None for _ in range(3)]
    lib.check(lib.c.vpho_force_eval(assets.handle, capi.ptr(v), capi.ptr(sc), capi.ptr(w), capi.ptr(m), capi.ptr(fc),
                                    capi.ptr(cone), capi.ptr(g), capi.ptr(c), n, 1, capi.ptr(terms), capi.ptr(outs[0]),
                                    capi.ptr(outs[1]), capi.ptr(outs[2]), capi.stream_of(v)), "vpho_force_eval")
    return (terms, *outs) if return_forces else terms


def force_optimize(assets: Assets, vert3d: torch.Tensor, force_contact: torch.Tensor, gravity: torch.Tensor, com: torch.Tensor,
                   is_grasped: Optional[torch.Tensor] = None, n_iter: int = 3000, switch_iter: int = 300, lr: float = 1e-3,
                   return_losses: bool = False):
    """`ForceOptimizer.optimize_batch` for one batch (lib/engine/force_optimization.py:110-207) in one persistent kernel.
    vert3d (n,778,3), force_contact (n,32), gravity / com (n,3) or (n,1,3), already in the flipped frame (:134-137).
    -> dict(scale (n,32), weight (n,32,8), force_local, force_global (n,32,3) [, losses (n_iter,5)])."""
    lib = assets.lib
    v = vert3d.reshape(-1, 778, 3).contiguous().float()
    n, dev = v.shape[0], v.device
    fc = force_contact.contiguous().float()
    g, c = gravity.reshape(n, 3).contiguous().float(), com.reshape(n, 3).contiguous().float()
    ig = None if is_grasped is None else is_grasped.to(torch.uint8).contiguous()
    cone = cone_anchor_table().to(dev)
    out = dict(scale=torch.empty((n, 32), dtype=torch.float32, device=dev), weight=torch.empty((n, 32, 8), dtype=torch.float32, device=dev),
               force_local=torch.empty((n, 32, 3), dtype=torch.float32, device=dev),
               force_global=torch.empty((n, 32, 3), dtype=torch.float32, device=dev))
    losses = torch.empty((n_iter, 5), dtype=torch.float32, device=dev) if return_losses else None
    ws = torch.empty(max(int(lib.c.vpho_force_optimize_workspace_bytes(n)), 256), dtype=torch.uint8, device=dev)
    lib.check(lib.c.vpho_force_optimize(assets.handle, capi.ptr(v), capi.ptr(fc), capi.ptr(g), capi.ptr(c), capi.ptr(ig), capi.ptr(cone),
                                        n, int(n_iter), int(switch_iter), C.c_float(lr), capi.ptr(out["scale"]), capi.ptr(out["weight"]),
                                        capi.ptr(out["force_local"]), capi.ptr(out["force_global"]), capi.ptr(losses), capi.ptr(ws),
                                        ws.numel(), capi.stream_of(v)), "vpho_force_optimize")
    if return_losses:
        out["losses"] = losses
    return out


def hand_pa_metrics(pd_joint, gt_joint, pd_vert, gt_vert, lib=None) -> torch.Tensor:
    """Per-image Procrustes-aligned hand errors in mm, (n, 23) = PA-MJE, PA-MVE, JE[21]
    (TesterHand.criterion_MJE_PAMJE, lib/engine/test.py:657-679; rigid_align_AtoB, lib/utils/transform_fn.py:43-66)."""
    lib = lib or capi.lib()
    f = lambda t: t.contiguous().float()   # noqa: E731
    pj, gj, pv, gv = f(pd_joint), f(gt_joint), f(pd_vert), f(gt_vert)
    n = pj.shape[0]
    out = torch.empty((n, 23), dtype=torch.float32, device=pj.device)
    lib.check(lib.c.vpho_hand_pa_metrics(capi.ptr(pj), capi.ptr(gj), capi.ptr(pv), capi.ptr(gv), n, capi.ptr(out),
                                         capi.stream_of(pj)), "vpho_hand_pa_metrics")
    return out


def pose_metrics(assets: Assets, pd_joint, gt_joint, pd_vert, gt_vert, pd_obj6d, gt_obj6d, obj_name) -> torch.Tensor:
    """Per-image final pose error in mm, (n, 4) = MJE, MVE (TesterHand, lib/engine/test.py:657-679), ADD, ADD-S
    (TesterObject.criterion_ADD_REP, lib/engine/test.py:413-442), computed on the device so that the metric gather is the
    only traffic leaving the GPU."""
    lib = assets.lib
    f = lambda t: t.contiguous().float()   # noqa: E731
    pj, gj, pv, gv = f(pd_joint), f(gt_joint), f(pd_vert), f(gt_vert)
    po, go = pd_obj6d.contiguous().double(), gt_obj6d.contiguous().double()
    n, dev = pj.shape[0], pj.device
    ids = assets.ids(obj_name, dev)
    out = torch.empty((n, 4), dtype=torch.float32, device=dev)
    lib.check(lib.c.vpho_pose_metrics(assets.handle, capi.ptr(pj), capi.ptr(gj), capi.ptr(pv), capi.ptr(gv), capi.ptr(po),
                                      capi.ptr(go), capi.ptr(ids), n, capi.ptr(out), capi.stream_of(pj)), "vpho_pose_metrics")
    return out


OBJ_METRIC_COLS = ("MCE", "OCE", "MCE2", "SMCE", "ADD", "ADDS", "REP", "CD", "FSCORE@2mm", "FSCORE@5mm", "FSCORE@10mm",
                   "FSCORE@2cm", "FSCORE@5cm", "FSCORE@10cm", "ADD01d", "ADDS01d", "REP5")


class ObjectMetrics:
    """`TesterObject` (lib/engine/test.py:196-584) on the device: every per-candidate object metric of the reference's
    evaluate loop in one launch.  tables: bbox3d (n_obj,8,3), diameter (n_obj,), sym_R (n_obj,K,3,3), sym_t (n_obj,K,3)
    [metres, already padded with identities as TesterObject.__init__ does], optional sym_count, optional `verts`
    (n_obj,Q,3) for the F-score / Chamfer terms (default: the assets' sampled surface)."""

    def __init__(self, assets: Assets, tables: Dict[str, np.ndarray]):
        self.assets, self.lib = assets, assets.lib
        box = np.ascontiguousarray(tables["bbox3d"], dtype=np.float32).reshape(assets.n_obj, 8, 3)
        diam = np.ascontiguousarray(tables["diameter"], dtype=np.float32).reshape(assets.n_obj)
        sR = np.ascontiguousarray(tables["sym_R"], dtype=np.float64)
        K = sR.shape[1]
        st = np.ascontiguousarray(tables["sym_t"], dtype=np.float64).reshape(assets.n_obj, K, 3)
        cnt = tables.get("sym_count")
        cnt = None if cnt is None else np.ascontiguousarray(cnt, dtype=np.int32)
        fv = tables.get("verts_full")
        fv = None if fv is None else np.ascontiguousarray(fv, dtype=np.float32)
        h = C.c_void_p()
        self.lib.check(self.lib.c.vpho_objmetrics_create(
            assets.n_obj, capi.host_ptr(box), capi.host_ptr(diam), K, capi.host_ptr(sR), capi.host_ptr(st),
            None if cnt is None else capi.host_ptr(cnt), 0 if fv is None else fv.shape[1],
            None if fv is None else capi.host_ptr(fv), C.byref(h)), "vpho_objmetrics_create")
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.c.vpho_objmetrics_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def __call__(self, pd_rt: torch.Tensor, gt_rt: torch.Tensor, obj_name, cam_intr: torch.Tensor) -> torch.Tensor:
        """pd_rt (n, C, 3, 4) or (n, 3, 4), gt_rt (n, 3, 4), cam_intr (n, 3, 3) -> (n, C, 17) float64 (`OBJ_METRIC_COLS`)."""
        squeeze = pd_rt.dim() == 3
        pd = (pd_rt[:, None] if squeeze else pd_rt).contiguous().double()
        gt = gt_rt.contiguous().double()
        n, Cn, dev = pd.shape[0], pd.shape[1], pd.device
        K = cam_intr.contiguous().float()
        ids = self.assets.ids(obj_name, dev)
        out = torch.empty((n, Cn, len(OBJ_METRIC_COLS)), dtype=torch.float64, device=dev)
        self.lib.check(self.lib.c.vpho_object_metrics(self.assets.handle, self.handle, capi.ptr(pd), capi.ptr(gt),
                                                      capi.ptr(ids, torch.int32), capi.ptr(K), n, Cn, capi.ptr(out),
                                                      capi.stream_of(pd)), "vpho_object_metrics")
        return out[:, 0] if squeeze else out


def obj_6d_to_rt(pose6d: torch.Tensor, root_joint: torch.Tensor) -> torch.Tensor:
    """`Trainer.__postprocess_obj_rt` (lib/engine/train_diff_hand_obj.py:593-596): (bs, ..., 9) rot6d + translation ->
    (bs, ..., 3, 4) [R | t + root_joint]; tensor glue in the input dtype (Gram-Schmidt of pytorch3d's
    rotation_6d_to_matrix)."""
    a1, a2 = pose6d[..., :3], pose6d[..., 3:6]
    b1 = torch.nn.functional.normalize(a1, dim=-1)
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    R = torch.stack((b1, b2, b3), dim=-2)
    rj = root_joint.reshape(root_joint.shape[0], *([1] * (pose6d.dim() - 2)), 3).to(pose6d.dtype)
    t = pose6d[..., 6:] + rj
    return torch.cat([R, t[..., None]], dim=-1)
