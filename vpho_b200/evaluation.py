"""The step after the hot path in `Trainer.evaluate` (lib/engine/train_diff_hand_obj.py:224-269, 333-357), on the device.

The reference moves every prediction to the host (`to_numpy`), runs `TesterHand` / `TesterObject` in numpy (with a
`.cuda()` round trip per image for the object distances) and gathers python dicts of metric arrays with
`gather_for_metrics(use_gather_object=True)` (pickle -> byte tensor -> NCCL all_gather).  Here the same per-image metrics
are computed by `libvpho_b200.so` kernels and packed into one fixed-width float64 row per image -- the ONLY payload of the
final NCCL gather and the only metric traffic that leaves the GPU.

Row layout (`EVAL_COLS`): for each evaluated hand prediction (`agg_candidate`, `one_candidate` = first diffusion sample,
and `regression` when the batch carries it; test_diff_hand :454-490) MJE, PA_MJE, MVE, PAMVE, JE[21] in millimetres
(`TesterHand`, lib/engine/test.py:585-680); for each evaluated object pose (`mean_pose` = aggregated, `one_candidate`;
test_diff_object :492-514) the 17 columns of `ObjectMetrics` (`TesterObject`, lib/engine/test.py:196-584; metres / px).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import capi
from .aggregation import OBJ_METRIC_COLS, Assets, ObjectMetrics, obj_6d_to_rt

HAND_COLS = ["MJE", "PA_MJE", "MVE", "PAMVE"] + [f"MJE_{i}" for i in range(21)]


def symmetry_tables(model_infos, max_sym_disc_step: float = 0.01):
    """The padded symmetry-transform tables `TesterObject.__init__` builds from assets_models_info.json
    (lib/engine/test.py:205-232 with get_symmetry_transformations :97-152): per object the identity and its discrete
    symmetries, each composed with the continuous symmetries discretised into ceil(pi / max_sym_disc_step) rotations about
    the given axis; shorter lists are padded with identities; translations millimetres -> metres.
    -> sym_R (N, K, 3, 3), sym_t (N, K, 3) float64, sym_count (N,) int32  (inputs of `ObjectMetrics`)."""
    import math
    import numpy as np

    def about_axis(angle, axis):
        a = np.asarray(axis[:3], np.float64)
        a = a / math.sqrt(float(a @ a))
        c, s = math.cos(angle), math.sin(angle)
        skew = np.array([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]])
        return np.diag([c, c, c]) + np.outer(a, a) * (1.0 - c) + skew * s

    per = []
    for info in model_infos:
        discrete = [(np.eye(3), np.zeros(3))]
        for sym in info.get("symmetries_discrete", []):
            m = np.asarray(sym, np.float64).reshape(4, 4)
            discrete.append((m[:3, :3], m[:3, 3]))
        continuous = []
        for sym in info.get("symmetries_continuous", []):
            off = np.asarray(sym["offset"], np.float64).reshape(3)
            n_steps = int(np.ceil(np.pi / max_sym_disc_step))
            for i in range(1, n_steps):
                R = about_axis(i * (2.0 * np.pi / n_steps), sym["axis"])
                continuous.append((R, off - R @ off))
        combos = [(Rc @ Rd, Rc @ td + tc) for Rd, td in discrete for Rc, tc in continuous] if continuous else discrete
        per.append(combos)
    K = max(len(c) for c in per)
    sym_R = np.tile(np.eye(3), (len(per), K, 1, 1))
    sym_t = np.zeros((len(per), K, 3))
    for i, combos in enumerate(per):
        for k, (R, t) in enumerate(combos):
            sym_R[i, k], sym_t[i, k] = R, t / 1000.0
    return sym_R, sym_t, np.asarray([len(c) for c in per], np.int32)


def postprocess_hand_vert(vert: torch.Tensor, root_joint: torch.Tensor, is_right: torch.Tensor) -> torch.Tensor:
    """`Trainer.__postprocess_hand_vert` (train_diff_hand_obj.py:598-602): un-flip left hands, add the root joint.
    vert (bs, ..., V, 3) wrist-relative in the flipped frame -> camera frame.  Returns a new tensor."""
    bs = vert.shape[0]
    sign = torch.where(is_right.reshape(bs, *([1] * (vert.dim() - 1))), 1.0, -1.0).to(vert.dtype)
    out = vert.clone()
    out[..., 0] = out[..., 0] * sign[..., 0]
    return out + root_joint.reshape(bs, *([1] * (vert.dim() - 2)), 3).to(vert.dtype)


def hand_metrics_mm(pd_joint, gt_joint, pd_vert, gt_vert, lib=None) -> torch.Tensor:
    """(n, 25) float32 = MJE, PA_MJE, MVE, PAMVE, JE[21] in mm for one prediction per image (TesterHand.__call__)."""
    lib = lib or capi.lib()
    f = lambda t: t.contiguous().float()   # noqa: E731
    pj, gj, pv, gv = f(pd_joint), f(gt_joint), f(pd_vert), f(gt_vert)
    n = pj.shape[0]
    out = torch.empty((n, 25), dtype=torch.float32, device=pj.device)
    lib.check(lib.c.vpho_hand_metrics(capi.ptr(pj), capi.ptr(gj), capi.ptr(pv), capi.ptr(gv), n, capi.ptr(out),
                                      capi.stream_of(pj)), "vpho_hand_metrics")
    return out


class EvalRecorder:
    """Builds the per-image evaluation row of a predict() result on the device."""

    def __init__(self, assets: Assets, metric_tables: Dict, with_regression: bool = False):
        self.assets = assets
        self.obj_metrics = ObjectMetrics(assets, metric_tables)
        self.hand_sets = ["agg_candidate", "one_candidate"] + (["regression"] if with_regression else [])
        self.obj_sets = ["mean_pose", "one_candidate"]
        self._ws: Optional[torch.Tensor] = None
        self.cols: List[str] = [f"hand/{s}/{c}" for s in self.hand_sets for c in HAND_COLS] + \
                               [f"obj/{s}/{c}" for s in self.obj_sets for c in OBJ_METRIC_COLS]

    @property
    def width(self) -> int:
        return len(self.cols)

    @torch.no_grad()
    def __call__(self, pd: Dict, batch: Dict, reg_vert: Optional[torch.Tensor] = None,
                 reg_joint: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pd: output of VphoHotPath.predict.  batch: device tensors root_joint (bs,3), is_right (bs,), gt_joint (bs,21,3),
        gt_hand_vert (bs,778,3), gt_obj_rt (bs,3,4), cam_intr (bs,3,3), obj_id / obj_name.  -> (bs, width) float64.
        One `vpho_eval_record` call (five launches on the current stream): no intermediate tensor leaves the library."""
        lib = self.assets.lib
        with_reg = "regression" in self.hand_sets
        if with_reg and (reg_vert is None or reg_joint is None):
            raise ValueError("this recorder was built with_regression=True: pass reg_vert and reg_joint")
        f32 = lambda t: t.contiguous().float()      # noqa: E731  (no-ops for predict()'s own outputs)
        f64 = lambda t: t.contiguous().double()     # noqa: E731
        aj, av = f32(pd["agg_hand_joint"]), f32(pd["agg_hand_vert"])
        cj, cv = f32(pd["diff_final_hand_joint"]), f32(pd["diff_final_hand_vert"])
        ao, co = f64(pd["agg_obj_6d"]), f64(pd["diff_final_obj_6d"])
        bs, S, dev = cj.shape[0], cj.shape[1], cj.device
        root, right = f32(batch["root_joint"]), batch["is_right"].contiguous().to(torch.uint8)
        gj, gv, grt, K = f32(batch["gt_joint"]), f32(batch["gt_hand_vert"]), f64(batch["gt_obj_rt"]), f32(batch["cam_intr"])
        ids = self.assets.ids(batch["obj_id"] if "obj_id" in batch else batch["obj_name"], dev)
        rj, rv = (f32(reg_joint), f32(reg_vert)) if with_reg else (None, None)
        out = torch.empty((bs, self.width), dtype=torch.float64, device=dev)
        need = int(lib.c.vpho_eval_record_workspace_bytes(bs))
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        P = capi.ptr
        args = capi.EvalRecordArgs(
            bs=bs, S=S, agg_hand_joint=P(aj), agg_hand_vert=P(av), cand_hand_joint=P(cj), cand_hand_vert=P(cv),
            reg_hand_joint=P(rj) if with_reg else None, reg_hand_vert=P(rv) if with_reg else None, agg_obj_6d=P(ao),
            cand_obj_6d=P(co), root_joint=P(root), is_right=P(right), gt_joint=P(gj), gt_vert=P(gv), gt_obj_rt=P(grt),
            cam_intr=P(K), obj_id=P(ids), out=P(out))
        lib.check(lib.c.vpho_eval_record(self.assets.handle, self.obj_metrics.handle, C.byref(args), P(self._ws),
                                         self._ws.numel(), capi.stream_of(out)), "vpho_eval_record")
        return out

    @torch.no_grad()
    def unfused(self, pd: Dict, batch: Dict, reg_vert: Optional[torch.Tensor] = None,
                reg_joint: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The same row assembled from the individual entry points (`vpho_hand_metrics`, `vpho_object_metrics`) with the
        post-processing in torch: the cross-check of `vpho_eval_record` in tests/test_evaluation.py."""
        root, is_right = batch["root_joint"], batch["is_right"].bool()
        lib = self.assets.lib
        parts = []
        hand_preds = {"agg_candidate": (pd["agg_hand_joint"], pd["agg_hand_vert"]),
                      "one_candidate": (pd["diff_final_hand_joint"][:, 0], pd["diff_final_hand_vert"][:, 0]),
                      "regression": (reg_joint, reg_vert)}
        for s in self.hand_sets:
            j, v = hand_preds[s]
            j, v = postprocess_hand_vert(j, root, is_right), postprocess_hand_vert(v, root, is_right)
            parts.append(hand_metrics_mm(j, batch["gt_joint"], v, batch["gt_hand_vert"], lib=lib).double())
        ids = batch["obj_id"] if "obj_id" in batch else batch["obj_name"]
        obj6d = torch.stack([pd["agg_obj_6d"].double(), pd["diff_final_obj_6d"][:, 0].double()], dim=1)    # (bs, 2, 9)
        rt = obj_6d_to_rt(obj6d, root.double())
        om = self.obj_metrics(rt, batch["gt_obj_rt"], ids, batch["cam_intr"])                               # (bs, 2, 17)
        parts.append(om.reshape(om.shape[0], -1))
        return torch.cat(parts, dim=1).contiguous()


def summarize(table: torch.Tensor, cols: List[str]) -> Dict[str, float]:
    """Rank-0 reduction of the gathered table: instance means, in the units of the reference's printed tables (mm for the
    hand and for object distances, px for REP, % for rates; TesterObject.format, test.py:569-583)."""
    t = table.double().mean(0)
    out = {}
    for c, v in zip(cols, t.tolist()):
        name = c.rsplit("/", 1)[1]
        if c.startswith("obj/") and name in ("MCE", "OCE", "MCE2", "SMCE", "ADD", "ADDS", "CD"):
            v *= 1000.0
        elif c.startswith("obj/") and (name.startswith("FSCORE") or name in ("ADD01d", "ADDS01d", "REP5")):
            v *= 100.0
        out[c] = v
    return out
