"""TEST INFRASTRUCTURE -- CPU oracle of the object-pose metrics of `TesterObject` (lib/engine/test.py:196-584), the step
right after the hot path in `Trainer.evaluate` (lib/engine/train_diff_hand_obj.py:236-259).  SURVEY.md §8(f) N3.

Restated per (image, candidate) so that one call covers the reference's loop over images:
  MCE, OCE            criterion_MCE_OCE   lib/engine/test.py:354-375   (8 box corners, numpy float64)
  SMCE                criterion_SMCE      lib/engine/test.py:377-399   (min over the object's symmetry transforms)
  MCE2                criterion_MCE2      lib/engine/test.py:401-417 -> compute_obj_metrics_dexycb :155-193 (axis-aligned
                                          boxes of the two posed clouds, float32)
  ADD, ADD-S, REP     criterion_ADD_REP   lib/engine/test.py:419-451   (float32 distances; float64 projection)
  F-score x6, CD      criterion_FSCORE    lib/engine/test.py:453-503
  ADD01d, ADDS01d     cal_ADD01d          lib/engine/test.py:505-518 ;  REP5  cal_REP5 :520-521
Symmetry tables: get_symmetry_transformations (lib/engine/test.py:97-152, from ArtiBoost / BOP toolkit) and the padding
of TesterObject.__init__ (:205-232).

Declared deviation (same as SURVEY.md Appendix A.4): torch.cdist runs in the exact mode.
Only tests/ and oracle/make_golden.py import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch

COLS = ("MCE", "OCE", "MCE2", "SMCE", "ADD", "ADDS", "REP", "CD", "FSCORE@2mm", "FSCORE@5mm", "FSCORE@10mm", "FSCORE@2cm",
        "FSCORE@5cm", "FSCORE@10cm", "ADD01d", "ADDS01d", "REP5")
F_THRESHOLDS = (0.002, 0.005, 0.010, 0.020, 0.050, 0.100)
MAX_SYM_DISC_STEP = 0.01          # TesterObject.max_sym_disc_step (test.py:206)


def axis_rotation(angle: float, direction: Sequence[float]) -> np.ndarray:
    """3x3 rotation by `angle` about `direction` (rotation_matrix, test.py:60-94, origin-centred part)."""
    d = np.array(direction[:3], dtype=np.float64)
    d = d / math.sqrt(float(np.dot(d, d)))
    s, c = math.sin(angle), math.cos(angle)
    R = np.diag([c, c, c]) + np.outer(d, d) * (1.0 - c)
    d = d * s
    return R + np.array([[0.0, -d[2], d[1]], [d[2], 0.0, -d[0]], [-d[1], d[0], 0.0]])


def symmetry_transformations(model_info: dict, max_sym_disc_step: float = MAX_SYM_DISC_STEP) -> List[Dict[str, np.ndarray]]:
    """test.py:97-152: identity + discrete symmetries, each composed with the discretised continuous ones."""
    disc = [(np.eye(3), np.zeros((3, 1)))]
    for sym in model_info.get("symmetries_discrete", []):
        m = np.reshape(sym, (4, 4))
        disc.append((m[:3, :3], m[:3, 3].reshape(3, 1)))
    cont = []
    for sym in model_info.get("symmetries_continuous", []):
        axis = np.array(sym["axis"])
        offset = np.array(sym["offset"]).reshape(3, 1)
        steps = int(np.ceil(np.pi / max_sym_disc_step))
        step = 2.0 * np.pi / steps
        for i in range(1, steps):
            R = axis_rotation(i * step, axis)
            cont.append((R, -R.dot(offset) + offset))
    out = []
    for Rd, td in disc:
        if cont:
            for Rc, tc in cont:
                out.append({"R": Rc.dot(Rd), "t": Rc.dot(td) + tc})
        else:
            out.append({"R": Rd, "t": td})
    return out


def symmetry_tables(model_infos: Sequence[dict]):
    """TesterObject.__init__ (test.py:208-232): per-object transforms padded with identities to a common count;
    translations mm -> m.  -> R (N, K, 3, 3), t (N, K, 3) float64, count (N,)."""
    per = [symmetry_transformations(mi) for mi in model_infos]
    K = max(len(p) for p in per)
    R = np.tile(np.eye(3), (len(per), K, 1, 1))
    t = np.zeros((len(per), K, 3))
    for i, p in enumerate(per):
        for k, tr in enumerate(p):
            R[i, k] = tr["R"]
            t[i, k] = tr["t"].reshape(3) / 1000.0
    return R, t, np.array([len(p) for p in per], np.int32)


def _pose(points: np.ndarray, rt: np.ndarray) -> np.ndarray:
    """points (P,3), rt (...,3,4) -> (...,P,3):  p R^T + t  in float64 (the einsum of test.py:366,391,411,435,474)."""
    return np.einsum("ni,...ij->...nj", points, np.swapaxes(rt[..., :3], -1, -2)) + rt[..., 3][..., None, :]


def _nearest(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """(C,P,3), (Q,3) float32 -> (C,P) distance of every a-point to its nearest b-point, exact (non-matmul) cdist."""
    return torch.cdist(a, b[None].expand(a.shape[0], -1, -1), compute_mode="donot_use_mm_for_euclid_dist").min(-1)[0]


def object_metrics(tables: dict, pd_rt: np.ndarray, gt_rt: np.ndarray, obj_id: Sequence[int], cam_intr: np.ndarray) -> np.ndarray:
    """tables: bbox3d (N,8,3), verts_sampled (N,P,3), verts (N,Q,3) [F-score / CD cloud], diameter (N,), sym_R (N,K,3,3),
    sym_t (N,K,3).  pd_rt (n,C,3,4), gt_rt (n,3,4), cam_intr (n,3,3)  ->  (n, C, 17) float64 in the order of COLS
    (metres, pixels, fractions, 0/1 flags)."""
    pd_rt, gt_rt, cam_intr = np.asarray(pd_rt, np.float64), np.asarray(gt_rt, np.float64), np.asarray(cam_intr, np.float64)
    n, C = pd_rt.shape[:2]
    out = np.zeros((n, C, len(COLS)))
    for i in range(n):
        o = int(obj_id[i])
        box = np.asarray(tables["bbox3d"][o], np.float64)
        pd_box, gt_box = _pose(box, pd_rt[i]), _pose(box, gt_rt[i])
        out[i, :, 0] = np.linalg.norm(pd_box - gt_box, axis=-1).mean(-1)
        out[i, :, 1] = np.linalg.norm(pd_box.mean(-2) - gt_box.mean(-2), axis=-1)
        # SMCE: symmetric copies of the box, posed by the ground truth; best (smallest) mean corner distance
        sR, st = np.asarray(tables["sym_R"][o], np.float64), np.asarray(tables["sym_t"][o], np.float64)
        sym_box = np.einsum("ni,kji->knj", box, sR) + st[:, None, :]
        gt_sym = np.einsum("kni,ij->knj", sym_box, gt_rt[i][:, :3].T) + gt_rt[i][:, 3][None, None, :]
        out[i, :, 3] = np.linalg.norm(pd_box[:, None] - gt_sym[None], axis=-1).mean(-1).min(-1)
        # sampled-surface metrics (float32 clouds, as `.float().cuda()` makes them)
        vs = np.asarray(tables["verts_sampled"][o], np.float64)
        pd_v, gt_v = _pose(vs, pd_rt[i]), _pose(vs, gt_rt[i])
        pd_t, gt_t = torch.from_numpy(pd_v).float(), torch.from_numpy(gt_v).float()
        out[i, :, 5] = _nearest(pd_t, gt_t).mean(-1).numpy()
        out[i, :, 4] = torch.norm(pd_t - gt_t, dim=-1).mean(-1).numpy()
        K = cam_intr[i]
        pd_p = (np.einsum("...ni,ij->...nj", pd_v, K.T) / (pd_v[..., 2:3] + 1e-7))[..., :2]
        gt_p = (np.einsum("ni,ij->nj", gt_v, K.T) / (gt_v[..., 2:3] + 1e-7))[..., :2]
        out[i, :, 6] = np.linalg.norm(pd_p - gt_p, axis=-1).mean(-1)
        # MCE2: axis-aligned boxes of the two posed clouds (compute_obj_metrics_dexycb), float32
        ci = torch.tensor([[0, 1, 0, 0, 1, 0, 1, 1], [0, 0, 1, 0, 1, 1, 0, 1], [0, 0, 0, 1, 0, 1, 1, 1]])

        def aabb(m):                                      # (C,P,3) -> (C,8,3)
            mm = torch.stack([m.min(dim=1)[0], m.max(dim=1)[0]], dim=2)          # (C,3,2)
            return torch.stack([mm[:, 0, ci[0]], mm[:, 1, ci[1]], mm[:, 2, ci[2]]], dim=2)
        out[i, :, 2] = (aabb(pd_t) - aabb(gt_t[None]).float()).norm(2, -1).mean(-1).numpy()
        # F-score / Chamfer on the full cloud
        vf = np.asarray(tables["verts"][o], np.float64)
        pf = torch.from_numpy(_pose(vf, pd_rt[i])).float()
        gf = torch.from_numpy(_pose(vf, gt_rt[i])).float()
        dmat = torch.cdist(pf, gf[None].expand(C, -1, -1), compute_mode="donot_use_mm_for_euclid_dist")
        d_pg, d_gp = dmat.min(dim=2)[0], dmat.min(dim=1)[0]
        out[i, :, 7] = (0.5 * (d_pg.mean(dim=1) + d_gp.mean(dim=1))).numpy()
        for j, th in enumerate(F_THRESHOLDS):
            prec, rec = (d_pg < th).float().mean(dim=1), (d_gp < th).float().mean(dim=1)
            out[i, :, 8 + j] = ((2 * prec * rec) / (prec + rec + 1e-6)).numpy()
        diam = float(tables["diameter"][o])
        out[i, :, 14] = out[i, :, 4] <= diam * 0.1
        out[i, :, 15] = out[i, :, 5] <= diam * 0.1
        out[i, :, 16] = out[i, :, 6] < 5
    return out


def synthetic_metric_tables(objects: dict, seed: int = 7) -> dict:
    """Per-object tables the reference reads from YCB_MESHES / assets_models_info.json, manufactured from the synthetic
    point tables: box corners and diameter of the sampled surface, and a mix of symmetry classes (none, one discrete
    half-turn, a continuous axis) so that every branch of get_symmetry_transformations is exercised."""
    verts = np.asarray(objects["verts_sampled"], np.float64)
    N = verts.shape[0]
    lo, hi = verts.min(1), verts.max(1)
    corners = np.array([[(hi if (c >> a) & 1 else lo)[:, a] for a in range(3)] for c in range(8)])     # (8,3,N)
    bbox3d = np.transpose(corners, (2, 0, 1))
    diameter = np.sqrt(((hi - lo) ** 2).sum(-1))
    infos = []
    for i in range(N):
        mi = {"diameter": float(diameter[i] * 1000)}
        if i % 3 == 1:
            mi["symmetries_discrete"] = [[-1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]]
        if i % 3 == 2:
            mi["symmetries_continuous"] = [{"axis": [0, 0, 1], "offset": [0, 0, 0]}]
        if i % 6 == 5:
            mi["symmetries_discrete"] = [[1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1]]
        infos.append(mi)
    sR, st, cnt = symmetry_tables(infos)
    return {"bbox3d": bbox3d.astype(np.float32), "verts_sampled": np.asarray(objects["verts_sampled"], np.float32),
            "verts": np.asarray(objects["verts_sampled"], np.float32), "diameter": diameter.astype(np.float32),
            "sym_R": sR, "sym_t": st, "sym_count": cnt, "model_info": infos}
