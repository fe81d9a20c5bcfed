"""TEST INFRASTRUCTURE -- CPU restatement (PyTorch, float32) of the aggregation modes the predict branch does not use
(SURVEY.md §8f row N4): lib/model/aggregation.py

    HandAggregator.select_topk_hand_by_observed_heatmap_and_fuse_by_index   :180-284
    HandAggregator.select_by_heatmap :82-113, select_by_heatmap_cascade(_n_level) :115-178, 469-535,
    select_by_2D_pt :286-377, average_all :379-424, random :426-467
    ObjectAggregator.select_by_heatmap :646-659, select_by_heatmap_cascade (is_force_selection=False) :661-722,
    select_by_2D_pt :1001-1052, average_all :1054-1082, random :1084-1112

Pinned against the reference's own `HandAggregator` / `ObjectAggregator` classes (oracle/make_golden_modes.py ->
tests/golden/agg_modes.npz, and live in the build container: tests/test_agg_modes.py) under the same declared rule as the
hot-path fixtures: `Tensor.topk` canonicalised to (value descending, index ascending).
Only tests/ may import this file.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from .vpho_oracle import (MANO_JOINT_LEVEL, MANO_PARAMS_LEVEL, OracleMano, OracleObject, average_quaternion, canonical_topk,
                          obj_fuse_topk, obj_heat_topk, project, sample_heat)   # (puts oracle/shims on sys.path)
from pytorch3d.transforms.rotation_conversions import axis_angle_to_quaternion, quaternion_to_axis_angle   # noqa: E402  oracle/shims


def hand_level(mano: OracleMano, pose, shape, root_joint, cam, heatmap, bbox, K, fuse_index, observe_index, independent: bool,
               is_weight: bool) -> Dict:
    """select_topk_hand_by_observed_heatmap_and_fuse_by_index (:180-284); `pose` (bs*n, 48) is updated in place."""
    bs = heatmap.shape[0]
    vert, joint = mano(pose, shape)
    vert, joint = vert.reshape(bs, -1, 778, 3), joint.reshape(bs, -1, 21, 3)
    pt2d = project(joint + root_joint[:, None, None], cam)
    bb = bbox[:, None, None, :]
    pt2d = 2 * (pt2d - bb[..., :2]) / (bb[..., 2:] - bb[..., :2]) - 1
    heat_val = sample_heat(heatmap, pt2d, observe_index)
    pose3 = pose.reshape(bs, -1, 48)
    fuse_index = list(fuse_index)
    if not independent:
        score = heat_val.sum(dim=-1)
        val, topk = canonical_topk(score, K, dim=1)
        weight = (val + 1e-8) / (val.sum(dim=1, keepdim=True) + 1e-8)
        bidx = torch.arange(bs)[:, None].repeat(1, K)
        topk_aa = pose3[bidx, topk][:, :, fuse_index].reshape(bs, K, -1, 3)
        quat = axis_angle_to_quaternion(topk_aa).permute(0, 2, 1, 3)
        fq = average_quaternion(quat, weight[:, None].expand(-1, quat.shape[1], -1) if is_weight else None)
        topk_vert, topk_joint = vert[bidx, topk], joint[bidx, topk]
    else:
        M, N = len(observe_index), len(fuse_index)
        score = heat_val.reshape(bs, -1, M // (N // 3), N // 3).mean(dim=-2)
        val, topk = canonical_topk(score, K, dim=1)
        weight = ((val + 1e-8) / (val.sum(dim=1, keepdim=True) + 1e-8)).permute(0, 2, 1)
        i1 = torch.arange(bs)[:, None, None].repeat(1, K, N // 3)
        i2 = torch.tensor(fuse_index, dtype=torch.long)[None, None].repeat(bs, K, 1).reshape(bs, K, -1, 3)[:, :, :, 0] // 3
        topk_aa = pose.reshape(bs, -1, 16, 3)[i1, topk, i2]
        quat = axis_angle_to_quaternion(topk_aa).permute(0, 2, 1, 3)
        fq = average_quaternion(quat, weight if is_weight else None)
        topk_vert, topk_joint = vert[i1, topk], joint[i1, topk]
    faa = quaternion_to_axis_angle(fq).reshape(bs, -1)
    pose3[:, :, fuse_index] = pose3[:, :, fuse_index] * 0 + faa[:, None]
    return {"score": score, "val": val, "topk": topk, "fused_idx_pose": faa, "topk_idx_pose_aa": topk_aa, "fused_pose": pose3,
            "topk_vert": topk_vert, "topk_joint": topk_joint, "vert": vert, "joint": joint}


def _finish(mano: OracleMano, fused_pose, shape, bs):
    sh = shape.reshape(bs, -1, 10)[:, 0]
    fv, fj = mano(fused_pose, sh)
    return torch.cat((fused_pose, sh), dim=-1), fv.reshape(bs, 778, 3), fj.reshape(bs, 21, 3)


def hand_heatmap(mano, pose, shape, root_joint, cam, heatmap, bbox, K, is_weight) -> Dict:
    """select_by_heatmap (:82-113)"""
    bs = root_joint.shape[0]
    fd = hand_level(mano, pose.clone(), shape, root_joint, cam, heatmap, bbox, K, list(range(48)), list(range(21)), False, is_weight)
    m, fv, fj = _finish(mano, fd["fused_pose"][:, 0], shape, bs)
    return {"topk": fd["topk"], "diff_topk_joint": fd["topk_joint"], "agg_hand_mano": m, "agg_vert": fv, "agg_joint": fj, "val": fd["val"],
            "score": fd["score"]}


def hand_cascade_n_level(mano, pose, pose_regression, shape, root_joint, cam, heatmap, bbox, K, is_weight, n_level,
                         use_regression: bool = True) -> Dict:
    """select_by_heatmap_cascade_n_level (:469-535); n_level = 4 is select_by_heatmap_cascade (:115-178)"""
    bs = root_joint.shape[0]
    pose = pose.clone().reshape(bs, -1, 48)
    shape_c = shape.clone().reshape(bs, -1, 10)
    n_cand = pose.shape[1]
    if use_regression:
        pose = torch.cat((pose, torch.zeros_like(pose) + pose_regression[:, None].clone()), dim=1)
        shape_c = shape_c.repeat(1, 2, 1)
    pose, shape_c = pose.reshape(-1, 48), shape_c.reshape(-1, 10)
    fds = []
    for lv in range(min(n_level, 4)):
        fuse_idx = MANO_PARAMS_LEVEL[lv]
        observe = [j for l in range(lv + 1, 5) for j in MANO_JOINT_LEVEL[l]]
        if use_regression and lv == 0:
            p3 = pose.view(bs, -1, 48)
            p3[:, n_cand:, fuse_idx] = p3[:, :n_cand, fuse_idx]
        fd = hand_level(mano, pose, shape_c, root_joint, cam, heatmap, bbox, K, fuse_idx, observe, lv != 0, is_weight)
        pose = fd["fused_pose"].reshape(-1, 48)
        fds.append(fd)
    m, fv, fj = _finish(mano, fds[-1]["fused_pose"][:, 0], shape, bs)
    return {"topk": fds[-1]["topk"], "diff_topk_joint": fds[-1]["topk_joint"], "agg_hand_mano": m, "agg_vert": fv, "agg_joint": fj,
            "levels": fds}


def hand_2d_pt(mano, pose, shape, root_joint, cam, heatmap, bbox, K, mode: str) -> Dict:
    """select_by_2D_pt (:286-377)"""
    bs, J, H, W = heatmap.shape
    vert, joint = mano(pose, shape)
    vert, joint = vert.reshape(bs, -1, 778, 3), joint.reshape(bs, -1, 21, 3)
    pose16 = pose.reshape(bs, -1, 16, 3)
    pt2d = project(joint + root_joint[:, None, None], cam)
    bb = bbox[:, None, None, :]
    pt2d = 2 * (pt2d - bb[..., :2]) / (bb[..., 2:] - bb[..., :2]) - 1
    X, Y = torch.arange(W) / (W - 1) * 2 - 1, torch.arange(H) / (H - 1) * 2 - 1
    XX, YY = torch.meshgrid(X, Y, indexing="ij")            # the reference's default
    XX, YY = XX[None, None].repeat(bs, J, 1, 1).reshape(bs, J, -1), YY[None, None].repeat(bs, J, 1, 1).reshape(bs, J, -1)
    ind = torch.argmax(heatmap.reshape(bs, J, -1), dim=-1)
    i1, i2 = torch.arange(bs)[:, None].repeat(1, J), torch.arange(J)[None].repeat(bs, 1)
    pt_hm = torch.stack([XX[i1, i2, ind], YY[i1, i2, ind]], dim=-1)
    score = -torch.norm(pt2d - pt_hm[:, None], dim=-1)
    if "pose" in mode:
        val, topk = canonical_topk(score.sum(-1), K, dim=1)
        bidx = torch.arange(bs)[:, None].repeat(1, K)
        quat = axis_angle_to_quaternion(pose16[bidx, topk]).permute(0, 2, 1, 3)
        fused = quaternion_to_axis_angle(average_quaternion(quat)).reshape(bs, -1)
        m, fv, fj = _finish(mano, fused, shape, bs)
        return {"topk": topk, "diff_topk_joint": joint[bidx, topk], "agg_hand_mano": m, "agg_vert": fv, "agg_joint": fj, "score": score}
    val, topk = canonical_topk(score, K, dim=1)
    b3 = torch.arange(bs)[:, None, None].repeat(1, K, J)
    j3 = torch.arange(J)[None, None].repeat(bs, K, 1)
    tj = joint[b3, topk, j3]
    return {"topk": topk, "diff_topk_joint": tj, "agg_hand_mano": torch.zeros(bs, 58), "agg_vert": torch.zeros(bs, 778, 3),
            "agg_joint": tj.mean(dim=1), "score": score}


def hand_average_all(mano, pose, shape, bs) -> Dict:
    """average_all (:379-424)"""
    quat = axis_angle_to_quaternion(pose.reshape(bs, -1, 16, 3)).permute(0, 2, 1, 3)
    fused = quaternion_to_axis_angle(average_quaternion(quat)).reshape(bs, -1)
    m, fv, fj = _finish(mano, fused, shape, bs)
    return {"agg_hand_mano": m, "agg_vert": fv, "agg_joint": fj}


def hand_random(mano, pose, shape, bs) -> Dict:
    """random (:426-467): candidate 0"""
    m, fv, fj = _finish(mano, pose.reshape(bs, -1, 48)[:, 0], shape, bs)
    return {"agg_hand_mano": m, "agg_vert": fv, "agg_joint": fj}


def _obj_verts(obj: OracleObject, fused, root_joint, obj_name, is_right):
    p = fused.clone()
    p[..., 6:] = p[..., 6:] + root_joint
    return obj.flip_pt3d(obj(p, obj_name, data_name="verts"), is_right)


def obj_heatmap(obj: OracleObject, pose6d, root_joint, obj_name, cam, heatmap, bbox, k, is_right) -> Dict:
    """ObjectAggregator.select_by_heatmap (:646-659)"""
    topk, weight, _ = obj_heat_topk(obj, pose6d, root_joint, obj_name, cam, heatmap, bbox, k, is_right)
    fused = obj_fuse_topk(topk, pose6d).float()
    return {"agg_6d": fused, "agg_obj_vert": _obj_verts(obj, fused, root_joint, obj_name, is_right), "topk": topk}


def obj_cascade_plain(obj: OracleObject, pose6d, root_joint, obj_name, cam, heatmap, bbox, k, is_right, is_weight) -> Dict:
    """ObjectAggregator.select_by_heatmap_cascade with is_force_selection=False (:661-722)"""
    ori = pose6d.clone()
    sel = lambda p: obj_heat_topk(obj, p, root_joint, obj_name, cam, heatmap, bbox, k, is_right)[:2]   # noqa: E731
    topk, w = sel(ori)
    trans1 = obj_fuse_topk(topk, ori, w if is_weight else None)[:, 6:]
    p = ori.clone()
    p[..., 6:] = ori[..., 6:] * 0 + trans1[:, None]
    topk, w = sel(p)
    rot1 = obj_fuse_topk(topk, p, w if is_weight else None)[:, :6]
    p = ori.clone()
    p[..., :6] = ori[..., :6] * 0 + rot1[:, None]
    topk_t2, _ = sel(p)
    p = ori.clone()
    p[..., 6:] = ori[..., 6:] * 0 + trans1[:, None]
    topk_r2, _ = sel(p)
    trans2 = obj_fuse_topk(topk_t2, p)[:, 6:]          # fused on the LAST pose set, as the reference does (:713)
    rot2 = obj_fuse_topk(topk_r2, p)[:, :6]
    fused = torch.cat([rot2, trans2], dim=-1).float()
    return {"agg_6d": fused, "pose6d_candidate": p, "agg_obj_vert": _obj_verts(obj, fused, root_joint, obj_name, is_right)}


def obj_2d_pt_pose(obj: OracleObject, pose6d, root_joint, obj_name, cam, heatmap, bbox, k, is_right) -> Dict:
    """ObjectAggregator.select_by_2D_pt, '2D_pt_pose' (:1001-1052)"""
    bs, J, H, W = heatmap.shape
    p = pose6d.clone().float()
    p[..., 6:] = p[..., 6:] + root_joint.unsqueeze(1)
    pts = obj.flip_pt3d(obj(p, obj_name), is_right)
    pt2d = project(pts, cam)
    bb = bbox[:, None, None, :]
    pt2d = 2 * (pt2d - bb[..., :2]) / (bb[..., 2:] - bb[..., :2]) - 1
    X, Y = torch.arange(W) / (W - 1) * 2 - 1, torch.arange(H) / (H - 1) * 2 - 1
    XX, YY = torch.meshgrid(X, Y, indexing="ij")
    XX, YY = XX[None, None].repeat(bs, J, 1, 1).reshape(bs, J, -1), YY[None, None].repeat(bs, J, 1, 1).reshape(bs, J, -1)
    ind = torch.argmax(heatmap.reshape(bs, J, -1), dim=-1)
    i1, i2 = torch.arange(bs)[:, None].repeat(1, J), torch.arange(J)[None].repeat(bs, 1)
    pt_hm = torch.stack([XX[i1, i2, ind], YY[i1, i2, ind]], dim=-1)
    score = (-torch.norm(pt2d - pt_hm[:, None], dim=-1)).sum(-1)
    _, topk = canonical_topk(score, k, dim=1)
    fused = obj_fuse_topk(topk, pose6d).float()
    return {"agg_6d": fused, "agg_obj_vert": _obj_verts(obj, fused, root_joint, obj_name, is_right), "topk": topk, "score": score}


def obj_average_first_k(obj: OracleObject, pose6d, root_joint, obj_name, k, is_right) -> Dict:
    """ObjectAggregator.average_all (:1054-1082): fuse_topk over the first k candidates"""
    bs = pose6d.shape[0]
    fused = obj_fuse_topk(torch.arange(k)[None].repeat(bs, 1), pose6d).float()
    return {"agg_6d": fused, "agg_obj_vert": _obj_verts(obj, fused, root_joint, obj_name, is_right)}


def obj_random(obj: OracleObject, pose6d, root_joint, obj_name, is_right) -> Dict:
    """ObjectAggregator.random (:1084-1112): fuse_topk over candidate 0"""
    bs = pose6d.shape[0]
    fused = obj_fuse_topk(torch.zeros(bs, 1, dtype=torch.long), pose6d).float()
    return {"agg_6d": fused, "agg_obj_vert": _obj_verts(obj, fused, root_joint, obj_name, is_right)}
