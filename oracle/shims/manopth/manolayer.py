"""TEST INFRASTRUCTURE (oracle shim) -- restatement of hassony2/manopth `ManoLayer.forward`
(use_pca=False, flat_hand_mean=True, center_idx=0, side='right' -- the configuration fixed at
/root/reference lib/model/head_mano.py:48-55).  manopth is a third-party dependency of the reference that is
neither vendored under /root/reference nor pinned in environment.yaml; the arithmetic below follows the
published upstream functions `manolayer.ManoLayer.forward`, `rodrigues_layer.batch_rodrigues / quat2mat`,
`tensutils.th_posemap_axisang / th_with_zeros / subtract_flat_id` as frozen in SURVEY.md §3.2 and A.9.

The licensed MANO_RIGHT.pkl is not available offline, so the model tensors are injected through
`set_model(dict)` before the layer is constructed (same shapes as the pickle's fields).
"""
import numpy as np
import torch
from torch.nn import Module

_MODEL = None

TIP_IDS_RIGHT = [745, 317, 444, 556, 673]
JOINT_REORDER_21 = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]
CHAIN_REORDER_16 = [0, 1, 6, 11, 2, 7, 12, 3, 8, 13, 4, 9, 14, 5, 10, 15]


def set_model(model: dict) -> None:
    """model: v_template (778,3), shapedirs (778,3,10), posedirs (778,3,135), J_regressor (16,778), weights (778,16)."""
    global _MODEL
    _MODEL = {k: np.asarray(v) for k, v in model.items()}


def quat2mat(quat):
    norm_quat = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = norm_quat[:, 0], norm_quat[:, 1], norm_quat[:, 2], norm_quat[:, 3]
    batch_size = quat.size(0)
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz = w * x, w * y, w * z
    xy, xz, yz = x * y, x * z, y * z
    rot = torch.stack(
        [w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
         2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
         2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(batch_size, 3, 3)
    return rot


def batch_rodrigues(axisang):
    axisang_norm = torch.norm(axisang + 1e-8, p=2, dim=1)
    angle = torch.unsqueeze(axisang_norm, -1)
    axisang_normalized = torch.div(axisang, angle)
    angle = angle * 0.5
    v_cos = torch.cos(angle)
    v_sin = torch.sin(angle)
    quat = torch.cat([v_cos, v_sin * axisang_normalized], dim=1)
    rot_mat = quat2mat(quat)
    return rot_mat.view(rot_mat.shape[0], 9)


def th_with_zeros(tensor):
    batch_size = tensor.shape[0]
    padding = tensor.new([0.0, 0.0, 0.0, 1.0])
    padding.requires_grad = False
    return torch.cat([tensor, padding.view(1, 1, 4).repeat(batch_size, 1, 1)], 1)


class ManoLayer(Module):
    def __init__(self, center_idx=None, flat_hand_mean=True, ncomps=6, side="right", mano_root="",
                 use_pca=True, root_rot_mode="axisang", joint_rot_mode="axisang", robust_rot=False, **kw):
        super().__init__()
        if _MODEL is None:
            raise RuntimeError("oracle manopth shim: call set_model(...) before constructing ManoLayer")
        assert not use_pca and flat_hand_mean and side == "right" and ncomps == 45
        self.center_idx = center_idx
        self.side = side
        m = _MODEL
        self.register_buffer("th_v_template", torch.from_numpy(m["v_template"]).float().unsqueeze(0))
        self.register_buffer("th_shapedirs", torch.from_numpy(m["shapedirs"]).float())
        self.register_buffer("th_posedirs", torch.from_numpy(m["posedirs"]).float())
        self.register_buffer("th_J_regressor", torch.from_numpy(m["J_regressor"]).float())
        self.register_buffer("th_weights", torch.from_numpy(m["weights"]).float())
        self.register_buffer("th_hands_mean", torch.zeros(1, 45))

    def forward(self, th_pose_coeffs, th_betas=torch.zeros(1), th_trans=torch.zeros(1)):
        batch_size = th_pose_coeffs.shape[0]
        th_full_pose = torch.cat([th_pose_coeffs[:, :3], self.th_hands_mean + th_pose_coeffs[:, 3:48]], 1)
        # th_posemap_axisang
        rot_mats = batch_rodrigues(th_full_pose.contiguous().view(-1, 3)).view(batch_size, 16 * 9)
        th_rot_map = rot_mats[:, 9:]
        # subtract_flat_id
        id_flat = torch.eye(3, dtype=rot_mats.dtype, device=rot_mats.device).view(1, 9).repeat(batch_size, 15)
        th_pose_map = th_rot_map - id_flat
        root_rot = rot_mats[:, :9].view(batch_size, 3, 3)

        th_v_shaped = torch.matmul(self.th_shapedirs, th_betas.transpose(1, 0)).permute(2, 0, 1) + self.th_v_template
        th_j = torch.matmul(self.th_J_regressor, th_v_shaped)
        th_v_posed = th_v_shaped + torch.matmul(self.th_posedirs, th_pose_map.transpose(0, 1)).permute(2, 0, 1)

        root_j = th_j[:, 0, :].contiguous().view(batch_size, 3, 1)
        root_trans = th_with_zeros(torch.cat([root_rot, root_j], 2))
        all_rots = th_rot_map.view(th_rot_map.shape[0], 15, 3, 3)
        lev1_idxs, lev2_idxs, lev3_idxs = [1, 4, 7, 10, 13], [2, 5, 8, 11, 14], [3, 6, 9, 12, 15]
        lev1_rots = all_rots[:, [idx - 1 for idx in lev1_idxs]]
        lev2_rots = all_rots[:, [idx - 1 for idx in lev2_idxs]]
        lev3_rots = all_rots[:, [idx - 1 for idx in lev3_idxs]]
        lev1_j, lev2_j, lev3_j = th_j[:, lev1_idxs], th_j[:, lev2_idxs], th_j[:, lev3_idxs]

        all_transforms = [root_trans.unsqueeze(1)]
        lev1_j_rel = lev1_j - root_j.transpose(1, 2)
        lev1_rel = th_with_zeros(torch.cat([lev1_rots, lev1_j_rel.unsqueeze(3)], 3).view(-1, 3, 4))
        root_trans_flt = root_trans.unsqueeze(1).repeat(1, 5, 1, 1).view(root_trans.shape[0] * 5, 4, 4)
        lev1_flt = torch.matmul(root_trans_flt, lev1_rel)
        all_transforms.append(lev1_flt.view(all_rots.shape[0], 5, 4, 4))

        lev2_j_rel = lev2_j - lev1_j
        lev2_rel = th_with_zeros(torch.cat([lev2_rots, lev2_j_rel.unsqueeze(3)], 3).view(-1, 3, 4))
        lev2_flt = torch.matmul(lev1_flt, lev2_rel)
        all_transforms.append(lev2_flt.view(all_rots.shape[0], 5, 4, 4))

        lev3_j_rel = lev3_j - lev2_j
        lev3_rel = th_with_zeros(torch.cat([lev3_rots, lev3_j_rel.unsqueeze(3)], 3).view(-1, 3, 4))
        lev3_flt = torch.matmul(lev2_flt, lev3_rel)
        all_transforms.append(lev3_flt.view(all_rots.shape[0], 5, 4, 4))

        th_results = torch.cat(all_transforms, 1)[:, CHAIN_REORDER_16]
        th_results_global = th_results

        joint_js = torch.cat([th_j, th_j.new_zeros(th_j.shape[0], 16, 1)], 2)
        tmp2 = torch.matmul(th_results, joint_js.unsqueeze(3))
        th_results2 = (th_results - torch.cat([tmp2.new_zeros(*tmp2.shape[:2], 4, 3), tmp2], 3)).permute(0, 2, 3, 1)

        th_T = torch.matmul(th_results2, self.th_weights.transpose(0, 1))
        th_rest_shape_h = torch.cat(
            [th_v_posed.transpose(2, 1),
             torch.ones((batch_size, 1, th_v_posed.shape[1]), dtype=th_T.dtype, device=th_T.device)], 1)
        th_verts = (th_T * th_rest_shape_h.unsqueeze(1)).sum(2).transpose(2, 1)
        th_verts = th_verts[:, :, :3]
        th_jtr = th_results_global[:, :, :3, 3]
        tips = th_verts[:, TIP_IDS_RIGHT]
        th_jtr = torch.cat([th_jtr, tips], 1)
        th_jtr = th_jtr[:, JOINT_REORDER_21]

        if th_trans is None or bool(torch.norm(th_trans) == 0):
            if self.center_idx is not None:
                center_joint = th_jtr[:, self.center_idx].unsqueeze(1)
                th_jtr = th_jtr - center_joint
                th_verts = th_verts - center_joint
        else:
            th_jtr = th_jtr + th_trans.unsqueeze(1)
            th_verts = th_verts + th_trans.unsqueeze(1)
        # manopth scales to millimetres
        th_verts = th_verts * 1000
        th_jtr = th_jtr * 1000
        return th_verts, th_jtr
