"""TEST INFRASTRUCTURE -- seeded input builders shared by oracle/make_golden.py and the tests, so the reference run that
minted tests/golden/*.npz and the later oracle / CUDA runs consume identical inputs."""
from __future__ import annotations

import numpy as np
import torch

from vpho_b200 import synthetic as syn

SIGMA_T0 = 0.01 * (50 / 0.01) ** 0.65


def assets():
    mano = syn.make_mano_model()
    return mano, syn.make_anchor_assets(mano), syn.make_object_tables()


def sampler_case(head: str, bs: int, S: int, last_std: float, seed: int):
    st = syn.make_denoiser_state(head, seed, last_std=last_std)
    g = torch.Generator().manual_seed(100 + seed)
    enc = torch.relu(torch.randn(bs, 1024, generator=g))
    D = 96 if head == "mano_pose" else 9
    init = torch.randn(bs * S, D, generator=g) * SIGMA_T0
    return st, enc, init


def aggregate_case(bs: int, S: int, seed: int, spread: float = 0.25):
    """Clustered hand / object candidates around a hidden pose (what a trained sampler produces) + the synthetic batch."""
    mano, anch, objs = assets()
    batch = syn.make_eval_batch(bs, seed=seed, sample_num=S, mano=mano, objects=objs)
    T = lambda k: torch.from_numpy(np.asarray(batch[k]))  # noqa: E731
    g = torch.Generator().manual_seed(300 + seed)
    true_pose = torch.cat([T("true_wrist"), torch.randn(bs, 45, generator=g) * 0.2], 1)
    pose_diff = (true_pose[:, None] + torch.randn(bs, S, 48, generator=g) * spread).reshape(-1, 48).float()
    shape = T("pd_mano_shape")[:, None].repeat(1, S, 1).reshape(-1, 10)
    rot = torch.randn(bs, S, 6, generator=g, dtype=torch.float64)
    tr = T("true_obj_rot").double()[:, None, :2, :].reshape(bs, 1, 6)
    rot[:, ::2] = tr + 0.15 * rot[:, ::2]
    tt = T("true_obj_trans").double()[:, None] + 0.02 * torch.randn(bs, S, 3, generator=g, dtype=torch.float64)
    kw = dict(cam_intrinsic=T("cam_intr_crop_flip"), root_joint_flip=T("root_joint_flip"), root_joint=T("root_joint"),
              is_right=T("is_right"), force_local=T("force_local"), is_grasped=T("is_grasped"),
              hand_pose_diff=pose_diff, hand_pose_regression=T("pd_mano_pose"), hand_shape=shape,
              hand_heatmap=T("hm_hand"), hand_bbox=T("bbox_hand"), obj_pose6d=torch.cat([rot, tt], -1),
              obj_heatmap=T("hm_obj"), obj_bbox=T("bbox_obj_rect"), obj_name=list(batch["obj_name"]))
    return kw, batch, true_pose


def clone_kw(kw: dict) -> dict:
    return {k: (v.clone() if isinstance(v, torch.Tensor) else (list(v) if isinstance(v, list) else v)) for k, v in kw.items()}


def fingerprint(*tensors) -> float:
    """Cheap order-sensitive checksum used to detect drift of the seeded generators between machines."""
    acc = 0.0
    for t in tensors:
        a = np.asarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float64).reshape(-1)
        w = np.cos(np.arange(a.size, dtype=np.float64) * 0.37)
        acc += float((a * w).sum())
    return acc


def e2e_priors(kind: str, bs: int, S: int, batch: dict, seed: int):
    """Prior tensors for the end-to-end tests: 'random' = N(0, sigma(T0)^2) as in the README run; 'clustered' = tight
    clusters of valid 6D poses around a hidden pose (what a trained sampler converges to)."""
    from pytorch3d.transforms.rotation_conversions import axis_angle_to_matrix, matrix_to_rotation_6d   # oracle/shims
    g = torch.Generator().manual_seed(seed)
    if kind == "random":
        return torch.randn(bs * S, 96, generator=g) * SIGMA_T0, torch.randn(bs * S, 9, generator=g) * SIGMA_T0
    T = lambda k: torch.from_numpy(np.asarray(batch[k]))  # noqa: E731
    true_pose = torch.cat([T("true_wrist"), torch.randn(bs, 45, generator=g) * 0.2], 1)
    aa = (true_pose[:, None] + torch.randn(bs, S, 48, generator=g) * 0.2).reshape(bs * S, 16, 3)
    ph = matrix_to_rotation_6d(axis_angle_to_matrix(aa)).reshape(bs * S, 96) * (1 + 0.05 * torch.randn(bs * S, 1, generator=g))
    rot = T("true_obj_rot")[:, None, :2, :].reshape(bs, 1, 6) + 0.1 * torch.randn(bs, S, 6, generator=g)
    tr = T("true_obj_trans")[:, None] + 0.02 * torch.randn(bs, S, 3, generator=g)
    return ph.float().contiguous(), torch.cat([rot, tr], -1).reshape(bs * S, 9).float().contiguous()


def object_metric_case(n: int = 6, C: int = 5, seed: int = 11):
    """Seeded inputs of the object-metric fixture: ground-truth poses in front of the camera, candidates scattered around
    them from near-exact to far off (so every F-score threshold and both 0.1-diameter flags see both outcomes)."""
    from scipy.spatial.transform import Rotation as Rot
    rng = np.random.default_rng(seed)
    R_gt = Rot.random(n, random_state=int(rng.integers(1 << 30))).as_matrix()
    t_gt = rng.normal(size=(n, 3)) * 0.05 + np.array([0.02, -0.03, 0.6])
    gt = np.concatenate([R_gt, t_gt[..., None]], -1)
    spread = np.array([1e-3, 0.02, 0.1, 0.4, 1.5])[:C]
    d = Rot.from_rotvec((rng.normal(size=(n, C, 3)) * spread[None, :, None]).reshape(-1, 3)).as_matrix().reshape(n, C, 3, 3)
    tt = t_gt[:, None] + rng.normal(size=(n, C, 3)) * (spread[None, :, None] * 0.05)
    pd = np.concatenate([d @ R_gt[:, None], tt[..., None]], -1)
    ids = np.array([0, 1, 2, 5, 20, 8][:n], np.int32)
    K = np.tile(np.array([[600.0, 0, 128], [0, 610.0, 126], [0, 0, 1]], np.float32), (n, 1, 1))
    return {"pd_rt": pd, "gt_rt": gt, "obj_id": ids, "cam_intr": K}
