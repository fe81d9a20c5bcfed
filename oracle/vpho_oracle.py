"""TEST INFRASTRUCTURE -- CPU oracle for the VPHO evaluation hot path.

A plain-PyTorch/NumPy/SciPy restatement of what the reference executes between lib/model/VPHO.py:236 and
:304 (sample -> MANO -> visual/physical scoring -> top-k -> aggregation).  It exists ONLY to check the
CUDA path: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it; the product package `vpho_b200/` never does.

Parity status: the reference ships no golden vectors for this path, so the oracle is pinned
against outputs of the reference's own files executed in the build container
(`oracle/make_golden.py` -> tests/golden/*.npz, checked by tests/test_oracle_golden.py, and live by
tests/test_oracle_vs_reference.py when /root/reference exists).

Third-party arithmetic that is not under /root/reference and is restated in `oracle/shims/`:
manopth ManoLayer (unpinned), pytorch3d rotation conversions (unpinned; 0.7.8 semantics frozen).
`scipy.integrate.solve_ivp(method='RK45')` (reference pins scipy==1.12.0; this image has 1.18.1) and
`torch.nn.functional.grid_sample`, `torch.topk`, `torch.linalg.eigh` are called directly, as the reference does.

Declared deviations from the unmodified reference (SURVEY.md §8c "oracle rules"):
  (0) `torch.cdist` runs with compute_mode='donot_use_mm_for_euclid_dist' (Appendix A.4);
  (ii) top-k is canonicalised to (value desc, index asc) by a stable sort;
  numpy>=2 promotion applies to `0.5*g^2*score` (float64 product).
"""
from __future__ import annotations

import os
import sys
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F
from scipy import integrate

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
if _SHIMS not in sys.path:
    sys.path.insert(0, _SHIMS)

from pytorch3d.transforms.rotation_conversions import (  # noqa: E402  (oracle/shims)
    axis_angle_to_quaternion, matrix_to_axis_angle, matrix_to_quaternion, matrix_to_rotation_6d,
    quaternion_to_axis_angle, quaternion_to_matrix, rotation_6d_to_matrix,
)
import manopth.manolayer as _mano_shim  # noqa: E402

# ---------------------------------------------------------------------------------------------------------
# index tables (lib/utils/hand_fn.py:240-274, lib/model/aggregation.py:584-590, lib/utils/physics_fn.py:124-169)
# ---------------------------------------------------------------------------------------------------------
MANO_PARAMS_LEVEL = {
    0: [0, 1, 2],
    1: [39, 40, 41, 3, 4, 5, 12, 13, 14, 30, 31, 32, 21, 22, 23],
    2: [42, 43, 44, 6, 7, 8, 15, 16, 17, 33, 34, 35, 24, 25, 26],
    3: [45, 46, 47, 9, 10, 11, 18, 19, 20, 36, 37, 38, 27, 28, 29],
}
MANO_JOINT_LEVEL = {0: [0], 1: [1, 5, 9, 13, 17], 2: [2, 6, 10, 14, 18], 3: [3, 7, 11, 15, 19], 4: [4, 8, 12, 16, 20]}
SKELETON_LEVEL = {
    0: [[0, 1], [0, 5], [0, 9], [0, 13], [0, 17]],
    1: [[1, 2], [5, 6], [9, 10], [13, 14], [17, 18]],
    2: [[2, 3], [6, 7], [10, 11], [14, 15], [18, 19]],
    3: [[3, 4], [7, 8], [11, 12], [15, 16], [19, 20]],
}
FINGER_FORCE_LEVEL = [[1, 2, 3, 4], [8, 9, 10, 11], [14, 15, 16, 17], [21, 22, 23, 24], [28, 29, 30, 31]]


def anchor_skeleton_table() -> np.ndarray:
    """(32,2) joint pairs giving each force anchor's bone direction (lib/utils/physics_fn.py:127-169)."""
    label = [5, 12, 19, 18, 26, 25, 6, 0, 7, 13, 20, 27, 1, 8, 14, 21, 28,
             2, 3, 4, 9, 11, 10, 15, 17, 16, 22, 24, 23, 29, 31, 30]
    S = SKELETON_LEVEL
    sk = [S[0][1], S[0][2], S[0][3], S[0][3], S[0][4], S[0][4],
          S[0][0], S[0][0], S[1][1], S[1][2], S[1][3], S[1][4],
          S[2][0], S[2][1], S[2][2], S[2][3], S[2][4],
          S[3][0], S[3][0], S[3][0], S[3][1], S[3][1], S[3][1], S[3][2], S[3][2], S[3][2],
          S[3][3], S[3][3], S[3][3], S[3][4], S[3][4], S[3][4]]
    sk = np.asarray(sk)
    return sk[np.argsort(np.asarray(label))]


# Geometry dtype of the scoring / aggregation stage.  The reference casts the (float64) object poses to float32 before any
# geometry (`pose6d.clone().float()`, aggregation.py:753,958,1282).  `float64_shadow()` switches those casts -- and the
# MANO / object / anchor tables -- to float64: the SAME algorithm evaluated without FP32 rounding.  It is used only to
# measure how far the reference's own FP32 result sits from the exact one (tests/test_headline_parity.py), i.e. to
# derive the parity tolerance of ill-conditioned cases instead of choosing it.
_GEOM = [torch.float32]


def _geom(x: torch.Tensor) -> torch.Tensor:
    return x.to(_GEOM[0])


class float64_shadow:
    def __enter__(self):
        _GEOM.append(torch.float64)
        _GEOM[0] = torch.float64
        return self

    def __exit__(self, *exc):
        _GEOM.pop()
        _GEOM[0] = torch.float32
        return False


def canonical_topk(x: torch.Tensor, k: int, dim: int = 1):
    """torch.topk with the canonical tie-break (value desc, index asc) -- oracle rule (ii)."""
    order = torch.sort(-x, dim=dim, stable=True)[1]
    idx = order.narrow(dim, 0, k)
    return torch.gather(x, dim, idx), idx


# ---------------------------------------------------------------------------------------------------------
# S1/S2: VE-SDE, score network, probability-flow ODE sampler
# ---------------------------------------------------------------------------------------------------------
SIGMA_MIN, SIGMA_MAX, SAMPLING_EPS = 0.01, 50, 1e-5   # lib/model/sde.py:90-97


def ve_marginal_std(t):
    """lib/model/sde.py:15-18."""
    return SIGMA_MIN * (SIGMA_MAX / SIGMA_MIN) ** t


def ve_sde(t: torch.Tensor):
    """lib/model/sde.py:20-24 -- returns (drift, diffusion)."""
    sigma = SIGMA_MIN * (SIGMA_MAX / SIGMA_MIN) ** t
    drift = torch.tensor(0)
    diffusion = sigma * torch.sqrt(torch.tensor(2 * (np.log(SIGMA_MAX) - np.log(SIGMA_MIN)), device=t.device))
    return drift, diffusion


class OracleDenoiser:
    """BaseDenoiser.forward (lib/model/denoiser.py:68-82) with GaussianFourierProjection (:19-31),
    ManoPoseHead2 / ObjHead2 (:166-179, :234-247) and ParallelLinear (lib/model/parallel_linear.py:27-35)."""

    def __init__(self, state: Dict[str, np.ndarray]):
        self.p = {k: torch.as_tensor(np.asarray(v)).float() for k, v in state.items()}
        self.n = self.p["head.head.0.weight"].shape[0]
        self.out_dim = 3 * self.n
        self.calls = 0

    def __call__(self, data: dict) -> torch.Tensor:
        p = self.p
        self.calls += 1
        feat, t, x = data["feat"], data["t"], data["sampled_pose"]
        tt = t.squeeze(1)
        x_proj = tt[:, None] * p["t_encoder.0.W"][None, :] * 2 * np.pi
        four = torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)
        t_feat = torch.relu(F.linear(four, p["t_encoder.1.weight"], p["t_encoder.1.bias"]))
        h = torch.relu(F.linear(x, p["pose_encoder.0.weight"], p["pose_encoder.0.bias"]))
        pose_feat = torch.relu(F.linear(h, p["pose_encoder.2.weight"], p["pose_encoder.2.bias"]))
        total = torch.cat([t_feat, pose_feat, feat], dim=-1)
        y = torch.einsum("bc,ncd->bnd", total, p["head.head.0.weight"]) + p["head.head.0.bias"]
        y = torch.relu(y)
        y = torch.einsum("bnc,ncd->bnd", y, p["head.head.2.weight"]) + p["head.head.2.bias"]
        out = y.reshape(-1, self.out_dim)
        std = ve_marginal_std(t)
        return out / (std + 1e-7)


def oracle_sample(den: OracleDenoiser, feat: torch.Tensor, T0: float, init_x: torch.Tensor, num_steps: int = 50,
                  rtol: float = 3e-3, atol: float = 3e-4, eps: float = SAMPLING_EPS, denoise: bool = True):
    """cond_ode_sampler (lib/model/score_based_model.py:45-105) with the prior draw passed in as `init_x`
    (= randn*sigma(T0), lib/model/sde.py:26-28) so both sides integrate from the same seeded tensor.
    Returns xs (N,num_steps,D) f64, x (N,D) f64, info dict."""
    data = {"feat": feat}
    batch_size, pose_dim = init_x.shape
    shape = init_x.shape
    nan_seen = [False]

    def score_eval(d):
        score = den(d)
        if torch.any(torch.isnan(score)):
            nan_seen[0] = True
            score = torch.nan_to_num_(score, nan=0.0, posinf=0.0, neginf=0.0)
        return score.numpy().reshape((-1,))

    def ode_func(t, x):
        x = torch.tensor(x.reshape(-1, pose_dim)).float()
        time_steps = torch.ones(batch_size).unsqueeze(-1) * t
        drift, diffusion = ve_sde(torch.tensor(t))
        drift = drift.numpy()
        diffusion = diffusion.numpy()
        data["sampled_pose"] = x
        data["t"] = time_steps
        return drift - 0.5 * (diffusion ** 2) * score_eval(data)

    t_eval = np.linspace(T0, eps, num_steps)
    calls0 = den.calls
    with torch.no_grad():
        res = integrate.solve_ivp(ode_func, (T0, eps), init_x.reshape(-1).numpy(), rtol=rtol, atol=atol,
                                  method="RK45", t_eval=t_eval, max_step=10)
        xs = torch.tensor(res.y).T.view(-1, batch_size, pose_dim)
        x = torch.tensor(res.y[:, -1]).reshape(shape)
        if denoise:
            vec_eps = torch.ones((x.shape[0], 1)) * eps
            drift, diffusion = ve_sde(vec_eps)
            data["sampled_pose"] = x.float()
            data["t"] = vec_eps
            grad = den(data)
            drift = drift - diffusion ** 2 * grad
            x = x + drift * ((1 - eps) / num_steps)
    info = {"nfev": int(res.nfev), "net_calls": den.calls - calls0, "status": int(res.status), "nan": nan_seen[0]}
    return xs.permute(1, 0, 2), x, info


# ---------------------------------------------------------------------------------------------------------
# M1/M2: MANO layer and the 6D -> axis-angle post-processing
# ---------------------------------------------------------------------------------------------------------
class OracleMano:
    """HeadMano.get_hand_verts (lib/model/head_mano.py:78-87) over the manopth restatement (oracle/shims)."""

    def __init__(self, model: Dict[str, np.ndarray], dtype=torch.float32):
        _mano_shim.set_model(model)
        self.layer = _mano_shim.ManoLayer(ncomps=45, center_idx=0, flat_hand_mean=True, side="right",
                                          mano_root="", use_pca=False)
        if dtype != torch.float32:          # float64 shadow: the model tensors themselves stay the float32 values
            self.layer = self.layer.to(dtype)
        self.calls = 0

    def __call__(self, pose: torch.Tensor, shape: torch.Tensor):
        self.calls += int(pose.shape[0])
        verts, joints = self.layer(th_pose_coeffs=pose, th_betas=shape)
        return verts / 1000, joints / 1000


def postprocess_diffusion_hand(hand_final: torch.Tensor, pd_mano_shape: torch.Tensor, sample_num: int,
                               hand_inprocess: Optional[torch.Tensor] = None):
    """vpho_net.postprocess_diffusion_hand, branch 'mano_pose' (lib/model/VPHO.py:306-331)."""
    bs = pd_mano_shape.shape[0]
    hf = hand_final.reshape(bs, sample_num, 16, 6)
    hf = matrix_to_axis_angle(rotation_6d_to_matrix(hf)).reshape(bs, sample_num, 48)
    hf = torch.cat((hf, pd_mano_shape[:, None].repeat(1, sample_num, 1)), dim=-1).reshape(-1, 58)
    if hand_inprocess is None:
        return None, hf
    n_in = hand_inprocess.shape[1]
    hi = hand_inprocess.reshape(-1, sample_num, n_in, 16, 6)
    hi = matrix_to_axis_angle(rotation_6d_to_matrix(hi)).reshape(bs, sample_num, n_in, 48)
    hi = torch.cat((hi, pd_mano_shape[:, None, None].repeat(1, sample_num, n_in, 1)), dim=-1).reshape(-1, n_in, 58)
    return hi, hf


# ---------------------------------------------------------------------------------------------------------
# O1: object point layer
# ---------------------------------------------------------------------------------------------------------
class OracleObject:
    """HeadObject.forward / flip_pt3d (lib/model/head_object.py:36-67)."""

    def __init__(self, tables: Dict[str, object], dtype=torch.float32):
        self.names = list(tables["names"])
        self.tab = {
            "keypoint": torch.as_tensor(np.asarray(tables["kpt3d"])).float().to(dtype),
            "verts": torch.as_tensor(np.asarray(tables["verts_sampled"])).float().to(dtype),
            "CoM": torch.as_tensor(np.asarray(tables["CoM"])).float()[:, None].to(dtype),
        }

    def __call__(self, pose: torch.Tensor, name: Sequence[str], data_name: str = "keypoint"):
        ids = [self.names.index(n) for n in name]
        pts = self.tab[data_name][ids]
        rotmat = rotation_6d_to_matrix(pose[..., :6])
        new_verts = torch.einsum("bvi,b...ji->b...vj", pts, rotmat)
        return new_verts + pose[..., 6:].unsqueeze(-2)

    @staticmethod
    def flip_pt3d(pt3d: torch.Tensor, is_right: torch.Tensor):
        idx = torch.arange(pt3d.shape[0])[~is_right]
        pt3d[idx, ..., 0] = pt3d[idx, ..., 0] * -1
        return pt3d


# ---------------------------------------------------------------------------------------------------------
# P1: force anchors, local -> global forces
# ---------------------------------------------------------------------------------------------------------
class OracleAnchors:
    """ForceAnchor.__call__ + Vert2Joint (lib/utils/physics_fn.py:224-257, lib/utils/hand_fn.py:427-448) and
    from_local_to_global (lib/model/physics.py:362-371)."""

    def __init__(self, anchors: Dict[str, np.ndarray], dtype=torch.float32):
        self.face = torch.as_tensor(np.asarray(anchors["face_vertex_idx"]).reshape(-1)).long()
        aw = np.asarray(anchors["anchor_weight"], np.float64)
        aw = np.concatenate([np.ones((aw.shape[0], 1)), aw], axis=1)
        self.anchor_weight = torch.from_numpy(aw).float().to(dtype)
        self.vert2joint = torch.as_tensor(np.asarray(anchors["vert2joint"])).float().to(dtype)
        self.skel = torch.as_tensor(anchor_skeleton_table()).long()

    def points_and_frames(self, vertices: torch.Tensor):
        iv = vertices[..., self.face, :].reshape(vertices.shape[:-2] + (-1, 3, 3))
        b1 = iv[..., :, 1, :] - iv[..., :, 0, :]
        b2 = iv[..., :, 2, :] - iv[..., :, 0, :]
        joints = torch.einsum("...ij,ki->...kj", vertices, self.vert2joint)
        dy = joints[..., self.skel[:, 1], :] - joints[..., self.skel[:, 0], :]
        dz = torch.cross(b1, b2, dim=-1)
        dz = dz / (torch.norm(dz, dim=-1, keepdim=True) + 1e-8)
        dy = dy / (torch.norm(dy, dim=-1, keepdim=True) + 1e-8)
        dx = torch.cross(dy, dz, dim=-1)
        dy = torch.cross(dz, dx, dim=-1)
        dy = dy / (torch.norm(dy, dim=-1, keepdim=True) + 1e-8)
        frame = torch.stack([dx, dy, dz], dim=-1)
        w = self.anchor_weight
        pts = w[:, 1:2] * b1 + w[:, 2:3] * b2 + iv[..., :, 0, :]
        return pts, frame

    def from_local_to_global(self, force_local: torch.Tensor, hand_vert: torch.Tensor):
        pts, frame = self.points_and_frames(hand_vert)
        force_global = torch.einsum("...bi,...bji->...bj", force_local, frame)
        return pts, force_global


# ---------------------------------------------------------------------------------------------------------
# V1/V2/T1/T2: projection, bicubic heat sampling, quaternion averaging
# ---------------------------------------------------------------------------------------------------------
def project(pt3d_cam: torch.Tensor, cam_intrinsic: torch.Tensor):
    """project_point_by_cam_intrinsic (lib/model/aggregation.py:24-32)."""
    j2 = torch.einsum("b...ij,blj->b...il", pt3d_cam, cam_intrinsic)
    return j2[..., :2] / j2[..., 2:]


def sample_heat(heatmap: torch.Tensor, pt2d_norm: torch.Tensor, observe: Sequence[int]):
    """Heat sampling loops (lib/model/aggregation.py:206-213, 767-775): bicubic grid_sample per observed map."""
    vals = []
    for i in observe:
        g = pt2d_norm[:, :, [i]]
        v = F.grid_sample(heatmap[:, [i]], g, align_corners=False, mode="bicubic").squeeze(1)
        vals.append(v)
    return torch.concat(vals, dim=-1)


def average_quaternion(Q: torch.Tensor, W: Optional[torch.Tensor] = None):
    """lib/utils/transform_fn.py:101-125."""
    shape = Q.shape
    if W is None:
        W = torch.ones_like(Q[..., 0])
    weight_sum = W.sum(dim=-1, keepdim=True)
    oq = ((Q[..., 0:1] > 0).float() - 0.5) * 2 * Q
    A = torch.einsum("...ni,...nj->...nij", oq, oq)
    A = torch.sum(torch.einsum("...nij,...n->...nij", A, W), -3)
    A = A / weight_sum.reshape(*shape[:-2], 1, 1)
    q = torch.linalg.eigh(A)[1][..., -1]
    return ((q[..., 0:1] > 0).float() - 0.5) * 2 * q


def average_rot6d(rot6d: torch.Tensor, weights: Optional[torch.Tensor] = None):
    """lib/model/aggregation.py:50-56."""
    if weights is None:
        weights = torch.ones_like(rot6d[..., 0]) / rot6d.shape[-2]
    quat = matrix_to_quaternion(rotation_6d_to_matrix(rot6d))
    qm = average_quaternion(quat, weights)
    return matrix_to_rotation_6d(quaternion_to_matrix(qm))


def exact_cdist(a, b):
    return torch.cdist(a, b, p=2, compute_mode="donot_use_mm_for_euclid_dist")


# ---------------------------------------------------------------------------------------------------------
# A1: hand heat-map cascade
# ---------------------------------------------------------------------------------------------------------
def _hand_level(mano: OracleMano, pose, shape, root_joint, cam, heatmap, bbox, K, fuse_index, observe_index,
                independent: bool):
    """select_topk_hand_by_observed_heatmap_and_fuse_by_index (lib/model/aggregation.py:180-284), is_weight=True."""
    bs = heatmap.shape[0]
    vert, joint = mano(pose, shape)
    joint = joint.reshape(bs, -1, 21, 3)
    joint_cam = joint + root_joint[:, None, None]
    pt2d = project(joint_cam, cam)
    bb = bbox[:, None, None, :]
    pt2d = pt2d - bb[..., :2]
    pt2d = 2 * pt2d / (bb[..., 2:] - bb[..., :2]) - 1
    heat_val = sample_heat(heatmap, pt2d, observe_index)
    pose3 = pose.reshape(bs, -1, 48)
    out = {}
    if not independent:
        score = heat_val.sum(dim=-1)
        val, topk = canonical_topk(score, K, dim=1)
        weight = (val + 1e-8) / (val.sum(dim=1, keepdim=True) + 1e-8)
        bidx = torch.arange(bs)[:, None].repeat(1, K)
        tp = pose3[bidx, topk][:, :, fuse_index].reshape(bs, K, -1, 3)
        quat = axis_angle_to_quaternion(tp).permute(0, 2, 1, 3)
        fq = average_quaternion(quat, weight[:, None])
        faa = quaternion_to_axis_angle(fq).reshape(bs, -1)
        topk_aa = tp
    else:
        M, N = len(observe_index), len(fuse_index)
        n_obs = M // (N // 3)
        score = heat_val.reshape(bs, -1, n_obs, N // 3).mean(dim=-2)
        val, topk = canonical_topk(score, K, dim=1)
        weight = (val + 1e-8) / (val.sum(dim=1, keepdim=True) + 1e-8)
        weight = weight.permute(0, 2, 1)
        pose16 = pose.reshape(bs, -1, 16, 3)
        i1 = torch.arange(bs)[:, None, None].repeat(1, K, N // 3)
        i2 = torch.tensor(fuse_index, dtype=torch.long)[None, None].repeat(bs, K, 1).reshape(bs, K, -1, 3)[:, :, :, 0] // 3
        tp = pose16[i1, topk, i2]
        quat = axis_angle_to_quaternion(tp).permute(0, 2, 1, 3)
        fq = average_quaternion(quat, weight)
        faa = quaternion_to_axis_angle(fq).reshape(bs, -1)
        topk_aa = tp
    fused = pose3.clone()
    fused[:, :, fuse_index] = fused[:, :, fuse_index] * 0 + faa[:, None]
    out.update(score=score, val=val, topk=topk, weight=weight, fused_idx_pose=faa, topk_idx_pose_aa=topk_aa,
               fused_pose=fused, joint=joint)
    return out


def hand_cascade(mano: OracleMano, pose_diff, pose_regression, shape, root_joint, cam, heatmap, bbox, K):
    """HandAggregator.select_by_heatmap_cascade (lib/model/aggregation.py:115-178), use_regression_as_candidate=True."""
    bs = root_joint.shape[0]
    pose = pose_diff.clone().reshape(bs, -1, 48)
    S = pose.shape[1]
    extra = torch.zeros_like(pose) + pose_regression[:, None].clone()
    pose = torch.cat((pose, extra), dim=1).reshape(-1, 48)
    shape2 = shape.clone().reshape(bs, -1, 10).repeat(1, 2, 1).reshape(-1, 10)
    levels = []
    for lv in range(4):
        fuse_idx = MANO_PARAMS_LEVEL[lv]
        observe = []
        for j in range(lv + 1, 5):
            observe.extend(MANO_JOINT_LEVEL[j])
        if lv == 0:
            pose = pose.view(bs, -1, 48)
            pose[:, S:, fuse_idx] = pose[:, :S, fuse_idx]
            pose = pose.reshape(-1, 48)
        d = _hand_level(mano, pose, shape2, root_joint, cam, heatmap, bbox, K, fuse_idx, observe, independent=lv != 0)
        pose = d["fused_pose"].reshape(-1, 48)
        levels.append(d)
    fused_pose = levels[-1]["fused_pose"][:, 0]
    shp = shape.reshape(bs, -1, 10)[:, 0]
    fused_mano = torch.cat((fused_pose, shp), dim=-1)
    fv, fj = mano(fused_pose, shp)
    return {"agg_hand_mano": fused_mano, "agg_vert": fv.reshape(bs, 778, 3), "agg_joint": fj.reshape(bs, 21, 3),
            "levels": levels}


# ---------------------------------------------------------------------------------------------------------
# object branch: heat-map selection, physics3 selection, fusion
# ---------------------------------------------------------------------------------------------------------
def obj_heat_topk(obj: OracleObject, pose6d, root_joint, obj_name, cam, heatmap, bbox, k, is_right):
    """ObjectAggregator.select_topk_object_by_heatmap (lib/model/aggregation.py:742-780)."""
    p = _geom(pose6d.clone())
    p[..., 6:] = p[..., 6:] + root_joint.unsqueeze(1)
    pts = obj(p, obj_name)
    pts = obj.flip_pt3d(pts, is_right)
    pt2d = project(pts, cam)
    bb = bbox[:, None, None, :]
    pt2d = pt2d - bb[..., :2]
    pt2d = 2 * pt2d / (bb[..., 2:] - bb[..., :2]) - 1
    J = heatmap.shape[1]
    hv = sample_heat(heatmap, pt2d, list(range(J))).sum(dim=-1)
    val, topk = canonical_topk(hv, k, dim=1)
    weight = (val + 1e-8) / (val.sum(dim=1, keepdim=True) + 1e-8)
    return topk, weight, hv


def obj_fuse_topk(topk, pose6d, weight=None):
    """ObjectAggregator.fuse_topk (lib/model/aggregation.py:729-740)."""
    bs = pose6d.size(0)
    bidx = torch.arange(bs)[:, None].repeat(1, topk.size(1))
    sel = pose6d[bidx, topk]
    if weight is None:
        trans = sel[:, :, 6:].mean(dim=1)
    else:
        trans = (sel[:, :, 6:] * weight[:, :, None]).sum(dim=1)
    rot = average_rot6d(sel[..., :6], weights=weight)
    return torch.cat([rot, trans], dim=-1)


def obj_physics3_topk(obj: OracleObject, pose6d, root_joint, obj_name, is_right, force_point, force_global, k):
    """ObjectAggregator.select_topk_object_by_physics3 (lib/model/aggregation.py:947-997); the two cdist passes of
    cdist_memory_save / nn_for_r_memory_save (:1115-1142) are kept as in the reference."""
    p = _geom(pose6d.clone())
    p[..., 6:] = p[..., 6:] + root_joint.unsqueeze(1)
    ov = obj.flip_pt3d(obj(p, obj_name, data_name="verts"), is_right)
    oc = obj.flip_pt3d(obj(p, obj_name, data_name="CoM"), is_right)
    fnorm = force_global.norm(dim=-1)
    fw = fnorm / fnorm.sum(dim=-1, keepdim=True)
    fp = force_point[:, None]
    bs = fp.shape[0]
    d_ls, r_ls = [], []
    for i in range(bs):
        d_ls.append(exact_cdist(fp[i], ov[i]).min(dim=-1)[0])
    cd = torch.stack(d_ls, 0)
    score = (cd * fw[:, None]).sum(-1)
    fdir = force_global / fnorm[:, :, None]
    for i in range(bs):
        am = exact_cdist(fp[i], ov[i]).min(dim=-1)[1]
        ind = torch.arange(am.shape[0])[:, None].repeat(1, am.shape[1])
        r_ls.append(fp[i] - ov[i][ind, am])
    r = torch.stack(r_ls, 0) - oc
    L = torch.cross(fdir[:, None].expand_as(r), r, dim=-1).sum(-2).norm(dim=-1)
    score = -(score * L)
    val, topk = canonical_topk(score, k, dim=1)
    weight = torch.ones_like(val)
    weight = weight / weight.sum(dim=1, keepdim=True)
    return topk, weight, score


def hand_physics_select(mano: OracleMano, anchors: OracleAnchors, pose, root_joint_flip, obj_vert, obj_com,
                        force_local, K):
    """HandAggregator.select_by_physics (lib/model/aggregation.py:537-626)."""
    bs = pose.shape[0]
    p = pose.reshape(-1, 58)
    vert, _ = mano(p[:, :48], p[:, 48:])
    vert = vert.reshape(bs, -1, 778, 3)
    vert_cam = vert + root_joint_flip[:, None, None]
    C = vert_cam.shape[1]
    fl = force_local[:, None].repeat(1, C, 1, 1).reshape(-1, 32, 3)
    fpt, fg = anchors.from_local_to_global(fl, vert_cam.reshape(-1, 778, 3))
    fpt, fg = fpt.reshape(bs, -1, 32, 3), fg.reshape(bs, -1, 32, 3)
    fnorm = fg.norm(dim=-1)
    fw = fnorm / fnorm.sum(dim=-1, keepdim=True)
    cd = torch.stack([exact_cdist(fpt[i], obj_vert[i]).min(dim=-1)[0] for i in range(bs)], 0)
    score = fw * cd
    fgn = fg / fnorm[:, :, :, None]
    # dead code in the reference (`L` is computed from force directions and never used, :569-576); kept for cost parity
    r_ls = []
    for i in range(bs):
        am = exact_cdist(fgn[i], obj_vert[i][None]).min(dim=-1)[1]
        r_ls.append(fgn[i] - obj_vert[i][am])
    _ = torch.cross(fgn, torch.stack(r_ls, 0) - obj_com[:, None], dim=-1).sum(-2).norm(dim=-1)
    I = fgn.sum(-2).norm(dim=-1)
    score = -(score * I[:, :, None])
    topks, fscores = [], []
    for f in range(5):
        fs = score[:, :, FINGER_FORCE_LEVEL[f]].sum(dim=-1)
        _, tk = canonical_topk(fs, K, dim=1)
        topks.append(tk)
        fscores.append(fs)
    topks = torch.stack(topks, 1)
    fscores = torch.stack(fscores, 1)
    bidx = torch.arange(bs)[:, None].repeat(1, K)
    fuse_pose = pose[:, 0].clone()
    for f in range(5):
        fuse_idx = MANO_PARAMS_LEVEL[2][3 * f:3 * f + 3] + MANO_PARAMS_LEVEL[3][3 * f:3 * f + 3]
        tp = pose[:, :, :48][bidx, topks[:, f]][:, :, fuse_idx].reshape(-1, K, 2, 3)
        q = axis_angle_to_quaternion(tp).permute(0, 2, 1, 3)
        q = average_quaternion(q)
        fuse_pose[:, fuse_idx] = quaternion_to_axis_angle(q).reshape(-1, 6)
    fv, fj = mano(fuse_pose[:, :48], fuse_pose[:, 48:])
    return {"agg_pose": fuse_pose, "agg_vert": fv, "agg_joint": fj, "finger_topk": topks, "finger_score": fscores}


def hoi_aggregate(mano: OracleMano, obj: OracleObject, anchors: OracleAnchors, *, cam_intrinsic, root_joint_flip,
                  root_joint, is_right, force_local, is_grasped, hand_pose_diff, hand_pose_regression, hand_shape,
                  hand_heatmap, hand_bbox, hand_topk, obj_pose6d, obj_heatmap, obj_bbox, obj_topk, obj_name,
                  phy_topk: int = 5):
    """HOI_Aggregator.__call__ (lib/model/aggregation.py:1167-1353)."""
    hs = hand_cascade(mano, hand_pose_diff, hand_pose_regression, hand_shape, root_joint_flip, cam_intrinsic,
                      hand_heatmap, hand_bbox, hand_topk)
    agg_mano = hs["agg_hand_mano"]
    hand_vert = hs["agg_vert"] + root_joint_flip[:, None]
    force_point, force_global = anchors.from_local_to_global(force_local, hand_vert)

    t_topk, t_w, t_score = obj_heat_topk(obj, obj_pose6d, root_joint, obj_name, cam_intrinsic, obj_heatmap, obj_bbox,
                                         obj_topk, is_right)
    fused = obj_fuse_topk(t_topk, obj_pose6d, t_w)
    transl_fused = fused[:, 6:]
    upd = obj_pose6d.clone()
    upd[..., 6:] = transl_fused[:, None]
    r_topk, _, r_score = obj_heat_topk(obj, upd, root_joint, obj_name, cam_intrinsic, obj_heatmap, obj_bbox, obj_topk,
                                       is_right)
    bs = obj_pose6d.size(0)
    bidx = torch.arange(bs)[:, None].repeat(1, obj_topk)
    c_t = obj_pose6d[bidx, t_topk][:, :, 6:][:, :, None].repeat(1, 1, obj_topk, 1)
    c_r = obj_pose6d[bidx, r_topk][:, :, :6][:, None, :].repeat(1, obj_topk, 1, 1)
    cand = torch.cat([c_r, c_t], dim=-1).reshape(bs, -1, 9)

    p_topk, p_w, p_score = obj_physics3_topk(obj, cand, root_joint, obj_name, is_right, force_point, force_global,
                                             phy_topk)
    h_topk, h_w, h_score = obj_heat_topk(obj, cand, root_joint, obj_name, cam_intrinsic, obj_heatmap, obj_bbox,
                                         phy_topk, is_right)
    ung = torch.arange(bs)[~is_grasped]
    new_topk = p_topk.clone()
    new_topk[ung] = h_topk[ung]
    new_w = p_w.clone()
    new_w[ung] = h_w[ung]
    pose6d_fused = obj_fuse_topk(new_topk, cand, new_w)

    p = _geom(pose6d_fused.clone())
    p[..., 6:] = p[..., 6:] + root_joint
    ov = obj.flip_pt3d(obj(p, obj_name, data_name="verts"), is_right)
    oc = obj.flip_pt3d(obj(p, obj_name, data_name="CoM"), is_right)

    l3 = agg_mano[:, MANO_PARAMS_LEVEL[2]].reshape(-1, 1, 5, 3)
    l4 = hs["levels"][3]["topk_idx_pose_aa"][:, :hand_topk]
    l4 = torch.cat([l4, agg_mano[:, MANO_PARAMS_LEVEL[3]].reshape(-1, 1, 5, 3)], dim=1)
    nc = hand_topk + 1
    newp = torch.cat([l3[:, :, None].repeat(1, 1, nc, 1, 1), l4[:, None]], dim=-2).reshape(bs, nc, -1, 3)
    cpose = agg_mano.clone()[:, None, :48].repeat(1, nc, 1)
    cpose[:, :, MANO_PARAMS_LEVEL[2]] = newp[:, :, :5].reshape(-1, nc, 15)
    cpose[:, :, MANO_PARAMS_LEVEL[3]] = newp[:, :, 5:].reshape(-1, nc, 15)
    cpose = torch.cat([cpose, agg_mano[:, None, 48:].repeat(1, nc, 1)], dim=-1)
    hp = hand_physics_select(mano, anchors, cpose, root_joint_flip, ov, oc, force_local, phy_topk)
    return {
        "obj_agg_6d": pose6d_fused, "pose6d_candidate": cand, "agg_obj_vert": ov,
        "hand_agg_mano": hp["agg_pose"], "hand_agg_vert": hp["agg_vert"], "hand_agg_joint": hp["agg_joint"],
        # extra diagnostics for stage-wise parity (not part of the reference's return dict)
        "_dbg": {
            "cascade": hs, "force_point": force_point, "force_global": force_global,
            "obj_transl_topk": t_topk, "obj_transl_score": t_score, "obj_transl_weight": t_w,
            "obj_transl_fused": transl_fused, "obj_rot_topk": r_topk, "obj_rot_score": r_score,
            "phys_topk": p_topk, "phys_score": p_score, "heat5_topk": h_topk, "heat5_score": h_score,
            "heat5_weight": h_w, "obj_com": oc, "hand_cand_pose": cpose,
            "finger_topk": hp["finger_topk"], "finger_score": hp["finger_score"],
        },
    }


# ---------------------------------------------------------------------------------------------------------
# G1: predict-branch glue (lib/model/VPHO.py:228-304) and the "final pose error" metrics
# ---------------------------------------------------------------------------------------------------------
def oracle_predict(batch: dict, den_hand: OracleDenoiser, den_obj: OracleDenoiser, mano: OracleMano,
                   obj: OracleObject, anchors: OracleAnchors, *, init_x_hand, init_x_obj, sample_num=100,
                   sampling_steps=50, T0=0.65, topk_hand=30, topk_obj=10, with_inprocess=False, timing=None):
    import time
    t = lambda a: torch.as_tensor(np.asarray(a))  # noqa: E731
    bs = t(batch["encoding_hand"]).shape[0]
    enc_h, enc_o = t(batch["encoding_hand"]).float(), t(batch["encoding_obj"]).float()
    pd_pose, pd_shape = t(batch["pd_mano_pose"]).float(), t(batch["pd_mano_shape"]).float()
    out = {}
    t0 = time.perf_counter()
    xs_h, x_h, info_h = oracle_sample(den_hand, enc_h[:, None].repeat(1, sample_num, 1).reshape(-1, enc_h.shape[1]),
                                      T0, init_x_hand, sampling_steps)
    t1 = time.perf_counter()
    xs_hf, x_hf = xs_h.float(), x_h.float()
    inproc, final_mano = postprocess_diffusion_hand(x_hf, pd_shape, sample_num, xs_hf if with_inprocess else None)
    out["diff_final_hand_mano"] = final_mano.reshape(bs, sample_num, 58)
    if with_inprocess:
        out["diff_inprocess_hand_mano"] = inproc.reshape(bs, sample_num, -1, 58)
    fv, fj = mano(final_mano[:, :48], final_mano[:, 48:])
    out["diff_final_hand_vert"] = fv.reshape(bs, sample_num, 778, 3)
    out["diff_final_hand_joint"] = fj.reshape(bs, sample_num, 21, 3)
    t2 = time.perf_counter()
    xs_o, x_o, info_o = oracle_sample(den_obj, enc_o[:, None].repeat(1, sample_num, 1).reshape(-1, enc_o.shape[1]),
                                      T0, init_x_obj, sampling_steps)
    out["diff_final_obj_6d"] = x_o.reshape(bs, sample_num, 9)
    if with_inprocess:
        out["diff_inprocess_obj_6d"] = xs_o.reshape(bs, sample_num, -1, 9)
    t3 = time.perf_counter()
    sel = hoi_aggregate(
        mano, obj, anchors, cam_intrinsic=t(batch["cam_intr_crop_flip"]).float(),
        root_joint_flip=t(batch["root_joint_flip"]).float(), root_joint=t(batch["root_joint"]).float(),
        is_right=t(batch["is_right"]).bool(), force_local=t(batch["force_local"]).float(),
        is_grasped=t(batch["is_grasped"]).bool(), hand_pose_diff=final_mano[:, :48].clone(),
        hand_pose_regression=pd_pose, hand_shape=final_mano[:, 48:], hand_heatmap=t(batch["hm_hand"]).float(),
        hand_bbox=t(batch["bbox_hand"]).float(), hand_topk=topk_hand, obj_pose6d=out["diff_final_obj_6d"],
        obj_heatmap=t(batch["hm_obj"]).float(), obj_bbox=t(batch["bbox_obj_rect"]).float(), obj_topk=topk_obj,
        obj_name=list(batch["obj_name"]))
    t4 = time.perf_counter()
    out["agg_obj_6d"] = sel["obj_agg_6d"]
    out["agg_hand_mano"] = sel["hand_agg_mano"]
    out["agg_hand_vert"] = sel["hand_agg_vert"]
    out["agg_hand_joint"] = sel["hand_agg_joint"]
    out["_sel"] = sel
    out["_info"] = {"hand": info_h, "obj": info_o}
    if timing is not None:
        timing.update(hand_sample=t1 - t0, mano_finals=t2 - t1, obj_sample=t3 - t2, aggregate=t4 - t3)
    return out


def cone_anchor_table(friction_coeff: float = 0.8) -> torch.Tensor:
    """`HeadPhysics.anchor` (lib/model/physics.py:692-698) with xy scaled as get_local_force does (:549-550)."""
    num_anchor = 8
    a = torch.arange(0, 2 * torch.pi, 2 * torch.pi / num_anchor)[:num_anchor]
    anchor = torch.stack([torch.cos(a), torch.sin(a), torch.ones_like(a)], dim=-1) / num_anchor
    anchor = anchor.clone()
    anchor[:, :2] *= friction_coeff
    return anchor


def get_local_force(scale: torch.Tensor, weight: torch.Tensor, friction_coeff: float = 0.8):
    """HeadPhysics.get_local_force (lib/model/physics.py:546-557)."""
    scale = torch.abs(scale)
    weight = torch.softmax(weight, dim=-1)
    anchor = cone_anchor_table(friction_coeff).to(weight.dtype)[None, :].repeat_interleave(scale.size(-1), dim=0)
    direction = torch.einsum("...ij,ijk->...ik", weight, anchor)
    direction = direction / (direction.norm(dim=-1, keepdim=True) + 1e-8)
    return direction * scale[..., None]


def force_eval_terms(anchors: "OracleAnchors", vert3d, scale, weight, contact_mask, force_contact, gravity, com):
    """Per-sample forward terms of one ForceOptimizer iteration (lib/engine/force_optimization.py:141-171), before the
    batch means and loss weights: |sum f + g|, (sum f).(-g), |sum (p - CoM) x f|, mean_j (log|c_j/s_j| mask_j)^2.
    vert3d (n,778,3); scale (n,32); weight (n,32,8); contact_mask (n,32) bool; force_contact (n,32); gravity, com (n,1,3)."""
    scale = scale * contact_mask
    force_local = get_local_force(scale, weight)
    force_point, force_global = anchors.from_local_to_global(force_local, vert3d)
    resultant = (force_global.sum(1, keepdim=True) + gravity).squeeze(1)
    t0 = torch.norm(resultant, dim=-1)
    t1 = torch.einsum("...i,...i->...", force_global.sum(1, keepdim=True), -1 * gravity).squeeze(-1)
    moment = torch.cross(force_point - com, force_global, dim=-1).sum(1)
    t2 = torch.norm(moment, dim=-1)
    scale_norm = scale / (scale.norm(dim=-1, keepdim=True) + 1e-8)
    fc_norm = force_contact / (force_contact.norm(dim=-1, keepdim=True) + 1e-8)
    dist = torch.log(torch.abs(fc_norm / (scale_norm + 1e-8)) + 1e-8) * contact_mask
    t3 = (dist ** 2).mean(dim=-1)
    return torch.stack([t0, t1, t2, t3], dim=-1), force_local, force_point, force_global


def hand_pose_error_mm(pd_joint, gt_joint, pd_vert, gt_vert):
    """MJE / MVE as in TesterHand (lib/engine/test.py:657-679): mean Euclidean distance, metres -> mm."""
    mje = (pd_joint - gt_joint).norm(dim=-1).mean(dim=-1) * 1000
    mve = (pd_vert - gt_vert).norm(dim=-1).mean(dim=-1) * 1000
    return mje, mve


def rigid_align_AtoB(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """Similarity Procrustes alignment of A onto B, `rigid_transform_3D_AtoB` + `rigid_align_AtoB`
    (lib/utils/transform_fn.py:43-66), statement by statement in the dtype of the inputs."""
    n, dim = A.shape
    centroid_A = np.mean(A, axis=0)
    centroid_B = np.mean(B, axis=0)
    H = np.dot(np.transpose(A - centroid_A), B - centroid_B) / n
    U, s, V = np.linalg.svd(H)
    R = np.dot(np.transpose(V), np.transpose(U))
    if np.linalg.det(R) < 0:
        s[-1] = -s[-1]
        V[2] = -V[2]
        R = np.dot(np.transpose(V), np.transpose(U))
    varP = np.var(A, axis=0).sum()
    c = 1 / varP * np.sum(s)
    t = -np.dot(c * R, np.transpose(centroid_A)) + np.transpose(centroid_B)
    return np.transpose(np.dot(c * R, np.transpose(A))) + t


def hand_pa_error_mm(pd_joint, gt_joint, pd_vert, gt_vert):
    """PA-MJE / PA-MVE and the per-joint errors JE of `TesterHand.criterion_MJE_PAMJE` (lib/engine/test.py:657-679):
    -> (pa_mje (n,), pa_mve (n,), je (n, 21)) in mm."""
    pj, gj, pv, gv = (np.asarray(t, np.float32) for t in (pd_joint, gt_joint, pd_vert, gt_vert))
    pa_mje = np.stack([np.linalg.norm(gj[i] - rigid_align_AtoB(pj[i], gj[i]), axis=-1).mean(-1) for i in range(pj.shape[0])])
    pa_mve = np.stack([np.linalg.norm(gv[i] - rigid_align_AtoB(pv[i], gv[i]), axis=-1).mean(-1) for i in range(pv.shape[0])])
    je = np.linalg.norm(gj - pj, axis=-1)
    return pa_mje * 1000, pa_mve * 1000, je * 1000


def object_add_mm(obj: OracleObject, pd_6d, gt_6d, obj_name):
    """ADD / ADD-S of TesterObject.criterion_ADD_REP (lib/engine/test.py:413-442) on the sampled vertices."""
    pv = obj(pd_6d.float(), obj_name, data_name="verts")
    gv = obj(gt_6d.float(), obj_name, data_name="verts")
    add = (pv - gv).norm(dim=-1).mean(dim=-1) * 1000
    adds = torch.stack([exact_cdist(pv[i], gv[i]).min(dim=-1)[0].mean() for i in range(pv.shape[0])]) * 1000
    return add, adds


def force_optimize(anchors: "OracleAnchors", vert3d, force_contact, gravity, com, n_iter: int = 3000, switch_iter: int = 300,
                   lr: float = 1e-3, trace: bool = False):
    """ForceOptimizer.optimize_batch for ONE batch (lib/engine/force_optimization.py:110-207), autograd + torch.optim.AdamW
    exactly as the reference builds them (get_optimizer :33-37: two AdamW optimisers, betas (0.9, 0.999), eps 1e-8,
    lr 1e-3, default weight decay; init_param :42-44: scale 0.05, weight 0).  Inputs already in the flipped frame
    (:134-137).  Returns the parameters after the last step and the forces of the last iteration's forward pass (what the
    reference saves, :185-206), plus the loss trace when asked."""
    from torch import nn, optim
    bs = vert3d.shape[0]
    contact_mask = force_contact > 0.1
    dt = vert3d.dtype               # float32 as the reference; float64 inputs give the shadow run the tests derive their bar from
    scale_p = nn.Parameter(torch.ones(bs, 32, dtype=dt) * 0.05)
    weight_p = nn.Parameter(torch.zeros(bs, 32, 8, dtype=dt))
    opt1 = optim.AdamW([weight_p], betas=(0.9, 0.999), eps=1e-8, lr=lr)
    opt2 = optim.AdamW([scale_p, weight_p], betas=(0.9, 0.999), eps=1e-8, lr=lr)
    gravity = gravity.reshape(bs, 1, 3)
    com = com.reshape(bs, 1, 3)
    losses = []
    force_local = force_global = None
    for i in range(n_iter):
        scale = scale_p.clone() * contact_mask
        weight = weight_p.clone()
        force_local = get_local_force(scale, weight)
        force_point, force_global = anchors.from_local_to_global(force_local, vert3d)
        resultant = (force_global.sum(1, keepdim=True) + gravity).squeeze(1)
        force_loss = torch.norm(resultant, dim=-1).mean()
        sum_weight = force_loss.detach()
        resultant2 = force_global.sum(1, keepdim=True)
        cos_proj = torch.einsum("...i,...i->...", resultant2, -1 * gravity)
        gravity_loss = F.mse_loss(cos_proj, torch.ones_like(cos_proj))
        moment = torch.cross(force_point - com, force_global, dim=-1).sum(1)
        moment_loss = torch.norm(moment, dim=-1).mean() * 30
        moment_loss = moment_loss / (100 * sum_weight ** 2 + 1e-8)
        scale_norm = scale / (scale.norm(dim=-1, keepdim=True).detach() + 1e-8).detach()
        fc_norm = force_contact / (force_contact.norm(dim=-1, keepdim=True).detach() + 1e-8)
        dist = torch.log(torch.abs(fc_norm / (scale_norm + 1e-8)) + 1e-8) * contact_mask
        dist_loss = (dist ** 2).mean() * 0.1
        dist_loss = dist_loss / (1000 * sum_weight ** 2 + 1e-8)
        if i < switch_iter:
            loss, opt = gravity_loss, opt1
        else:
            loss, opt = force_loss + moment_loss + dist_loss, opt2
        opt.zero_grad()
        loss.backward()
        opt.step()
        if trace:
            losses.append([float(x.detach()) for x in (loss, force_loss, gravity_loss, moment_loss, dist_loss)])
    out = {"scale": scale_p.detach().clone(), "weight": weight_p.detach().clone(), "force_local": force_local.detach(),
           "force_global": force_global.detach()}
    if trace:
        out["losses"] = torch.tensor(losses)
    return out
