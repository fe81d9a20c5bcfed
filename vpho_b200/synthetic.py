"""Seeded synthetic assets and DexYCB-shaped eval batches.

Nothing here is a kernel or a fallback: it only manufactures *inputs* with the shapes and
dtypes the reference's hot path consumes, because the licensed assets (MANO_RIGHT.pkl,
asset/2021_CVPR_CPF/anchor/*, asset/ours/vert2joint.pkl, the YCB meshes) and DexYCB itself are
not available offline.  A user with the real assets passes them through the same dict layouts.

Layouts produced (all numpy, float32 unless noted):
  mano model   : v_template (778,3), shapedirs (778,3,10), posedirs (778,3,135),
                 J_regressor (16,778), weights (778,16)            [manopth ManoLayer buffers]
  anchors      : face_vertex_idx (32,3) int32, anchor_weight (32,2), vert2joint (21,778)
                 [lib/utils/physics_fn.py:121-140, lib/utils/hand_fn.py:427-433]
  object tables: names[21], kpt3d (21,27,3), verts_sampled (21,2048,3), CoM (21,3)
                 [lib/model/head_object.py:9-34, lib/utils/misc_fn.py:42-67]
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np

# manopth kinematic order: 0 wrist; (1,2,3) index; (4,5,6) middle; (7,8,9) pinky; (10,11,12) ring; (13,14,15) thumb
_FINGER_BASE = {"index": 1, "middle": 4, "pinky": 7, "ring": 10, "thumb": 13}
# manopth fingertip vertex ids for the right hand, order thumb,index,middle,ring,pinky
TIP_VERTS = (745, 317, 444, 556, 673)
_TIP_OF_FINGER = {"thumb": 745, "index": 317, "middle": 444, "ring": 556, "pinky": 673}

YCB_NAMES: List[str] = [
    "002_master_chef_can", "003_cracker_box", "004_sugar_box", "005_tomato_soup_can",
    "006_mustard_bottle", "007_tuna_fish_can", "008_pudding_box", "009_gelatin_box",
    "010_potted_meat_can", "011_banana", "019_pitcher_base", "021_bleach_cleanser",
    "024_bowl", "025_mug", "035_power_drill", "036_wood_block", "037_scissors",
    "040_large_marker", "051_large_clamp", "052_extra_large_clamp", "061_foam_brick",
]


def _rest_skeleton() -> Dict[str, np.ndarray]:
    """Rest joint positions (metres) of a right hand, palm facing -z, fingers along +x."""
    J = np.zeros((16, 3), np.float64)
    tips = {}
    spec = {  # finger: (mcp position, direction, segment lengths)
        "index": ((0.088, 0.028, 0.0), (1.0, 0.10, 0.0), (0.034, 0.024, 0.022)),
        "middle": ((0.092, 0.004, 0.0), (1.0, 0.00, 0.0), (0.038, 0.027, 0.023)),
        "ring": ((0.086, -0.018, 0.0), (1.0, -0.10, 0.0), (0.035, 0.025, 0.022)),
        "pinky": ((0.074, -0.038, 0.0), (1.0, -0.22, 0.0), (0.026, 0.018, 0.019)),
        "thumb": ((0.026, 0.034, -0.012), (0.70, 0.68, -0.20), (0.036, 0.030, 0.026)),
    }
    for name, (mcp, d, seg) in spec.items():
        b = _FINGER_BASE[name]
        d = np.asarray(d, np.float64)
        d /= np.linalg.norm(d)
        p = np.asarray(mcp, np.float64)
        J[b] = p
        J[b + 1] = p + d * seg[0]
        J[b + 2] = J[b + 1] + d * seg[1]
        tips[name] = J[b + 2] + d * seg[2]
    return {"J": J, "tips": tips}


def make_mano_model(seed: int = 1234, max_influences: int = 4) -> Dict[str, np.ndarray]:
    """A 778-vertex pseudo-hand with MANO's tensor shapes (SURVEY.md §8d 'MANO model').
    max_influences: non-zero skinning weights per vertex.  SMPL-family models (MANO is one) constrain every vertex to at most
    4 influencing joints (SMPL, Loper et al. 2015, sec. 3: blend weights kept sparse for compatibility with rendering
    engines), so 4 is the default; None gives a fully dense 778 x 16 matrix (the kernels' dense skinning path, kept under
    test)."""
    rng = np.random.default_rng(seed)
    sk = _rest_skeleton()
    J, tips = sk["J"], sk["tips"]
    # bones: (parent joint, child position, child joint or -1)
    bones = []
    for name, b in _FINGER_BASE.items():
        bones.append((0, J[b], b))            # palm: wrist -> mcp
        bones.append((b, J[b + 1], b + 1))
        bones.append((b + 1, J[b + 2], b + 2))
        bones.append((b + 2, tips[name], -1))
    nb = len(bones)
    V = 778
    v = np.zeros((V, 3), np.float64)
    w = np.zeros((V, 16), np.float64)
    bone_of = rng.integers(0, nb, V)
    u = rng.random(V)
    for i in range(V):
        pj, cpos, cj = bones[bone_of[i]]
        p0 = J[pj]
        axis = cpos - p0
        L = np.linalg.norm(axis)
        a = axis / L
        # radial offset orthogonal to the bone
        r = rng.normal(size=3)
        r -= a * (r @ a)
        r /= np.linalg.norm(r) + 1e-12
        rad = 0.011 if pj == 0 else 0.0075
        v[i] = p0 + axis * u[i] + r * rad * (0.6 + 0.4 * rng.random())
        # skinning weights: parent joint dominates, blends to child joint / grand-parent
        ww = np.full(16, 1e-3) * rng.random(16)
        if max_influences is not None:
            # keep the noise only on the strongest `max_influences - 2` joints besides the bone's own two
            own = {pj} if cj < 0 else {pj, cj}
            others = sorted((j for j in range(16) if j not in own), key=lambda j: -ww[j])
            for j in others[max(0, max_influences - len(own)):]:
                ww[j] = 0.0
        ww[pj] += 1.0 - 0.5 * u[i]
        if cj >= 0:
            ww[cj] += 0.5 * u[i]
        else:
            ww[pj] += 0.5 * u[i]
        w[i] = ww / ww.sum()
    # fingertips sit exactly on the extended finger ends
    for name, vid in _TIP_OF_FINGER.items():
        v[vid] = tips[name]
        b = _FINGER_BASE[name]
        ww = np.zeros(16)
        ww[b + 2] = 1.0
        w[vid] = ww
    # joint regressor: each joint regressed from the vertices closest to it
    Jreg = np.zeros((16, V), np.float64)
    for j in range(16):
        d = np.linalg.norm(v - J[j], axis=1)
        idx = np.argsort(d)[:24]
        cw = rng.random(24) + 0.2
        Jreg[j, idx] = cw / cw.sum()
    shapedirs = rng.normal(size=(V, 3, 10)) * 0.0025
    posedirs = rng.normal(size=(V, 3, 135)) * 0.0010
    return {
        "v_template": v.astype(np.float32),
        "shapedirs": shapedirs.astype(np.float32),
        "posedirs": posedirs.astype(np.float32),
        "J_regressor": Jreg.astype(np.float32),
        "weights": w.astype(np.float32),
    }


def make_anchor_assets(mano: Dict[str, np.ndarray], seed: int = 4321) -> Dict[str, np.ndarray]:
    """32 force anchors as vertex triples + barycentric-like weights, and the dense 21x778
    vert2joint matrix (stand-ins for asset/2021_CVPR_CPF/anchor/* and asset/ours/vert2joint.pkl)."""
    rng = np.random.default_rng(seed)
    v = mano["v_template"].astype(np.float64)
    V = v.shape[0]
    face = np.zeros((32, 3), np.int32)
    for a in range(32):
        c = rng.integers(0, V)
        d = np.linalg.norm(v - v[c], axis=1)
        near = np.argsort(d)[1:12]
        pick = rng.choice(near, 2, replace=False)
        face[a] = (c, pick[0], pick[1])
    anchor_weight = (rng.random((32, 2)) * 0.5).astype(np.float32)
    # vert2joint: 16 regressed joints in the 21-joint output order + 5 one-hot fingertips
    Jreg = mano["J_regressor"].astype(np.float64)
    tips = np.zeros((5, V))
    for k, vid in enumerate(TIP_VERTS):
        tips[k, vid] = 1.0
    j21 = np.concatenate([Jreg, tips], 0)
    order = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]
    v2j = j21[order]
    # the real asset is a fitted, slightly noisy regressor ("not precise", hand_fn.py:449)
    noise = rng.random((21, V)) * (rng.random((21, V)) > 0.985) * 0.02
    v2j = v2j + noise
    v2j /= v2j.sum(1, keepdims=True)
    return {
        "face_vertex_idx": face,
        "anchor_weight": anchor_weight,
        "vert2joint": v2j.astype(np.float32),
    }


def make_object_tables(seed: int = 777, n_verts: int = 2048) -> Dict[str, object]:
    """Per-object point tables for the 21 YCB names (head_object.py:9-34)."""
    rng = np.random.default_rng(seed)
    n = len(YCB_NAMES)
    kpt = np.zeros((n, 27, 3), np.float32)
    verts = np.zeros((n, n_verts, 3), np.float32)
    com = np.zeros((n, 3), np.float32)
    for o in range(n):
        half = rng.uniform(0.025, 0.125, 3)            # 5-25 cm box
        # FPS-like: jittered points on an ellipsoid/box blend surface
        p = rng.normal(size=(n_verts, 3))
        p /= np.linalg.norm(p, axis=1, keepdims=True)
        boxy = p / np.max(np.abs(p), axis=1, keepdims=True)
        mix = rng.uniform(0.2, 0.8)
        s = (mix * p + (1 - mix) * boxy) * half
        s += rng.normal(size=3) * 0.004                # mesh origin is not the centroid
        verts[o] = s.astype(np.float32)
        mn, mx = verts[o].min(0), verts[o].max(0)
        k = []
        for i in range(3):
            for j in range(3):
                for l in range(3):
                    wgt = np.array([i, j, l], np.float32) / 2
                    k.append(mn + wgt * (mx - mn))
        kpt[o] = np.stack(k, 0)
        com[o] = verts[o].astype(np.float64).mean(0).astype(np.float32)
    return {"names": list(YCB_NAMES), "kpt3d": kpt, "verts_sampled": verts, "CoM": com}


def make_denoiser_state(head: str, seed: int = 0, last_std: float = 0.05) -> Dict[str, np.ndarray]:
    """State-dict of BaseDenoiser (denoiser.py:33-66) with the reference's init
    (nn.Linear N(0,.01^2)/bias 0 via init_weights VPHO.py:34-45; ParallelLinear kaiming-uniform,
    parallel_linear.py:19-25) and the zero-initialised last layer re-drawn N(0,last_std^2)
    (SURVEY.md §8c-iv).  Keys follow the reference module tree."""
    assert head in ("mano_pose", "obj")
    D, n = (96, 32) if head == "mano_pose" else (9, 3)
    rng = np.random.default_rng(seed + (0 if head == "mano_pose" else 1000003))
    f32 = np.float32
    st = {}
    st["t_encoder.0.W"] = (rng.normal(size=64) * 30.0).astype(f32)
    st["t_encoder.1.weight"] = (rng.normal(size=(128, 128)) * 0.01).astype(f32)
    st["t_encoder.1.bias"] = np.zeros(128, f32)
    st["pose_encoder.0.weight"] = (rng.normal(size=(256, D)) * 0.01).astype(f32)
    st["pose_encoder.0.bias"] = np.zeros(256, f32)
    st["pose_encoder.2.weight"] = (rng.normal(size=(256, 256)) * 0.01).astype(f32)
    st["pose_encoder.2.bias"] = np.zeros(256, f32)
    # ParallelLinear(1408,256,n): weight (n,1408,256); torch's fan_in for a 3-D tensor is size(1)*prod(size[2:])
    fan_in = 1408 * 256
    bound_w = math.sqrt(6.0 / ((1 + 5.0) * fan_in))  # kaiming_uniform(a=sqrt(5)): gain*sqrt(3/fan_in)
    st["head.head.0.weight"] = rng.uniform(-bound_w, bound_w, size=(n, 1408, 256)).astype(f32)
    st["head.head.0.bias"] = rng.uniform(-1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in), size=(n, 256)).astype(f32)
    st["head.head.2.weight"] = (rng.normal(size=(n, 256, 3)) * last_std).astype(f32)
    st["head.head.2.bias"] = (rng.normal(size=(n, 3)) * last_std).astype(f32)
    return st


def _rodrigues_np(aa: np.ndarray) -> np.ndarray:
    th = np.linalg.norm(aa, axis=-1, keepdims=True)
    k = aa / np.maximum(th, 1e-12)
    K = np.zeros(aa.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s, c = np.sin(th)[..., None], np.cos(th)[..., None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def _gauss_heatmaps(uv: np.ndarray, size: int, sigma: float, rng, noise: float) -> np.ndarray:
    """uv (bs,J,2) in heat-map pixel units -> (bs,J,size,size) sums of Gaussians + U(0,noise)."""
    bs, J, _ = uv.shape
    ys, xs = np.mgrid[0:size, 0:size].astype(np.float64)
    hm = np.exp(-((xs[None, None] - uv[..., 0, None, None]) ** 2 + (ys[None, None] - uv[..., 1, None, None]) ** 2)
                / (2 * sigma ** 2))
    hm += rng.random((bs, J, size, size)) * noise
    return hm.astype(np.float32)


def make_eval_batch(bs: int, seed: int = 0, sample_num: int = 100,
                    mano: Dict[str, np.ndarray] | None = None,
                    objects: Dict[str, object] | None = None) -> Dict[str, object]:
    """Synthetic stand-ins for everything `vpho_net.forward(mode='predict')` computes *before* it
    enters the hot path (VPHO.py:112-173) plus the dataset fields the hot path reads
    (dexycb6.py:471-509): encodings, heat-maps, regressed MANO, local forces, camera, boxes."""
    rng = np.random.default_rng(seed)
    mano = mano or make_mano_model()
    objects = objects or make_object_tables()
    f32 = np.float32
    out: Dict[str, object] = {}
    out["encoding_hand"] = np.maximum(rng.normal(size=(bs, 1024)), 0).astype(f32)
    out["encoding_obj"] = np.maximum(rng.normal(size=(bs, 1024)), 0).astype(f32)
    out["pd_mano_pose"] = (rng.normal(size=(bs, 48)) * 0.3).astype(f32)
    out["pd_mano_shape"] = rng.normal(size=(bs, 10)).astype(f32)
    fx = rng.uniform(500, 700, bs)
    K = np.zeros((bs, 3, 3), np.float64)
    K[:, 0, 0] = fx
    K[:, 1, 1] = fx
    K[:, 0, 2] = 128 + rng.normal(size=bs) * 4
    K[:, 1, 2] = 128 + rng.normal(size=bs) * 4
    K[:, 2, 2] = 1
    out["cam_intr_crop_flip"] = K.astype(f32)
    root = np.stack([rng.normal(size=bs) * 0.02, rng.normal(size=bs) * 0.02, rng.uniform(0.4, 0.9, bs)], 1)
    is_right = rng.random(bs) < 0.7
    out["is_right"] = is_right
    out["is_grasped"] = rng.random(bs) < 0.7
    out["root_joint"] = root.astype(f32)
    rf = root.copy()
    rf[~is_right, 0] *= -1
    out["root_joint_flip"] = rf.astype(f32)
    obj_id = rng.integers(0, len(objects["names"]), bs)
    out["obj_id"] = obj_id.astype(np.int32)
    out["obj_name"] = [objects["names"][i] for i in obj_id]

    # hidden "true" hand pose -> approximate joints (rest skeleton rotated by the wrist) for the heat-maps
    sk = _rest_skeleton()
    j21 = np.zeros((21, 3))
    order = {"thumb": 1, "index": 5, "middle": 9, "ring": 13, "pinky": 17}
    for name, b in _FINGER_BASE.items():
        o = order[name]
        j21[o:o + 3] = sk["J"][b:b + 3]
        j21[o + 3] = sk["tips"][name]
    true_wrist = rng.normal(size=(bs, 3)) * 0.3
    R = _rodrigues_np(true_wrist)
    jc = np.einsum("bij,kj->bki", R, j21) + rf[:, None]
    uvw = np.einsum("bkj,bij->bki", jc, K)
    uv = uvw[..., :2] / uvw[..., 2:]
    lo, hi = uv.min(1), uv.max(1)
    c, half = (lo + hi) / 2, np.maximum((hi - lo).max(1, keepdims=True) * 0.6, 20.0)
    bbox_hand = np.concatenate([c - half, c + half], 1)
    out["bbox_hand"] = bbox_hand.astype(f32)
    uv_hm = (uv - bbox_hand[:, None, :2]) / (bbox_hand[:, None, 2:] - bbox_hand[:, None, :2]) * 64 - 0.5
    out["hm_hand"] = _gauss_heatmaps(uv_hm, 64, 2.0, rng, 0.05)

    # hidden "true" object pose close to the hand
    t_rel = rng.normal(size=(bs, 3)) * 0.02 + np.array([0.07, 0.0, 0.02])
    rot = _rodrigues_np(rng.normal(size=(bs, 3)) * 0.8)
    kp = objects["kpt3d"][obj_id].astype(np.float64)
    kc = np.einsum("bij,bkj->bki", rot, kp) + t_rel[:, None] + root[:, None]
    kc[~is_right, :, 0] *= -1
    uvw = np.einsum("bkj,bij->bki", kc, K)
    uv = uvw[..., :2] / uvw[..., 2:]
    lo, hi = uv.min(1), uv.max(1)
    c, half = (lo + hi) / 2, np.maximum((hi - lo).max(1, keepdims=True) * 0.6, 20.0)
    bbox_obj = np.concatenate([c - half, c + half], 1)
    out["bbox_obj_rect"] = bbox_obj.astype(f32)
    uv_hm = (uv - bbox_obj[:, None, :2]) / (bbox_obj[:, None, 2:] - bbox_obj[:, None, :2]) * 64 - 0.5
    out["hm_obj"] = _gauss_heatmaps(uv_hm, 64, 2.0, rng, 0.05)
    out["true_obj_rot"] = rot.astype(f32)
    out["true_obj_trans"] = t_rel.astype(f32)
    out["true_wrist"] = true_wrist.astype(f32)

    # local contact forces = get_local_force(random scale, softmaxed random weights)  (physics.py:546-557)
    scale = np.abs(rng.normal(size=(bs, 32)) * 0.5)
    wgt = rng.normal(size=(bs, 32, 8))
    wgt = np.exp(wgt) / np.exp(wgt).sum(-1, keepdims=True)        # fc_weight's Softmax
    wgt = np.exp(wgt) / np.exp(wgt).sum(-1, keepdims=True)        # softmax again inside get_local_force
    th = np.arange(8) * 2 * np.pi / 8
    anchor = np.stack([np.cos(th) * 0.8, np.sin(th) * 0.8, np.ones(8)], -1) / 8
    d = wgt @ anchor
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    out["force_local"] = (d * scale[..., None]).astype(f32)
    out["sample_num"] = sample_num
    return out


def make_eval_ground_truth(batch: Dict[str, object], head_mano, objects: Dict[str, object]):
    """Ground-truth fields `Trainer.evaluate` reads from the dataset (gt_joint, gt_hand_vert, gt_obj_rt, cam_intr;
    lib/engine/train_diff_hand_obj.py:236-257) for a synthetic batch: the hidden "true" wrist pose with flat fingers posed
    by the MANO layer, and the hidden object pose, both in the un-flipped camera frame.  `head_mano`: the product's
    `HeadMano` (the meshes are made by the same CUDA layer; this runs outside any timed region)."""
    import torch
    bs = len(batch["obj_name"])
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    T = lambda k: torch.from_numpy(np.asarray(batch[k])).to(dev)     # noqa: E731
    pose = torch.cat([T("true_wrist").float(), torch.zeros(bs, 45, device=dev)], 1)
    v, j = head_mano.get_hand_verts(pose=pose, shape=T("pd_mano_shape").float())
    is_right = T("is_right").bool()
    sign = torch.where(is_right, 1.0, -1.0)[:, None]
    root = T("root_joint").float()
    v, j = v.clone(), j.clone()
    v[..., 0] *= sign
    j[..., 0] *= sign
    gt_rt = torch.cat([T("true_obj_rot").double(), (T("true_obj_trans").double() + root.double())[..., None]], -1)
    K = T("cam_intr_crop_flip").float()
    return {"gt_joint": j + root[:, None], "gt_hand_vert": v + root[:, None], "gt_obj_rt": gt_rt, "cam_intr": K}


def make_metric_tables(objects: Dict[str, object]) -> Dict[str, object]:
    """Synthetic stand-ins for what `TesterObject` reads per object (lib/engine/test.py:196-232): YCB_MESHES[name]['bbox3d'
    / 'diameter'] and asset/2023_NIPS_DeepSimHO/assets_models_info.json -- box corners and diagonal of the sampled surface and
    a mix of symmetry classes (none / one discrete half-turn / a continuous axis / both)."""
    from .evaluation import symmetry_tables
    verts = np.asarray(objects["verts_sampled"], np.float64)
    N = verts.shape[0]
    lo, hi = verts.min(1), verts.max(1)
    corners = np.array([[(hi if (c >> a) & 1 else lo)[:, a] for a in range(3)] for c in range(8)])     # (8, 3, N)
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
    return {"bbox3d": np.transpose(corners, (2, 0, 1)).astype(np.float32), "diameter": diameter.astype(np.float32),
            "sym_R": sR, "sym_t": st, "sym_count": cnt, "model_info": infos}



# ---------------------------------------------------------------------------------------------------------------------------
# N1: synthetic weights / inputs of the modules that produce the hot path's inputs (VPHO.py:56-74,129-178)
# ---------------------------------------------------------------------------------------------------------------------------
PRODUCER_DIMS = {"C": 256, "heat_hid": 128, "Jh": 21, "Jo": 27, "enc_hid": 256, "roi": 32, "mano_layers": (1024, 512),
                 "d_model": 512, "ff": 2048, "phys_hid": 512}
# toy dimensions for the CPU emulator tests: same module tree, every width shrunk
PRODUCER_DIMS_TOY = {"C": 6, "heat_hid": 8, "Jh": 3, "Jo": 4, "enc_hid": 8, "roi": 16, "mano_layers": (12, 10),
                     "d_model": 16, "ff": 24, "phys_hid": 16}


def make_producer_state(seed: int = 0, dims: Dict[str, object] = PRODUCER_DIMS) -> Dict[str, np.ndarray]:
    """State dict (reference key layout: `vpho_net.state_dict()` minus backbone and denoisers) of head_hm_hand / head_hm_obj
    (HeadHeatmap2(256, 21|27, 128)), encoder_hand / encoder_obj (Encoder(277|283, 256)), head_mano (HeadMano(1024, [1024, 512])),
    cross_hand / cross_obj (CrossModule(8, 512)) and head_physics (HeadPhysics(512)), buffers included
    (`cross_*.pose_embedder.pe` cross_module.py:70-78, `head_physics.anchor` physics.py:692-698).  The reference's
    `init_weights` (VPHO.py:34-45: conv N(0, .001^2), linear N(0, .01^2)) makes every activation vanish, which would make a
    parity test vacuous; weights are drawn with fan-in scaling instead and the BatchNorm running statistics / affine terms are
    random, so that every layer carries O(1) signal.  Shapes and names are the reference's (`dims` shrinks them for the
    emulator tests)."""
    import torch
    rng = np.random.default_rng(900001 + seed)
    f32 = np.float32
    st: Dict[str, np.ndarray] = {}
    C, hh, eh, roi = dims["C"], dims["heat_hid"], dims["enc_hid"], dims["roi"]
    dm, ff, ph = dims["d_model"], dims["ff"], dims["phys_hid"]
    enc_dim = eh * (roi // 16) ** 2
    in_hw = roi // 4
    proj_dim = int(dm / (in_hw ** 2 / 32))           # cross_module.py:95

    def conv(name, co, ci, k, bias=True, gain=1.0):
        st[name + ".weight"] = (rng.normal(size=(co, ci, k, k)) * gain / math.sqrt(ci * k * k)).astype(f32)
        if bias:
            st[name + ".bias"] = (rng.normal(size=co) * 0.1).astype(f32)

    def bn(name, c):
        st[name + ".weight"] = rng.uniform(0.6, 1.4, size=c).astype(f32)
        st[name + ".bias"] = (rng.normal(size=c) * 0.1).astype(f32)
        st[name + ".running_mean"] = (rng.normal(size=c) * 0.2).astype(f32)
        st[name + ".running_var"] = rng.uniform(0.5, 1.5, size=c).astype(f32)
        st[name + ".num_batches_tracked"] = np.asarray(100, np.int64)

    def lin(name, co, ci, gain=1.0):
        st[name + ".weight"] = (rng.normal(size=(co, ci)) * gain / math.sqrt(ci)).astype(f32)
        st[name + ".bias"] = (rng.normal(size=co) * 0.1).astype(f32)

    for p, out in (("head_hm_hand", dims["Jh"]), ("head_hm_obj", dims["Jo"])):
        conv(p + ".conv_layers.0", hh, C, 3)
        conv(p + ".conv_layers.1", hh, hh, 3)
        bn(p + ".conv_layers.2", hh)
        st[p + ".deconv_layers.0.weight"] = (rng.normal(size=(hh, hh // 2, 4, 4)) / math.sqrt(hh * 4)).astype(f32)
        bn(p + ".deconv_layers.1", hh // 2)
        conv(p + ".final_layer", out, hh // 2, 1)
    for p, cin in (("encoder_hand", C + dims["Jh"]), ("encoder_obj", C + dims["Jo"])):
        conv(p + ".project", eh, cin, 1)
        for r in range(8):
            q = f"{p}.reg.{r}"
            bn(q + ".bn", eh)
            conv(q + ".conv1", eh // 2, eh, 1, gain=1.4)
            bn(q + ".bn1", eh // 2)
            conv(q + ".conv2", eh // 2, eh // 2, 3, gain=1.4)
            bn(q + ".bn2", eh // 2)
            conv(q + ".conv3", eh, eh // 2, 1, gain=0.5)
    h1, h2 = dims["mano_layers"]
    lin("head_mano.base_layer.0", h1, enc_dim, 1.4)
    lin("head_mano.base_layer.2", h2, h1, 1.4)
    lin("head_mano.fc_pose", 96, h2)
    lin("head_mano.fc_shape", 10, h2)
    pe = torch.zeros(5000, dm)
    position = torch.arange(0, 5000, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dm, 2).float() * (-math.log(10000.0) / dm))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    for p in ("cross_hand", "cross_obj"):
        conv(p + ".proj_hand", proj_dim, eh, 3)
        conv(p + ".proj_obj", proj_dim, eh, 3)
        lin(p + ".gravity_proj", dm, 63)
        st[p + ".pose_embedder.pe"] = pe.unsqueeze(0).transpose(0, 1).contiguous().numpy()
        a = p + ".attn.layers.0"
        st[a + ".self_attn.in_proj_weight"] = (rng.normal(size=(3 * dm, dm)) / math.sqrt(dm)).astype(f32)
        st[a + ".self_attn.in_proj_bias"] = (rng.normal(size=3 * dm) * 0.1).astype(f32)
        lin(a + ".self_attn.out_proj", dm, dm)
        lin(a + ".linear1", ff, dm, 1.4)
        lin(a + ".linear2", dm, ff)
        for n in ("norm1", "norm2"):
            st[f"{a}.{n}.weight"] = rng.uniform(0.6, 1.4, size=dm).astype(f32)
            st[f"{a}.{n}.bias"] = (rng.normal(size=dm) * 0.1).astype(f32)
    for q, out in (("fc_scale", 1), ("fc_weight", 8), ("fc_CoM", 3)):
        lin(f"head_physics.{q}.0", ph, dm, 1.4)
        lin(f"head_physics.{q}.2", out, ph)
    ang = torch.arange(0, 2 * torch.pi, 2 * torch.pi / 8)[:8]
    st["head_physics.anchor"] = (torch.stack([torch.cos(ang), torch.sin(ang), torch.ones_like(ang)], dim=-1) / 8).numpy()
    return st


def make_producer_inputs(bs: int, seed: int = 0, roi: int = 32, C: int = 256) -> Dict[str, np.ndarray]:
    """RoI-aligned FPN features (post-ReLU-like, >= 0) and the per-image scalars the producers read (VPHO.py:113-160)."""
    rng = np.random.default_rng(424242 + seed)
    f32 = np.float32

    def feat():
        return np.maximum(rng.normal(size=(bs, C, roi, roi)), 0).astype(f32)

    def boxes():
        c = rng.uniform(90, 166, size=(bs, 2))
        wh = rng.uniform(40, 90, size=(bs, 2))
        tight = np.concatenate([c - wh / 2, c + wh / 2], 1)
        side = wh.max(1, keepdims=True) * rng.uniform(1.0, 1.3, size=(bs, 1))
        rect = np.concatenate([c - side / 2, c + side / 2], 1)
        return tight.astype(f32), rect.astype(f32)
    bh, bhr = boxes()
    bo, bor = boxes()
    g = rng.normal(size=(bs, 1, 3))
    g = 9.8 * g / np.linalg.norm(g, axis=-1, keepdims=True)
    return {"hf_hr": feat(), "of_or_rect": feat(), "hf_hr_rect": feat(), "bbox_hand": bh, "bbox_hand_rect": bhr,
            "bbox_obj": bo, "bbox_obj_rect": bor, "is_right": rng.random(bs) < 0.7, "gravity": g.astype(f32)}
