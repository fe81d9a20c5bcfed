"""Accuracy of the 3xTF32 tcgen05 kernels when the pose-feature term dominates (conditioning weights zeroed, larger
pose-encoder weights): tensor-core path vs FP32-SIMT path vs the float64-evaluated network."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpho_b200 import synthetic as syn  # noqa: E402
from vpho_b200.score_based_model import Denoiser  # noqa: E402


def f64_network(st, x, t, feat):
    p = {k: torch.from_numpy(np.asarray(v)).double() for k, v in st.items()}
    tt = torch.full((x.shape[0],), float(np.float32(t)), dtype=torch.float64)
    xp = (tt[:, None].float() * p["t_encoder.0.W"].float()[None] * 2 * np.pi).double()
    four = torch.cat([torch.sin(xp), torch.cos(xp)], -1)
    tf = torch.relu(four @ p["t_encoder.1.weight"].T + p["t_encoder.1.bias"])
    h = torch.relu(x.double() @ p["pose_encoder.0.weight"].T + p["pose_encoder.0.bias"])
    pf = torch.relu(h @ p["pose_encoder.2.weight"].T + p["pose_encoder.2.bias"])
    tot = torch.cat([tf, pf, feat.double()], -1)
    y = torch.relu(torch.einsum("bc,ncd->bnd", tot, p["head.head.0.weight"]) + p["head.head.0.bias"])
    y = torch.einsum("bnc,ncd->bnd", y, p["head.head.2.weight"]) + p["head.head.2.bias"]
    std = float(np.float32(0.01) * np.float32(5000.0 ** float(np.float32(t)))) + 1e-7
    return y.reshape(x.shape[0], -1) / std


for scale in (0.01, 0.3):
    st = syn.make_denoiser_state("mano_pose", 0)
    rng = np.random.default_rng(5)
    st["pose_encoder.0.weight"] = (rng.normal(size=(256, 96)) * scale).astype(np.float32)
    st["pose_encoder.2.weight"] = (rng.normal(size=(256, 256)) * scale).astype(np.float32)
    if scale > 0.1:
        st["head.head.0.weight"] = st["head.head.0.weight"].copy()
        st["head.head.0.weight"][:, 384:, :] = 0          # conditioning slice off: the K = 256 pose term is everything
    os.environ["VPHO_HEAD_GEMM"], os.environ["VPHO_POSE_ENCODER"] = "simt", "simt"
    d_simt = Denoiser(st)
    os.environ.pop("VPHO_POSE_ENCODER")
    os.environ["VPHO_HEAD_GEMM"] = "tf32"
    d_tf32 = Denoiser(st)
    os.environ.pop("VPHO_HEAD_GEMM")
    d_tc = Denoiser(st)
    g = torch.Generator().manual_seed(0)
    enc = torch.relu(torch.randn(4, 1024, generator=g))
    x = torch.randn(400, 96, generator=g) * 2.5
    feat = enc[:, None].repeat(1, 100, 1).reshape(-1, 1024)
    for t in (0.65, 0.1):
        data = {"feat_unique": enc.cuda(), "sampled_pose": x.cuda(), "t": torch.full((400, 1), t).cuda()}
        a, b, c3 = d_tc(data).cpu().double(), d_simt(data).cpu().double(), d_tf32(data).cpu().double()
        r = f64_network(st, x, t, feat)
        n = r.norm()
        print(f"pose-weight scale {scale} t={t}: fp16x3-vs-f64 {((a - r).norm() / n).item():.3e}  tf32x3-vs-f64 {((c3 - r).norm() / n).item():.3e}  "
              f"simt-vs-f64 {((b - r).norm() / n).item():.3e}", flush=True)
