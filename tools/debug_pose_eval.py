"""Tensor-core vs strict-FP32 score evaluation, error pattern by row / column (debugging aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpho_b200 import synthetic as syn
from vpho_b200.score_based_model import Denoiser
for head, rows in (("obj", 400), ("mano_pose", 400), ("obj", 6400), ("mano_pose", 6400)):
    st = syn.make_denoiser_state(head, 0, last_std=0.05)
    den, ref = Denoiser(st), Denoiser(st, strict_fp32=True)
    g = torch.Generator().manual_seed(1)
    S = 100
    enc = torch.relu(torch.randn(rows // S, 1024, generator=g)).cuda()
    x = (torch.randn(rows, den.out_dim, generator=g) * 2.5).cuda()
    data = {"feat_unique": enc, "sampled_pose": x, "t": torch.full((rows, 1), 0.31, device="cuda")}
    a, b = den(data), ref(data)
    torch.cuda.synchronize()
    err = (a - b).abs()
    print(head, rows, "rel", ((a - b).norm() / b.norm()).item(), "max", err.max().item(), "nan", torch.isnan(a).sum().item())
    bad_rows = (err.max(dim=1).values > 1e-3 * b.abs().max()).nonzero().flatten()
    print("  bad rows:", bad_rows.numel(), bad_rows[:20].tolist(), "... mod 128:", sorted(set((bad_rows % 128).tolist()))[:40])
    bad_cols = (err.max(dim=0).values > 1e-3 * b.abs().max()).nonzero().flatten()
    print("  bad cols:", bad_cols.tolist())
