"""Cross-checks the tcgen05 3xTF32 head GEMM against the FP32-SIMT head GEMM and the CPU oracle; times both."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vpho_oracle as O  # noqa: E402
from vpho_b200 import synthetic as syn  # noqa: E402
from vpho_b200.score_based_model import Denoiser, ScoreBasedModelAgent  # noqa: E402


def mk(head, mode):
    os.environ["VPHO_HEAD_GEMM"] = {"simt": "simt", "tf32": "tf32", "tc_head_only": "tf32"}.get(mode, "tc")
    os.environ["VPHO_POSE_ENCODER"] = "simt" if mode in ("simt", "tc_head_only") else "tc"
    return Denoiser(syn.make_denoiser_state(head, 0))


for head, D in (("obj", 9), ("mano_pose", 96)):
    d_tc, d_simt = mk(head, "tc"), mk(head, "simt")
    od = O.OracleDenoiser(syn.make_denoiser_state(head, 0))
    for bs, S in ((2, 100), (64, 100), (3, 37)):
        g = torch.Generator().manual_seed(bs)
        enc = torch.relu(torch.randn(bs, 1024, generator=g))
        x = torch.randn(bs * S, D, generator=g) * 2.5
        feat = enc[:, None].repeat(1, S, 1).reshape(-1, 1024)
        for t in (0.65, 0.2):
            tt = torch.ones(bs * S, 1) * t
            data = {"feat_unique": enc.cuda(), "sampled_pose": x.cuda(), "t": tt.cuda()}
            a = d_tc(data).cpu()
            b = d_simt(data).cpu()
            line = f"{head} bs={bs} S={S} t={t}: tc-vs-simt rel {((a - b).norm() / b.norm()).item():.3e} max {(a - b).abs().max().item():.3e}"
            if bs <= 3:
                o = od({"feat": feat, "sampled_pose": x, "t": tt})
                line += f" | tc-vs-oracle {((a - o).norm() / o.norm()).item():.3e} simt-vs-oracle {((b - o).norm() / o.norm()).item():.3e}"
            print(line, flush=True)
    enc = torch.relu(torch.randn(64, 1024)).cuda()
    for name, den in (("tc (3xFP16 head)", d_tc), ("tf32 (3xTF32 head)", mk(head, "tf32")), ("simt", d_simt)):
        agent = ScoreBasedModelAgent(50, 100)
        data = {"feat_unique": enc, "n_rows": 6400}
        for _ in range(3):
            agent.sample(data, den, 0.65, return_inprocess=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            agent.sample(data, den, 0.65, return_inprocess=False)
        e1.record()
        torch.cuda.synchronize()
        print(f"{head} sample() with {name}: {e0.elapsed_time(e1) / 5:.3f} ms, net_calls {agent.last_info['net_calls']}", flush=True)
