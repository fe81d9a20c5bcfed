"""NOTE: this times the BLOCKING stand-alone `sample()` (a host read of the status word inside every call, attempts issued
in two rounds); the predict path uses the deferred, paired form -- see tools/graph_probe.py / bench.py for that.
Times one sample() of each denoiser at the README shape (bs 64 x 100 candidates, 50 output points) with CUDA
events and reports achieved FLOP/s on the factored-minimum FLOP count of SURVEY.md §8d."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpho_b200 import synthetic as syn  # noqa: E402
from vpho_b200.score_based_model import Denoiser, ScoreBasedModelAgent  # noqa: E402

FLOP = {"mano_pose": 4423680, "obj": 533504}


def main():
    bs, S = int(os.environ.get("BS", 64)), 100
    for head in ("mano_pose", "obj"):
        den = Denoiser(syn.make_denoiser_state(head, 0))
        g = torch.Generator().manual_seed(1)
        enc = torch.relu(torch.randn(bs, 1024, generator=g)).cuda()
        agent = ScoreBasedModelAgent(50, S)
        data = {"feat_unique": enc, "n_rows": bs * S}
        for xs_on in (True, False):
            for _ in range(3):
                agent.sample(data, den, 0.65, return_inprocess=xs_on)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                agent.sample(data, den, 0.65, return_inprocess=xs_on)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            calls = agent.last_info["net_calls"]
            fl = bs * S * FLOP[head] * calls
            print(json.dumps({"case": "sample", "head": head, "bs": bs, "xs": xs_on, "ms": round(ms, 3),
                              "net_calls": calls, "attempts": agent.last_info["attempts"],
                              "TFLOPs_factored": round(fl / ms / 1e9, 2), "cand_per_s": round(bs * S / ms * 1e3)}))


if __name__ == "__main__":
    main()
