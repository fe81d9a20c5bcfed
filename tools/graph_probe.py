"""Does replaying the paired sampler from a CUDA graph beat enqueuing its ~100 launches one by one?"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vpho_b200.vpho import VphoHotPath  # noqa: E402

mano, anchors, objects, batch, prior_h, prior_o, st_h, st_o = bench.make_inputs(64, 0)
hp = VphoHotPath(mano, anchors, objects, st_h, st_o)
batch["obj_id"] = np.asarray(batch["obj_id"], np.int32)
res = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in batch.items() if isinstance(v, np.ndarray)}
ph, po = prior_h.cuda(), prior_o.cuda()
for _ in range(5):
    hp.predict(res, prior_hand=ph, prior_obj=po)
torch.cuda.synchronize()
agent, S, bs = hp.score_agent, hp.sample_num, 64
da = {"feat_unique": res["encoding_hand"], "n_rows": bs * S}
db = {"feat_unique": res["encoding_obj"], "n_rows": bs * S}


def run():
    return agent.sample_pair(da, hp.denoiser_hand, db, hp.denoiser_obj, hp.sample_T0, prior_a=ph, prior_b=po)


def timeit(fn, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for _ in range(3):
    run()
print(f"paired sampler, stream launches: {timeit(run):.3f} ms")
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        run()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = run()
torch.cuda.synchronize()
print(f"paired sampler, graph replay:    {timeit(g.replay):.3f} ms")
ref = run()
torch.cuda.synchronize()
g.replay()
torch.cuda.synchronize()
print("graph output equals stream output:", torch.equal(out[0][1], ref[0][1]), torch.equal(out[1][1], ref[1][1]))
