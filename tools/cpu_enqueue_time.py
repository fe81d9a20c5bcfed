"""How long does the host take to ENQUEUE one step (no device sync inside)?  If this approaches the device time of a step,
the path is launch-bound and CUDA graphs are the next lever."""
import os
import sys
import time

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
enq, tot = [], []
for _ in range(20):
    t0 = time.perf_counter()
    pd, pend = hp._predict_once(res, ph, po, True)          # enqueue only
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    enq.append((t1 - t0) * 1e3)
    tot.append((t2 - t0) * 1e3)
    del pd
print(f"host enqueue of one step: median {np.median(enq):.2f} ms (min {min(enq):.2f}); enqueue + device: median {np.median(tot):.2f} ms")

# back-to-back steps without the per-step status read (what a one-step-lookahead evaluation loop achieves): the host runs
# ahead of the device, so the device never waits for Python between steps
for depth in (1, 2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    pending = []
    e0.record()
    for i in range(n):
        pending.append(hp._predict_once(res, ph, po, True))
        if len(pending) >= depth:
            pd, pend = pending.pop(0)
            st = torch.stack([p.counters for p in pend]).cpu()
            del pd
    while pending:
        pd, pend = pending.pop(0)
        st = torch.stack([p.counters for p in pend]).cpu()
        del pd
    e1.record()
    torch.cuda.synchronize()
    print(f"lookahead depth {depth}: {e0.elapsed_time(e1) / n:.3f} ms per step")
