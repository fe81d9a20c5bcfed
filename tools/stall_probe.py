"""Where do the rare 10-100 ms stalls of the enqueueing thread come from?  Runs the pipelined resident loop for a few hundred
batches, times every sub-call of predict_begin and reports the slow ones together with the allocator's cudaMalloc count."""
import gc
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
log = []


def wrap(obj, name):
    f = getattr(obj, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        dt = (time.perf_counter() - t0) * 1e3
        if dt > 3.0:
            log.append((name, round(dt, 2), torch.cuda.memory_stats()["num_device_alloc"]))
        return r
    setattr(obj, name, g)


wrap(hp.score_agent, "sample_pair")
wrap(hp, "_snapshot")
wrap(hp, "_downstream")
wrap(hp, "postprocess_diffusion_hand")
wrap(hp.head_mano, "get_hand_verts")
s = torch.cuda.Stream(priority=-2)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
with torch.cuda.stream(s):
    for _ in range(5):
        hp.predict(res, prior_hand=ph, prior_obj=po)
    torch.cuda.synchronize()
    gc.collect()
    gc.disable()
    ticket = None
    slow = []
    a0 = torch.cuda.memory_stats()["num_device_alloc"]
    for i in range(n):
        t0 = time.perf_counter()
        nxt = hp.predict_begin(res, prior_hand=ph, prior_obj=po)
        t1 = time.perf_counter()
        if ticket is not None:
            hp.predict_end(ticket)
        t2 = time.perf_counter()
        if (t1 - t0) * 1e3 > 3.0 or (t2 - t1) * 1e3 > 6.0:
            slow.append((i, round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2), torch.cuda.memory_stats()["num_device_alloc"] - a0))
        ticket = nxt
        if i % 100 == 99:
            ms = torch.cuda.memory_stats()
            print(f"i={i}: allocated {torch.cuda.memory_allocated() >> 20} MB, reserved {torch.cuda.memory_reserved() >> 20} MB, "
                  f"active {ms['active_bytes.all.current'] >> 20} MB, inactive_split {ms['inactive_split_bytes.all.current'] >> 20} MB, "
                  f"cudaMallocs {ms['num_device_alloc'] - a0}")
    hp.predict_end(ticket)
    torch.cuda.synchronize()
print("batches", n, "cudaMalloc calls during the loop", torch.cuda.memory_stats()["num_device_alloc"] - a0,
      "reserved MB", torch.cuda.memory_reserved() >> 20)
print("slow iterations (i, begin ms, end ms, cudaMallocs so far):", slow)
print("slow sub-calls (name, ms, cudaMalloc count):", log)
