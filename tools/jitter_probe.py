"""GPU probe: per-batch times of the one-ahead pipelined loop, with / without the NVML clock-sampling thread, with the
allocator's device-allocation count.  usage: python tools/jitter_probe.py [steps]"""
import gc
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from vpho_b200.vpho import VphoHotPath  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-2))
mano, anchors, objects, batch, prior_h, prior_o, st_h, st_o = bench.make_inputs(bench.BS, seed=0)
batch["obj_id"] = np.asarray(batch["obj_id"], np.int32)
res = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in batch.items() if isinstance(v, np.ndarray)}
ph, po = prior_h.to(dev), prior_o.to(dev)
hp = VphoHotPath(mano, anchors, objects, st_h, st_o, sample_num=bench.S, sampling_steps=bench.STEPS_ODE, sample_T0=bench.T0,
                 topk_hand=bench.K_HAND, topk_obj=bench.K_OBJ)


def loop(name, warm=6):
    tk = None
    for _ in range(warm):
        nt = hp.predict_begin(res, prior_hand=ph, prior_obj=po)
        if tk is not None:
            hp.predict_end(tk)
        tk = nt
    VphoHotPath.join(hp.predict_end(tk))
    torch.cuda.synchronize()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    host = []
    a0 = torch.cuda.memory_stats()["num_device_alloc"]
    last = tk = None
    for i in range(K):
        marks[i].record()
        h0 = time.perf_counter()
        nt = hp.predict_begin(res, prior_hand=ph, prior_obj=po)
        h1 = time.perf_counter()
        if tk is not None:
            last = None
            last = hp.predict_end(tk)
        host.append((round((h1 - h0) * 1e3, 2), round((time.perf_counter() - h1) * 1e3, 2)))
        tk = nt
    last = None
    last = hp.predict_end(tk)
    VphoHotPath.join(last)
    marks[K].record()
    torch.cuda.synchronize()
    per = [round(marks[i].elapsed_time(marks[i + 1]), 2) for i in range(K)]
    print(name, "total %.4f ms/step; median %.3f; device allocs during loop %d; reserved %.2f GB" % (
        marks[0].elapsed_time(marks[K]) / K, float(np.median(per)), torch.cuda.memory_stats()["num_device_alloc"] - a0,
        torch.cuda.memory_reserved() / 2**30), flush=True)
    print("  per step", per, flush=True)
    print("  host (begin ms, end ms)", host[:12], flush=True)


loop("baseline")
loop("baseline again")
gc.disable()
loop("gc disabled")
gc.enable()
cs = bench.ClockSampler(0)
cs.start()
time.sleep(0.2)
loop("with NVML thread (20 ms)")
cs.stop_flag.set()
cs.join(timeout=2)
loop("after NVML thread stopped")
