"""GPU probe: throughput of K back-to-back batches, joined vs pipelined (predict(defer_join=True)) under different stream
priorities.  usage: python tools/pipeline_probe.py [steps]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from vpho_b200.vpho import VphoHotPath  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda", 0)
mano, anchors, objects, batch, prior_h, prior_o, st_h, st_o = bench.make_inputs(bench.BS, seed=0)
import numpy as np  # noqa: E402
batch["obj_id"] = np.asarray(batch["obj_id"], np.int32)
res = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in batch.items() if isinstance(v, np.ndarray)}
ph, po = prior_h.to(dev), prior_o.to(dev)


def run(name, defer, compute_prio=None, agg_prio=-1, mesh_prio=0):
    hp = VphoHotPath(mano, anchors, objects, st_h, st_o, sample_num=bench.S, sampling_steps=bench.STEPS_ODE, sample_T0=bench.T0,
                     topk_hand=bench.K_HAND, topk_obj=bench.K_OBJ)
    hp._agg_stream = [torch.cuda.Stream(device=dev, priority=agg_prio) for _ in range(2)]
    hp._side_stream2 = [torch.cuda.Stream(device=dev, priority=mesh_prio) for _ in range(2)]
    cs = torch.cuda.Stream(device=dev, priority=compute_prio) if compute_prio is not None else torch.cuda.current_stream()
    with torch.cuda.stream(cs):
        for _ in range(4):
            VphoHotPath.join(hp.predict(res, prior_hand=ph, prior_obj=po, defer_join=defer))
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        last = None
        for _ in range(K):
            last = None
            last = hp.predict(res, prior_hand=ph, prior_obj=po, defer_join=defer)
        VphoHotPath.join(last)
        t1.record()
        torch.cuda.synchronize()
        w = time.perf_counter() - w0
    print("%-46s %.4f ms/step (device)  %.4f ms/step (wall)" % (name, t0.elapsed_time(t1) / K, w * 1e3 / K), flush=True)


def run_ahead(name, compute_prio=None, agg_prio=None, mesh_prio=None):
    hp = VphoHotPath(mano, anchors, objects, st_h, st_o, sample_num=bench.S, sampling_steps=bench.STEPS_ODE, sample_T0=bench.T0,
                     topk_hand=bench.K_HAND, topk_obj=bench.K_OBJ)
    if agg_prio is not None:
        hp._agg_stream = [torch.cuda.Stream(device=dev, priority=agg_prio) for _ in range(2)]
        hp._side_stream2 = [torch.cuda.Stream(device=dev, priority=mesh_prio) for _ in range(2)]
    cs = torch.cuda.Stream(device=dev, priority=compute_prio) if compute_prio is not None else torch.cuda.current_stream()
    with torch.cuda.stream(cs):
        for _ in range(4):
            VphoHotPath.join(hp.predict(res, prior_hand=ph, prior_obj=po, defer_join=True))
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        last = tk = None
        for _ in range(K):
            nt = hp.predict_begin(res, prior_hand=ph, prior_obj=po, defer_join=True)
            if tk is not None:
                last = None
                last = hp.predict_end(tk)
            tk = nt
        last = None
        last = hp.predict_end(tk)
        VphoHotPath.join(last)
        t1.record()
        torch.cuda.synchronize()
        w = time.perf_counter() - w0
    print("%-46s %.4f ms/step (device)  %.4f ms/step (wall)" % (name, t0.elapsed_time(t1) / K, w * 1e3 / K), flush=True)


if len(sys.argv) > 2 and sys.argv[2] == "prio":
    for _ in range(2):
        run_ahead("one-ahead, compute -2, agg -1, mesh 0 (default)", -2)
        run_ahead("one-ahead, compute -2, agg -2, mesh 0", -2, -2, 0)
        run_ahead("one-ahead, compute -2, agg -3, mesh 0", -2, -3, 0)
        run_ahead("one-ahead, compute -2, agg -1, mesh -1", -2, -1, -1)
        run_ahead("one-ahead, compute -2, agg -2, mesh -2", -2, -2, -2)
    sys.exit(0)
run("joined", False)
run_ahead("one-ahead, compute -2 (agg -1, mesh 0)", -2)
run_ahead("one-ahead, default stream (agg -1, mesh 0)")
run_ahead("one-ahead, compute -3 (agg -2, mesh -1)", -3)
run("pipelined, default stream, agg -1, mesh 0", True)
run("pipelined, compute -2, agg -1, mesh 0", True, -2, -1, 0)
run("pipelined, compute -2, agg 0, mesh 0", True, -2, 0, 0)
run("pipelined, compute -3, agg -2, mesh -1", True, -3, -2, -1)
run("pipelined, default stream, agg 0, mesh 0", True, None, 0, 0)
run("joined again", False)
