"""GPU probe: device time of the evaluation record (vpho_eval_record) and of its parts."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from vpho_b200 import synthetic as syn  # noqa: E402
from vpho_b200.evaluation import EvalRecorder, hand_metrics_mm, postprocess_hand_vert  # noqa: E402
from vpho_b200.aggregation import obj_6d_to_rt  # noqa: E402
from vpho_b200.vpho import VphoHotPath  # noqa: E402

dev = torch.device("cuda", 0)
mano, anchors, objects, batch, prior_h, prior_o, st_h, st_o = bench.make_inputs(bench.BS, seed=0)
batch["obj_id"] = np.asarray(batch["obj_id"], np.int32)
hp = VphoHotPath(mano, anchors, objects, st_h, st_o, sample_num=bench.S, sampling_steps=bench.STEPS_ODE, sample_T0=bench.T0,
                 topk_hand=bench.K_HAND, topk_obj=bench.K_OBJ)
rec = EvalRecorder(hp.assets, syn.make_metric_tables(objects))
gt = syn.make_eval_ground_truth(batch, hp.head_mano, objects)
res = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in batch.items() if isinstance(v, np.ndarray)}
res.update({k: v.to(dev) for k, v in gt.items()})
pd = hp.predict(res, prior_hand=prior_h.to(dev), prior_obj=prior_o.to(dev))
torch.cuda.synchronize()


def timeit(name, fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    print("%-40s %.4f ms" % (name, a.elapsed_time(b) / n), flush=True)


timeit("vpho_eval_record (fused)", lambda: rec(pd, res))
root, right = res["root_joint"], res["is_right"].bool()
j = postprocess_hand_vert(pd["agg_hand_joint"], root, right)
v = postprocess_hand_vert(pd["agg_hand_vert"], root, right)
timeit("hand metrics, 64 rows", lambda: hand_metrics_mm(j, res["gt_joint"], v, res["gt_hand_vert"], lib=hp.lib))
o6 = torch.stack([pd["agg_obj_6d"].double(), pd["diff_final_obj_6d"][:, 0].double()], dim=1)
rt = obj_6d_to_rt(o6, root.double())
timeit("object metrics, 64 x 2 poses", lambda: rec.obj_metrics(rt, res["gt_obj_rt"], res["obj_id"], res["cam_intr"]))
