"""Runs N steps of the hot path at the headline shape (for ncu captures; not a benchmark)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from vpho_b200.vpho import VphoHotPath  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mano, anchors, objects, batch, ph, po, st_h, st_o = bench.make_inputs(bench.BS, 0)
hp = VphoHotPath(mano, anchors, objects, st_h, st_o, sample_num=bench.S, sampling_steps=bench.STEPS_ODE, topk_hand=bench.K_HAND,
                 topk_obj=bench.K_OBJ)
batch["obj_id"] = np.asarray(batch["obj_id"], np.int32)
dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in batch.items() if isinstance(v, np.ndarray)}
ph, po = ph.cuda(), po.cuda()
for _ in range(n):
    pd = hp.predict(dev, prior_hand=ph, prior_obj=po)
torch.cuda.synchronize()
print("ok", hp.last_info)
