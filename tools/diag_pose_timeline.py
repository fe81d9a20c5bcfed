"""Timeline of CTA 0 of the LAST k_pose_tc launch of a full paired sampler pass at the headline shape (bs 64 x 100), i.e.
with the float64 RK stage combination in the stage-input phase, plus the stand-alone evaluation mode.
Needs the stamped twin:  python -m vpho_b200.build --timeline   (libvpho_b200_timeline.so; the product library has no stamps)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpho_b200 import build as vb, capi  # noqa: E402

capi.LIB_PATH = vb.LIB_TIMELINE
import bench  # noqa: E402
from vpho_b200.vpho import VphoHotPath  # noqa: E402

lib = capi.lib()
fn = lib.c.vpho_debug_tc_clocks
fn.restype = C.c_int
fn.argtypes = [C.c_int, C.c_void_p, C.c_int]
mano, anchors, objects, batch, prior_h, prior_o, st_h, st_o = bench.make_inputs(64, 0)
hp = VphoHotPath(mano, anchors, objects, st_h, st_o)
batch["obj_id"] = np.asarray(batch["obj_id"], np.int32)
res = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in batch.items() if isinstance(v, np.ndarray)}
ph, po = prior_h.cuda(), prior_o.cuda()
agent, S, bs = hp.score_agent, hp.sample_num, 64
da = {"feat_unique": res["encoding_hand"], "n_rows": bs * S}
db = {"feat_unique": res["encoding_obj"], "n_rows": bs * S}


def run():
    return agent.sample_pair(da, hp.denoiser_hand, db, hp.denoiser_obj, hp.sample_T0, prior_a=ph, prior_b=po)


def show(tag):
    buf = np.zeros(3 * 2048, np.uint64)
    fn(0, buf.ctypes.data, buf.size)
    prod, mma, comp = (buf[i * 2048 + 1024:(i + 1) * 2048].astype(np.int64) for i in range(3))
    t0 = prod[0]
    n1 = int(os.environ.get("NK1", "3"))
    n2 = 4
    print(f"== {tag} (ns after the PDL wait of CTA 0, thread 0)")
    print("producer: empty-wait done per chunk:", (prod[1:1 + n1 + n2] - t0).tolist(), "kernel end:", int(prod[1 + n1 + n2] - t0))
    print("mma gemm1 (x ok, W ok) per chunk:", (mma[0:2 * n1] - t0).reshape(-1, 2).tolist(), "d1 committed", int(mma[2 * n1] - t0))
    print("mma gemm2 (W ok, A ok) per chunk:", (mma[1 + 2 * n1:1 + 2 * n1 + 2 * n2] - t0).reshape(-1, 2).tolist(), "d2 committed",
          int(mma[1 + 2 * n1 + 2 * n2] - t0))
    print("compute: d1 seen", int(comp[0] - t0), "restaged chunks:", (comp[1:1 + n2] - t0).tolist(), "d2 seen", int(comp[1 + n2] - t0),
          "done", int(comp[2 + n2] - t0))
    ex = buf[2 * 2048 + 1100:2 * 2048 + 1114].astype(np.int64) - t0
    print("chunk 0 of thread 128: unit 0 stored", int(ex[12]), "unit 1 stored", int(ex[13]))
    print("fine stamps (thread 128): entry", int(ex[0]), "after pdl wait", int(ex[1]), "scalars seen", int(ex[2]), "x chunk arrived (not last)", int(ex[3]),
          "last x chunk arrived", int(ex[4]), "| restage: rowmax stored", int(ex[5]), "bar passed", int(ex[6]), "| epilogue: tmem ld done", int(ex[7]),
          "rowmax stored", int(ex[8]), "bar passed", int(ex[9]), "staged", int(ex[10]), "bar passed", int(ex[11]))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print("paired sampler (timeline build, stamps off): %.3f ms" % (e0.elapsed_time(e1) / 10))
fn(1, None, 0)
run()
torch.cuda.synchronize()
show("last pose call of a paired sampler pass (mode final: x = y)")
for st in (1, 3, 6):
    fn(100 + st, None, 0)
    run()
    torch.cuda.synchronize()
    show(f"last call of RK stage {st} in a paired sampler pass ({st} K slots enter)")
# a mid-integration stage: run the blocking single sampler of the hand only, whose last stage call is s = 6 of the last attempt
from vpho_b200.score_based_model import Denoiser  # noqa: E402
g = torch.Generator().manual_seed(0)
x = (torch.randn(6400, 96, generator=g) * 2.5).cuda()
data = {"feat_unique": res["encoding_hand"], "sampled_pose": x, "t": torch.full((6400, 1), 0.3, device="cuda")}
for _ in range(3):
    hp.denoiser_hand(data)
torch.cuda.synchronize()
fn(1, None, 0)
hp.denoiser_hand(data)
torch.cuda.synchronize()
show("stand-alone evaluation (hand only, 50 CTAs)")
