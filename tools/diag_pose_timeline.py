"""Timeline of CTA 0 of one tensor-core pose-encoder launch.  Needs VPHO_TC_TIMELINE=1 python -m vpho_b200.build --force
(the pose kernel stamps slots 1024.. of each role, the head GEMM slots 0..)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpho_b200 import capi, synthetic as syn  # noqa: E402
from vpho_b200.score_based_model import Denoiser  # noqa: E402

lib = capi.lib()
fn = lib.c.vpho_debug_tc_clocks
fn.restype = C.c_int
fn.argtypes = [C.c_int, C.c_void_p, C.c_int]
den = Denoiser(syn.make_denoiser_state("mano_pose", 0))
g = torch.Generator().manual_seed(0)
enc = torch.relu(torch.randn(64, 1024, generator=g)).cuda()
x = (torch.randn(6400, 96, generator=g) * 2.5).cuda()
data = {"feat_unique": enc, "sampled_pose": x, "t": torch.full((6400, 1), 0.3, device="cuda")}
for _ in range(3):
    den(data)
torch.cuda.synchronize()
fn(1, None, 0)
den(data)
torch.cuda.synchronize()
buf = np.zeros(3 * 2048, np.uint64)
fn(0, buf.ctypes.data, buf.size)
prod, mma, comp = (buf[i * 2048 + 1024:(i + 1) * 2048].astype(np.int64) for i in range(3))
t0 = prod[0]
# default build: GEMM 1 = 3 chunks of 32 k (3xTF32), GEMM 2 = 4 chunks of 64 k (3xFP16); VPHO_POSE_GEMM2=tf32: 8 chunks
n2 = 8 if os.environ.get("VPHO_POSE_GEMM2") == "tf32" else 4
print("producer: start 0; empty-wait done per chunk:", (prod[1:4 + n2] - t0).tolist(), "kernel end:", int(prod[4 + n2] - t0))
print("mma: x ready", int(mma[0] - t0), "gemm1 full-ok:", (mma[1:4] - t0).tolist(), "d1 committed", int(mma[4] - t0))
print("mma gemm2 (W ok, A ok) per chunk:", (mma[5:5 + 2 * n2] - t0).reshape(-1, 2).tolist(), "d2 committed", int(mma[5 + 2 * n2] - t0))
print("compute: d1 seen", int(comp[0] - t0), "restaged chunks:", (comp[1:1 + n2] - t0).tolist(), "d2 seen", int(comp[1 + n2] - t0),
      "done", int(comp[2 + n2] - t0))
