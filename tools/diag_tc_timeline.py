"""Timeline of CTA 0 of one tensor-core head GEMM launch (producer / MMA issuer / epilogue), from %globaltimer stamps.
Needs a library built with the stamps compiled in:  VPHO_TC_TIMELINE=1 python -m vpho_b200.build --force"""
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
prod, mma, epi = (buf[i * 2048:(i + 1) * 2048].astype(np.int64) for i in range(3))
t0 = min(v for v in np.concatenate([prod, mma, epi]) if v > 0)
half = os.environ.get("VPHO_HEAD_GEMM") != "tf32"
ch = 4 if half else 8
print("chunks per item", ch)
n_it = 11
mm = mma[:n_it * (2 + 2 * ch)].reshape(n_it, 2 + 2 * ch) - t0
wait_empty = (mm[:, 1] - mm[:, 0]).sum()
wait_full = (mm[:, 2::2] - np.concatenate([mm[:, 1:2], mm[:, 3:-1:2]], axis=1)).sum()
issue = (mm[:, 3::2] - mm[:, 2::2]).sum()
print("MMA lane over", n_it, "items: span", mm[-1, -1] - mm[0, 0], "ns; waiting tmem_empty", wait_empty, "ns; waiting full", wait_full,
      "ns; issuing", issue, "ns")
print("per item span:", (mm[:, -1] - mm[:, 0]).tolist())
print("per item wait tmem_empty:", (mm[:, 1] - mm[:, 0]).tolist())
for item in range(int(os.environ.get("ITEMS", "3"))):
    m = mma[item * (2 + 2 * ch):(item + 1) * (2 + 2 * ch)] - t0
    e = epi[(item // 2) * 10:(item // 2 + 1) * 10] - t0      # group 0 stamps even items only
    p = prod[item * 2 * ch:(item + 1) * 2 * ch] - t0
    print(f"item {item}")
    print("  producer (empty-wait done, issued) per chunk:", p.reshape(-1, 2).tolist())
    print("  mma: start", m[0], "tmem_empty ok", m[1], "per chunk (full ok, committed):", m[2:].reshape(-1, 2).tolist())
    print("  epilogue: start", e[0], "after bar1", e[1], "tmem_full ok", e[2], "fenced", e[3], "after each tcgen05.ld", e[4:8].tolist(),
          "after math+arrive", e[8], "after emit", e[9])
