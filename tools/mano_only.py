"""Runs the 6400-candidate MANO forward a few times (for ncu captures)."""
import sys

import torch

sys.path.insert(0, ".")
from vpho_b200 import synthetic as syn  # noqa: E402
from vpho_b200.head_mano import HeadMano  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6400
hm = HeadMano(syn.make_mano_model())
g = torch.Generator().manual_seed(0)
p, s = (torch.randn(n, 48, generator=g) * 0.4).cuda(), torch.randn(n, 10, generator=g).cuda()
for _ in range(4):
    v, j = hm.get_hand_verts(pose=p, shape=s)
torch.cuda.synchronize()
print("ok", float(v.abs().mean()))
