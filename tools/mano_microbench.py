"""BASELINE config 3(a): MANO LBS microbench -- candidates in {6400..409600}, verts materialised / joints only.
Prints one JSON line per case with achieved HBM GB/s on the algorithmic 9 820 B/candidate."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpho_b200 import synthetic as syn  # noqa: E402
from vpho_b200.head_mano import HeadMano  # noqa: E402


def main():
    peaks = {}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    hbm = peaks.get("hbm_gbs", 6650.0)
    hm = HeadMano(syn.make_mano_model())
    for n in (6400, 25600, 102400, 409600):
        g = torch.Generator(device="cuda").manual_seed(n)
        pose = torch.randn(n, 48, device="cuda", generator=g) * 0.5
        shape = torch.randn(n, 10, device="cuda", generator=g)
        for need_verts in (True, False):
            for _ in range(3):
                hm.get_hand_verts(pose=pose, shape=shape, need_verts=need_verts)
            torch.cuda.synchronize()
            reps = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                hm.get_hand_verts(pose=pose, shape=shape, need_verts=need_verts)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            nbytes = n * (232 + (9336 if need_verts else 0) + 252)
            print(json.dumps({"case": "mano_lbs", "n": n, "verts": need_verts, "ms": round(ms, 4),
                              "GBps": round(nbytes / ms / 1e6, 1), "frac_hbm": round(nbytes / ms / 1e6 / hbm, 3),
                              "GFLOPs": round(n * 1.176e6 / ms / 1e6, 1), "note": "includes torch.empty of outputs"}))


if __name__ == "__main__":
    main()
