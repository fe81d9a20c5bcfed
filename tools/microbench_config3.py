"""BASELINE config 3: MANO LBS + hand-object contact scoring microbench.
candidates in {6400, 25600, 102400, 409600} x object points in {2048, 4096, 8192}:
  (a) LBS, vertices materialised (HBM roofline: 9 820 algorithmic bytes per candidate);
  (b) LBS + 32 force anchors + anchor contact scoring (the reference's semantics, aggregation.py:553-590);
  (c) dense 778-vertex x object-point nearest distance (stress superset; 8 FLOP per pair, FP32 CUDA cores).
One JSON line per case (CUDA events, 3 warm-up + 5 timed launches, inputs larger than L2 for the big cases)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpho_b200 import synthetic as syn  # noqa: E402
from vpho_b200.aggregation import Assets, HeadPhysics, anchor_contact, vertex_contact  # noqa: E402
from vpho_b200.head_mano import HeadMano  # noqa: E402


def timed(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
        if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
    mano = syn.make_mano_model()
    anch, objs = syn.make_anchor_assets(mano), syn.make_object_tables()
    hm, phys = HeadMano(mano), HeadPhysics(Assets(anch, objs))
    Cn = 100
    for n in (6400, 25600, 102400, 409600):
        G = n // Cn
        g = torch.Generator(device="cuda").manual_seed(n)
        pose = torch.randn(n, 48, device="cuda", generator=g) * 0.4
        shape = torch.randn(n, 10, device="cuda", generator=g)
        root = torch.tensor([0.03, -0.02, 0.6], device="cuda")
        fl = torch.rand(G, Cn, 32, 3, device="cuda", generator=g) * 0.3
        verts = torch.empty((n, 778, 3), device="cuda")
        joints = torch.empty((n, 21, 3), device="cuda")

        def lbs():
            from vpho_b200 import capi
            hm.lib.check(hm.lib.c.vpho_mano_forward(hm.handle, capi.ptr(pose), capi.ptr(shape), n, capi.ptr(verts), capi.ptr(joints),
                                                    capi.stream_of(pose)), "mano")
        ms = timed(lbs)
        print(json.dumps({"case": "a_lbs_verts", "candidates": n, "ms": round(ms, 4), "GBps": round(n * 9820 / ms / 1e6, 1),
                          "frac_hbm": round(n * 9820 / ms / 1e6 / peaks["hbm_gbs"], 4), "TFLOPs": round(n * 1.176e6 / ms / 1e9, 2)}), flush=True)
        for P in (2048, 4096, 8192):
            obj = root + torch.tensor([0.07, 0.0, 0.02], device="cuda") + 0.05 * torch.randn(G, P, 3, device="cuda", generator=g)
            vc = verts.view(G, Cn, 778, 3)

            def fused():
                lbs()
                fp, fgl = phys.from_local_to_global(fl, vc)      # verts are wrist-centred here: geometry-only timing
                anchor_contact(fp, fgl, obj)
            ms_b = timed(fused)
            pairs_b = n * 32 * P
            print(json.dumps({"case": "b_lbs_anchor_contact", "candidates": n, "points": P, "ms": round(ms_b, 4),
                              "cand_per_s": round(n / ms_b * 1e3), "Gpairs_per_s": round(pairs_b / ms_b / 1e6, 1)}), flush=True)
            if n <= 25600:
                ms_c = timed(lambda: vertex_contact(vc, obj), reps=3)
                pairs_c = n * 778 * P
                print(json.dumps({"case": "c_dense_vertex_contact", "candidates": n, "points": P, "ms": round(ms_c, 3),
                                  "Gpairs_per_s": round(pairs_c / ms_c / 1e6, 1), "TFLOPs_8_per_pair": round(pairs_c * 8 / ms_c / 1e9, 2)}),
                      flush=True)


if __name__ == "__main__":
    main()
