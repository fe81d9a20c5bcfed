"""Times `vpho_heads_forward` (N1: heat-map heads, encoders, regression head, cross modules, physics head) at the README batch
on one GPU: CUDA events on the launching stream, inputs (3 x 67 MB of RoI features) larger than half of L2 and re-read every
call, algorithmic FLOP from the layer shapes.  Usage: python tools/producers_bench.py [bs] [reps] [strict]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpho_b200 import capi, synthetic as syn  # noqa: E402
from vpho_b200.producers import FeatureHeads  # noqa: E402


def producers_flop(d, bs):
    """2 x multiply-adds of every dense layer (VPHO.py:129-176 at the given widths)."""
    roi, C, hh, eh, dm, ff, ph = d["roi"], d["C"], d["heat_hid"], d["enc_hid"], d["d_model"], d["ff"], d["phys_hid"]
    px = roi * roi
    f = 0
    for J in (d["Jh"], d["Jo"]):
        f += px * (9 * C * hh + 9 * hh * hh) + 4 * px * (4 * hh * (hh // 2)) + 4 * px * (hh // 2) * J          # heat-map head
        f += px * (C + J) * eh                                                                                   # encoder project
        p = px
        for _ in range(4):
            f += 2 * p * (eh * (eh // 2) + 9 * (eh // 2) ** 2 + (eh // 2) * eh)
            p //= 4
    h1, h2 = d["mano_layers"]
    enc_dim = eh * (roi // 16) ** 2
    f += enc_dim * h1 + h1 * h2 + h2 * 106
    proj = int(dm / ((roi // 4) ** 2 / 32))
    tok = 65
    per_cross = 2 * (roi // 4) ** 2 * 9 * eh * proj + tok * (3 * dm * dm + dm * dm + 2 * dm * ff) + 63 * dm + tok * 2 * bs * dm
    f += 2 * per_cross
    f += 32 * (3 * dm * ph + ph * 12)
    return 2 * f * bs


def main():
    bs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    strict = len(sys.argv) > 3 and sys.argv[3] == "strict"
    d = syn.PRODUCER_DIMS
    fh = FeatureHeads(syn.make_producer_state(0))
    inp = syn.make_producer_inputs(bs, 1)
    T = {k: torch.from_numpy(np.asarray(v)).cuda() for k, v in inp.items()}
    for _ in range(3):
        out = fh(T["hf_hr"], T["of_or_rect"], T["hf_hr_rect"], T, strict_fp32=strict)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    l0 = capi.lib().c.vpho_launch_count()
    ev[0].record()
    for i in range(reps):
        out = fh(T["hf_hr"], T["of_or_rect"], T["hf_hr_rect"], T, strict_fp32=strict)
        ev[i + 1].record()
    torch.cuda.synchronize()
    launches = (capi.lib().c.vpho_launch_count() - l0) // reps
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    med = ms[len(ms) // 2]
    fp32 = (capi.c_float * 1)()
    f16 = (capi.c_float * 1)()
    capi.lib().c.vpho_measure_peaks(fp32, f16, 3, None)
    flop = producers_flop(d, bs)
    peak = fp32[0] if strict else f16[0]
    print(json.dumps({"what": "vpho_heads_forward (N1 producers)", "path": "fp32 simt" if strict else "tcgen05 3xFP16", "bs": bs, "ms_median": round(med, 4), "ms_min": round(ms[0], 4),
                      "images_per_s": round(bs / med * 1e3, 1), "launches_per_call": int(launches), "gflop": round(flop / 1e9, 2),
                      "tflops": round(flop / med / 1e9, 2), "fp32_fma_peak_tflops_measured": round(fp32[0], 2),
                      "f16_umma_peak_tflops_measured": round(f16[0], 2),
                      "frac_of_peak": round(flop / med / 1e9 / peak, 4),
                      "peak": "measured FP32 FMA" if strict else "measured kind::f16 UMMA (each algorithmic FLOP is 3 UMMAs)",
                      "checksum": float(out["encoding_hand"].double().sum().item())}))


if __name__ == "__main__":
    main()
