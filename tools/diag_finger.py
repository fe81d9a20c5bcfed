import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_e2e import _run
hp, pd, ref, batch, assets = _run(None, "cuda", "clustered", 1, 100, 30, 10, 50, seed=11)
d = hp.hoi_aggregator.last_debug; od = ref["_sel"]["_dbg"]
ours_s, ref_s = d["finger_score"].cpu()[0], od["finger_score"][0]
ours_t, ref_t = d["finger_topk"].cpu()[0], od["finger_topk"][0]
torch.set_printoptions(precision=7, linewidth=200)
for f in range(5):
    print("finger", f, "ours topk", ours_t[f].tolist(), "ref topk", ref_t[f].tolist())
    print("  ref scores sorted:", torch.sort(ref_s[f], descending=True)[0][:8])
    print("  ours scores at ref order:", ours_s[f][torch.sort(ref_s[f], descending=True)[1][:8]])
    print("  max |ours-ref| / max|ref|:", ((ours_s[f] - ref_s[f]).abs().max() / ref_s[f].abs().max()).item())
print("cascade pose diff", (d["cascade_pose"].cpu() - od["cascade"]["agg_hand_mano"][:, :48]).abs().max().item())
print("obj fused diff", (pd["agg_obj_6d"].cpu() - ref["agg_obj_6d"]).abs().max().item())
print("hand cand pose diff", (od["hand_cand_pose"][0, :, :48] - 0).shape)
for lv in range(4):
    L = od["cascade"]["levels"][lv]; tk = L["topk"]; tk = tk[..., None] if tk.dim() == 2 else tk
    nf = tk.shape[-1]
    ot = d["hand_topk"][lv].cpu()[:, :nf].permute(0, 2, 1)
    print("level", lv, "topk equal", (ot == tk).float().mean().item())
