"""Merge the per-test parity reports the -m gpu headline tests write (gpurun_out/parity_bs64_*.json) into
profiles/<round>_parity_bs64.json.  usage: python tools/make_parity_report.py [src_dir] [dst]"""
import glob
import json
import os
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_parity_bs64.json"
out = {"shape": "bs=64 x sample_num=100 x 50 ODE output points, topk_hand 30 / topk_obj 10, bench.py seeds (make_inputs(64, 0))",
       "how": "tests/test_headline_parity.py on a B200 (-m gpu); oracle = oracle/vpho_oracle.py in float32, noise floor = its float64 "
              "shadow + 3 runs on +-1-ulp perturbed inputs (tests/sensitivity.py); a list is 'exact' when the index lists are equal, "
              "'near_tie' when every disagreement lies inside the derived band, 'not_judged' when the list ranks candidates whose "
              "inputs already differ upstream (parent joints fused differently inside the reference's own noise), 'bad' otherwise",
       "cases": {}}
for f in sorted(glob.glob(os.path.join(src, "parity_bs64_*.json"))):
    d = json.load(open(f))
    d.pop("clean_mask", None)
    out["cases"][os.path.basename(f)[len("parity_bs64_"):-5]] = d
out["summary"] = {k: {q: v[q] for q in ("lists", "exact", "near_tie", "not_judged", "bad", "clean_images")} for k, v in out["cases"].items()}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out["summary"], indent=1))
