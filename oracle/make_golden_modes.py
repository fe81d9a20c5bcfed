"""TEST INFRASTRUCTURE -- mints tests/golden/agg_modes.npz by running the REFERENCE'S OWN `HandAggregator` /
`ObjectAggregator` classes (lib/model/aggregation.py, imported unmodified through oracle/reference_loader.py; build container
only) in every mode row N4 covers:

    python -m oracle.make_golden_modes

Declared rule, as for the hot-path fixtures: `Tensor.topk` canonicalised to (value descending, index ascending).
('heatmap' is minted with is_weight=False: with weights the reference's own shape assertion in average_quaternion fails for
more than one fused joint, transform_fn.py:113.)
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cases                                     # noqa: E402
from oracle.reference_loader import load_reference           # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASE = dict(bs=3, S=16, seed=1, K=6, Ko=4)


def canonical_topk(self, k, dim=-1, largest=True, sorted=True):
    order = torch.sort(-self if largest else self, dim=dim, stable=True)[1].narrow(dim, 0, k)
    return torch.return_types.topk((torch.gather(self, dim, order), order))


def reference_modes(ref) -> dict:
    """name -> tensor for every mode, from the reference's classes."""
    hm = ref.head_mano.HeadMano(in_dim=1024)
    ho = ref.head_object.HeadObject()
    hand = ref.aggregation.HandAggregator(hm.get_hand_verts)
    obj = ref.aggregation.ObjectAggregator(ho)
    kw, batch, _ = cases.aggregate_case(CASE["bs"], CASE["S"], CASE["seed"])
    K, Ko = CASE["K"], CASE["Ko"]
    hk = lambda: dict(pose=kw["hand_pose_diff"].clone(), shape=kw["hand_shape"].clone(), root_joint=kw["root_joint_flip"],   # noqa: E731
                      cam_intrinsic=kw["cam_intrinsic"], heatmap=kw["hand_heatmap"], bbox=kw["hand_bbox"], k=K,
                      pose_regression=kw["hand_pose_regression"].clone())
    ok = lambda: dict(pose6d=kw["obj_pose6d"].clone(), root_joint=kw["root_joint"], obj_name=list(kw["obj_name"]),            # noqa: E731
                      cam_intrinsic=kw["cam_intrinsic"], heatmap=kw["obj_heatmap"], bbox=kw["obj_bbox"], k=Ko,
                      is_right=kw["is_right"])
    out = {}
    _topk = torch.Tensor.topk
    torch.Tensor.topk = canonical_topk
    try:
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            runs = {
                "heatmap": hand(mode="heatmap", is_weight=False, **hk()),
                "cascade4": hand(mode="heatmap_cascade", is_weight=True, use_regression_as_candidate=True, **hk()),
                "nlevel2": hand(mode="heatmap_cascade_n_level", n_level=2, is_weight=True, use_regression_as_candidate=True, **hk()),
                "nlevel3u": hand(mode="heatmap_cascade_n_level", n_level=3, is_weight=False, use_regression_as_candidate=True, **hk()),
                "pt_pose": hand(mode="2D_pt_pose", **hk()),
                "pt_joint": hand(mode="2D_pt_joint", **hk()),
                "average_all": hand(mode="average_all", **hk()),
                "random": hand(mode="random", **hk()),
            }
            for name, r in runs.items():
                for key in ("agg_hand_mano", "agg_vert", "agg_joint", "topk"):
                    if isinstance(r.get(key), torch.Tensor):
                        out[f"hand_{name}_{key}"] = r[key].numpy()
            r = obj(mode="heatmap", **ok())
            out["obj_heatmap_agg_6d"], out["obj_heatmap_agg_obj_vert"] = r["agg_6d"].numpy(), r["agg_obj_vert"].numpy()
            for name in ("2D_pt_pose", "average_all", "random"):
                r = obj(mode=name, **ok())
                out[f"obj_{name}_agg_6d"], out[f"obj_{name}_agg_obj_vert"] = r["agg_6d"].numpy(), r["agg_obj_vert"].numpy()
            for w in (False, True):
                r = obj(mode="heatmap_cascade", is_weight=w, is_force_selection=False, **ok())
                out[f"obj_cascade_w{int(w)}_agg_6d"] = r["agg_6d"].numpy()
                out[f"obj_cascade_w{int(w)}_agg_obj_vert"] = r["agg_obj_vert"].numpy()
    finally:
        torch.Tensor.topk = _topk
    out["fp"] = np.float64(cases.fingerprint(kw["hand_pose_diff"], kw["obj_pose6d"], kw["hand_heatmap"]))
    return out


def main():
    mano, anch, objs = cases.assets()
    ref = load_reference(mano, anch, objs)
    out = reference_modes(ref)
    np.savez_compressed(os.path.join(OUT, "agg_modes.npz"), **CASE, **out)
    print("agg_modes.npz:", sorted(out))


if __name__ == "__main__":
    main()
