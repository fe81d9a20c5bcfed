"""TEST INFRASTRUCTURE -- imports the reference's OWN hot-path files, unmodified, from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
`oracle/make_golden.py` to mint the committed fixtures under tests/golden/ and by
tests/test_oracle_vs_reference.py to validate the travelling restatement `oracle/vpho_oracle.py`.
Nothing in the product package, bench.py's timed arm or the `-m gpu` tests may import this module.

What is shimmed: `ipdb`, `pytorch3d.transforms`, `pytorch3d.ops.knn`, `manopth.manolayer`
(oracle/shims/*), a fake `lib.dataset.base` exposing `YCB_MESHES`, `sys.argv` for lib/configs/args.py, and
a temporary CWD holding synthetic `asset/ours/vert2joint.pkl` + `asset/2021_CVPR_CPF/anchor/*`.
Files executed verbatim from the reference: lib/model/{sde,parallel_linear,denoiser,score_based_model,
aggregation,head_mano,head_object,physics}.py, lib/utils/{hand_fn,physics_fn,transform_fn}.py, lib/configs/args.py,
lib/engine/test.py (TesterObject, with a synthetic asset/2023_NIPS_DeepSimHO/assets_models_info.json).
"""
from __future__ import annotations

import os
import pickle
import sys
import tempfile
import types
from types import SimpleNamespace

import numpy as np

REFERENCE_ROOT = "/root/reference"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_loaded = None


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "model"))


def load_reference(mano: dict, anchors: dict, objects: dict, *, sample_num=100, topk_hand=30, topk_obj=10,
                   sampling_steps=50, sample_T0=0.65) -> SimpleNamespace:
    """Import the reference modules once per process (module-level singletons read assets at import)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("/root/reference is not present on this machine")
    argv = sys.argv
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="vpho_oracle_assets_")
    try:
        sys.argv = ["oracle", "--sample_num", str(sample_num), "--topk_hand", str(topk_hand), "--topk_obj",
                    str(topk_obj), "--sampling_steps", str(sampling_steps), "--sample_T0", str(sample_T0)]
        for p in (_SHIMS, REFERENCE_ROOT):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.makedirs(os.path.join(tmp, "asset", "ours"))
        adir = os.path.join(tmp, "asset", "2021_CVPR_CPF", "anchor")
        os.makedirs(adir)
        with open(os.path.join(tmp, "asset", "ours", "vert2joint.pkl"), "wb") as f:
            pickle.dump({"vert2joint": np.asarray(anchors["vert2joint"], np.float32)}, f)
        np.savetxt(os.path.join(adir, "face_vertex_idx.txt"), anchors["face_vertex_idx"], fmt="%d")
        np.savetxt(os.path.join(adir, "anchor_weight.txt"), np.asarray(anchors["anchor_weight"], np.float64), fmt="%.9e")
        np.savetxt(os.path.join(adir, "merged_vertex_assignment.txt"), np.zeros(778, np.int32), fmt="%d")
        with open(os.path.join(adir, "anchor_mapping_path.pkl"), "wb") as f:
            pickle.dump({}, f)
        os.chdir(tmp)

        # tables TesterObject reads (lib/engine/test.py:196-232): box corners, diameter, symmetry info (synthetic)
        import json
        from oracle.object_metrics import synthetic_metric_tables
        mt = synthetic_metric_tables(objects)
        os.makedirs(os.path.join(tmp, "asset", "2023_NIPS_DeepSimHO"))
        with open(os.path.join(tmp, "asset", "2023_NIPS_DeepSimHO", "assets_models_info.json"), "w") as f:
            json.dump({str(i + 1): mi for i, mi in enumerate(mt["model_info"])}, f)
        base = types.ModuleType("lib.dataset.base")
        ycb = {}
        for i, n in enumerate(objects["names"]):
            ycb[n] = {"kpt3d": np.asarray(objects["kpt3d"][i]), "shift": np.eye(4)[:3],
                      "verts_sampled": np.asarray(objects["verts_sampled"][i]),
                      "CoM": np.asarray(objects["CoM"][i]), "verts": np.asarray(objects["verts_sampled"][i]),
                      "bbox3d": np.asarray(mt["bbox3d"][i]), "diameter": float(mt["diameter"][i])}
        base.YCB_MESHES = ycb
        base.YCB_CLASSES = {i + 1: n for i, n in enumerate(objects["names"])}
        base.YCB_ID = {n: i + 1 for i, n in enumerate(objects["names"])}
        sys.modules["lib.dataset.base"] = base

        import manopth.manolayer as shim_mano  # noqa: E402  (oracle/shims)
        shim_mano.set_model(mano)

        import lib.model.sde as ref_sde
        import lib.model.denoiser as ref_denoiser
        import lib.model.score_based_model as ref_sbm
        import lib.model.head_mano as ref_head_mano
        import lib.model.head_object as ref_head_object
        import lib.model.physics as ref_physics
        import lib.model.aggregation as ref_aggregation
        import lib.utils.transform_fn as ref_transform_fn
        import lib.utils.hand_fn as ref_hand_fn
        import lib.utils.physics_fn as ref_physics_fn
        from lib.configs.args import cfg
        import lib.engine.test as ref_test                # TesterObject / TesterHand (reads the json above at construction)
        tester_object = ref_test.TesterObject()
    finally:
        os.chdir(cwd)
        sys.argv = argv
    _loaded = SimpleNamespace(sde=ref_sde, denoiser=ref_denoiser, sbm=ref_sbm, head_mano=ref_head_mano,
                              head_object=ref_head_object, physics=ref_physics, aggregation=ref_aggregation,
                              transform_fn=ref_transform_fn, hand_fn=ref_hand_fn, physics_fn=ref_physics_fn,
                              cfg=cfg, asset_dir=tmp, test=ref_test, tester_object=tester_object, metric_tables=mt)
    return _loaded
