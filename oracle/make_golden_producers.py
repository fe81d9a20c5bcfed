"""TEST INFRASTRUCTURE -- mints tests/golden/producers.npz by running the REFERENCE'S OWN module classes and glue functions
(build container only; /root/reference must be present):

    python -m oracle.make_golden_producers

Classes imported unmodified: lib/model/head_inplane.py HeadHeatmap2, lib/model/encoding.py Encoder, lib/model/head_mano.py
HeadMano, lib/model/cross_module.py CrossModule, lib/model/physics.py HeadPhysics.  lib/model/VPHO.py cannot be imported
(it pulls timm and the FPN backbone), so the two glue functions the producers need -- `vpho_net.align_hm_to_bbox_rectangle`
and `flip_tensor_by_mask_index` / `flip_point3d_by_mask_index` (VPHO.py:333-364) -- are cut out of that file's syntax tree
and executed as they stand; the wiring between them (VPHO.py:129-176) is followed line by line in `reference_forward`.
Inputs and weights are regenerated from seeds by vpho_b200.synthetic; the fixture stores the outputs (heat-maps on a 4x4
stride plus float64 sums, everything else in full).
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cases                      # noqa: E402
from oracle.reference_loader import REFERENCE_ROOT, load_reference   # noqa: E402
from vpho_b200 import synthetic as syn        # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "producers.npz")
BS, STATE_SEED, INPUT_SEED = 3, 0, 1


def vpho_glue():
    """The reference's own glue functions, executed from the source text of lib/model/VPHO.py."""
    import torch.nn.functional as F
    src = open(os.path.join(REFERENCE_ROOT, "lib", "model", "VPHO.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "F": F}
    want = {"flip_tensor_by_mask_index", "flip_point3d_by_mask_index"}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            exec(compile(ast.Module([node], []), "VPHO.py", "exec"), ns)
        if isinstance(node, ast.ClassDef) and node.name == "vpho_net":
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == "align_hm_to_bbox_rectangle":
                    exec(compile(ast.Module([sub], []), "VPHO.py", "exec"), ns)
    return ns


def reference_modules(ref, st):
    sys.path.insert(0, REFERENCE_ROOT)
    import lib.model.cross_module as CM
    import lib.model.encoding as EN
    import lib.model.head_inplane as HI
    T = {k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}

    def sub(p):
        return {k[len(p) + 1:]: v for k, v in T.items() if k.startswith(p + ".")}
    m = types.SimpleNamespace()
    m.head_hm_hand = HI.HeadHeatmap2(256, 21, 128)
    m.head_hm_obj = HI.HeadHeatmap2(256, 27, 128)
    m.encoder_hand = EN.Encoder(256 + 21, 256, size_input_feature=(32, 32))
    m.encoder_obj = EN.Encoder(256 + 27, 256, size_input_feature=(32, 32))
    m.head_mano = ref.head_mano.HeadMano(in_dim=1024, is_output_contact=False)
    m.cross_hand = CM.CrossModule(8, 512)
    m.cross_obj = CM.CrossModule(8, 512)
    m.head_physics = ref.physics.HeadPhysics(hid_dim=512)
    for name in ("head_hm_hand", "head_hm_obj", "encoder_hand", "encoder_obj", "head_mano", "cross_hand", "cross_obj", "head_physics"):
        mod = getattr(m, name).eval()
        missing, unexpected = mod.load_state_dict(sub(name), strict=False)
        assert not unexpected and all(k.startswith("mano_layer.") for k in missing), (name, missing, unexpected)
    return m


@torch.no_grad()
def reference_forward(m, glue, inp):
    """lib/model/VPHO.py:129-176 with the reference's modules and glue."""
    import torch.nn.functional as F
    data = {k: torch.from_numpy(np.asarray(v)) for k, v in inp.items()}
    hf_hr, of_or_rect, hf_hr_rect = data["hf_hr"], data["of_or_rect"], data["hf_hr_rect"]
    self = types.SimpleNamespace(cfg=types.SimpleNamespace(heatmap_size=64))
    pd_hm_hand = m.head_hm_hand(hf_hr)
    pd_hm_obj = m.head_hm_obj(of_or_rect)
    pd_hm_hand_rect = glue["align_hm_to_bbox_rectangle"](self, pd_hm_hand, data["bbox_hand"], data["bbox_hand_rect"])
    pd_hm_obj_rect = glue["align_hm_to_bbox_rectangle"](self, pd_hm_obj, data["bbox_obj"], data["bbox_obj_rect"])
    of_or_rect = glue["flip_tensor_by_mask_index"](of_or_rect, is_flip=~data["is_right"])
    pd_hm_obj_rect_ori = glue["flip_tensor_by_mask_index"](pd_hm_obj_rect, is_flip=~data["is_right"])
    pd_hm_hand_rs = F.interpolate(pd_hm_hand_rect, size=hf_hr.shape[-2:], mode="bilinear", align_corners=False)
    pd_hm_obj_ori_rs = F.interpolate(pd_hm_obj_rect_ori, size=hf_hr.shape[-2:], mode="bilinear", align_corners=False)
    encoding_hand, enc_hand_ls = m.encoder_hand(torch.cat((hf_hr_rect, pd_hm_hand_rs), dim=1))
    encoding_obj, enc_obj_ls = m.encoder_obj(torch.cat((of_or_rect, pd_hm_obj_ori_rs), dim=1))
    pd_mano_pose, pd_mano_shape = m.head_mano(encoding_hand)
    gravity_flipped = glue["flip_point3d_by_mask_index"](data["gravity"], is_flip=~data["is_right"])
    enc_phy_hand, _, _ = m.cross_hand(enc_hand_ls[1], enc_obj_ls[1].detach(), gravity_flipped)
    _, enc_phy_obj, _ = m.cross_obj(enc_hand_ls[1].detach(), enc_obj_ls[1], gravity_flipped)
    pd_phy_dt = m.head_physics(enc_phy_hand, enc_phy_obj)
    return {"hand_heatmap": pd_hm_hand, "obj_heatmap": pd_hm_obj, "encoding_hand": encoding_hand, "encoding_obj": encoding_obj,
            "mano_pose": pd_mano_pose, "mano_shape": pd_mano_shape, "force_local": pd_phy_dt["force_local"],
            "force_scale": pd_phy_dt["scale"], "force_weight": pd_phy_dt["weight"], "CoM": pd_phy_dt["CoM"],
            "enc_phy_hand": enc_phy_hand, "enc_phy_obj": enc_phy_obj}


def pack(out):
    """What the fixture keeps of a result dict (shared with tests/test_producers.py)."""
    keep = {}
    for k, v in out.items():
        a = np.asarray(v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v)
        if k.endswith("heatmap"):
            keep[k + "_strided"] = a[:, :, ::4, ::4].copy()
            keep[k + "_sum"] = np.asarray([a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()])
        elif k in ("hm_hand_rs", "hm_obj_rs"):
            continue
        else:
            keep[k] = a
    return keep


def main():
    mano, anch, objs = cases.assets()
    ref = load_reference(mano, anch, objs)
    st = syn.make_producer_state(STATE_SEED)
    inp = syn.make_producer_inputs(BS, INPUT_SEED)
    out = reference_forward(reference_modules(ref, st), vpho_glue(), inp)
    np.savez_compressed(OUT, bs=BS, state_seed=STATE_SEED, input_seed=INPUT_SEED,
                        fp=cases.fingerprint(inp["hf_hr"], st["encoder_obj.reg.7.conv3.weight"]), **pack(out))
    print("written", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
