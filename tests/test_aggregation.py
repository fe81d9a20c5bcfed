"""Scoring / top-k / aggregation parity: `HOI_Aggregator` over the CUDA kernels (C ABI `vpho_hoi_aggregate`) against
the oracle's `hoi_aggregate` (restating lib/model/aggregation.py:1167-1353) and against the golden outputs minted from
the reference's own files.  Tolerances: selections bit-exact up to near-ties (tests/parity.py); fused vertices / joints
2e-6 m (= 2e-3 mm, FP32 at 0.1-0.9 m camera depth); MANO parameters 2e-5 rad; object pose 1e-6."""
import os

import numpy as np
import pytest
import torch

from oracle import cases
from oracle import vpho_oracle as O
from tests import parity
from vpho_b200.aggregation import Assets, HOI_Aggregator, HeadObject, HeadPhysics
from vpho_b200.head_mano import HeadMano

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(lib, dev, bs, S, Kh, Ko, seed, spread=0.25):
    mano, anch, objs = cases.assets()
    kw, batch, _ = cases.aggregate_case(bs, S, seed, spread)
    kw.update(hand_topk=Kh, obj_topk=Ko)
    ref = O.hoi_aggregate(O.OracleMano(mano), O.OracleObject(objs), O.OracleAnchors(anch), **cases.clone_kw(kw))
    agg = HOI_Aggregator(HeadMano(mano, lib=lib), Assets(anch, objs, lib=lib), debug=True)
    out = agg(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()})
    return out, agg.last_debug, ref, kw


def _check_types(out, bs, S, Ko):
    assert out["obj_agg_6d"].dtype == torch.float64 and tuple(out["obj_agg_6d"].shape) == (bs, 9)
    assert out["pose6d_candidate"].dtype == torch.float64 and tuple(out["pose6d_candidate"].shape) == (bs, Ko * Ko, 9)
    assert out["agg_obj_vert"].dtype == torch.float32 and tuple(out["agg_obj_vert"].shape) == (bs, 2048, 3)
    assert tuple(out["hand_agg_mano"].shape) == (bs, 58) and tuple(out["hand_agg_vert"].shape) == (bs, 778, 3)
    assert tuple(out["hand_agg_joint"].shape) == (bs, 21, 3)


def test_aggregation_emulated_small(emu_lib):
    out, dbg, ref, _ = _run(emu_lib, "cpu", bs=3, S=16, Kh=6, Ko=4, seed=1)
    _check_types(out, 3, 16, 4)
    rep = parity.check_hoi_against_oracle(out, dbg, ref)
    assert rep["clean_images"] >= 2


def test_aggregation_emulated_matches_reference_golden(emu_lib):
    g = np.load(os.path.join(GOLD, "aggregate_small.npz"))
    out, dbg, ref, kw = _run(emu_lib, "cpu", int(g["bs"]), int(g["S"]), int(g["Kh"]), int(g["Ko"]), int(g["seed"]))
    assert abs(cases.fingerprint(kw["hand_pose_diff"], kw["obj_pose6d"], kw["hand_heatmap"]) - float(g["fp"])) < 1e-6
    for k, tol in (("hand_agg_vert", 2e-6), ("hand_agg_joint", 2e-6), ("agg_obj_vert", 2e-6), ("obj_agg_6d", 1e-6)):
        assert np.abs(out[k].cpu().numpy().astype(np.float64) - g[k]).max() <= tol, k


def test_object_points_and_force_anchors_emulated(emu_lib):
    mano, anch, objs = cases.assets()
    assets = Assets(anch, objs, lib=emu_lib)
    ho, oo = HeadObject(assets), O.OracleObject(objs)
    g = torch.Generator().manual_seed(0)
    pose = torch.randn(3, 5, 9, generator=g)
    names = [objs["names"][i] for i in (0, 7, 20)]
    is_right = torch.tensor([True, False, True])
    for dn in ("keypoint", "verts", "CoM"):
        a = ho(pose, names, data_name=dn)
        b = oo(pose, names, data_name=dn)
        assert (a - b).abs().max().item() < 1e-6
        a2 = ho(pose, names, data_name=dn, is_right=is_right)
        assert (a2 - oo.flip_pt3d(b.clone(), is_right)).abs().max().item() < 1e-6
    verts = torch.from_numpy(mano["v_template"])[None].repeat(2, 1, 1) + torch.tensor([0.02, -0.01, 0.6])
    fl = torch.randn(2, 32, 3, generator=g)
    fp, fg = HeadPhysics(assets).from_local_to_global(fl, verts)
    fp2, fg2 = O.OracleAnchors(anch).from_local_to_global(fl, verts)
    assert (fp - fp2).abs().max().item() < 1e-6 and (fg - fg2).abs().max().item() < 1e-4


@pytest.mark.gpu
def test_object_points_and_force_anchors_cuda(cuda_lib):
    mano, anch, objs = cases.assets()
    assets = Assets(anch, objs)
    ho, oo = HeadObject(assets), O.OracleObject(objs)
    g = torch.Generator().manual_seed(1)
    pose = torch.randn(4, 100, 9, generator=g)
    names = [objs["names"][i] for i in (3, 0, 20, 11)]
    is_right = torch.tensor([True, False, False, True])
    for dn in ("keypoint", "verts", "CoM"):
        b = oo(pose, names, data_name=dn)
        assert (ho(pose.cuda(), names, data_name=dn).cpu() - b).abs().max().item() < 2e-6
        a2 = ho(pose.cuda(), names, data_name=dn, is_right=is_right.cuda()).cpu()
        assert (a2 - oo.flip_pt3d(b.clone(), is_right)).abs().max().item() < 2e-6
    verts = torch.from_numpy(mano["v_template"])[None].repeat(5, 1, 1) + torch.tensor([0.02, -0.01, 0.6])
    fl = torch.randn(5, 32, 3, generator=g)
    fp, fg = HeadPhysics(assets).from_local_to_global(fl.cuda(), verts.cuda())
    fp2, fg2 = O.OracleAnchors(anch).from_local_to_global(fl, verts)
    assert (fp.cpu() - fp2).abs().max().item() < 1e-6 and (fg.cpu() - fg2).abs().max().item() < 1e-4


def _ties(lib, dev, bs, S, Kh, Ko):
    """All diffusion candidates of an image identical (and equal to the regression pose) -> every score of every list
    ties exactly; the canonical order (value desc, index asc) must then return index order 0..K-1 everywhere, and a
    partial tie (two distinct groups) must keep index order inside each group."""
    mano, anch, objs = cases.assets()
    kw, _, _ = cases.aggregate_case(bs, S, 3)
    pd = kw["hand_pose_diff"].reshape(bs, S, 48)
    kw["hand_pose_diff"] = pd[:, :1].repeat(1, S, 1).reshape(-1, 48)
    kw["hand_pose_regression"] = pd[:, 0].clone()
    kw["obj_pose6d"] = kw["obj_pose6d"][:, :1].repeat(1, S, 1)
    kw.update(hand_topk=Kh, obj_topk=Ko)
    agg = HOI_Aggregator(HeadMano(mano, lib=lib), Assets(anch, objs, lib=lib), debug=True)
    agg(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()})
    d = {k: v.cpu() for k, v in agg.last_debug.items()}
    for b in range(bs):
        for lv in range(4):
            for f in range(1 if lv == 0 else 5):
                assert d["hand_topk"][lv, b, f].tolist() == list(range(Kh)), (lv, b, f)
        assert d["obj_topk"][0, b, :Ko].tolist() == list(range(Ko)) and d["obj_topk"][1, b, :Ko].tolist() == list(range(Ko))
        assert d["obj_topk"][3, b, :5].tolist() == list(range(5))          # heat-map ranking of the K x K identical poses
        for f in range(5):
            # candidates 0..Kh-1 carry the identical top-k DIP parameters (exact ties); candidate Kh (the FUSED parameters,
            # a quaternion average that is not bit-identical to its inputs) may rank anywhere
            dup = [i for i in d["finger_topk"][b, f].tolist() if i != Kh]
            assert dup == list(range(len(dup))), (b, f, d["finger_topk"][b, f].tolist())
    # two groups: odd candidates get a second pose; whichever group scores higher comes first, in index order
    pd2 = pd[:, :1].repeat(1, S, 1)
    pd2[:, 1::2] = pd[:, 1:2]
    kw["hand_pose_diff"] = pd2.reshape(-1, 48)
    agg(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()})
    tk = agg.last_debug["hand_topk"].cpu()[0, :, 0, :Kh]
    sc = agg.last_debug["hand_score"].cpu()[0, :, :, 0]
    for b in range(bs):
        order = tk[b].tolist()
        vals = sc[b][tk[b].long()]
        assert bool((vals[:-1] >= vals[1:]).all())
        for i in range(len(order) - 1):
            if vals[i] == vals[i + 1]:
                assert order[i] < order[i + 1], (b, order)


def test_topk_ties_follow_index_order(emu_lib):
    _ties(emu_lib, "cpu", 1, 16, 6, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("bs,S,Kh,Ko", [(2, 16, 6, 4), (3, 100, 30, 10)])
def test_topk_ties_follow_index_order_cuda(cuda_lib, bs, S, Kh, Ko):
    _ties(None, "cuda", bs, S, Kh, Ko)


@pytest.mark.gpu
@pytest.mark.parametrize("bs,S,Kh,Ko,seed", [(3, 16, 6, 4, 1), (8, 100, 30, 10, 2), (5, 37, 12, 5, 3), (2, 200, 30, 10, 4),
                                               (1, 300, 64, 16, 5), (2, 5, 5, 3, 6)])
def test_aggregation_cuda(cuda_lib, bs, S, Kh, Ko, seed):
    out, dbg, ref, _ = _run(None, "cuda", bs, S, Kh, Ko, seed)
    _check_types(out, bs, S, Ko)
    rep = parity.check_hoi_against_oracle(out, dbg, ref)
    print("parity report", rep)
    assert rep["clean_images"] >= bs - 1 - bs // 4


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["aggregate_small", "aggregate_readme"])
def test_aggregation_cuda_matches_reference_golden(cuda_lib, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    out, dbg, ref, kw = _run(None, "cuda", int(g["bs"]), int(g["S"]), int(g["Kh"]), int(g["Ko"]), int(g["seed"]))
    assert abs(cases.fingerprint(kw["hand_pose_diff"], kw["obj_pose6d"], kw["hand_heatmap"]) - float(g["fp"])) < 1e-6
    rep = parity.check_hoi_against_oracle(out, dbg, ref)
    if rep["clean_images"] == int(g["bs"]):
        for k, tol in (("hand_agg_vert", 2e-6), ("hand_agg_joint", 2e-6), ("agg_obj_vert", 2e-6), ("obj_agg_6d", 1e-6)):
            assert np.abs(out[k].cpu().numpy().astype(np.float64) - g[k]).max() <= tol, k


@pytest.mark.gpu
def test_aggregation_cuda_rejects_bad_arguments(cuda_lib):
    from vpho_b200 import capi
    mano, anch, objs = cases.assets()
    kw, _, _ = cases.aggregate_case(1, 16, 3)
    kw.update(hand_topk=40, obj_topk=4)     # topk_hand > 2*S
    agg = HOI_Aggregator(HeadMano(mano), Assets(anch, objs))
    with pytest.raises(capi.VphoError):
        agg(**{k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()})
