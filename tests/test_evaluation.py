"""The on-device evaluation record (`vpho_b200.evaluation.EvalRecorder` <- Trainer.evaluate's metric step,
lib/engine/train_diff_hand_obj.py:224-269; TesterHand / TesterObject, lib/engine/test.py) against the oracle restatements
applied the way the reference applies them: postprocess (un-flip, add root; :578-602), then TesterHand on the aggregated /
first-candidate hand and TesterObject on the aggregated / first-candidate object pose (:454-514)."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle import object_metrics as OM
from oracle import vpho_oracle as O
from vpho_b200 import synthetic as syn
from vpho_b200.aggregation import Assets
from vpho_b200.evaluation import HAND_COLS, EvalRecorder, summarize
from vpho_b200.head_mano import HeadMano


def _fake_predict(bs, S, g, mano_layer, batch, dev):
    """A predict()-shaped dict without running the sampler: candidate hands / object poses scattered around the truth."""
    T = lambda k: torch.from_numpy(np.asarray(batch[k])).to(dev)   # noqa: E731
    pose = torch.cat([T("true_wrist").float(), torch.zeros(bs, 45, device=dev)], 1)
    shape = T("pd_mano_shape").float()
    cand = pose[:, None] + 0.1 * torch.randn(bs, S, 48, generator=g).to(dev)
    v, j = mano_layer.get_hand_verts(pose=cand.reshape(-1, 48), shape=shape[:, None].repeat(1, S, 1).reshape(-1, 10))
    av, aj = mano_layer.get_hand_verts(pose=cand.mean(1), shape=shape)
    gt6 = torch.cat([T("true_obj_rot")[:, :2].reshape(bs, 6).double(), T("true_obj_trans").double()], 1)
    o6 = gt6[:, None] + 0.05 * torch.randn(bs, S, 9, generator=g, dtype=torch.float64).to(dev)
    return {"agg_hand_joint": aj, "agg_hand_vert": av, "diff_final_hand_joint": j.reshape(bs, S, 21, 3),
            "diff_final_hand_vert": v.reshape(bs, S, 778, 3), "agg_obj_6d": o6.mean(1), "diff_final_obj_6d": o6}


def _check(lib, dev, bs, S):
    mano, anch, objs = cases.assets()
    tables = OM.synthetic_metric_tables(objs)
    assets, hm = Assets(anch, objs, lib=lib), HeadMano(mano, lib=lib)
    batch = syn.make_eval_batch(bs, seed=3, sample_num=S, mano=mano, objects=objs)
    gt = syn.make_eval_ground_truth(batch, hm, objs)
    gt = {k: v.to(dev) for k, v in gt.items()}
    g = torch.Generator().manual_seed(5)
    pd = _fake_predict(bs, S, g, hm, batch, dev)
    dbatch = {k: torch.from_numpy(np.asarray(batch[k])).to(dev) for k in ("root_joint", "is_right", "obj_id")}
    dbatch.update(gt)
    rec = EvalRecorder(assets, tables)
    row = rec(pd, dbatch).cpu()
    assert tuple(row.shape) == (bs, rec.width) and row.dtype == torch.float64 and rec.width == 2 * 25 + 2 * 17
    # ---- oracle side, following the reference's steps
    root, is_right = torch.from_numpy(batch["root_joint"]), torch.from_numpy(batch["is_right"])

    def post(x):          # __postprocess_hand_vert
        x = x.cpu().clone()
        idx = torch.arange(bs)[~is_right]
        x[idx, ..., 0] = -x[idx, ..., 0]
        return torch.einsum("b...i,bi->b...i", torch.ones_like(x), root) + x
    gj, gv = gt["gt_joint"].cpu(), gt["gt_hand_vert"].cpu()
    col = 0
    for pj, pv in ((pd["agg_hand_joint"], pd["agg_hand_vert"]), (pd["diff_final_hand_joint"][:, 0], pd["diff_final_hand_vert"][:, 0])):
        pj, pv = post(pj), post(pv)
        mje, mve = O.hand_pose_error_mm(pj, gj, pv, gv)
        pa = O.hand_pa_error_mm(pj, gj, pv, gv)
        ref = np.concatenate([np.stack([np.asarray(mje), pa[0], np.asarray(mve), pa[1]], 1), pa[2]], 1)
        assert np.abs(row[:, col:col + 25].numpy() - ref).max() < 1e-3        # mm
        col += 25
    o6 = torch.stack([pd["agg_obj_6d"], pd["diff_final_obj_6d"][:, 0]], 1).cpu()
    from pytorch3d.transforms.rotation_conversions import rotation_6d_to_matrix      # oracle/shims
    rt = torch.cat([rotation_6d_to_matrix(o6[..., :6]), (o6[..., 6:] + root[:, None].double())[..., None]], -1)
    ref_o = OM.object_metrics(tables, rt.numpy(), gt["gt_obj_rt"].cpu().numpy(), batch["obj_id"], gt["cam_intr"].cpu().numpy())
    ours_o = row[:, col:].numpy().reshape(bs, 2, 17)
    assert np.abs(ours_o[..., [0, 1, 3, 6]] - ref_o[..., [0, 1, 3, 6]]).max() < 1e-8
    assert (np.abs(ours_o[..., [2, 4, 5, 7]] - ref_o[..., [2, 4, 5, 7]]) / np.abs(ref_o[..., [2, 4, 5, 7]])).max() < 5e-6
    assert np.abs(ours_o[..., 8:14] - ref_o[..., 8:14]).max() <= 1.5 / 2048
    # the fused call agrees with the row assembled from the individual entry points; with the regression hand too
    un = rec.unfused(pd, dbatch).cpu()
    assert (row[:, :50] - un[:, :50]).abs().max() < 1e-4 and (row[:, 50:] - un[:, 50:]).abs().max() < 1e-9
    rec3 = EvalRecorder(assets, tables, with_regression=True)
    row3 = rec3(pd, dbatch, reg_vert=pd["agg_hand_vert"], reg_joint=pd["agg_hand_joint"]).cpu()
    assert rec3.width == 3 * 25 + 2 * 17 and torch.equal(row3[:, :50], row[:, :50]) and torch.equal(row3[:, 75:], row[:, 50:])
    assert torch.equal(row3[:, 50:75], row[:, :25])
    with pytest.raises(ValueError):
        rec3(pd, dbatch)
    s = summarize(row, rec.cols)
    assert abs(s["hand/agg_candidate/MJE"] - float(row[:, 0].mean())) < 1e-12 and len(s) == rec.width
    assert rec.cols[0] == "hand/agg_candidate/MJE" and rec.cols[25 + 3] == "hand/one_candidate/" + HAND_COLS[3]


def test_evaluation_record_emulated(emu_lib):
    _check(emu_lib, "cpu", 2, 3)


@pytest.mark.gpu
def test_evaluation_record_cuda(cuda_lib):
    _check(None, "cuda", 16, 20)
