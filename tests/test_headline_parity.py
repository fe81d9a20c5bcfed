"""Parity AT THE HEADLINE CONFIGURATION (BASELINE.json configs[1]: batch 64 x sample_num 100 x 50 sampling points, top-k
30 / 10) with bench.py's own inputs and seeds (`bench.make_inputs(64, 0)`), in the regime bench.py runs ("random":
N(0, sigma(T0)^2) priors) and in the regime a trained sampler produces ("clustered").  Reference path:
lib/model/VPHO.py:236-304 (predict branch), lib/model/score_based_model.py:45-105, lib/model/aggregation.py:1167-1353.

  (a) both samplers in lock-step (`sample_pair`) vs `oracle_sample`: identical controller trajectory (nfev, accepted,
      rejected) and |x - x_oracle| <= 2e-5 max(1, |x|) on the finals and on all 50 output points;
  (b) `HOI_Aggregator` on the ORACLE's finals vs the oracle's `hoi_aggregate`: selections exact up to near-ties, values of
      every image with exact selections within a bar derived from the oracle's own float64 / +-1-ulp shadow runs
      (tests/sensitivity.py);
  (c) `VphoHotPath.predict` vs `oracle_predict` end to end, same derived bar, plus the north-star "final pose error within
      1e-3 mm" (MJE / MVE / ADD / ADD-S against the synthetic ground truth) on the images with exact selections.

The parity report of (b) and (c) is written to $VPHO_PARITY_REPORT_DIR (default gpurun_out/) as parity_bs64_*.json BEFORE
anything is asserted; the committed copies live under profiles/.  The same code runs at a toy shape on the CPU emulator
build in the `-m "not gpu"` suite.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cases
from oracle import vpho_oracle as O
from tests import parity
from tests import sensitivity as sens

HEADLINE = dict(bs=64, S=100, steps=50, kh=30, ko=10)
TOY = dict(bs=2, S=12, steps=5, kh=5, ko=4)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CACHE = {}


def _inputs(kind, cfg):
    """bench.py's batch / weights / prior draws; 'clustered' swaps the priors for tight pose clusters."""
    import bench
    from vpho_b200 import synthetic as syn
    bs, S = cfg["bs"], cfg["S"]
    if S == bench.S:
        mano, anchors, objects, batch, ph, po, st_h, st_o = bench.make_inputs(bs, 0)
    else:
        mano, anchors, objects = cases.assets()
        batch = syn.make_eval_batch(bs, seed=0, sample_num=S, mano=mano, objects=objects)
        st_h, st_o = syn.make_denoiser_state("mano_pose", 0), syn.make_denoiser_state("obj", 0)
        ph, po = cases.e2e_priors("random", bs, S, batch, 0)
    if kind == "clustered":
        ph, po = cases.e2e_priors("clustered", bs, S, batch, 0)
    return (mano, anchors, objects), batch, ph, po, st_h, st_o


def agg_kwargs(batch, final_mano, obj6d, cfg):
    T = lambda k: torch.from_numpy(np.asarray(batch[k]))  # noqa: E731
    return dict(cam_intrinsic=T("cam_intr_crop_flip").float(), root_joint_flip=T("root_joint_flip").float(),
                root_joint=T("root_joint").float(), is_right=T("is_right").bool(), force_local=T("force_local").float(),
                is_grasped=T("is_grasped").bool(), hand_pose_diff=final_mano[:, :48].clone(),
                hand_pose_regression=T("pd_mano_pose").float(), hand_shape=final_mano[:, 48:].clone(),
                hand_heatmap=T("hm_hand").float(), hand_bbox=T("bbox_hand").float(), hand_topk=cfg["kh"],
                obj_pose6d=obj6d.clone(), obj_heatmap=T("hm_obj").float(), obj_bbox=T("bbox_obj_rect").float(),
                obj_topk=cfg["ko"], obj_name=list(batch["obj_name"]))


def _oracle(kind, cfg):
    """oracle_predict + the reference's own rounding-noise floor of the aggregation stage; once per (regime, shape)."""
    key = (kind, cfg["bs"], cfg["S"])
    if key not in _CACHE:
        torch.set_num_threads(os.cpu_count() or 1)
        assets, batch, ph, po, st_h, st_o = _inputs(kind, cfg)
        mano, anchors, objects = assets
        ref = O.oracle_predict(batch, O.OracleDenoiser(st_h), O.OracleDenoiser(st_o), O.OracleMano(mano),
                               O.OracleObject(objects), O.OracleAnchors(anchors), init_x_hand=ph, init_x_obj=po,
                               sample_num=cfg["S"], sampling_steps=cfg["steps"], topk_hand=cfg["kh"], topk_obj=cfg["ko"],
                               with_inprocess=False)
        kw = agg_kwargs(batch, ref["diff_final_hand_mano"].reshape(-1, 58), ref["diff_final_obj_6d"], cfg)
        floor = sens.oracle_sensitivity(assets, kw, ref["_sel"], n_ulp=3, seed=0)
        _CACHE[key] = (ref, kw, floor)
    return _CACHE[key]


def _write_report(name, rep):
    d = os.environ.get("VPHO_PARITY_REPORT_DIR", os.path.join(ROOT, "gpurun_out"))
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name), "w") as f:
            json.dump(rep, f, indent=1, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o))
    except OSError:
        pass


def _floor_summary(floor):
    return {k: {"max_over_shadows": sens.stats(floor[k]), "f64_shadow": sens.stats(floor["_runs"][0][k]),
                "ulp_runs": [sens.stats(r[k]) for r in floor["_runs"][1:]]} for k in sens.KEYS}


def _pose_errors(assets, batch, hand_joint, hand_vert, obj6d):
    """Final pose error vs the synthetic ground truth, mm: MJE, MVE (TesterHand, lib/engine/test.py:657-679), ADD, ADD-S
    (TesterObject.criterion_ADD_REP, lib/engine/test.py:413-442)."""
    mano, anchors, objects = assets
    om, oo = O.OracleMano(mano), O.OracleObject(objects)
    T = lambda k: torch.from_numpy(np.asarray(batch[k]))  # noqa: E731
    bs = hand_joint.shape[0]
    gt_pose = torch.cat([T("true_wrist"), torch.zeros(bs, 45)], 1)
    gv, gj = om(gt_pose, T("pd_mano_shape"))
    mje, mve = O.hand_pose_error_mm(hand_joint.float(), gj, hand_vert.float(), gv)
    gt6d = torch.cat([T("true_obj_rot")[:, :2].reshape(bs, 6), T("true_obj_trans")], 1)
    add, adds = O.object_add_mm(oo, obj6d, gt6d, batch["obj_name"])
    return torch.stack([torch.as_tensor(x).reshape(bs).double() for x in (mje, mve, add, adds)], 1)


def _brief(rep):
    return {k: rep[k] for k in ("lists", "exact", "near_tie", "not_judged", "bad", "clean_images")}


def _check_aggregation(lib, dev, kind, cfg, tag):
    from vpho_b200.aggregation import Assets, HOI_Aggregator
    from vpho_b200.head_mano import HeadMano
    ref, kw, floor = _oracle(kind, cfg)
    assets, batch, *_ = _inputs(kind, cfg)
    mano, anchors, objects = assets
    agg = HOI_Aggregator(HeadMano(mano, lib=lib), Assets(anchors, objects, lib=lib), debug=True)
    out = agg(**{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in cases.clone_kw(kw).items()})
    if dev != "cpu":
        torch.cuda.synchronize()
    rep = parity.check_hoi_derived(out, agg.last_debug, ref["_sel"], floor, c=10.0)
    rep["oracle_noise_floor"] = _floor_summary(floor)
    rep["what"] = f"HOI_Aggregator on the oracle's finals, {cfg}, {kind} priors, bench seeds"
    _write_report(f"parity_{tag}_aggregate_{kind}.json", rep)
    print("parity report (b)", kind, _brief(rep))
    assert rep["bad"] == 0, rep["bad_lists"]
    assert not rep["violations"], rep["violations"][:5]
    assert rep["pose6d_candidate_err_clean"] == 0.0           # pure gathers of identical inputs
    assert rep["clean_images"] >= cfg["bs"] // 2
    return rep


def _check_predict(lib, dev, kind, cfg, tag):
    from vpho_b200.vpho import VphoHotPath, to_device
    ref, kw, floor = _oracle(kind, cfg)
    assets, batch, ph, po, st_h, st_o = _inputs(kind, cfg)
    mano, anchors, objects = assets
    hp = VphoHotPath(mano, anchors, objects, st_h, st_o, sample_num=cfg["S"], sampling_steps=cfg["steps"], topk_hand=cfg["kh"],
                     topk_obj=cfg["ko"], debug=True, lib=lib)
    pd = hp.predict(to_device(batch, dev), prior_hand=ph, prior_obj=po, with_inprocess=False)
    if dev != "cpu":
        torch.cuda.synchronize()
    for side in ("hand", "obj"):
        for k in ("nfev", "net_calls"):
            assert hp.last_info[side][k] == ref["_info"][side][k], (side, k)
    rel = lambda a, b: ((a.cpu().double() - b.double()).norm() / b.double().norm()).item()   # noqa: E731
    upstream = {"diff_final_hand_vert_rel": rel(pd["diff_final_hand_vert"], ref["diff_final_hand_vert"]),
                "diff_final_hand_joint_rel": rel(pd["diff_final_hand_joint"], ref["diff_final_hand_joint"]),
                "diff_final_obj_6d_abs": (pd["diff_final_obj_6d"].cpu() - ref["diff_final_obj_6d"]).abs().max().item()}
    rep = parity.check_hoi_derived(pd["_sel"], hp.hoi_aggregator.last_debug, ref["_sel"], floor, c=10.0)
    rep["upstream"] = upstream
    rep["oracle_noise_floor"] = _floor_summary(floor)
    rep["sampler"] = {"ours": hp.last_info, "oracle": ref["_info"]}
    # final pose error (north star: within 1e-3 mm of the oracle's) on the images whose selections all matched exactly;
    # the oracle's own float64 shadow gives the same difference for the reference against itself
    e_ours = _pose_errors(assets, batch, pd["agg_hand_joint"].cpu(), pd["agg_hand_vert"].cpu(), pd["agg_obj_6d"].cpu())
    e_ref = _pose_errors(assets, batch, ref["agg_hand_joint"], ref["agg_hand_vert"], ref["agg_obj_6d"])
    sh = floor["_f64"]
    e_f64 = _pose_errors(assets, batch, sh["hand_agg_joint"], sh["hand_agg_vert"], sh["obj_agg_6d"])
    cm = torch.tensor(rep["clean_mask"])
    d_ours, d_ref = (e_ours - e_ref).abs(), (e_f64 - e_ref).abs()
    rep["pose_error_mm"] = {"cols": ["MJE", "MVE", "ADD", "ADD-S"], "ours_minus_oracle_clean": sens.stats(d_ours[cm]),
                            "oracle_f64_minus_oracle_clean": sens.stats(d_ref[cm]), "ours_minus_oracle_all": sens.stats(d_ours)}
    rep["what"] = f"VphoHotPath.predict vs oracle_predict, {cfg}, {kind} priors, bench seeds"
    _write_report(f"parity_{tag}_predict_{kind}.json", rep)
    print("parity report (c)", kind, _brief(rep), upstream, rep["pose_error_mm"])
    assert upstream["diff_final_hand_vert_rel"] < 1e-5 and upstream["diff_final_hand_joint_rel"] < 1e-5
    assert upstream["diff_final_obj_6d_abs"] < 2e-5
    assert rep["bad"] == 0, rep["bad_lists"]
    assert not rep["violations"], rep["violations"][:5]
    assert rep["clean_images"] >= cfg["bs"] // 2
    if cm.any():
        tol = torch.clamp(10.0 * d_ref[cm], min=1e-3)          # 1e-3 mm, or 10x the reference's own FP32-vs-FP64 difference
        assert bool((d_ours[cm] <= tol).all()), (d_ours[cm] - tol).max().item()
    return rep


def test_headline_checks_on_the_emulator_at_a_toy_shape(emu_lib):
    _check_predict(emu_lib, "cpu", "clustered", TOY, "toy")


@pytest.mark.gpu
def test_headline_sampler_pair_matches_oracle(cuda_lib):
    """(a) at bs = 64: 6400 rows x 96 / x 9, 800 + 75 head-GEMM work items over 73 CTA pairs, every pipeline stage and both
    accumulator buffers cycling many times per CTA."""
    from vpho_b200.score_based_model import Denoiser, ScoreBasedModelAgent
    cfg = HEADLINE
    BS, S, STEPS = cfg["bs"], cfg["S"], cfg["steps"]
    assets, batch, ph, po, st_h, st_o = _inputs("random", cfg)
    enc_h = torch.from_numpy(np.asarray(batch["encoding_hand"])).float()
    enc_o = torch.from_numpy(np.asarray(batch["encoding_obj"])).float()
    den_h, den_o = Denoiser(st_h), Denoiser(st_o)
    agent = ScoreBasedModelAgent(sampling_steps=STEPS, sample_num=S)
    da = {"feat_unique": enc_h.cuda(), "n_rows": BS * S}
    db = {"feat_unique": enc_o.cuda(), "n_rows": BS * S}
    (xs_h, x_h, pend_h), (xs_o, x_o, pend_o) = agent.sample_pair(da, den_h, db, den_o, 0.65, prior_a=ph, prior_b=po)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(64):
        torch.cuda.synchronize()
        if all([pend_h.resolve(), pend_o.resolve()]):
            break
        pend_h.pair.advance(4, stream)
    else:
        raise AssertionError("pair sampler did not finish")
    torch.set_num_threads(os.cpu_count() or 1)
    for (xs, x, pend), st, enc, prior in (((xs_h, x_h, pend_h), st_h, enc_h, ph), ((xs_o, x_o, pend_o), st_o, enc_o, po)):
        feat = enc[:, None].repeat(1, S, 1).reshape(-1, 1024)
        xs2, x2, info = O.oracle_sample(O.OracleDenoiser(st), feat, 0.65, prior, STEPS)
        assert pend.info["status"] == 1 and info["status"] == 0
        assert pend.info["nfev"] == info["nfev"] and pend.info["net_calls"] == info["net_calls"]
        assert ((x.cpu() - x2).abs() <= 2e-5 * x2.abs().clamp(min=1)).all()
        assert ((xs.cpu() - xs2).abs() <= 2e-5 * xs2.abs().clamp(min=1)).all()
        del xs2


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["clustered", "random"])
def test_headline_aggregation_on_oracle_finals(cuda_lib, kind):
    """(b): identical inputs on both sides (the oracle's own sampler finals)."""
    _check_aggregation(None, "cuda", kind, HEADLINE, "bs64")


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["clustered", "random"])
def test_headline_predict_matches_oracle(cuda_lib, kind):
    """(c): the whole predict branch at the benchmarked shape, bench seeds."""
    _check_predict(None, "cuda", kind, HEADLINE, "bs64")
