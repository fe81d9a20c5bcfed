"""Pins the travelling CPU oracle (oracle/vpho_oracle.py) against the golden vectors minted from the reference's OWN
files by oracle/make_golden.py (tests/golden/*.npz).  Runs anywhere (no /root/reference, no GPU).
Same machine class -> results agree to the last bit; a small tolerance absorbs a different host's BLAS kernels."""
import os

import numpy as np
import pytest
import torch

from oracle import cases
from oracle import vpho_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,tol", [("sampler_obj", 5e-6), ("sampler_obj_stiff", 2e-4), ("sampler_hand", 5e-6)])
def test_oracle_sampler_matches_reference_golden(name, tol):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    head, bs, S = str(g["head"]), int(g["bs"]), int(g["S"])
    st, enc, init = cases.sampler_case(head, bs, S, float(g["last_std"]), int(g["seed"]))
    assert abs(cases.fingerprint(enc, init) - float(g["fp"])) < 1e-6, "seeded input generators drifted"
    den = O.OracleDenoiser(st)
    feat = enc[:, None].repeat(1, S, 1).reshape(-1, 1024)
    ev = den({"feat": feat, "sampled_pose": init, "t": torch.ones(feat.shape[0], 1) * 0.31})
    assert np.abs(ev.numpy() - g["eval_t031"]).max() <= 1e-5 * np.abs(g["eval_t031"]).max()
    xs, x, info = O.oracle_sample(den, feat, 0.65, init, int(g["steps"]))
    assert info["net_calls"] == int(g["net_calls"])
    assert (np.abs(x.numpy() - g["x"]) <= tol * np.maximum(1, np.abs(g["x"]))).all()
    assert (np.abs(xs.numpy() - g["xs"]) <= tol * np.maximum(1, np.abs(g["xs"]))).all()


def test_oracle_mano_matches_reference_golden(assets):
    g = np.load(os.path.join(GOLD, "mano.npz"))
    v, j = O.OracleMano(assets["mano"])(torch.from_numpy(g["pose"]), torch.from_numpy(g["shape"]))
    assert np.abs(v.numpy() - g["verts"]).max() < 5e-7 and np.abs(j.numpy() - g["joints"]).max() < 5e-7


@pytest.mark.parametrize("name", ["aggregate_small", "aggregate_readme"])
def test_oracle_aggregation_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    mano, anch, objs = cases.assets()
    kw, _, _ = cases.aggregate_case(int(g["bs"]), int(g["S"]), int(g["seed"]))
    kw.update(hand_topk=int(g["Kh"]), obj_topk=int(g["Ko"]))
    assert abs(cases.fingerprint(kw["hand_pose_diff"], kw["obj_pose6d"], kw["hand_heatmap"]) - float(g["fp"])) < 1e-6
    r = O.hoi_aggregate(O.OracleMano(mano), O.OracleObject(objs), O.OracleAnchors(anch), **kw)
    assert np.array_equal(r["pose6d_candidate"].numpy(), g["pose6d_candidate"])
    for k, tol in (("obj_agg_6d", 1e-7), ("agg_obj_vert", 1e-6), ("hand_agg_mano", 2e-5), ("hand_agg_vert", 2e-6),
                   ("hand_agg_joint", 2e-6)):
        assert np.abs(r[k].numpy().astype(np.float64) - g[k]).max() <= tol, k


def test_canonical_topk_is_value_desc_index_asc():
    x = torch.tensor([[1.0, 3.0, 3.0, 0.5, 3.0, 1.0]])
    v, i = O.canonical_topk(x, 4)
    assert i.tolist() == [[1, 2, 4, 0]] and v.tolist() == [[3.0, 3.0, 3.0, 1.0]]


def test_quaternion_average_of_identical_inputs_is_the_input():
    q = torch.nn.functional.normalize(torch.randn(3, 1, 4, generator=torch.Generator().manual_seed(0)), dim=-1).repeat(1, 7, 1)
    out = O.average_quaternion(q)
    ref = q[:, 0] * torch.sign(q[:, 0, :1])
    assert (out - ref).abs().max().item() < 1e-6
