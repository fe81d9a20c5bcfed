"""Live check of the oracle restatement against the reference's own files (build container only: skipped where
/root/reference is absent, e.g. on the GPU box).  Every output must be bit-identical, under the declared oracle rules
(exact-mode cdist, canonical top-k)."""
import copy

import numpy as np
import pytest
import torch

from oracle import cases
from oracle import vpho_oracle as O
from oracle.reference_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    mano, anch, objs = cases.assets()
    return load_reference(mano, anch, objs)


def test_sampler_bit_identical(ref):
    _, marg, _, _, _ = ref.sde.init_sde("ve")
    st, enc, init = cases.sampler_case("obj", 2, 10, 0.5, 7)
    den = ref.denoiser.BaseDenoiser(marg, head="obj").eval()
    den.load_state_dict({k: torch.from_numpy(v) for k, v in st.items()})
    agent = ref.sbm.ScoreBasedModelAgent()
    agent.prior_fn = lambda shape, T: init.clone()
    feat = enc[:, None].repeat(1, 10, 1).reshape(-1, 1024)
    xs, x = agent.sample({"feat": feat}, den, 0.65)
    xs2, x2, _ = O.oracle_sample(O.OracleDenoiser(st), feat, 0.65, init, ref.cfg.sampling_steps)
    assert torch.equal(xs, xs2) and torch.equal(x, x2)


def test_mano_and_aggregation_bit_identical(ref):
    mano, anch, objs = cases.assets()
    hm = ref.head_mano.HeadMano(in_dim=1024)
    g = torch.Generator().manual_seed(1)
    p, s = torch.randn(5, 48, generator=g) * 0.5, torch.randn(5, 10, generator=g)
    v, j = hm.get_hand_verts(pose=p, shape=s)
    v2, j2 = O.OracleMano(mano)(p, s)
    assert torch.equal(v, v2) and torch.equal(j, j2)
    agg = ref.aggregation.HOI_Aggregator(hm.get_hand_verts, ref.head_object.HeadObject(), ref.physics.HeadPhysics(hid_dim=512))
    kw, _, _ = cases.aggregate_case(2, 20, 9)
    kw.update(hand_topk=8, obj_topk=5)
    _cd, _tk = torch.cdist, torch.Tensor.topk

    def canonical(self, k, dim=-1, largest=True, sorted=True):
        order = torch.sort(-self, dim=dim, stable=True)[1].narrow(dim, 0, k)
        return torch.return_types.topk((torch.gather(self, dim, order), order))
    torch.cdist = lambda a, b, p=2.0, compute_mode=None: _cd(a, b, p=p, compute_mode="donot_use_mm_for_euclid_dist")
    torch.Tensor.topk = canonical
    try:
        with torch.no_grad():
            r = agg(**cases.clone_kw(kw))
    finally:
        torch.cdist, torch.Tensor.topk = _cd, _tk
    o = O.hoi_aggregate(O.OracleMano(mano), O.OracleObject(objs), O.OracleAnchors(anch), **cases.clone_kw(kw))
    for k in ("obj_agg_6d", "pose6d_candidate", "agg_obj_vert", "hand_agg_mano", "hand_agg_vert", "hand_agg_joint"):
        assert r[k].dtype == o[k].dtype and torch.equal(r[k], o[k]), k


def test_procrustes_alignment_bit_identical(ref):
    """oracle.rigid_align_AtoB vs the reference's lib/utils/transform_fn.py:43-66 (used by TesterHand's PA metrics)"""
    rng = np.random.default_rng(3)
    for n in (21, 778):
        for dt in (np.float32, np.float64):
            B = rng.normal(size=(n, 3)).astype(dt) * 0.05
            A = (B @ np.linalg.qr(rng.normal(size=(3, 3)))[0].astype(dt)) * dt(1.3) + rng.normal(size=(n, 3)).astype(dt) * dt(0.004)
            assert np.array_equal(O.rigid_align_AtoB(A.copy(), B.copy()), ref.transform_fn.rigid_align_AtoB(A.copy(), B.copy()))
