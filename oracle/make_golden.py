"""TEST INFRASTRUCTURE -- mints tests/golden/*.npz by running the REFERENCE'S OWN FILES (unmodified, imported from
/root/reference through oracle/reference_loader.py) on seeded inputs.  Run in the build container only:

    python -m oracle.make_golden

The reference ships no golden vectors for this path; these fixtures are what pins the travelling oracle
(oracle/vpho_oracle.py) and, through it, the CUDA path.  Declared oracle rules applied to the reference run:
  (0)  torch.cdist forced to compute_mode='donot_use_mm_for_euclid_dist' (Appendix A.4);
  (ii) Tensor.topk replaced by the canonical (value desc, index asc) selection -- torch.topk's tie order is unspecified
       and the cascade always contains exact ties (duplicate regression candidates, zero heat outside the map), so the
       unpatched reference is not reproducible; how many fixture outputs change without the patch is printed;
  the sampler's prior draw is monkey-patched to the seeded tensor.
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases  # noqa: E402
from oracle.reference_loader import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def object_metrics_fixture(ref, objs):
    """TesterObject's own criteria (lib/engine/test.py:354-521), called per image exactly as TesterObject.__call__ does
    (:250-279), on seeded poses.  `.cuda()` is a no-op here (no GPU in the build container) and torch.cdist runs in the
    exact mode (oracle rule 0)."""
    inp = cases.object_metric_case()
    T = ref.tester_object
    names = [objs["names"][i] for i in inp["obj_id"]]
    _cuda, _cd = torch.Tensor.cuda, torch.cdist
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cdist = lambda a, b, p=2.0, compute_mode=None: _cd(a, b, p=p, compute_mode="donot_use_mm_for_euclid_dist")
    rows = []
    try:
        for i, nm in enumerate(names):
            pd, gt, K = inp["pd_rt"][i], inp["gt_rt"][i], inp["cam_intr"][i]
            mce, oce = T.criterion_MCE_OCE(pd, gt, nm)
            smce = T.criterion_SMCE(pd, gt, nm)
            mce2 = np.array([T.criterion_MCE2(pd[c], gt, nm) for c in range(pd.shape[0])]).reshape(-1)
            add, adds, rep = T.criterion_ADD_REP(pd, gt, nm, K)
            fs, cd = T.criterion_FSCORE(pd, gt, nm)
            a01, s01 = T.cal_ADD01d(add, adds, nm)
            rows.append(np.stack([mce, oce, mce2, smce, add, adds, rep, cd] +
                                 [fs[k] for k in ("FSCORE@2mm", "FSCORE@5mm", "FSCORE@10mm", "FSCORE@2cm", "FSCORE@5cm", "FSCORE@10cm")] +
                                 [a01, s01, T.cal_REP5(rep)], -1).astype(np.float64))
    finally:
        torch.Tensor.cuda, torch.cdist = _cuda, _cd
    np.savez_compressed(os.path.join(OUT, "object_metrics.npz"), metrics=np.stack(rows), fp=cases.fingerprint(inp["pd_rt"], inp["gt_rt"]))
    print("object_metrics", np.stack(rows).shape)


def main():
    torch.set_num_threads(8)
    mano, anch, objs = cases.assets()
    ref = load_reference(mano, anch, objs)
    os.makedirs(OUT, exist_ok=True)
    object_metrics_fixture(ref, objs)
    if "--only-object-metrics" in sys.argv:
        return
    _, marg, sde_fn, eps, T = ref.sde.init_sde("ve")
    agent = ref.sbm.ScoreBasedModelAgent()

    def ref_denoiser(head, st):
        d = ref.denoiser.BaseDenoiser(marg, head=head).eval()
        d.load_state_dict({k: torch.from_numpy(v) for k, v in st.items()})
        return d

    # ---- sampler fixtures: (head, bs, S, last_std, seed, steps)
    for name, (head, bs, S, last_std, seed, steps) in {
            "sampler_obj": ("obj", 2, 12, 0.05, 0, 11),
            "sampler_obj_stiff": ("obj", 2, 6, 3.0, 1, 7),
            "sampler_hand": ("mano_pose", 1, 8, 0.05, 2, 5)}.items():
        st, enc, init = cases.sampler_case(head, bs, S, last_std, seed)
        den = ref_denoiser(head, st)
        feat = enc[:, None].repeat(1, S, 1).reshape(-1, 1024)
        agent.prior_fn = lambda shape, T, _i=init: _i.clone()
        agent.cfg.sampling_steps = steps
        calls = [0]
        orig = den.forward

        def counted(data, _o=orig):
            calls[0] += 1
            return _o(data)
        den.forward = counted
        xs, x = agent.sample({"feat": feat}, den, 0.65)
        tt = torch.ones(feat.shape[0], 1) * 0.31
        with torch.no_grad():
            ev = orig({"feat": feat, "sampled_pose": init, "t": tt})
        np.savez_compressed(os.path.join(OUT, name + ".npz"), head=head, bs=bs, S=S, last_std=last_std, seed=seed, steps=steps,
                            xs=xs.numpy(), x=x.numpy(), net_calls=calls[0], eval_t031=ev.numpy(),
                            fp=cases.fingerprint(enc, init))
        print(name, "net_calls", calls[0], "x", tuple(x.shape))
    agent.cfg.sampling_steps = 50

    # ---- MANO fixture
    hm = ref.head_mano.HeadMano(in_dim=1024)
    g = torch.Generator().manual_seed(42)
    pose, shape = torch.randn(7, 48, generator=g) * 0.6, torch.randn(7, 10, generator=g)
    v, j = hm.get_hand_verts(pose=pose, shape=shape)
    np.savez_compressed(os.path.join(OUT, "mano.npz"), pose=pose.numpy(), shape=shape.numpy(), verts=v.numpy(), joints=j.numpy())

    # ---- aggregation fixtures
    ho, hp = ref.head_object.HeadObject(), ref.physics.HeadPhysics(hid_dim=512)
    agg = ref.aggregation.HOI_Aggregator(hm.get_hand_verts, ho, hp)
    _cd = torch.cdist
    _topk = torch.Tensor.topk

    def canonical_topk(self, k, dim=-1, largest=True, sorted=True):
        order = torch.sort(-self if largest else self, dim=dim, stable=True)[1].narrow(dim, 0, k)
        return torch.return_types.topk((torch.gather(self, dim, order), order))
    for name, (bs, S, Kh, Ko, seed) in {"aggregate_small": (3, 16, 6, 4, 1), "aggregate_readme": (2, 100, 30, 10, 2)}.items():
        kw, batch, _ = cases.aggregate_case(bs, S, seed)
        kw.update(hand_topk=Kh, obj_topk=Ko)
        torch.cdist = lambda a, b, p=2.0, compute_mode=None: _cd(a, b, p=p, compute_mode="donot_use_mm_for_euclid_dist")
        try:
            with torch.no_grad():
                r_plain = agg(**cases.clone_kw(kw))
                torch.Tensor.topk = canonical_topk
                r = agg(**cases.clone_kw(kw))
        finally:
            torch.cdist = _cd
            torch.Tensor.topk = _topk
        changed = [k for k in r if isinstance(r[k], torch.Tensor) and not torch.equal(r[k], r_plain[k])]
        print(name, "outputs that depend on torch.topk's tie order:", changed)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), bs=bs, S=S, Kh=Kh, Ko=Ko, seed=seed,
                            fp=cases.fingerprint(kw["hand_pose_diff"], kw["obj_pose6d"], kw["hand_heatmap"]),
                            **{k: r[k].numpy() for k in ("obj_agg_6d", "pose6d_candidate", "agg_obj_vert", "hand_agg_mano",
                                                         "hand_agg_vert", "hand_agg_joint")})
        print(name, "done")
    # informational: how often does the reference's default (mm-based) cdist change the selections?
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
