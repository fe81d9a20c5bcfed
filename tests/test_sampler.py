"""Sampler parity: the device-resident RK45 probability-flow sampler + score network (csrc/sampler.cu, C ABI
`vpho_score_eval` / `vpho_sample_*`) against the oracle restatement of `cond_ode_sampler`
(lib/model/score_based_model.py:45-105) which calls scipy's RK45 exactly as the reference does.

Tolerances (written here, float32 network / float64 state):
  * one network call: 5e-6 relative L2 (different FP32 summation order than MKL's sgemm);
  * sampler output: identical controller trajectory (nfev, accepted, rejected bit-equal) and
    |x - x_oracle| <= 2e-5 * max(1, |x|) element-wise (2e-4 for the deliberately stiff last_std >= 0.5 cases, whose
    100x larger scores amplify the float32 rounding of the network through the ODE).
"""
import numpy as np
import pytest
import torch

from oracle import vpho_oracle as O
from vpho_b200 import synthetic as syn
from vpho_b200.score_based_model import Denoiser, ScoreBasedModelAgent, ve_prior_std


def _setup(head, lib, bs, S, last_std, seed=0, dev="cpu"):
    st = syn.make_denoiser_state(head, seed, last_std=last_std)
    den, od = Denoiser(st, lib=lib), O.OracleDenoiser(st)
    g = torch.Generator().manual_seed(100 + seed)
    enc = torch.relu(torch.randn(bs, 1024, generator=g))
    feat = enc[:, None].repeat(1, S, 1).reshape(-1, 1024)
    return den, od, enc, feat, g


def _check_eval(den, od, feat, g, S, dev):
    D = den.out_dim
    x = torch.randn(feat.shape[0], D, generator=g) * 2.5
    for t in (0.65, 0.31, 1e-5):
        tt = torch.ones(feat.shape[0], 1) * t
        o1 = den({"feat": feat.to(dev), "sampled_pose": x.to(dev), "t": tt.to(dev), "sample_num": S}).cpu()
        o2 = od({"feat": feat, "sampled_pose": x, "t": tt})
        assert ((o1 - o2).norm() / o2.norm()).item() < 5e-6


def _check_sample(den, od, feat, S, dev, seed, steps=50, T0=0.65, expect_reject=None, tol=2e-5):
    agent = ScoreBasedModelAgent(sampling_steps=steps, sample_num=S)
    torch.manual_seed(seed)
    xs, x = agent.sample({"feat": feat.to(dev)}, den, T0)
    torch.manual_seed(seed)
    init = torch.randn(feat.shape[0], den.out_dim) * ve_prior_std(T0)     # the same global-RNG draw (sde.py:26-28)
    xs2, x2, info = O.oracle_sample(od, feat, T0, init, steps)
    li = agent.last_info
    assert li["status"] == 1 and info["status"] == 0
    assert li["nfev"] == info["nfev"] and li["net_calls"] == info["net_calls"]
    assert xs.dtype == torch.float64 and x.dtype == torch.float64
    assert tuple(xs.shape) == tuple(xs2.shape) and tuple(x.shape) == tuple(x2.shape)
    assert ((x.cpu() - x2).abs() <= tol * x2.abs().clamp(min=1)).all()
    assert ((xs.cpu() - xs2).abs() <= tol * xs2.abs().clamp(min=1)).all()
    assert torch.equal(xs[:, 0].cpu().float(), init)                      # first t_eval point is the prior draw
    if expect_reject is not None:
        assert (li["rejected"] > 0) == expect_reject
    return li


def test_sampler_emulated_object_head(emu_lib):
    den, od, enc, feat, g = _setup("obj", emu_lib, bs=2, S=12, last_std=0.05)
    _check_eval(den, od, feat, g, 12, "cpu")
    _check_sample(den, od, feat, 12, "cpu", seed=3)


def test_sampler_emulated_stiff_scores_need_more_attempts(emu_lib):
    # a much larger last layer makes the ODE move: more accepted steps than the first enqueue (continue path)
    den, od, enc, feat, g = _setup("obj", emu_lib, bs=2, S=6, last_std=3.0)
    li = _check_sample(den, od, feat, 6, "cpu", seed=4, steps=11, tol=2e-4)
    assert li["attempts"] >= 6


def test_sampler_emulated_ragged_rows(emu_lib):
    # N = 7 rows is not a multiple of any tile; rows_per_feat = 1 (no repeat structure)
    st = syn.make_denoiser_state("obj", 1, last_std=0.05)
    den, od = Denoiser(st, lib=emu_lib), O.OracleDenoiser(st)
    g = torch.Generator().manual_seed(9)
    feat = torch.relu(torch.randn(7, 1024, generator=g))
    _check_eval(den, od, feat, g, 0, "cpu")
    _check_sample(den, od, feat, 0, "cpu", seed=5, steps=5)


@pytest.mark.gpu
@pytest.mark.parametrize("head,bs,S,last_std", [("obj", 4, 100, 0.05), ("mano_pose", 4, 100, 0.05),
                                                ("mano_pose", 3, 37, 0.5), ("obj", 5, 100, 3.0)])
def test_sampler_cuda(cuda_lib, head, bs, S, last_std):
    den, od, enc, feat, g = _setup(head, None, bs, S, last_std)
    _check_eval(den, od, feat, g, S, "cuda")
    _check_sample(den, od, feat, S, "cuda", seed=bs, tol=2e-4 if last_std >= 0.5 else 2e-5)


@pytest.mark.gpu
def test_sampler_cuda_feat_unique_equals_repeated(cuda_lib):
    den, od, enc, feat, g = _setup("mano_pose", None, 3, 100, 0.05)
    agent = ScoreBasedModelAgent(50, 100)
    torch.manual_seed(1)
    xs1, x1 = agent.sample({"feat": feat.cuda()}, den, 0.65)
    torch.manual_seed(1)
    xs2, x2 = agent.sample({"feat_unique": enc.cuda(), "n_rows": 300}, den, 0.65)
    assert torch.equal(x1, x2) and torch.equal(xs1, xs2)


@pytest.mark.gpu
def test_sampler_cuda_readme_batch_runs(cuda_lib):
    # README shape (bs 64 x 100): finishes, finite, deterministic across two runs
    den, od, enc, feat, g = _setup("mano_pose", None, 64, 100, 0.05)
    agent = ScoreBasedModelAgent(50, 100)
    outs = []
    for _ in range(2):
        torch.manual_seed(2)
        _, x = agent.sample({"feat_unique": enc.cuda(), "n_rows": 6400}, den, 0.65, return_inprocess=False)
        outs.append(x)
    assert torch.isfinite(outs[0]).all() and torch.equal(outs[0], outs[1])


def _nan_state(head="obj"):
    st = syn.make_denoiser_state(head, 3, last_std=0.05)
    st["head.head.2.bias"] = st["head.head.2.bias"].copy()
    st["head.head.2.bias"][1, 2] = np.nan        # one output dimension of every row becomes NaN
    return st


def _check_nan_semantics(lib, dev):
    """NaN scores are replaced by 0 inside the ODE right-hand side (score_based_model.py:69-71) but NOT in the final
    predictor step (:95-104), so the affected dimension keeps its prior value through the integration and ends NaN."""
    st = _nan_state()
    den, od = Denoiser(st, lib=lib), O.OracleDenoiser(st)
    g = torch.Generator().manual_seed(1)
    feat = torch.relu(torch.randn(6, 1024, generator=g))
    agent = ScoreBasedModelAgent(sampling_steps=7, sample_num=0)
    torch.manual_seed(3)
    xs, x = agent.sample({"feat": feat.to(dev)}, den, 0.65)
    torch.manual_seed(3)
    init = torch.randn(6, 9) * ve_prior_std(0.65)
    xs2, x2, info = O.oracle_sample(od, feat, 0.65, init, 7)
    assert agent.last_info["nan"] and info["nan"]
    assert agent.last_info["nfev"] == info["nfev"]
    xs, x = xs.cpu(), x.cpu()
    assert torch.equal(torch.isnan(x), torch.isnan(x2)) and torch.isnan(x[:, 5]).all()
    assert not torch.isnan(xs).any() and not torch.isnan(xs2).any()
    ok = ~torch.isnan(x2)
    assert ((x[ok] - x2[ok]).abs() <= 2e-5 * x2[ok].abs().clamp(min=1)).all()
    assert ((xs - xs2).abs() <= 2e-5 * xs2.abs().clamp(min=1)).all()
    assert torch.equal(xs[:, -1, 5].float(), init[:, 5])          # the NaN dimension never moved during the ODE


def test_sampler_emulated_nan_scores_follow_reference_semantics(emu_lib):
    _check_nan_semantics(emu_lib, "cpu")


@pytest.mark.gpu
def test_sampler_cuda_nan_scores_follow_reference_semantics(cuda_lib):
    _check_nan_semantics(None, "cuda")


@pytest.mark.gpu
def test_sampler_cuda_empty_batch(cuda_lib):
    den = Denoiser(syn.make_denoiser_state("obj", 0))
    agent = ScoreBasedModelAgent(5, 0)
    xs, x = agent.sample({"feat": torch.zeros(0, 1024, device="cuda")}, den, 0.65)
    assert tuple(x.shape) == (0, 9) and tuple(xs.shape) == (0, 5, 9)


@pytest.mark.gpu
def test_sampler_cuda_simt_and_tensor_core_paths_agree(cuda_lib):
    st = syn.make_denoiser_state("mano_pose", 0)
    d_simt = Denoiser(st, strict_fp32=True)          # VPHO_DENOISER_STRICT_FP32: the FP32 SIMT kernels
    d_tc = Denoiser(st)
    g = torch.Generator().manual_seed(0)
    enc = torch.relu(torch.randn(5, 1024, generator=g)).cuda()
    x = (torch.randn(500, 96, generator=g) * 2.5).cuda()
    for t in (0.65, 0.1):
        data = {"feat_unique": enc, "sampled_pose": x, "t": torch.full((500, 1), t, device="cuda")}
        a, b = d_tc(data), d_simt(data)
        assert ((a - b).norm() / b.norm()).item() < 5e-7        # 3xTF32 keeps FP32-class accuracy


@pytest.mark.gpu
@pytest.mark.parametrize("name,tol", [("sampler_obj", 2e-5), ("sampler_obj_stiff", 2e-4), ("sampler_hand", 2e-5)])
def test_sampler_cuda_matches_reference_golden(cuda_lib, name, tol):
    """CUDA sampler against the fixtures minted from the reference's own files (tests/golden, oracle/make_golden.py)."""
    import os
    from oracle import cases
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    head, bs, S = str(g["head"]), int(g["bs"]), int(g["S"])
    st, enc, init = cases.sampler_case(head, bs, S, float(g["last_std"]), int(g["seed"]))
    assert abs(cases.fingerprint(enc, init) - float(g["fp"])) < 1e-6
    den = Denoiser(st)
    agent = ScoreBasedModelAgent(sampling_steps=int(g["steps"]), sample_num=S)
    xs, x = agent.sample({"feat_unique": enc.cuda(), "n_rows": bs * S}, den, 0.65, prior=init)
    assert agent.last_info["net_calls"] == int(g["net_calls"])
    assert (np.abs(x.cpu().numpy() - g["x"]) <= tol * np.maximum(1, np.abs(g["x"]))).all()
    assert (np.abs(xs.cpu().numpy() - g["xs"]) <= tol * np.maximum(1, np.abs(g["xs"]))).all()


@pytest.mark.gpu
@pytest.mark.parametrize("bs_a,bs_b,S,std_a,std_b", [(4, 4, 100, 0.05, 0.05), (3, 5, 30, 0.05, 0.5), (2, 1, 7, 0.5, 0.05)])
def test_sampler_cuda_pair_is_bit_identical_to_separate_samplers(cuda_lib, bs_a, bs_b, S, std_a, std_b):
    """`sample_pair` (hand and object integrations advanced in lock-step through shared launches, `vpho_sample_pair_*`)
    must reproduce two separate `sample()` calls bit for bit -- including when one integration needs more RK attempts
    than the other (stiff scores on one side), odd tile counts and different batch sizes."""
    den_a, _, enc_a, _, g = _setup("mano_pose", cuda_lib, bs_a, S, std_a, seed=1)
    den_b, _, enc_b, _, _ = _setup("obj", cuda_lib, bs_b, S, std_b, seed=2)
    agent = ScoreBasedModelAgent(sampling_steps=20, sample_num=S)
    pa = torch.randn(bs_a * S, den_a.out_dim, generator=g) * ve_prior_std(0.65)
    pb = torch.randn(bs_b * S, den_b.out_dim, generator=g) * ve_prior_std(0.65)
    da = {"feat_unique": enc_a.cuda(), "n_rows": bs_a * S}
    db = {"feat_unique": enc_b.cuda(), "n_rows": bs_b * S}
    xs_a, x_a = agent.sample(da, den_a, 0.65, prior=pa)
    info_a = dict(agent.last_info)
    xs_b, x_b = agent.sample(db, den_b, 0.65, prior=pb)
    info_b = dict(agent.last_info)
    for _ in range(8):
        (ys_a, y_a, pend_a), (ys_b, y_b, pend_b) = agent.sample_pair(da, den_a, db, den_b, 0.65, prior_a=pa, prior_b=pb)
        torch.cuda.synchronize()
        if all([pend_a.resolve(), pend_b.resolve()]):
            break
    else:
        raise AssertionError("pair sampler did not converge")
    for k in ("nfev", "accepted", "rejected"):
        assert pend_a.info[k] == info_a[k] and pend_b.info[k] == info_b[k]
    assert torch.equal(y_a, x_a) and torch.equal(y_b, x_b)
    assert torch.equal(ys_a, xs_a) and torch.equal(ys_b, xs_b)
    for _ in range(8):
        (fs_a, f_a, pend_a), (fs_b, f_b, pend_b) = agent.sample_pair(da, den_a, db, den_b, 0.65, prior_a=pa, prior_b=pb,
                                                                     inprocess_float32=(True, False))
        torch.cuda.synchronize()
        if all([pend_a.resolve(), pend_b.resolve()]):
            break
    assert fs_a.dtype == torch.float32 and torch.equal(fs_a, xs_a.float()) and torch.equal(fs_b, xs_b) and torch.equal(f_a, x_a)


def test_sampler_emulated_pair_matches_separate_and_tolerates_an_empty_job(emu_lib):
    """host logic of `vpho_sample_pair_*` on the CPU emulator (the error-norm / post-step kernels serve both jobs there
    too): bit-identical to separate samplers; a job without rows drops out"""
    S = 5
    den_a, _, enc_a, _, g = _setup("obj", emu_lib, 2, S, 0.05, seed=3)
    den_b, _, enc_b, _, _ = _setup("obj", emu_lib, 1, S, 0.5, seed=4)
    agent = ScoreBasedModelAgent(sampling_steps=6, sample_num=S)
    pa = torch.randn(2 * S, den_a.out_dim, generator=g) * ve_prior_std(0.65)
    pb = torch.randn(1 * S, den_b.out_dim, generator=g) * ve_prior_std(0.65)
    da, db = {"feat_unique": enc_a, "n_rows": 2 * S}, {"feat_unique": enc_b, "n_rows": S}
    xs_a, x_a = agent.sample(da, den_a, 0.65, prior=pa)
    xs_b, x_b = agent.sample(db, den_b, 0.65, prior=pb)
    for _ in range(8):
        (ys_a, y_a, pend_a), (ys_b, y_b, pend_b) = agent.sample_pair(da, den_a, db, den_b, 0.65, prior_a=pa, prior_b=pb)
        if all([pend_a.resolve(), pend_b.resolve()]):
            break
    else:
        raise AssertionError("pair sampler did not converge")
    assert torch.equal(y_a, x_a) and torch.equal(y_b, x_b) and torch.equal(ys_a, xs_a) and torch.equal(ys_b, xs_b)
    # float32 trajectory of one job (`vpho_sample_args.xs_f32`): exactly `.float()` of the float64 one
    for _ in range(8):
        (fs_a, f_a, pend_a), (fs_b, f_b, pend_b) = agent.sample_pair(da, den_a, db, den_b, 0.65, prior_a=pa, prior_b=pb,
                                                                     inprocess_float32=(True, False))
        if all([pend_a.resolve(), pend_b.resolve()]):
            break
    assert fs_a.dtype == torch.float32 and torch.equal(fs_a, xs_a.float()) and torch.equal(fs_b, xs_b) and torch.equal(f_a, x_a)
    # second job empty
    empty = {"feat_unique": enc_b[:0], "n_rows": 0}
    for _ in range(8):
        (zs_a, z_a, pend_a), (zs_e, z_e, pend_e) = agent.sample_pair(da, den_a, empty, den_b, 0.65, prior_a=pa,
                                                                     prior_b=pb[:0])
        if all([pend_a.resolve(), pend_e.resolve()]):
            break
    assert torch.equal(z_a, x_a) and z_e.shape == (0, den_b.out_dim)


def test_spare_attempt_policy_host_logic():
    """The number of RK attempts enqueued up front: last batch's count plus one spare, the spare dropped once SPARE_WINDOW
    batches in a row needed the same count and restored by the first batch that breaks the pattern (host logic only)."""
    from types import SimpleNamespace
    from vpho_b200.score_based_model import SPARE_WINDOW, PendingSample
    den, agent = SimpleNamespace(), SimpleNamespace(spare_attempt=True, last_info=None)

    def feed(attempts, status=1):
        p = PendingSample(agent, den, None, 1, attempts, None)
        p.pair = object()
        return p.resolve([status, 6 * attempts + 2, attempts, 0, 0, attempts, 0, 0])

    for i in range(SPARE_WINDOW - 1):
        assert feed(3) and den.attempts_hint == 4
    assert feed(3) and den.attempts_hint == 3               # stable: no spare
    assert feed(3) and den.attempts_hint == 3
    assert not feed(3, status=0) and den.attempts_hint == 3  # unfinished with the enqueued budget: the caller continues it
    assert feed(4) and den.attempts_hint == 5               # pattern broken: spare is back
    for i in range(SPARE_WINDOW - 2):
        assert feed(4) and den.attempts_hint == 5
    assert feed(4) and den.attempts_hint == 4
    agent.spare_attempt = False
    assert feed(5) and den.attempts_hint == 5
