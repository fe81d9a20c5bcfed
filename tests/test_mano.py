"""MANO layer parity: CUDA kernel (C ABI `vpho_mano_forward`) vs the oracle restatement of manopth
(`HeadMano.get_hand_verts`, lib/model/head_mano.py:78-87).  Tolerance: 1e-5 relative, FP32 (north_star)."""
import numpy as np
import pytest
import torch

from oracle import vpho_oracle as O
from vpho_b200.head_mano import HeadMano

REL_TOL = 1e-5


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def _check(hm, om, n, dev, seed, scale=0.5):
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(n, 48, generator=g) * scale
    s = torch.randn(n, 10, generator=g)
    v, j = hm.get_hand_verts(pose=p.to(dev), shape=s.to(dev))
    v2, j2 = om(p, s)
    assert v.shape == (n, 778, 3) and j.shape == (n, 21, 3)
    assert _rel(v.cpu(), v2) < REL_TOL and _rel(j.cpu(), j2) < REL_TOL
    # per-element: relative to the hand's size (~0.1 m) so wrist-centred zeros do not blow up the ratio
    assert (v.cpu() - v2).abs().max().item() < 1e-5 * 0.2
    # joints only (no vertices): the FP32 SIMT path.  The 16 kinematic joints are the same arithmetic on both paths; the 5
    # fingertips are vertices, which the full path blends on tensor cores (3xFP16 split, FP32 accumulate)
    _, j3 = hm.get_hand_verts(pose=p.to(dev), shape=s.to(dev), need_verts=False)
    tips = [4, 8, 12, 16, 20]
    kin = [i for i in range(21) if i not in tips]
    assert torch.equal(j3[:, kin], j[:, kin])
    assert (j3[:, tips] - j[:, tips]).abs().max().item() < 1e-6 if n else True


def test_mano_emulated_kernel(assets, emu_lib):
    hm, om = HeadMano(assets["mano"], lib=emu_lib), O.OracleMano(assets["mano"])
    for n in (1, 5, 19):
        _check(hm, om, n, "cpu", seed=n)


def test_mano_zero_pose_is_template(assets, emu_lib):
    hm = HeadMano(assets["mano"], lib=emu_lib)
    v, j = hm.get_hand_verts(pose=torch.zeros(1, 48), shape=torch.zeros(1, 10))
    m = assets["mano"]
    wrist = (m["J_regressor"].astype(np.float64) @ m["v_template"].astype(np.float64))[0]
    ref = torch.from_numpy((m["v_template"] - wrist).astype(np.float32))
    assert (v[0] - ref).abs().max().item() < 2e-7


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 3, 64, 1984, 6400, 12800])
def test_mano_cuda(assets, cuda_lib, n):
    hm, om = HeadMano(assets["mano"]), O.OracleMano(assets["mano"])
    _check(hm, om, n, "cuda", seed=n)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 63, 64, 65, 200, 1984, 6400])
def test_mano_cuda_tensor_core_blend_matches_the_fp32_kernel(assets, cuda_lib, n):
    """tcgen05 blend (3xFP16 operand split) vs the strict FP32 SIMT kernel (VPHO_MANO_STRICT_FP32) on the same inputs:
    every candidate count exercises a different grid split (candidate tiles of 64 x vertex-tile groups)."""
    hm = HeadMano(assets["mano"])
    g = torch.Generator().manual_seed(n)
    p = (torch.randn(n, 48, generator=g) * 0.7).cuda()
    s = (torch.randn(n, 10, generator=g) * 1.5).cuda()
    v1, j1 = hm.get_hand_verts(pose=p, shape=s)
    v2, j2 = hm.get_hand_verts(pose=p, shape=s, strict_fp32=True)
    assert (v1 - v2).abs().max().item() < 5e-7 and (j1 - j2).abs().max().item() < 5e-7
    assert _rel(v1, v2) < 2e-6


@pytest.mark.gpu
def test_mano_cuda_large_angles_and_empty(assets, cuda_lib):
    hm, om = HeadMano(assets["mano"]), O.OracleMano(assets["mano"])
    _check(hm, om, 257, "cuda", seed=7, scale=2.5)
    v, j = hm.get_hand_verts(pose=torch.zeros(0, 48, device="cuda"), shape=torch.zeros(0, 10, device="cuda"))
    assert v.shape == (0, 778, 3) and j.shape == (0, 21, 3)


def _dense_and_mixed_models():
    """the default synthetic model has <= 4 influences per vertex (as SMPL-family models do); the kernels also carry a
    dense 16-joint path: a fully dense weight matrix, and one where only a few vertices are dense (a warp whose vertices are
    not all sparse takes the dense path, its neighbours the sparse one)"""
    from vpho_b200 import synthetic as syn
    dense = syn.make_mano_model(max_influences=None)
    mixed = syn.make_mano_model()
    mixed = dict(mixed, weights=mixed["weights"].copy())
    for vid in (5, 100, 101, 640, 777):
        mixed["weights"][vid] = dense["weights"][vid]
    return {"dense": dense, "mixed": mixed}


def test_mano_emulated_dense_weights(emu_lib):
    for name, m in _dense_and_mixed_models().items():
        assert (m["weights"] != 0).sum(1).max() == 16
        _check(HeadMano(m, lib=emu_lib), O.OracleMano(m), 7, "cpu", seed=3)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["dense", "mixed"])
@pytest.mark.parametrize("n", [3, 200, 6400])
def test_mano_cuda_dense_weights(cuda_lib, kind, n):
    m = _dense_and_mixed_models()[kind]
    hm = HeadMano(m)
    _check(hm, O.OracleMano(m), n, "cuda", seed=n)
    g = torch.Generator().manual_seed(n + 1)
    p, s = (torch.randn(n, 48, generator=g) * 0.7).cuda(), torch.randn(n, 10, generator=g).cuda()
    v1, j1 = hm.get_hand_verts(pose=p, shape=s)
    v2, j2 = hm.get_hand_verts(pose=p, shape=s, strict_fp32=True)
    assert (v1 - v2).abs().max().item() < 5e-7 and (j1 - j2).abs().max().item() < 5e-7
