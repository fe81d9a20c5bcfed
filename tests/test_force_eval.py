"""BASELINE config 5 parity: pseudo-force evaluation (C ABI `vpho_force_eval`) against the oracle's restatement of one
ForceOptimizer iteration's forward math (lib/engine/force_optimization.py:141-171; `get_local_force`
lib/model/physics.py:546-557 is additionally checked bit-identical against the reference's own method when
/root/reference is present).  Tolerances: forces / points 2e-5 relative to the force scale (bone directions come from a
778-term FP32 regression in camera coordinates), the four terms 1e-4 relative."""
import pytest
import torch

from oracle import cases
from oracle import vpho_oracle as O
from vpho_b200.aggregation import Assets, force_eval
from vpho_b200.head_mano import HeadMano


def _case(lib, dev, n, seed):
    mano, anch, objs = cases.assets()
    g = torch.Generator().manual_seed(seed)
    pose, shape = torch.randn(n, 48, generator=g) * 0.4, torch.randn(n, 10, generator=g)
    verts, _ = HeadMano(mano, lib=lib).get_hand_verts(pose=pose.to(dev), shape=shape.to(dev))
    vert3d = verts.cpu() + torch.tensor([0.02, -0.03, 0.6])
    scale = torch.randn(n, 32, generator=g) * 0.5
    weight = torch.randn(n, 32, 8, generator=g)
    fc = torch.rand(n, 32, generator=g)
    mask = fc > 0.1
    gravity = torch.nn.functional.normalize(torch.randn(n, 1, 3, generator=g), dim=-1)
    com = torch.tensor([0.08, -0.02, 0.62]) + 0.02 * torch.randn(n, 1, 3, generator=g)
    t_ref, fl_ref, fp_ref, fg_ref = O.force_eval_terms(O.OracleAnchors(anch), vert3d, scale, weight, mask, fc, gravity, com)
    t, fl, fp, fg = force_eval(Assets(anch, objs, lib=lib), vert3d.to(dev), scale.to(dev), weight.to(dev), mask.to(dev),
                               fc.to(dev), gravity.to(dev), com.to(dev), return_forces=True)
    assert (fl.cpu() - fl_ref).abs().max().item() < 1e-6
    assert (fp.cpu() - fp_ref).abs().max().item() < 2e-7
    assert (fg.cpu() - fg_ref).abs().max().item() < 1e-4 * fg_ref.abs().max().item()
    for k in range(4):
        assert (t.cpu()[:, k] - t_ref[:, k]).abs().max().item() < 1e-4 * t_ref[:, k].abs().max().item() + 1e-6, k


def test_force_eval_emulated(emu_lib):
    _case(emu_lib, "cpu", 5, 0)


def test_oracle_get_local_force_matches_reference():
    from oracle.reference_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("/root/reference not present")
    mano, anch, objs = cases.assets()
    ref = load_reference(mano, anch, objs)
    hp = ref.physics.HeadPhysics(hid_dim=512)
    g = torch.Generator().manual_seed(1)
    scale, weight = torch.randn(6, 32, generator=g), torch.randn(6, 32, 8, generator=g)
    assert torch.equal(hp.get_local_force(scale, weight), O.get_local_force(scale, weight))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 64, 6400])
def test_force_eval_cuda(cuda_lib, n):
    _case(None, "cuda", n, n)
