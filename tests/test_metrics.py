"""Final pose-error metrics on the device (C ABI `vpho_pose_metrics`) against the oracle's restatement of TesterHand
MJE / MVE (lib/engine/test.py:657-679) and TesterObject ADD / ADD-S (lib/engine/test.py:413-442).  Bar: 1e-3 mm."""
import pytest
import torch

from oracle import cases
from oracle import vpho_oracle as O
from vpho_b200.aggregation import Assets, hand_pa_metrics, pose_metrics


def _case(lib, dev, n, seed):
    mano, anch, objs = cases.assets()
    g = torch.Generator().manual_seed(seed)
    gj, gv = torch.randn(n, 21, 3, generator=g) * 0.05, torch.randn(n, 778, 3, generator=g) * 0.05
    pj, pv = gj + 0.01 * torch.randn(n, 21, 3, generator=g), gv + 0.01 * torch.randn(n, 778, 3, generator=g)
    go = torch.cat([torch.randn(n, 6, generator=g, dtype=torch.float64), 0.1 * torch.randn(n, 3, generator=g, dtype=torch.float64)], 1)
    po = go + 0.02 * torch.randn(n, 9, generator=g, dtype=torch.float64)
    names = [objs["names"][int(i)] for i in torch.randint(0, 21, (n,), generator=g)]
    mje, mve = O.hand_pose_error_mm(pj, gj, pv, gv)
    add, adds = O.object_add_mm(O.OracleObject(objs), po, go, names)
    ref = torch.stack([mje, mve, add, adds], 1)
    out = pose_metrics(Assets(anch, objs, lib=lib), pj.to(dev), gj.to(dev), pv.to(dev), gv.to(dev), po.to(dev), go.to(dev), names)
    assert (out.cpu() - ref).abs().max().item() < 1e-3


def test_metrics_emulated(emu_lib):
    _case(emu_lib, "cpu", 2, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 64])
def test_metrics_cuda(cuda_lib, n):
    _case(None, "cuda", n, n)


def _pa_case(lib, dev, n, seed, degenerate=False):
    """PA-MJE / PA-MVE / JE against the oracle restatement of criterion_MJE_PAMJE + rigid_align_AtoB.  Bar: 1e-3 mm
    (the reference runs numpy float32; the device accumulates the moments in float64)."""
    import numpy as np
    g = torch.Generator().manual_seed(seed)
    gj, gv = torch.randn(n, 21, 3, generator=g) * 0.05, torch.randn(n, 778, 3, generator=g) * 0.05
    # predictions: rotated / scaled / shifted ground truth plus noise, so that the alignment has real work to do
    q = torch.linalg.qr(torch.randn(n, 3, 3, generator=g))[0]
    if degenerate:
        q[:, :, 2] = -q[:, :, 2] * torch.sign(torch.linalg.det(q))[:, None]       # improper: exercises the reflection fix
    pj = 1.2 * gj @ q + 0.03 + 0.004 * torch.randn(n, 21, 3, generator=g)
    pv = 0.9 * gv @ q - 0.02 + 0.004 * torch.randn(n, 778, 3, generator=g)
    ref = np.concatenate([np.stack(O.hand_pa_error_mm(pj, gj, pv, gv)[:2], 1), O.hand_pa_error_mm(pj, gj, pv, gv)[2]], 1)
    out = hand_pa_metrics(pj.to(dev), gj.to(dev), pv.to(dev), gv.to(dev), lib=lib).cpu().numpy()
    assert np.abs(out - ref).max() < 1e-3, np.abs(out - ref).max()


def test_pa_metrics_emulated(emu_lib):
    _pa_case(emu_lib, "cpu", 2, 0)
    _pa_case(emu_lib, "cpu", 2, 1, degenerate=True)


@pytest.mark.gpu
@pytest.mark.parametrize("n,deg", [(1, False), (64, False), (16, True)])
def test_pa_metrics_cuda(cuda_lib, n, deg):
    _pa_case(cuda_lib, "cuda", n, n, degenerate=deg)
