"""`vpho_postprocess_hand` (one fused pass: float64 6D rotations -> float32 axis-angle + shape, [rows][steps][58]) against
the oracle's restatement of `vpho_net.postprocess_diffusion_hand` (lib/model/VPHO.py:306-331).  Bar: 1e-4 rad worst case / 5e-6 at the 99.9th
percentile on the axis-angle parameters away from theta = pi, 2e-5 on the rotation matrices everywhere, exact on the shape copy and on the output layout."""
import pytest
import torch

from oracle import vpho_oracle as O
from oracle.shims.pytorch3d.transforms.rotation_conversions import axis_angle_to_matrix
from vpho_b200 import capi


def _run(lib, dev, bs, S, n_steps, seed):
    g = torch.Generator().manual_seed(seed)
    n = bs * S
    xs = torch.randn(n_steps, n, 96, generator=g, dtype=torch.float64)          # sampler storage order
    shape = torch.randn(bs, 10, generator=g)
    out = torch.empty((n, n_steps, 58), dtype=torch.float32, device=dev)
    xd, sd = xs.to(dev), shape.to(dev)
    lib.check(lib.c.vpho_postprocess_hand(capi.ptr(xd), n_steps, n, S, capi.ptr(sd), capi.ptr(out), capi.stream_of(xd)),
              "vpho_postprocess_hand")
    if dev != "cpu":
        torch.cuda.synchronize()
    ref_in, ref_fin = O.postprocess_diffusion_hand(xs[-1].float(), shape, S, xs.permute(1, 0, 2).float())
    out = out.cpu()
    assert torch.equal(out[..., 48:], ref_in[..., 48:])
    # the rotation-vector map is ill-conditioned towards theta = pi (and its sign is free there): compare the parameters
    # directly away from pi, and everything as rotation matrices
    a, r = out[..., :48].reshape(-1, 3), ref_in[..., :48].reshape(-1, 3)
    away = r.norm(dim=1) < 2.5
    d = (a - r).abs()[away].flatten()
    # random 6D inputs include nearly collinear column pairs (ill-conditioned Gram-Schmidt in float32): the bulk is
    # at rounding level, the worst of ~10^6 rotations within 1e-4
    assert d.max().item() < 1e-4 and torch.quantile(d[:: max(1, d.numel() // 100000)], 0.999).item() < 5e-6
    assert (axis_angle_to_matrix(a) - axis_angle_to_matrix(r)).abs().max().item() < 2e-5
    return out, ref_fin


def test_postprocess_emulated(emu_lib):
    out, ref_fin = _run(emu_lib, "cpu", 2, 3, 4, 0)
    assert (axis_angle_to_matrix(out[:, -1, :48].reshape(-1, 3)) - axis_angle_to_matrix(ref_fin[:, :48].reshape(-1, 3))).abs().max().item() < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("bs,S,n_steps", [(1, 1, 1), (3, 5, 7), (8, 100, 50)])
def test_postprocess_cuda(cuda_lib, bs, S, n_steps):
    _run(cuda_lib, "cuda", bs, S, n_steps, bs)


@pytest.mark.gpu
def test_postprocess_matches_composed_path(cuda_lib):
    """bit-equal to `.float()` -> vpho_rot6d_to_axis_angle -> cat, the three-kernel path it replaces"""
    lib = cuda_lib
    g = torch.Generator().manual_seed(5)
    n, T = 12, 6
    xs = torch.randn(T, n, 96, generator=g, dtype=torch.float64).cuda()
    shape = torch.randn(3, 10, generator=g).cuda()
    out = torch.empty((n, T, 58), dtype=torch.float32, device="cuda")
    lib.check(lib.c.vpho_postprocess_hand(capi.ptr(xs), T, n, 4, capi.ptr(shape), capi.ptr(out), capi.stream_of(xs)), "pp")
    x32 = xs.float().contiguous()
    aa = torch.empty((T * n * 16, 3), dtype=torch.float32, device="cuda")
    lib.check(lib.c.vpho_rot6d_to_axis_angle(capi.ptr(x32), T * n * 16, capi.ptr(aa), capi.stream_of(x32)), "aa")
    comp = torch.cat((aa.reshape(T, n, 48).permute(1, 0, 2), shape.repeat_interleave(4, 0)[:, None].expand(n, T, 10)), -1)
    assert torch.equal(out, comp)
    # float32 trajectory in: same result as the float64 one after `.float()`
    out32 = torch.empty_like(out)
    lib.check(lib.c.vpho_postprocess_hand_f32(capi.ptr(x32), T, n, 4, capi.ptr(shape), capi.ptr(out32), capi.stream_of(x32)), "pp32")
    assert torch.equal(out32, out)
    # empty inputs are accepted
    lib.check(lib.c.vpho_postprocess_hand(None, 0, 0, 4, None, None, None), "empty")
