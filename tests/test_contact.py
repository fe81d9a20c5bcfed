"""BASELINE config 3 parity: MANO LBS -> force anchors -> anchor/vertex contact scoring (C ABI `vpho_anchor_contact`,
`vpho_vertex_contact`) against the oracle's restatement of `select_by_physics` (lib/model/aggregation.py:553-590) and
exact-mode `torch.cdist`.  Distances: 2e-7 m absolute (FP32 at 0.6 m depth); finger scores: 1e-4 relative to the scale."""
import pytest
import torch

from oracle import cases
from oracle import vpho_oracle as O
from vpho_b200.aggregation import Assets, HeadPhysics, anchor_contact, vertex_contact
from vpho_b200.head_mano import HeadMano


def _case(lib, dev, G, Cn, P, seed):
    mano, anch, objs = cases.assets()
    g = torch.Generator().manual_seed(seed)
    pose = torch.randn(G * Cn, 48, generator=g) * 0.4
    shape = torch.randn(G * Cn, 10, generator=g)
    root = torch.tensor([0.03, -0.02, 0.6]) + 0.02 * torch.randn(G, 1, 3, generator=g)
    obj = root + torch.tensor([0.07, 0.0, 0.02]) + 0.05 * torch.randn(G, P, 3, generator=g)
    fl = torch.randn(G, 32, 3, generator=g).abs() * 0.3
    hm = HeadMano(mano, lib=lib)
    verts, _ = hm.get_hand_verts(pose=pose.to(dev), shape=shape.to(dev))
    verts_cam = verts.reshape(G, Cn, 778, 3) + root.to(dev)[:, None]
    fl_rep = fl[:, None].repeat(1, Cn, 1, 1)
    fp, fg = HeadPhysics(Assets(anch, objs, lib=lib)).from_local_to_global(fl_rep.to(dev), verts_cam)
    dist, score = anchor_contact(fp, fg, obj.to(dev), lib=lib)
    vd = vertex_contact(verts_cam, obj.to(dev), lib=lib)
    # oracle on the same posed vertices
    oa = O.OracleAnchors(anch)
    vc = verts_cam.cpu()
    fp2, fg2 = oa.from_local_to_global(fl_rep.reshape(-1, 32, 3), vc.reshape(-1, 778, 3))
    fp2, fg2 = fp2.reshape(G, Cn, 32, 3), fg2.reshape(G, Cn, 32, 3)
    cd = torch.stack([O.exact_cdist(fp2[i], obj[i]).min(dim=-1)[0] for i in range(G)], 0)
    fn = fg2.norm(dim=-1)
    sc = -((fn / fn.sum(-1, keepdim=True)) * cd * (fg2 / fn[..., None]).sum(-2).norm(dim=-1)[..., None])
    fs = torch.stack([sc[..., f].sum(-1) for f in O.FINGER_FORCE_LEVEL], -1)
    vd2 = torch.stack([O.exact_cdist(vc[i], obj[i]).min(dim=-1)[0] for i in range(G)], 0)
    assert (dist.cpu() - cd).abs().max().item() < 2e-7 + 2e-5 * cd.max().item()
    assert (score.cpu() - fs).abs().max().item() < 1e-4 * fs.abs().max().item()
    assert (vd.cpu() - vd2).abs().max().item() < 2e-7


def test_contact_emulated(emu_lib):
    _case(emu_lib, "cpu", G=2, Cn=3, P=300, seed=0)


@pytest.mark.gpu
@pytest.mark.parametrize("G,Cn,P", [(2, 31, 2048), (3, 7, 4096), (1, 5, 8192), (2, 3, 1000)])
def test_contact_cuda(cuda_lib, G, Cn, P):
    _case(None, "cuda", G, Cn, P, seed=P)
