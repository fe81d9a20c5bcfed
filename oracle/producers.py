"""TEST INFRASTRUCTURE -- CPU restatement (PyTorch functional ops, float32) of the modules that PRODUCE the hot path's
inputs, i.e. everything `vpho_net.forward` does between the four RoI-aligned feature maps and the
`score_agent.sample` calls (lib/model/VPHO.py:129-178):

    HeadHeatmap2.forward                lib/model/head_inplane.py:42-107
    vpho_net.align_hm_to_bbox_rectangle lib/model/VPHO.py:333-346
    flip_tensor_by_mask_index           lib/model/VPHO.py:349-357
    Encoder.forward / Residual.forward  lib/model/encoding.py:5-73
    HeadMano.forward                    lib/model/head_mano.py:61-76
    CrossModule.forward                 lib/model/cross_module.py:91-137 (PosEmbedder :9-46, PositionalEncoding :49-88)
    HeadPhysics.forward                 lib/model/physics.py:648-721, get_local_force :546-557

All modules are in eval mode (BatchNorm uses running statistics, dropout is the identity).  Weights come in as ONE dict
keyed exactly like `vpho_net.state_dict()` (`head_hm_hand.conv_layers.0.weight`, `encoder_obj.reg.3.bn1.running_var`,
`cross_hand.attn.layers.0.self_attn.in_proj_weight`, ...), which is what `vpho_b200.checkpoint` hands over as `rest`.
Pinned against the reference's own module classes by tests/test_oracle_vs_reference.py (live, build container only) and
through tests/golden/producers_*.npz (minted by oracle/make_golden.py from those classes).

Reference quirks restated on purpose:
  * `HeadHeatmap2(act=nn.LeakyReLU(True))`: the positional argument is `negative_slope`, so the activation after the second
    convolution's BatchNorm is LeakyReLU(slope=1.0) = identity (head_inplane.py:43).
  * `align_hm_to_bbox_rectangle` builds its grid with `torch.meshgrid(..., indexing='ij')` and stacks (xx, yy): the sampled map
    is the TRANSPOSE of a plain re-crop (VPHO.py:336-345).
  * `nn.TransformerEncoderLayer(d_model, nhead=2)` is used with its default `batch_first=False` on a (bs, 65, 512) tensor:
    attention runs ACROSS THE IMAGES OF THE BATCH for each of the 65 token positions, and `PositionalEncoding` adds
    `pe[image_index]` (cross_module.py:107-110,131-133).  Results therefore depend on batch composition, like the sampler's.
  * `HeadPhysics.fc_weight` ends in a Softmax and `get_local_force` applies softmax again (physics.py:659-664,548).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from .shims.pytorch3d.transforms.rotation_conversions import matrix_to_axis_angle, rotation_6d_to_matrix

BN_EPS = 1e-5
LN_EPS = 1e-5


def _bn(x, st, p):
    return F.batch_norm(x, st[p + ".running_mean"], st[p + ".running_var"], st[p + ".weight"], st[p + ".bias"], False, 0.1, BN_EPS)


def _conv(x, st, p, pad):
    return F.conv2d(x, st[p + ".weight"], st.get(p + ".bias"), padding=pad)


def _lin(x, st, p):
    return F.linear(x, st[p + ".weight"], st[p + ".bias"])


def head_heatmap2(x, st, p):
    """HeadHeatmap2(256, out, 128): conv3x3, conv3x3+BN+LeakyReLU(1.0), deconv4x4/2+BN+ReLU, conv1x1 (head_inplane.py:99-104)."""
    x = _conv(x, st, p + ".conv_layers.0", 1)
    x = _conv(x, st, p + ".conv_layers.1", 1)
    x = F.leaky_relu(_bn(x, st, p + ".conv_layers.2"), 1.0)
    x = F.conv_transpose2d(x, st[p + ".deconv_layers.0.weight"], None, stride=2, padding=1, output_padding=0)
    x = F.relu(_bn(x, st, p + ".deconv_layers.1"))
    return _conv(x, st, p + ".final_layer", 0)


def align_hm_to_bbox_rectangle(hm, bbox, bbox_rect):
    """VPHO.py:333-346."""
    n = hm.shape[-1]
    xx, yy = torch.meshgrid(torch.arange(n), torch.arange(n), indexing="ij")
    xx = xx / (n - 1) * 2 - 1
    yy = yy / (n - 1) * 2 - 1
    rel = (bbox_rect[:, 2:] - bbox_rect[:, :2]) / (bbox[:, 2:] - bbox[:, :2])
    xx = xx * rel[:, 0][:, None, None]
    yy = yy * rel[:, 1][:, None, None]
    return F.grid_sample(hm, torch.stack((xx, yy), dim=-1), mode="bilinear", align_corners=False)


def flip_by_mask(t, is_flip):
    """flip_tensor_by_mask_index (VPHO.py:349-357): flip the last axis of the flagged images."""
    return torch.stack([t[i].flip(-1) if bool(f) else t[i] for i, f in enumerate(is_flip)], dim=0)


def residual(x, st, p):
    """Residual(numIn == numOut) (encoding.py:21-36); LeakyReLU default slope 0.01."""
    out = F.leaky_relu(_bn(x, st, p + ".bn"), 0.01)
    out = _conv(out, st, p + ".conv1", 0)
    out = F.leaky_relu(_bn(out, st, p + ".bn1"), 0.01)
    out = _conv(out, st, p + ".conv2", 1)
    out = F.leaky_relu(_bn(out, st, p + ".bn2"), 0.01)
    out = _conv(out, st, p + ".conv3", 0)
    res = _conv(x, st, p + ".conv4", 0) if (p + ".conv4.weight") in st else x
    return out + res


def encoder(x, st, p, n_block=4, n_mod=2):
    """Encoder.forward (encoding.py:58-73) -> (flattened (bs, 1024), [pooled maps])."""
    x = _conv(x, st, p + ".project", 0)
    xs = []
    for i in range(n_block):
        for j in range(n_mod):
            x = residual(x, st, f"{p}.reg.{i * n_mod + j}")
        x = F.max_pool2d(x, 2, 2)
        xs.append(x)
    return x.flatten(1), xs


def head_mano_regression(enc, st, p="head_mano"):
    """HeadMano.forward (head_mano.py:61-76) -> (axis-angle pose (bs, 48), shape (bs, 10))."""
    h = F.leaky_relu(_lin(enc, st, p + ".base_layer.0"), 0.01)
    h = F.leaky_relu(_lin(h, st, p + ".base_layer.2"), 0.01)
    r6 = _lin(h, st, p + ".fc_pose").reshape(enc.shape[0], -1, 6)
    aa = matrix_to_axis_angle(rotation_6d_to_matrix(r6)).reshape(enc.shape[0], -1)
    return aa, _lin(h, st, p + ".fc_shape")


def gravity_embedding(g, multires=10):
    """PosEmbedder(3, 10).embed (cross_module.py:9-46): [x, sin(x f0), cos(x f0), sin(x f1), ...], f = 2^0 .. 2^9."""
    out = [g]
    for f in 2.0 ** torch.linspace(0.0, multires - 1, steps=multires):
        out += [torch.sin(g * f), torch.cos(g * f)]
    return torch.cat(out, -1)


def positional_table(n, d):
    """PositionalEncoding.pe[:n, 0] (cross_module.py:70-78)."""
    pe = torch.zeros(n, d)
    pos = torch.arange(0, n, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def encoder_layer(x, st, p, nhead=2):
    """nn.TransformerEncoderLayer (post-norm, ReLU, dim_feedforward 2048, eval) on x (L, N, E) -- L is the attended axis."""
    L, N, E = x.shape
    hd = E // nhead
    qkv = F.linear(x, st[p + ".self_attn.in_proj_weight"], st[p + ".self_attn.in_proj_bias"])
    q, k, v = qkv.split(E, dim=-1)

    def heads(t):                                   # (L, N, E) -> (N, nhead, L, hd), as F.multi_head_attention_forward
        return t.reshape(L, N * nhead, hd).transpose(0, 1).reshape(N, nhead, L, hd)
    o = F.scaled_dot_product_attention(heads(q), heads(k), heads(v))          # softmax(q k^T / sqrt(hd)) v
    o = o.permute(2, 0, 1, 3).reshape(L, N, E)
    o = _lin(o, st, p + ".self_attn.out_proj")
    x = F.layer_norm(x + o, (E,), st[p + ".norm1.weight"], st[p + ".norm1.bias"], LN_EPS)
    f = _lin(F.relu(_lin(x, st, p + ".linear1")), st, p + ".linear2")
    return F.layer_norm(x + f, (E,), st[p + ".norm2.weight"], st[p + ".norm2.bias"], LN_EPS)


def cross_module(x_hand, x_obj, gravity, st, p, num_force=32):
    """CrossModule.forward (cross_module.py:119-137) -> (y_hand, y_obj, y_gravity)."""
    bs = x_hand.shape[0]
    xh = _conv(x_hand, st, p + ".proj_hand", 1).view(bs, num_force, -1)
    xo = _conv(x_obj, st, p + ".proj_obj", 1).view(bs, num_force, -1)
    g = _lin(gravity_embedding(gravity), st, p + ".gravity_proj")
    x = torch.cat([xh, xo, g], dim=1)                                   # (bs, 65, 512)
    x = x + positional_table(bs, x.shape[-1])[:, None, :]                 # pe[:x.size(0)] with x.size(0) == bs
    x = encoder_layer(x, st, p + ".attn.layers.0")
    return torch.split(x, [num_force, num_force, 1], dim=1)


def get_local_force(scale, weight, friction=0.8, num_anchor=8):
    """HeadForce2.get_local_force (physics.py:546-557) with the anchors of HeadPhysics.__init__ (:692-698)."""
    a = torch.arange(0, 2 * torch.pi, 2 * torch.pi / num_anchor)[:num_anchor]
    anchor = torch.stack([torch.cos(a), torch.sin(a), torch.ones_like(a)], dim=-1) / num_anchor
    anchor = anchor.clone()
    anchor[:, :2] *= friction
    w = torch.softmax(weight, dim=-1)
    d = torch.einsum("...j,jk->...k", w, anchor)
    d = d / (d.norm(dim=-1, keepdim=True) + 1e-8)
    return d * torch.abs(scale)[..., None]


def head_physics(x_hand, x_obj, st, p="head_physics"):
    """HeadPhysics.forward (physics.py:700-721)."""
    def mlp(x, q):
        return _lin(F.leaky_relu(_lin(x, st, f"{p}.{q}.0"), 0.01), st, f"{p}.{q}.2")
    scale = mlp(x_hand, "fc_scale").squeeze(-1)
    weight = torch.softmax(mlp(x_obj, "fc_weight"), dim=-1)
    return {"force_local": get_local_force(scale, weight), "scale": scale, "weight": weight, "CoM": mlp(x_obj, "fc_CoM")}


@torch.no_grad()
def oracle_producers(st: Dict[str, torch.Tensor], hf_hr, of_or_rect, hf_hr_rect, bbox_hand, bbox_hand_rect, bbox_obj,
                     bbox_obj_rect, is_right, gravity) -> Dict[str, torch.Tensor]:
    """VPHO.py:129-178 from the RoI-aligned maps (bs, 256, 32, 32) to everything the predict branch consumes.
    (`of_or`, the tight object RoI, only lends its spatial size to F.interpolate there and is not an input.)"""
    st = {k: torch.as_tensor(v) for k, v in st.items()}
    as_t = lambda a: torch.as_tensor(a)                                                       # noqa: E731
    hf_hr, of_or_rect, hf_hr_rect = as_t(hf_hr).float(), as_t(of_or_rect).float(), as_t(hf_hr_rect).float()
    bbox_hand, bbox_hand_rect = as_t(bbox_hand).float(), as_t(bbox_hand_rect).float()
    bbox_obj, bbox_obj_rect = as_t(bbox_obj).float(), as_t(bbox_obj_rect).float()
    is_right, gravity = as_t(is_right).bool(), as_t(gravity).float()
    flip = ~is_right
    hm_hand = head_heatmap2(hf_hr, st, "head_hm_hand")
    hm_obj = head_heatmap2(of_or_rect, st, "head_hm_obj")
    hm_hand_rect = align_hm_to_bbox_rectangle(hm_hand, bbox_hand, bbox_hand_rect)
    hm_obj_rect = align_hm_to_bbox_rectangle(hm_obj, bbox_obj, bbox_obj_rect)
    of_flip = flip_by_mask(of_or_rect, flip)
    hm_obj_rect_ori = flip_by_mask(hm_obj_rect, flip)
    size = hf_hr.shape[-2:]
    hm_hand_rs = F.interpolate(hm_hand_rect, size=size, mode="bilinear", align_corners=False)
    hm_obj_rs = F.interpolate(hm_obj_rect_ori, size=size, mode="bilinear", align_corners=False)
    enc_hand, hand_ls = encoder(torch.cat((hf_hr_rect, hm_hand_rs), dim=1), st, "encoder_hand")
    enc_obj, obj_ls = encoder(torch.cat((of_flip, hm_obj_rs), dim=1), st, "encoder_obj")
    pose, shape = head_mano_regression(enc_hand, st)
    g = gravity.clone()
    g[flip, ..., 0] *= -1                                                                   # flip_point3d_by_mask_index
    phy_hand, _, _ = cross_module(hand_ls[1], obj_ls[1], g, st, "cross_hand")
    _, phy_obj, _ = cross_module(hand_ls[1], obj_ls[1], g, st, "cross_obj")
    phy = head_physics(phy_hand, phy_obj, st)
    return {"hand_heatmap": hm_hand, "obj_heatmap": hm_obj, "encoding_hand": enc_hand, "encoding_obj": enc_obj,
            "mano_pose": pose, "mano_shape": shape, "force_local": phy["force_local"], "force_scale": phy["scale"],
            "force_weight": phy["weight"], "CoM": phy["CoM"], "enc_phy_hand": phy_hand, "enc_phy_obj": phy_obj,
            "hm_hand_rs": hm_hand_rs, "hm_obj_rs": hm_obj_rs}
