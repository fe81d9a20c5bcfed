"""Diagnostics of the tensor-core MANO blend (VPHO_MANO_DEBUG_BLEND): compares the blended rest pose with a float64 torch
evaluation, per coordinate / vertex tile / coefficient group."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vpho_b200 import capi, synthetic as syn  # noqa: E402
from vpho_b200.head_mano import HeadMano  # noqa: E402

m = syn.make_mano_model()
hm = HeadMano(m)
lib = hm.lib


def rodrigues(aa):          # manopth batch_rodrigues (quaternion form)
    ang = (aa + 1e-8).norm(dim=1, keepdim=True)
    nrm = aa / ang
    q = torch.cat([torch.cos(ang / 2), torch.sin(ang / 2) * nrm], 1)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([w*w+x*x-y*y-z*z, 2*x*y-2*w*z, 2*w*y+2*x*z, 2*w*z+2*x*y, w*w-x*x+y*y-z*z, 2*y*z-2*w*x,
                        2*x*z-2*w*y, 2*w*x+2*y*z, w*w-x*x-y*y+z*z], 1)


def run(n, pose, shape, flags):
    v = torch.empty((n, 778, 3), device="cuda")
    j = torch.empty((n, 21, 3), device="cuda")
    lib.check(lib.c.vpho_mano_forward_ex(hm.handle, capi.ptr(pose), capi.ptr(shape), n, capi.ptr(v), capi.ptr(j), flags, None), "mano")
    torch.cuda.synchronize()
    return v.cpu().double(), j.cpu().double()


def ref_blend(pose, shape):
    n = pose.shape[0]
    R = rodrigues(pose.double().reshape(-1, 3)).reshape(n, 16, 9)
    pm = (R[:, 1:] - torch.eye(3, dtype=torch.float64).reshape(1, 1, 9)).reshape(n, 135)
    sd = torch.from_numpy(m["shapedirs"]).double()
    pdirs = torch.from_numpy(m["posedirs"]).double()
    vt = torch.from_numpy(m["v_template"]).double()
    return vt[None] + torch.einsum("vdk,nk->nvd", sd, shape.double()) + torch.einsum("vdk,nk->nvd", pdirs, pm)


g = torch.Generator().manual_seed(0)
for n, case in ((1, "full"), (3, "shape_only"), (3, "pose_only"), (70, "full"), (6400, "full")):
    pose = torch.randn(n, 48, generator=g) * 0.5
    shape = torch.randn(n, 10, generator=g)
    if case == "shape_only":
        pose = pose * 0
    if case == "pose_only":
        shape = shape * 0
    pc, sc_ = pose.cuda().contiguous(), shape.cuda().contiguous()
    vb, _ = run(n, pc, sc_, 2)
    rb = ref_blend(pose, shape)
    err = (vb - rb).abs()
    print(f"[{case} n={n}] blend max err {err.max():.3e} (blend offsets max {((rb - torch.from_numpy(m['v_template']).double()).abs().max()):.3e})")
    print("   per coord:", [f"{err[..., d].max():.2e}" for d in range(3)], " per vtile:", [f"{err[:, t*128:(t+1)*128].max():.2e}" for t in range(7)])
    print("   per candidate (first 8):", [f"{err[c].max():.2e}" for c in range(min(n, 8))], " worst cand", int(err.reshape(n, -1).max(1)[0].argmax()))
    if err.max() > 1e-6:
        # is it a pure scale / subset problem?  regress our offsets on the reference offsets
        ours_off = (vb - torch.from_numpy(m["v_template"]).double())[0].reshape(-1)
        ref_off = (rb - torch.from_numpy(m["v_template"]).double())[0].reshape(-1)
        print("   cand 0: <ours,ref>/<ref,ref> =", float((ours_off @ ref_off) / (ref_off @ ref_off)), " |ours|/|ref| =", float(ours_off.norm() / ref_off.norm()))
    v1, j1 = run(n, pc, sc_, 0)
    v2, j2 = run(n, pc, sc_, 1)
    print(f"   TC vs SIMT: verts {float((v1 - v2).abs().max()):.3e} joints {float((j1 - j2).abs().max()):.3e}")
# timing
for n in (6400,):
    pose = (torch.randn(n, 48, generator=g) * 0.5).cuda()
    shape = torch.randn(n, 10, generator=g).cuda()
    for flags in (0, 1):
        v = torch.empty((n, 778, 3), device="cuda")
        j = torch.empty((n, 21, 3), device="cuda")
        for _ in range(3):
            lib.c.vpho_mano_forward_ex(hm.handle, capi.ptr(pose), capi.ptr(shape), n, capi.ptr(v), capi.ptr(j), flags, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.c.vpho_mano_forward_ex(hm.handle, capi.ptr(pose), capi.ptr(shape), n, capi.ptr(v), capi.ptr(j), flags, None)
        e1.record()
        torch.cuda.synchronize()
        print(f"n={n} flags={flags}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
