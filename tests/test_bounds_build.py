"""The -DVPHO_DEBUG_BOUNDS twin of the library (`python -m vpho_b200.build --bounds` -> libvpho_b200_bounds.so): every
global-memory index formed by the tensor-core kernels (k_pose_tc, k_head_tc, k_mano_tc) and the K-slot writes is asserted on
the device and a violation traps.  compute-sanitizer cannot run tcgen05 kernels, so this is the out-of-bounds check of the
GPU suite: the hot path at ragged shapes (rows not a multiple of the 128-row tile, odd image counts, a final partial MANO
tile) must run clean AND give bit-for-bit the product library's results."""
import ctypes
import os
import re

import pytest
import torch

from oracle import cases
from vpho_b200 import build as B
from vpho_b200 import capi
from vpho_b200 import synthetic as syn


def test_bounds_twin_exports_the_same_abi():
    if not os.path.exists(B.LIB_BOUNDS):
        pytest.skip("libvpho_b200_bounds.so not built here (python -m vpho_b200.build --bounds)")
    lib = ctypes.CDLL(B.LIB_BOUNDS)
    hdr = open(os.path.join(B.ROOT, "include", "vpho_b200.h")).read()
    missing = [n for n in sorted(set(re.findall(r"\b(vpho_[a-z0-9_]+)\s*\(", hdr))) if not hasattr(lib, n)]
    assert not missing, missing


@pytest.mark.gpu
@pytest.mark.parametrize("bs,S,steps,Kh,Ko", [(3, 100, 10, 30, 10), (1, 37, 6, 7, 5), (5, 16, 5, 6, 4)])
def test_hot_path_runs_clean_under_bounds_asserts(cuda_lib, bs, S, steps, Kh, Ko):
    from vpho_b200.vpho import VphoHotPath, to_device
    assert os.path.exists(B.LIB_BOUNDS), "build the twin with `python -m vpho_b200.build --bounds` (build() of __graft_entry__ does)"
    twin = capi.Library(B.LIB_BOUNDS)
    mano, anch, objs = cases.assets()
    batch = syn.make_eval_batch(bs, seed=bs, sample_num=S, mano=mano, objects=objs)
    st_h, st_o = syn.make_denoiser_state("mano_pose", 0), syn.make_denoiser_state("obj", 0)
    ph, po = cases.e2e_priors("clustered", bs, S, batch, bs)
    outs = []
    for lib in (None, twin):
        hp = VphoHotPath(mano, anch, objs, st_h, st_o, sample_num=S, sampling_steps=steps, topk_hand=Kh, topk_obj=Ko, lib=lib)
        pd = hp.predict(to_device(batch, "cuda"), prior_hand=ph, prior_obj=po)
        torch.cuda.synchronize()          # a trapped assert surfaces here as a CUDA error
        outs.append(pd)
    for k in ("diff_final_hand_mano", "diff_final_obj_6d", "diff_final_hand_vert", "diff_final_hand_joint", "agg_hand_vert",
              "agg_hand_joint", "agg_obj_6d", "diff_inprocess_hand_mano"):
        assert torch.equal(outs[0][k], outs[1][k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 63, 64, 65, 777])
def test_mano_tc_runs_clean_under_bounds_asserts(cuda_lib, n):
    from vpho_b200.head_mano import HeadMano
    twin = capi.Library(B.LIB_BOUNDS)
    mano, _, _ = cases.assets()
    g = torch.Generator().manual_seed(n)
    pose, shape = (torch.randn(n, 48, generator=g) * 0.4).cuda(), torch.randn(n, 10, generator=g).cuda()
    v0, j0 = HeadMano(mano).get_hand_verts(pose=pose, shape=shape)
    v1, j1 = HeadMano(mano, lib=twin).get_hand_verts(pose=pose, shape=shape)
    torch.cuda.synchronize()
    assert torch.equal(v0, v1) and torch.equal(j0, j1)
