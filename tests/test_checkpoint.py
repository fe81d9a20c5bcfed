"""`vpho_b200.checkpoint`: the denoiser weights of what the reference's trainer writes (`accel.save_state` directory with
model.safetensors / pytorch_model.bin, or `final_model.pt`; lib/engine/base_trainer.py:81-96)."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_loader as RL
from vpho_b200 import checkpoint as ck
from vpho_b200 import synthetic as syn


def _full_state(prefix=""):
    """a vpho_net-shaped state dict: both denoisers plus a few of the modules that are not on the path"""
    st = {}
    for name, head, seed in (("denoiser_hand", "mano_pose", 1), ("denoiser_obj", "obj", 2)):
        for k, v in syn.make_denoiser_state(head, seed).items():
            st[f"{prefix}{name}.{k}"] = torch.from_numpy(v)
    st[prefix + "head_hm_hand.conv.weight"] = torch.randn(4, 4)
    st[prefix + "feature_extractor.layer.bias"] = torch.randn(7)
    return st


@pytest.mark.parametrize("kind", ["safetensors_dir", "bin_dir", "final_model", "ddp_prefix"])
def test_load_denoiser_states(tmp_path, kind):
    st = _full_state("module." if kind == "ddp_prefix" else "")
    if kind == "safetensors_dir":
        from safetensors.torch import save_file
        d = tmp_path / "epoch_3.state"
        d.mkdir()
        save_file(st, str(d / "model.safetensors"))
        (d / "optimizer.bin").write_bytes(b"x")          # accelerate also writes optimizer / rng files next to it
        path = str(d)
    elif kind == "bin_dir":
        d = tmp_path / "epoch_3.state"
        d.mkdir()
        torch.save(st, str(d / "pytorch_model.bin"))
        path = str(d)
    else:
        path = str(tmp_path / "final_model.pt")
        torch.save(st, path)
    hand, obj, rest = ck.load_denoiser_states(path)
    for got, head, seed in ((hand, "mano_pose", 1), (obj, "obj", 2)):
        want = syn.make_denoiser_state(head, seed)
        assert set(got) == set(want) == set(ck.DENOISER_KEYS)
        for k in want:
            assert got[k].dtype == np.float32 and got[k].flags["C_CONTIGUOUS"] and np.array_equal(got[k], want[k]), k
    assert set(rest) == {"head_hm_hand.conv.weight", "feature_extractor.layer.bias"}


def test_errors(tmp_path):
    with pytest.raises(ck.CheckpointError):
        ck.load_denoiser_states(str(tmp_path / "nothing"))
    (tmp_path / "empty.state").mkdir()
    with pytest.raises(ck.CheckpointError):
        ck.load_denoiser_states(str(tmp_path / "empty.state"))
    st = _full_state()
    del st["denoiser_obj.head.head.2.bias"]
    torch.save(st, str(tmp_path / "a.pt"))
    with pytest.raises(ck.CheckpointError, match="missing"):
        ck.load_denoiser_states(str(tmp_path / "a.pt"))
    st = _full_state()
    st["denoiser_hand.pose_encoder.0.weight"] = torch.zeros(256, 95)
    torch.save(st, str(tmp_path / "b.pt"))
    with pytest.raises(ck.CheckpointError, match="shape"):
        ck.load_denoiser_states(str(tmp_path / "b.pt"))
    st = _full_state()
    st["denoiser_hand.t_encoder.1.bias"][3] = float("nan")
    torch.save(st, str(tmp_path / "c.pt"))
    with pytest.raises(ck.CheckpointError, match="non-finite"):
        ck.load_denoiser_states(str(tmp_path / "c.pt"))


@pytest.mark.skipif(not RL.reference_available(), reason="/root/reference is not present on this box")
def test_key_names_are_the_reference_modules():
    """The reference's own BaseDenoiser state_dict has exactly the keys / shapes the loader expects, under the attribute names
    vpho_net gives the two denoisers (lib/model/VPHO.py:58-59)."""
    import re
    from oracle import cases
    mano, anch, objs = cases.assets()
    ref = RL.load_reference(mano, anch, objs)
    src = open(os.path.join(RL.REFERENCE_ROOT, "lib/model/VPHO.py")).read()
    assert re.search(r"self\.denoiser_hand\s*=\s*BaseDenoiser\(", src) and re.search(r"self\.denoiser_obj\s*=\s*BaseDenoiser\(", src)
    marg = ref.sbm.ScoreBasedModelAgent().marginal_prob_fn
    for head, D, n in (("mano_pose", 96, 32), ("obj", 9, 3)):
        sd = ref.denoiser.BaseDenoiser(marg, head=head).state_dict()
        assert set(sd) == set(ck.DENOISER_KEYS)
        assert tuple(sd["head.head.0.weight"].shape) == (n, 1408, 256) and tuple(sd["pose_encoder.0.weight"].shape) == (256, D)
        # and a checkpoint written from the reference modules loads
    full = {}
    for name, head in (("denoiser_hand", "mano_pose"), ("denoiser_obj", "obj")):
        full.update({f"{name}.{k}": v for k, v in ref.denoiser.BaseDenoiser(marg, head=head).state_dict().items()})
    hand, obj, rest = ck.split_state(full)
    assert not rest and hand["head.head.2.weight"].shape == (32, 256, 3) and obj["head.head.2.weight"].shape == (3, 256, 3)
