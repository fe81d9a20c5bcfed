"""Load the hot path's weights from what the reference's trainer writes.

`BaseTrainer.save_checkpoint` (lib/engine/base_trainer.py:85-89) calls `accel.save_state(dir)`: accelerate writes the
prepared model's `state_dict()` as `<dir>/model.safetensors` (or `pytorch_model.bin` with safe_serialization off; with several
prepared models `model_1.safetensors`, ...).  `BaseTrainer.save_model` (:91-96) writes `final_model.pt` =
`torch.save(model.state_dict())`.  `load_checkpoint` (:81-83) reads them back with `strict=False`.

The model is `vpho_net` (lib/model/VPHO.py:48-84); the hot path needs two of its sub-modules:

    denoiser_hand.*   BaseDenoiser(head='mano_pose')   (lib/model/denoiser.py:33-66)
    denoiser_obj.*    BaseDenoiser(head='obj')

Everything else in the file (feature extractor, heat-map / regression heads, cross modules) produces the path's INPUTS and is
returned untouched under `rest` for the caller that runs those modules.  A DistributedDataParallel / torch.compile wrapper
prefix (`module.`, `_orig_mod.`) is stripped.  Nothing here touches the GPU.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np
import torch

DENOISER_KEYS = ("t_encoder.0.W", "t_encoder.1.weight", "t_encoder.1.bias", "pose_encoder.0.weight", "pose_encoder.0.bias",
                 "pose_encoder.2.weight", "pose_encoder.2.bias", "head.head.0.weight", "head.head.0.bias", "head.head.2.weight",
                 "head.head.2.bias")
_WRAPPERS = ("module.", "_orig_mod.")


class CheckpointError(RuntimeError):
    pass


def _read_file(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device="cpu")
    obj = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(obj, dict) and "state_dict" in obj and isinstance(obj["state_dict"], dict):
        obj = obj["state_dict"]
    if not isinstance(obj, dict):
        raise CheckpointError(f"{path}: expected a state dict, got {type(obj).__name__}")
    return obj


def find_model_file(path: str) -> str:
    """`path`: an `accel.save_state` directory, or a file (`final_model.pt`, `model.safetensors`, `pytorch_model.bin`)."""
    if os.path.isfile(path):
        return path
    if not os.path.isdir(path):
        raise CheckpointError(f"{path}: no such checkpoint")
    for name in ("model.safetensors", "pytorch_model.bin", "final_model.pt"):
        f = os.path.join(path, name)
        if os.path.isfile(f):
            return f
    raise CheckpointError(f"{path}: no model.safetensors / pytorch_model.bin / final_model.pt inside")


def _strip(key: str) -> str:
    changed = True
    while changed:
        changed = False
        for w in _WRAPPERS:
            if key.startswith(w):
                key, changed = key[len(w):], True
    return key


def split_state(state: Dict[str, torch.Tensor]) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray], Dict[str, torch.Tensor]]:
    """-> (denoiser_hand state, denoiser_obj state, rest) with the sub-module prefix removed from the first two and their
    tensors as contiguous float32 numpy arrays (what `Denoiser(state)` takes).  Shapes are checked against the module
    definitions (denoiser.py:33-66, parallel_linear.py:10-25): the hand denoiser has 32 heads over 96 inputs, the object
    denoiser 3 heads over 9."""
    hand, obj, rest = {}, {}, {}
    for k, v in state.items():
        k = _strip(k)
        if k.startswith("denoiser_hand."):
            hand[k[len("denoiser_hand."):]] = v
        elif k.startswith("denoiser_obj."):
            obj[k[len("denoiser_obj."):]] = v
        else:
            rest[k] = v

    def finish(sub, name, D, n):
        missing = [k for k in DENOISER_KEYS if k not in sub]
        if missing:
            raise CheckpointError(f"{name}: missing keys {missing}")
        out = {k: np.ascontiguousarray(sub[k].detach().to(torch.float32).cpu().numpy()) for k in DENOISER_KEYS}
        want = {"t_encoder.0.W": (64,), "t_encoder.1.weight": (128, 128), "t_encoder.1.bias": (128,),
                "pose_encoder.0.weight": (256, D), "pose_encoder.0.bias": (256,), "pose_encoder.2.weight": (256, 256),
                "pose_encoder.2.bias": (256,), "head.head.0.weight": (n, 1408, 256), "head.head.0.bias": (n, 256),
                "head.head.2.weight": (n, 256, 3), "head.head.2.bias": (n, 3)}
        for k, shp in want.items():
            if tuple(out[k].shape) != shp:
                raise CheckpointError(f"{name}.{k}: shape {tuple(out[k].shape)}, expected {shp}")
            if not np.isfinite(out[k]).all():
                raise CheckpointError(f"{name}.{k}: non-finite values")
        return out
    return finish(hand, "denoiser_hand", 96, 32), finish(obj, "denoiser_obj", 9, 3), rest


def load_denoiser_states(path: str):
    """-> (denoiser_hand state, denoiser_obj state, rest) from an `accel.save_state` directory or a state-dict file."""
    return split_state(_read_file(find_model_file(path)))


def hot_path_from_checkpoint(path: str, mano_model: Dict, anchors: Dict, objects: Dict, **kwargs):
    """`VphoHotPath` with the denoisers of a reference checkpoint; kwargs as `VphoHotPath(...)` (sample_num, sampling_steps,
    sample_T0, topk_hand, topk_obj: cfg.sample_num / sampling_steps / sample_T0 / topk_hand / topk_obj of the reference's
    yaml).  The MANO model, anchor tables and object assets are files of their own in the reference too
    (asset/mano_v1_2, asset/anchor, the YCB meshes) and are passed in."""
    from .vpho import VphoHotPath
    hand, obj, _ = load_denoiser_states(path)
    return VphoHotPath(mano_model, anchors, objects, hand, obj, **kwargs)
