"""N1: the feature-side producers -- heat-map heads, encoders, regression head, cross modules, physics head
(lib/model/VPHO.py:129-178).  CPU: the oracle restatement against the fixture minted from the reference's own classes
(oracle/make_golden_producers.py) and, live, against those classes; the CUDA sources on the SIMT emulator at toy widths.
GPU: the product library through the C ABI against the oracle at the reference's dimensions.

Tolerance (floating point, written here): every layer accumulates in FP32 in a different order than the CPU reference, over
up to ~30 chained layers; outputs are compared as max |ours - oracle| <= 2e-5 * max(1, max |oracle|) per tensor (measured:
~1e-6 on the emulator, a few 1e-6 on the GPU)."""
import numpy as np
import pytest
import torch

from oracle import producers as P
from oracle.make_golden_producers import BS, INPUT_SEED, OUT, STATE_SEED, pack
from oracle.reference_loader import reference_available
from vpho_b200 import synthetic as syn

REL = 2e-5
PAIRS = (("hand_heatmap", "hand_heatmap"), ("obj_heatmap", "obj_heatmap"), ("encoding_hand", "encoding_hand"),
         ("encoding_obj", "encoding_obj"), ("mano_pose", "mano_pose"), ("mano_shape", "mano_shape"),
         ("enc_phy_hand", "enc_phy_hand"), ("enc_phy_obj", "enc_phy_obj"), ("scale", "force_scale"), ("weight", "force_weight"),
         ("CoM", "CoM"), ("force_local", "force_local"))


def _compare(out, ref, rel=REL):
    worst = {}
    for k, rk in PAIRS:
        a, b = out[k].detach().cpu().double(), ref[rk].double()
        assert a.shape == b.shape, (k, a.shape, b.shape)
        assert torch.isfinite(a).all(), k
        err, scale = (a - b).abs().max().item(), max(1.0, b.abs().max().item())
        worst[k] = err / scale
        assert err <= rel * scale, (k, err, scale)
    return worst


def test_oracle_matches_reference_fixture():
    """oracle/producers.py reproduces what the reference's own module classes + glue produced (bit-identical on the
    convolutional chain; the transformer layer goes through the same F.scaled_dot_product_attention)."""
    g = np.load(OUT)
    assert int(g["bs"]) == BS
    st = syn.make_producer_state(STATE_SEED)
    inp = syn.make_producer_inputs(BS, INPUT_SEED)
    from oracle import cases
    assert abs(cases.fingerprint(inp["hf_hr"], st["encoder_obj.reg.7.conv3.weight"]) - float(g["fp"])) < 1e-6, "seeded generators drifted"
    mine = pack(P.oracle_producers(st, **inp))
    for k in g.files:
        if k in ("bs", "state_seed", "input_seed", "fp"):
            continue
        a, b = np.asarray(mine[k], np.float64), np.asarray(g[k], np.float64)
        assert a.shape == b.shape, k
        assert np.abs(a - b).max() <= 2e-6 * max(1.0, np.abs(b).max()), (k, np.abs(a - b).max())


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present")
def test_oracle_vs_reference_classes_live():
    from oracle import cases
    from oracle import make_golden_producers as G
    from oracle.reference_loader import load_reference
    mano, anch, objs = cases.assets()
    ref = load_reference(mano, anch, objs)
    st = syn.make_producer_state(5)
    inp = syn.make_producer_inputs(2, 9)
    theirs = G.reference_forward(G.reference_modules(ref, st), G.vpho_glue(), inp)
    mine = P.oracle_producers(st, **inp)
    for k in ("hand_heatmap", "obj_heatmap", "encoding_hand", "encoding_obj", "mano_pose", "mano_shape"):
        assert torch.equal(theirs[k], mine[k]), k                     # convolutional chain: bit-identical
    for k in ("enc_phy_hand", "enc_phy_obj", "force_local", "force_scale", "force_weight", "CoM"):
        assert (theirs[k] - mine[k]).abs().max().item() <= 2e-6 * max(1.0, theirs[k].abs().max().item()), k


def _run(lib, dims, bs, seed):
    from vpho_b200.producers import FeatureHeads
    st = syn.make_producer_state(seed, dims)
    inp = syn.make_producer_inputs(bs, seed + 1, roi=dims["roi"], C=dims["C"])
    dev = "cuda" if lib.path.endswith("libvpho_b200.so") else "cpu"
    T = {k: torch.from_numpy(np.asarray(v)).to(dev) for k, v in inp.items()}
    out = FeatureHeads(st, lib=lib)(T["hf_hr"], T["of_or_rect"], T["hf_hr_rect"], T, debug=True)
    if dev == "cuda":
        torch.cuda.synchronize()
    return out, P.oracle_producers(st, **inp)


def test_emulated_toy_chain(emu_lib):
    """The same CUDA sources (implicit-GEMM convolutions, transposed-convolution phases, attention, layer norm, physics tail)
    on the CPU emulator at toy widths, left and right hands mixed."""
    out, ref = _run(emu_lib, syn.PRODUCER_DIMS_TOY, 3, 3)
    _compare(out, ref)


def test_create_rejects_incomplete_state(emu_lib):
    from vpho_b200 import capi
    from vpho_b200.producers import FeatureHeads
    st = syn.make_producer_state(1, syn.PRODUCER_DIMS_TOY)
    del st["encoder_obj.reg.5.bn2.running_var"]
    with pytest.raises(capi.VphoError):
        FeatureHeads(st, lib=emu_lib)
    st = syn.make_producer_state(1, syn.PRODUCER_DIMS_TOY)
    st["head_mano.fc_shape.weight"] = st["head_mano.fc_shape.weight"][:, :-1].copy()
    with pytest.raises(capi.VphoError):
        FeatureHeads(st, lib=emu_lib)


@pytest.mark.gpu
@pytest.mark.parametrize("bs,seed", [(1, 0), (3, 0), (5, 7)])
def test_producers_cuda(cuda_lib, bs, seed):
    out, ref = _run(cuda_lib, syn.PRODUCER_DIMS, bs, seed)
    print("producers parity bs", bs, _compare(out, ref))


@pytest.mark.gpu
def test_producers_cuda_fixture(cuda_lib):
    """The CUDA path against the fixture minted from the reference's own classes."""
    out, _ = _run(cuda_lib, syn.PRODUCER_DIMS, BS, STATE_SEED)      # inputs: seed + 1 == INPUT_SEED
    assert INPUT_SEED == STATE_SEED + 1
    g = np.load(OUT)
    mine = pack({rk: out[k] for k, rk in PAIRS})
    for k in g.files:
        if k in ("bs", "state_seed", "input_seed", "fp"):
            continue
        a, b = np.asarray(mine[k], np.float64), np.asarray(g[k], np.float64)
        tol = REL * max(1.0, np.abs(b).max()) * (a.size if k.endswith("_sum") else 1)
        assert np.abs(a - b).max() <= tol, (k, np.abs(a - b).max())


@pytest.mark.gpu
def test_producers_cuda_headline_batch(cuda_lib):
    """bs = 64 (README batch): attention runs across the 64 images of the batch."""
    out, ref = _run(cuda_lib, syn.PRODUCER_DIMS, 64, 11)
    print("producers parity bs 64", _compare(out, ref))
