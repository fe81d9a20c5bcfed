"""Object-pose metrics on the device (`ObjectMetrics` <- TesterObject, lib/engine/test.py:196-584; C ABI
`vpho_object_metrics`) against (1) the fixture minted from the reference's OWN TesterObject methods
(tests/golden/object_metrics.npz, oracle/make_golden.py) and (2) the travelling restatement oracle/object_metrics.py.

Tolerances: float64 columns (MCE, OCE, SMCE, REP) 1e-9 relative; float32 columns (MCE2, ADD, ADD-S, CD) 2e-6 relative
(different FP32 summation order over 2048 points); F-scores within 1.5 / n_points (a point whose nearest distance sits
within one FP32 ulp of a threshold may fall on either side); the 0/1 flags exact away from their thresholds."""
import os

import numpy as np
import pytest
import torch

from oracle import cases
from oracle import object_metrics as OM
from vpho_b200.aggregation import Assets, ObjectMetrics, obj_6d_to_rt

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "object_metrics.npz")
F64_COLS, F32_COLS, FS_COLS, FLAG_COLS = [0, 1, 3, 6], [2, 4, 5, 7], list(range(8, 14)), [14, 15, 16]


def _compare(ours: np.ndarray, ref: np.ndarray, n_pts: int, diameter: np.ndarray):
    rel = lambda a, b: np.abs(a - b) / np.maximum(np.abs(b), 1e-12)   # noqa: E731
    assert rel(ours[..., F64_COLS], ref[..., F64_COLS]).max() < 1e-9
    assert rel(ours[..., F32_COLS], ref[..., F32_COLS]).max() < 2e-6
    assert np.abs(ours[..., FS_COLS] - ref[..., FS_COLS]).max() <= 1.5 / n_pts
    # flags: exact unless the value sits within 1e-5 relative of its threshold
    thr = np.stack([diameter * 0.1, diameter * 0.1, np.full_like(diameter, 5.0)], -1)[:, None, :]
    val = ref[..., [4, 5, 6]]
    safe = np.abs(val - thr) > 1e-5 * thr
    assert (ours[..., FLAG_COLS][safe] == ref[..., FLAG_COLS][safe]).all()


def _run(lib, dev, inp=None):
    mano, anch, objs = cases.assets()
    tables = OM.synthetic_metric_tables(objs)
    inp = inp or cases.object_metric_case()
    om = ObjectMetrics(Assets(anch, objs, lib=lib), tables)
    T = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    out = om(T(inp["pd_rt"]), T(inp["gt_rt"]), T(inp["obj_id"]), T(inp["cam_intr"]))
    return out.cpu().numpy(), tables, inp, objs


def _check_all(lib, dev):
    ours, tables, inp, objs = _run(lib, dev)
    assert ours.shape == (6, 5, 17) and ours.dtype == np.float64
    g = np.load(GOLD)
    assert abs(cases.fingerprint(inp["pd_rt"], inp["gt_rt"]) - float(g["fp"])) < 1e-6
    diam = tables["diameter"][inp["obj_id"]].astype(np.float64)
    _compare(ours, g["metrics"], objs["verts_sampled"].shape[1], diam)                     # the reference's own numbers
    ref = OM.object_metrics(tables, inp["pd_rt"], inp["gt_rt"], inp["obj_id"], inp["cam_intr"])
    assert np.array_equal(ref, g["metrics"])                                                # oracle pinned to the fixture
    # every F-score threshold and flag is exercised in both directions by the fixture
    assert (g["metrics"][..., FS_COLS].min(axis=(0, 1)) < 0.5).all() and (g["metrics"][..., FS_COLS].max(axis=(0, 1)) > 0.5).all()
    assert set(np.unique(g["metrics"][..., 14])) == {0.0, 1.0}


def test_object_metrics_oracle_matches_reference_fixture():
    mano, anch, objs = cases.assets()
    tables = OM.synthetic_metric_tables(objs)
    inp = cases.object_metric_case()
    g = np.load(GOLD)
    ref = OM.object_metrics(tables, inp["pd_rt"], inp["gt_rt"], inp["obj_id"], inp["cam_intr"])
    assert np.array_equal(ref, g["metrics"])
    assert tables["sym_R"].shape[1] == 628          # identity x 314 continuous steps x one discrete half-turn (object 5)


def test_product_symmetry_tables_equal_the_oracle_tables():
    # vpho_b200.evaluation.symmetry_tables (what a user feeds ObjectMetrics) vs the oracle restatement, itself checked
    # against TesterObject.R / .t of the reference in the build container (oracle/make_golden.py)
    from vpho_b200 import synthetic as syn
    mano, anch, objs = cases.assets()
    a, b = syn.make_metric_tables(objs), OM.synthetic_metric_tables(objs)
    assert np.abs(a["sym_R"] - b["sym_R"]).max() < 1e-15 and np.abs(a["sym_t"] - b["sym_t"]).max() < 1e-18
    assert np.array_equal(a["sym_count"], b["sym_count"]) and np.array_equal(a["bbox3d"], b["bbox3d"])
    assert np.array_equal(a["diameter"], b["diameter"])


def test_object_metrics_emulated(emu_lib):
    inp = cases.object_metric_case()
    small = {k: v[:2, :2] if k == "pd_rt" else v[:2] for k, v in inp.items()}      # the emulator is slow: 2 images x 2 candidates
    ours, tables, _, objs = _run(emu_lib, "cpu", small)
    g = np.load(GOLD)["metrics"][:2, :2]
    _compare(ours, g, objs["verts_sampled"].shape[1], tables["diameter"][small["obj_id"]].astype(np.float64))


@pytest.mark.gpu
def test_object_metrics_cuda(cuda_lib):
    _check_all(None, "cuda")


@pytest.mark.gpu
def test_object_metrics_cuda_readme_batch(cuda_lib):
    """64 images x (100 finals + the aggregate): what one evaluate() batch needs; against the oracle on a sample of rows,
    and the single-candidate call form."""
    mano, anch, objs = cases.assets()
    tables = OM.synthetic_metric_tables(objs)
    om = ObjectMetrics(Assets(anch, objs), tables)
    g = torch.Generator().manual_seed(3)
    n, C = 64, 101
    gt6 = torch.cat([torch.randn(n, 6, generator=g, dtype=torch.float64), torch.randn(n, 3, generator=g, dtype=torch.float64) * 0.05], -1)
    pd6 = gt6[:, None] + torch.randn(n, C, 9, generator=g, dtype=torch.float64) * torch.tensor([0.3] * 6 + [0.02] * 3, dtype=torch.float64)
    root = torch.tensor([0.0, 0.0, 0.6], dtype=torch.float64).repeat(n, 1)
    pd_rt, gt_rt = obj_6d_to_rt(pd6, root), obj_6d_to_rt(gt6, root)
    ids = torch.randint(0, 21, (n,), generator=g, dtype=torch.int32)
    K = torch.tensor([[600.0, 0, 128], [0, 600.0, 128], [0, 0, 1]]).repeat(n, 1, 1)
    out = om(pd_rt.cuda(), gt_rt.cuda(), ids.cuda(), K.cuda()).cpu().numpy()
    assert out.shape == (n, C, 17) and np.isfinite(out).all()
    rows = [0, 17, 63]
    ref = OM.object_metrics(tables, pd_rt[rows].numpy(), gt_rt[rows].numpy(), ids[rows].numpy(), K[rows].numpy())
    _compare(out[rows], ref, objs["verts_sampled"].shape[1], tables["diameter"][ids[rows].numpy()].astype(np.float64))
    one = om(pd_rt[:, 5].cuda(), gt_rt.cuda(), ids.cuda(), K.cuda()).cpu().numpy()
    assert one.shape == (n, 17) and np.array_equal(one, out[:, 5])
