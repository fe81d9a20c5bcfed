"""Host-side mirror of the reference's score-based sampling interface over the CUDA sampler.

  * `Denoiser`              <- `BaseDenoiser` (lib/model/denoiser.py:33-82), evaluation only
  * `ScoreBasedModelAgent`  <- `ScoreBasedModelAgent` (lib/model/score_based_model.py:109-149): `sample(data, denoiser,
                               T0, init_x=None) -> (xs (N, steps, D) f64, x (N, D) f64)`, `get_score(data, denoiser)`

Everything numeric happens in `libvpho_b200.so` (csrc/sampler.cu); this file only allocates outputs, forwards
pointers and polls the device-side controller's status word.
"""
from __future__ import annotations

import collections
import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import capi

SIGMA_MIN, SIGMA_MAX, SAMPLING_EPS = 0.01, 50, 1e-5   # lib/model/sde.py:90-97 ('ve')
RTOL, ATOL, MAX_STEP = 3e-3, 3e-4, 10.0               # lib/model/score_based_model.py:51-52,91

_KEYS = ("t_encoder.0.W", "t_encoder.1.weight", "t_encoder.1.bias", "pose_encoder.0.weight", "pose_encoder.0.bias",
         "pose_encoder.2.weight", "pose_encoder.2.bias", "head.head.0.weight", "head.head.0.bias",
         "head.head.2.weight", "head.head.2.bias")


def ve_prior_std(T: float) -> float:
    """sigma(T) as a python float (lib/model/sde.py:15-18,26-28)."""
    return SIGMA_MIN * (SIGMA_MAX / SIGMA_MIN) ** T


class Denoiser:
    """Packed score network on the device.  `state` uses the reference's state-dict keys (denoiser.py:33-66)."""

    STRICT_FP32 = 1      # VPHO_DENOISER_STRICT_FP32 (include/vpho_b200.h)

    def __init__(self, state: Dict[str, object], lib: Optional[capi.Library] = None, strict_fp32: bool = False):
        """strict_fp32: FP32 SIMT kernels instead of the tcgen05 path (used to cross-check the tensor-core kernels).  The
        path is fixed at creation; creation FAILS when the tensor-core resources cannot be built (no silent downgrade)."""
        self.lib = lib or capi.lib()
        arrs = []
        for k in _KEYS:
            v = state[k]
            if isinstance(v, torch.Tensor):
                v = v.detach().cpu().numpy()
            arrs.append(np.ascontiguousarray(v, dtype=np.float32))
        n = arrs[7].shape[0]
        assert arrs[7].shape == (n, 1408, 256) and arrs[9].shape == (n, 256, 3) and arrs[3].shape == (256, 3 * n)
        self.n_heads = n
        self.out_dim = 3 * n
        h = C.c_void_p()
        self.lib.check(self.lib.c.vpho_denoiser_create_ex(n, *[capi.host_ptr(a) for a in arrs],
                                                         self.STRICT_FP32 if strict_fp32 else 0, C.byref(h)),
                       "vpho_denoiser_create_ex")
        self.handle = h
        self._ws: Dict[int, tuple] = {}
        self.calls = 0   # network evaluations issued by the last sample() (nfev + 1)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.c.vpho_denoiser_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def workspace(self, n_rows: int, rows_per_feat: int, n_eval: int, device, slot: int = 0) -> torch.Tensor:
        """One cached workspace per `slot` (a pipelined loop alternates two: the previous batch's integration may still
        have to be continued on its own workspace after the next batch's has been enqueued)."""
        key = (n_rows, rows_per_feat, n_eval, str(device))
        held = self._ws.get(slot)
        if held is None or held[0] != key:
            nbytes = self.lib.c.vpho_sample_workspace_bytes(self.n_heads, n_rows, rows_per_feat, n_eval)
            held = (key, torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device))
            self._ws[slot] = held   # keep one per slot
        return held[1]

    def __call__(self, data: dict) -> torch.Tensor:
        """`BaseDenoiser.forward` for the sampling call pattern: every row shares one time value."""
        x = data["sampled_pose"].contiguous().float()
        t = data["t"]
        t0 = float(t.reshape(-1)[0].item()) if isinstance(t, torch.Tensor) else float(t)
        feat, rpf = _unique_feat(data, x.shape[0])
        out = torch.empty_like(x)
        ws = self.workspace(x.shape[0], rpf, 1, x.device)
        st = self.lib.c.vpho_score_eval(self.handle, capi.ptr(x), C.c_float(t0), capi.ptr(feat), x.shape[0], rpf,
                                        capi.ptr(out), capi.ptr(ws), ws.numel(), capi.stream_of(x))
        self.lib.check(st, "vpho_score_eval")
        return out


def _unique_feat(data: dict, n_rows: int):
    """Returns (feat (R,1024) contiguous f32, rows_per_feat).  `data['feat_unique']` (one row per image) avoids the
    check; a plain repeated `data['feat']` (VPHO.py:239) is recognised by comparing each group with its first row."""
    if "feat_unique" in data:
        fu = data["feat_unique"].contiguous().float()
        assert n_rows % max(fu.shape[0], 1) == 0
        return fu, (n_rows // fu.shape[0] if fu.shape[0] else 1)
    feat = data["feat"].contiguous().float()
    assert feat.shape[0] == n_rows
    S = int(data.get("sample_num", 0) or 0)
    if S > 1 and n_rows % S == 0:
        g = feat.view(n_rows // S, S, -1)
        if bool((g == g[:, :1]).all().item()):
            return g[:, 0].contiguous(), S
    return feat, 1


SPARE_WINDOW = 8      # batches with an identical attempt count after which the spare RK attempt is no longer enqueued


class PendingSample:
    """Status of one enqueued sample(); `resolve` interprets the controller's counters once they are on the host."""

    def __init__(self, agent, denoiser, counters, n_rows, attempts_enqueued, keep_alive):
        self.agent, self.denoiser, self.counters, self.n_rows = agent, denoiser, counters, n_rows
        self.attempts_enqueued = attempts_enqueued
        self._keep_alive = keep_alive      # inputs read by stream-ordered kernels
        self.info: dict = {}

    def resolve(self, c=None) -> bool:
        """True: the integration finished and the outputs are valid.  False: it needs more attempts than were enqueued
        (the hint is raised; re-issue the sample).  Raises on a failed integration."""
        if c is None:
            c = self.counters.cpu().tolist() if self.n_rows else [1, 0, 0, 0, 0, 0, 0, 0]
        self.info = {"status": c[0], "nfev": c[1], "accepted": c[2], "rejected": c[3], "nan": bool(c[4]),
                     "attempts": c[5], "net_calls": c[1] + 1}
        self.agent.last_info = self.info
        self.denoiser.calls = c[1] + 1
        if c[0] < 0:
            raise capi.VphoError("RK45: required step size is less than spacing between numbers")
        if c[0] == 0:
            if getattr(self, "pair", None) is None:     # stand-alone deferred sample(): the caller re-issues with a larger budget
                self.denoiser.attempts_hint = max(self.attempts_enqueued, c[5]) + 4
            return False
        if c[4]:
            print("\033[31mWarning: NaN detected in score evaluation. \033[0m")   # score_based_model.py:70
        # steady state: enqueue what the last batch needed plus one spare attempt (14 no-op launches, ~35 us) -- unless the
        # last SPARE_WINDOW batches all needed the same number: then the spare is dropped until a batch breaks the pattern
        # (that batch is continued on its own workspace by the caller, `PendingPair.advance`, and the spare comes back)
        hist = getattr(self.denoiser, "_attempt_hist", None)
        if hist is None:
            hist = self.denoiser._attempt_hist = collections.deque(maxlen=SPARE_WINDOW)
        hist.append(int(c[5]))
        stable = len(hist) == SPARE_WINDOW and min(hist) == max(hist)
        self.denoiser.attempts_hint = max(1, c[5]) + (1 if (self.agent.spare_attempt and not stable) else 0)
        return True


class PendingPair:
    """Two samplers advancing in lock-step on the same workspaces.  `advance(n)` enqueues `n` more RK attempts for both
    (`vpho_sample_pair_continue`; a finished integration ignores them) followed by the final predictor step
    (`vpho_sample_pair_finish`, which runs only for an integration that has reached t = eps)."""

    def __init__(self, lib, args_a, args_b, pendings):
        self.lib, self.args_a, self.args_b, self.pendings = lib, args_a, args_b, pendings

    def advance(self, attempts: int, stream) -> None:
        pa, pb = C.byref(self.args_a), C.byref(self.args_b)
        self.lib.check(self.lib.c.vpho_sample_pair_continue(pa, pb, int(attempts), stream), "vpho_sample_pair_continue")
        self.lib.check(self.lib.c.vpho_sample_pair_finish(pa, pb, stream), "vpho_sample_pair_finish")
        for p in self.pendings:
            p.attempts_enqueued += int(attempts)


class ScoreBasedModelAgent:
    """Drop-in for the evaluation-time methods of the reference's `ScoreBasedModelAgent`."""

    def __init__(self, sampling_steps: int = 50, sample_num: int = 100, sampler: str = "ode",
                 first_attempts: int = 6):
        if sampler != "ode":
            raise NotImplementedError("Only ode sampler is supported for now.")
        self.sampling_steps = sampling_steps
        self.sample_num = sample_num
        self.sampling_eps = SAMPLING_EPS
        self.T = 1.0
        self.first_attempts = first_attempts
        self.spare_attempt = False     # VphoHotPath enables it together with deferred status checks
        self.last_info: dict = {}

    def prior_fn(self, shape, T):
        """ve_prior (lib/model/sde.py:26-28): CPU draw from torch's global generator, as the reference does."""
        return torch.randn(*shape) * ve_prior_std(T)

    @torch.no_grad()
    def sample(self, data: dict, denoiser: Denoiser, T0: float, init_x: Optional[torch.Tensor] = None,
               return_inprocess: bool = True, prior: Optional[torch.Tensor] = None, defer_check: bool = False):
        """-> (in_process (N, steps, D) float64 [permuted view, as the reference returns it], final (N, D) float64).

        The integration runs entirely on the device.  By default the 8-int status word is read back before returning
        (one host sync) and more RK attempts are enqueued if the controller has not reached t = eps yet.  With
        `defer_check=True` nothing is read back: a third value, a `PendingSample`, is returned and the caller must call
        `resolve()` on it once the stream has been synchronised for another reason (VphoHotPath.predict does this once
        per batch); `resolve()` says whether the outputs are valid or the sample has to be re-issued."""
        device = (data["feat_unique"] if "feat_unique" in data else data["feat"]).device
        D = denoiser.out_dim
        n_rows = int(data["n_rows"]) if "n_rows" in data else int(data["feat"].shape[0])
        # `prior` (optional): the prior draw randn*sigma(T0) supplied by the caller ("sampler noise fed from the same
        # seeded tensor"); otherwise drawn from torch's global CPU generator exactly as the reference does.
        prior = self.prior_fn((n_rows, D), T0).to(device) if prior is None else prior.to(device)
        x0 = prior if init_x is None else init_x.to(device) + prior      # score_based_model.py:62
        x0 = x0.contiguous().float()
        d2 = dict(data)
        d2.setdefault("sample_num", self.sample_num)
        feat, rpf = _unique_feat(d2, n_rows)
        n_eval = self.sampling_steps
        lib = denoiser.lib
        ws = denoiser.workspace(n_rows, rpf, n_eval, device)
        xs = torch.empty((n_eval, n_rows, D), dtype=torch.float64, device=device) if return_inprocess else None
        x = torch.empty((n_rows, D), dtype=torch.float64, device=device)
        counters = torch.zeros(8, dtype=torch.int32, device=device)
        stream = capi.stream_of(x0)
        hint = getattr(denoiser, "attempts_hint", None) or self.first_attempts
        st = lib.c.vpho_sample_begin(denoiser.handle, capi.ptr(feat), n_rows, rpf, capi.ptr(x0), float(T0),
                                     float(self.sampling_eps), None, n_eval, RTOL, ATOL, MAX_STEP, n_eval, hint,
                                     capi.ptr(xs), capi.ptr(x), capi.ptr(counters), capi.ptr(ws), ws.numel(), stream)
        lib.check(st, "vpho_sample_begin")
        lib.check(lib.c.vpho_sample_finish(denoiser.handle, n_rows, rpf, n_eval, capi.ptr(ws), ws.numel(), stream),
                  "vpho_sample_finish")
        pending = PendingSample(self, denoiser, counters, n_rows, hint, (feat, x0))
        xs_view = None if xs is None else xs.permute(1, 0, 2)
        if defer_check:
            return xs_view, x, pending
        while True:
            c = counters.cpu().tolist() if n_rows else [1, 0, 0, 0, 0, 0, 0, 0]
            if c[0] != 0:
                break
            lib.check(lib.c.vpho_sample_continue(denoiser.handle, n_rows, rpf, n_eval, 4, capi.ptr(ws), ws.numel(),
                                                 stream), "vpho_sample_continue")
            lib.check(lib.c.vpho_sample_finish(denoiser.handle, n_rows, rpf, n_eval, capi.ptr(ws), ws.numel(), stream),
                      "vpho_sample_finish")
        pending.resolve(c)
        return xs_view, x

    @torch.no_grad()
    def sample_pair(self, data_a: dict, denoiser_a: Denoiser, data_b: dict, denoiser_b: Denoiser, T0: float,
                    return_inprocess: bool = True, prior_a: Optional[torch.Tensor] = None,
                    prior_b: Optional[torch.Tensor] = None, inprocess_float32=(False, False), ws_slot: int = 0):
        """The two `sample()` calls of `vpho_net.forward(mode='predict')` (hand, then object: VPHO.py:239-262) advanced in
        lock-step through shared kernel launches (`vpho_sample_pair_*`).  Priors are drawn in the reference's order
        (a, then b).  Always deferred: -> ((xs_a, x_a, pending_a), (xs_b, x_b, pending_b)); results are bit-identical
        to two separate `sample(..., defer_check=True)` calls.  `inprocess_float32[i]`: return sampler i's trajectory
        already rounded to float32 (what `vpho_net` keeps of the hand's: `.float()`, VPHO.py:243) -- no float64 copy is
        written."""
        jobs = []
        for (data, den, prior), f32 in zip(((data_a, denoiser_a, prior_a), (data_b, denoiser_b, prior_b)), inprocess_float32):
            device = (data["feat_unique"] if "feat_unique" in data else data["feat"]).device
            D = den.out_dim
            n_rows = int(data["n_rows"]) if "n_rows" in data else int(data["feat"].shape[0])
            prior = self.prior_fn((n_rows, D), T0).to(device) if prior is None else prior.to(device)
            x0 = prior.contiguous().float()
            d2 = dict(data)
            d2.setdefault("sample_num", self.sample_num)
            feat, rpf = _unique_feat(d2, n_rows)
            n_eval = self.sampling_steps
            ws = den.workspace(n_rows, rpf, n_eval, device, ws_slot)
            xs = (torch.empty((n_eval, n_rows, D), dtype=torch.float32 if f32 else torch.float64, device=device)
                  if return_inprocess else None)
            # zeros, not empty: the final predictor step writes `x` only once the integration has reached t = eps, and
            # the speculative downstream work of `VphoHotPath.predict` must never consume uninitialised memory
            x = torch.zeros((n_rows, D), dtype=torch.float64, device=device)
            counters = torch.zeros(8, dtype=torch.int32, device=device)
            args = capi.SampleArgs(den.handle, capi.ptr(feat), n_rows, rpf, capi.ptr(x0), float(T0), float(self.sampling_eps),
                                   None, n_eval, RTOL, ATOL, MAX_STEP, n_eval, capi.ptr(None if f32 else xs), capi.ptr(x),
                                   capi.ptr(counters), capi.ptr(ws), ws.numel(), capi.ptr(xs if f32 else None))
            jobs.append((den, args, xs, x, counters, n_rows, (feat, x0, ws)))
        lib = denoiser_a.lib
        stream = capi.stream_of(jobs[0][3])
        hint = max((getattr(j[0], "attempts_hint", None) or self.first_attempts) for j in jobs)
        pa, pb = C.byref(jobs[0][1]), C.byref(jobs[1][1])
        lib.check(lib.c.vpho_sample_pair_begin(pa, pb, hint, stream), "vpho_sample_pair_begin")
        lib.check(lib.c.vpho_sample_pair_finish(pa, pb, stream), "vpho_sample_pair_finish")
        out = []
        for den, args, xs, x, counters, n_rows, keep in jobs:
            pending = PendingSample(self, den, counters, n_rows, hint, keep)
            out.append((None if xs is None else xs.permute(1, 0, 2), x, pending))
        pair = PendingPair(lib, jobs[0][1], jobs[1][1], (out[0][2], out[1][2]))
        out[0][2].pair = out[1][2].pair = pair
        return out[0], out[1]

    def get_score(self, data, denoiser):
        return denoiser(data)
