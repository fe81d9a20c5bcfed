"""Evaluation hot path of `vpho_net.forward(data, mode='predict')` (lib/model/VPHO.py:228-304) on the CUDA kernels.

`VphoHotPath.predict` takes what the reference's feature extractor produces per image (VPHO.py:112-173: encodings,
heat-maps, regressed MANO, local forces) plus the dataset fields the hot path reads (dexycb6.py:471-509) and returns
the same `pd_dt` keys with the same shapes/dtypes.  All arithmetic is in `libvpho_b200.so`; there is no CPU path.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import capi
from .aggregation import Assets, HeadObject, HeadPhysics, HOI_Aggregator
from .head_mano import HeadMano
from .score_based_model import Denoiser, ScoreBasedModelAgent

_TENSOR_KEYS = ("encoding_hand", "encoding_obj", "pd_mano_pose", "pd_mano_shape", "hm_hand", "hm_obj", "force_local",
                "cam_intr_crop_flip", "root_joint_flip", "root_joint", "is_right", "is_grasped", "bbox_hand",
                "bbox_obj_rect")


class VphoHotPath:
    def __init__(self, mano_model: Dict, anchors: Dict, objects: Dict, denoiser_hand_state: Dict,
                 denoiser_obj_state: Dict, *, sample_num: int = 100, sampling_steps: int = 50, sample_T0: float = 0.65,
                 topk_hand: int = 30, topk_obj: int = 10, lib: Optional[capi.Library] = None, debug: bool = False):
        self.lib = lib or capi.lib()
        self.sample_num, self.sampling_steps, self.sample_T0 = sample_num, sampling_steps, sample_T0
        self.topk_hand, self.topk_obj = topk_hand, topk_obj
        self.head_mano = HeadMano(mano_model, lib=self.lib)
        self.assets = Assets(anchors, objects, lib=self.lib)
        self.head_obj = HeadObject(self.assets)
        self.head_physics = HeadPhysics(self.assets)
        self.denoiser_hand = Denoiser(denoiser_hand_state, lib=self.lib)
        self.denoiser_obj = Denoiser(denoiser_obj_state, lib=self.lib)
        self.score_agent = ScoreBasedModelAgent(sampling_steps=sampling_steps, sample_num=sample_num)
        self.hoi_aggregator = HOI_Aggregator(self.head_mano, self.assets, debug=debug)
        # Two batches may be in flight: each slot has its own aggregator (workspace, asset handle with its fork / join
        # events) and its own aggregation stream, so that the latency-bound aggregation chains of consecutive batches
        # overlap each other instead of queueing on one stream (measured: 2.77 -> 2.72 ms per batch).  `hoi_aggregator` is the one the latest batch used.
        self.agg_slots = 2                # 1: every batch's aggregation on the same stream / workspace (tools/pipeline_probe.py)
        self.side_slots = 1               # the same for the output-only stream (measured: a second one gains nothing)
        self._hoi_slots = [self.hoi_aggregator, None]
        self._asset_src = (anchors, objects, debug)
        self.last_info: dict = {}
        self._slot_done = [[], []]        # events completing the library-stream work of each slot's latest batch
        self._side_stream2 = [None, None]
        self._agg_stream = [None, None]
        self._status_host = None
        self._ws_slot = 0
        self._events: list = []          # ring of reusable CUDA events (creating one per batch costs a driver call each)
        self._event_i = 0
        # True: everything on the caller's stream (no output-only side stream, no high-priority aggregation stream);
        # bench.py's serialised per-kernel pass uses it together with vpho_set_pdl(0).  Results are identical.
        self.serialize = False

    # ---- vpho_net.postprocess_diffusion_hand, branch 'mano_pose' (VPHO.py:306-331) ----
    def postprocess_diffusion_hand(self, hand_inprocess, hand_final, pd_mano_shape):
        """One fused pass per tensor (`vpho_postprocess_hand`): float64 6D rotations -> float32 axis-angle + shape.
        hand_inprocess: the sampler's (N, steps, 96) permuted view of its [steps][N][96] buffer, or None."""
        S = self.sample_num
        bs = pd_mano_shape.shape[0]
        dev = hand_final.device
        shape = pd_mano_shape.contiguous().float()

        def fused(x, n_steps, n_rows):
            out = torch.empty((n_rows, n_steps, 58), dtype=torch.float32, device=dev)
            fn = self.lib.c.vpho_postprocess_hand_f32 if x.dtype == torch.float32 else self.lib.c.vpho_postprocess_hand
            self.lib.check(fn(capi.ptr(x), n_steps, n_rows, S, capi.ptr(shape), capi.ptr(out), capi.stream_of(x)),
                           "vpho_postprocess_hand")
            return out

        hf = fused(hand_final.contiguous().double(), 1, bs * S).reshape(-1, 58)
        hi = None
        if hand_inprocess is not None:
            n_in = hand_inprocess.shape[1]
            native = hand_inprocess.permute(1, 0, 2)              # storage order [steps][N][96]
            if not native.is_contiguous() or native.dtype not in (torch.float64, torch.float32):
                native = native.contiguous().double()
            hi = fused(native, n_in, bs * S)
        return hi, hf

    # ---- N1: vpho_net.forward from the RoI-aligned feature maps (VPHO.py:129-304) ----
    def attach_feature_heads(self, state: Dict) -> "VphoHotPath":
        """`state`: the reference's state dict (or the `rest` part `vpho_b200.checkpoint.load_denoiser_states` returns) holding
        head_hm_*, encoder_*, head_mano, cross_*, head_physics.  After this `predict_from_features` is available."""
        from .producers import FeatureHeads
        self.feature_heads = FeatureHeads(state, lib=self.lib)
        return self

    @torch.no_grad()
    def predict_from_features(self, hf_hr: torch.Tensor, of_or_rect: torch.Tensor, hf_hr_rect: torch.Tensor, batch: Dict, **kw) -> Dict:
        """The predict branch of `vpho_net.forward` from the four RoI-aligned maps on (VPHO.py:129-304; `of_or` is not
        consumed there): the producers (`FeatureHeads`, one C call) write encodings, heat-maps, regression pose and local
        forces on the device and `predict` consumes them in place -- no heat-map ever crosses PCIe.  `batch` carries the
        dataset fields (bbox_hand[_rect], bbox_obj[_rect], is_right, gravity, cam_intr_crop_flip, root_joint[_flip], is_grasped,
        obj_id / obj_name).  Also returns `reg_hand_vert` / `reg_hand_joint` / `hand_heatmap` / `obj_heatmap` / `force_local`
        like the reference's pd_dt (VPHO.py:230-234)."""
        if getattr(self, "feature_heads", None) is None:
            raise capi.VphoError("predict_from_features: call attach_feature_heads(state) first")
        f = self.feature_heads(hf_hr, of_or_rect, hf_hr_rect, batch)
        b = dict(batch)
        b.update(encoding_hand=f["encoding_hand"], encoding_obj=f["encoding_obj"], pd_mano_pose=f["mano_pose"],
                 pd_mano_shape=f["mano_shape"], hm_hand=f["hand_heatmap"], hm_obj=f["obj_heatmap"], force_local=f["force_local"])
        pd = self.predict(to_device(b, hf_hr.device), **kw)
        rv, rj = self.head_mano.get_hand_verts(pose=f["mano_pose"], shape=f["mano_shape"])
        pd.update(reg_hand_vert=rv, reg_hand_joint=rj, hand_heatmap=f["hand_heatmap"], obj_heatmap=f["obj_heatmap"],
                  force_local=f["force_local"])
        return pd

    @torch.no_grad()
    def predict(self, batch: Dict, *, prior_hand: Optional[torch.Tensor] = None, prior_obj: Optional[torch.Tensor] = None,
                with_inprocess: bool = True, prefetch=None, defer_join: bool = False) -> Dict:
        """The whole batch (both samplers, MANO, scoring, aggregation) is enqueued without a host synchronisation and the
        two samplers' status words are read once at the end.  The number of RK attempts enqueued up front adapts to the
        previous batch (what it needed plus one spare).  If an integration needed more, it is CONTINUED on the same
        workspaces (`vpho_sample_pair_continue` + `_finish`, as the reference's solve_ivp simply keeps stepping) until both
        controllers report completion -- there is no attempt cap -- and only then is the downstream work enqueued again on
        the now valid samples.  Results are identical either way.

        defer_join=True (pipelined evaluation loops): the caller's stream is NOT made to wait for the aggregation and the
        output-only work, which run on the library's own streams; `predict` returns as soon as both samplers have
        finished, so the next batch's samplers (tensor-core bound, one CTA per SM) start while this batch's aggregation
        (a chain of small latency-bound kernels) fills the idle SM time between them.  `pd["_done"]` holds the events
        that complete the outputs; call `VphoHotPath.join(pd)` before reading them on another stream.

        `predict_begin` / `predict_end` split the call at its one host synchronisation, so that a loop can enqueue batch
        i+1 before it waits for batch i's status (no idle gap on the GPU between two batches' samplers)."""
        return self.predict_end(self.predict_begin(batch, prior_hand=prior_hand, prior_obj=prior_obj, with_inprocess=with_inprocess,
                                                   prefetch=prefetch, defer_join=defer_join))

    @torch.no_grad()
    def predict_begin(self, batch: Dict, *, prior_hand: Optional[torch.Tensor] = None, prior_obj: Optional[torch.Tensor] = None,
                      with_inprocess: bool = True, prefetch=None, defer_join: bool = True) -> Dict:
        """Enqueue one batch (samplers, speculative downstream work, the caller's `prefetch` hook); no host wait.  At most
        TWO batches may be in flight (begin, begin, end, begin, end, ...): the sampler workspaces alternate between two
        slots.  -> ticket for `predict_end`."""
        if prior_hand is None or prior_obj is None:
            # draw both priors up front, in the reference's order (hand, then object)
            from .score_based_model import ve_prior_std
            S, bs = self.sample_num, batch["encoding_hand"].shape[0]
            if prior_hand is None:
                prior_hand = torch.randn(bs * S, self.denoiser_hand.out_dim) * ve_prior_std(self.sample_T0)
            if prior_obj is None:
                prior_obj = torch.randn(bs * S, self.denoiser_obj.out_dim) * ve_prior_std(self.sample_T0)
        self.score_agent.spare_attempt = True
        S = self.sample_num
        enc_h, enc_o = batch["encoding_hand"], batch["encoding_obj"]
        bs = enc_h.shape[0]
        self._ws_slot ^= 1
        # Back-pressure: this slot's previous batch (two batches ago) must have finished its aggregation and output-only work
        # before this batch's samplers start.  Without it the low-priority library streams fall behind the samplers without
        # bound in a loop that never joins them (their backlog is then paid at the end, and every block they still hold is
        # one the caching allocator cannot reuse: it answers with a fresh cudaMalloc -- milliseconds of host stall -- every
        # other batch and the pool grows by gigabytes).
        if enc_h.is_cuda:
            main = torch.cuda.current_stream(enc_h.device)
            for ev in self._slot_done[self._ws_slot]:
                main.wait_event(ev)
            self._slot_done[self._ws_slot] = []
        # The hand and the object integrations issue the same sequence of network calls: they advance in lock-step
        # through shared kernel launches (`sample_pair`), the object's work items filling the SMs the hand's leave idle.
        samples = self.score_agent.sample_pair(
            {"feat_unique": enc_h, "n_rows": bs * S}, self.denoiser_hand, {"feat_unique": enc_o, "n_rows": bs * S},
            self.denoiser_obj, self.sample_T0, return_inprocess=with_inprocess, prior_a=prior_hand, prior_b=prior_obj,
            inprocess_float32=(True, False),      # the hand trajectory is only ever used as float32 (VPHO.py:243)
            ws_slot=self._ws_slot)
        t = {"batch": batch, "samples": samples, "pend": (samples[0][2], samples[1][2]), "issue": 0,
             "with_inprocess": with_inprocess, "prefetch": prefetch, "defer_join": defer_join, "slot": self._ws_slot}
        self._issue(t)
        return t

    def _issue(self, t: Dict) -> None:
        t["status"] = self._snapshot(t["pend"], t["slot"])     # stream-ordered behind the samplers only: the host wakes when THEY finish
        t["pd"] = self._downstream(t["batch"], t["samples"], t["with_inprocess"], t["defer_join"], t["slot"])   # speculative on the first pass
        if t["prefetch"] is not None:
            # caller hook `prefetch(pd, issue)`, run after the batch is enqueued and before the host blocks on its
            # status: the place to enqueue device-to-host reads of `pd` (stream-ordered behind the aggregation) and, for
            # issue == 0, the next batch's host-to-device copies on another stream.
            # When the samplers had to be continued (issue > 0, rare) it is called again with the new outputs.
            t["prefetch"](t["pd"], t["issue"])
        t["issue"] += 1

    @torch.no_grad()
    def predict_end(self, t: Dict) -> Dict:
        """The one host synchronisation of the batch (its samplers' status words); -> `pd`."""
        pend = t["pend"]
        while True:
            done = self._await(pend, t["status"])
            if done:
                return t["pd"]
            # slow path: keep stepping both integrations until they finish, then redo the downstream work once
            stream = capi.stream_of(t["samples"][0][1])
            while not done:
                pend[0].pair.advance(4, stream)
                done = self._await(pend, self._snapshot(pend, t["slot"]))
            self._issue(t)

    def _event(self):
        """Next event of a ring of 32.  Re-recording an event a stale `pd` still refers to only makes a late join wait for
        later work of the same stream, which is still correct."""
        if len(self._events) < 32:
            self._events.append(torch.cuda.Event())
            return self._events[-1]
        self._event_i = (self._event_i + 1) % 32
        return self._events[self._event_i]

    def _snapshot(self, pend, slot: int = 0):
        """Enqueue the read of both controllers' counters (pinned host buffer of this batch's slot + event on the current
        stream)."""
        dev_c = torch.stack([p.counters for p in pend])
        if not dev_c.is_cuda:
            return dev_c, None
        if self._status_host is None:
            self._status_host = [torch.empty((2, 8), dtype=torch.int32).pin_memory() for _ in range(2)]
        host = self._status_host[slot & 1]
        host.copy_(dev_c, non_blocking=True)
        ev = self._event()
        ev.record(torch.cuda.current_stream(dev_c.device))
        return host, ev

    def _await(self, pend, status) -> bool:
        host, ev = status
        if ev is not None:
            ev.synchronize()
        ok = [p.resolve(c) for p, c in zip(pend, host.tolist())]
        self.last_info = {"hand": pend[0].info, "obj": pend[1].info}
        return all(ok)

    @staticmethod
    def _priorities(main):
        """(aggregation stream, output-only stream) priorities.  Caller on a default-priority stream: the aggregation chain
        goes ahead of the pending candidate-mesh CTAs (-1, 0).  Caller on a high-priority stream (what a pipelined loop
        should use: the samplers are the long pole, the previous batch's aggregation has a whole sampler phase of slack):
        one and two levels below the caller.  Measured on a B200, 64 x 100 x 50 (tools/pipeline_probe.py): pipelined
        2.85 ms/batch with the caller at -2, 3.04 at 0; joined 3.08-3.23."""
        p = int(getattr(main, "priority", 0) or 0)
        if p >= 0:
            return -1, 0
        return min(0, p + 1), min(0, p + 2)

    @staticmethod
    def join(pd: Dict, stream=None) -> Dict:
        """Make `stream` (default: the current one) wait for the outputs of a `predict(..., defer_join=True)` result."""
        evs = pd.pop("_done", None)
        if evs:
            stream = stream or torch.cuda.current_stream(evs[0][1])
            for ev, _ in evs:
                stream.wait_event(ev)
        return pd

    def _aggregator(self, slot: int):
        if self._hoi_slots[slot] is None:
            anchors, objects, debug = self._asset_src
            self._hoi_slots[slot] = HOI_Aggregator(self.head_mano, Assets(anchors, objects, lib=self.lib), debug=debug)
        self.hoi_aggregator = self._hoi_slots[slot]
        return self.hoi_aggregator

    def _downstream(self, batch: Dict, samples, with_inprocess: bool, defer_join: bool = False, slot: int = 0) -> Dict:
        """Everything after the two `sample()` calls of the predict branch (VPHO.py:243-304).  batch: tensors on the CUDA
        device (see `to_device`)."""
        S = self.sample_num
        (xs_h, x_h, _), (xs_o, x_o, _) = samples
        enc_h = batch["encoding_hand"]
        bs = enc_h.shape[0]
        pd_mano_pose, pd_mano_shape = batch["pd_mano_pose"], batch["pd_mano_shape"]
        pd = {"hand_heatmap": batch["hm_hand"], "obj_heatmap": batch["hm_obj"], "force_local": batch["force_local"]}
        main = torch.cuda.current_stream(enc_h.device) if enc_h.is_cuda else None
        # The aggregator needs only the final hand poses.  Everything else computed from the hand sampler's output is
        # output-only (the in-process trajectory for visualisation, the 6400 posed meshes of the candidates): it runs on a
        # side stream, concurrently with the aggregation.
        _, final_mano = self.postprocess_diffusion_hand(None, x_h, pd_mano_shape)
        pd["diff_final_hand_mano"] = final_mano.reshape(bs, S, 58)

        def output_only_work():
            if with_inprocess:
                inproc, _ = self.postprocess_diffusion_hand(xs_h, x_h, pd_mano_shape)
                pd["diff_inprocess_hand_mano"] = inproc.reshape(bs, S, -1, 58)
                # VPHO.py:250: every 10th output point of the first candidate of the first image (visualisation)
                iv, ij = self.head_mano.get_hand_verts(pose=inproc[0, ::10, :48], shape=inproc[0, ::10, 48:])
                pd["diff_inprocess_hand_vert"], pd["diff_inprocess_hand_joint"] = iv.reshape(-1, 778, 3), ij.reshape(-1, 21, 3)
            fv, fj = self.head_mano.get_hand_verts(pose=final_mano[:, :48], shape=final_mano[:, 48:])
            pd["diff_final_hand_vert"] = fv.reshape(bs, S, 778, 3)
            pd["diff_final_hand_joint"] = fj.reshape(bs, S, 21, 3)

        side2 = None
        if main is not None and not self.serialize:
            sslot = slot if self.side_slots > 1 else 0
            if self._side_stream2[sslot] is None:
                self._side_stream2[sslot] = torch.cuda.Stream(device=enc_h.device, priority=self._priorities(main)[1])
            side2 = self._side_stream2[sslot]
            side2.wait_stream(main)
            for t in (xs_h, x_h, final_mano):
                if t is not None:
                    t.record_stream(side2)
            with torch.cuda.stream(side2):
                output_only_work()
            for k in ("diff_inprocess_hand_mano", "diff_inprocess_hand_vert", "diff_inprocess_hand_joint",
                      "diff_final_hand_vert", "diff_final_hand_joint"):
                if k in pd:
                    pd[k].record_stream(main)
        else:
            output_only_work()

        if with_inprocess:
            pd["diff_inprocess_obj_6d"] = xs_o.reshape(bs, S, -1, 9)
        pd["diff_final_obj_6d"] = x_o.reshape(bs, S, 9)

        aslot = slot if self.agg_slots > 1 else 0
        hoi = self._aggregator(aslot)

        def aggregate():
            return hoi(
                cam_intrinsic=batch["cam_intr_crop_flip"], root_joint_flip=batch["root_joint_flip"],
                root_joint=batch["root_joint"], is_right=batch["is_right"], force_local=batch["force_local"],
                is_grasped=batch["is_grasped"], hand_pose_diff=final_mano[:, :48], hand_pose_regression=pd_mano_pose,
                hand_shape=final_mano[:, 48:], hand_heatmap=batch["hm_hand"], hand_bbox=batch["bbox_hand"],
                hand_topk=self.topk_hand, obj_pose6d=pd["diff_final_obj_6d"], obj_heatmap=batch["hm_obj"],
                obj_bbox=batch["bbox_obj_rect"], obj_topk=self.topk_obj,
                obj_name=batch["obj_id"] if "obj_id" in batch else batch["obj_name"])

        # The aggregation is a chain of short kernels on the critical path; the output-only stream floods the GPU with
        # the 6400 candidate meshes at the same time.  Running the chain on a high-priority stream lets its CTAs be
        # placed ahead of the pending mesh CTAs whenever SM slots free up.
        if side2 is not None:
            if self._agg_stream[aslot] is None:
                self._agg_stream[aslot] = torch.cuda.Stream(device=enc_h.device, priority=self._priorities(main)[0])
            hs = self._agg_stream[aslot]
            hs.wait_stream(main)
            if defer_join:
                # nothing on the caller's stream orders these reads any more: tell the allocator who uses the tensors
                for t in [final_mano, x_o] + [v for v in batch.values() if torch.is_tensor(v) and v.is_cuda]:
                    t.record_stream(hs)
            with torch.cuda.stream(hs):
                sel = aggregate()
            if not defer_join:
                main.wait_stream(hs)
            for v in sel.values():
                if torch.is_tensor(v):
                    v.record_stream(main)
        else:
            sel = aggregate()
        pd["agg_obj_6d"] = sel["obj_agg_6d"]
        pd["agg_hand_mano"] = sel["hand_agg_mano"]
        pd["agg_hand_vert"] = sel["hand_agg_vert"]
        pd["agg_hand_joint"] = sel["hand_agg_joint"]
        pd["_sel"] = sel
        if side2 is not None:
            if defer_join:
                pd["_done"] = []
                for st in (hs, side2):
                    ev = self._event()
                    ev.record(st)
                    pd["_done"].append((ev, enc_h.device))
                self._slot_done[slot] = [ev for ev, _ in pd["_done"]]
            else:
                main.wait_stream(side2)
        return pd

    __call__ = predict


def to_device(batch: Dict, device="cuda", non_blocking: bool = True) -> Dict:
    """Host dict (numpy arrays / tensors) -> tensors on `device`; `obj_name` becomes an int32 id tensor when `obj_id`
    is present.  Mirrors the `to_device(batch)` step of Trainer.evaluate (train_diff_hand_obj.py:214)."""
    import numpy as np
    out = {}
    for k, v in batch.items():
        if isinstance(v, np.ndarray):
            v = torch.from_numpy(v)
        if isinstance(v, torch.Tensor):
            if k == "obj_id":
                v = v.to(torch.int32)          # the kernels read `const int32_t*`
            out[k] = v.to(device, non_blocking=non_blocking)
        else:
            out[k] = v
    return out
