"""Data-parallel gradient exchange for the drop-in modules: one process per GPU, dialogues sharded
across ranks, ONE summed gradient allreduce per step.

The reference has no distributed code at all (its ``import torch.distributed`` at train.py:12 is
dead); this is the hookup the north star asks for.  Design:
  * gradients are exchanged in a few flat fp32 buckets, ordered by the order in which autograd finishes them
    (``observe_grad_order``: head -> recurrence weights -> input projections -> encoders);
  * a post-accumulate-grad hook per parameter counts a bucket down; when it reaches zero the bucket's gradients are
    packed into the flat buffer by ONE multi-tensor copy (``p.grad`` become views of it); its ``all_reduce(SUM)`` is
    enqueued in ``finish()`` (default) or, with ``overlap=True``, right away on the NCCL stream next to the rest of the
    backward — on B200 that costs more than it hides (see ``__init__``); ``finish()`` waits for all buckets;
  * parameters the model never uses (e.g. ``encoder_*.pos_ffn.fc``) are left out, so
    their ``.grad`` stays ``None`` exactly as in the single-process reference step (Adam with weight
    decay would otherwise move them);
  * loss scaling is the caller's: scale the shard loss by n_shard/n_global and SUM-reduce to equal
    the single-process step on the concatenated batch (loss.py:21 divides by the batch's utterances).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


# Registered-but-never-applied parameters of the drop-in modules: the reference constructs them, so they must
# exist for state_dict / RNG-order compatibility, but no forward path touches them and their grad stays None.
#   every model:   encoder_*.pos_ffn.fc.*                                   (encoder.py:99 builds it, :111 never calls it)
#   MARN1_sps:     marn_cell_{f,b}.lstm_s.*, marn_cell_{f,b}.crossatt_a2l.*, marn_cell_{f,b}.crossatt_l2a.Wv
#                  (lsthm_sps.py:136,153: constructed; :59-72 uses only Wq, Wk of crossatt_l2a)
_UNUSED_PATTERNS = (r"\.pos_ffn\.fc\.", r"^marn_cell_[fb]\.lstm_s\.", r"^marn_cell_[fb]\.crossatt_a2l\.",
                    r"^marn_cell_[fb]\.crossatt_l2a\.Wv$")


# Per model class (looked up along the MRO, so MARN1_no_en inherits MARN1_nsps's):
#   MARN1_onlysp:  linear.* (lsthm_onlysp.py: constructed, never called), marn_cell_{f,b}.lstm_q0/q1.* (leftovers of lsthm_sps)
#   MARN1_nsps:    fc2.* (lsthm_nsps.py:352: resid_a is computed and discarded), marn_cell_{f,b}.gru_l.* (line 156: never called)
#   MARN1_no_en:   additionally encoder_l.* (lsthm_no_en.py:287 constructs it; :306/:309 — its only calls — are commented out)
_UNUSED_BY_CLASS = {"MARN1_onlysp": (r"^linear\.", r"^marn_cell_[fb]\.lstm_q[01]\."),
                    "MARN1_nsps": (r"^fc2\.", r"^marn_cell_[fb]\.gru_l\."),
                    "MARN1_no_en": (r"^encoder_l\.",)}


def unused_parameter_names(model: torch.nn.Module) -> List[str]:
    """Names of the parameters no forward path uses (their gradient must stay None, as in the single-process reference
    step: Adam with weight decay would otherwise move them).  tests/test_host_logic.py pins this list to the
    ``grad is None`` set of the reference-generated fixtures."""
    import re
    extra = tuple(p for c in type(model).__mro__ for p in _UNUSED_BY_CLASS.get(c.__name__, ()))
    pats = [re.compile(p) for p in _UNUSED_PATTERNS + extra]
    return [n for n, _ in model.named_parameters() if any(p.search(n) for p in pats)]


def observe_grad_order(model: torch.nn.Module, step_fn) -> List[str]:
    """Run ``step_fn()`` (one forward + backward of ``model``) and return the parameter names in the order autograd
    finished their gradients.  Pass the result as ``GradAllReducer(..., order=...)``: buckets laid out in THAT order fill
    front to back during the backward, so every allreduce but the last overlaps the remaining backward kernels (registration
    order is a poor guess: the drop-in MARN registers its encoders last, but their gradients are also the last to finish).
    The gradients this dry run leaves behind are dropped."""
    names = {id(p): n for n, p in model.named_parameters()}
    seen: List[str] = []
    handles = [p.register_post_accumulate_grad_hook(lambda q: seen.append(names[id(q)]))
               for p in model.parameters() if p.requires_grad]
    try:
        step_fn()
    finally:
        for h in handles:
            h.remove()
    model.zero_grad(set_to_none=True)
    return seen


class GradAllReducer:
    def __init__(self, model: torch.nn.Module, world_size: int, bucket_bytes: int = 4 << 20,
                 group: Optional[dist.ProcessGroup] = None, flatten_params: bool = False,
                 order: Optional[List[str]] = None, overlap: bool = False):
        self.world, self.group = world_size, group
        # overlap=False (default): every bucket is reduced in finish(), after the backward.  True: a bucket's allreduce is
        # enqueued the moment it is complete.  Measured on 8 x B200 (profiles/r02/scale_ab_n8.log): the 6.8 MB of gradients are
        # ~0.15 ms of NVLink time at the end of a 14.7 ms step, while NCCL kernels running NEXT TO the backward take SMs from
        # its persistent / cooperative kernels (one CTA per SM): 15.22 ms per step overlapped vs 14.81 ms not (1 GPU: 14.65).
        self.overlap = overlap
        self.flatten_params = flatten_params
        self.param_buckets: List[torch.Tensor] = []   # flat parameter storage per bucket (FusedAdam steps on these)
        skip = set(unused_parameter_names(model))
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad and n not in skip]
        # without an observed order: autograd finishes gradients roughly in reverse registration order of use, walk the
        # parameters backwards so bucket 0 fills first.  With ``order`` (observe_grad_order): exactly that order; parameters it
        # does not name (no gradient in the observed step) go last.  Every rank must pass the same order.
        named = named[::-1]
        if order is not None:
            pos = {n: i for i, n in enumerate(order)}
            named.sort(key=lambda np_: pos.get(np_[0], len(pos)))     # stable: unnamed ones keep their relative order
        self.buckets: List[torch.Tensor] = []
        self._members: List[List[torch.nn.Parameter]] = []
        self._bucket_of: Dict[int, int] = {}
        self._home: Dict[int, tuple] = {}
        self._view: Dict[int, torch.Tensor] = {}      # the parameter-shaped view of its slot in the flat bucket
        cur, cur_bytes = [], 0
        for n, p in named:
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self._add_bucket(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._add_bucket(cur)
        self._pending = [0] * len(self.buckets)
        self._handles: List = []
        self.fire_order: List[int] = []
        self._fired = [False] * len(self.buckets)
        self.skipped = sorted(skip)
        for b, members in enumerate(self._members):
            for p in members:
                p.register_post_accumulate_grad_hook(self._make_hook(b))
        self.zero_grad()

    @staticmethod
    def _offsets(params):
        """Start offset of each parameter inside its flat bucket, padded to 8 floats so every view keeps the
        16/32-byte alignment the kernels ask for."""
        offs, o = [], 0
        for p in params:
            offs.append(o)
            o += (p.numel() + 7) // 8 * 8
        return offs, o

    def _add_bucket(self, params):
        offs, total = self._offsets(params)
        flat = torch.zeros(total, device=params[0].device, dtype=params[0].dtype)
        for p, o in zip(params, offs):
            self._bucket_of[id(p)] = len(self.buckets)
            self._home[id(p)] = (len(self.buckets), o)          # where this parameter's gradient must live
            self._view[id(p)] = flat[o:o + p.numel()].view_as(p)
        self.buckets.append(flat)
        self._members.append(list(params))
        if self.flatten_params:
            # parameters of a bucket share one flat storage too (p.data become views; values are preserved), so the
            # optimizer is one fused launch per bucket (the padding gaps hold zeros and stay zero under Adam)
            pflat = torch.zeros(total, device=params[0].device, dtype=params[0].dtype)
            for p, o in zip(params, offs):
                pflat[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = pflat[o:o + p.numel()].view_as(p)
            self.param_buckets.append(pflat)

    def _expected_ptr(self, p) -> int:
        return self._view[id(p)].data_ptr()

    def _gather(self, b):
        """Move the gradients autograd produced for bucket ``b`` into their slots of the flat bucket — ONE multi-tensor copy
        for the whole bucket — and make ``p.grad`` the views.  Members without a gradient in this step get a zero slot (they
        take part in the sum) and keep ``grad = None``."""
        src, dst = [], []
        if self.buckets[b].is_cuda:
            cs = torch.cuda.current_stream(self.buckets[b].device)
            for st, ev in self._xstream[b]:     # gradients finished on other streams (see the hook)
                if st != cs:
                    cs.wait_event(ev)
            self._xstream[b] = []
        with torch.no_grad():
            for p in self._members[b]:
                v = self._view[id(p)]
                if p.grad is None:
                    v.zero_()
                elif p.grad.data_ptr() != v.data_ptr():
                    src.append(p.grad)
                    dst.append(v)
            if src:
                torch._foreach_copy_(dst, src)
                # The sources were allocated on whatever stream produced them and are read HERE, possibly on another stream:
                # dropping them now would hand their memory back to the producing stream's pool, which may overwrite it before
                # this copy has run (seen as two corrupted recurrence-weight gradients when a bucket completed on a side stream
                # while the caller's stream was still busy).  Keep them until the next zero_grad(): by then backward() has
                # joined every stream into the caller's.  (No record_stream: see mab_net._EncoderBranches.)
                self._keepalive.extend(src)
        for p in self._members[b]:
            if p.grad is not None:
                p.grad = self._view[id(p)]
        if self.buckets[b].is_cuda and cs != self._home_stream:
            ev = torch.cuda.Event()             # packed on a side stream: finish() orders the caller's stream behind it
            ev.record(cs)
            self._packed_events.append(ev)

    def _reduce(self, b):
        self._fired[b] = True
        nvtx = torch.cuda.nvtx if self.buckets[b].is_cuda else None
        if nvtx is not None:
            nvtx.range_push(f"lsthm/K6_allreduce_bucket{b}")
        self._handles.append(dist.all_reduce(self.buckets[b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        if nvtx is not None:
            nvtx.range_pop()

    def _make_hook(self, b):
        def hook(param):
            self._pending[b] -= 1
            if self._pending[b] < 0:
                raise RuntimeError("GradAllReducer: a bucket received more gradients than it has members (zero_grad() not called?)")
            if param.grad is not None and param.grad.is_cuda:
                # Branches of the model may run on their own streams (mab_net.MabNet.encode), and autograd replays a branch's
                # backward on the stream it ran on: this gradient is only ordered with the stream this hook runs on.  Leave an
                # event behind for whichever stream ends up packing the bucket.
                # (Also for gradients finished on the training loop's own stream: the bucket may be packed on a side stream.)
                cs = torch.cuda.current_stream(param.grad.device)
                ev = torch.cuda.Event()
                ev.record(cs)
                self._xstream[b].append((cs, ev))
            if self._pending[b] == 0:           # the bucket is complete: pack it, and start its allreduce behind the backward
                self.fire_order.append(b)
                self._gather(b)
                self._gathered[b] = True
                if self.world > 1 and self.overlap:
                    self._reduce(b)
        return hook

    def zero_grad(self):
        """Drop the gradients (``p.grad = None``: autograd then hands its freshly computed gradient tensors over instead of
        launching one accumulate-add per parameter into a zeroed buffer — ~100 tiny kernels per step for the MARN models) and
        re-arm the per-bucket counters.  ``optimizer.zero_grad()`` / ``model.zero_grad()`` with either ``set_to_none`` are
        equivalent as far as the gradients go, but only this call re-arms the counters."""
        for b, members in enumerate(self._members):
            self._pending[b] = len(members)
            for p in members:
                p.grad = None
        self._handles = []
        self._fired = [False] * len(self.buckets)
        self._gathered = [False] * len(self.buckets)
        self.fire_order = []                # bucket indices in the order they completed during this step's backward
        # the training loop's stream (the one zero_grad() / finish() are called on); events of gradients finished elsewhere
        cuda = bool(self.buckets) and self.buckets[0].is_cuda
        self._home_stream = torch.cuda.current_stream(self.buckets[0].device) if cuda else None
        self._xstream: List[List] = [[] for _ in self.buckets]
        self._keepalive: List[torch.Tensor] = []
        self._packed_events: List = []

    def finish(self):
        """Block the current stream until every bucket's allreduce is complete.  Afterwards ``p.grad`` of every bucketed
        parameter that received a gradient is a view of its flat bucket holding the SUM over ranks."""
        if self._packed_events:                 # buckets packed on side streams: order this stream (and the allreduce) behind them
            cs = torch.cuda.current_stream(self.buckets[0].device)
            for ev in self._packed_events:
                cs.wait_event(ev)
            self._packed_events = []
        for b in range(len(self.buckets)):
            if not self._gathered[b]:           # a bucket whose hooks did not all fire (a parameter unused in THIS step)
                self._gather(b)
                self._gathered[b] = True
            if self.world > 1 and not self._fired[b]:
                self._reduce(b)
        for h in self._handles:
            h.wait()
        self._handles = []


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas, eps, weight_decay) semantics (L2-style decay, no amsgrad; what the reference's
    trainer builds at model_trainer.py:82) as ONE fused CUDA launch per gradient bucket of a
    ``GradAllReducer(..., flatten_params=True)``.  Parameters that are not in a bucket (never-used ones whose grad
    is None in the reference) are not touched, exactly as torch skips ``grad is None``."""

    def __init__(self, reducer: GradAllReducer, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if not reducer.flatten_params:
            raise RuntimeError("FusedAdam needs GradAllReducer(..., flatten_params=True)")
        self.reducer = reducer
        # a torch.optim.Optimizer with ONE param_group: lr schedulers (the reference's StepLR, model_trainer.py:83) read and
        # write param_groups[0]["lr"]; step() takes every hyper-parameter from there
        super().__init__([p for members in reducer._members for p in members],
                         dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.step_count = 0
        self.exp_avg = [torch.zeros_like(b) for b in reducer.param_buckets]
        self.exp_avg_sq = [torch.zeros_like(b) for b in reducer.param_buckets]

    @property
    def lr(self) -> float:
        return self.param_groups[0]["lr"]

    def step(self, closure=None) -> None:
        from . import _lib
        g0 = self.param_groups[0]
        self.step_count += 1
        for p, g, m, v in zip(self.reducer.param_buckets, self.reducer.buckets, self.exp_avg, self.exp_avg_sq):
            _lib.adam_step(p, g, m, v, g0["lr"], g0["betas"][0], g0["betas"][1], g0["eps"], g0["weight_decay"], self.step_count)

    def state_dict(self) -> dict:
        """Checkpointable optimizer state (the reference saves weights only, model_trainer.py:170-171; resuming a run needs this)."""
        g0 = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        return {"step": self.step_count, "param_group": g0, "exp_avg": [t.clone() for t in self.exp_avg],
                "exp_avg_sq": [t.clone() for t in self.exp_avg_sq]}

    def load_state_dict(self, sd: dict) -> None:
        if len(sd["exp_avg"]) != len(self.exp_avg) or any(a.shape != b.shape for a, b in zip(sd["exp_avg"], self.exp_avg)):
            raise RuntimeError("FusedAdam.load_state_dict: bucket layout differs from the checkpoint's")
        self.step_count = int(sd["step"])
        self.param_groups[0].update(sd["param_group"])
        for dst, src in zip(self.exp_avg, sd["exp_avg"]):
            dst.copy_(src)
        for dst, src in zip(self.exp_avg_sq, sd["exp_avg_sq"]):
            dst.copy_(src)

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.reducer.zero_grad()          # drops the gradients and re-arms the buckets (see GradAllReducer.zero_grad)
