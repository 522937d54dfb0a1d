"""Input side of the hot path: what `ModelTrainer.train_network` does with a collated batch before it
calls the model (model_trainer.py:99-105) — move the tensors to the device, average the four RoBERTa layers and
concatenate the acoustic features — as a double-buffered pinned-memory feeder plus one fused device pass.

    feeder = DeviceFeeder(loader, device)            # loader yields the reference's collate_fn tuples
    for x, qmask, umask, label in feeder:            # x [L, B, 1124] is already assembled on the device
        logp, _, _ = model(x, qmask, umask)

``LengthBucketBatchSampler`` + ``collate_dialogues`` are the host side in front of it: length-bucketed batches (across
steps; the shards of one step share one padded length) collated straight into pinned buffers in the reference's layout.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Sequence, Tuple

import torch

from . import _lib


def assemble_input(r1: torch.Tensor, r2: torch.Tensor, r3: torch.Tensor, r4: torch.Tensor, acouf: torch.Tensor) -> torch.Tensor:
    """``torch.cat(((r1 + r2 + r3 + r4) / 4, acouf), dim=-1)`` (model_trainer.py:104-105) in one kernel; bit-identical."""
    ts = [t.contiguous() for t in (r1, r2, r3, r4, acouf)]
    if not all(t.is_cuda and t.dtype == torch.float32 for t in ts):
        raise RuntimeError("assemble_input: float32 CUDA tensors only (there is no CPU path)")
    return _lib.assemble_input(*ts)


class DeviceFeeder:
    """Iterates a loader of reference-style batches ``(r1, r2, r3, r4, visuf, acouf, qmask, umask, label, ...)``
    (dataloader.py:29-47) and yields ``(x, qmask, umask, label)`` on ``device``: batch i+1 is staged through pinned
    host buffers and copied on a side stream while batch i is being consumed."""

    def __init__(self, loader: Iterable[Sequence[torch.Tensor]], device: torch.device):
        self.loader, self.device = loader, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)

    def _stage(self, batch):
        r1, r2, r3, r4, _visuf, acouf, qmask, umask, label = batch[:9]
        host = [t if t.is_pinned() else t.pin_memory() for t in (r1, r2, r3, r4, acouf, qmask, umask, label)]
        with torch.cuda.stream(self.stream):
            dev = [t.to(self.device, non_blocking=True) for t in host]
            x = assemble_input(*dev[:5])
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return (x, dev[5], dev[6], dev[7]), ev, host

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]]:
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            (out, ev, _host), nxt = nxt, None
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                pass
            torch.cuda.current_stream(self.device).wait_event(ev)
            for t in out:
                t.record_stream(torch.cuda.current_stream(self.device))
            yield out


# ------------------------------------------------------------------------------------------------
# Length-bucketed batching + pinned collate.  The reference draws batches with a
# SubsetRandomSampler (dataloader.py:143-150) and pads every batch to its longest dialogue (pad_sequence,
# dataloader.py:45-47): with IEMOCAP's length spread (res.csv: mean 52, sigma 17, max 110) a random batch of 32
# is ~45 % padding, all of which the recurrence kernels still step through.  Bucketing is across STEPS only:
# the shards of one global step are always padded to the same length.
# ------------------------------------------------------------------------------------------------
class LengthBucketBatchSampler:
    """``batch_sampler`` for ``torch.utils.data.DataLoader``: every epoch shuffles the dialogue indices, cuts them into
    pools of ``pool_batches`` global batches, sorts each pool by length and emits its batches in random order — every
    dialogue exactly once per epoch, batches of similar length, different composition every epoch.  With ``world > 1``
    rank ``rank`` receives every ``world``-th dialogue of the length-sorted global batch (so shard workloads match);
    ``pad_to`` of the same step is identical on all ranks (``global_max_len(step)``)."""

    def __init__(self, lengths: Sequence[int], batch_size: int, pool_batches: int = 8, shuffle: bool = True, seed: int = 0,
                 world: int = 1, rank: int = 0, drop_last: bool = False):
        if batch_size < 1 or world < 1 or not 0 <= rank < world or batch_size % world:
            raise ValueError("batch_size must be a positive multiple of world, 0 <= rank < world")
        self.lengths = [int(n) for n in lengths]
        self.batch_size, self.pool_batches, self.shuffle, self.seed = batch_size, max(1, pool_batches), shuffle, seed
        self.world, self.rank, self.drop_last, self.epoch = world, rank, drop_last, 0
        self._plan = None

    def set_epoch(self, epoch: int) -> None:
        self.epoch, self._plan = epoch, None

    def _global_batches(self):
        if self._plan is not None:
            return self._plan
        g = torch.Generator().manual_seed(self.seed + self.epoch)
        n = len(self.lengths)
        order = torch.randperm(n, generator=g).tolist() if self.shuffle else list(range(n))
        pool = self.batch_size * self.pool_batches
        batches = []
        for s in range(0, n, pool):
            chunk = sorted(order[s:s + pool], key=lambda i: (self.lengths[i], i))
            batches += [chunk[b:b + self.batch_size] for b in range(0, len(chunk), self.batch_size)]
        if self.drop_last:
            batches = [b for b in batches if len(b) == self.batch_size]
        if self.shuffle:
            batches = [batches[i] for i in torch.randperm(len(batches), generator=g).tolist()]
        self._plan = batches
        return batches

    def global_max_len(self, step: int) -> int:
        return max(self.lengths[i] for i in self._global_batches()[step])

    def __len__(self) -> int:
        return len(self._global_batches())

    def __iter__(self):
        for b in self._global_batches():
            yield b[self.rank::self.world] if self.world > 1 else b

    def padding_fraction(self) -> float:
        """Padded positions / all positions over one epoch of global batches (what the kernels step through for nothing)."""
        pad = tot = 0
        for b in self._global_batches():
            L = max(self.lengths[i] for i in b)
            tot += L * len(b)
            pad += L * len(b) - sum(self.lengths[i] for i in b)
        return pad / max(tot, 1)


def collate_dialogues(samples, pad_to: int = 0, pin: bool = True):
    """The reference's ``collate_fn`` (dataloader.py:45-47) written straight into pinned host buffers: fields 0-6
    (four RoBERTa layers, visual, acoustic, qmask) are padded time-major ``[L,B,d]``, fields 7-8 (umask, label) batch-major
    ``[B,L]``, the rest (dialogue ids) stay lists.  ``pad_to`` > 0 pads to that length instead of the batch maximum (shards
    of one global step must agree).  Values are identical to ``pad_sequence``."""
    B = len(samples)
    L = max(max(int(s[0].shape[0]) for s in samples), int(pad_to))
    pin = pin and torch.cuda.is_available()
    out = []
    for f in range(len(samples[0])):
        first = samples[0][f]
        if not torch.is_tensor(first):
            out.append([s[f] for s in samples])
            continue
        tail = tuple(first.shape[1:])
        shape = (L, B) + tail if f < 7 else (B, L) + tail
        buf = torch.zeros(shape, dtype=first.dtype, pin_memory=pin)
        for b, s in enumerate(samples):
            n = int(s[f].shape[0])
            if f < 7:
                buf[:n, b] = s[f]
            else:
                buf[b, :n] = s[f]
        out.append(buf)
    return out
