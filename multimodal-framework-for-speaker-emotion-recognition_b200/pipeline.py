"""Input side of the hot path: what `ModelTrainer.train_network` does with a collated batch before it
calls the model (model_trainer.py:99-105) — move the tensors to the device, average the four RoBERTa layers and
concatenate the acoustic features — as a double-buffered pinned-memory feeder plus one fused device pass.

    feeder = DeviceFeeder(loader, device)            # loader yields the reference's collate_fn tuples
    for x, qmask, umask, label in feeder:            # x [L, B, 1124] is already assembled on the device
        logp, _, _ = model(x, qmask, umask)
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
