"""Side-stream host->device prefetch of clip batches, the B200 counterpart of the reference's
`data_prefetcher` (data_utils/DataLoad.py:304-336): while the forward of batch i runs on the compute
stream, the pinned host buffer of batch i+1 is copied on a copy stream into a small ring of
pre-allocated device buffers, so the PCIe transfer is hidden behind the kernels and no device memory
is allocated in the steady state."""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional

import torch


class ClipPrefetcher:
    def __init__(self, batches: Iterable[torch.Tensor], device: torch.device, depth: int = 2):
        self.it: Iterator[torch.Tensor] = iter(batches)
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.nslots = max(2, depth + 1)
        self.bufs: List[Optional[torch.Tensor]] = [None] * self.nslots
        self.ready = [torch.cuda.Event() for _ in range(self.nslots)]       # copy finished
        self.free = [torch.cuda.Event() for _ in range(self.nslots)]        # compute finished reading
        self.shapes: List[Optional[torch.Size]] = [None] * self.nslots
        self.pending: List[int] = []
        self.head = 0
        self.in_use: Optional[int] = None
        for ev in self.free:
            ev.record(torch.cuda.current_stream(device))
        for _ in range(self.nslots - 1):
            self._issue()

    def _issue(self) -> None:
        try:
            host = next(self.it)
        except StopIteration:
            return
        slot = self.head
        self.head = (self.head + 1) % self.nslots
        buf = self.bufs[slot]
        if buf is None or buf.numel() < host.numel() or buf.dtype != host.dtype:
            buf = torch.empty(host.numel(), dtype=host.dtype, device=self.device)
            self.bufs[slot] = buf
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])
            buf[: host.numel()].view(host.shape).copy_(host, non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        self.shapes[slot] = host.shape
        self.pending.append(slot)

    def next(self) -> Optional[torch.Tensor]:
        cur = torch.cuda.current_stream(self.device)
        if self.in_use is not None:                       # the previous batch has been consumed by now
            self.free[self.in_use].record(cur)
            self.in_use = None
            self._issue()
        if not self.pending:
            return None
        slot = self.pending.pop(0)
        cur.wait_event(self.ready[slot])
        self.in_use = slot
        shape = self.shapes[slot]
        return self.bufs[slot][: shape.numel()].view(shape)

    def __iter__(self):
        while True:
            b = self.next()
            if b is None:
                return
            yield b
