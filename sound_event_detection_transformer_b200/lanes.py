"""Eval lanes: n independent copies of an eval-mode model, each with its own CUDA stream (and so its own workspace and
CUDA graph), driven round-robin from one host thread.  Batch i runs on lane i % n; with n = 2 the under-occupied phases of
one forward (the 21-query decoder, kernel tails, the stem) overlap with the next batch's backbone.  Every batch still is one
complete forward of the reference's `SEDT.forward` (sedt/sedt.py:60-100); only the order in which the GPU interleaves the
kernels of consecutive batches changes, never a result.

    lanes = EvalLanes([model_a, model_b])
    lanes.fork()                       # lane streams wait for work already queued on the caller's stream
    for i, x in enumerate(batches):
        with lanes.stream(i):          # everything in here (prefetch wait, forward, result copy) is ordered on lane i % n
            out = lanes.model(i)(x)
    lanes.join()                       # the caller's stream waits for every lane
"""
from __future__ import annotations

from typing import List, Sequence

import torch


class EvalLanes:
    def __init__(self, models: Sequence[torch.nn.Module], device: torch.device | None = None):
        if len(models) < 1:
            raise ValueError("EvalLanes needs at least one model")
        if len({id(m) for m in models}) != len(models):
            raise ValueError("EvalLanes: every lane needs its own model object (workspace and CUDA graph are per model)")
        for m in models:
            if m.training:
                raise ValueError("EvalLanes runs eval-mode models; call .eval() first")
        self.models: List[torch.nn.Module] = list(models)
        self.device = device if device is not None else next(models[0].parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("EvalLanes needs CUDA models: there is no CPU path")
        # one lane: the caller's own stream (exactly the plain single-stream loop)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in models] if len(models) > 1 else [None]

    def __len__(self) -> int:
        return len(self.models)

    def model(self, i: int) -> torch.nn.Module:
        return self.models[i % len(self.models)]

    def stream(self, i: int):
        s = self.streams[i % len(self.models)]
        return torch.cuda.stream(s) if s is not None else torch.cuda.stream(torch.cuda.current_stream(self.device))

    def fork(self) -> None:
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            if s is not None:
                s.wait_stream(cur)

    def join(self) -> None:
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            if s is not None:
                cur.wait_stream(s)
