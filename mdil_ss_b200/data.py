"""Host -> device input pipeline of the training loop (the reference's DataLoader hands pinned CPU tensors to
``.cuda()`` inside the step, train_new_task_step2.py:281-283: copy and compute are serial there).

``DevicePrefetcher`` issues the copies of the NEXT batch on a side stream while the current step runs, so the
63 MB / step of a 6 x 3 x 512 x 1024 fp32 batch (+ int64 labels) cost no step time.  Every batch is still copied from
host memory every step; only the overlap changes."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch


class DevicePrefetcher:
    def __init__(self, device: torch.device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._pending: Optional[Tuple[Tuple[torch.Tensor, ...], torch.cuda.Event]] = None

    def put(self, *host_tensors: torch.Tensor) -> None:
        """Start the asynchronous copy of one batch (pinned host tensors) to the device."""
        with torch.cuda.stream(self.stream):
            dev = tuple(t.to(self.device, non_blocking=True) for t in host_tensors)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (dev, ev)

    def get(self) -> Sequence[torch.Tensor]:
        """The batch whose copy was started by the last ``put`` (the consumer stream waits for it on the device)."""
        if self._pending is None:
            raise RuntimeError("DevicePrefetcher.get() without a pending put()")
        dev, ev = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)     # allocated on the copy stream, consumed on the compute stream
        return dev
