"""Input staging — pinned-host double buffering ahead of the training step.

The reference feeds the model from a host-side pipeline (examples/ml_perf/dataloader.py:67-133 builds the batches,
examples/ml_perf/main.py:35-106 hands them to the trainer).  On the B200 path the step itself is ~8 ms, and the ids of a
multi-hot Criteo batch are large (ml_perf hotness sum 214 x int64 x 65536 examples = 112 MB, ~4 ms of PCIe), so the
host->device copy of batch t+1 has to run under the compute of batch t:

    for ids, labels in prefetch(host_batches, depth=2):      # device tensors, already resident
        model.train_on_batch(ids, labels, optimizer)

`prefetch` owns `depth` slots of (pinned host, device) buffers and one copy stream.  Batches that are already pinned are
copied straight from where they are; pageable batches go through the slot's pinned buffers first.  Ordering is carried by
CUDA events only (the consumer's stream waits for a slot's copy, the copy stream waits until the consumer has released the
slot), so nothing blocks the host except the optional pageable->pinned memcpy.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Sequence

import torch


class _Slot:
    def __init__(self):
        self.pinned = None
        self.dev = None
        self.ready = torch.cuda.Event()
        self.free = None


def prefetch(batches: Iterable[Sequence[torch.Tensor]], depth: int = 2, device: str | torch.device = "cuda") -> Iterator[tuple]:
    """Yields tuples of DEVICE tensors, one per host batch (a sequence of CPU tensors of fixed shapes), with the copies of
    the next `depth - 1` batches in flight on a side stream.  A yielded batch stays valid until the next one is requested."""
    if depth < 1:
        raise ValueError("prefetch: depth must be >= 1")
    dev = torch.device(device)
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [_Slot() for _ in range(depth)]
    it = iter(batches)
    issued = []                       # slots with a copy in flight, oldest first

    def issue(slot: _Slot) -> bool:
        try:
            host = next(it)
        except StopIteration:
            return False
        host = tuple(host)
        if slot.dev is None:
            slot.dev = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host]
            slot.pinned = [None if t.is_pinned() else torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in host]
        with torch.cuda.stream(copy_stream):
            if slot.free is not None:
                copy_stream.wait_event(slot.free)              # the consumer is done with this slot's device buffers
            for t, p, d in zip(host, slot.pinned, slot.dev):
                if not t.is_pinned():
                    if p is None:
                        p = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                    p.copy_(t)                                  # host memcpy into the pinned staging buffer
                    t = p
                d.copy_(t, non_blocking=True)
            slot.ready.record(copy_stream)
        return True

    for s in slots:
        if issue(s):
            issued.append(s)
        else:
            break
    prev = None
    while True:
        main = torch.cuda.current_stream(dev)
        if prev is not None:                                    # everything enqueued on the batch just consumed ...
            prev.free = torch.cuda.Event()
            prev.free.record(main)                              # ... precedes the reuse of its slot
            if issue(prev):
                issued.append(prev)
        if not issued:
            return
        cur = issued.pop(0)
        main.wait_event(cur.ready)
        yield tuple(cur.dev)
        prev = cur
