"""Host -> device feeder for packed dialogue batches (the step right before the hot path).

The reference collates on the CPU (``ERCCollate``, track_mm/mmbase.py:344-455), and its DataLoader hands the trainer
pageable tensors that ``accelerate`` moves to the GPU synchronously at the top of every step
(lumo/trainer/trainer.py:315-327).  At the scale of BASELINE config 5 one step's inputs are 6 GB, i.e. ~110 ms of
PCIe time against ~17 ms of kernels, so the copy has to overlap the previous step's compute:

  * ``DeviceFeeder`` owns ``depth`` sets of device buffers and a dedicated copy stream;
  * ``submit(batch)`` enqueues the H2D copies of a dict of (pinned) host tensors on the copy stream as soon as the
    buffer set is free (an event recorded by ``release``), and returns immediately;
  * ``get()`` makes the CURRENT stream wait for the oldest submitted batch and returns its device tensors;
  * ``release()`` marks that batch's buffers reusable once the work queued so far on the current stream is done.

Nothing here computes; it is stream/event plumbing around ``Tensor.copy_(non_blocking=True)``.
"""
from collections import deque

import torch


def pin(batch):
    """Pinned copies of a dict of CPU tensors (what a DataLoader with ``pin_memory=True`` hands out)."""
    return {k: (v if v.is_pinned() else v.pin_memory()) for k, v in batch.items()}


class DeviceFeeder:
    def __init__(self, device, depth=2, copy_streams=1, chunk_bytes=256 << 20):
        """``copy_streams`` > 1 splits large tensors into ``chunk_bytes`` pieces issued round-robin on several streams
        (more than one copy engine busy on the H2D direction); 1 = one stream, one copy per tensor."""
        assert depth >= 1 and copy_streams >= 1
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.extra_streams = [torch.cuda.Stream(device=self.device) for _ in range(copy_streams - 1)]
        self.chunk_bytes = chunk_bytes
        self._store = [dict() for _ in range(depth)]         # slot -> {key: flat uint8 device storage (high-water mark)}
        self._bufs = [dict() for _ in range(depth)]          # slot -> {key: view of the storage shaped like the last batch}
        self._free = [None] * depth                          # slot -> event after which the slot may be overwritten
        self._ready = deque()                                # (slot, event, keys) in submission order
        self._in_use = deque()                               # slots handed out by get() and not yet released
        self._next = 0
        self.h2d_bytes = 0

    def submit(self, batch):
        """Start copying ``batch`` (dict of host tensors; pinned memory makes the copies asynchronous)."""
        if len(self._ready) + len(self._in_use) >= self.depth:
            raise RuntimeError("DeviceFeeder: all %d buffer sets are in flight; call get()/release() first" % self.depth)
        slot = self._next
        self._next = (self._next + 1) % self.depth
        store, bufs = self._store[slot], self._bufs[slot]
        streams = [self.copy_stream] + self.extra_streams
        grew = False
        for k, h in batch.items():
            # Device storage is sized to the HIGH-WATER mark per key and handed out as views, so the reference collate's
            # changing Lmax (mmbase.py:354-455 pads every batch to its own longest dialogue) does not reallocate.
            nbytes = h.numel() * h.element_size()
            flat = store.get(k)
            if flat is None or flat.numel() < nbytes:
                if flat is not None:
                    for st in streams:                       # the outgoing block may still be read by queued copies
                        flat.record_stream(st)
                store[k] = flat = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
                grew = True
            bufs[k] = flat[:nbytes].view(h.dtype).view(h.shape) if nbytes else torch.empty(h.shape, dtype=h.dtype, device=self.device)
        if grew:
            # The caching allocator may have handed back a block that kernels ALREADY QUEUED on the caller's stream still
            # read (it was freed on that stream, which is all the allocator tracks).  The copies below run on other
            # streams, so they must not start before the caller's stream has reached this point.
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            for st in streams:
                st.wait_event(ev)
        for st in streams:
            if self._free[slot] is not None:
                st.wait_event(self._free[slot])
        turn = 0
        for k, h in batch.items():
            nbytes = h.numel() * h.element_size()
            self.h2d_bytes += nbytes
            if len(streams) == 1 or nbytes <= self.chunk_bytes or h.dim() == 0 or not h.is_contiguous():
                with torch.cuda.stream(streams[0]):
                    bufs[k].copy_(h, non_blocking=True)
                continue
            rows = h.size(0)
            step = max(1, int(rows * self.chunk_bytes // nbytes))
            for r0 in range(0, rows, step):
                with torch.cuda.stream(streams[turn % len(streams)]):
                    bufs[k][r0:r0 + step].copy_(h[r0:r0 + step], non_blocking=True)
                turn += 1
        for st in self.extra_streams:                       # the batch is ready when every stream is done
            e = torch.cuda.Event()
            e.record(st)
            self.copy_stream.wait_event(e)
        with torch.cuda.stream(self.copy_stream):
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._ready.append((slot, ev, tuple(batch.keys())))
        return slot

    def get(self):
        """Device tensors of the oldest submitted batch; the current stream waits for its copies."""
        slot, ev, keys = self._ready.popleft()
        torch.cuda.current_stream(self.device).wait_event(ev)
        self._in_use.append(slot)
        return {k: self._bufs[slot][k] for k in keys}

    def release(self):
        """The oldest batch handed out by get() may be overwritten once the current stream reaches this point."""
        slot = self._in_use.popleft()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[slot] = ev
