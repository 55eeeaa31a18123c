"""Device-side batch prefetch: the B200 equivalent of VinceSolver.prefetch_batches / start_prefetch
(/root/reference/solvers/vince_solver.py:340-384).

The reference moves every batch to the GPU from a background Python thread (`val.to(self.model.device)`, :352-355) so
that the copy of batch i+1 overlaps the training step of batch i.  Here the same overlap is expressed with CUDA
streams instead of a thread: host batches (pinned memory) are copied on a dedicated copy stream into one of
`depth` device buffer sets, and the compute stream waits on the copy's event before it reads a buffer and records a
"released" event when the step that consumed it has been issued.  No tensor of the batch is ever re-allocated, so the
encoder plan's static arena and (later) CUDA graphs see stable pointers.

    pf = BatchPrefetcher(device, depth=2)
    pf.submit(host_batch)                 # async H2D of every tensor value on the copy stream
    batch = pf.next()                     # dict with device tensors; compute stream is ordered after the copy
    ... run the step on `batch` ...
    pf.release(batch)                     # the buffers may be overwritten once the work issued so far has finished

Non-tensor values (`data_source`, `num_frames`, `batch_types`, ...) are passed through unchanged, and - like the
reference (:356) - the host copy of `queue_data` is kept under "queue_data_cpu".
"""
import collections

import torch


class BatchPrefetcher:
    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchPrefetcher copies to a CUDA device (got %s); there is no CPU path" % self.device)
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = [dict(buffers={}, ready=torch.cuda.Event(), released=None) for _ in range(depth)]
        self._free = collections.deque(range(depth))
        self._pending = collections.deque()          # slot indices submitted but not yet handed out
        self.h2d_bytes = 0                           # bytes copied by the last submit()

    def can_submit(self):
        return bool(self._free)

    def submit(self, host_batch):
        """Start the H2D copy of one host batch.  Tensor values should be pinned for the copy to be asynchronous."""
        if not self._free:
            raise RuntimeError("BatchPrefetcher: all %d buffer sets are in flight; call release() first" % self.depth)
        idx = self._free.popleft()
        slot = self._slots[idx]
        out = {}
        nbytes = 0
        with torch.cuda.stream(self.copy_stream):
            if slot["released"] is not None:
                self.copy_stream.wait_event(slot["released"])       # the step that last read this slot is done
            for key, val in host_batch.items():
                if isinstance(val, torch.Tensor):
                    buf = slot["buffers"].get(key)
                    if buf is None or buf.shape != val.shape or buf.dtype != val.dtype:
                        buf = torch.empty(val.shape, dtype=val.dtype, device=self.device)
                        slot["buffers"][key] = buf
                    buf.copy_(val, non_blocking=True)
                    nbytes += val.numel() * val.element_size()
                    out[key] = buf
                else:
                    out[key] = val
            slot["ready"].record(self.copy_stream)
        if "queue_data" in host_batch:
            out["queue_data_cpu"] = host_batch["queue_data"]
        out["_prefetch_slot"] = idx
        slot["batch"] = out
        self._pending.append(idx)
        self.h2d_bytes = nbytes
        return nbytes

    def next(self):
        """Oldest submitted batch; the current stream is made to wait for its copy (no host synchronisation)."""
        if not self._pending:
            raise RuntimeError("BatchPrefetcher.next(): nothing submitted")
        idx = self._pending.popleft()
        slot = self._slots[idx]
        torch.cuda.current_stream(self.device).wait_event(slot["ready"])
        return slot.pop("batch")

    def release(self, batch):
        """Call after the last kernel reading `batch` has been issued on the current stream."""
        idx = batch["_prefetch_slot"]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._slots[idx]["released"] = ev
        self._free.append(idx)
