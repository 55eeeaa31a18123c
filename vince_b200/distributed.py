"""Multi-GPU key exchange: one process per GPU, full weight + queue replica each, and ONE collective
per step - an NCCL all-gather of the new keys [B, D] of every rank, in rank order, into every replica's ring buffer
(enqueue == all-gather).  Replaces the reference's single-process nn.DataParallel (models/vince_model.py:35,125),
whose implicit gather moved [B, C, 7, 7] features to GPU 0 every forward.

`torch.distributed` is used for rendezvous only (broadcasting the ncclUniqueId); the data path is
libvince_b200's own communicator (vince_comm_init / vince_allgather_enqueue).

`ring_slices` is the pure host-side arithmetic of where gathered rows land (shared with StorageQueue and tested on
CPU, including with a world_size-2 gloo group).
"""
import ctypes

import torch

from . import _lib, ops


def ring_slices(tail, n, K):
    """Rows [0, n) enqueued at `tail` into a ring of K rows, following storage_queue.py:31-49 (n <= K):
    returns ([(src_row, dst_row, count), ...], new_tail, wrapped)."""
    if n > K:
        raise ValueError("ring_slices: %d rows do not fit a queue of %d in one call" % (n, K))
    if n == 0:
        return [], tail, False
    if tail + n > K:
        first = K - tail
        out = []
        if first > 0:
            out.append((0, tail, first))
        out.append((first, 0, n - first))
        return out, n - first, True
    return [(0, tail, n)], tail + n, False


class KeyGather:
    """All-gather + enqueue of the per-rank keys.  Build once per process after init_process_group()."""

    def __init__(self, device, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self.comm = ctypes.c_void_p()
        self._scratch = None
        lib = _lib.lib()
        uid = torch.zeros((128,), dtype=torch.uint8)
        if self.rank == 0:
            buf = (ctypes.c_uint8 * 128)()
            _lib.check(lib.vince_comm_unique_id(buf), "vince_comm_unique_id")
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        backend = dist.get_backend(group)
        if backend == "nccl":
            uid = uid.to(self.device)
        dist.broadcast(uid, src=0, group=group)
        raw = bytes(uid.cpu().tolist())
        with torch.cuda.device(self.device):
            _lib.check(lib.vince_comm_init(ctypes.byref(self.comm), raw, self.world, self.rank), "vince_comm_init")

    def enqueue(self, queue, keys, item_images=None, data_source=None, ema=None):
        """queue: vince_b200.StorageQueue; keys: [n_local, D] fp32 CUDA.  Every rank ends with identical queues.
        ema=(table_dev, n_chunks, momentum): also apply the momentum EMA in the scatter kernel's launch."""
        n_local, D = keys.shape
        total = n_local * self.world
        if self._scratch is None or self._scratch.numel() < total * D:
            self._scratch = torch.empty((total * D,), device=self.device, dtype=torch.float32)
        if queue._shadow_is_stale():
            queue._refresh_shadow()
        table, n_chunks, momentum = ema if ema is not None else (None, 0, 0.0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().vince_allgather_enqueue_ema(
                self.comm, ops._ptr(keys.detach().contiguous(), torch.float32, "keys"), n_local, D,
                ops._ptr(queue.vector_queue, torch.float32, "queue"),
                ops._ptr(queue.vector_queue_tf32, torch.float32, "queue_tf32"), queue.maxsize, queue.current_tail,
                ops._ptr(self._scratch, torch.float32, "scratch"),
                ops._ptr(table, torch.uint8, "ema table") if n_chunks else None, n_chunks, float(momentum),
                float(1 - momentum), ops._stream()), "vince_allgather_enqueue_ema")
        images = item_images if item_images is not None else [None] * total
        queue.bookkeep(total, images, data_source)

    def close(self):
        if self.comm:
            _lib.lib().vince_comm_destroy(self.comm)
            self.comm = ctypes.c_void_p()
