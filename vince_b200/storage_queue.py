"""Ring-buffer queue of momentum-encoder keys, resident in HBM.

Drop-in for /root/reference/utils/storage_queue.py:4-56 - same constructor, attributes (`maxsize`, `feat_size`,
`vector_queue`, `image_queue`, `data_source_queue`, `current_tail`, `full`) and methods (`clear`, `enqueue`,
`dequeue`, `__len__`), including the wrap-around recursion of enqueue (:35-43; `full` is set only when a wrap
happens, and tail == K is a legal resting state).

Differences, all additive:
  * the copies are done by the coalesced `vince_ema_enqueue` kernel, which in the same pass maintains
    `vector_queue_tf32`, a copy of the queue rounded (RN) to TF32 that the fused InfoNCE kernel streams through the
    tensor cores; `dequeue()` returns it under the extra key "queue_vectors_tf32";
  * the shadow follows out-of-band writes too: code that assigns `queue.vector_queue = ...` or mutates it with torch
    ops (`vector_queue.copy_(...)`, checkpoint restore) changes the tensor's identity / version counter, which
    `dequeue()` notices and answers by re-rounding the shadow;
  * multi-GPU: `vince_b200.distributed.KeyGather.enqueue(queue, keys, ...)` enqueues the rank-ordered all-gather of
    every rank's keys.
"""
import torch

from . import ops


class StorageQueue(object):
    def __init__(self, maxsize, feat_size, device=None, dtype=torch.float32):
        if dtype != torch.float32:
            raise TypeError("vince_b200.StorageQueue stores fp32 (the reference's dtype)")
        self.maxsize = maxsize
        self.feat_size = feat_size
        self.device = device
        self.dtype = dtype
        self._init_buffers()

    def _init_buffers(self):
        dev = torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("vince_b200.StorageQueue lives in GPU memory; got device %s (no CPU fallback)" % dev)
        if self.feat_size % 4 != 0:
            raise ValueError("feat_size must be a multiple of 4")
        # storage_queue.py:10-12 / :21-25: unit-norm random rows (initialisation only, not on the hot path)
        self.vector_queue = torch.nn.functional.normalize(
            torch.randn((self.maxsize, self.feat_size), device=dev, requires_grad=False, dtype=self.dtype), dim=-1
        ).contiguous()
        self.vector_queue_tf32 = torch.empty_like(self.vector_queue)
        self._refresh_shadow()
        self.image_queue = [None for _ in range(self.maxsize)]
        self.data_source_queue = [None for _ in range(self.maxsize)]
        self.current_tail = 0
        self.full = False

    def _refresh_shadow(self):
        vq = self.vector_queue
        if self.vector_queue_tf32 is None or self.vector_queue_tf32.shape != vq.shape or \
                self.vector_queue_tf32.device != vq.device:
            self.vector_queue_tf32 = torch.empty_like(vq)
        with torch.no_grad(), torch.cuda.device(vq.device):
            ops.round_tf32(vq.contiguous(), self.vector_queue_tf32)
        self._shadow_key = (vq.data_ptr(), vq._version)

    def _shadow_is_stale(self):
        # our own kernels write through raw pointers and never touch the version counter; any torch-level write does
        vq = self.vector_queue
        return self._shadow_key != (vq.data_ptr(), vq._version)

    def __len__(self):
        return len(self.image_queue)

    def clear(self):
        self._init_buffers()

    def load(self, vectors):
        """Overwrite the whole buffer (tests / checkpointing); keeps the TF32 shadow coherent."""
        with torch.no_grad():
            self.vector_queue.copy_(vectors)
        self._refresh_shadow()

    def enqueue(self, items, item_images, data_source):
        assert len(items) == len(item_images)
        num_items = items.shape[0]
        if self.current_tail + num_items > self.maxsize:
            num_start = self.maxsize - self.current_tail
            if num_start > 0:
                self._copy_rows(items[:num_start], self.current_tail)
                self.image_queue[self.current_tail:] = item_images[:num_start]
                self.data_source_queue[self.current_tail:] = [data_source] * num_start
            self.current_tail = 0
            self.full = True
            self.enqueue(items[num_start:], item_images[num_start:], data_source)
        else:
            self._copy_rows(items, self.current_tail)
            self.image_queue[self.current_tail: self.current_tail + num_items] = item_images
            self.data_source_queue[self.current_tail: self.current_tail + num_items] = [data_source] * num_items
            self.current_tail += num_items

    def _copy_rows(self, rows, at):
        if rows.shape[0] == 0:
            return
        rows = rows.detach()
        if rows.dtype != torch.float32 or rows.shape[1] != self.feat_size:
            raise ValueError("enqueue: expected fp32 rows of width %d" % self.feat_size)
        if self._shadow_is_stale():
            self._refresh_shadow()
        with torch.cuda.device(self.vector_queue.device):
            ops.ema_enqueue(None, 0, 0.0, self.vector_queue, self.vector_queue_tf32, rows.contiguous(), at)

    def bookkeep(self, n, item_images, data_source):
        """Advance tail/full and the host-side lists for n rows whose vectors were already written by a fused
        launch (VinceQueueModel.vince_update(..., enqueue=...) or the all-gather path)."""
        t = self.current_tail
        if t + n > self.maxsize:
            first = self.maxsize - t
            if first > 0:
                self.image_queue[t:] = item_images[:first]
                self.data_source_queue[t:] = [data_source] * first
            rest = n - first
            self.image_queue[0:rest] = item_images[first:]
            self.data_source_queue[0:rest] = [data_source] * rest
            self.current_tail = rest
            self.full = True
        else:
            self.image_queue[t:t + n] = item_images
            self.data_source_queue[t:t + n] = [data_source] * n
            self.current_tail = t + n

    def dequeue(self):
        if self._shadow_is_stale():
            self._refresh_shadow()
        return {
            "queue_vectors": self.vector_queue.detach(),
            "queue_images": self.image_queue,
            "queue_data_sources": self.data_source_queue,
            "queue_vectors_tf32": self.vector_queue_tf32,
        }
