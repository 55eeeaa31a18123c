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


class CrossGpuShuffle:
    """MoCo's cross-GPU batch shuffle ("shuffle-BN") for one-process-per-GPU training.

    The reference shuffles the batch with ONE global randperm and lets nn.DataParallel split the shuffled batch
    across the GPUs (models/vince_model.py:137-142 + :35,125), so every GPU's train-mode BatchNorm normalises a random
    mix of clips and the key encoder cannot read "which frames belong together" off the batch statistics; the outputs
    are un-shuffled afterwards (:184-192).  With one process per GPU that is a permutation of the GLOBAL batch of
    world x B frames: rank r forwards frames perm[r*B:(r+1)*B], wherever they live.

    Every rank draws the same permutation from a shared-seed CPU generator (no communication), so all (source,
    destination) counts are known everywhere and the exchange is one variable-split all-to-all of the frames (NCCL over
    NVLink; 1 byte/pixel with uint8 frames) before the forward and one all-to-all of each [B, ...] output back.
    `exchange(x)` -> (x_shuffled, ctx); `restore(y, ctx)` returns rows to their owner, in original order.
    """

    def __init__(self, group=None, seed=0):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.gen = torch.Generator().manual_seed(seed)         # identical stream on every rank

    @staticmethod
    def plan(perm, world, rank, B):
        """Pure index arithmetic (tested on CPU).  perm: global permutation [world*B]; rank r forwards frames
        perm[r*B:(r+1)*B].  Returns (send_order [B] local indices grouped by destination, send_counts [world],
        recv_counts [world], place [B]: received row i (grouped by source, then by position in the source's send order)
        is row place[i] of this rank's shuffled batch)."""
        perm = perm.to(torch.int64)
        n = world * B
        pos_of = torch.empty(n, dtype=torch.int64)
        pos_of[perm] = torch.arange(n, dtype=torch.int64)       # global frame g sits at position pos_of[g] of the shuffle
        mine = torch.arange(rank * B, (rank + 1) * B, dtype=torch.int64)
        pos = pos_of[mine]                                      # where my frames go
        order = torch.argsort(pos)                              # by destination rank, then by position there
        send_order = order
        send_counts = torch.bincount(pos // B, minlength=world)
        want = perm[rank * B:(rank + 1) * B]                    # the frames I forward, in shuffled order
        src = want // B
        recv_counts = torch.bincount(src, minlength=world)
        # rows arrive grouped by source; within a source in increasing shuffled position (the sender's order)
        place = torch.argsort(src, stable=True)
        return send_order, send_counts.tolist(), recv_counts.tolist(), place

    def _a2a(self, x, in_splits, out_splits):
        out = torch.empty((sum(out_splits),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        self.dist.all_to_all_single(out, x.contiguous(), output_split_sizes=out_splits, input_split_sizes=in_splits,
                                    group=self.group)
        return out

    def exchange(self, x):
        """x: [B, ...] this rank's frames.  Returns (shuffled local batch [B, ...], ctx)."""
        B = x.shape[0]
        perm = torch.randperm(self.world * B, generator=self.gen)
        send_order, send_counts, recv_counts, place = self.plan(perm, self.world, self.rank, B)
        dev = x.device
        send_order_d, place_d = send_order.to(dev), place.to(dev)
        got = self._a2a(x.index_select(0, send_order_d), send_counts, recv_counts)
        shuffled = torch.empty_like(got)
        shuffled.index_copy_(0, place_d, got)
        return shuffled, (send_order_d, send_counts, recv_counts, place_d, perm)

    def restore(self, y, ctx):
        """y: [B, ...] outputs for the shuffled batch.  Returns the outputs of THIS rank's own frames, original order."""
        send_order_d, send_counts, recv_counts, place_d, _ = ctx
        back = self._a2a(y.index_select(0, place_d), recv_counts, send_counts)
        out = torch.empty_like(back)
        out.index_copy_(0, send_order_d, back)
        return out
