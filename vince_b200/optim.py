"""Fused multi-tensor SGD + gradient all-reduce for the B200 training step.

`FusedSGD` is a drop-in for the optimiser the reference solver builds at /root/reference/solvers/vince_solver.py:252-256
(`torch.optim.SGD(model.parameters(), lr, momentum=0.9, weight_decay=1e-4)`): same constructor arguments, `param_groups`
(so solver_runner.py:36-43's learning-rate warm-up, which rewrites `param_group["lr"]`, keeps working), `zero_grad()` and
`step()` - but one kernel launch (`vince_sgd_step`) over a device-resident chunk table instead of torch's per-tensor
loops, with torch.optim.SGD's exact update rule:
    d = grad + weight_decay * p;  buf = d (first step) | momentum * buf + d;  p -= lr * buf.

`GradAllReduce` sums the flat gradient buffer of `vince_b200.backward.GradSlots` across ranks with ONE ncclAllReduce
(`vince_allreduce_sum`, the communicator of `KeyGather`) and folds the 1/world average into the optimiser's grad scale.
"""
import numpy as np
import torch

from . import _lib, ops


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        if dampening != 0.0 or nesterov:
            raise NotImplementedError("FusedSGD implements the reference's configuration: dampening=0, nesterov=False")
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        super().__init__(params, defaults)
        self.grad_scale = 1.0          # multiplied into every gradient (1/world after a summed all-reduce)
        self._tables = {}
        self._steps = 0

    def _table(self, gi, group):
        plist = [p for p in group["params"] if p.grad is not None]
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in plist)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1], cached[2], False
        chunks, fresh = [], False
        for p in plist:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                raise RuntimeError("FusedSGD: parameters and gradients must be contiguous fp32 CUDA tensors")
            st = self.state[p]
            buf_ptr = 0
            if group["momentum"] != 0:
                # torch.optim.SGD starts the buffer as a copy of the first d; a ZERO buffer under the general rule
                # gives the same bits (momentum * 0 + d == d), so no first-step special case is needed
                if "momentum_buffer" not in st:
                    st["momentum_buffer"] = torch.zeros_like(p)
                buf_ptr = st["momentum_buffer"].data_ptr()
            n = p.numel()
            for off in range(0, n, 16384):
                chunks.append((p.data_ptr() + 4 * off, p.grad.data_ptr() + 4 * off, buf_ptr + 4 * off if buf_ptr else 0,
                               min(16384, n - off)))
        arr = np.array(chunks, dtype=np.int64).reshape(-1, 4)
        table = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(plist[0].device) if plist else None
        self._tables[gi] = (key, table, len(chunks))
        return table, len(chunks), True

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            table, n, _ = self._table(gi, group)
            if not n:
                continue
            with torch.cuda.device(table.device):
                _lib.check(_lib.lib().vince_sgd_step(ops._ptr(table, torch.uint8, "sgd table"), n, float(group["lr"]),
                                                     float(group["momentum"]), float(group["weight_decay"]),
                                                     float(self.grad_scale), 0, ops._stream()), "vince_sgd_step")
        self._steps += 1
        return loss


class GradAllReduce:
    """model.grad_sync hook: sum the flat gradient buffer over all ranks (one ncclAllReduce on the current stream).
    Pair it with FusedSGD.grad_scale = 1 / world (or divide the loss) to average."""

    def __init__(self, key_gather):
        self.gather = key_gather            # vince_b200.distributed.KeyGather owns the communicator

    def __call__(self, slots):
        flat = slots.flat
        with torch.cuda.device(flat.device):
            _lib.check(_lib.lib().vince_allreduce_sum(self.gather.comm, ops._ptr(flat, torch.float32, "grads"),
                                                      flat.numel(), ops._stream()), "vince_allreduce_sum")
