"""Data-parallel gradient exchange for the hot path.

The reference trains under ``MMDistributedDataParallel(..., broadcast_buffers=False)``
(mmdet3d/apis/train.py:105-109): one replica per GPU, gradients averaged over the ranks, nothing else
exchanged. The step of this path is launch-bound in places (about 2900 kernels in 35 ms), so the
per-parameter autograd hooks, bucket copies and mid-backward NCCL kernels of the generic wrapper cost more
than they hide. ``GradientExchange`` does the same averaging with one flat buffer:

  backward (plain module, no hooks)  ->  one multi-tensor copy of every gradient into ``flat``
  ->  ONE all-reduce(avg) of ``flat`` over NCCL / NVSwitch  ->  ``p.grad`` re-pointed at views of ``flat``

so the optimizer and the gradient clip read the averaged values without a copy back.
"""
import torch
import torch.distributed as dist


class GradientExchange:
    """Averages the gradients of ``params`` over the process group in one collective per step.

    Every parameter that requires grad must have received a gradient when ``exchange()`` is called (freeze
    structurally unused parameters first, as bench.py and tests/test_ddp_gloo.py do) - a missing one raises,
    it is never silently treated as zero.
    """

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        self.group = group
        self.world = dist.get_world_size(group)
        first = self.params[0]
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=first.dtype, device=first.device)
        self.views = []
        off = 0
        for p in self.params:
            assert p.dtype == first.dtype and p.device == first.device
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        backend = dist.get_backend(group)
        self._avg = dist.ReduceOp.AVG if backend == "nccl" else None

    @staticmethod
    def broadcast_initial_state(module, src=0, group=None):
        """Every replica starts from rank ``src``'s parameters and buffers (what the generic wrapper does once at
        construction); one coalesced broadcast per dtype."""
        by_dtype = {}
        for t in list(module.parameters()) + list(module.buffers()):
            by_dtype.setdefault(t.dtype, []).append(t.data)
        for tensors in by_dtype.values():
            flat = torch.cat([t.reshape(-1) for t in tensors])
            dist.broadcast(flat, src=src, group=group)
            off = 0
            for t in tensors:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def exchange(self):
        grads = []
        for p in self.params:
            if p.grad is None:
                raise RuntimeError("GradientExchange: a trainable parameter of shape %s received no gradient"
                                   % (tuple(p.shape),))
            grads.append(p.grad)
        # gradients accumulated in place since the last exchange (zero_grad(set_to_none=False)) already live in flat
        todo = [(v, g) for v, g in zip(self.views, grads) if g.data_ptr() != v.data_ptr()]
        if todo:
            torch._foreach_copy_([v for v, _ in todo], [g for _, g in todo])
        if self._avg is not None:
            dist.all_reduce(self.flat, op=self._avg, group=self.group)
        else:
            dist.all_reduce(self.flat, group=self.group)
            self.flat.div_(self.world)
        for p, v in zip(self.params, self.views):
            p.grad = v
