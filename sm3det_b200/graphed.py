"""One forward+backward of a backbone captured into a CUDA graph.

The backbones issue thousands of short kernels per step through the C-ABI (LSKNet-S at 4 images per GPU: ~5900 launches for
~40 ms of GPU work), so at small per-GPU batches the step is bound by the host's launch rate, not by the device.  Every op
on the path is capture-safe by construction: no host read-backs (the MoE plan, expert tile counts, SyncBN counts and the
dropout seed all stay in device memory), all scratch comes from torch's caching allocator, and the only random numbers are
torch's graph-safe philox draws (gating noise, drop-path masks, the dropout seed tensor read by ``sm3_dropout_dev``).

    step = GraphedStep(lambda x: fwd_bwd(net, x), [example_x], net.parameters())
    loss = step(x)            # copies x into the static input, replays; gradients of this step are in p.grad

Semantics of a replay: gradients are REPLACED (as after ``zero_grad(set_to_none=True)`` + backward), the weight operand images
are re-split from the current weights inside the graph (so an optimizer step between replays is honoured), running statistics
and RNG state advance exactly as in eager mode.  Do not call ``zero_grad(set_to_none=True)`` between replays: the .grad
tensors live in the graph's memory pool and are rewritten by the next replay.
"""
from typing import Callable, Iterable, List, Sequence

import torch

from . import _lib


class GraphedStep:
    def __init__(self, step_fn: Callable[..., torch.Tensor], example_inputs: Sequence[torch.Tensor],
                 parameters: Iterable[torch.nn.Parameter], warmup: int = 3, invalidate: Sequence = (),
                 capture_error_mode: str = 'global'):
        """step_fn(*inputs) must run forward AND backward and return a tensor (or tuple of tensors) to read after a replay.
        ``invalidate``: objects with an ``invalidate()`` method (the backbones' PackCache) cleared before capture so that
        the weight split kernels are part of the graph.  ``capture_error_mode='thread_local'`` when other threads issue CUDA
        calls during capture (NCCL's watchdog in a multi-GPU job)."""
        self.params: List[torch.nn.Parameter] = [p for p in parameters if p.requires_grad]
        self.static_inputs = [x.detach().clone() for x in example_inputs]
        self._fn = step_fn
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):                     # lazy initialisation (function attributes, allocator pools) off the graph
            for _ in range(max(int(warmup), 1)):
                step_fn(*self.static_inputs)
                for p in self.params:
                    p.grad = None
        cur.wait_stream(side)
        torch.cuda.synchronize()
        for c in invalidate:
            c.invalidate()
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.LAUNCHES
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.static_output = step_fn(*self.static_inputs)
        self.launches_per_replay = _lib.LAUNCHES - before          # C-ABI kernel launches recorded in the graph
        self.replays = 0

    def __call__(self, *inputs: torch.Tensor):
        for s, x in zip(self.static_inputs, inputs):
            if x.data_ptr() != s.data_ptr():
                s.copy_(x, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.static_output


def allreduce_gradients(parameters: Iterable[torch.nn.Parameter], group=None, average: bool = True, skip: Sequence[str] = (),
                        named: Sequence = ()):
    """Data-parallel gradient sync as ONE flat all-reduce (capturable in a CUDA graph, unlike DDP's bucket hooks).

    All gradients are packed into one buffer, summed over the group (NVLS / NVLink ring, NCCL's choice) and unpacked in
    place.  ConvNeXt-T e8 carries 136 M parameters = 544 MB: measured on 4 GPUs the whole un-overlapped sync (collective,
    pack / unpack copies, rank skew) costs 3.9 ms of a 48.6 ms step, no more than DDP's bucket-by-bucket overlap cost when
    its NCCL kernels queued behind persistent GEMMs that own every SM.  ``named``:
    optional (name, parameter) pairs with ``skip`` = names left out (the expert parameters an expert-parallel rank owns)."""
    import torch.distributed as dist
    if named:
        skip = set(skip)
        ps = [p for n, p in named if n not in skip and p.grad is not None]
    else:
        ps = [p for p in parameters if p.grad is not None]
    if not ps or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    grads = [p.grad for p in ps]
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, group=group)
    if average:
        flat.div_(dist.get_world_size(group))
    torch._foreach_copy_(grads, list(torch._utils._unflatten_dense_tensors(flat, grads)))
    return flat.numel()
