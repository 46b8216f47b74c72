"""world-2 gloo worker: sm3det_b200.graphed.allreduce_gradients averages every gradient except the skipped names."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sm3det_b200.graphed import allreduce_gradients  # noqa: E402


def main():
    dist.init_process_group('gloo')
    rank, W = dist.get_rank(), dist.get_world_size()
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    net[1].bias.grad = None                                     # parameters without a gradient are left alone
    named = list(net.named_parameters())
    n = allreduce_gradients(None, named=named, skip=['0.bias'])
    assert n == 5 * 7 + 7 * 3
    mean = sum(range(1, W + 1)) / W
    assert torch.allclose(net[0].weight.grad, torch.full((7, 5), mean * 1))
    assert torch.allclose(net[0].bias.grad, torch.full((7,), float(rank + 1) * 2))       # skipped: still the local value
    assert torch.allclose(net[1].weight.grad, torch.full((3, 7), mean * 3))
    assert net[1].bias.grad is None
    print(f'rank {rank}: grad sync ok')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
