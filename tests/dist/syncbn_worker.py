"""world_size-2 gloo worker (CPU): the SyncBN statistic combination of sm3det_b200.lsk_functional (shifted sums all-reduced
across ranks) must reproduce torch's BatchNorm statistics over the concatenated batch.  Ragged per-rank row counts."""
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, sys.argv[1])
from sm3det_b200.lsk_functional import bn_batch_stats  # noqa: E402


def main():
    dist.init_process_group('gloo')
    rank, W = dist.get_rank(), dist.get_world_size()
    C = 40
    rows = 500 + 123 * rank
    g = torch.Generator().manual_seed(3 + rank)
    x = torch.randn(rows, C, generator=g) * 2.5 + 1.3
    rm = torch.randn(C, generator=torch.Generator().manual_seed(99)) * 0.2          # identical on every rank
    s1 = (x - rm).sum(0)
    s2 = ((x - rm) ** 2).sum(0)
    mean, var, n = bn_batch_stats(s1, s2, rows, rm, sync=True)
    xs = [None] * W
    dist.all_gather_object(xs, x)
    full = torch.cat(xs)
    assert n == full.shape[0]
    assert torch.allclose(mean, full.mean(0), atol=1e-5) and torch.allclose(var, full.var(0, unbiased=False), rtol=1e-4, atol=1e-5)
    # and the running-stat update BatchNormFn applies equals torch's on the full batch
    rmt, rvt = rm.clone(), torch.ones(C)
    F.batch_norm(full.t().reshape(1, C, -1), rmt, rvt, None, None, True, 0.1, 1e-5)
    rm2 = rm * 0.9 + 0.1 * mean
    rv2 = torch.ones(C) * 0.9 + 0.1 * var * (n / (n - 1))
    assert torch.allclose(rm2, rmt, atol=1e-5) and torch.allclose(rv2, rvt, rtol=1e-4, atol=1e-5)
    print(f'rank {rank}: syncbn stats ok')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
