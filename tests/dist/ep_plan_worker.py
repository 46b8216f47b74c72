"""world_size-2 gloo worker (CPU): checks the expert-parallel exchange plan (sm3det_b200.expert_parallel._build_plan) by
emulating the peer gathers with all_gathered buffers.  Launched by tests/test_ep_plan.py through torch.distributed.run."""
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, sys.argv[1])
from sm3det_b200.expert_parallel import EPContext, _build_plan  # noqa: E402


def local_plan(top_idx, E):
    """CPU mirror of sm3_moe_plan / sm3_moe_assign: expert-major slots, segments padded to 128 rows."""
    T, k = top_idx.shape
    counts = torch.bincount(top_idx.reshape(-1), minlength=E).to(torch.int32)
    tiles = (counts + 127) // 128
    seg_begin = (torch.cumsum(tiles, 0) - tiles) * 128
    max_tiles = (T * k + 127) // 128 + E
    tile_group = torch.full((max_tiles,), 12345, dtype=torch.int32)          # garbage beyond the live tiles, as on the GPU
    tile_group[:int(tiles.sum())] = torch.repeat_interleave(torch.arange(E, dtype=torch.int32), tiles.long())
    pair_token = torch.full((max_tiles * 128,), -1, dtype=torch.int32)
    slot_of = torch.empty((T, k), dtype=torch.int32)
    cursor = seg_begin.clone()
    for t in range(T):
        for j in range(k):
            e = int(top_idx[t, j])
            slot_of[t, j] = cursor[e]
            pair_token[cursor[e]] = t
            cursor[e] += 1
    return counts, seg_begin.to(torch.int32), tile_group, torch.tensor([int(tiles.sum())], dtype=torch.int32), pair_token, slot_of


def main():
    dist.init_process_group('gloo')
    rank, W = dist.get_rank(), dist.get_world_size()
    E, k, C = 4, 2, 8
    T = 300 + 77 * rank                       # ragged: different token counts per rank
    g = torch.Generator().manual_seed(100 + rank)
    v = torch.randn(T, C, generator=g)
    logits = torch.randn(T, E, generator=g)
    if rank == 1:
        logits[:, 3] = -1e9                   # an expert that receives nothing from this rank
    top_idx = logits.topk(k, dim=1).indices
    counts, seg_begin, tile_group, num_tiles, pair_token, slot_of = local_plan(top_idx, E)
    R_s = pair_token.numel()
    ctx = EPContext.__new__(EPContext)
    ctx.group, ctx.world, ctx.rank = dist.group.WORLD, W, rank
    meta = [torch.empty(2, E, dtype=torch.int32) for _ in range(W)]
    dist.all_gather(meta, torch.stack([counts, seg_begin]))
    allm = torch.stack(meta)
    P = _build_plan(ctx, allm[:, 0], allm[:, 1], tile_group, num_tiles, pair_token, E, R_s, torch.device('cpu'))
    # every rank publishes (v, pair list): emulate the P2P reads with all_gather_object
    pub = [None] * W
    dist.all_gather_object(pub, (v, pair_token))
    R_d = P['R_d']
    xr = torch.zeros(max(R_d, 1), C)
    for r in range(R_d):
        s = int(P['src_rank'][r])
        if s >= 0:
            tok = int(pub[s][1][int(P['src_slot'][r])])
            assert tok >= 0
            xr[r] = pub[s][0][tok]
    # (1) expert-side rows: segment of local expert el holds exactly the tokens every source routed to it, source-major
    El = E // W
    for el in range(El):
        gidx = rank * El + el
        want = []
        for s in range(W):                    # source-major: the tokens rank s routed to expert gidx, in its slot order
            lo, n = int(allm[s, 1, gidx]), int(allm[s, 0, gidx])
            want.append(pub[s][0][pub[s][1][lo:lo + n].long()])
        want = torch.cat(want)
        b, e = int(P['seg_begin'][el]), int(P['seg_end'][el])
        assert e - b == want.shape[0] and torch.equal(xr[b:e], want), (rank, el)
        assert b % 128 == 0
        assert all(int(P['tile_group'][i]) == el for i in range(b // 128, (e + 127) // 128))
    assert int(P['num_tiles']) * 128 == R_d
    # (2) combine descriptors: my slot l must find, on its owner, the row that was gathered from my token
    rows = [None] * W
    dist.all_gather_object(rows, xr)
    for l in range(R_s):
        d = int(P['comb_rank'][l])
        if pair_token[l] < 0 or l >= int(num_tiles) * 128:
            assert d == -1
            continue
        assert d == int(tile_group[l // 128]) // El
        assert torch.equal(rows[d][int(P['comb_row'][l])], v[int(pair_token[l])]), (rank, l)
    # (3) every pair is covered exactly once
    n_pairs = [None] * W
    dist.all_gather_object(n_pairs, int((P['src_rank'][:R_d] >= 0).sum()) if R_d else 0)
    tot = torch.tensor([T * k])
    dist.all_reduce(tot)
    assert sum(n_pairs) == int(tot)
    print(f'rank {rank}: ep plan ok (R_d={R_d}, pairs here {n_pairs[rank]})')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
