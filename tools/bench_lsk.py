#!/usr/bin/env python
"""Side measurement (not the bench.py contract): fwd+bwd images/s of BASELINE config 5 -- SM3Det LSKNet-S MoE backbone
(configs/SM3Det/SM3Det_lsk_s.py:13-25), bs=4 per GPU, 1024x1024, fp32, noisy gating + dropout as configured.
  python tools/bench_lsk.py [--batch 4] [--size 1024] [--steps 5]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sm3det_b200 import LSKNet_moe_MultiInput  # noqa: E402
from sm3det_b200.synth import make_images  # noqa: E402

KW = dict(MoE_Block_inds_fc1=[[], [0], [0, 2], [0]], MoE_Block_inds_fc2=[[], [0], [0, 2], [0]], num_experts=4, top_k=2,
          embed_dims=[64, 128, 320, 512], depths=[2, 2, 4, 2], drop_rate=0.1, drop_path_rate=0.,
          norm_cfg=dict(type='SyncBN', requires_grad=True))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--size', type=int, default=1024)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    a = ap.parse_args()
    torch.manual_seed(0)
    net = LSKNet_moe_MultiInput(**KW).cuda().train()
    x = make_images(a.batch, a.size, a.size, seed=3).cuda()

    def step():
        outs, loss = net(x)
        (sum(o.mean() for o in outs) + loss).backward()
        net.zero_grad(set_to_none=True)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({'metric': 'LSKNet-S MoE backbone images/s (fwd+bwd)', 'value': a.batch / ms * 1e3, 'ms_per_step': ms,
                      'batch': a.batch, 'size': a.size, 'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == '__main__':
    main()
