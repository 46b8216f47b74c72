#!/usr/bin/env python
"""Diagnostic (GPU box): LSKNet-MoE CUDA backbone vs the live CPU oracle for config variants; prints the worst gradient
error per top-level module group, in forward order.  usage: python tools/diag_lsk.py [size]"""
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.cases import lsk_injections, upstream_grads          # noqa: E402
from oracle.lsk_moe_oracle import LskConfig, lsk_backbone_forward, lsk_param_shapes   # noqa: E402
from sm3det_b200 import LSKNet_moe_MultiInput                               # noqa: E402
from sm3det_b200.synth import make_images, make_state_dict                  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_lsk_gpu import inject                                             # noqa: E402


def rel(a, b):
    return float((a.detach().float().cpu() - b.detach().float().cpu()).abs().max() / (b.detach().abs().max() + 1e-30))


def run(tag, size, mode, kw, n=1):
    cfg = LskConfig(**kw)
    sd = make_state_dict(lsk_param_shapes(cfg), 0, True)
    net = LSKNet_moe_MultiInput(norm_cfg=dict(type='SyncBN', requires_grad=True), **kw)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    x = make_images(n, size, size, seed=1234)
    gold = dict(img=(n, size, size), mode=mode)
    noise, drops = lsk_injections(cfg, gold)
    inject(net, cfg, noise, drops)
    rec = []
    res = net(x.cuda(), record=rec)
    has_loss = isinstance(res, tuple) and len(res) == 2 and isinstance(res[0], tuple)
    outs, loss = res if has_loss else (res, None)
    forced = [r['top_idx'].cpu().long() for r in rec] or None
    skip = ('running_', 'num_batches', '.mean', '.std')
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not any(t in k for t in skip) else v) for k, v in sd.items()}
    resc = lsk_backbone_forward(sdo, cfg, x, train=True, noise=noise, drop_masks=drops, bn_state={}, forced_idx=forced)
    oc, lc = resc if has_loss else (resc, None)
    ups = upstream_grads([o.detach().cpu() for o in outs])
    (sum((o * g.cuda()).sum() for o, g in zip(outs, ups)) + (loss if has_loss else 0.0)).backward()
    (sum((o * g).sum() for o, g in zip(oc, ups)) + (lc if has_loss else 0.0)).backward()
    groups = OrderedDict()
    for name, p in net.named_parameters():
        want = sdo[name].grad
        if want is None:
            want = torch.zeros_like(sdo[name])
        if float(want.abs().max()) < 1e-7:
            continue
        parts = name.split('.')
        key = '.'.join(parts[:2]) if parts[0].startswith('block') else parts[0]
        e = rel(p.grad, want)
        if e > groups.get(key, (0.0, ''))[0]:
            groups[key] = (e, name)
    print(f'== {tag} n={n} size {size} {mode}: fwd {[f"{rel(a, b):.1e}" for a, b in zip(outs, oc)]}')
    print('   ' + '  '.join(f'{k}:{v[0]:.1e}' for k, v in groups.items()))


if __name__ == '__main__':
    M = dict(embed_dims=[64, 64, 128, 128], depths=[1, 1, 2, 1], mlp_ratios=[4, 4, 2, 2])
    for n, size in ((2, 64), (1, 64), (2, 128), (1, 128), (2, 256), (1, 96)):
        run('mini dense', size, 'train', dict(M), n=n)
