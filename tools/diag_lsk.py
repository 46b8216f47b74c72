#!/usr/bin/env python
"""Diagnostic (GPU box): LSKNet-S (config 5 widths) CUDA backbone vs the live CPU oracle at several image sizes / modes,
printing every parameter whose gradient is off.  usage: python tools/diag_lsk.py [sizes...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.cases import LSK_S_KW, lsk_injections, upstream_grads          # noqa: E402
from oracle.lsk_moe_oracle import LskConfig, lsk_backbone_forward, lsk_param_shapes   # noqa: E402
from sm3det_b200 import LSKNet_moe_MultiInput                               # noqa: E402
from sm3det_b200.synth import make_images, make_state_dict                  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_lsk_gpu import inject                                             # noqa: E402


def rel(a, b):
    return float((a.detach().float().cpu() - b.detach().float().cpu()).abs().max() / (b.detach().abs().max() + 1e-30))


def run(size, mode, kw):
    cfg = LskConfig(**kw)
    sd = make_state_dict(lsk_param_shapes(cfg), 0, True)
    net = LSKNet_moe_MultiInput(norm_cfg=dict(type='SyncBN', requires_grad=True), **kw)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    x = make_images(1, size, size, seed=1234)
    gold = dict(img=(1, size, size), mode=mode)
    noise, drops = lsk_injections(cfg, gold)
    inject(net, cfg, noise, drops)
    rec = []
    outs, loss = net(x.cuda(), record=rec)
    forced = [r['top_idx'].cpu().long() for r in rec]
    skip = ('running_', 'num_batches', '.mean', '.std')
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not any(t in k for t in skip) else v) for k, v in sd.items()}
    oc, lc = lsk_backbone_forward(sdo, cfg, x, train=True, noise=noise, drop_masks=drops, bn_state={}, forced_idx=forced)
    ups = upstream_grads([o.detach().cpu() for o in outs])
    (sum((o * g.cuda()).sum() for o, g in zip(outs, ups)) + loss).backward()
    (sum((o * g).sum() for o, g in zip(oc, ups)) + lc).backward()
    print(f'== size {size} mode {mode} drop {kw.get("drop_rate", 0)}: fwd {[f"{rel(a, b):.1e}" for a, b in zip(outs, oc)]} loss {loss.item():.6f} / {lc.item():.6f}')
    bad = 0
    for name, p in net.named_parameters():
        want = sdo[name].grad
        if want is None:
            want = torch.zeros_like(sdo[name])
        e = rel(p.grad, want)
        if e > 3e-3 and float(want.abs().max()) > 1e-7:
            bad += 1
            if bad <= 25:
                print(f'   {name:60s} err {e:.2e}  |ref|max {float(want.abs().max()):.2e}')
    print(f'   parameters off: {bad}')


if __name__ == '__main__':
    sizes = [int(a) for a in sys.argv[1:]] or [256, 1024]
    for s in sizes:
        run(s, 'train', dict(LSK_S_KW, noisy_gating=False))
        run(s, 'train_noisy', dict(LSK_S_KW, drop_rate=0.1))
