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


def run_hooks(size, n):
    """gradient at every stage output (loss term + what the next stage sends back), CUDA vs oracle"""
    kw = dict(embed_dims=[64, 64, 128, 128], depths=[1, 1, 2, 1], mlp_ratios=[4, 4, 2, 2])
    cfg = LskConfig(**kw)
    sd = make_state_dict(lsk_param_shapes(cfg), 0, True)
    net = LSKNet_moe_MultiInput(norm_cfg=dict(type='SyncBN', requires_grad=True), **kw)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    x = make_images(n, size, size, seed=1234)
    xg = x.cuda().requires_grad_(True)
    outs = net(xg)
    for o in outs:
        o.retain_grad()
    skip = ('running_', 'num_batches', '.mean', '.std')
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not any(t in k for t in skip) else v) for k, v in sd.items()}
    xc = x.clone().requires_grad_(True)
    oc = lsk_backbone_forward(sdo, cfg, xc, train=True, bn_state={})
    for o in oc:
        o.retain_grad()
    ups = upstream_grads([o.detach().cpu() for o in outs])
    sum((o * g.cuda()).sum() for o, g in zip(outs, ups)).backward()
    sum((o * g).sum() for o, g in zip(oc, ups)).backward()
    print(f'== hooks n={n} size {size}: d(out_i) ' + '  '.join(f'{rel(a.grad, b.grad):.1e}' for a, b in zip(outs, oc)) + f'  d(image) {rel(xg.grad, xc.grad):.1e}')
    # contribution of the next stage alone: grad - upstream
    print('   next-stage part: ' + '  '.join(f'{rel(a.grad.cpu() - g, b.grad - g):.1e}' for a, b, g in zip(outs, oc, ups)))
    for o, b, g in zip(outs[:3], oc[:3], ups[:3]):
        d_gpu, d_ref = (o.grad.cpu() - g)[0], (b.grad - g)[0]
        print('   sample GPU', d_gpu.flatten()[:4].tolist(), 'ref', d_ref.flatten()[:4].tolist(), 'ratio of norms', float(d_gpu.norm() / d_ref.norm()))


def run_internal(size, n):
    """gradients at every Block output / patch-embed output, CUDA (NHWC) vs oracle (NCHW), in forward order"""
    import oracle.lsk_moe_oracle as O
    from sm3det_b200 import lsk_backbone as LB
    kw = dict(embed_dims=[64, 64, 128, 128], depths=[1, 1, 2, 1], mlp_ratios=[4, 4, 2, 2])
    cfg = LskConfig(**kw)
    sd = make_state_dict(lsk_param_shapes(cfg), 0, True)
    net = LSKNet_moe_MultiInput(norm_cfg=dict(type='SyncBN', requires_grad=True), **kw)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    got = []

    def keep(name):
        def hook(mod, inp, out):
            t = out[0] if isinstance(out, tuple) else out
            if torch.is_tensor(t) and t.requires_grad:
                t.retain_grad()
                got.append((name, t))
        return hook
    for name, m in net.named_modules():
        if isinstance(m, (LB.Block, LB.OverlapPatchEmbed)):
            m.register_forward_hook(keep(name))
    x = make_images(n, size, size, seed=1234)
    outs = net(x.cuda())
    skip = ('running_', 'num_batches', '.mean', '.std')
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not any(t in k for t in skip) else v) for k, v in sd.items()}
    ref = []
    o_block, o_bn = O.block, O.batch_norm

    def block(xx, sd_, p, *a, **k):
        r = o_block(xx, sd_, p, *a, **k)
        r[0].retain_grad()
        ref.append((p.rstrip('.'), r[0]))
        return r

    def bn(xx, sd_, p, *a, **k):
        r = o_bn(xx, sd_, p, *a, **k)
        if p.startswith('patch_embed') and p != 'patch_embed1.':
            r.retain_grad()
            ref.append((p.split('.')[0], r))
        return r
    O.block, O.batch_norm = block, bn
    try:
        oc = lsk_backbone_forward(sdo, cfg, x, train=True, bn_state={})
    finally:
        O.block, O.batch_norm = o_block, o_bn
    ups = upstream_grads([o.detach().cpu() for o in outs])
    sum((o * g.cuda()).sum() for o, g in zip(outs, ups)).backward()
    sum((o * g).sum() for o, g in zip(oc, ups)).backward()
    refd = dict(ref)
    print(f'== internal n={n} size {size}')
    for name, t in got:
        if name in refd and refd[name].grad is not None and t.grad is not None:
            r = refd[name]
            print(f'   {name:16s} fwd {rel(t.permute(0, 3, 1, 2), r):.1e}  grad {rel(t.grad.permute(0, 3, 1, 2), r.grad):.1e}  strides {tuple(t.grad.stride())} shape {tuple(t.grad.shape)}')


def run_selfcheck(size, n):
    """No oracle: take the CUDA net's own tensors around the stage-3 -> stage-4 boundary and recompute each backward step
    with torch on the CPU."""
    import torch.nn.functional as F
    from sm3det_b200 import lsk_backbone as LB
    from sm3det_b200 import lsk_functional as LF
    kw = dict(embed_dims=[64, 64, 128, 128], depths=[1, 1, 2, 1], mlp_ratios=[4, 4, 2, 2])
    cfg = LskConfig(**kw)
    sd = make_state_dict(lsk_param_shapes(cfg), 0, True)
    net = LSKNet_moe_MultiInput(norm_cfg=dict(type='SyncBN', requires_grad=True), **kw)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    cap = {}
    pe = net.patch_embed4
    rm0, rv0 = pe.norm.running_mean.clone(), pe.norm.running_var.clone()
    o_apply = LF.PatchEmbedFn.apply

    def pe_forward(x, nchw):
        cap['y3'] = x
        conv = o_apply(x, pe.proj.weight, pe.proj.bias, pe.proj.stride[0], nchw)
        conv.retain_grad()
        cap['conv'] = conv
        out = LB._bn(pe.norm, conv)
        out.retain_grad()
        cap['bn'] = out
        return out
    pe.forward = pe_forward
    x = make_images(n, size, size, seed=1234)
    outs = net(x.cuda())
    for o in outs:
        o.retain_grad()
    ups = upstream_grads([o.detach().cpu() for o in outs])
    sum((o * g.cuda()).sum() for o, g in zip(outs, ups)).backward()
    y3, conv, bn = cap['y3'], cap['conv'], cap['bn']
    print(f'== selfcheck n={n} size {size}: y3 is outs[2]: {y3 is outs[2]}  shapes y3 {tuple(y3.shape)} conv {tuple(conv.shape)}')
    # (1) BatchNorm backward from the CUDA d(bn)
    c_cpu = conv.detach().cpu().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    w_bn, b_bn = pe.norm.weight.detach().cpu().requires_grad_(True), pe.norm.bias.detach().cpu().requires_grad_(True)
    z = F.batch_norm(c_cpu, rm0.cpu().clone(), rv0.cpu().clone(), w_bn, b_bn, True, 0.1, pe.norm.eps)
    z.backward(bn.grad.cpu().permute(0, 3, 1, 2).contiguous())
    print(f'   BN fwd {rel(bn.permute(0, 3, 1, 2), z):.1e}  BN bwd d(conv) {rel(conv.grad.permute(0, 3, 1, 2), c_cpu.grad):.1e}  dgamma {rel(pe.norm.weight.grad, w_bn.grad):.1e}')
    # (2) conv backward from the CUDA d(conv)
    y_cpu = y3.detach().cpu().requires_grad_(True)
    w_c = pe.proj.weight.detach().cpu().requires_grad_(True)
    F.conv2d(y_cpu, w_c, pe.proj.bias.detach().cpu(), stride=2, padding=1).backward(conv.grad.cpu().permute(0, 3, 1, 2).contiguous())
    dx_gpu = y3.grad.cpu() - ups[2]
    print(f'   conv bwd: dx {rel(dx_gpu, y_cpu.grad):.1e}  dw {rel(pe.proj.weight.grad, w_c.grad):.1e}   (y3.grad strides {tuple(y3.grad.stride())})')


def run_both(size, n):
    import torch.nn.functional as F
    import oracle.lsk_moe_oracle as O
    from sm3det_b200 import lsk_backbone as LB
    from sm3det_b200 import lsk_functional as LF
    kw = dict(embed_dims=[64, 64, 128, 128], depths=[1, 1, 2, 1], mlp_ratios=[4, 4, 2, 2])
    cfg = LskConfig(**kw)
    sd = make_state_dict(lsk_param_shapes(cfg), 0, True)
    net = LSKNet_moe_MultiInput(norm_cfg=dict(type='SyncBN', requires_grad=True), **kw)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    cap = {}
    pe = net.patch_embed4
    o_apply = LF.PatchEmbedFn.apply

    def pe_forward(x, nchw):
        conv = o_apply(x, pe.proj.weight, pe.proj.bias, pe.proj.stride[0], nchw)
        conv.retain_grad()
        cap['conv'] = conv
        out = LB._bn(pe.norm, conv)
        out.retain_grad()
        cap['bn'] = out
        return out
    pe.forward = pe_forward
    x = make_images(n, size, size, seed=1234)
    outs = net(x.cuda())
    for o in outs:
        o.retain_grad()
    skip = ('running_', 'num_batches', '.mean', '.std')
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not any(t in k for t in skip) else v) for k, v in sd.items()}
    ocap = {}
    o_bn = O.batch_norm

    def bn(xx, sd_, p, *a, **k):
        r = o_bn(xx, sd_, p, *a, **k)
        if p == 'patch_embed4.norm.':
            xx.retain_grad(); r.retain_grad(); ocap['conv'] = xx; ocap['bn'] = r
        return r
    O.batch_norm = bn
    try:
        oc = lsk_backbone_forward(sdo, cfg, x, train=True, bn_state={})
    finally:
        O.batch_norm = o_bn
    for o in oc:
        o.retain_grad()
    ups = upstream_grads([o.detach().cpu() for o in outs])
    sum((o * g.cuda()).sum() for o, g in zip(outs, ups)).backward()
    sum((o * g).sum() for o, g in zip(oc, ups)).backward()
    P = lambda t: t.permute(0, 3, 1, 2)
    print(f'== both n={n} size {size}: conv4 fwd {rel(P(cap["conv"]), ocap["conv"]):.1e}  bn4 fwd {rel(P(cap["bn"]), ocap["bn"]):.1e}  '
          f'd(bn4) {rel(P(cap["bn"].grad), ocap["bn"].grad):.1e}  d(conv4) {rel(P(cap["conv"].grad), ocap["conv"].grad):.1e}  '
          f'd(out3) {rel(outs[2].grad, oc[2].grad):.1e}  out3 fwd {rel(outs[2], oc[2]):.1e}')
    gb, ob = P(cap['bn'].grad).cpu(), ocap['bn'].grad
    print('   d(bn4) first values GPU', gb.flatten()[:4].tolist(), 'oracle', ob.flatten()[:4].tolist())
    gc, oc_ = P(cap['conv'].grad).cpu(), ocap['conv'].grad
    print('   d(conv4) first values GPU', gc.flatten()[:4].tolist(), 'oracle', oc_.flatten()[:4].tolist())


if __name__ == '__main__':
    run_both(64, 2)
    run_both(64, 1)
    run_both(256, 1)
