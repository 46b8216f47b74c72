#!/usr/bin/env python
"""Diagnostic (GPU box): LSKNet-MoE CUDA backbone vs the live CPU oracle for config variants; prints the worst gradient
error per top-level module group, in forward order.  usage: python tests/diag/diag_lsk.py [size]"""
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.cases import lsk_injections, upstream_grads          # noqa: E402
from oracle.lsk_moe_oracle import LskConfig, lsk_backbone_forward, lsk_param_shapes   # noqa: E402
from sm3det_b200 import LSKNet_moe_MultiInput                               # noqa: E402
from sm3det_b200.synth import make_images, make_state_dict                  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests'))
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


def run_params(size, n, mode='train_noisy', dev='cuda'):
    """Config 5 (real LSKNet-S widths): product vs the oracle run ON THE GPU in float64 and float32 (teacher-forced routing).
    Prints every parameter whose gradient differs by more than 1e-3, in forward order, and oracle-fp32 vs oracle-fp64 beside
    it (the conditioning of that gradient)."""
    from oracle.cases import LSK_S_KW
    kw = dict(LSK_S_KW, drop_rate=0.1)
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
    outs, loss = res
    forced = [r['top_idx'].long() for r in rec]
    ups = upstream_grads([o.detach().cpu() for o in outs])
    (sum((o * g.cuda()).sum() for o, g in zip(outs, ups)) + loss).backward()
    skip = ('running_', 'num_batches', '.mean', '.std')

    def oracle(dt):
        sdo = {}
        for k, v in sd.items():
            v = v.to(dev)
            if v.is_floating_point():
                v = v.to(dt)
                if not any(t in k for t in skip):
                    v = v.clone().requires_grad_(True)
            sdo[k] = v
        r = lsk_backbone_forward(sdo, cfg, x.to(dev).to(dt), train=True, noise=[t.to(dev).to(dt) for t in noise] if noise else None,
                                 drop_masks=[t.to(dev).to(dt) for t in drops] if drops else None, bn_state={}, forced_idx=[f.to(dev) for f in forced])
        oc, lc = r
        (sum((o * g.to(dev).to(dt)).sum() for o, g in zip(oc, ups)) + lc).backward()
        return sdo, oc
    s64, o64 = oracle(torch.float64)
    s32, o32 = oracle(torch.float32)
    print(f'== params n={n} size {size} {mode}: fwd product-vs-f64 {[f"{rel(a, b):.1e}" for a, b in zip(outs, o64)]}  oracle32-vs-f64 {[f"{rel(a, b):.1e}" for a, b in zip(o32, o64)]}')
    for name, p in net.named_parameters():
        w = s64[name].grad
        if w is None or float(w.abs().max()) < 1e-9:
            continue
        e, e32 = rel(p.grad, w), rel(s32[name].grad, w)
        if e > 1e-3 or e32 > 1e-3:
            print(f'   {name:58s} product {e:.1e}   oracle-fp32 {e32:.1e}   |g|max {float(w.abs().max()):.2e}')


def run_sgu(size, n, mode='train_noisy'):
    """Config 5: gradients of every tensor inside each LSK spatial gating unit, product vs oracle (fp64 on the GPU)."""
    import torch.nn.functional as F
    import oracle.lsk_moe_oracle as O
    from oracle.cases import LSK_S_KW
    from sm3det_b200 import lsk_backbone as LB
    from sm3det_b200 import lsk_functional as LF
    kw = dict(LSK_S_KW, drop_rate=0.1)
    cfg = LskConfig(**kw)
    sd = make_state_dict(lsk_param_shapes(cfg), 0, True)
    net = LSKNet_moe_MultiInput(norm_cfg=dict(type='SyncBN', requires_grad=True), **kw)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    got = []

    def p_forward(self, x):
        t = {}
        t['x'] = x
        t['attn1'] = LF.DWConvFn.apply(x, self.conv0.weight, self.conv0.bias, 5, 1)
        t['attn2'] = LF.DWConvFn.apply(t['attn1'], self.conv_spatial.weight, self.conv_spatial.bias, 7, 3)
        t['a1'] = LB._conv1x1(self.conv1, t['attn1'])
        t['a2'] = LB._conv1x1(self.conv2, t['attn2'])
        t['sel'] = LF.LSKSelectFn.apply(t['a1'], t['a2'], self.conv_squeeze.weight, self.conv_squeeze.bias)
        t['attn'] = LB._conv1x1(self.conv, t['sel'])
        t['out'] = LF.MulFn.apply(x, t['attn'])
        for v in t.values():
            if v.requires_grad:
                v.retain_grad()
        got.append(t)
        return t['out']
    o_fwd = LB.LSKblock.forward
    LB.LSKblock.forward = p_forward
    x = make_images(n, size, size, seed=1234)
    gold = dict(img=(n, size, size), mode=mode)
    noise, drops = lsk_injections(cfg, gold)
    inject(net, cfg, noise, drops)
    rec = []
    try:
        outs, loss = net(x.cuda(), record=rec)
    finally:
        LB.LSKblock.forward = o_fwd
    forced = [r['top_idx'].long() for r in rec]
    ups = upstream_grads([o.detach().cpu() for o in outs])
    (sum((o * g.cuda()).sum() for o, g in zip(outs, ups)) + loss).backward()
    ref = []

    def o_lsk_block(x, sd, p):
        c = x.shape[1]
        t = {'p': p, 'x': x}
        t['attn1'] = F.conv2d(x, sd[p + 'conv0.weight'], sd[p + 'conv0.bias'], padding=2, groups=c)
        t['attn2'] = F.conv2d(t['attn1'], sd[p + 'conv_spatial.weight'], sd[p + 'conv_spatial.bias'], padding=9, groups=c, dilation=3)
        t['a1'] = F.conv2d(t['attn1'], sd[p + 'conv1.weight'], sd[p + 'conv1.bias'])
        t['a2'] = F.conv2d(t['attn2'], sd[p + 'conv2.weight'], sd[p + 'conv2.bias'])
        attn = torch.cat([t['a1'], t['a2']], dim=1)
        agg = torch.cat([attn.mean(1, keepdim=True), attn.max(1, keepdim=True)[0]], 1)
        top2 = attn.detach().topk(2, dim=1).values
        t['gap'] = ((top2[:, 0] - top2[:, 1]) / attn.detach().abs().amax(1).clamp_min(1e-20)).flatten()
        sig = F.conv2d(agg, sd[p + 'conv_squeeze.weight'], sd[p + 'conv_squeeze.bias'], padding=3).sigmoid()
        t['sel'] = t['a1'] * sig[:, 0:1] + t['a2'] * sig[:, 1:2]
        t['attn'] = F.conv2d(t['sel'], sd[p + 'conv.weight'], sd[p + 'conv.bias'])
        t['out'] = x * t['attn']
        for k, v in t.items():
            if torch.is_tensor(v) and v.requires_grad:
                v.retain_grad()
        ref.append(t)
        return t['out']
    skip = ('running_', 'num_batches', '.mean', '.std')
    dt, dev = torch.float64, 'cuda'
    sdo = {}
    for k, v in sd.items():
        v = v.to(dev)
        if v.is_floating_point():
            v = v.to(dt)
            if not any(t in k for t in skip):
                v = v.clone().requires_grad_(True)
        sdo[k] = v
    o_orig = O.lsk_block
    O.lsk_block = o_lsk_block
    try:
        oc, lc = lsk_backbone_forward(sdo, cfg, x.to(dev).to(dt), train=True, noise=[t.to(dev).to(dt) for t in noise] if noise else None,
                                      drop_masks=[t.to(dev).to(dt) for t in drops] if drops else None, bn_state={}, forced_idx=[f.to(dev) for f in forced])
    finally:
        O.lsk_block = o_orig
    (sum((o * g.to(dev).to(dt)).sum() for o, g in zip(oc, ups)) + lc).backward()
    print(f'== sgu n={n} size {size}')
    P = lambda t: t.permute(0, 3, 1, 2)
    for g, r in zip(got, ref):
        gap = r['gap']
        line = f"   {r['p']:42s} shape {tuple(r['a1'].shape)}  near-ties(<1e-4) {int((gap < 1e-4).sum())}/{gap.numel()}"
        for k in ('out', 'attn', 'sel', 'a1', 'a2', 'attn2', 'attn1', 'x'):
            if g[k].grad is None or r[k].grad is None:
                line += f'  d{k} -'
                continue
            line += f'  d{k} {rel(P(g[k].grad), r[k].grad):.1e}'
        line += '  | fwd a1 %.1e sel %.1e' % (rel(P(g['a1']), r['a1']), rel(P(g['sel']), r['sel']))
        print(line)


if __name__ == '__main__':
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for a in sys.argv[1:] or ['768']:
        run_sgu(int(a), 2)
