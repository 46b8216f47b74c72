"""End-to-end parity of the CUDA backbone against the CPU oracle (and the reference-generated goldens)."""
import glob
import os

import pytest
import torch

from oracle.cases import CASES, make_noise, upstream_grads
from oracle.convnext_moe_oracle import OracleConfig, backbone_forward, param_shapes
from oracle.gen_golden import moe_token_counts
from sm3det_b200.synth import make_images, make_state_dict

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
TOL = 1e-3   # north-star tolerance: max-norm relative, fp32


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def build(kw, weights='trained', seed=0):
    from sm3det_b200 import ConvNeXt_moe_MultiInput
    cfg = OracleConfig(**kw)
    sd = make_state_dict(param_shapes(cfg), seed, weights == 'trained')
    net = ConvNeXt_moe_MultiInput(**kw)
    missing = net.load_state_dict(sd, strict=True)
    return cfg, sd, net.cuda()


def routing_flips(rec_gpu, rec_cpu):
    """number of tokens whose top-k set differs, and whether every flip is a numerical near-tie"""
    flips, ok = 0, True
    for g, c in zip(rec_gpu, rec_cpu):
        gi = g['top_idx'].cpu().long().sort(dim=1).values
        ci = c['top_idx'].sort(dim=1).values
        m = (gi != ci).any(dim=1)
        flips += int(m.sum())
        if m.any():
            lg = c['logits'][m]
            top = lg.topk(min(gi.shape[1] + 1, lg.shape[1]), dim=-1).values
            gap = (top[:, :-1] - top[:, 1:]).min(dim=1).values
            ok = ok and bool((gap < 2e-3 * lg.abs().max()).all())
    return flips, ok


@pytest.mark.parametrize('path', sorted(p for p in glob.glob(os.path.join(GOLD, '*.pt')) if not os.path.basename(p).startswith(('lsk_', 'van_'))), ids=lambda p: os.path.basename(p)[:-3])
def test_forward_matches_reference_golden(path):
    gold = torch.load(path, weights_only=False)
    if gold['mode'] == 'train_noisy':
        pytest.skip('noisy forward covered by test_train_noisy_forward')
    cfg, sd, net = build(gold['kw'], gold['weights'])
    n, h, w = gold['img']
    x = make_images(n, h, w, seed=1234).cuda()
    net.train(gold['mode'] != 'eval')
    rec = []
    with torch.no_grad():
        res = net(x, record=rec)
    has_loss = 'gate_loss' in gold
    outs, loss = res if has_loss else (res, None)
    st = gold['stride']
    flips = 0
    for r, g in zip(rec, gold['moe']):
        flips += int((r['top_idx'].cpu().long().sort(1).values != g['top_idx'].long().sort(1).values).any(1).sum())
    errs = [rel(o[:, :, ::st, ::st], g) for o, g in zip(outs, gold['outs'])]
    print(os.path.basename(path), 'rel errs', errs, 'flips', flips, 'loss', None if loss is None else (loss.item(), gold['gate_loss'].item()))
    if flips == 0:
        assert max(errs) < TOL
        if has_loss:
            assert abs(loss.item() - gold['gate_loss'].item()) <= 1e-4 * abs(gold['gate_loss'].item()) + 1e-8
    else:
        assert flips <= 2, 'more routing flips than fp32 near-ties can explain'


@pytest.mark.parametrize('name', ['mini_moe_e4k2_eval', 'mini_moe_e8k3_eval', 'mini_moe_e6k1_eval', 'mini_moe_e2k2_eval'])
def test_moe_layers_match_oracle(name):
    spec = CASES[name]
    cfg, sd, net = build(spec['kw'])
    n, h, w = spec['img']
    x = make_images(n, h, w, seed=77)
    net.eval()
    rec_g, rec_c = [], []
    with torch.no_grad():
        outs_g, loss_g = net(x.cuda(), record=rec_g)
        outs_c, loss_c = backbone_forward(sd, cfg, x, record=rec_c)
    flips, near_tie = routing_flips(rec_g, rec_c)
    assert near_tie, 'routing differs on a token that is not a near-tie'
    print(name, 'flips', flips, [rel(a, b) for a, b in zip(outs_g, outs_c)], loss_g.item(), loss_c.item())
    if flips == 0:
        for a, b in zip(outs_g, outs_c):
            assert rel(a, b) < TOL
        assert abs(loss_g.item() - loss_c.item()) <= 1e-4 * abs(loss_c.item()) + 1e-8
        # pre-gamma MoE outputs (what layer scale would otherwise hide)
        for g, c in zip(rec_g, rec_c):
            assert rel(g['y'], c['y']) < TOL
            assert rel(g['importance'], c['importance']) < 1e-4


def test_list_input_and_plain_class():
    from sm3det_b200 import ConvNeXt_moe
    kw = dict(arch=dict(depths=[1, 1, 2, 1], channels=[32, 64, 96, 128]), MoE_Block_inds=[[], [], [1], []], num_experts=4, top_k=2)
    cfg, sd, net = build(kw)
    xs = [make_images(1, 64, 64, seed=s) for s in (1, 2, 3)]
    net.eval()
    with torch.no_grad():
        a = net([t.cuda() for t in xs], ['sar', 'rgb', 'ifr'])
        b = net(torch.cat(xs).cuda())
    assert isinstance(a, tuple) and len(a) == 2 and len(a[0]) == 4
    for p, q in zip(a[0], b[0]):
        assert torch.equal(p, q) and p.is_contiguous()
    assert a[0][0].shape == (3, 32, 16, 16) and a[0][3].shape == (3, 128, 2, 2)
    # plain class: tuple only when dense, (tuple, loss) with MoE
    cfgp = OracleConfig(multi_input=False, **kw)
    sdp = make_state_dict(param_shapes(cfgp), 1, True)
    netp = ConvNeXt_moe(**kw)
    netp.load_state_dict(sdp, strict=True)
    netp = netp.cuda().eval()
    x = make_images(2, 64, 64, seed=5)
    with torch.no_grad():
        og, lg = netp(x.cuda())
        oc, lc = backbone_forward(sdp, cfgp, x)
    for p, q in zip(og, oc):
        assert rel(p, q) < TOL
    dense = ConvNeXt_moe(arch=kw['arch']).cuda().eval()
    with torch.no_grad():
        r = dense(x.cuda())
    assert isinstance(r, tuple) and len(r) == 4 and torch.is_tensor(r[0])


@pytest.mark.parametrize('name', ['mini_dense', 'mini_moe_e4k2_train_clean', 'mini2_moe_e8k2_train_clean',
                                  'mini_moe_e4k2_train_noisy', 'mini_moe_e8k3_noisy'])
def test_backward_matches_oracle(name):
    if name == 'mini_moe_e8k3_noisy':
        spec = dict(CASES['mini_moe_e8k3_eval'], mode='train_noisy')
    else:
        spec = CASES[name]
    kw = dict(spec['kw'])
    noisy = spec['mode'] == 'train_noisy'
    if not noisy:
        kw['noisy_gating'] = False
    cfg, sd, net = build(kw)
    n, h, w = spec['img']
    x = make_images(n, h, w, seed=1234)
    net.train()
    noise = None
    if noisy:
        noise = make_noise(cfg, moe_token_counts(cfg, n, h, w))
        for m, nz in zip([m for m in net.modules() if m.__class__.__name__ == 'MoE_layer'], noise):
            m._injected_noise = nz
    rec_g, rec_c = [], []
    res_g = net(x.cuda(), record=rec_g)
    sdg = {k: (v.clone().requires_grad_(True) if 'ffn.mean' not in k and 'ffn.std' not in k else v) for k, v in sd.items()}
    res_c = backbone_forward(sdg, cfg, x, train=True, noise=noise, record=rec_c)
    has_loss = isinstance(res_c, tuple) and len(res_c) == 2 and isinstance(res_c[0], tuple)
    og, lg = res_g if has_loss else (res_g, None)
    oc, lc = res_c if has_loss else (res_c, None)
    flips, near_tie = routing_flips(rec_g, rec_c)
    assert near_tie
    ups = upstream_grads(oc)
    (sum((o * u.cuda()).sum() for o, u in zip(og, ups)) + (lg if has_loss else 0.0)).backward()
    (sum((o * u).sum() for o, u in zip(oc, ups)) + (lc if has_loss else 0.0)).backward()
    worst = {}
    for pname, p in net.named_parameters():
        ref = sdg[pname].grad
        if ref is None:
            ref = torch.zeros_like(sdg[pname])
        assert p.grad is not None, f'{pname}: every parameter must receive a (possibly zero) gradient (DDP)'
        if 'w_noise' in pname and not noisy:
            continue
        e = (p.grad.cpu() - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
        worst[pname] = e
    bad = {k: v for k, v in worst.items() if v > 2e-3}
    top = sorted(worst.items(), key=lambda kv: -kv[1])[:8]
    print(name, 'flips', flips, 'fwd', [rel(a, b) for a, b in zip(og, oc)], 'worst grads', top)
    if flips == 0:
        assert not bad, bad


def test_train_noisy_forward():
    spec = CASES['mini_moe_e4k2_train_noisy']
    gold = torch.load(os.path.join(GOLD, 'mini_moe_e4k2_train_noisy.pt'), weights_only=False)
    cfg, sd, net = build(spec['kw'])
    n, h, w = spec['img']
    x = make_images(n, h, w, seed=1234)
    noise = make_noise(cfg, moe_token_counts(cfg, n, h, w))
    moe_layers = [m for m in net.modules() if m.__class__.__name__ == 'MoE_layer']
    for m, nz in zip(moe_layers, noise):
        m._injected_noise = nz
    net.train()
    rec = []
    with torch.no_grad():
        outs, loss = net(x.cuda(), record=rec)
    flips = sum(int((r['top_idx'].cpu().long().sort(1).values != g['top_idx'].long().sort(1).values).any(1).sum())
                for r, g in zip(rec, gold['moe']))
    errs = [rel(o, g) for o, g in zip(outs, gold['outs'])]
    print('noisy fwd', errs, flips, loss.item(), gold['gate_loss'].item())
    if flips == 0:
        assert max(errs) < TOL
        assert abs(loss.item() - gold['gate_loss'].item()) <= 1e-4 * abs(gold['gate_loss'].item())
        for r, g in zip(rec, gold['moe']):
            assert rel(r['load'], g['load']) < 1e-4


def test_mixed_precision_mode():
    """AMP recipe: under torch.autocast the GEMMs run single-pass bf16; the result stays within bf16
    accuracy of the fp32 oracle, differs from the fp32-accurate mode, and the next plain forward is fp32-accurate again."""
    spec = CASES['mini_moe_e4k2_train_clean']
    cfg, sd, net = build(spec['kw'])
    net.train()
    n, h, w = spec['img']
    x = make_images(n, h, w, seed=1234).cuda()
    with torch.no_grad():
        ref, _ = backbone_forward(sd, cfg, x.cpu(), train=True)
    from sm3det_b200 import ops
    try:
        o32, l32 = net(x)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            o16, l16 = net(x)
        (sum(o.float().square().mean() for o in o16) + l16).backward()
    finally:
        ops.set_gemm_precision('fp32')            # never leak the bf16 mode into the other tests of this process
    assert all(p.grad is None or torch.isfinite(p.grad).all() for p in net.parameters())
    e32 = max(rel(a, b) for a, b in zip(o32, ref))
    l2 = lambda a, b: ((a.detach().float().cpu() - b).norm() / b.norm()).item()     # a routing flip moves single tokens a lot
    e16 = max(l2(a, b) for a, b in zip(o16, ref))
    print('fp32-mode err', e32, 'bf16-mode rel-L2 err', e16)
    assert e32 < TOL
    assert 1e-4 < e16 < 5e-2
    o32b, _ = net(x)
    assert max(rel(a, b) for a, b in zip(o32b, ref)) < TOL
