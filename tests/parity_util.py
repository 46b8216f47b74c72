"""Shared machinery of the end-to-end parity tests (ConvNeXt-MoE family).

How a comparison is made rigorous in the presence of routing flips
------------------------------------------------------------------
The router's top-k is bit-exact on identical inputs (tests/test_ops_gpu.py), but end to end the router sees LayerNorm
outputs that differ from the CPU oracle's in the last bits, so on a token whose (k)-th and (k+1)-th logits are a numerical
tie the two sides may pick different experts -- and that token (plus everything a 7x7 depthwise conv spreads it to) then
legitimately differs by O(1).  Round 1 skipped the numeric assertions whenever that happened.  Now:

  1. the oracle is run *teacher-forced* to the CUDA path's routing (``forced_idx``, a test-only hook): at every MoE layer it
     still computes its OWN logits and top-k from inputs that already contain the CUDA path's upstream decisions, so a
     difference at layer L is that layer's own numerics, not the echo of an upstream flip (with 36 MoE layers in a row --
     config 4 -- one legitimate tie flip otherwise cascades into real routing changes downstream);
  2. every such per-layer flip must be a near-tie: the oracle's (k)-vs-(k+1) logit gap on that token must be below
     ``GAP_TOL x max|logit|``, and flips must be rare;
  3. outputs, pre-gamma MoE outputs, importance / load / gate loss and every parameter gradient of the forced oracle are
     compared with the CUDA path on ALL elements -- no masking, no early exit;
  4. when the routing also equals the reference's own (the fixture's indices) everywhere, the CUDA path is additionally
     compared with the reference-generated fixture values directly.
The oracle itself is pinned bit-for-bit to the unmodified reference by oracle/gen_golden.py / tests/test_oracle.py.
"""
import torch

from oracle.cases import make_noise, upstream_grads
from oracle.convnext_moe_oracle import OracleConfig, backbone_forward, param_shapes, tie_da_weights
from oracle.gen_golden import moe_token_counts
from sm3det_b200.synth import make_images, make_state_dict

TOL = 1e-3        # north-star tolerance: max-norm relative, fp32
GRAD_TOL = 2e-3   # parameter gradients (they pass through up to 36 GEMM pairs twice)
TEMP_TOL = 1e-2   # w_gate.temperature: one scalar = a sum over every token's logits with heavy cancellation
GAP_TOL = 1e-3    # a flip is a near-tie if the oracle's (k)-vs-(k+1) gap < GAP_TOL * max|logit| of that layer
MAX_FLIP_FRACTION = 2e-3


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def build(kw, weights='trained', seed=0, cls=None, da=False):
    from sm3det_b200 import ConvNeXt_DA_MultiInput, ConvNeXt_moe_MultiInput
    cfg = OracleConfig(da=da, **kw)
    sd = make_state_dict(param_shapes(cfg), seed, weights == 'trained')
    if da:
        tie_da_weights(sd)                     # the reference shares one Sequential between the three dataset gates
        cls = cls or ConvNeXt_DA_MultiInput
    net = (cls or ConvNeXt_moe_MultiInput)(**kw)
    net.load_state_dict(sd, strict=True)
    return cfg, sd, net.cuda()


def flipped_tokens(idx_gpu, idx_ref):
    """bool [T]: the top-k expert SET differs."""
    return (idx_gpu.cpu().long().sort(dim=1).values != idx_ref.cpu().long().sort(dim=1).values).any(dim=1)


def assert_flips_are_near_ties(rec_gpu, ref_layers, what=''):
    """ref_layers: per MoE layer a dict with top_idx and either (gap, logit_scale) [fixture] or logits [live oracle].
    Returns the total number of flipped tokens."""
    total = tokens = 0
    assert len(rec_gpu) == len(ref_layers), (len(rec_gpu), len(ref_layers))
    for li, (g, c) in enumerate(zip(rec_gpu, ref_layers)):
        m = flipped_tokens(g['top_idx'], c['top_idx'])
        tokens += m.numel()
        if not m.any():
            continue
        total += int(m.sum())
        if 'gap' in c:
            gap, scale = c['gap'][m].float(), c['logit_scale']
        else:
            lg = c['logits']
            k = g['top_idx'].shape[1]
            top = lg[m].topk(min(k + 1, lg.shape[1]), dim=-1).values
            gap, scale = (top[:, k - 1] - top[:, k]), float(lg.abs().max())
        worst = float(gap.max())
        assert worst < GAP_TOL * scale, (f'{what} MoE layer {li}: routing differs on a token whose (k)-vs-(k+1) logit gap is '
                                         f'{worst:.3e} (max|logit| {scale:.3e}) -- not a numerical tie')
    assert total <= max(2, MAX_FLIP_FRACTION * tokens), f'{what}: {total} routing flips in {tokens} tokens'
    return total


def run_case(kw, img, mode, weights='trained', gold=None, img_seed=1234, backward=None, check_pre_gamma=True, datasets=None):
    """Full comparison of the CUDA backbone with the (teacher-forced) oracle, and with ``gold`` when given.
    Returns a dict of the measured errors (printed by the callers)."""
    kw = dict(kw)
    da = datasets is not None              # ConvNeXt_DA_MultiInput: `datasets` names the DALayer gate per batch / per sample
    cfg, sd, net = build(kw, weights, da=da)
    n, h, w = img
    x = make_images(n, h, w, seed=img_seed)
    okw, gkw = {}, {}
    xg = x.cuda()
    if da:
        okw, gkw = dict(datasets=list(datasets)), dict(datasets=list(datasets))
        if len(datasets) > 1:
            x = [x[i:i + 1] for i in range(n)]
            xg = [t.cuda() for t in x]
    train = mode != 'eval'
    noisy = mode == 'train_noisy'
    if backward is None:
        backward = train
    net.train(train)
    noise = None
    if noisy:
        noise = make_noise(cfg, moe_token_counts(cfg, n, h, w))
        for m, nz in zip([m for m in net.modules() if m.__class__.__name__ == 'MoE_layer'], noise):
            m._injected_noise = nz
    rec_g = []
    with torch.set_grad_enabled(backward):
        res_g = net(xg, record=rec_g, **gkw)
    has_loss = isinstance(res_g, tuple) and len(res_g) == 2 and isinstance(res_g[0], tuple)
    og, lg = res_g if has_loss else (res_g, None)

    # (1) the oracle, teacher-forced to the CUDA path's routing
    forced = [r['top_idx'].cpu().long() for r in rec_g] or None
    rec_c, pre_c = [], []
    if backward:
        sdo = {k: (v.clone().requires_grad_(True) if 'ffn.mean' not in k and 'ffn.std' not in k else v) for k, v in sd.items()}
        if da:
            tie_da_weights(sdo)
    else:
        sdo = sd
    with torch.set_grad_enabled(backward):
        res_c = backbone_forward(sdo, cfg, x, train=train, noise=noise, record=rec_c, pre_gamma=pre_c, forced_idx=forced, **okw)
    oc, lc = res_c if has_loss else (res_c, None)
    # (2) per layer: the CUDA choice vs the oracle's own top-k on the same (forced-upstream) inputs -- ties only
    flips = 0
    if rec_g:
        own = [dict(top_idx=c['logits'].topk(g['top_idx'].shape[1], dim=-1).indices, logits=c['logits']) for g, c in zip(rec_g, rec_c)]
        flips = assert_flips_are_near_ties(rec_g, own, what=str(gold['name'] if gold else kw.get('arch')))
    # routing vs the reference's un-forced decisions (fixture): equal unless a tie flipped somewhere upstream
    gold_flips = 0
    if gold is not None and rec_g:
        gold_flips = sum(int(flipped_tokens(g['top_idx'], c['top_idx']).sum()) for g, c in zip(rec_g, gold['moe']))
        assert flips > 0 or gold_flips == 0, 'routing differs from the fixture although every layer agrees with the oracle'
    errs = dict(flips=flips, gold_flips=gold_flips, fwd=[rel(a, b) for a, b in zip(og, oc)])
    assert max(errs['fwd']) < TOL, errs
    if has_loss:
        errs['loss'] = (lg.item(), lc.item())
        assert abs(lg.item() - lc.item()) <= 1e-4 * abs(lc.item()) + 1e-8, errs
        for g, c in zip(rec_g, rec_c):
            assert torch.equal(g['top_idx'].cpu().long().sort(1).values, c['top_idx'].sort(1).values)
            assert rel(g['importance'], c['importance']) < 1e-4
            assert rel(g['load'], c['load']) < 1e-4
            if check_pre_gamma and g.get('y') is not None:
                assert rel(g['y'], c['y']) < TOL            # pre-gamma MoE output (what layer scale would otherwise hide)

    # (4) the reference-generated fixture itself, when routing agrees everywhere
    if gold is not None and gold_flips == 0:
        st = gold['stride']
        errs['gold_fwd'] = [rel(o[:, :, ::st, ::st], g) for o, g in zip(og, gold['outs'])]
        assert max(errs['gold_fwd']) < TOL, errs
        if has_loss:
            assert abs(lg.item() - gold['gate_loss'].item()) <= 1e-4 * abs(gold['gate_loss'].item()) + 1e-8
        if gold['moe'] and 'gap' not in gold['moe'][0]:
            for r, g in zip(rec_g, gold['moe']):
                assert rel(r['load'], g['load']) < 1e-4

    if backward:
        ups = upstream_grads(oc)
        (sum((o * u.cuda()).sum() for o, u in zip(og, ups)) + (lg if has_loss else 0.0)).backward()
        (sum((o * u).sum() for o, u in zip(oc, ups)) + (lc if has_loss else 0.0)).backward()
        worst = {}
        for pname, p in net.named_parameters():
            ref = sdo[pname].grad
            if ref is None:
                ref = torch.zeros_like(sdo[pname])
            assert p.grad is not None, f'{pname}: every parameter must receive a (possibly zero) gradient (DDP)'
            if 'w_noise' in pname and not noisy:
                continue
            worst[pname] = (p.grad.cpu() - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
        errs['worst_grads'] = sorted(worst.items(), key=lambda kv: -kv[1])[:6]
        bad = {k: v for k, v in worst.items() if v > (TEMP_TOL if k.endswith('w_gate.temperature') else GRAD_TOL)}
        assert not bad, bad
        if gold is not None and gold_flips == 0 and 'grads' in gold:
            for pname, dg in gold['grads'].items():
                got = dict(net.named_parameters())[pname].grad.detach().float().cpu().reshape(-1)
                if 'full' in dg:
                    want, have = dg['full'], got
                else:
                    want, have = dg['sample'], got[dg['idx']]
                    assert abs(got.double().norm().item() - dg['l2']) <= GRAD_TOL * dg['l2'] + 1e-12, pname
                scale = max(float(want.abs().max()), dg.get('l2', 0.0) / max(1.0, got.numel() ** 0.5), 1e-12)
                assert float((have - want).abs().max()) <= GRAD_TOL * scale * 4, pname
    return errs
