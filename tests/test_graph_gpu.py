"""sm3det_b200.graphed.GraphedStep: a captured forward+backward replays to the same loss and gradients as eager launches
(test infrastructure may use the oracle's case tables; the product path under test is the CUDA library)."""
import pytest
import torch

from oracle.cases import CASES, LSK_CASES

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.detach().float() - b.detach().float()).abs().max() / (b.detach().abs().max() + 1e-30))


def fwd_bwd(net):
    def step(x):
        outs, loss = net(x)
        tot = sum(o.square().mean() for o in outs) + loss
        tot.backward()
        return tot.detach()
    return step


def grads(net):
    return {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}


def test_graphed_convnext_moe_step_matches_eager():
    from parity_util import build
    from sm3det_b200.graphed import GraphedStep
    from sm3det_b200.synth import make_images
    spec = CASES['mini2_moe_e8k2_train_clean']
    _, _, net = build(spec['kw'])
    net.train()
    xs = [make_images(*spec['img'], seed=s).cuda() for s in (1, 2, 3)]
    step = fwd_bwd(net)
    want = []
    for x in xs:
        net.zero_grad(set_to_none=True)
        want.append((step(x).clone(), grads(net)))
    net.zero_grad(set_to_none=True)
    g = GraphedStep(step, [xs[0]], net.parameters(), invalidate=[m._packs for m in net.modules() if hasattr(m, '_packs')])
    assert g.launches_per_replay > 100
    for x, (loss, gr) in zip(xs, want):
        got = g(x)
        torch.cuda.synchronize()
        assert abs(got.item() - loss.item()) <= 1e-5 * abs(loss.item())
        now = grads(net)
        assert set(now) == set(gr)
        worst = max((rel(now[k], gr[k]) / (5.0 if k.endswith('temperature') else 1.0), k) for k in gr if float(gr[k].abs().max()) > 1e-8)
        assert worst[0] < 3e-4, worst             # atomics reorder sums; nothing else differs between the two launch modes
    # an optimizer-style in-place weight update between replays is honoured (operand images are re-split inside the graph)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.01)
    got = g(xs[0]).clone()
    net.zero_grad(set_to_none=True)
    ref = step(xs[0])
    assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item()) and abs(got.item() - want[0][0].item()) > 1e-6 * abs(ref.item())


def test_graphed_lsk_step_draws_fresh_noise_and_dropout_masks():
    from test_lsk_gpu import build
    from sm3det_b200.graphed import GraphedStep
    from sm3det_b200.synth import make_images
    spec = LSK_CASES['lsk_mini_moe_e3k1_train_noisy_drop']
    _, _, net = build(spec['kw'])
    net.train()
    x = make_images(*spec['img'], seed=5).cuda()
    step = fwd_bwd(net)
    g = GraphedStep(step, [x], net.parameters())
    losses = []
    for _ in range(4):
        losses.append(g(x).item())
        assert all(torch.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)
    assert len({round(v, 10) for v in losses}) == 4, losses      # dropout seed, gating noise and drop-path masks change per replay
    # with every source of randomness off the replay equals eager
    spec = LSK_CASES['lsk_mini_moe_e4k2_train_clean']
    _, _, net = build(spec['kw'])
    net.train()
    x = make_images(*spec['img'], seed=6).cuda()
    step = fwd_bwd(net)
    rm0 = {k: v.clone() for k, v in net.state_dict().items() if 'running_' in k}
    ref = step(x).clone()
    gr = grads(net)
    net.zero_grad(set_to_none=True)
    net.load_state_dict(rm0, strict=False)                       # BatchNorm running statistics advance on every pass
    g = GraphedStep(step, [x], net.parameters(), warmup=1)
    net.load_state_dict(rm0, strict=False)
    got = g(x)
    torch.cuda.synchronize()
    assert abs(got.item() - ref.item()) <= 2e-5 * abs(ref.item())
    now = grads(net)
    # (w_gate.temperature: a scalar sum over all tokens with heavy cancellation; atomics reorder it between runs)
    worst = max((rel(now[k], gr[k]) / (5.0 if k.endswith('temperature') else 1.0), k) for k in gr if float(gr[k].abs().max()) > 1e-8)
    assert worst[0] < 5e-4, worst
