"""CPU oracle: functional restatement of SM3Det's ConvNeXt-MoE backbone forward.

TEST INFRASTRUCTURE ONLY -- not part of the product.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this file; the product package
``sm3det_b200`` never does (it fails loudly when its CUDA library is missing).

What it restates (all line numbers: /root/reference/mmrotate/models/backbones/convnext_moe.py):
  LayerNorm2d.forward :34-47, CosineTopKGate.forward :99-106, MoE_layer.cv_squared :140-147,
  _gates_to_load :149-150, _prob_in_top_k :152-174, noisy_top_k_gating :194-223,
  MoE_layer.forward :226-248, SparseDispatcher :250-293, ConvNeXtBlock._inner_forward :343-372,
  FFN.forward :397-405, ConvNeXt_moe.forward :582-600, ConvNeXt_moe_MultiInput.forward :794-820.

The arithmetic of the reference lives in PyTorch (third party, not under /root/reference); this file
therefore issues the *same torch CPU ops in the same order* on a plain ``state_dict`` (so it is also a
fair CPU timing baseline, ``cpu_baseline.kind = "port"``), is differentiable through autograd exactly
like the reference, and additionally exposes the per-layer routing decisions the parity tests need.

Parity pinning: the reference ships NO tests or golden vectors for this path,
so the oracle is pinned against outputs of the reference module itself, executed in place through
``oracle/ref_shim.py`` by ``oracle/gen_golden.py``; the resulting fixtures live in ``tests/golden/``
and ``tests/test_oracle.py`` re-checks the oracle against them (and against the live reference when
``/root/reference`` is present).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

ARCH_SETTINGS = {  # convnext_moe.py:409-454
    'atto': dict(depths=[2, 2, 6, 2], channels=[40, 80, 160, 320]),
    'femto': dict(depths=[2, 2, 6, 2], channels=[48, 96, 192, 384]),
    'pico': dict(depths=[2, 2, 6, 2], channels=[64, 128, 256, 512]),
    'nano': dict(depths=[2, 2, 8, 2], channels=[80, 160, 320, 640]),
    'tiny': dict(depths=[3, 3, 9, 3], channels=[96, 192, 384, 768]),
    'small': dict(depths=[3, 3, 27, 3], channels=[96, 192, 384, 768]),
    'base': dict(depths=[3, 3, 27, 3], channels=[128, 256, 512, 1024]),
    'swin_large': dict(depths=[2, 2, 18, 2], channels=[192, 384, 768, 1536]),
    'large': dict(depths=[3, 3, 27, 3], channels=[192, 384, 768, 1536]),
    'xlarge': dict(depths=[3, 3, 27, 3], channels=[256, 512, 1024, 2048]),
    'huge': dict(depths=[3, 3, 27, 3], channels=[352, 704, 1408, 2816]),
}


@dataclass
class OracleConfig:
    """Constructor kwargs of ConvNeXt_moe(_MultiInput) that change the math (:456-485, :733-762)."""
    arch: object = 'tiny'
    in_channels: int = 3
    stem_patch_size: int = 4
    MoE_Block_inds: Sequence[Sequence[int]] = field(default_factory=lambda: [[], [], [], []])
    num_experts: int = 2
    top_k: int = 2
    noisy_gating: bool = True
    gate: str = 'cosine'
    out_indices: Sequence[int] = (0, 1, 2, 3)
    drop_path_rate: float = 0.0
    layer_scale_init_value: float = 1e-6
    eps: float = 1e-6
    multi_input: bool = True   # ConvNeXt_moe_MultiInput (stem conv under dataset_stems.single)
    da: bool = False           # ConvNeXt_DA_MultiInput (convnext_moe_DA.py): a DALayer gates every block's branch per dataset
    da_reduction: int = 16     # DALayer(channel, reduction=16) :296

    @property
    def depths(self):
        a = ARCH_SETTINGS[self.arch] if isinstance(self.arch, str) else self.arch
        return list(a['depths'])

    @property
    def channels(self):
        a = ARCH_SETTINGS[self.arch] if isinstance(self.arch, str) else self.arch
        return list(a['channels'])

    def moe_blocks(self, stage: int) -> List[int]:
        d = self.depths[stage]
        return [q for q in self.MoE_Block_inds[stage] if q < d]  # :559


def param_shapes(cfg: OracleConfig) -> Dict[str, tuple]:
    """state_dict key -> shape, parameters and buffers, in module order."""
    sh: Dict[str, tuple] = {}
    C = cfg.channels
    ps = cfg.stem_patch_size
    if cfg.multi_input:
        sh['downsample_layers.0.0.weight'] = (C[0],)
        sh['downsample_layers.0.0.bias'] = (C[0],)
    else:
        sh['downsample_layers.0.0.weight'] = (C[0], cfg.in_channels, ps, ps)
        sh['downsample_layers.0.0.bias'] = (C[0],)
        sh['downsample_layers.0.1.weight'] = (C[0],)
        sh['downsample_layers.0.1.bias'] = (C[0],)
    for i in range(1, 4):
        sh[f'downsample_layers.{i}.0.weight'] = (C[i - 1],)
        sh[f'downsample_layers.{i}.0.bias'] = (C[i - 1],)
        sh[f'downsample_layers.{i}.1.weight'] = (C[i], C[i - 1], 2, 2)
        sh[f'downsample_layers.{i}.1.bias'] = (C[i],)
    for i in range(4):
        c = C[i]
        moe = cfg.moe_blocks(i)
        for j in range(cfg.depths[i]):
            p = f'stages.{i}.{j}.'
            sh[p + 'gamma'] = (c,)
            sh[p + 'depthwise_conv.weight'] = (c, 1, 7, 7)
            sh[p + 'depthwise_conv.bias'] = (c,)
            sh[p + 'norm.weight'] = (c,)
            sh[p + 'norm.bias'] = (c,)
            if j in moe:
                sh[p + 'ffn.w_noise'] = (c, cfg.num_experts)
                sh[p + 'ffn.mean'] = (1,)
                sh[p + 'ffn.std'] = (1,)
                for e in range(cfg.num_experts):
                    q = p + f'ffn.experts.{e}.'
                    sh[q + 'pointwise_conv1.weight'] = (4 * c, c)
                    sh[q + 'pointwise_conv1.bias'] = (4 * c,)
                    sh[q + 'pointwise_conv2.weight'] = (c, 4 * c)
                    sh[q + 'pointwise_conv2.bias'] = (c,)
                if cfg.gate == 'cosine':
                    P = min(c // 2, 256)
                    sh[p + 'ffn.w_gate.temperature'] = (1,)
                    sh[p + 'ffn.w_gate.sim_matrix'] = (P, cfg.num_experts)
                    sh[p + 'ffn.w_gate.cosine_projector.weight'] = (P, c)
                    sh[p + 'ffn.w_gate.cosine_projector.bias'] = (P,)
                else:
                    sh[p + 'ffn.w_gate'] = (c, cfg.num_experts)
            else:
                sh[p + 'ffn.pointwise_conv1.weight'] = (4 * c, c)
                sh[p + 'ffn.pointwise_conv1.bias'] = (4 * c,)
                sh[p + 'ffn.pointwise_conv2.weight'] = (c, 4 * c)
                sh[p + 'ffn.pointwise_conv2.bias'] = (c,)
            if cfg.da:
                # DALayer.fc = ModuleList([Sequential(...)] * 3) (convnext_moe_DA.py:299-304): ONE Sequential registered three
                # times, so the state_dict lists the same two weights under fc.0 / fc.1 / fc.2
                for m in range(3):
                    sh[p + f'DA.fc.{m}.0.weight'] = (c // cfg.da_reduction, c)
                    sh[p + f'DA.fc.{m}.2.weight'] = (c, c // cfg.da_reduction)
        if i in cfg.out_indices:
            sh[f'norm{i}.weight'] = (c,)
            sh[f'norm{i}.bias'] = (c,)
    if cfg.multi_input:
        sh['dataset_stems.single.weight'] = (C[0], cfg.in_channels, ps, ps)
        sh['dataset_stems.single.bias'] = (C[0],)
    return sh


# ------------------------------------------------------------------------------------------------
def layernorm2d(x, w, b, eps, data_format='channel_first'):
    """LayerNorm2d.forward :34-47."""
    if data_format == 'channel_last':
        return F.layer_norm(x, (x.shape[-1],), w, b, eps)
    x = x.permute(0, 2, 3, 1)
    x = F.layer_norm(x, (x.shape[-1],), w, b, eps)
    return x.permute(0, 3, 1, 2).contiguous()


def ffn(x, sd, p):
    """FFN.forward :397-405 (use_grn=False in every SM3Det config)."""
    x = F.linear(x, sd[p + 'pointwise_conv1.weight'], sd[p + 'pointwise_conv1.bias'])
    x = F.gelu(x)
    return F.linear(x, sd[p + 'pointwise_conv2.weight'], sd[p + 'pointwise_conv2.bias'])


def cosine_gate(x, sd, p):
    """CosineTopKGate.forward :99-106; clamp_max = log(1/0.01) :96."""
    proj = F.linear(x, sd[p + 'cosine_projector.weight'], sd[p + 'cosine_projector.bias'])
    logits = torch.matmul(F.normalize(proj, dim=1), F.normalize(sd[p + 'sim_matrix'], dim=0))
    clamp_max = torch.log(torch.tensor(1. / 0.01)).item()
    logit_scale = torch.clamp(sd[p + 'temperature'], max=clamp_max).exp()
    return logits * logit_scale


def cv_squared(x):
    """MoE_layer.cv_squared :140-147."""
    eps = 1e-10
    if x.shape[0] == 1:
        return torch.Tensor([0])
    return x.float().var() / (x.float().mean() ** 2 + eps)


def _normal_cdf(v):
    # torch.distributions.Normal(0, 1).cdf :170-172
    return 0.5 * (1 + torch.erf(v / math.sqrt(2)))


def prob_in_top_k(clean, noisy, stddev, noisy_top_values, k):
    """MoE_layer._prob_in_top_k :152-174."""
    batch = clean.size(0)
    m = noisy_top_values.size(1)
    flat = noisy_top_values.flatten()
    pos_in = (torch.arange(batch) * m + k).to(flat.device)                    # .to(device): :157
    thr_in = torch.unsqueeze(torch.gather(flat, 0, pos_in), 1)
    is_in = torch.gt(noisy, thr_in)
    thr_out = torch.unsqueeze(torch.gather(flat, 0, pos_in - 1), 1)
    prob_in = _normal_cdf((clean - thr_in) / stddev)
    prob_out = _normal_cdf((clean - thr_out) / stddev)
    return torch.where(is_in, prob_in, prob_out)


_FORCE = None   # iterator of forced top-k index tensors (set by *_forward(forced_idx=...)); test-only


def noisy_top_k_gating(x, sd, p, cfg: OracleConfig, train: bool, noise=None, noise_epsilon=1e-2):
    """MoE_layer.noisy_top_k_gating :194-223.  ``noise`` (optional [T,E]) replaces randn_like :203."""
    global _FORCE
    E, k = cfg.num_experts, cfg.top_k
    if cfg.gate == 'linear':
        clean = x @ sd[p + 'w_gate']
    else:
        clean = cosine_gate(x, sd, p + 'w_gate.')
    noisy = None
    stddev = None
    if cfg.noisy_gating and train:
        raw = x @ sd[p + 'w_noise']
        stddev = (F.softplus(raw) + noise_epsilon) * train
        eps_t = torch.randn_like(clean) if noise is None else noise
        noisy = clean + eps_t * stddev
        logits = noisy
    else:
        logits = clean
    top_logits, top_idx = logits.topk(min(k + 1, E), dim=-1)
    top_k_logits = top_logits[:, :k]
    top_k_idx = top_idx[:, :k]
    if _FORCE is not None:
        # TEST-ONLY teacher forcing (no reference counterpart): route with the supplied top-k indices so the CUDA
        # path and this oracle can be compared element-wise on tokens whose (k)-vs-(k+1) logits are a numerical tie
        top_k_idx = next(_FORCE).long()
        top_k_logits = logits.gather(1, top_k_idx)
    top_k_gates = torch.softmax(top_k_logits, -1)
    zeros = torch.zeros_like(logits, requires_grad=True)
    gates = zeros.scatter(-1, top_k_idx, top_k_gates)
    if cfg.noisy_gating and k < E and train:
        load = prob_in_top_k(clean, noisy, stddev, top_logits, k).sum(0)
    else:
        load = (gates > 0).sum(0)
    return gates, load, dict(logits=logits, top_idx=top_k_idx, top_gates=top_k_gates)


def moe_layer(x, sd, p, cfg: OracleConfig, train: bool, noise=None, loss_coef=1e-2, record=None):
    """MoE_layer.forward :226-248 with SparseDispatcher :250-293."""
    x_shape = x.shape
    x = x.reshape(-1, x.shape[-1])
    gates, load, info = noisy_top_k_gating(x, sd, p, cfg, train, noise)
    importance = gates.sum(dim=0)
    loss = cv_squared(importance) + cv_squared(load)
    loss = loss * loss_coef
    # SparseDispatcher.__init__ :252-262
    sorted_experts, index_sorted = torch.nonzero(gates).sort(0)
    _, expert_index = sorted_experts.split(1, dim=1)
    batch_index = sorted_experts[index_sorted[:, 1], 0]
    part_sizes = list((gates > 0).sum(0).cpu().numpy())
    gates_exp = gates[batch_index.flatten()]
    nonzero_gates = torch.gather(gates_exp, 1, expert_index)
    # dispatch :264-266
    inp_exp = x[batch_index].squeeze(1)
    expert_inputs = torch.split(inp_exp, part_sizes, dim=0)
    outs = [ffn(expert_inputs[e], sd, p + f'experts.{e}.').reshape(-1, x_shape[-1])
            for e in range(cfg.num_experts)]
    # combine :269-284
    stitched = torch.cat(outs, 0).mul(nonzero_gates)
    zeros = torch.zeros(gates.size(0), outs[-1].size(1), requires_grad=True, device=stitched.device, dtype=stitched.dtype)   # :277
    y = zeros.index_add(0, batch_index, stitched.float())
    if record is not None:
        record.append(dict(prefix=p, x=x.detach(), top_idx=info['top_idx'].detach(),
                           top_gates=info['top_gates'].detach(), logits=info['logits'].detach(),
                           importance=importance.detach(), load=load.detach().float(),
                           loss=loss.detach(), y=y.detach()))
    return y.reshape(x_shape), loss


def drop_path(x, rate, train, mask=None):
    """timm DropPath semantics (reference imports it at :6,27; applied :338-339,370)."""
    if rate == 0.0 or not train:
        return x
    keep = 1.0 - rate
    if mask is None:
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0:
            mask = mask / keep
    return x * mask


DA_INDEX = {'sar': 0, 'rgb': 1, 'ifr': 2}      # DALayer.dataset_DA, convnext_moe_DA.py:305


def tie_da_weights(sd):
    """The three DALayer.fc entries are one module in the reference: make fc.1 / fc.2 the same tensors as fc.0 (what
    load_state_dict leaves behind is the LAST entry's values, so fc.2 wins).  Returns sd (modified in place)."""
    for k in list(sd):
        if '.DA.fc.2.' in k:
            sd[k.replace('.DA.fc.2.', '.DA.fc.0.')] = sd[k]
            sd[k.replace('.DA.fc.2.', '.DA.fc.1.')] = sd[k]
    return sd


def da_layer(x, sd, p, datasets):
    """DALayer.forward convnext_moe_DA.py:307-319: squeeze (global average pool) -> Linear/ReLU/Linear/Sigmoid chosen by the
    sample's dataset -> channel-wise scale.  One dataset name = the whole batch; otherwise one name per sample."""
    b, c = x.shape[:2]
    y = F.adaptive_avg_pool2d(x, 1).view(b, c)

    def fc(m, v):
        return torch.sigmoid(F.linear(F.relu(F.linear(v, sd[p + f'fc.{m}.0.weight'])), sd[p + f'fc.{m}.2.weight']))
    if len(datasets) == 1:
        s = fc(DA_INDEX[datasets[0]], y).view(b, c, 1, 1)
    else:
        s = torch.cat([fc(DA_INDEX[d], row.view(1, c)) for row, d in zip(y, datasets)], dim=1).view(b, c, 1, 1)
    return x * s.expand_as(x)


def convnext_block(x, sd, p, cfg: OracleConfig, is_moe: bool, dpr: float, train: bool,
                   noise=None, dp_mask=None, record=None, pre_gamma=None, datasets=None):
    """ConvNeXtBlock._inner_forward :343-372 (linear_pw_conv=True path); with cfg.da the DA variant
    convnext_moe_DA.py:374-403 (the branch is gated by DALayer before drop_path and the shortcut)."""
    shortcut = x
    C = x.shape[1]
    x = F.conv2d(x, sd[p + 'depthwise_conv.weight'], sd[p + 'depthwise_conv.bias'], padding=3,
                 groups=C)
    x = x.permute(0, 2, 3, 1)
    x = layernorm2d(x, sd[p + 'norm.weight'], sd[p + 'norm.bias'], cfg.eps, 'channel_last')
    loss = None
    if is_moe:
        x, loss = moe_layer(x, sd, p + 'ffn.', cfg, train, noise, record=record)
    else:
        x = ffn(x, sd, p + 'ffn.')
    if pre_gamma is not None:
        pre_gamma.append(x.detach())
    x = x.permute(0, 3, 1, 2)
    if cfg.layer_scale_init_value > 0:
        x = x.mul(sd[p + 'gamma'].view(1, -1, 1, 1))
    if cfg.da:
        x = da_layer(x, sd, p + 'DA.', datasets)                           # convnext_moe_DA.py:400
    x = shortcut + drop_path(x, dpr, train, dp_mask)
    return x, loss


def backbone_forward(sd: Dict[str, torch.Tensor], cfg: OracleConfig, x, train: bool = False,
                     noise: Optional[List[torch.Tensor]] = None,
                     dp_masks: Optional[List[torch.Tensor]] = None,
                     record: Optional[list] = None, pre_gamma: Optional[list] = None,
                     forced_idx: Optional[List[torch.Tensor]] = None, datasets: Optional[Sequence[str]] = None):
    """ConvNeXt_moe.forward :582-600 / ConvNeXt_moe_MultiInput.forward :794-820.

    ``forced_idx`` (test-only): per MoE layer, in forward order, a [T,k] index tensor that replaces the layer's own
    top-k choice (see noisy_top_k_gating).

    ``x``: Tensor [N,3,H,W] or list of such (concatenated on N, :798-800).
    Returns ``tuple(outs)`` or ``(tuple(outs), gate_loss)`` when at least one MoE block ran.
    """
    global _FORCE
    _FORCE = None if forced_idx is None else iter(forced_idx)
    try:
        return _backbone_forward(sd, cfg, x, train, noise, dp_masks, record, pre_gamma, datasets)
    finally:
        _FORCE = None


def _backbone_forward(sd, cfg, x, train, noise, dp_masks, record, pre_gamma, datasets=None):
    if isinstance(x, (list, tuple)):
        x = torch.cat(list(x), dim=0)
    C, depths = cfg.channels, cfg.depths
    ps = cfg.stem_patch_size
    total = sum(depths)
    dpr = [v.item() for v in torch.linspace(0, cfg.drop_path_rate, total)]  # :523-526
    if cfg.multi_input:
        x = F.conv2d(x, sd['dataset_stems.single.weight'], sd['dataset_stems.single.bias'], stride=ps)
    outs, gate_losses = [], []
    blk = 0
    moe_i = 0
    for i in range(4):
        if i == 0:
            if cfg.multi_input:
                x = layernorm2d(x, sd['downsample_layers.0.0.weight'], sd['downsample_layers.0.0.bias'], cfg.eps)
            else:
                x = F.conv2d(x, sd['downsample_layers.0.0.weight'], sd['downsample_layers.0.0.bias'], stride=ps)
                x = layernorm2d(x, sd['downsample_layers.0.1.weight'], sd['downsample_layers.0.1.bias'], cfg.eps)
        else:
            x = layernorm2d(x, sd[f'downsample_layers.{i}.0.weight'], sd[f'downsample_layers.{i}.0.bias'], cfg.eps)
            x = F.conv2d(x, sd[f'downsample_layers.{i}.1.weight'], sd[f'downsample_layers.{i}.1.bias'], stride=2)
        moe = cfg.moe_blocks(i)
        for j in range(depths[i]):
            is_moe = j in moe
            nz = None
            if is_moe and noise is not None:
                nz = noise[moe_i]
            x, loss = convnext_block(x, sd, f'stages.{i}.{j}.', cfg, is_moe, dpr[blk], train, nz,
                                     None if dp_masks is None else dp_masks[blk], record, pre_gamma, datasets)
            if is_moe:
                moe_i += 1
            if loss is not None:
                gate_losses.append(loss)
            blk += 1
        if i in cfg.out_indices:
            outs.append(layernorm2d(x, sd[f'norm{i}.weight'], sd[f'norm{i}.bias'], cfg.eps))
    if len(gate_losses) > 0:
        return tuple(outs), sum(gate_losses) / len(gate_losses)
    return tuple(outs)
