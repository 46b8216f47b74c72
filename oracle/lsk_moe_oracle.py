"""CPU oracle: functional restatement of SM3Det's LSKNet-MoE backbone forward (BASELINE config 5).

TEST INFRASTRUCTURE ONLY -- not part of the product.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline legs may import this file.

What it restates (all line numbers: /root/reference/mmrotate/models/backbones/lsk_moe.py):
  CosineTopKGate.forward :73-78, MoE_layer.cv_squared :108-115, _prob_in_top_k :120-142,
  noisy_top_k_gating :163-192, MoE_layer.forward :195-228 (NCHW in/out, single Conv2d(in,out,1)
  experts, empty experts skipped :215-217), SparseDispatcher :230-273 (combine WITHOUT .float() :263),
  Mlp.forward :303-318, DWConv :580-587, LSKblock.forward :329-343, Attention.forward :355-363,
  Block.forward :387-396, OverlapPatchEmbed.forward :413-418, LSKNet_moe.forward_features :541-559,
  LSKNet_moe_MultiInput.__init__ :689-695 (stem conv under dataset_stems, patch_embed1 = BN only),
  forward_features :716-739, forward :740-765.

Same torch CPU ops in the same order on a plain ``state_dict``; differentiable through autograd
exactly like the reference.  BatchNorm running statistics are functional: ``bn_state`` (a dict of
clones) receives the in-place updates F.batch_norm makes in training mode.

Parity pinning: the reference ships no tests for this path; ``oracle/gen_golden.py`` runs the real
``lsk_moe.py`` through ``oracle/ref_shim.py`` and asserts this file reproduces it bit-for-bit.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from .convnext_moe_oracle import cosine_gate, cv_squared, prob_in_top_k


@dataclass
class LskConfig:
    """Constructor kwargs of LSKNet_moe_MultiInput that change the math (lsk_moe.py:602-628)."""
    MoE_Block_inds_fc1: Sequence[Sequence[int]] = field(default_factory=lambda: [[], [], [], []])
    MoE_Block_inds_fc2: Sequence[Sequence[int]] = field(default_factory=lambda: [[], [], [], []])
    num_experts: int = 2
    top_k: int = 2
    noisy_gating: bool = True
    gate: str = 'cosine'
    in_channels: int = 3
    embed_dims: Sequence[int] = (32, 64, 160, 256)
    mlp_ratios: Sequence[int] = (8, 8, 4, 4)
    depths: Sequence[int] = (3, 3, 5, 2)
    drop_rate: float = 0.0
    drop_path_rate: float = 0.0
    num_stages: int = 4
    bn_eps: float = 1e-5          # nn.BatchNorm2d default (build_norm_layer BN/SyncBN)
    bn_momentum: float = 0.1
    ln_eps: float = 1e-6          # norm_layer=partial(nn.LayerNorm, eps=1e-6) :606
    layer_scale_init_value: float = 1e-2   # :381
    multi_input: bool = True
    spatial_unit: str = 'lsk'     # 'lka' = VAN_moe (van_moe.py: same file with LKA :319-333 as the gating unit)

    def moe_fc1(self, stage):
        return [q for q in self.MoE_Block_inds_fc1[stage] if q < self.depths[stage]]   # :451

    def moe_fc2(self, stage):
        return [q for q in self.MoE_Block_inds_fc2[stage] if q < self.depths[stage]]   # :452


def _bn_shapes(sh, p, c):
    sh[p + 'weight'] = (c,)
    sh[p + 'bias'] = (c,)
    sh[p + 'running_mean'] = (c,)
    sh[p + 'running_var'] = (c,)
    sh[p + 'num_batches_tracked'] = ()


def _moe_shapes(sh, p, cin, cout, cfg):
    sh[p + 'w_noise'] = (cin, cfg.num_experts)
    sh[p + 'mean'] = (1,)
    sh[p + 'std'] = (1,)
    for e in range(cfg.num_experts):
        sh[p + f'experts.{e}.weight'] = (cout, cin, 1, 1)
        sh[p + f'experts.{e}.bias'] = (cout,)
    if cfg.gate == 'cosine':
        P = min(cin // 2, 256)
        sh[p + 'w_gate.temperature'] = (1,)
        sh[p + 'w_gate.sim_matrix'] = (P, cfg.num_experts)
        sh[p + 'w_gate.cosine_projector.weight'] = (P, cin)
        sh[p + 'w_gate.cosine_projector.bias'] = (P,)
    else:
        sh[p + 'w_gate'] = (cin, cfg.num_experts)


def lsk_param_shapes(cfg: LskConfig) -> Dict[str, tuple]:
    """state_dict key -> shape (LSK keys), parameters and buffers."""
    sh: Dict[str, tuple] = {}
    D = list(cfg.embed_dims)
    for i in range(cfg.num_stages):
        c = D[i]
        pe = f'patch_embed{i + 1}.'
        if i == 0 and cfg.multi_input:
            _bn_shapes(sh, pe, c)                                  # patch_embed1 = BN only :692-695
        else:
            ks = 7 if i == 0 else 3
            sh[pe + 'proj.weight'] = (c, cfg.in_channels if i == 0 else D[i - 1], ks, ks)
            sh[pe + 'proj.bias'] = (c,)
            _bn_shapes(sh, pe + 'norm.', c)
        hid = int(c * cfg.mlp_ratios[i])
        for j in range(cfg.depths[i]):
            p = f'block{i + 1}.{j}.'
            sh[p + 'layer_scale_1'] = (c,)
            sh[p + 'layer_scale_2'] = (c,)
            _bn_shapes(sh, p + 'norm1.', c)
            _bn_shapes(sh, p + 'norm2.', c)
            a = p + 'attn.'
            sh[a + 'proj_1.weight'] = (c, c, 1, 1); sh[a + 'proj_1.bias'] = (c,)
            g = a + 'spatial_gating_unit.'
            sh[g + 'conv0.weight'] = (c, 1, 5, 5); sh[g + 'conv0.bias'] = (c,)
            sh[g + 'conv_spatial.weight'] = (c, 1, 7, 7); sh[g + 'conv_spatial.bias'] = (c,)
            if cfg.spatial_unit == 'lka':
                sh[g + 'conv1.weight'] = (c, c, 1, 1); sh[g + 'conv1.bias'] = (c,)
            else:
                sh[g + 'conv1.weight'] = (c // 2, c, 1, 1); sh[g + 'conv1.bias'] = (c // 2,)
                sh[g + 'conv2.weight'] = (c // 2, c, 1, 1); sh[g + 'conv2.bias'] = (c // 2,)
                sh[g + 'conv_squeeze.weight'] = (2, 2, 7, 7); sh[g + 'conv_squeeze.bias'] = (2,)
                sh[g + 'conv.weight'] = (c, c // 2, 1, 1); sh[g + 'conv.bias'] = (c,)
            sh[a + 'proj_2.weight'] = (c, c, 1, 1); sh[a + 'proj_2.bias'] = (c,)
            m = p + 'mlp.'
            if j in cfg.moe_fc1(i):
                _moe_shapes(sh, m + 'fc1.', c, hid, cfg)
            else:
                sh[m + 'fc1.weight'] = (hid, c, 1, 1); sh[m + 'fc1.bias'] = (hid,)
            sh[m + 'dwconv.dwconv.weight'] = (hid, 1, 3, 3); sh[m + 'dwconv.dwconv.bias'] = (hid,)
            if j in cfg.moe_fc2(i):
                _moe_shapes(sh, m + 'fc2.', hid, c, cfg)
            else:
                sh[m + 'fc2.weight'] = (c, hid, 1, 1); sh[m + 'fc2.bias'] = (c,)
        sh[f'norm{i + 1}.weight'] = (c,)
        sh[f'norm{i + 1}.bias'] = (c,)
    if cfg.multi_input:
        sh['dataset_stems.single.weight'] = (D[0], cfg.in_channels, 7, 7)
        sh['dataset_stems.single.bias'] = (D[0],)
    return sh


# ------------------------------------------------------------------------------------------------
def batch_norm(x, sd, p, cfg: LskConfig, train: bool, bn_state: Optional[dict]):
    """nn.BatchNorm2d.forward (what build_norm_layer returns for BN / SyncBN at world size 1)."""
    rm, rv = sd[p + 'running_mean'], sd[p + 'running_var']
    if train:
        rm, rv = rm.detach().clone(), rv.detach().clone()
    y = F.batch_norm(x, rm, rv, sd[p + 'weight'], sd[p + 'bias'], train, cfg.bn_momentum, cfg.bn_eps)
    if train and bn_state is not None:
        bn_state[p + 'running_mean'] = rm
        bn_state[p + 'running_var'] = rv
    return y


_FORCE = None   # iterator of forced top-k index tensors (set by *_forward(forced_idx=...)); test-only
_FORCE_AMAX = None   # iterator of forced LSK channel-argmax tensors [N*H*W] (NHWC token order) per LSKblock; test-only
_AMAX_RECORD = None  # list receiving {own, forced, gap, scale} per LSKblock when forcing


def noisy_top_k_gating(x, sd, p, cfg: LskConfig, train: bool, noise=None, noise_epsilon=1e-2):
    """MoE_layer.noisy_top_k_gating lsk_moe.py:163-192."""
    global _FORCE
    E, k = cfg.num_experts, cfg.top_k
    if cfg.gate == 'linear':
        clean = x @ sd[p + 'w_gate']
    else:
        clean = cosine_gate(x, sd, p + 'w_gate.')
    noisy = stddev = None
    if cfg.noisy_gating and train:
        raw = x @ sd[p + 'w_noise']
        stddev = (F.softplus(raw) + noise_epsilon) * train
        eps_t = torch.randn_like(clean) if noise is None else noise
        noisy = clean + eps_t * stddev
        logits = noisy
    else:
        logits = clean
    top_logits, top_idx = logits.topk(min(k + 1, E), dim=-1)
    top_k_logits, top_k_idx = top_logits[:, :k], top_idx[:, :k]
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


def moe_conv_layer(x, sd, p, cfg: LskConfig, train: bool, noise=None, loss_coef=1e-2, record=None):
    """MoE_layer.forward lsk_moe.py:195-228 + SparseDispatcher :230-273.  x: NCHW -> (NCHW, loss)."""
    x = x.permute(0, 2, 3, 1)
    x_shape = x.shape
    x = x.reshape(-1, x.shape[-1])
    gates, load, info = noisy_top_k_gating(x, sd, p, cfg, train, noise)
    importance = gates.sum(dim=0)
    loss = cv_squared(importance) + cv_squared(load)
    loss = loss * loss_coef
    sorted_experts, index_sorted = torch.nonzero(gates).sort(0)
    _, expert_index = sorted_experts.split(1, dim=1)
    batch_index = sorted_experts[index_sorted[:, 1], 0]
    part_sizes = list((gates > 0).sum(0).cpu().numpy())
    gates_exp = gates[batch_index.flatten()]
    nonzero_gates = torch.gather(gates_exp, 1, expert_index)
    inp_exp = x[batch_index].squeeze(1)
    expert_inputs = torch.split(inp_exp, part_sizes, dim=0)
    outs = []
    for e in range(cfg.num_experts):
        if expert_inputs[e].shape[0] != 0:                                     # :215
            o = F.conv2d(expert_inputs[e].reshape(-1, x_shape[-1], 1, 1), sd[p + f'experts.{e}.weight'],
                         sd[p + f'experts.{e}.bias'])
            outs.append(o.reshape(expert_inputs[e].shape[0], -1))
    stitched = torch.cat(outs, 0).mul(nonzero_gates)
    zeros = torch.zeros(gates.size(0), outs[-1].size(1), requires_grad=True, device=stitched.device, dtype=stitched.dtype)
    y = zeros.index_add(0, batch_index, stitched)                              # :263 (no .float())
    if record is not None:
        record.append(dict(prefix=p, x=x.detach(), top_idx=info['top_idx'].detach(),
                           top_gates=info['top_gates'].detach(), logits=info['logits'].detach(),
                           importance=importance.detach(), load=load.detach().float(),
                           loss=loss.detach(), y=y.detach()))
    y = y.reshape(x_shape[0], x_shape[1], x_shape[2], -1)
    return y.permute(0, 3, 1, 2).contiguous(), loss


def dropout(x, rate, train, masks):
    """nn.Dropout (Mlp.drop :300, applied :311,316).  ``masks``: iterator of pre-scaled keep masks."""
    if rate == 0.0 or not train:
        return x
    if masks is None:
        return F.dropout(x, rate, True)
    return x * next(masks)


def mlp(x, sd, p, cfg: LskConfig, moe1: bool, moe2: bool, train: bool, noise_it, drop_it, record):
    """Mlp.forward :303-318."""
    loss = []
    if moe1:
        x, l1 = moe_conv_layer(x, sd, p + 'fc1.', cfg, train, None if noise_it is None else next(noise_it), record=record)
        loss.append(l1)
    else:
        x = F.conv2d(x, sd[p + 'fc1.weight'], sd[p + 'fc1.bias'])
    x = F.conv2d(x, sd[p + 'dwconv.dwconv.weight'], sd[p + 'dwconv.dwconv.bias'], padding=1, groups=x.shape[1])
    x = F.gelu(x)
    x = dropout(x, cfg.drop_rate, train, drop_it)
    if moe2:
        x, l2 = moe_conv_layer(x, sd, p + 'fc2.', cfg, train, None if noise_it is None else next(noise_it), record=record)
        loss.append(l2)
    else:
        x = F.conv2d(x, sd[p + 'fc2.weight'], sd[p + 'fc2.bias'])
    x = dropout(x, cfg.drop_rate, train, drop_it)
    if len(loss) > 0:
        return x, sum(loss) / len(loss)
    return x, None


def lsk_block(x, sd, p):
    """LSKblock.forward :329-343."""
    c = x.shape[1]
    attn1 = F.conv2d(x, sd[p + 'conv0.weight'], sd[p + 'conv0.bias'], padding=2, groups=c)
    attn2 = F.conv2d(attn1, sd[p + 'conv_spatial.weight'], sd[p + 'conv_spatial.bias'], padding=9, groups=c, dilation=3)
    attn1 = F.conv2d(attn1, sd[p + 'conv1.weight'], sd[p + 'conv1.bias'])
    attn2 = F.conv2d(attn2, sd[p + 'conv2.weight'], sd[p + 'conv2.bias'])
    attn = torch.cat([attn1, attn2], dim=1)
    avg_attn = torch.mean(attn, dim=1, keepdim=True)
    max_attn, own_idx = torch.max(attn, dim=1, keepdim=True)
    forced = None
    if _FORCE_AMAX is not None:
        # test-only teacher forcing of the channel argmax (same idea as _FORCE for the router): the max feature and its
        # gradient follow the channel the CUDA path selected, so near-tie flips do not hide behind a loose tolerance
        forced = next(_FORCE_AMAX).to(attn.device).long().reshape(own_idx.shape)
    if _AMAX_RECORD is not None:
        d = attn.detach()
        f = own_idx if forced is None else forced
        _AMAX_RECORD.append(dict(prefix=p, own=own_idx.detach().flatten(), forced=f.detach().flatten(),
                                 gap=(max_attn.detach() - d.gather(1, f)).flatten(), scale=d.abs().amax(1).flatten()))
    if forced is not None:
        max_attn = attn.gather(1, forced)
    agg = torch.cat([avg_attn, max_attn], dim=1)
    sig = F.conv2d(agg, sd[p + 'conv_squeeze.weight'], sd[p + 'conv_squeeze.bias'], padding=3).sigmoid()
    attn = attn1 * sig[:, 0, :, :].unsqueeze(1) + attn2 * sig[:, 1, :, :].unsqueeze(1)
    attn = F.conv2d(attn, sd[p + 'conv.weight'], sd[p + 'conv.bias'])
    return x * attn


def lka(x, sd, p):
    """LKA.forward van_moe.py:327-333."""
    c = x.shape[1]
    u = x.clone()
    attn = F.conv2d(x, sd[p + 'conv0.weight'], sd[p + 'conv0.bias'], padding=2, groups=c)
    attn = F.conv2d(attn, sd[p + 'conv_spatial.weight'], sd[p + 'conv_spatial.bias'], padding=9, groups=c, dilation=3)
    attn = F.conv2d(attn, sd[p + 'conv1.weight'], sd[p + 'conv1.bias'])
    return u * attn


def attention(x, sd, p, unit='lsk'):
    """Attention.forward :355-363."""
    shortcut = x.clone()
    x = F.conv2d(x, sd[p + 'proj_1.weight'], sd[p + 'proj_1.bias'])
    x = F.gelu(x)
    x = lsk_block(x, sd, p + 'spatial_gating_unit.') if unit == 'lsk' else lka(x, sd, p + 'spatial_gating_unit.')
    x = F.conv2d(x, sd[p + 'proj_2.weight'], sd[p + 'proj_2.bias'])
    return x + shortcut


def drop_path(x, rate, train, mask=None):
    if rate == 0.0 or not train:
        return x
    keep = 1.0 - rate
    if mask is None:
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0:
            mask = mask / keep
    return x * mask


def block(x, sd, p, cfg: LskConfig, moe1, moe2, dpr, train, bn_state, noise_it, drop_it, dp_mask, record):
    """Block.forward :387-396."""
    ls1 = sd[p + 'layer_scale_1'].unsqueeze(-1).unsqueeze(-1)
    ls2 = sd[p + 'layer_scale_2'].unsqueeze(-1).unsqueeze(-1)
    x = x + drop_path(ls1 * attention(batch_norm(x, sd, p + 'norm1.', cfg, train, bn_state), sd, p + 'attn.', cfg.spatial_unit),
                      dpr, train, dp_mask)
    y, loss = mlp(batch_norm(x, sd, p + 'norm2.', cfg, train, bn_state), sd, p + 'mlp.', cfg, moe1, moe2, train,
                  noise_it, drop_it, record)
    x = x + drop_path(ls2 * y, dpr, train, dp_mask)
    return x, loss


def lsk_backbone_forward(sd: Dict[str, torch.Tensor], cfg: LskConfig, x, train: bool = False,
                         noise: Optional[List[torch.Tensor]] = None, drop_masks: Optional[List[torch.Tensor]] = None,
                         dp_masks: Optional[List[torch.Tensor]] = None, record: Optional[list] = None,
                         bn_state: Optional[dict] = None, forced_idx: Optional[List[torch.Tensor]] = None,
                         forced_amax: Optional[List[torch.Tensor]] = None, amax_record: Optional[list] = None):
    """LSKNet_moe_MultiInput.forward :740-765 (datasets=None path) + forward_features :716-739,
    or LSKNet_moe.forward_features :541-559 when ``cfg.multi_input`` is False."""
    global _FORCE, _FORCE_AMAX, _AMAX_RECORD
    _FORCE = None if forced_idx is None else iter(forced_idx)
    _FORCE_AMAX = None if forced_amax is None else iter(forced_amax)
    _AMAX_RECORD = amax_record
    try:
        return _lsk_backbone_forward(sd, cfg, x, train, noise, drop_masks, dp_masks, record, bn_state)
    finally:
        _FORCE = _FORCE_AMAX = _AMAX_RECORD = None


def _lsk_backbone_forward(sd, cfg, x, train, noise, drop_masks, dp_masks, record, bn_state):
    if isinstance(x, (list, tuple)):
        x = torch.cat(list(x), dim=0)
    D, depths = list(cfg.embed_dims), list(cfg.depths)
    dpr = [v.item() for v in torch.linspace(0, cfg.drop_path_rate, sum(depths))]   # :446
    noise_it = None if noise is None else iter(noise)
    drop_it = None if drop_masks is None else iter(drop_masks)
    if cfg.multi_input:
        x = F.conv2d(x, sd['dataset_stems.single.weight'], sd['dataset_stems.single.bias'], stride=4, padding=3)
    B = x.shape[0]
    outs, gate_losses = [], []
    cur = 0
    for i in range(cfg.num_stages):
        pe = f'patch_embed{i + 1}.'
        if i == 0 and cfg.multi_input:
            x = batch_norm(x, sd, pe, cfg, train, bn_state)
        else:
            ks = 7 if i == 0 else 3
            x = F.conv2d(x, sd[pe + 'proj.weight'], sd[pe + 'proj.bias'], stride=4 if i == 0 else 2, padding=ks // 2)
            x = batch_norm(x, sd, pe + 'norm.', cfg, train, bn_state)
        H, W = x.shape[2], x.shape[3]
        for j in range(depths[i]):
            x, loss = block(x, sd, f'block{i + 1}.{j}.', cfg, j in cfg.moe_fc1(i), j in cfg.moe_fc2(i), dpr[cur + j],
                            train, bn_state, noise_it, drop_it, None if dp_masks is None else dp_masks[cur + j], record)
            if loss is not None:
                gate_losses.append(loss)
        cur += depths[i]
        x = x.flatten(2).transpose(1, 2)
        x = F.layer_norm(x, (D[i],), sd[f'norm{i + 1}.weight'], sd[f'norm{i + 1}.bias'], cfg.ln_eps)
        x = x.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
        outs.append(x)
    if len(gate_losses) > 0:
        return tuple(outs), sum(gate_losses) / len(gate_losses)
    return tuple(outs)
