"""Deterministic synthetic inputs and weights (there is no dataset / checkpoint access).

Images follow SURVEY.md section 8(d): already-normalised tensors with SAR-like (speckled, one
channel replicated), RGB-like (3 low-passed channels + noise) and IR-like (smooth + hot spots)
statistics, mixed 2:1:1 as ``source_ratio=[2,1,1]`` (reference configs/SM3Det/SM3Det_convnext_t.py:10).

Weights are generated per state_dict key from ``crc32(key) ^ seed`` so that the CPU oracle, the
reference module (build container only) and the CUDA module can all be filled with bit-identical
values without shipping a checkpoint.
"""
import math
import zlib
from typing import Dict, Mapping, Sequence

import torch
import torch.nn.functional as F


def _gen(seed: int, tag: str) -> torch.Generator:
    g = torch.Generator(device='cpu')
    g.manual_seed((zlib.crc32(tag.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def make_images(n: int, h: int, w: int, seed: int = 1234, modality: str = 'mix') -> torch.Tensor:
    """fp32 CPU tensor [n,3,h,w].  modality in {'sar','rgb','ifr','mix','randn'}."""
    if modality == 'mix':
        order = ['sar', 'sar', 'rgb', 'ifr']
        return torch.cat([make_images(1, h, w, seed + 17 * i, order[i % 4]) for i in range(n)], 0)
    g = _gen(seed, modality)
    if modality == 'randn':
        return torch.randn(n, 3, h, w, generator=g)
    lo = max(1, min(8, h // 4, w // 4))

    def smooth(c):
        z = torch.randn(n, c, max(1, h // lo), max(1, w // lo), generator=g)
        return F.interpolate(z, size=(h, w), mode='bilinear', align_corners=False)

    if modality == 'sar':
        base = smooth(1).abs() + 0.2
        speckle = torch.empty(n, 1, h, w).exponential_(1.0, generator=g)
        for _ in range(3):
            speckle = speckle + torch.empty(n, 1, h, w).exponential_(1.0, generator=g)
        x = base * speckle / 4.0                      # Gamma(4, 1/4) multiplicative speckle
        x = (x - x.mean()) / (x.std() + 1e-6)
        return x.expand(n, 3, h, w).contiguous()
    if modality == 'rgb':
        x = smooth(3) + 0.3 * torch.randn(n, 3, h, w, generator=g)
        return (x - x.mean()) / (x.std() + 1e-6)
    if modality == 'ifr':
        x = smooth(1)
        hot = (torch.rand(n, 1, h, w, generator=g) > 0.999).float() * 6.0
        x = x + F.max_pool2d(hot, 5, 1, 2)
        x = (x - x.mean()) / (x.std() + 1e-6)
        return x.expand(n, 3, h, w).contiguous()
    raise ValueError(modality)


def make_state_dict(shapes: Mapping[str, Sequence[int]], seed: int = 0,
                    trained_like: bool = True) -> Dict[str, torch.Tensor]:
    """Fill every key of ``shapes`` (state_dict key -> shape) with seeded values.

    trained_like=False mimics the reference constructor (gamma = 1e-6, LN = (1, 0), trunc-normal-ish
    weights, temperature = ln 2); trained_like=True perturbs what training would move so that layer
    scale no longer hides FFN/MoE errors (SURVEY.md section 7 'hard parts').
    """
    sd = {}
    for key, shape in shapes.items():
        shape = tuple(shape)
        g = _gen(seed, key)
        leaf = key.split('.')[-1]
        is_norm = ('.norm.' in key or key.startswith('norm') or
                   (key.startswith('downsample_layers') and len(shape) == 1 and
                    not key.startswith(('downsample_layers.1.1', 'downsample_layers.2.1', 'downsample_layers.3.1'))))
        if key.startswith('downsample_layers.0.0') and len(shapes[key.rsplit('.', 1)[0] + '.weight']) == 4:
            is_norm = False                                    # plain ConvNeXt_moe: .0.0 is the stem conv
        if (key.rsplit('.', 1)[0] + '.running_mean') in shapes:
            is_norm = True                                     # BatchNorm affine (LSKNet: norm1/2, patch_embed*.norm)
        if leaf == 'gamma':
            t = torch.rand(shape, generator=g) * 0.9 + 0.1 if trained_like else torch.full(shape, 1e-6)
        elif leaf.startswith('layer_scale_'):                  # lsk_moe.py:381-385 (init 1e-2)
            t = torch.rand(shape, generator=g) * 0.9 + 0.1 if trained_like else torch.full(shape, 1e-2)
        elif leaf == 'running_mean':
            t = torch.randn(shape, generator=g) * 0.1 if trained_like else torch.zeros(shape)
        elif leaf == 'running_var':
            t = torch.rand(shape, generator=g) + 0.5 if trained_like else torch.ones(shape)
        elif leaf == 'num_batches_tracked':
            sd[key] = torch.tensor(0, dtype=torch.long)
            continue
        elif leaf == 'temperature':
            t = torch.full(shape, math.log(10.0) if trained_like else math.log(2.0))
        elif leaf == 'sim_matrix':
            t = torch.randn(shape, generator=g) * (1.0 if trained_like else 0.01)
        elif leaf == 'w_noise':
            t = torch.randn(shape, generator=g) * 0.05 if trained_like else torch.zeros(shape)
        elif leaf == 'mean':
            t = torch.zeros(shape)
        elif leaf == 'std':
            t = torch.ones(shape)
        elif is_norm:
            if leaf == 'weight':
                t = torch.rand(shape, generator=g) + 0.5 if trained_like else torch.ones(shape)
            else:
                t = torch.randn(shape, generator=g) * 0.1 if trained_like else torch.zeros(shape)
        elif leaf == 'weight' or leaf == 'w_gate':
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            std = (1.0 / math.sqrt(max(fan_in, 1))) if trained_like else 0.02
            t = torch.randn(shape, generator=g) * std
        elif leaf == 'bias':
            t = torch.randn(shape, generator=g) * (0.05 if trained_like else 0.0)
        else:
            raise KeyError(f'unknown state_dict key kind: {key}')
        sd[key] = t.float().contiguous()
    return sd


def state_dict_checksum(sd: Mapping[str, torch.Tensor]) -> float:
    """Order-independent fingerprint used by the golden fixtures to detect RNG drift."""
    return float(sum(v.double().abs().sum().item() * (1 + (zlib.crc32(k.encode()) % 97)) for k, v in sd.items()))
