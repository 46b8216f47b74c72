#!/usr/bin/env python
"""Finite-difference check of the LSKNet-MoE CPU oracle's parameter gradients (float64, mini widths, 64^2 images).

Why this exists: at BATCH SIZE 1, torch 2.11 CPU autograd returns gradients for this op sequence (= the reference's own
lsk_moe.py ops, executed by oracle/lsk_moe_oracle.py and bit-identically by the unmodified reference) that DISAGREE with
central finite differences of the same float64 forward, while at batch size 2 they agree to 7 digits:

    n=2 patch_embed4.proj.weight(3, 5, 1, 1): FD -1.969265e-02  autograd -1.969265e-02
    n=1 patch_embed4.proj.weight(3, 5, 1, 1): FD -2.524361e-02  autograd +3.568306e-02
    n=1 patch_embed2.norm.weight(9,):         FD -2.106801e-01  autograd +7.772075e-01

Every individual op passes torch.autograd.gradcheck at N = 1, so this is a framework problem of the composite graph, not of
the reference's math.  The CUDA path agrees with the finite differences (tests/diag/diag_lsk.py on the GPU box), so the LSKNet
full-size fixtures are generated with batch 2 (oracle/cases.py) -- config 5 runs 4 images per GPU anyway.
"""
import os
import sys

import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.cases import upstream_grads
from oracle.lsk_moe_oracle import LskConfig, lsk_backbone_forward, lsk_param_shapes
from sm3det_b200.synth import make_images, make_state_dict
torch.set_default_dtype(torch.float64)
kw = dict(embed_dims=[64, 64, 128, 128], depths=[1, 1, 2, 1], mlp_ratios=[4, 4, 2, 2])
cfg = LskConfig(**kw)
base = {k: (v.double() if v.is_floating_point() else v) for k, v in make_state_dict(lsk_param_shapes(cfg), 0, True).items()}
skip = ('running_', 'num_batches', '.mean', '.std')
for n in (2, 1):
    x = make_images(n, 64, 64, seed=1234).double()
    ups = None
    def loss_of(sd):
        global ups
        oc = lsk_backbone_forward(sd, cfg, x, train=True, bn_state={})
        if ups is None:
            ups = [u.double() for u in upstream_grads([o.detach().float() for o in oc])]
        return sum((o * g).sum() for o, g in zip(oc, ups))
    sdo = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not any(t in k for t in skip) else v) for k, v in base.items()}
    loss_of(sdo).backward()
    for key, idx in (('patch_embed4.proj.weight', (3, 5, 1, 1)), ('block1.0.layer_scale_1', (7,)), ('block3.1.mlp.fc1.weight', (2, 3, 0, 0)), ('patch_embed2.norm.weight', (9,))):
        outs = []
        for sgn in (1, -1):
            sd2 = {k: v.clone() for k, v in base.items()}
            sd2[key][idx] += sgn * 1e-6
            with torch.no_grad():
                outs.append(float(loss_of(sd2)))
        fd = (outs[0] - outs[1]) / 2e-6
        print(f'n={n} {key}{idx}: FD {fd:+.6e}  autograd {float(sdo[key].grad[idx]):+.6e}')
