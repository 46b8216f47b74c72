"""CPU oracle for MultitaskFPN.  TEST INFRASTRUCTURE ONLY.

Restates mmrotate/models/necks/Multitask_FPN.py:108-162 (forward) for norm_cfg = act_cfg = None (ConvModule = biased
Conv2d), size-based nearest upsampling, extra levels as stride-2 3x3 convs; same torch CPU ops in the same order.
Pinned by tests/test_fpn.py against the unmodified reference module executed through oracle/ref_shim.py."""
import torch.nn.functional as F


def fpn_param_shapes(in_channels, out_channels, num_outs, extra_level=0, add_extra_convs=False):
    sh = {}
    n = len(in_channels)
    for i in range(n):
        sh[f'lateral_convs.{i}.conv.weight'] = (out_channels, in_channels[i], 1, 1)
        sh[f'lateral_convs.{i}.conv.bias'] = (out_channels,)
        sh[f'fpn_convs.{i}.conv.weight'] = (out_channels, out_channels, 3, 3)
        sh[f'fpn_convs.{i}.conv.bias'] = (out_channels,)
    extra = num_outs - n + extra_level
    if add_extra_convs and extra >= 1:
        for i in range(extra):
            cin = in_channels[-1] if (i == 0 and add_extra_convs in (True, 'on_input')) else out_channels
            sh[f'fpn_convs.{n + i}.conv.weight'] = (out_channels, cin, 3, 3)
            sh[f'fpn_convs.{n + i}.conv.bias'] = (out_channels,)
    return sh


def fpn_forward(sd, inputs, num_ins, num_outs, start_level=0, add_extra_convs='on_output'):
    """Multitask_FPN.py:108-162."""
    conv = lambda k, x, **kw: F.conv2d(x, sd[k + '.conv.weight'], sd[k + '.conv.bias'], **kw)
    laterals = [conv(f'lateral_convs.{i}', inputs[i]) for i in range(start_level, num_ins)]          # :115-118
    used = len(laterals)
    for i in range(used - 1, 0, -1):                                                                  # :121-134
        laterals[i - 1] = laterals[i - 1] + F.interpolate(laterals[i], size=laterals[i - 1].shape[2:], mode='nearest')
    outs = [conv(f'fpn_convs.{i + start_level}', laterals[i], padding=1) for i in range(used)]      # :138-140
    if num_outs > len(outs):                                                                          # :142-161
        assert add_extra_convs
        src = {'on_input': inputs[num_ins - 1], 'on_lateral': laterals[-1], 'on_output': outs[-1]}[add_extra_convs]
        outs.append(conv(f'fpn_convs.{used + start_level}', src, stride=2, padding=1))
        for i in range(used + 1, num_outs):
            outs.append(conv(f'fpn_convs.{i + start_level}', outs[-1], stride=2, padding=1))
    return tuple(outs)
