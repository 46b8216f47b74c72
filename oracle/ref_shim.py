"""Import shims that execute the UNMODIFIED reference backbone files in place.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sm3det_b200/`` may import this module.

The reference (``/root/reference/mmrotate/models/backbones/convnext_moe.py``) imports
``timm``, ``mmengine``, ``mmcv`` and ``mmrotate.models.builder`` -- none of which are installed in
this image.  We pre-seed ``sys.modules`` with minimal stand-ins and load the
reference source file with ``importlib`` under its real dotted name, so its relative import
``from ..builder import ROTATED_BACKBONES`` resolves to our stub.  No reference source is copied.

``/root/reference`` exists only in the build container, never on the GPU box: this module is used
by ``oracle/gen_golden.py`` (fixture generation + pinning the restated oracle) and by the
``not gpu`` tests that are skipped when the reference tree is absent.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("SM3DET_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mmrotate/models/backbones/convnext_moe.py"))


class _DropPath(nn.Module):
    """timm.models.layers.DropPath semantics: per-sample Bernoulli(keep), scaled by 1/keep."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


class _Registry:
    def __init__(self):
        self.module_dict = {}

    def register_module(self, *args, **kwargs):
        def deco(cls):
            self.module_dict[cls.__name__] = cls
            return cls
        return deco


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave as a package
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


_REGISTRY = _Registry()


def _install_shims():
    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def constant_init(module, val, bias=0):
        if getattr(module, "weight", None) is not None:
            nn.init.constant_(module.weight, val)
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias, bias)

    def trunc_normal_init(module, mean=0, std=1, a=-2, b=2, bias=0):
        if getattr(module, "weight", None) is not None:
            nn.init.trunc_normal_(module.weight, mean, std, a, b)
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias, bias)

    def normal_init(module, mean=0, std=1, bias=0):
        if getattr(module, "weight", None) is not None:
            nn.init.normal_(module.weight, mean, std)
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias, bias)

    def build_activation_layer(cfg):
        assert cfg["type"] == "GELU", cfg
        return nn.GELU()

    def build_norm_layer(cfg, num_features, postfix=""):
        t = cfg["type"]
        if t in ("BN", "SyncBN"):
            # CPU oracle: SyncBN needs a process group; plain BN is numerically identical at world 1
            return "bn" + str(postfix), nn.BatchNorm2d(num_features)
        if t == "LN":
            return "ln" + str(postfix), nn.LayerNorm(num_features)
        raise KeyError(t)

    _mod("timm"); _mod("timm.models")
    _mod("timm.models.layers", DropPath=_DropPath, trunc_normal_=nn.init.trunc_normal_,
         to_2tuple=to_2tuple)
    _mod("mmengine"); _mod("mmengine.runner")
    _mod("mmengine.model", ModuleList=nn.ModuleList, Sequential=nn.Sequential, BaseModule=_BaseModule)
    _mod("mmengine.logging", MMLogger=type("MMLogger", (), {
        "get_current_instance": staticmethod(lambda: None)}))
    _mod("mmengine.runner.checkpoint", CheckpointLoader=type("CheckpointLoader", (), {}))
    class ConvModule(nn.Module):
        """mmcv.cnn.ConvModule with norm_cfg = act_cfg = None (what MultitaskFPN builds in every SM3Det config)."""

        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, conv_cfg=None, norm_cfg=None,
                     act_cfg=None, inplace=False):
            super().__init__()
            assert conv_cfg is None and norm_cfg is None and act_cfg is None
            self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding)

        def forward(self, x):
            return self.conv(x)

    def auto_fp16(*a, **k):
        return lambda f: f

    _mod("mmcv")
    _mod("mmcv.cnn", ConvModule=ConvModule, build_activation_layer=build_activation_layer,
         build_norm_layer=build_norm_layer, constant_init=constant_init,
         trunc_normal_init=trunc_normal_init, normal_init=normal_init)
    _mod("mmcv.cnn.utils")
    _mod("mmcv.cnn.utils.weight_init", constant_init=constant_init, normal_init=normal_init,
         trunc_normal_init=trunc_normal_init)
    _mod("mmcv.runner", BaseModule=_BaseModule, _load_checkpoint=None, load_state_dict=None, auto_fp16=auto_fp16)
    _mod("mmcv.utils", to_2tuple=to_2tuple)
    _mod("mmrotate"); _mod("mmrotate.models"); _mod("mmrotate.models.backbones")
    _mod("mmrotate.models.necks")
    _mod("mmrotate.models.builder", ROTATED_BACKBONES=_REGISTRY, ROTATED_NECKS=_REGISTRY)
    _mod("mmrotate.utils", get_root_logger=lambda *a, **k: None)


def load_reference_module(name="convnext_moe", package="backbones"):
    """Execute ``<reference>/mmrotate/models/<package>/<name>.py`` in place and return the module."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    full = f"mmrotate.models.{package}.{name}"
    if full in sys.modules and getattr(sys.modules[full], "__file__", None):
        return sys.modules[full]
    _install_shims()
    path = os.path.join(REFERENCE_ROOT, "mmrotate/models", package, name + ".py")
    spec = importlib.util.spec_from_file_location(full, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def build_reference_backbone(cls_name="ConvNeXt_moe_MultiInput", seed=0, module="convnext_moe", **kwargs):
    """Construct the reference class under a fixed seed (it never calls init_weights())."""
    mod = load_reference_module(module)
    torch.manual_seed(seed)
    net = getattr(mod, cls_name)(**kwargs)
    return net
