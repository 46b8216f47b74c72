"""Registry / BaseModule shims with the mmcv surface the backbone is built through.

Reference: ``ROTATED_BACKBONES = mmdet.models.builder.MODELS`` (mmrotate/models/builder.py:4-6);
``Registry.build -> build_from_cfg`` pops ``type`` and calls ``cls(**cfg)``, re-raising as
``type(e)(f'{cls.__name__}: {e}')`` (mmcv/mmcv/utils/registry.py:57-72); ``BaseModule`` keeps
``init_cfg`` and offers ``init_weights()`` (mmcv/mmcv/runner/base_module.py:15-213).

When mmcv / mmrotate are importable the classes are ALSO registered into the real registries by
``register_into_mmrotate()`` so ``configs/SM3Det/*.py`` build them unchanged; neither package is
required (none is installed in this image).
"""
import copy

import torch.nn as nn


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def __contains__(self, key):
        return key in self._module_dict

    def _register(self, cls, name=None, force=False):
        name = name or cls.__name__
        if not force and name in self._module_dict:
            raise KeyError(f'{name} is already registered in {self._name}')
        self._module_dict[name] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def _deco(cls):
            self._register(cls, name, force)
            return cls
        return _deco

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict):
            raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
        if 'type' not in cfg:
            raise KeyError(f'`cfg` must contain the key "type", but got {cfg}')
        args = copy.deepcopy(dict(cfg))
        if default_args is not None:
            for k, v in default_args.items():
                args.setdefault(k, v)
        obj_type = args.pop('type')
        if isinstance(obj_type, str):
            cls = self.get(obj_type)
            if cls is None:
                raise KeyError(f'{obj_type} is not in the {self._name} registry')
        elif isinstance(obj_type, type):
            cls = obj_type
        else:
            raise TypeError(f'type must be a str or valid type, but got {type(obj_type)}')
        try:
            return cls(**args)
        except Exception as e:
            raise type(e)(f'{cls.__name__}: {e}')


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    @property
    def is_init(self):
        return self._is_init

    def init_weights(self):
        self._is_init = True


ROTATED_BACKBONES = Registry('models')
BACKBONES = ROTATED_BACKBONES


def build_backbone(cfg):
    return ROTATED_BACKBONES.build(cfg)


def register_into_mmrotate():
    """Best effort: also expose the classes through the real mmrotate/mmdet registry if present."""
    try:
        from mmrotate.models.builder import ROTATED_BACKBONES as real   # type: ignore
    except Exception:
        return False
    for name, cls in ROTATED_BACKBONES.module_dict.items():
        real.register_module(name=name, force=True, module=cls)
    return True
