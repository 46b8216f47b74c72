"""`custom_imports=dict(imports=['sm3det_b200.mmrotate_plugin'])` hook: registers the CUDA backbones into the
real mmrotate registry (overriding the stock Python classes of the same name) when mmrotate is importable."""
from .registry import register_into_mmrotate

REGISTERED = register_into_mmrotate()
