import os, sys, time, torch
sys.path.insert(0, '.')
import bench
from oracle.convnext_moe_oracle import OracleConfig, param_shapes
from sm3det_b200.synth import make_images, make_state_dict
cfg = OracleConfig(**bench.MODEL_KW); sd = make_state_dict(param_shapes(cfg), 0, True)
x = make_images(1, 1024, 1024, seed=1234)
for th in (16, 32, 64, 128):
    torch.set_num_threads(th)
    bench.cpu_reference_step(sd, cfg, make_images(1, 128, 128, seed=1))
    t0 = time.perf_counter(); bench.cpu_reference_step(sd, cfg, x); dt = time.perf_counter() - t0
    print(f'threads {th}: fwd+bwd 1 image {dt:.2f} s', flush=True)
