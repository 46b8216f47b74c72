#!/usr/bin/env python
"""Benchmark of the SM3Det ConvNeXt-MoE backbone hot path (BASELINE.json metric: backbone images/s @1024^2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = forward + backward of the backbone over one synthetic batch (BASELINE configs[1]:
ConvNeXt-T, E=8 top-2, MoE in the last two stages every other block, 8 x 3 x 1024 x 1024 per GPU,
fp32).  N > 1 runs one process per GPU (torchrun), DistributedDataParallel over NCCL, per-GPU batch
fixed (weak scaling); the timed region is bracketed by barrier + synchronize and the max over ranks
is reported.  `--impl reference` times the reference's own CPU implementation of the same path (the
oracle port: identical torch CPU ops in the reference's order) on the box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL_KW = dict(arch='tiny', MoE_Block_inds=[[], [], [0, 2, 4, 6, 8], [0, 2]], num_experts=8, top_k=2,
                noisy_gating=False, drop_path_rate=0.0)
METRIC = 'backbone images/sec @1024^2 (fwd+bwd)'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=8, help='images per GPU per step')
    ap.add_argument('--size', type=int, default=1024)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--amp', action='store_true', help='NOT the headline: run the mixed-precision recipe (single-pass bf16 GEMMs)')
    ap.add_argument('--expert-parallel', action='store_true',
                    help='N > 1 only, not the headline config: shard the experts over the ranks (NVLink peer-memory dispatch)')
    ap.add_argument('--cpu-images', type=int, default=1, help='images in the bounded CPU sample')
    return ap.parse_args()


def workload_config(args, world):
    return {'workload': f'SM3Det ConvNeXt-T e8t2 last-2-blocks MoE backbone, fwd+bwd, bs={args.batch}/GPU x {world} GPU, '
                        f'{args.size}x{args.size}x3 synthetic SAR/RGB/IR 2:1:1, fp32, noisy_gating=False, drop_path=0',
            'arch': 'ConvNeXt-T', 'num_experts': 8, 'top_k': 2, 'moe_blocks': MODEL_KW['MoE_Block_inds'],
            'per_gpu_batch': args.batch, 'global_batch': args.batch * world, 'image': args.size,
            'parallelism': f'dp{world}' + ('+ep' if getattr(args, 'expert_parallel', False) and world > 1 else ''), 'l2': 'inputs and activations exceed L2 (>=100 MB per tensor); no flush needed'}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '200'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith('active')})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def cpu_reference_step(sd, cfg, x):
    """fwd+bwd of the oracle port on the host cores (train mode, clean gating) -- test infrastructure."""
    from oracle.convnext_moe_oracle import backbone_forward
    sdg = {k: (v.clone().requires_grad_(True) if 'ffn.mean' not in k and 'ffn.std' not in k else v) for k, v in sd.items()}
    outs, loss = backbone_forward(sdg, cfg, x, train=True)
    (sum(o.mean() for o in outs) + loss).backward()


def time_cpu_reference(args, images, steps, warmup):
    from oracle.convnext_moe_oracle import OracleConfig, param_shapes
    from sm3det_b200.synth import make_images, make_state_dict
    # torch's CPU kernels stop scaling (and then collapse: 143 s/img at 128 threads vs 1.5 s at 16 on the 128-thread
    # B200 host, profiles/r01_cpu_threads.txt) long before the box runs out of cores: use the best-performing count.
    threads = int(os.environ.get('SM3_CPU_THREADS', min(16, os.cpu_count() or 1)))
    torch.set_num_threads(threads)
    cfg = OracleConfig(**MODEL_KW)
    sd = make_state_dict(param_shapes(cfg), 0, True)
    x = make_images(images, args.size, args.size, seed=1234)
    cpu_reference_step(sd, cfg, make_images(1, 128, 128, seed=1))     # thread-pool / allocator warm-up, not timed
    for _ in range(warmup):
        cpu_reference_step(sd, cfg, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(sd, cfg, x)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return images / dt, dt, threads


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    ips, dt, threads = time_cpu_reference(args, args.cpu_images, steps, warmup)
    cfg = workload_config(args, args.gpus)
    line = {'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': 'img/s', 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': ips, 'unit': 'img/s', 'cores': threads, 'kind': 'port',
                             'sample': f'{args.cpu_images} images of the workload per step (fwd+bwd, torch CPU fp32, '
                                       f'{threads} threads); reference python executes the same ops'},
            'e2e': {'value': ips, 'unit': 'img/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def gemm_roofline(net, x, peaks):
    """Instrumented pass: CUDA-event time of every tensor-core GEMM launch in one fwd+bwd step and its
    algorithmic FLOPs (2*M*N*K; grouped launches count live rows only via the plan's token count)."""
    from sm3det_b200 import ops
    rec = []
    orig = ops.gemm

    def timed(**kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(**kw)
        e1.record()
        rows = kw['M']
        if kw.get('sched', 0) == ops.SCHED_GROUPED:
            rows = kw.get('_live_rows', rows)
        flops = 2.0 * rows * kw['N'] * kw['K']
        rec.append((e0, e1, flops, kw['M'], kw['N'], kw['K'], kw.get('sched', 0), 4.0 * (rows * kw['K'] + kw['N'] * kw['K'] + rows * kw['N'])))
        return r

    ops.gemm = timed
    try:
        outs, loss = net(x)
        (sum(o.mean() for o in outs) + loss).backward()
        torch.cuda.synchronize()
    finally:
        ops.gemm = orig
    tot_ms = sum(a.elapsed_time(b) for a, b, *_ in rec)
    if os.environ.get('SM3_GEMM_TABLE'):
        with open(os.environ['SM3_GEMM_TABLE'], 'w') as f:
            for a, b, fl, M, N, K, sched, _ in rec:
                ms = a.elapsed_time(b)
                f.write(f'M={M} N={N} K={K} sched={sched} ms={ms:.4f} tflops={fl / ms * 1e-9:.1f}\n')
    # grouped / split-K launches: M (or K) is the padded pair space; close enough for the aggregate (pad <= 1.5 %)
    tot_flops = sum(r[2] for r in rec)
    peak = peaks.get('bf16_tflops_sustained') or peaks.get('bf16_tflops') or 1590.0
    ach = tot_flops / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
    return {'bound': 'tensor', 'kernel': 'gemm_bf16x3_kernel (all launches of one fwd+bwd step)', 'achieved': ach,
            'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak, 'traffic': None, 'launches': len(rec),
            'algorithmic_bytes_per_launch': sum(r[7] for r in rec) / max(len(rec), 1),
            'gemm_ms_per_step': tot_ms,
            'note': 'algorithmic fp32 FLOPs; each costs 3 bf16 tensor-core MACs (hi*hi+hi*lo+lo*hi), so the tensor pipe '
                    'runs at 3x this rate; peak = measured cuBLAS bf16 (sustained) from MEASURED_PEAKS.json'
                    if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 'peak = fallback 1.59 PFLOP/s'}


def moe_roofline(net, x, peaks):
    """BASELINE metric (2): achieved HBM GB/s of the MoE dispatch(+expert) path of one forward, against the measured copy
    bandwidth.  Timed with CUDA events around (a) the whole 3-kernel-sequence the north star names -- dispatch gather
    (fused into the operand pack), grouped expert GEMM pair, combine scatter -- and (b) the dispatch kernels alone
    (gather-pack + combine).  Algorithmic bytes: SURVEY.md 8(d): (5k+1)*T*C*4 + E*(8C^2+5C)*4 + 8kT for the sequence,
    (3k+1)*T*C*4 for gather + scatter."""
    from sm3det_b200 import ops
    seq, disp = [], []
    o_assign, o_combine, o_pack = ops.moe_assign, ops.moe_combine, ops.pack_act
    state = {}

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def assign(top_idx, plan, *, T, E, k):
        r = o_assign(top_idx, plan, T=T, E=E, k=k)
        state.update(t0=ev(), T=T, E=E, k=k)
        return r

    def pack(x_, **kw):
        if kw.get('row_index') is not None and 't0' in state:
            a = ev(); r = o_pack(x_, **kw); b = ev()
            state['gather'] = (a, b)
            return r
        return o_pack(x_, **kw)

    def combine(o, slot_of, top_idx, gate, gamma, resid, row_scale, *, T, Cc, k, want_y=False):
        a = ev()
        r = o_combine(o, slot_of, top_idx, gate, gamma, resid, row_scale, T=T, Cc=Cc, k=k, want_y=want_y)
        b = ev()
        E = state['E']
        seq.append((state.pop('t0'), b, ((5 * k + 1) * T * Cc * 4 + E * (8 * Cc * Cc + 5 * Cc) * 4 + 8 * k * T), 16.0 * k * T * Cc * Cc))
        g = state.pop('gather', None)
        if g is not None:
            disp.append((g[0], g[1], a, b, (3 * k + 1) * T * Cc * 4))
        return r

    ops.moe_assign, ops.moe_combine, ops.pack_act = assign, combine, pack
    try:
        with torch.no_grad():
            net(x)
        torch.cuda.synchronize()
    finally:
        ops.moe_assign, ops.moe_combine, ops.pack_act = o_assign, o_combine, o_pack
    peak = peaks.get('hbm_gbs') or 6550.0
    ms = sum(a.elapsed_time(b) for a, b, *_ in seq)
    by = sum(r[2] for r in seq)
    fl = sum(r[3] for r in seq)
    dms = sum(a.elapsed_time(b) + c.elapsed_time(d) for a, b, c, d, _ in disp)
    dby = sum(r[4] for r in disp)
    out = {'bound': 'hbm', 'kernel': 'MoE dispatch+expert path (gather-pack -> grouped GEMM x2 -> combine), all MoE layers of one forward',
           'layers': len(seq), 'achieved': by / (ms * 1e-3) / 1e9 if ms else 0.0, 'peak': peak, 'unit': 'GB/s',
           'ms': ms, 'algorithmic_bytes': by, 'expert_tflops': fl / (ms * 1e-3) / 1e12 if ms else 0.0,
           'note': 'the expert GEMM pair is tensor-bound (16kTC^2 FLOP on 3-pass split-bf16), so the sequence cannot reach the '
                   'HBM roofline; dispatch_only isolates the HBM-bound gather + scatter kernels'}
    out['frac'] = out['achieved'] / peak
    if dms:
        out['dispatch_only'] = {'achieved': dby / (dms * 1e-3) / 1e9, 'frac': dby / (dms * 1e-3) / 1e9 / peak, 'ms': dms,
                                'algorithmic_bytes': dby}
    return out


def run_ours(args):
    import torch.distributed as dist
    from sm3det_b200 import ConvNeXt_moe_MultiInput, _lib
    from sm3det_b200.synth import make_images, make_state_dict

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl')
    lib = _lib.load()
    assert lib.sm3_device_supported() == 1, 'bench.py needs an sm_100 (B200) device'

    net = ConvNeXt_moe_MultiInput(**MODEL_KW)
    # seeded "trained-like" weights keyed by state_dict name (the product leg never touches oracle/)
    sd = make_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, 0, True)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    model = net
    if world > 1 and args.expert_parallel:
        from sm3det_b200.expert_parallel import enable_expert_parallel
        enable_expert_parallel(net, dist.new_group(list(range(world))))
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], broadcast_buffers=False,
                                                          gradient_as_bucket_view=True)
    B, S = args.batch, args.size
    host_x = make_images(B, S, S, seed=1234 + rank).pin_memory()
    dev_x = host_x.cuda()

    def step(x):
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=args.amp):
            outs, loss = model(x)
        tot = sum(o.float().mean() for o in outs) + loss
        tot.backward()
        return tot

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(dev_x)
        model.zero_grad(set_to_none=True)
    sampler = ClockSampler(local)
    # ---- device-resident timing -------------------------------------------------------------------
    _lib.LAUNCHES = 0
    sync()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(dev_x)
        model.zero_grad(set_to_none=True)
    e1.record()
    sync()
    launches = _lib.LAUNCHES
    ms = e0.elapsed_time(e1) / args.steps
    # ---- end to end: pinned host input -> device, result scalar back to the host, every step --------
    # Every step's batch is copied from pinned host memory inside the timed region; the copy of step i+1 is issued on
    # a side stream before step i computes (double-buffered), the way a data loader feeds the reference's train loop.
    h2d = host_x.numel() * 4
    copy_stream = torch.cuda.Stream()
    bufs = [torch.empty_like(dev_x), torch.empty_like(dev_x)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    main = torch.cuda.current_stream()

    def prefetch(slot, first_use):
        with torch.cuda.stream(copy_stream):
            if not first_use:
                copy_stream.wait_event(freed[slot])        # the step that last read this buffer has finished
            bufs[slot].copy_(host_x, non_blocking=True)
            ready[slot].record(copy_stream)

    sync()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    acc = 0.0
    # the 4-byte result of every step is copied to pinned host memory asynchronously and read one step later, so the
    # host keeps enqueueing step i+1 while step i runs (a blocking .item() per step drains the launch queue: ~1.5 ms/step)
    host_res = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    res_ready = [torch.cuda.Event(), torch.cuda.Event()]
    prefetch(0, True)
    for i in range(args.steps):
        cur = i & 1
        if i + 1 < args.steps:
            prefetch(cur ^ 1, i == 0)
        main.wait_event(ready[cur])
        tot = step(bufs[cur])
        freed[cur].record(main)
        host_res[cur].copy_(tot.detach(), non_blocking=True)      # D2H read of the step result (4 bytes)
        res_ready[cur].record(main)
        if i > 0:
            res_ready[cur ^ 1].synchronize()
            acc += float(host_res[cur ^ 1])
        model.zero_grad(set_to_none=True)
    res_ready[(args.steps - 1) & 1].synchronize()
    acc += float(host_res[(args.steps - 1) & 1])
    e3.record()
    sync()
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = e2.elapsed_time(e3) / args.steps
    t = torch.tensor([ms, ms_e2e], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        ep_mode = world > 1 and args.expert_parallel
        roof = roof_moe = None
        if not ep_mode and not args.amp:        # the instrumented extra passes are rank-0 only; expert parallelism needs every rank in each layer
            roof = gemm_roofline(net, dev_x, peaks)
            net.zero_grad(set_to_none=True)
            try:   # measured DRAM traffic of the GEMM launches of one step (ncu dram__bytes_read+write, profiles/)
                tr = json.load(open(os.path.join(ROOT, 'profiles', 'r01_gemm_traffic.json')))
                roof['traffic'] = tr['bytes_per_launch']
                roof['traffic_note'] = tr['note']
            except Exception:
                pass
            roof_moe = moe_roofline(net, dev_x, peaks)
        line = {'metric': METRIC, 'value': B * world / (ms * 1e-3), 'unit': 'img/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': ('bf16 GEMM operands (single pass), fp32 accumulate and fp32 elsewhere -- optional AMP recipe, not the headline'
                          if args.amp else 'f32 (bf16 hi+lo split operands on tcgen05, fp32 accumulate; SIMT fp32 elsewhere)'),
                'data': 'synthetic', 'config': workload_config(args, world), 'clocks': clocks,
                'e2e': {'value': B * world / (ms_e2e * 1e-3), 'unit': 'img/s', 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e},
                'gpu_launches': launches, 'roofline': roof, 'roofline_moe': roof_moe}
        if world == 1 and not args.no_cpu_baseline:
            ips, dt, threads = time_cpu_reference(args, args.cpu_images, 1, 0)
            line['cpu_baseline'] = {'value': ips, 'unit': 'img/s', 'cores': threads, 'kind': 'port',
                                    'sample': f'one fwd+bwd of {args.cpu_images} workload images (oracle port = the '
                                              f'reference\'s torch CPU ops), {threads} threads, {dt:.1f} s'}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
