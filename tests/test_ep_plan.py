"""world_size-2 gloo test (CPU) of the expert-parallel exchange plan: ragged token counts, an expert with no traffic from
one rank, 128-row padding, source-major segment order, and the slot -> (owner, row) map of the combine."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ep_exchange_plan_world2_gloo():
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29541',
                        os.path.join(ROOT, 'tests', 'dist', 'ep_plan_worker.py'), ROOT],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('ep plan ok') == 2, r.stdout[-2000:]


def test_syncbn_stat_combination_world2_gloo():
    """LSKNet SyncBN (BASELINE config 5): shifted-sum statistics all-reduced over 2 gloo ranks == full-batch BatchNorm."""
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29545',
                        os.path.join(ROOT, 'tests', 'dist', 'syncbn_worker.py'), ROOT],
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, MASTER_ADDR='127.0.0.1'))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('syncbn stats ok') == 2, r.stdout[-2000:]


def test_flat_gradient_allreduce_world2_gloo():
    """graphed.allreduce_gradients (the capturable replacement of DDP's bucket hooks): mean over ranks, skip list, None grads."""
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29546',
                        os.path.join(ROOT, 'tests', 'dist', 'gradsync_worker.py')],
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, MASTER_ADDR='127.0.0.1'))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('grad sync ok') == 2, r.stdout[-2000:]
