"""CPU: world_size-2 (gloo) run of render_image's tile shard + single packed all-gather."""
import os
import subprocess
import sys

from conftest import ROOT


def test_render_image_two_ranks_gloo():
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_worker.py"), "--backend", "gloo"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "DIST_OK gloo 2" in res.stdout


def test_training_gradient_exchange_two_ranks_gloo():
    """N > 1 training path: flat all-reduce of the dense gradients + in-place all-reduce of the
    table gradients, averaged over the ranks."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "dist_train_worker.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "TRAIN_EXCHANGE_OK gloo 2" in res.stdout
