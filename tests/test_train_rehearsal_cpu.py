"""CPU: rehearsal of the `-m gpu` tests of the training-step ops (tests/dryrun_train_gpu_tests_on_cpu.py, run in a
subprocess because it patches torch): the product's real Python wrappers, autograd Functions, `train_forward.level_loop`
and `ucnerf_b200.models.Model`, with the C-ABI calls served by the serial CPU instantiation of the same algorithm
templates (tests/cpu_harness.cpp), against the reference's vectors - including the reference's own
Model.forward(rand=True) forward + backward.  It checks the host side and the shared algorithm code; the CUDA
instantiation is what the real `-m gpu` run checks."""
import os
import subprocess
import sys

from conftest import ROOT


def test_training_gpu_tests_pass_when_rehearsed_on_cpu(harness):
    env = dict(os.environ, PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dryrun_train_gpu_tests_on_cpu.py")], env=env,
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "ALL DRY RUNS OK" in res.stdout
    assert res.stdout.count("\n  ok ") >= 24
