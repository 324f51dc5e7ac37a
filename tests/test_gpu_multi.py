"""Hardware multi-GPU parity of the sharded hypothesis sweep ("comparing to the 1-GPU result bit for bit").
Needs >= 2 visible GPUs (`gpurun --gpus N -- python -m pytest tests/test_gpu_multi.py -m gpu`); on a single-GPU box only the 1-rank
NCCL path runs (tests/test_gpu_at_size.py covers it at C4 size as well)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "multi_gpu_worker.py")


def _ngpu():
    import torch

    return torch.cuda.device_count()


def test_single_process_comm_all_visible_gpus():
    r = subprocess.run([sys.executable, WORKER, "--single"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "ok single-process" in r.stdout


@pytest.mark.parametrize("world", [2, 4, 8])
def test_torchrun_sharded_sweep_equals_single_rank(world):
    if _ngpu() < world:
        pytest.skip(f"{world} GPUs needed, {_ngpu()} visible")
    port = 29700 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok rank") == world
