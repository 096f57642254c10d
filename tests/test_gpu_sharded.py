"""Multi-GPU parity: the sharded register (NCCL exchanges) against the CPU oracle, one process per GPU.
Needs >= 2 GPUs on the box (gpurun --gpus 2); skipped otherwise."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_matches_oracle(engine, world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_port()), os.path.join(ROOT, "tests", "sharded_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    tail = (out.stdout + out.stderr)[-4000:]
    assert out.returncode == 0 and "SHARDED_OK" in out.stdout, tail


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_remote_passes_match_oracle(engine, world):
    """Experimental exchange-free scheduler (QB_SHARD_REMOTE=1): rank bits as tile bits, peers' amplitudes loaded and stored by
    the tile kernel itself (k2r_tile_pass_remote).  CPU-checked only so far; runs with QB_TEST_EXPERIMENTAL=1."""
    if os.environ.get("QB_TEST_EXPERIMENTAL") != "1":
        pytest.skip("experimental kernel: set QB_TEST_EXPERIMENTAL=1")
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_port()), os.path.join(ROOT, "tests", "sharded_worker.py")]
    env = dict(os.environ, QB_SHARD_REMOTE="1", QB_SHARDED_QUICK="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    tail = (out.stdout + out.stderr)[-4000:]
    assert out.returncode == 0 and "SHARDED_OK" in out.stdout, tail
    passes = [int(ln.split(":")[1]) for ln in out.stdout.splitlines() if ln.startswith("remote passes:")]
    assert passes and passes[0] > 0, tail
