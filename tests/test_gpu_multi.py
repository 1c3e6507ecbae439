"""N > 1 on real GPUs: the SyncGraphGroup exchanges (peer-memory kernels over NVLink and NCCL) against
the CPU oracle's SyncGraphGroup on the same split batches - tests/multi_gpu_worker.py under torchrun.
Skipped on a box with a single GPU (the CPU tier covers the host logic with gloo, tests/test_sync_gloo.py);
run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`.
"""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch

    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("mode", [2, 4], ids=["bf16x3", "bf16-shadows"])
def test_sync_exchanges_match_oracle(world, mode):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, MRN_TEST_GEMM_MODE=str(mode), OMP_NUM_THREADS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(lines[-1])
    print(json.dumps(res))
    assert r.returncode == 0 and res["ok"], res
    for name in ("peer", "nccl"):
        assert res["exchanges"][name]["replicas_identical"], name
    assert res["exchanges"]["peer"]["used_peer_exchange"]


@pytest.mark.parametrize("world", [2])
def test_async_group_on_two_gpus(world):
    """AsyncGraphGroup with N > 1: every rank trains without synchronising with the others - parameter
    fetches and gradient pushes go through per-shard device locks over peer memory (csrc/kernels/exchange.cu), one rank
    is deliberately slow.  scripts/async_check.py checks finite, falling costs and that after a final barrier + fetch
    every rank holds the same parameters (= the master shards)."""
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "async_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="4"))
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(lines[-1])
    print(json.dumps(res))
    assert r.returncode == 0 and res["ok"] and res["replicas_identical_after_fetch"], res
