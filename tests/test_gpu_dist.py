"""N > 1 parity under pytest: spawns one rank per GPU (2 ranks) and runs tests/dist_exchange_check.py - the hash exchange through
every transport (NCCL send/recv, SM stores into peer arenas, fenced two-context pipeline, split-phase copy-engine form) against the
oracle's partition function (M/operator/HashGenerator.java:41-46, M/operator/output/PagePartitioner.java:133-162), and the partitioned
join against the oracle join.  Skipped on a box with fewer than two GPUs."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    from trino_b200 import abi
    return abi.load_library().tgpu_device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(script, world, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, script), *args]
    return subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)


def test_two_rank_exchange_and_partitioned_join_match_oracle():
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    r = _torchrun("tests/dist_exchange_check.py", 2)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "dist_exchange_check ok" in r.stdout


def test_two_rank_workloads_match_oracle():
    # configs[3] / configs[4] / multi-GPU Q1 pipelines and the broadcast exchange on 2 ranks
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    r = _torchrun("tests/dist_workloads_check.py", 2)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "dist_workloads_check ok" in r.stdout
