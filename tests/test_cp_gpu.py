"""Context parallelism on real GPUs (needs >= 2 GPUs with peer access on one node; skipped on a single-GPU box): the views of one
scene sharded over 2 ranks must give bit-identical predictions to the single-GPU forward, eagerly and under CUDA-graph replay."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_context_parallel_two_gpus_bit_identical():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tools", "cp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "CP_CHECK PASS" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
