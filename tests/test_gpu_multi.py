"""Multi-GPU path of the library (NCCL communicator inside the C ABI, border reduction, block-distributed LU, distributed
Newton / OneStep loop).  Needs at least 2 GPUs on the box: skipped otherwise (the 1-GPU CI box runs the one-rank
communicator tests of test_gpu_border.py; the host-side partition logic is covered on CPU by test_multi_rank_gloo.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def torchrun(n, *args):
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
                        "127.0.0.1", "--master-port", "29547", *args], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("graph_partition,p2p", [(0, 1), (1, 1), (1, 0)])
def test_two_gpu_tran_equals_one_gpu_and_the_oracle(graph_partition, p2p):
    """p2p = 1: small collectives through the NVLink peer mailboxes; p2p = 0: the same through NCCL calls"""
    d = torchrun(2, os.path.join(ROOT, "scripts", "multi_gpu_tran.py"), "--rings", "40", "--stages", "31", "--tstop", "3e-10",
                 "--graph-partition", str(graph_partition), "--p2p", str(p2p))
    assert d["n_gpus"] == 2 and d["border_unknowns"] == 2
    assert d["single_gpu"]["identical_step_sequence_and_newton_counts"] and d["single_gpu"]["max_abs_waveform_diff"] < 1e-9
    assert all(r["waveforms_within_reltol_abstol"] and r["newton_counts_equal"] for r in d["oracle_single_ring_replay"])
    assert d["p2p_error"] == 0
