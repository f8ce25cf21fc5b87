"""Data-parallel correctness on real GPUs (needs >= 2 devices: `gpurun --gpus 2 -- python -m pytest tests -m gpu`).
Replaces the reference's nn.DataParallel wrap (train.py:517); see tests/dp_worker.py for what is checked."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("transport", ["peer", "peer-p2p", "nccl"])
def test_two_rank_graph_step_matches_one_gpu_global_batch(cuda, transport):
    """transport "peer": libvqacore's NVLink peer-memory all-reduce (vqa_peer_allreduce_f32; NVLS in-switch reduction
    when the fabric has a multicast mapping), "peer-p2p": the same with plain peer loads / stores, "nccl": dist.all_reduce."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", {"peer": "29731", "peer-p2p": "29735", "nccl": "29733"}[transport], os.path.join(ROOT, "tests", "dp_worker.py")]
    env = dict(os.environ, VQA_ALLREDUCE=transport.split("-")[0], VQA_PEER_SPIN_MS="20000",
               VQA_PEER_MC="0" if transport == "peer-p2p" else "1")      # "peer": force the NVLS path even at 2 ranks
    transport = transport.split("-")[0]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and "DP_OK transport %s" % transport in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
