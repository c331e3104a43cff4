"""N > 1 on real GPUs: runs tests/dist_worker.py under torch.distributed.run with one rank per GPU (NCCL).  Needs >= 2
visible GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`); skipped on a one-GPU box.  The host-side
logic of the round sync is covered on CPU (gloo, world_size 2) in tests/test_host_logic.py."""
import os
import subprocess
import sys

import pytest
import torch as th

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(th.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_trainer_global_batch_discriminator_and_round_sync():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    sys.stderr.write(r.stdout[-3000:] + r.stderr[-6000:])
    assert r.returncode == 0 and "DIST_OK" in r.stdout


@pytest.mark.skipif(th.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_member_parallel_ensemble_is_bit_identical_to_one_process():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29527", os.path.join(ROOT, "tests", "dist_pref_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    sys.stderr.write(r.stdout[-3000:] + r.stderr[-6000:])
    assert r.returncode == 0 and "DIST_PREF_OK" in r.stdout
