"""Multi-GPU parity in the GPU suite (-m gpu): when the box has at least two devices, spawn one NCCL rank per GPU
(two of them) and check that the sharded paths - MSEStep over row bands, render_bands, ShardedBatchStep - reproduce
the single-GPU results (tests/dist_worker.py).  Skips cleanly on a one-GPU box."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_two_rank_sharded_paths_equal_the_single_gpu_results():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs at least two GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'dist_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-2000:])
    rep = json.loads(lines[-1])
    assert rep['ok_all_ranks'] and rep['world'] == 2
    assert rep['band_step_image_bit_identical'] and rep['render_bands_bit_identical'] and rep['sharded_batch_images_bit_identical']
    print(rep)
