"""GPU data-parallel parity (needs >= 2 GPUs on the box: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`).
Two ranks x B/2 against the sharded CPU oracle (local BatchNorm statistics per rank, averaged gradients): see
tests/dist_parity_worker.py.  On a single-GPU box the tests are skipped (the gloo world-2 tests of
test_host_logic.py cover the host logic there)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('precision,mode', [('fp32', 'graph'), ('fp32', 'eager'), ('bf16', 'graph')])
def test_world2_matches_sharded_oracle(precision, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    port = 29600 + (os.getpid() + hash((precision, mode))) % 300
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'tests', 'dist_parity_worker.py'), precision, mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stderr[-4000:]
    assert 'DIST_PARITY_OK' in r.stdout, r.stdout[-4000:]
