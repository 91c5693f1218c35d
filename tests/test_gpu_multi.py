"""Multi-GPU paths over real NCCL ranks (self-skipping on a box with fewer than 2 GPUs): slab-sharded
volume and frame-sharded series through the public `segmentation_loop`.  The emulated-rank tests
(tests/test_gpu_slab.py) and the gloo tests (tests/test_host_logic.py) cover the same logic on one
device / on the CPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize('world', [2, 4])
def test_public_loop_over_nccl_ranks(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs, this box has {torch.cuda.device_count()}')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
                        '--master-addr', '127.0.0.1', '--master-port', str(29620 + world),
                        os.path.join(HERE, 'tools', 'nccl_worker.py'), str(tmp_path)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for k in range(world):
        assert f'rank {k} ok' in r.stdout
