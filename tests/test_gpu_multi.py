"""Multi-GPU tests (-m gpu): skipped on a box with fewer GPUs than the test needs (the driver's test box has
one); run by the builder with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`.

The sharded, streamed path under real NCCL (torchrun, one process per GPU): contiguous frame shards, the
last-seen exchange (R3:277,314), NCCL gather of the per-frame records and plane tilts - the gathered result
must be byte-identical to one sequential run (tests/multi_worker.py does the comparison on rank 0)."""
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
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("world,n_frames,batch", [(2, 37, 8), (2, 24, 5), (2, 40, 8), (4, 41, 4), (4, 48, 5)])
def test_streamed_shards_under_nccl_equal_one_sequential_run(world, n_frames, batch):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_worker.py"), str(n_frames), str(batch)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and lines, (p.returncode, p.stdout[-2000:], p.stderr[-4000:])
    rep = json.loads(lines[-1])
    assert rep["ok"] and rep["world"] == world and all(rep["report"][k] for k in ("pos3d", "pos_flags", "row_det", "plane")), rep
