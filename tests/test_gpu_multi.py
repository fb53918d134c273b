"""Two-GPU test (skipped unless >= 2 CUDA devices): tools/multigpu_check.py under torchrun -- distributed outputs must
equal the single-GPU outputs bitwise (gather exchange) / to 1e-6 (all-reduce of partials), incl. contigs that straddle ranks."""
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_two_gpu_parity(repo_root):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(repo_root / "tools" / "multigpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "gather bitwise-equal: True" in r.stdout and "module driver under torchrun x2 (gather): NPZ equals" in r.stdout
    assert "(allreduce): max |d|" in r.stdout
