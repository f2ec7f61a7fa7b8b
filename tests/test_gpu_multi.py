"""Multi-GPU parity over NCCL: needs a box with at least two GPUs (skipped otherwise; `gpurun --gpus N`).
Runs tools/multi_gpu_check.py under torchrun: the default run of BASELINE configs[1] (route.xml, 335,544,240 rays)
shared between the ranks -- whole launches and launches cut into ray ranges -- must give the dose map of the
reference-derived golden, and a smaller run must equal the same process's single-GPU run bit for bit."""
import os
import subprocess
import sys

import pytest

import uvrt_testlib as T

pytestmark = pytest.mark.gpu


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


def test_multi_gpu_run_equals_golden_and_single_gpu():
    n = _gpu_count()
    if n < 2:
        pytest.skip(f"{n} GPU(s) on this box: the NCCL path needs at least two")
    n = min(n, 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(T.ROOT, "tools", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_GPU_OK" in r.stdout, r.stdout[-3000:]
