"""-m gpu: multi-GPU parity (one process per GPU over NCCL) -- runs only where >= 2 GPUs are visible; the single-GPU round-end
box skips it.  tests/multi_gpu_check.py is the same check launched by hand with torchrun (see profiles/r1_multi_gpu_parity.txt)."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("pc,extra", [(0, []), (1, ["--dim", "2", "--nel", "8", "--degree", "4", "--reduction", "3"]), (1, ["--dim", "3", "--nel", "4", "--degree", "3", "--reduction", "2"]),
                                      # BASELINE configs[3] (c4): mixed orders -- ladder 9/6/3/1; 6 element layers per rank, so that the rings at degrees 9, 6, 3, 1, the extended
                                      # N = 1 elements and a coarsened superdomain all exist, with non-conforming faces between every pair of degrees
                                      (1, ["--dim", "3", "--nel", "12,4,4", "--degree", "9", "--reduction", "3", "--eps", "0.03"])])
def test_two_gpu_parity(pc, extra):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(29600 + pc + len(extra)),
           os.path.join(ROOT, "tests", "multi_gpu_check.py"), "--pc", str(pc)] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "MULTI_GPU_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
