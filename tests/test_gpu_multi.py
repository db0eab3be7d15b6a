"""Multi-GPU parity under pytest (-m gpu): tests/mgpu_check.py launched with torch.distributed.run on 2 / 4 / 8 GPUs of this
box, one rank per GPU.  Skipped (not failed) where the box has fewer GPUs.  Covers what a single GPU cannot: the halo exchange
over NVLink peer stores (fused one-launch SpMV, the four-launch variant, no overlap) and over NCCL, the rank-ordered dot
all-reduce, back-to-back exchanges (time_kernels), the boxes of add_imbalance, and -- against the UNMODIFIED reference's golden
record (tests/golden/big_*.npz) -- 128^3 at the same np, every array CRC, history and solution bit for bit.
ref/src/exchange_externals.hpp:44-161, ref/src/Vector_functions.hpp:218-222."""
import glob
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_mgpu_check_against_oracle_and_reference_goldens(nproc, oracle_built, product_lib):
    if _gpus() < nproc:
        pytest.skip("needs %d GPUs, this box has %d" % (nproc, _gpus()))
    imb = [p for p in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "mpi_*_np%d_imb*.npz" % nproc)))]
    big = os.path.join(ROOT, "tests", "golden", "big_128x128x128_np%d.npz" % nproc)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(29610 + nproc), os.path.join(ROOT, "tests", "mgpu_check.py"),
           "--sizes", ",".join(["2x2x2", "8x8x8", "5x3x7", "20x30x40", "48x48x48"] + imb), "--big", big if os.path.exists(big) else ""]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1700)
    tail = (out.stdout + out.stderr)[-4000:]
    assert out.returncode == 0 and "MGPU_CHECK OK" in out.stdout, tail
