"""Multi-GPU parity (BASELINE config 4 shape): TPC-H Q3 on 2 datanodes with two NCCL
redistributes, checked against the single-node oracle.  Needs >= 2 GPUs."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import ctypes
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return 0
        n = ctypes.c_int()
        return n.value if cu.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_q3_two_datanodes_with_nccl_redistribute():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "q3_multi.py"), "--sf", "1", "--orders", "30000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "OK q3 parity vs oracle" in r.stdout, r.stdout[-2000:]
