"""Multi-GPU parity (one process per GPU, NCCL halo exchange + Gram all-reduce) and the CPU-side world-size-2
check of the partition/halo host logic over gloo."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _torchrun(script, nproc, timeout=600, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", script)]
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=e)


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["nvlink-peer", "nccl"])
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_multi_gpu_parity(nproc, transport):
    """both halo transports: the fused pack->peer-store / wait->unpack kernels over NVLink peer memory (default) and
    the NCCL send/recv fallback (HXB200_HALO=nccl)"""
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    env = {"HXB200_EXPECT_TRANSPORT": transport}
    if transport == "nccl":
        env["HXB200_HALO"] = "nccl"
    r = _torchrun("mgpu_worker.py", nproc, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    print(r.stdout[-1500:])


def test_partition_halo_logic_gloo_world2():
    """Two CPU processes (gloo) each build ONLY their partition, exchange halos with torch.distributed
    following the MPIPatternP2P lists, and must reproduce the single-process oracle world."""
    r = _torchrun("gloo_worker.py", 2, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
