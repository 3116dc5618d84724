"""Slab decomposition with one PROCESS per rank (the production layout: torchrun, CUDA IPC peer buffers): two ranks
on two GPUs when the box has them, otherwise both ranks share the one device (IPC between processes on the same
device; the wait kernels then only make progress through time slicing, so the run is short)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_process_slab_run_matches_single_gpu():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "tests", "slab_dist_worker.py"), "gpu"]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SLAB_DIST_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
