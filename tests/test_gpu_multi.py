"""GPU (-m gpu): the sharded path.  With one GPU the single-rank slab path is checked against the oracle; with
two or more GPUs on the box the z-block decomposition with peer ghost planes must equal the single-GPU result
bit for bit (parity chain of SURVEY.md 8(d) cfg 4: GPU == oracle at small sizes, P-GPU == 1-GPU beyond)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_single_rank_slab_equals_oracle(gpu, oracle):
    from clap_b200.slab import ShardedVolume
    rng = np.random.default_rng(31)
    vol = (rng.integers(1, 6, (12, 20, 70)) * (rng.random((12, 20, 70)) < 0.3)).astype(np.uint8)
    want = vol.copy()
    s, b, n = oracle.ca3d_rule(7)
    wpop = oracle.ca3d_run(want, s, b, n, 5)
    sv = ShardedVolume(70, 20, 12, 0, 1, 5, int(vol.max()))
    sv.upload(vol)
    sv.prepare(7, 5)
    assert sv.run() == wpop
    got = np.empty_like(vol)
    sv.download(got)
    sv.close()
    assert np.array_equal(got, want)


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("shape", [(70, 24, 20, 6, 7, 2), (256, 64, 48, 8, 0, 4), (2048, 128, 64, 5, 7, 8)])
def test_sharded_equals_single_gpu(shape):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs at least two GPUs on the box")
    world = min(n, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(ROOT, "tests", "multi_gpu_worker.py")] + [str(v) for v in shape]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
