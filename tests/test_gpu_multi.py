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


# ---- the sharded path on ONE device: several ranks of one process, each launch on its share of the SMs ------------
# (ghost planes, peer pushes by the tiles' service warps, banked counters -- everything but the NVLink hop itself)

@pytest.mark.parametrize("case", [
    # d0   d1   d2  gens rule ranks block
    (70, 24, 20, 6, 7, 2, 2),           # z-blocks smaller than a tile: every tile has two edges
    (70, 24, 20, 6, 7, 2, 10),          # contiguous slabs
    (256, 64, 48, 8, 0, 4, 4),          # 4 ranks, one tile row per block
    (256, 64, 50, 9, 7, 3, 16),         # ragged last block, ragged last generation group
    (2048, 96, 64, 8, 7, 4, 16),        # BASELINE config-4 row width, the bench's block size
    (1000, 40, 33, 5, 2, 2, 4),         # 4 state planes (pyroclastic), ragged rows
    (300, 30, 12, 4, 7, 4, 1),          # single-plane blocks: every plane is an edge on both sides
])
@pytest.mark.parametrize("halo", ["ldst", "tma"])
def test_local_ranks_equal_single_gpu_and_oracle(gpu, oracle, monkeypatch, case, halo):
    """both halo-row paths of the service warp: ld / st batches (default) and TMA bulk copies (CLAPCA_HALO_LDST=0)"""
    from clap_b200.slab import LocalRanks
    d0, d1, d2, gens, rule, ranks, block = case
    if halo == "tma":
        monkeypatch.setenv("CLAPCA_HALO_LDST", "0")
    rng = np.random.default_rng(d0 * 7 + d2)
    full = (rng.integers(1, 6, (d2, d1, d0)) * (rng.random((d2, d1, d0)) < 0.3)).astype(np.uint8)
    want = full.copy()
    wpop = gpu.ca3d_run(want, rule, gens)
    if full.size <= 1 << 22:
        chk = full.copy()
        s, b, n = oracle.ca3d_rule(rule)
        assert oracle.ca3d_run(chk, s, b, n, gens) == wpop and np.array_equal(chk, want)
    # max_value covers the cells AND the value a born cell takes (pyroclastic: 9 -> 4 state planes)
    lr = LocalRanks(d0, d1, d2, ranks, gens, max(int(full.max()), gpu.ca3d_rule(rule).nr_states - 1), block)
    try:
        for rep in range(3):                # odd and even runs use the two banks of ghost counters
            lr.upload(full)
            assert lr.run(rule, gens) == wpop, f"rep {rep}: population"
            got = lr.download()
            assert np.array_equal(got, want), f"rep {rep}: {int((got != want).sum())} cells differ"
        from clap_b200 import synth
        assert np.array_equal(lr.plane_hashes(), synth.plane_hashes_numpy(np, want))
    finally:
        lr.close()


def test_local_ranks_ca3d_make_volume_with_255s(gpu, oracle):
    """the reference's own seed (ca3d_make leaves 255s) through the sharded path: 8 state planes"""
    from clap_b200.slab import LocalRanks
    full = oracle.ca3d_make(48, 24, 20, 42)
    assert full.max() == 255
    want = full.copy()
    s, b, n = oracle.ca3d_rule(7)
    wpop = oracle.ca3d_run(want, s, b, n, 6)
    lr = LocalRanks(48, 24, 20, 2, 6, 255, 4)
    try:
        lr.upload(full)
        assert lr.run(7, 6) == wpop
        assert np.array_equal(lr.download(), want)
    finally:
        lr.close()


@pytest.mark.parametrize("case", [
    # d0   d1   d2  gens rule ranks block
    (70, 24, 20, 6, 7, 2, 5),
    (256, 64, 50, 9, 7, 3, 10),         # ragged last block / generation group
    (2048, 96, 60, 8, 7, 4, 5),         # BASELINE config-4 row width, every tile has two z-block edges
    (1000, 40, 33, 5, 2, 2, 10),        # 4 state planes
])
def test_local_ranks_streamed_host_to_host_equals_resident(gpu, oracle, monkeypatch, case):
    """clapca_slab_run_streamed: per rank, upload / pack / every generation / unpack / download as one pipeline
    (layout items + the halo seed pushed by the pack items' service warps), against the single-GPU result"""
    import torch
    from clap_b200 import ClapcaError
    from clap_b200.slab import LocalRanks
    d0, d1, d2, gens, rule, ranks, block = case
    monkeypatch.setenv("CLAPCA_IO_CHUNK_PLANES", "3")
    rng = np.random.default_rng(d0 + d2)
    full = (rng.integers(1, 6, (d2, d1, d0)) * (rng.random((d2, d1, d0)) < 0.3)).astype(np.uint8)
    want = full.copy()
    wpop = gpu.ca3d_run(want, rule, gens)
    maxv = max(int(full.max()), gpu.ca3d_rule(rule).nr_states - 1)
    lr = LocalRanks(d0, d1, d2, ranks, gens, maxv, block)
    try:
        keep, hin, hout = [], [], []
        for vol in lr.ranks:
            n = max(1, vol.n_local) * d1 * d0
            a, b = torch.empty(n, dtype=torch.uint8, pin_memory=True), torch.empty(n, dtype=torch.uint8, pin_memory=True)
            keep += [a, b]
            hin.append(a.numpy())
            hout.append(b.numpy())
            if vol.n_local:
                hin[-1][:vol.n_local * d1 * d0] = full[vol.zglobal].reshape(-1)
        for rep in range(2):
            for b in hout:
                b[...] = 0xEE
            assert lr.run_streamed(rule, gens, hin, hout) == wpop, rep
            got = np.zeros_like(full)
            for vol, b in zip(lr.ranks, hout):
                if vol.n_local:
                    got[vol.zglobal] = b[:vol.n_local * d1 * d0].reshape(vol.n_local, d1, d0)
            assert np.array_equal(got, want), (rep, int((got != want).sum()))
        # a resident run on the same slabs afterwards (the claim order is rebuilt without layout items)
        lr.upload(full)
        assert lr.run(rule, gens) == wpop
        assert np.array_equal(lr.download(), want)
        # the bound given at create time is enforced by the pack items
        if maxv < 255 and d0 == 70:        # (the other ranks only notice through their watchdogs: once is enough)
            hin[0][0] = 200
            with pytest.raises(ClapcaError):
                lr.run_streamed(rule, gens, hin, hout)
    finally:
        lr.close()


def test_slab_prepare_rejects_cells_above_max_value(gpu):
    """the slab path must verify the caller's max_value (the pack kernel would drop the upper bits)"""
    from clap_b200 import ClapcaError
    from clap_b200.slab import ShardedVolume
    vol = np.zeros((4, 8, 40), np.uint8)
    vol[1, 2, 3] = 9
    sv = ShardedVolume(40, 8, 4, 0, 1, 3, 5)
    try:
        sv.upload(vol)
        with pytest.raises(ClapcaError):
            sv.prepare(7, 3)
    finally:
        sv.close()


@pytest.mark.slow
def test_local_ranks_on_a_slab_of_the_benched_volume(gpu):
    """2048 x 2048 rows of the benched seed volume (clap_b200/synth.py), 64 planes, 4 ranks in blocks of 16 planes
    against the single-GPU run: the configuration SCALE runs, at a depth one device holds four times over."""
    import torch
    from clap_b200 import synth
    from clap_b200.slab import LocalRanks
    d0 = d1 = 2048
    d2, gens = 64, 10
    full = synth.synth_torch(torch, d0, d1, 0, d2, "cuda:0").cpu().numpy()
    want = full.copy()
    wpop = gpu.ca3d_run(want, 7, gens)
    lr = LocalRanks(d0, d1, d2, 4, gens, 5, 16)
    try:
        lr.upload(full)
        assert lr.run(7, gens) == wpop
        assert np.array_equal(lr.plane_hashes(), synth.plane_hashes_numpy(np, want))
    finally:
        lr.close()


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
