"""CPU: host-side logic of the multi-GPU path -- z-block planning, and the IPC handle exchange / slab
reassembly through torch.distributed with the gloo backend (world_size 2, no GPU)."""
import os
import subprocess
import sys

import pytest

from clap_b200.slab import exchange_handles, local_planes, neighbour_ranks, plan_blocks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("d2,nranks,block", [(2048, 8, 16), (2048, 8, 256), (13, 3, 2), (9, 4, 5), (3, 4, 1),
                                             (100, 1, 16), (64, 2, 64)])
def test_plan_is_a_partition_and_blocks_chain_around_the_ring(d2, nranks, block):
    blocks = plan_blocks(d2, nranks, block)
    seen = []
    for j, (r, z0, z1) in enumerate(blocks):
        assert 0 <= r < nranks and z0 < z1 <= d2
        seen.extend(range(z0, z1))
        if j:                       # the block above always lives on the next rank of the ring
            prev_r = blocks[j - 1][0]
            assert r == neighbour_ranks(prev_r, nranks)[0] and blocks[j - 1][2] == z0
    assert seen == list(range(d2))
    allz = sorted(z for r in range(nranks) for z in local_planes(d2, nranks, block, r))
    assert allz == list(range(d2))
    if nranks == 1:
        assert blocks == [(0, 0, d2)]


def test_handle_exchange_picks_the_ring_neighbours():
    handles = [bytes([r]) * 64 for r in range(5)]
    for r in range(5):
        nxt, prv = exchange_handles(handles[r], r, 5, lambda payload: handles)
        assert nxt == handles[(r + 1) % 5] and prv == handles[(r - 1) % 5]
    with pytest.raises(ValueError):
        exchange_handles(handles[0], 0, 5, lambda payload: handles[:3])


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from clap_b200.slab import exchange_handles, local_planes, plan_blocks
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
def all_gather(payload):
    out = [None] * world
    dist.all_gather_object(out, payload)
    return out
mine = bytes([rank + 1]) * 64
nxt, prv = exchange_handles(mine, rank, world, all_gather)
assert nxt == bytes([(rank + 1) % world + 1]) * 64 and prv == bytes([(rank - 1) % world + 1]) * 64
# scatter a volume by the plan, "process" it locally, reassemble, reduce the population
d2, block = 23, 3
full = np.random.default_rng(0).integers(0, 4, (d2, 5, 7)).astype(np.uint8)
zs = local_planes(d2, world, block, rank)
piece = full[zs] + 1
pieces = [None] * world
dist.all_gather_object(pieces, (zs, piece))
got = np.zeros_like(full)
for z, arr in pieces:
    got[z] = arr
assert np.array_equal(got, full + 1)
pop = torch.tensor([int(np.count_nonzero(full[zs]))])
dist.all_reduce(pop)
assert int(pop[0]) == int(np.count_nonzero(full))
dist.barrier()
dist.destroy_process_group()
print("GLOO_OK", rank)
'''


def test_exchange_and_reassembly_over_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29633", str(script), ROOT]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.count("GLOO_OK") == 2, r.stdout[-2000:] + r.stderr[-2000:]
