"""CPU: host-side arithmetic that the GPU path relies on -- the lrand48() jump-ahead used by the device-side
seeding of ca2d_generate() (clap_b200/csrc/ca2d_layout.cuh, compiled here for the HOST with nvcc and checked
against step-by-step iteration and against glibc itself), and the z-block size rule of the sharded bench."""
import os
import shutil
import subprocess

import pytest

from clap_b200.slab import default_block_planes, local_planes, plan_blocks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

JUMP_SRC = r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "%s/clap_b200/csrc/ca2d_layout.cuh"
using namespace clapca;
int main()
{
    const unsigned long long M = (1ull << 48) - 1;
    int bad = 0;
    /* against step-by-step iteration from several states */
    const unsigned long long starts[3] = { 0x1234ABCD330Eull, 0ull, M };
    for (int s = 0; s < 3; s++) {
        unsigned long long y = starts[s];
        for (unsigned long long n = 1; n <= 70000; n++) {
            y = (y * 0x5DEECE66Dull + 0xBull) & M;
            if ((n %% 611 == 0 || n < 40) && r48_advance(starts[s], n) != y) bad++;
        }
    }
    /* jumps compose, also beyond 2^32 draws */
    const unsigned long long a = 268435456ull, b = 5000000019ull;
    if (r48_advance(r48_advance(starts[0], a), b) != r48_advance(starts[0], a + b)) bad++;
    if (r48_advance(starts[0], 0) != starts[0]) bad++;
    /* against glibc: srand48(seed); n x lrand48() */
    srand48(1234);
    for (int i = 0; i < 1000; i++) lrand48();
    unsigned long long x0 = ((1234ull & 0xffffffffull) << 16) | 0x330Eull;
    if ((unsigned long long)lrand48() != (r48_advance(x0, 1001) >> 17)) bad++;
    printf("bad %%d\n", bad);
    return bad != 0;
}
'''


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_rand48_jump_ahead_matches_iteration_and_glibc(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = tmp_path / "jump.cu"
    src.write_text(JUMP_SRC % ROOT)
    exe = tmp_path / "jump"
    subprocess.run([nvcc, "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(exe), str(src)],
                   check=True, capture_output=True, timeout=300)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "bad 0", r.stdout + r.stderr


def test_python_rand48_block_matches_scalar_stream():
    from clap_b200.ca import Rand48
    a, b = Rand48(77), Rand48(77)
    block = a.lrand48_block(5000)
    assert [int(v) for v in block[:200]] == [b.lrand48() for _ in range(200)]
    for _ in range(4800):
        b.lrand48()
    assert a.x == b.x


@pytest.mark.parametrize("d2", [1, 7, 64, 100, 1000, 2048, 4096])
@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
def test_default_block_planes(d2, nranks):
    b = default_block_planes(d2, nranks)
    per_rank = -(-d2 // nranks)
    assert 1 <= b <= d2
    blocks = plan_blocks(d2, nranks, b)
    assert sum(z1 - z0 for _, z0, z1 in blocks) == d2
    if nranks == 1:
        assert b == d2
        return
    assert b <= 48 and b <= per_rank
    loads = [sum(z1 - z0 for r, z0, z1 in blocks if r == k) for k in range(nranks)]
    if per_rank >= 5 * 6:               # room for five whole-tile blocks per rank: balanced within 15 %
        assert max(loads) <= 1.15 * d2 / nranks + 6, (b, loads)
        assert b % 6 == 0               # whole tiles
        assert min(sum(1 for r, _, _ in blocks if r == k) for k in range(nranks)) >= 2


def test_default_block_planes_of_the_scaling_bench():
    assert default_block_planes(2048, 2) == 48
    assert default_block_planes(2048, 4) == 48         # measured: 28.2 ms against 29.6 (42) and 30.5 (36)
    assert default_block_planes(2048, 8) == 48         # the measured optimum (profiles/r02_knobs_multi_wide_n8.txt)


def test_numa_binding_is_advisory_without_a_gpu():
    """bind_to_gpu_numa_node() must never fail a run: without NVML / a device it changes nothing and returns None"""
    from clap_b200.slab import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    got = bind_to_gpu_numa_node(0)
    assert got is None or got <= before
    if got is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
