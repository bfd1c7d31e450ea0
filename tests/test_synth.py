"""CPU: the benchmark's seed generator and the plane fingerprint (clap_b200/synth.py) -- numpy and torch twins agree,
any sub-volume is generated independently of the rest, and the committed anchor of the benched volume is coherent."""
import json
import os

import numpy as np
import pytest
import torch

from clap_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("shape", [(37, 19, 3, 9), (64, 64, 0, 5), (2048, 3, 2047, 2048), (1, 1, 0, 1)])
def test_numpy_and_torch_generators_agree(shape):
    d0, d1, z0, z1 = shape
    a = synth.synth_numpy(np, d0, d1, z0, z1)
    b = synth.synth_torch(torch, d0, d1, z0, z1, "cpu", chunk_planes=3).numpy()
    assert a.shape == (z1 - z0, d1, d0) and np.array_equal(a, b)
    assert a.max() <= 5


def test_sub_volumes_are_position_keyed():
    whole = synth.synth_numpy(np, 40, 30, 0, 12)
    assert np.array_equal(synth.synth_numpy(np, 40, 30, 5, 9), whole[5:9])
    assert not np.array_equal(synth.synth_numpy(np, 40, 30, 0, 12, seed=1), whole)


def test_density_and_value_distribution():
    v = synth.synth_numpy(np, 512, 512, 100, 104)
    assert abs((v != 0).mean() - 0.25) < 0.003
    counts = np.bincount(v.ravel(), minlength=6)[1:] / (v != 0).sum()
    assert np.all(np.abs(counts - 0.2) < 0.005)


def test_plane_hash_sees_position_and_value():
    v = synth.synth_numpy(np, 45, 7, 0, 3)
    h = synth.plane_hashes_numpy(np, v)
    assert h.dtype == np.uint64 and len(set(h.tolist())) == 3
    w = v.copy()
    w[1, 3, 4], w[1, 3, 5] = v[1, 3, 5], v[1, 3, 4]
    if v[1, 3, 4] != v[1, 3, 5]:
        assert synth.plane_hashes_numpy(np, w)[1] != h[1]           # swapped neighbours
    assert np.array_equal(synth.plane_hashes_numpy(np, w)[[0, 2]], h[[0, 2]])
    # reference value: splitmix64 of the single word (cell 0 = 1) keyed with k = 1
    one = np.zeros((1, 1, 8), np.uint8)
    one[0, 0, 0] = 1
    x = ((1 ^ 0x9E3779B97F4A7C15) + 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
    assert int(synth.plane_hashes_numpy(np, one)[0]) == x ^ (x >> 31)


def test_committed_anchor_of_the_benched_volume_matches_the_generator():
    """tests/golden/cfg4_planes_2048.json (the unmodified reference on the bottom 58 planes of the benched volume):
    the seed planes it was computed from are the planes the generator produces today"""
    path = os.path.join(G, "cfg4_planes_2048.json")
    if not os.path.exists(path):
        pytest.skip("anchor not generated")
    with open(path) as f:
        a = json.load(f)
    seed = synth.synth_numpy(np, a["side"], a["side"], 0, 2)
    assert ["%016x" % int(h) for h in synth.plane_hashes_numpy(np, seed)] == a["seed_plane_hashes"][:2]
    assert len(a["plane_hashes"]) == a["planes"] == 8 and a["generations"] == 50 and a["nca"] == 7
