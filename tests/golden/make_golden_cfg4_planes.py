"""Parity anchor for the BENCHED volume of BASELINE config 4 (chain links (ii)/(iii)): the unmodified
reference (oracle/_ref) run on the bottom slab of the very 2048 x 2048 x 2048 seed volume bench.py uses
(clap_b200/synth.py, seed 0xC1A9), for the full 50 generations.

Plane z of generation g depends on plane z+1 of generation g-1 and on nothing further up, so after G generations
planes 0 .. K-1 of the full volume are determined by planes 0 .. K-1+G of the seed: running the reference on a
slab of K + G planes (zero above, exactly as far as planes < K can tell) gives the true planes 0 .. K-1.  With
K = 8, G = 50: 58 planes = 2.4e8 cells (fits the reference's 32-bit index), ~4 minutes on one core.  The result is
committed as per-plane fingerprints (clapca_hash_planes / synth.plane_hashes_numpy); bench.py and the GPU tests
compare the corresponding planes of the single-GPU and of the sharded runs with them.

    python tests/golden/make_golden_cfg4_planes.py [side]
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle_lib  # noqa: E402
from clap_b200 import synth  # noqa: E402

K, G, RULE = 8, 50, 7


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    ref = oracle_lib.ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libclapref.so missing: make -C oracle ref (needs /root/reference)")
    vol = synth.synth_numpy(np, side, side, 0, K + G)
    seed_hashes = synth.plane_hashes_numpy(np, vol[:K])
    t0 = time.time()
    ref.ca3d_run(vol, RULE, G)
    dt = time.time() - t0
    out = {
        "what": f"planes 0..{K - 1} of ca3d_run(ca_coral, {G} generations) on the {side} x {side} x {side} synthetic "
                f"volume of clap_b200/synth.py (seed 0x{synth.SEED:X}), computed by the unmodified reference on the "
                f"bottom {K + G} planes",
        "side": side, "planes": K, "generations": G, "nca": RULE, "reference_seconds": round(dt, 1),
        "seed_plane_hashes": ["%016x" % int(h) for h in seed_hashes],
        "plane_hashes": ["%016x" % int(h) for h in synth.plane_hashes_numpy(np, vol[:K])],
        "plane_populations": [int(np.count_nonzero(vol[z])) for z in range(K)],
    }
    name = "cfg4_planes_%d.json" % side
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(name, out["plane_hashes"], dt)


if __name__ == "__main__":
    main()
