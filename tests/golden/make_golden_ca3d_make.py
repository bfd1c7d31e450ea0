"""Fingerprints of ca3d_make() (core/ca3d.c:41-99, 144-169) for small volumes, from oracle/_ref/libclapref.so = the
UNMODIFIED reference sources:

    python tests/golden/make_golden_ca3d_make.py        ->  tests/golden/ca3d_make_small.json

Small volumes are where ca3d_prune()'s order-dependent corner shows: an EMPTY cell with six occupied neighbours becomes
255 and counts as occupied for the cells the sweep visits after it (about a quarter of these seeds enclose such a cell).
The fingerprints pin the port (tests/test_oracle_golden.py) and through it clapca_grid_make3d (tests/test_gpu_parity.py).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402

SHAPES = [(6, 6, 6), (8, 8, 8), (7, 9, 8), (10, 10, 10), (5, 5, 5), (12, 6, 9)]
SEEDS = list(range(1, 41))


def main():
    ref = oracle_lib.ref()
    port = oracle_lib.port()
    if ref is None:
        raise SystemExit("oracle/_ref/libclapref.so missing: make -C oracle ref (needs /root/reference)")
    out = {"shapes": SHAPES, "seeds": SEEDS, "fnv1a64": {}, "cells_255_inside": 0}
    for d0, d1, d2 in SHAPES:
        hashes = []
        for seed in SEEDS:
            vol = ref.ca3d_make(d0, d1, d2, seed)
            hashes.append("%016x" % port.fnv(vol))
            out["cells_255_inside"] += int((vol[1:-1, 1:-1, 1:-1] == 255).sum())
        out["fnv1a64"]["%dx%dx%d" % (d0, d1, d2)] = hashes
    with open(os.path.join(HERE, "ca3d_make_small.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote ca3d_make_small.json:", sum(len(v) for v in out["fnv1a64"].values()), "volumes,",
          out["cells_255_inside"], "interior cells marked 255")


if __name__ == "__main__":
    main()
