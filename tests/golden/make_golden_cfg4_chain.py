"""Link (i) of the BASELINE config-4 parity chain: the unmodified reference (oracle/_ref) on a 256^3
volume for the FULL 50 generations -- 14 s per rule on one core, too slow for the suite, so pinned here once.
Input: the synthetic seed of the tests (P(alive) = 1/4, values 1..5) from numpy's default_rng(2048).

    python tests/golden/make_golden_cfg4_chain.py
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402


def chain_seed(side=256, seed=2048):
    rng = np.random.default_rng(seed)
    return (rng.integers(1, 6, (side, side, side)) * (rng.random((side, side, side)) < 0.25)).astype(np.uint8)


def main():
    ref, port = oracle_lib.ref(), oracle_lib.port()
    if ref is None:
        raise SystemExit("oracle/_ref/libclapref.so missing: make -C oracle ref (needs /root/reference)")
    vol0 = chain_seed()
    out = {"seed_fnv1a64": "%016x" % port.fnv(vol0), "seed_population": int(np.count_nonzero(vol0))}
    for nca, name in ((7, "ca_coral"), (0, "ca_445m")):
        vol = vol0.copy()
        t0 = time.time()
        pop = ref.ca3d_run(vol, nca, 50)
        out[name] = {"nca": nca, "generations": 50, "population": int(pop), "fnv1a64": "%016x" % port.fnv(vol),
                     "reference_seconds": round(time.time() - t0, 1)}
        print(name, out[name], flush=True)
    with open(os.path.join(HERE, "cfg4_chain_256.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
