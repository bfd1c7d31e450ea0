"""BASELINE config 3 at FULL size, from the unmodified reference (oracle/_ref/libclapref.so):

    srand48(1); ca2d_generate({born 5..8, surv 4..8, nr_states 1, decay, m1}, 16384, 100)      # core/ca2d.c:79-98

One core, about a quarter of an hour -- far too slow for the test suite, so the result is pinned here once as an
FNV-1a-64 fingerprint (+ population and a few row fingerprints after the seed and after 100 generations) in
cfg3_16384.json; tests/test_gpu_parity.py::test_ca2d_cfg3_full_size_reference_fingerprint compares the GPU run
against it bit for bit.  Also pins the multi-state ca_test rule at 4096^2 x 100 (about a minute).

    python tests/golden/make_golden_cfg3.py
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402


def fingerprint(ref_fnv, arr):
    side = arr.shape[0]
    rows = [0, 1, side // 3, side - 1]
    return {"fnv1a64": "%016x" % ref_fnv(arr), "population": int(np.count_nonzero(arr)), "sum": int(arr.sum(dtype=np.int64)),
            "rows": {str(r): "%016x" % ref_fnv(arr[r]) for r in rows}}


def main():
    ref, port = oracle_lib.ref(), oracle_lib.port()
    if ref is None:
        raise SystemExit("oracle/_ref/libclapref.so missing: make -C oracle ref (needs /root/reference)")
    out = {}
    cases = [("cave_bin_16384_x100_seed1", dict(born=0x1E0, surv=0x1F0, nr=1, decay=1, neigh=oracle_lib.NEIGH_M1), 16384, 100, 1),
             ("ca_test_4096_x100_seed7", dict(born=3 << 2, surv=3 << 7, nr=4, decay=1, neigh=oracle_lib.NEIGH_M1), 4096, 100, 7)]
    for name, ca, side, steps, seed in cases[::-1]:
        t0 = time.time()
        seeded = ref.ca2d_generate(side=side, steps=0, seed=seed, **ca)
        final = ref.ca2d_generate(side=side, steps=steps, seed=seed, **ca)
        out[name] = {"rule": {k: int(v) for k, v in ca.items()}, "side": side, "steps": steps, "srand48": seed,
                     "seed_grid": fingerprint(port.fnv, seeded), "final_grid": fingerprint(port.fnv, final),
                     "reference_seconds": round(time.time() - t0, 1)}
        print(name, out[name]["final_grid"]["fnv1a64"], out[name]["reference_seconds"], flush=True)
        with open(os.path.join(HERE, "cfg3_16384.json"), "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
