#!/usr/bin/env python
"""tools/ab_ca2d.py -- A/B of the 2D engines on BASELINE config 3 (16384^2 x 100, binary cave rule) inside ONE process:
the row engine (ca2d_bitplane.cuh) against the diagonal engine (ca2d_skew.cuh) with 1 and 2 words per lane.  Every
variant must leave the SAME grid (compared byte for byte on the device) with the unmodified reference's population
(tests/golden/cfg3_16384.json).  Prints kernel / total milliseconds (CUDA events inside the library) per variant.

    python tools/ab_ca2d.py [side [generations [steps]]]
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clap_b200                                         # noqa: E402
from clap_b200.ca import Rand48                          # noqa: E402
from clap_b200.rules import CellAutomaton                # noqa: E402
from clap_b200._lib import NEIGH_M1                      # noqa: E402


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    gens = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    clap_b200.init(0)
    dev = torch.device("cuda:0")
    ca = CellAutomaton("cave", 0x1E0, 0x1F0, 1, True, NEIGH_M1)
    grid = clap_b200.Grid(side, side, 1)
    grid.seed2d(ca, Rand48(1))
    seed = torch.empty(side * side, dtype=torch.uint8, device=dev)
    grid.download(seed.data_ptr())
    torch.cuda.synchronize()
    ref_pop = None
    if side == 16384 and gens == 100:
        with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cfg3_16384.json")) as f:
            ref_pop = json.load(f)["cave_bin_16384_x100_seed1"]["final_grid"]["population"]
    variants = [("row engine", {"CLAPCA_2D_SKEW": "0"}),
                ("diagonal, 2 words per lane", {"CLAPCA_2D_SKEW": "1", "CLAPCA_2D_SKEW_WPL": "2"}),
                ("diagonal, 1 word per lane", {"CLAPCA_2D_SKEW": "1", "CLAPCA_2D_SKEW_WPL": "1"})]
    first = None
    for name, env in variants:
        for k in ("CLAPCA_2D_SKEW", "CLAPCA_2D_SKEW_WPL"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ker, tot = [], []
        t0 = time.perf_counter()
        try:
            for i in range(2 + steps):
                grid.upload(seed.data_ptr())
                grid.run2d(ca, gens)
                st = grid.stats()
                if i >= 2:
                    ker.append(st["kernel_ms"]); tot.append(st["total_ms"])
        except Exception as e:          # noqa: BLE001 -- keep going: the other variants still say something
            print(f"{name:30s} FAILED: {e}", flush=True)
            continue
        out = torch.empty_like(seed)
        grid.download(out.data_ptr())
        torch.cuda.synchronize()
        pop = int(torch.count_nonzero(out))
        same = None if first is None else bool(torch.equal(out, first))
        if first is None:
            first = out
        print(f"{name:30s} engine={st['engine']:9s} workers={st['workers']:5d} kernel_ms min/med = {min(ker):7.3f} / "
              f"{sorted(ker)[len(ker) // 2]:7.3f}   total_ms med = {sorted(tot)[len(tot) // 2]:7.3f}   population {pop}"
              f"{'' if ref_pop is None else ' (reference %d: %s)' % (ref_pop, 'EQUAL' if pop == ref_pop else 'DIFFERENT')}"
              f"   grid equal to the first variant's: {same}   wall {time.perf_counter() - t0:.1f} s", flush=True)
    grid.close()


if __name__ == "__main__":
    main()
