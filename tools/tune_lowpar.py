#!/usr/bin/env python
"""Low-parallelism regime of the ca3d sweep kernel on ONE GPU.

At N GPUs a rank owns ~G/N sweeps per dependency level (key 2z+4g).  A single GPU running G/N generations
sees the same per-level parallelism (minus the NVLink hop), so the knobs that matter for 8-GPU scaling --
resident warps per SM, counter granularity -- can be tuned without an 8-GPU box:  `gens` = 50, 25, 12, 6
stands for N = 1, 2, 4, 8 at the headline 50 generations.

usage: tune_lowpar.py [side] ["gens list"] ["ctas list"] ["flag_rows list"] ["ENV=v ENV=v;ENV=v ..."]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import bench
import clap_b200
from clap_b200.rules import ca3d_rule


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    gens_list = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "50 25 12 6").split()]
    ctas_list = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "0 3 2").split()]
    flag_list = [int(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "8 4 2").split()]
    extras = (sys.argv[5] if len(sys.argv) > 5 else "").split(";")      # env settings, ';' separates variants
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    clap_b200.init(0)
    rule = ca3d_rule(7)
    seed_dev = bench.synth_planes(torch, side, side, 0, side, dev)
    torch.cuda.synchronize()
    grid = clap_b200.Grid(side, side, side)
    cells = side ** 3
    base = {}
    for extra in extras:
        for kv in extra.split():
            k, v = kv.split("=")
            os.environ[k] = v
        for gens in gens_list:
            for ctas in ctas_list:
                for fr in flag_list:
                    os.environ["CLAPCA_CTAS_PER_SM"] = str(ctas)
                    os.environ["CLAPCA_FLAG_ROWS"] = str(fr)
                    best, pop, st = 1e30, 0, None
                    for _ in range(3):
                        grid.upload(seed_dev.data_ptr())
                        pop = grid.run3d(rule, gens)
                        st = grid.stats()
                        best = min(best, st["kernel_ms"])
                    rate = cells * gens / (best * 1e-3) / 1e9
                    base.setdefault("rate", rate)
                    print(f"side {side} gens {gens:3d} ctas/sm {ctas} flag_rows {fr} {extra}: {best:8.3f} ms  "
                          f"{rate:7.1f} GCUPS ({rate / base['rate'] * 100:5.1f} % of first)  workers {st['workers']} "
                          f"pop {pop}", flush=True)


if __name__ == "__main__":
    main()
