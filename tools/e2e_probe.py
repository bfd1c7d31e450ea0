#!/usr/bin/env python
"""End-to-end (pinned host -> device -> pinned host) timing of ca3d_run on one GPU: the streamed pipeline of
clapca_grid_run3d_streamed against upload / run / download one after the other, over chunk sizes; and the
device-resident run with separate layout kernels against the layout items of the sweep launch.

usage: e2e_probe.py [side] [gens]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import bench
import clap_b200
from clap_b200.rules import ca3d_rule


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    gens = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    clap_b200.init(0)
    rule = ca3d_rule(7)
    cells = side ** 3
    seed_dev = bench.synth_planes(torch, side, side, 0, side, dev)
    host_in = torch.empty(cells, dtype=torch.uint8, pin_memory=True)
    host_out = torch.empty(cells, dtype=torch.uint8, pin_memory=True)
    host_ref = torch.empty(cells, dtype=torch.uint8, pin_memory=True)
    host_in.copy_(seed_dev.reshape(-1))
    torch.cuda.synchronize()
    grid = clap_b200.Grid(side, side, side)
    upd = cells * gens / 1e9

    print("== device-resident run: separate layout kernels vs layout items", flush=True)
    for fused in ("0", "1"):
        os.environ["CLAPCA_FUSED_LAYOUT"] = fused
        best_t, best_k, pop = 1e30, 1e30, 0
        for _ in range(4):
            grid.upload(seed_dev.data_ptr())
            pop = grid.run3d(rule, gens)
            st = grid.stats()
            best_t, best_k = min(best_t, st["total_ms"]), min(best_k, st["kernel_ms"])
        print(f"CLAPCA_FUSED_LAYOUT={fused}: total {best_t:8.3f} ms  sweep launch {best_k:8.3f} ms  "
              f"{upd / (best_t * 1e-3):7.1f} GCUPS  launches {st['launches']}  pop {pop}", flush=True)
    os.environ["CLAPCA_FUSED_LAYOUT"] = "0"
    grid.upload(seed_dev.data_ptr())
    rpop = grid.run3d(rule, gens)
    grid.download(host_ref.data_ptr())

    print("== host -> device -> host", flush=True)
    variants = [("CLAPCA_STREAMED=0", {"CLAPCA_STREAMED": "0"})]
    for mb in (8, 32, 128):
        variants.append((f"CLAPCA_STREAMED=1 CLAPCA_IO_CHUNK_MB={mb}", {"CLAPCA_STREAMED": "1", "CLAPCA_IO_CHUNK_MB": str(mb)}))
    for name, env in variants:
        os.environ.update(env)
        best = 1e30
        for i in range(4):
            host_out.zero_()
            t0 = time.perf_counter()
            pop = grid.run3d_streamed(rule, gens, host_in.data_ptr(), host_out.data_ptr(), max_value=5)
            dt = time.perf_counter() - t0
            if i:
                best = min(best, dt)
        st = grid.stats()
        ok = pop == rpop and torch.equal(host_out, host_ref)
        print(f"{name}: {best * 1e3:8.2f} ms  {upd / best:7.1f} GCUPS  streamed {st['streamed']}  "
              f"launch {st['kernel_ms']:8.2f} ms  {'bit-exact' if ok else 'MISMATCH'}  pop {pop}", flush=True)


if __name__ == "__main__":
    main()
