#!/usr/bin/env python
"""Sharded ca3d sweep (one process per GPU, launched with torchrun): kernel time over z-block size and sweep knobs,
all variants inside ONE set of processes (start-up and NCCL init are paid once).

usage: torchrun ... tools/multi_knobs.py [side] [gens] "BLOCK[,ENV=v,...];BLOCK[,ENV=v,...];..."
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

import bench
import clap_b200
from clap_b200.rules import ca3d_rule
from clap_b200.slab import ShardedVolume, plan_blocks, torch_all_gather_bytes


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    gens = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    variants = (sys.argv[3] if len(sys.argv) > 3 else "16;32;64").split(";")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    clap_b200.init(local)
    rule = ca3d_rule(7)
    touched = set()
    for var in variants:
        parts = var.split(",")
        block = int(parts[0])
        for k in touched:
            os.environ.pop(k, None)
        for kv in parts[1:]:
            k, v = kv.split("=")
            os.environ[k] = v
            touched.add(k)
        vol = ShardedVolume(side, side, side, rank, world, gens, 5, block, torch_all_gather_bytes(dist, dev))
        seed = torch.empty((max(1, vol.n_local), side, side), dtype=torch.uint8, device=dev)
        l = 0
        for r, z0, z1 in plan_blocks(side, world, block):
            if r == rank:
                seed[l:l + z1 - z0] = bench.synth_planes(torch, side, side, z0, z1, dev)
                l += z1 - z0
        torch.cuda.synchronize()
        best_k, best_t, pop = 1e30, 1e30, 0
        for _ in range(4):
            vol.upload(seed.data_ptr())
            dist.barrier()
            vol.prepare(rule, gens)
            torch.cuda.synchronize()
            dist.barrier()
            pop = vol.run()
            st = vol.stats()
            t = torch.tensor([st["kernel_ms"], st["total_ms"]], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best_k, best_t = min(best_k, float(t[0])), min(best_t, float(t[1]))
        tp = torch.tensor([pop], dtype=torch.int64, device=dev)
        dist.all_reduce(tp)
        # the cells, not just their count: per-plane fingerprints against the committed single-GPU run's
        import numpy as np
        hz = torch.zeros(side, dtype=torch.int64, device=dev)
        if vol.n_local:
            hz[torch.tensor(vol.zglobal, device=dev)] = torch.from_numpy(vol.plane_hashes().view(np.int64)).to(dev)
        dist.all_reduce(hz)
        if rank == 0:
            ev = bench.parity_evidence("ca3d_%d" % side, hz.cpu().numpy().view(np.uint64)) if gens == 50 and "ca3d_%d" % side in bench.WORKLOADS else {}
            print(f"N={world} side {side} gens {gens} block {block} {' '.join(parts[1:])}: sweep {best_k:8.3f} ms  step {best_t:8.3f} ms  "
                  f"{side ** 3 * gens / (best_t * 1e-3) / 1e9:7.1f} GCUPS  pop {int(tp[0])}  bit_equal_to_n1 {ev.get('bit_equal_to_n1')} "
                  f"reference_planes {ev.get('equal_to_unmodified_reference')}", flush=True)
        vol.close()
        del seed
        torch.cuda.empty_cache()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
