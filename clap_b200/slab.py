"""Multi-GPU ca3d: z-block slab decomposition, one process per GPU (torch.distributed for the plumbing).

The dependency structure of the reference's in-place sweep (core/ca3d.c:129-140) couples plane z of
generation g to plane z-1 of generation g and plane z+1 of generation g-1.  The volume is cut into
z-blocks dealt round-robin to the ranks; inside the fused sweep kernel the edge planes of a block store
their rows directly into the neighbouring GPU's ghost planes over NVLink (CUDA IPC peer mapping) and
raise its progress counters, so the halo exchange overlaps the interior updates row by row.  The only
collective on the path is the final all-reduce of the per-rank populations (ca3d_run's return value).
"""
import ctypes
from ctypes import byref, c_int, c_int64, c_void_p

import numpy as np

from . import _lib
from ._lib import RunStats, check
from .rules import CellAutomaton, ca3d_rule

DEFAULT_BLOCK_PLANES = 16


def default_block_planes(d2, nranks):
    """z-block size of the scaling bench.  Every z-block edge costs the team that owns it (its edge plane polls ghost
    tags and pushes rows over NVLink while 15 team-mates follow it row by row), every block adds pipeline depth:
    measured on 2 x B200 at 2048^3 x 50 (profiles/r01_knobs_multi_team_blocks_n2.txt) the sweep takes 85 / 85 / 80 /
    74 ms with blocks of 16 / 32 / 64 / 128 planes.  Hence: blocks of 128 planes, but at least two blocks per rank
    so that the fill of the rank pipeline, (nranks - 1) / (generations * blocks per rank), stays small."""
    per_rank = max(1, -(-int(d2) // max(1, int(nranks))))
    return max(1, min(128, max(16, per_rank // 2), per_rank))


# ---- pure host-side planning (mirrors SlabGeom in csrc/bp_plan.h; unit-tested on CPU) -----------------

def plan_blocks(d2, nranks, block_planes):
    """[(rank, z0, z1)] for every z-block, in global order: block j -> rank j % nranks."""
    if nranks == 1:
        return [(0, 0, d2)]
    out = []
    j = 0
    for z0 in range(0, d2, block_planes):
        out.append((j % nranks, z0, min(d2, z0 + block_planes)))
        j += 1
    return out


def local_planes(d2, nranks, block_planes, rank):
    """Global z of each plane rank `rank` owns, in its local storage order."""
    zs = []
    for r, z0, z1 in plan_blocks(d2, nranks, block_planes):
        if r == rank:
            zs.extend(range(z0, z1))
    return zs


def neighbour_ranks(rank, nranks):
    """(next, prev): owners of the block after / before any block of `rank`."""
    return (rank + 1) % nranks, (rank - 1 + nranks) % nranks


def exchange_handles(handle, rank, nranks, all_gather):
    """All-gather the 64-byte halo handles; returns (handle of next rank, handle of prev rank).
    `all_gather(bytes) -> list[bytes]` abstracts the collective (NCCL on the box, gloo in the CPU tests)."""
    handles = all_gather(bytes(handle))
    if len(handles) != nranks or any(len(h) != 64 for h in handles):
        raise ValueError("handle exchange returned malformed data")
    nxt, prv = neighbour_ranks(rank, nranks)
    return handles[nxt], handles[prv]


# ---- the device object ----------------------------------------------------------------------------------

class ShardedVolume:
    """This rank's share of a d0 x d1 x d2 ca3d volume (clapca_slab_*)."""

    def __init__(self, d0, d1, d2, rank, nranks, max_generations, max_value, block_planes=DEFAULT_BLOCK_PLANES,
                 all_gather=None):
        self._lib = _lib.lib()
        self.dims = (int(d0), int(d1), int(d2))
        self.rank, self.nranks = int(rank), int(nranks)
        self.block_planes = int(block_planes)
        h = c_void_p()
        check(self._lib, self._lib.clapca_slab_create(byref(h), d0, d1, d2, rank, nranks, block_planes,
                                                      max_generations, max_value))
        self._h = h
        n = c_int()
        check(self._lib, self._lib.clapca_slab_local_planes(self._h, byref(n)))
        self.n_local = n.value
        zmap = (c_int64 * max(1, self.n_local))()
        check(self._lib, self._lib.clapca_slab_plane_map(self._h, zmap))
        self.zglobal = [int(zmap[i]) for i in range(self.n_local)]
        assert self.zglobal == local_planes(d2, nranks, block_planes, rank)
        if nranks > 1:
            if all_gather is None:
                raise ValueError("nranks > 1 needs an all_gather callable for the IPC handle exchange")
            mine = ctypes.create_string_buffer(64)
            check(self._lib, self._lib.clapca_slab_ipc_handle(self._h, mine))
            nxt, prv = exchange_handles(mine.raw, rank, nranks, all_gather)
            check(self._lib, self._lib.clapca_slab_connect(self._h, ctypes.c_char_p(nxt), ctypes.c_char_p(prv)))
        else:
            check(self._lib, self._lib.clapca_slab_connect(self._h, None, None))

    def close(self):
        if self._h:
            self._lib.clapca_slab_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_ptr(self):
        return self._lib.clapca_slab_device_ptr(self._h)

    def upload(self, src):
        ptr = src.ctypes.data if isinstance(src, np.ndarray) else int(src)
        check(self._lib, self._lib.clapca_slab_upload(self._h, c_void_p(ptr)))

    def download(self, dst):
        ptr = dst.ctypes.data if isinstance(dst, np.ndarray) else int(dst)
        check(self._lib, self._lib.clapca_slab_download(self._h, c_void_p(ptr)))

    def prepare(self, rule, steps):
        rule = rule if isinstance(rule, CellAutomaton) else ca3d_rule(int(rule))
        check(self._lib, self._lib.clapca_slab_prepare(self._h, rule.surv_mask, rule.born_mask, rule.nr_states,
                                                       int(steps)))

    def run(self):
        pop = c_int64(0)
        check(self._lib, self._lib.clapca_slab_run(self._h, byref(pop)))
        return pop.value

    def stats(self):
        st = RunStats()
        check(self._lib, self._lib.clapca_slab_last_stats(self._h, byref(st)))
        return {"total_ms": st.total_ms, "kernel_ms": st.kernel_ms, "launches": st.launches, "planes": st.planes,
                "workers": st.workers}


def torch_all_gather_bytes(dist, device):
    """all_gather of a small byte string through torch.distributed (NCCL wants device tensors)."""
    import torch

    def gather(payload):
        t = torch.tensor(list(payload), dtype=torch.uint8, device=device)
        outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(outs, t)
        return [bytes(o.cpu().tolist()) for o in outs]

    return gather


def run_sharded_bench(args, workload, synth_planes):
    """bench.py body for WORLD_SIZE > 1: every rank owns its z-blocks of the volume; strong scaling."""
    import json
    import os
    import time

    import torch
    import torch.distributed as dist
    from bench import ClockSampler, measured_peak

    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    d0, d1, d2, gens, rule_index = workload
    rule = ca3d_rule(rule_index)
    block = int(os.environ.get("CLAPCA_BLOCK_PLANES", default_block_planes(d2, world)))

    vol = ShardedVolume(d0, d1, d2, rank, world, gens, 5, block, torch_all_gather_bytes(dist, dev))
    # synthetic seed, generated block by block so every rank produces exactly the cells a single GPU would
    seed_dev = torch.empty((max(1, vol.n_local), d1, d0), dtype=torch.uint8, device=dev)
    l = 0
    for r, z0, z1 in plan_blocks(d2, world, block):
        if r == rank:
            seed_dev[l:l + z1 - z0] = synth_planes(torch, d0, d1, z0, z1, dev)
            l += z1 - z0
    torch.cuda.synchronize()

    def step():
        vol.upload(seed_dev.data_ptr())
        dist.barrier()
        vol.prepare(rule, gens)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        pop = vol.run()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        return pop, vol.stats(), wall

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    tot, ker, wall_sum, launches, pop = 0.0, 0.0, 0.0, 0, 0
    for _ in range(args.steps):
        pop, st, wall = step()
        # device time of the step on this rank -> max over ranks (the sweep kernels overlap in time)
        t = torch.tensor([st["total_ms"], st["kernel_ms"], wall * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot += float(t[0]); ker += float(t[1]); wall_sum += float(t[2])
        launches += st["launches"]
    clocks = sampler.stop() if rank == 0 else None
    tp = torch.tensor([pop], dtype=torch.int64, device=dev)
    dist.all_reduce(tp)                                   # ca3d_run's return value: population of the whole volume
    tl = torch.tensor([launches], dtype=torch.int64, device=dev)
    dist.all_reduce(tl)

    # end to end: pinned host slabs in, pinned host slabs out
    e2e = None
    if not args.no_e2e:
        nbytes = vol.n_local * d0 * d1
        host_in = torch.empty(max(1, nbytes), dtype=torch.uint8, pin_memory=True)
        host_out = torch.empty(max(1, nbytes), dtype=torch.uint8, pin_memory=True)
        host_in[:nbytes].copy_(seed_dev.reshape(-1)[:nbytes])
        torch.cuda.synchronize()
        n_e2e = max(1, min(args.steps, 3))
        dts = []
        for i in range(1 + n_e2e):
            dist.barrier()
            t0 = time.perf_counter()
            vol.upload(host_in.data_ptr())
            vol.prepare(rule, gens)
            dist.barrier()
            vol.run()
            vol.download(host_out.data_ptr())
            dist.barrier()
            if i:
                dts.append(time.perf_counter() - t0)
        t = torch.tensor([sum(dts) / len(dts)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": d0 * d1 * d2 * gens / float(t[0]) / 1e9, "unit": "GCUPS",
               "h2d_bytes_per_step": d0 * d1 * d2, "d2h_bytes_per_step": d0 * d1 * d2 + 8 * world,
               "ms_per_step": float(t[0]) * 1e3, "steps": n_e2e}

    if rank == 0:
        updates = d0 * d1 * d2 * gens
        ms = tot / args.steps
        kms = ker / args.steps
        peak, peak_src = measured_peak()
        achieved = updates * 2.0 / world / (kms * 1e-3) / 1e9
        line = {
            "metric": "ca3d cell-updates/s", "value": updates / (ms * 1e-3) / 1e9, "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.workload}: ca3d_run {d0}x{d1}x{d2} uint8, {gens} generations, rule "
                                   f"{rule.name}, seed P(alive)=1/4 values 1..5",
                       "parallelism": f"z-blocks of {block} planes dealt round-robin to {world} GPUs, halo rows as "
                                      f"peer stores inside the sweep kernel (NVLink), population all-reduce (NCCL)",
                       "engine": "bitplane", "planes": st["planes"], "workers_per_gpu": st["workers"],
                       "l2": "per-GPU slab (%.1f GiB) is larger than L2" % (d0 * d1 * d2 / world / 2 ** 30),
                       "population": int(tp[0]), "wall_ms_per_step_sweep_max_rank": wall_sum / args.steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "ca3d_sweep_kernel (per GPU, all generations fused)",
                         "kernel_ms": kms, "algorithmic_bytes_per_update": 2.0, "peak_source": peak_src},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(tl[0]), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    vol.close()
    dist.destroy_process_group()
