"""Multi-GPU ca3d: z-block slab decomposition, one process per GPU (torch.distributed for the plumbing).

The dependency structure of the reference's in-place sweep (core/ca3d.c:129-140) couples plane z of
generation g to plane z-1 of generation g and plane z+1 of generation g-1.  The volume is cut into
z-blocks dealt round-robin to the ranks; inside the fused sweep kernel the service warp of every tile that
touches a z-block edge copies the finished rows into the neighbouring GPU's ghost plane over NVLink (CUDA IPC
peer mapping) and raises its progress counters, so the halo exchange overlaps the interior updates row by row.
The only collective on the path is the final all-reduce of the per-rank populations (ca3d_run's return value).

`LocalRanks` runs the same decomposition with several slabs of ONE process on one device (each rank's launch
keeps to its share of the SMs): the sharded path is exercised -- ghost planes, peer pushes, counters and all --
on a single-GPU box.
"""
import ctypes
from ctypes import byref, c_int, c_int64, c_void_p

import numpy as np

from . import _lib
from ._lib import RunStats, check
from .rules import CellAutomaton, ca3d_rule

DEFAULT_BLOCK_PLANES = 16


TILE_PLANES = 6     # planes per tile of the kernel sharded 3-plane volumes run (18 compute warps = 6 planes x 3 generations)


def default_block_planes(d2, nranks, tile_planes=TILE_PLANES, largest=48):
    """z-block size of the scaling bench: a multiple of the tile height (a ragged tile leaves most of a CTA idle), the
    largest one up to `largest` planes that still gives every rank five blocks and keeps the most loaded rank within
    15 % of its fair share.  Measured at 2048^3 x 50 with the final kernels (profiles/r02_knobs_multi_wide_n{2,4,8}.txt):
    8 GPUs, blocks of 24 / 30 / 36 / 42 / 48 planes: sweep 18.2 / 17.8 / 17.6 / 17.2 / 16.9 ms; 4 GPUs, 36 / 42 / 48:
    30.5 / 29.6 / 28.2 ms; 2 GPUs, 42 / 60: 49.6 / 49.0 ms -- z-block edges cost more than a few percent of imbalance
    (every edge adds the NVLink hop to the wave that carries a generation up the volume), and blocks that are too large
    leave the last ranks waiting for that wave (with the 15-warp kernel of earlier in the round: 50 planes 19.9 ms
    against 18.6 ms for 40)."""
    d2, nranks = int(d2), max(1, int(nranks))
    if nranks == 1:
        return d2
    per_rank = -(-d2 // nranks)
    fair = d2 / nranks
    best = None
    b = tile_planes
    while b <= min(largest, per_rank):
        loads = [0] * nranks
        for j, z0 in enumerate(range(0, d2, b)):
            loads[j % nranks] += min(b, d2 - z0)
        if per_rank >= 5 * b and max(loads) <= 1.15 * fair:
            best = b
        b += tile_planes
    if best is None:
        # small volumes: the largest whole-tile block that still gives every rank two blocks, else one block per rank
        b = (per_rank // 2) // tile_planes * tile_planes
        best = b if b >= tile_planes else max(1, min(per_rank, d2))
    return max(1, min(best, d2))


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
                 all_gather=None, connect=True):
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
        if not connect:
            return                      # the caller wires the ranks with connect_local()
        if nranks > 1:
            if all_gather is None:
                raise ValueError("nranks > 1 needs an all_gather callable for the IPC handle exchange")
            mine = ctypes.create_string_buffer(64)
            check(self._lib, self._lib.clapca_slab_ipc_handle(self._h, mine))
            nxt, prv = exchange_handles(mine.raw, rank, nranks, all_gather)
            check(self._lib, self._lib.clapca_slab_connect(self._h, ctypes.c_char_p(nxt), ctypes.c_char_p(prv)))
        else:
            check(self._lib, self._lib.clapca_slab_connect(self._h, None, None))

    def halo_ptr(self):
        return self._lib.clapca_slab_halo_ptr(self._h)

    def connect_local(self, halo_next, halo_prev, max_ctas=0):
        """same-process neighbours: their halo pointers instead of IPC handles (clapca_slab_connect_local)"""
        check(self._lib, self._lib.clapca_slab_connect_local(self._h, c_void_p(halo_next), c_void_p(halo_prev),
                                                             int(max_ctas)))

    def plane_hashes(self):
        """64-bit fingerprints of the local planes (local order), computed on the device"""
        from .ca import hash_planes
        return hash_planes(self.device_ptr(), self.dims[0] * self.dims[1], self.n_local)

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

    def prepare_streamed(self, rule, steps):
        rule = rule if isinstance(rule, CellAutomaton) else ca3d_rule(int(rule))
        check(self._lib, self._lib.clapca_slab_prepare_streamed(self._h, rule.surv_mask, rule.born_mask, rule.nr_states,
                                                                int(steps)))

    def run_streamed(self, host_in, host_out):
        """host -> device -> host as one pipeline (clapca_slab_run_streamed): this rank's planes, local order, in
        page-locked memory (numpy arrays over pinned storage or raw addresses)"""
        pin = host_in.ctypes.data if isinstance(host_in, np.ndarray) else int(host_in)
        pout = host_out.ctypes.data if isinstance(host_out, np.ndarray) else int(host_out)
        pop = c_int64(0)
        check(self._lib, self._lib.clapca_slab_run_streamed(self._h, c_void_p(pin), c_void_p(pout), byref(pop)))
        return pop.value

    def stats(self):
        st = RunStats()
        check(self._lib, self._lib.clapca_slab_last_stats(self._h, byref(st)))
        return {"total_ms": st.total_ms, "kernel_ms": st.kernel_ms, "launches": st.launches, "planes": st.planes,
                "workers": st.workers, "streamed": bool(st.streamed)}


def bind_to_gpu_numa_node(device_index):
    """Pin the calling process to the CPUs NVML reports as closest to `device_index` (its NUMA node), so that the pinned
    host slabs a rank allocates afterwards are first-touched next to its GPU's PCIe root.  On a two-socket 8-GPU host the
    streamed host -> host pipelines of all ranks otherwise share whatever node the launcher happened to run on.
    Returns the CPU set chosen, or None when NVML / the topology is not available (nothing is changed then)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        try:
            bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:       # noqa: BLE001 -- older torch: fall back to the NVML index
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        masks = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(masks) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:           # noqa: BLE001 -- advisory: never fail a run over an affinity hint
        return None


class LocalRanks:
    """`nranks` slabs of one process on ONE device: every rank owns its z-blocks, its halo region and its stream;
    prepare runs rank by rank (it also seeds the neighbours' ghost planes), the sweeps run concurrently from one
    thread per rank, each launch capped at its share of the SMs so that all of them are co-resident."""

    def __init__(self, d0, d1, d2, nranks, max_generations, max_value, block_planes, max_ctas=None):
        lib = _lib.lib()
        self.dims = (int(d0), int(d1), int(d2))
        self.nranks = int(nranks)
        if max_ctas is None:
            max_ctas = max(1, int(lib.clapca_sm_count()) // self.nranks)
        self.max_ctas = int(max_ctas)
        self.ranks = [ShardedVolume(d0, d1, d2, r, nranks, max_generations, max_value, block_planes, connect=False)
                      for r in range(self.nranks)]
        for r, vol in enumerate(self.ranks):
            nxt, prv = neighbour_ranks(r, self.nranks)
            vol.connect_local(self.ranks[nxt].halo_ptr(), self.ranks[prv].halo_ptr(), self.max_ctas)

    def close(self):
        for vol in self.ranks:
            vol.close()

    def upload(self, volume):
        """volume: (d2, d1, d0) uint8 numpy array of the WHOLE volume"""
        for vol in self.ranks:
            if vol.n_local:
                vol.upload(np.ascontiguousarray(volume[vol.zglobal]))

    def run(self, rule, steps):
        """prepare every rank, then all sweeps at once; returns the population of the whole volume"""
        import threading
        for vol in self.ranks:
            vol.prepare(rule, steps)
        pops, errs = [0] * self.nranks, []

        def work(r):
            try:
                pops[r] = self.ranks[r].run()
            except Exception as e:      # noqa: BLE001 -- re-raised below, on the caller's thread
                errs.append(e)

        ths = [threading.Thread(target=work, args=(r,)) for r in range(self.nranks)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]
        return sum(pops)

    def download(self):
        d0, d1, d2 = self.dims
        out = np.zeros((d2, d1, d0), np.uint8)
        for vol in self.ranks:
            if vol.n_local:
                buf = np.empty((vol.n_local, d1, d0), np.uint8)
                vol.download(buf)
                out[vol.zglobal] = buf
        return out

    def run_streamed(self, rule, steps, host_in, host_out):
        """every rank's host -> device -> host pipeline at once.  host_in / host_out: per-rank lists of pinned uint8
        buffers (this rank's planes in local order); returns the population of the whole volume"""
        import threading
        for vol in self.ranks:
            vol.prepare_streamed(rule, steps)
        pops, errs = [0] * self.nranks, []

        def work(r):
            try:
                pops[r] = self.ranks[r].run_streamed(host_in[r], host_out[r])
            except Exception as e:      # noqa: BLE001 -- re-raised below, on the caller's thread
                errs.append(e)

        ths = [threading.Thread(target=work, args=(r,)) for r in range(self.nranks)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]
        return sum(pops)

    def plane_hashes(self):
        out = np.zeros(self.dims[2], np.uint64)
        for vol in self.ranks:
            if vol.n_local:
                out[vol.zglobal] = vol.plane_hashes()
        return out


def torch_all_gather_bytes(dist, device):
    """all_gather of a small byte string through torch.distributed (NCCL wants device tensors)."""
    import torch

    def gather(payload):
        t = torch.tensor(list(payload), dtype=torch.uint8, device=device)
        outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(outs, t)
        return [bytes(o.cpu().tolist()) for o in outs]

    return gather
