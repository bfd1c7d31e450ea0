#!/usr/bin/env python
"""bench.py -- headline benchmark of the clap procedural-generation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Metric (BASELINE.json): CA cell-updates per second (GCUPS) for ca3d_run() -- core/ca3d.c:124-142 -- on a
synthetic 2048^3 uint8 volume, 50 generations of ca_coral (BASELINE config 4; fits one GPU, so it is also the
N = 1 workload).  One "step" = one full run of all generations over the volume.

  value      cells * generations / device time of the step (CUDA events on the library's stream, max over
             ranks), input already resident in HBM in the reference's uint8 layout; the layout pack/unpack
             kernels and the population count are INSIDE the timed region.
  e2e        same metric through the C-ABI grid calls with pinned HOST buffers: H2D of the seed volume and
             D2H of the result (+ population) inside the timed region, wall clock.
  roofline   the dominant kernel (the fused sweep kernel): algorithmic bytes = 2 B per cell update
              over its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference: the reference's own C (oracle/_ref/libclapref.so, built from the unmodified
             sources; oracle/port when that is absent) on one host core -- the path is single-threaded and
             sequentially dependent -- on a bounded sample of the same workload.

Rank 0 prints exactly one JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (d0, d1, d2, generations, rule index)
    "ca3d_2048": (2048, 2048, 2048, 50, 7),
    "ca3d_1024": (1024, 1024, 1024, 50, 7),
    "ca3d_512": (512, 512, 512, 50, 7),
    "ca3d_128": (128, 128, 128, 10, 7),
    # per-GPU shares of the 2048^3 volume as stand-alone volumes (scaling diagnostics)
    "ca3d_2048_z1024": (2048, 2048, 1024, 50, 7),
    "ca3d_2048_z256": (2048, 2048, 256, 50, 7),
}
# BASELINE config 3: binary cave-smoothing rule, 1 bit per cell in the 2D bit-plane engine
CA2D_WORKLOADS = {
    # name: (side, generations, born, surv, nr_states, decay)
    "ca2d_16384": (16384, 100, 0x1E0, 0x1F0, 1, True),
    "ca2d_4096": (4096, 100, 0x1E0, 0x1F0, 1, True),
    "ca2d_16384_cavetest": (16384, 100, 3 << 2, 3 << 7, 4, True),      # multi-state ca_test rule (terrain.c:391-398)
}
# BASELINE config 5 (+ the noise.c bake): per-lattice-point field evaluation
FIELD_WORKLOADS = {
    # name: (kind, size)
    "terrain_8192": ("terrain", 8192),      # terrain.c:447-467 heightmap, maze = ca2d_generate(&ca_test, 1024, 4)
    "terrain_1024": ("terrain", 1024),
    "terrain_mesh_8192": ("mesh", 8192),    # terrain.c:479-516 vertex / normal / uv / index buffers from the heightmap
    "terrain_mesh_1024": ("mesh", 1024),
    "noise_256": ("noise", 256),            # noise_grad3d_bake_rgba8(256, 4, 2.0, 0.5, 37.0, 0xc14d)
    "noise_64": ("noise", 64),              # the engine's default bake (noise.c:309-317)
}
SEED = 0xC1A9
CHUNK_PLANES = 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ca3d_2048", choices=sorted(WORKLOADS) + sorted(CA2D_WORKLOADS) + sorted(FIELD_WORKLOADS))
    ap.add_argument("--engine", default="auto", choices=["auto", "wavefront", "bitplane"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic seed volume: P(alive) = 1/4, values uniform 1..5, zero boundary outside (SURVEY.md 8d cfg 4).
# Generated per 64-plane chunk from a (seed, chunk) keyed generator so any slab owner produces the same
# cells as a single GPU would.
# ------------------------------------------------------------------------------------------------
def synth_planes(torch, d0, d1, z0, z1, device):
    out = torch.empty((z1 - z0, d1, d0), dtype=torch.uint8, device=device)
    z = z0
    while z < z1:
        c = z // CHUNK_PLANES
        cz0, cz1 = c * CHUNK_PLANES, (c + 1) * CHUNK_PLANES
        gen = torch.Generator(device=device)
        gen.manual_seed(SEED * 1000003 + c)
        r = torch.randint(0, 20, (CHUNK_PLANES, d1, d0), dtype=torch.uint8, device=device, generator=gen)
        r = torch.where(r < 5, r + 1, torch.zeros_like(r))
        lo, hi = max(z, cz0), min(z1, cz1)
        out[lo - z0:hi - z0] = r[lo - cz0:hi - cz0]
        z = hi
    return out


def synth_numpy(np, shape, seed):
    rng = np.random.default_rng(seed)
    r = rng.integers(0, 20, shape, dtype=np.uint8)
    return np.where(r < 5, r + 1, 0).astype(np.uint8)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic(workload):
    """DRAM bytes per launch of the dominant kernel, from the committed ncu capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return int(json.load(f)[workload]["bytes"])
    except (OSError, KeyError, ValueError):
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference C on one host core, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(np, rule_index, gens_full, budget_s):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    ref = oracle_lib.ref()
    side = 256
    vol = synth_numpy(np, (side, side, side), SEED)
    # ~50 MCUPS on one core -> one generation of 256^3 is ~0.35 s; size the sample to the budget
    gens = int(max(1, min(gens_full, budget_s / 0.40)))
    if ref is not None:
        kind = "reference"
        t0 = time.perf_counter()
        ref.ca3d_run(vol, rule_index, gens)
        dt = time.perf_counter() - t0
    else:
        kind = "port"
        port = oracle_lib.port()
        s, b, n = port.ca3d_rule(rule_index)
        t0 = time.perf_counter()
        port.ca3d_run(vol, s, b, n, gens)
        dt = time.perf_counter() - t0
    gcups = side ** 3 * gens / dt / 1e9
    return {"value": gcups, "unit": "GCUPS", "cores": 1, "kind": kind, "seconds": dt,
            "sample": f"ca3d_run ca_coral on a {side}^3 corner-sized synthetic volume (same P(alive)=1/4, values "
                      f"1..5), {gens} generations, {dt:.1f} s on one core of {os.cpu_count()} (reference path is "
                      f"single-threaded and sequentially dependent)"}


def run_reference_arm(args):
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload not in WORKLOADS:
        raise SystemExit("--impl reference times the headline ca3d workloads")
    d0, d1, d2, gens, rule = WORKLOADS[args.workload]
    total = args.steps + args.warmup
    budget = max(2.0, min(20.0, 150.0 / max(1, total)))
    vals, secs = [], []
    base = None
    for i in range(total):
        base = cpu_reference_sample(np, rule, gens, budget)
        if i >= args.warmup:
            vals.append(base["value"])
            secs.append(base["seconds"])
    v = sum(vals) / len(vals)
    base["value"] = v
    line = {
        "impl": "reference", "metric": "ca3d cell-updates/s", "value": v, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(secs) / len(secs) * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: ca3d_run {d0}x{d1}x{d2}, {gens} generations, rule ca_coral",
                   "note": "one step = the bounded sample described in cpu_baseline.sample, not the full volume"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# secondary workload: BASELINE config 3 (ca2d, bit-packed).  Same JSON keys; not the headline line.
# ------------------------------------------------------------------------------------------------
def run_ca2d(args, torch, clap_b200, dev, local):
    import numpy as np
    from clap_b200 import _lib as _lib_mod
    from clap_b200.rules import CellAutomaton
    from clap_b200._lib import NEIGH_M1
    side, gens, born, surv, nr, decay = CA2D_WORKLOADS[args.workload]
    ca = CellAutomaton(args.workload, born, surv, nr, decay, NEIGH_M1)
    cells = side * side
    updates = cells * gens
    # SURVEY 8(d) cfg 3 seed: srand48(1) + the reference's own fill loop (ca2d.c:86-90), drawn on the device
    # (clapca_grid_seed2d); a pristine copy is kept in a torch tensor for the state reset between steps
    from clap_b200.ca import Rand48
    grid = clap_b200.Grid(side, side, 1)
    grid.seed2d(ca, Rand48(1))
    stage = torch.empty(cells, dtype=torch.uint8, pin_memory=True)
    grid.download(stage.data_ptr())
    seed_dev = stage.to(dev).reshape(side, side)
    torch.cuda.synchronize()

    def step():
        grid.upload(seed_dev.data_ptr())
        grid.run2d(ca, gens)
        return grid.stats()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    tot_ms = ker_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        st = step()
        tot_ms += st["total_ms"]; ker_ms += st["kernel_ms"]; launches += st["launches"]
    torch.cuda.synchronize()
    clocks = sampler.stop()
    pop = grid.count()
    # BASELINE config 3 at full size was run ONCE through the unmodified reference (tests/golden/make_golden_cfg3.py):
    # the population of the final grid must be the reference's
    ref_pop = None
    if args.workload == "ca2d_16384":
        try:
            with open(os.path.join(ROOT, "tests", "golden", "cfg3_16384.json")) as f:
                ref_pop = json.load(f)["cave_bin_16384_x100_seed1"]["final_grid"]["population"]
        except (OSError, KeyError, ValueError):
            ref_pop = None
        if ref_pop is not None:
            assert pop == ref_pop, f"population {pop} differs from the reference's {ref_pop}"
    ms_per_step = tot_ms / args.steps
    kernel_ms = ker_ms / args.steps
    bytes_per_update = 0.25 if st["planes"] == 1 else 2.0
    peak, peak_src = measured_peak()
    achieved = updates * bytes_per_update / (kernel_ms * 1e-3) / 1e9
    e2e = None
    if not args.no_e2e:
        # the reference-facing call: ca2d_generate(ca, side, gens) after srand48(1) -- seeding and every generation
        # on the device, the finished grid copied back to a pinned host buffer (no H2D at all)
        out = stage
        torch.cuda.synchronize()
        n = max(1, min(args.steps, 3))
        from ctypes import byref, c_uint64, c_void_p
        lib = _lib_mod.lib()
        for i in range(1 + n):
            if i == 1:
                t0 = time.perf_counter()
            after = c_uint64(0)
            _lib_mod.check(lib, lib.clapca_ca2d_generate(c_void_p(out.data_ptr()), side, born, surv, nr, int(decay), NEIGH_M1,
                                                         gens, 0, Rand48(1).x, byref(after)))
        dt = (time.perf_counter() - t0) / n
        assert int(torch.count_nonzero(out)) == pop, "ca2d_generate result differs from the device-resident run"
        e2e = {"value": updates / dt / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": cells,
               "ms_per_step": dt * 1e3, "steps": n, "call": "clapca_ca2d_generate (device-side seeding + generations + D2H)"}
    cpu = None
    if not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        ref = oracle_lib.ref()
        cs = 2048
        arr = (np.random.default_rng(SEED).integers(0, 8, (cs, cs)) <= nr).astype(np.uint8) * (nr & 0xFF)
        k = int(max(1, min(gens, args.cpu_seconds / 0.12)))
        t0 = time.perf_counter()
        if ref is not None:
            ref.ca2d_step(arr, born, surv, nr, decay, NEIGH_M1, steps=k)
            kind = "reference"
        else:
            oracle_lib.port().ca2d_run(arr, born, surv, nr, decay, NEIGH_M1, k)
            kind = "port"
        dt = time.perf_counter() - t0
        cpu = {"value": cs * cs * k / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": kind,
               "sample": f"ca2d_step x{k} on a {cs}^2 grid of the same rule and density, {dt:.1f} s on one core of "
                         f"{os.cpu_count()} (single-threaded, sequentially dependent reference path)"}
    line = {
        "metric": "ca2d cell-updates/s", "value": updates / (ms_per_step * 1e-3) / 1e9, "unit": "GCUPS", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: ca2d_step x{gens} on {side}x{side} uint8, born 0x{born:x} surv 0x{surv:x} "
                               f"nr_states {nr} decay {int(decay)} m1, seed = srand48(1) + the reference fill loop (ca2d.c:86-90)",
                   "engine": st["engine"], "planes": st["planes"], "workers": st["workers"],
                   "l2": "the bit-packed grid is L2-resident by design; state is reset from a pristine uint8 device copy "
                         "(%.0f MiB, larger than L2) before every step" % (cells / 2 ** 20),
                   "population": pop, "population_of_the_unmodified_reference": ref_pop},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(args.workload), "kernel": "ca2d_sweep_kernel (all generations fused)", "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_update": bytes_per_update, "peak_source": peak_src,
                     "note": "dependency-latency bound: the in-place sweep order leaves a chain of side + 2*generations "
                             "row steps"},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# secondary workloads: BASELINE config 5 (terrain.c heightmap) and the noise.c gradient bake
# ------------------------------------------------------------------------------------------------
def run_fields(args, torch, clap_b200, dev, local):
    import ctypes
    import numpy as np
    from ctypes import byref, c_float, c_void_p
    from clap_b200 import _lib
    from clap_b200.ca import Rand48
    kind, size = FIELD_WORKLOADS[args.workload]
    lib = _lib.lib()
    peak, peak_src = measured_peak()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)       # > L2 (126 MB): written between steps
    TSEED, NSEED = 12345, 0xC14D

    if kind == "terrain":
        nr_v, mside = size, size // 8
        units = nr_v * nr_v
        maze = clap_b200.ca2d_generate(clap_b200.CA_TEST, mside, 4, Rand48(7))      # terrain.c:434
        d_maze = torch.from_numpy(maze).to(dev)
        d_map0 = torch.empty(units, dtype=torch.float32, device=dev)
        d_map = torch.empty(units, dtype=torch.float32, device=dev)
        bytes_per_unit, unit, metric = 8.0, "Mvertex/s", "terrain heightmap vertices/s"
        kernel = "terrain_heightmap_kernel (+ terrain_map0_kernel for the lattice)"
        desc = (f"{args.workload}: terrain.c:447-467 map fill, nr_v {nr_v}, seed {TSEED}, maze = ca2d_generate(&ca_test, "
                f"{mside}, 4) after srand48(7), 4 octaves")

        def step():
            a, b = c_float(), c_float()
            _lib.check(lib, lib.clapca_terrain_heightmap_device(c_void_p(d_map.data_ptr()), c_void_p(d_map0.data_ptr()),
                                                                TSEED, nr_v, 0.0, c_void_p(d_maze.data_ptr()), mside,
                                                                1.0, 4, byref(a), byref(b)))
            return a.value + b.value, b.value, 2 if os.environ.get("CLAPCA_TERRAIN_DIRECT") else 3

        host_out = torch.empty(units, dtype=torch.float32, pin_memory=True)

        def e2e_step():
            _lib.check(lib, lib.clapca_terrain_heightmap(c_void_p(host_out.data_ptr()), TSEED, nr_v, 0.0,
                                                         maze.ctypes.data_as(c_void_p), mside, 1.0, 4))
        h2d, d2h = maze.nbytes, units * 4
    elif kind == "mesh":
        nr_v, mside = size, size // 8
        units = nr_v * nr_v
        quads = (nr_v - 1) * (nr_v - 1)
        maze = clap_b200.ca2d_generate(clap_b200.CA_TEST, mside, 4, Rand48(7))
        d_maze = torch.from_numpy(maze).to(dev)
        d_map0 = torch.empty(units, dtype=torch.float32, device=dev)
        d_map = torch.empty(units, dtype=torch.float32, device=dev)
        _lib.check(lib, lib.clapca_terrain_heightmap_device(c_void_p(d_map.data_ptr()), c_void_p(d_map0.data_ptr()), TSEED,
                                                            nr_v, 0.0, c_void_p(d_maze.data_ptr()), mside, 1.0, 4, None, None))
        del d_map0
        d_vx = torch.empty(units * 3, dtype=torch.float32, device=dev)
        d_norm = torch.empty(units * 3, dtype=torch.float32, device=dev)
        d_tx = torch.empty(units * 2, dtype=torch.float32, device=dev)
        d_idx = torch.empty(quads * 6, dtype=torch.int16, device=dev)
        # per vertex: 4 B of map read, 12 + 12 + 8 B of vertex / normal / uv written, 12 B of indices per quad
        bytes_per_unit, unit, metric = 36.0 + 12.0 * quads / units, "Mvertex/s", "terrain mesh vertices/s"
        kernel = "terrain_mesh_vertex_kernel + terrain_mesh_index_kernel"
        desc = (f"{args.workload}: terrain.c:479-516 mesh buffers (vx, norm, tx, idx) of the {nr_v}^2 heightmap of "
                f"terrain_{nr_v}, origin (0,0,0), side {nr_v // 4}")
        side = float(nr_v // 4)

        def step():
            a = c_float()
            _lib.check(lib, lib.clapca_terrain_mesh_device(c_void_p(d_map.data_ptr()), nr_v, 0.0, 0.0, 0.0, side,
                                                           c_void_p(d_vx.data_ptr()), c_void_p(d_norm.data_ptr()),
                                                           c_void_p(d_tx.data_ptr()), c_void_p(d_idx.data_ptr()), byref(a)))
            return a.value, a.value, 2

        host_map = torch.empty(units, dtype=torch.float32, pin_memory=True)
        host_map.copy_(d_map)
        h_vx = torch.empty(units * 3, dtype=torch.float32, pin_memory=True)
        h_norm = torch.empty(units * 3, dtype=torch.float32, pin_memory=True)
        h_tx = torch.empty(units * 2, dtype=torch.float32, pin_memory=True)
        h_idx = torch.empty(quads * 6, dtype=torch.int16, pin_memory=True)

        def e2e_step():
            _lib.check(lib, lib.clapca_terrain_mesh(c_void_p(host_map.data_ptr()), nr_v, 0.0, 0.0, 0.0, side,
                                                    c_void_p(h_vx.data_ptr()), c_void_p(h_norm.data_ptr()),
                                                    c_void_p(h_tx.data_ptr()), c_void_p(h_idx.data_ptr())))
        h2d, d2h = units * 4, units * 32 + quads * 12
    else:
        units = size ** 3
        period = 37.0 if size == 256 else 64.0
        d_out = torch.empty(units, dtype=torch.int32, device=dev)
        bytes_per_unit, unit, metric = 4.0, "Mvoxel/s", "noise gradient bake voxels/s"
        kernel = "noise_bake_kernel"
        desc = f"{args.workload}: noise_grad3d_bake_rgba8({size}, 4, 2.0, 0.5, {period}, 0xc14d) (noise.c:222-270)"

        def step():
            a = c_float()
            _lib.check(lib, lib.clapca_noise_bake_device(c_void_p(d_out.data_ptr()), size, 4, 2.0, 0.5, period, NSEED,
                                                         byref(a)))
            return a.value, a.value, 1

        host_out = torch.empty(units, dtype=torch.int32, pin_memory=True)

        def e2e_step():
            _lib.check(lib, lib.clapca_noise_grad3d_bake_rgba8(c_void_p(host_out.data_ptr()), size, 4, 2.0, 0.5, period,
                                                               NSEED))
        h2d, d2h = 0, units * 4

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    tot_ms = ker_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t, k, n = step()
        tot_ms += t; ker_ms += k; launches += n
    clocks = sampler.stop()
    ms_per_step, kernel_ms = tot_ms / args.steps, ker_ms / args.steps
    achieved = units * bytes_per_unit / (kernel_ms * 1e-3) / 1e9

    e2e = None
    if not args.no_e2e:
        n = max(1, min(args.steps, 3))
        e2e_step()
        t0 = time.perf_counter()
        for _ in range(n):
            e2e_step()
        dt = (time.perf_counter() - t0) / n
        e2e = {"value": units / dt / 1e6, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": dt * 1e3, "steps": n}

    cpu = None
    if not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        ref, port = oracle_lib.ref(), oracle_lib.port()
        kind_cpu = "reference" if ref is not None else "port"
        if kind == "terrain":
            # the lattice in full (cheap), then as many vertex rows of the SAME map as the budget allows
            rows = int(max(8, min(nr_v, args.cpu_seconds * 1.9e6 / nr_v)))
            t0 = time.perf_counter()
            map0 = ref.terrain_map0(TSEED, nr_v) if ref is not None else port.terrain_map0(TSEED, nr_v)
            t1 = time.perf_counter()
            if ref is not None:
                buf, _ = ref._boxed(np.ascontiguousarray(maze, np.uint8))
                out = np.zeros((nr_v, nr_v), np.float32)
                ref.lib.ref_terrain_heightmap(TSEED, nr_v, map0.ctypes.data_as(c_void_p), 0.0, buf.ctypes.data + 12, 0,
                                              rows, out.ctypes.data_as(c_void_p))
            else:
                port.terrain_heightmap(map0, 0.0, maze, 0, rows)
            t2 = time.perf_counter()
            per_vertex = (t1 - t0) / units + (t2 - t1) / (rows * nr_v)
            cpu = {"value": 1.0 / per_vertex / 1e6, "unit": unit, "cores": 1, "kind": kind_cpu,
                   "sample": f"lattice map0 in full ({t1 - t0:.1f} s) + vertex rows 0..{rows - 1} of the same {nr_v}^2 map "
                             f"({t2 - t1:.1f} s) on one core of {os.cpu_count()}; per-vertex cost is uniform"}
        elif kind == "mesh":
            # the mesh stage has no reachable reference entry point (it sits inside terrain_init_square_landscape);
            # the oracle port restates it and its normals are pinned to the reference's calc_normal()
            hm = host_map.numpy().reshape(nr_v, nr_v)
            rows = int(max(8, min(nr_v, args.cpu_seconds * 2.0e7 / nr_v)))
            vx = np.zeros((units, 3), np.float32); nm = np.zeros((units, 3), np.float32)
            txx = np.zeros((units, 2), np.float32); ix = np.zeros(quads * 6, np.uint16)
            t0 = time.perf_counter()
            port.lib.ora_terrain_mesh(hm.ctypes.data_as(c_void_p), nr_v, 0.0, 0.0, 0.0, side, 0, rows,
                                      vx.ctypes.data_as(c_void_p), nm.ctypes.data_as(c_void_p),
                                      txx.ctypes.data_as(c_void_p), ix.ctypes.data_as(c_void_p))
            dt = time.perf_counter() - t0
            kind_cpu = "port"
            cpu = {"value": rows * nr_v / dt / 1e6, "unit": unit, "cores": 1, "kind": kind_cpu,
                   "sample": f"vertex rows 0..{rows - 1} of the same {nr_v}^2 mesh, {dt:.1f} s on one core of {os.cpu_count()}"}
        else:
            cs = min(size, 128)
            cperiod = period * cs / size            # same step (= eps) per voxel as the full workload
            t0 = time.perf_counter()
            if ref is not None:
                ref.noise_bake(cs, 4, 2.0, 0.5, cperiod, NSEED)
            else:
                port.noise_bake(cs, 4, 2.0, 0.5, cperiod, NSEED)
            dt = time.perf_counter() - t0
            cpu = {"value": cs ** 3 / dt / 1e6, "unit": unit, "cores": 1, "kind": kind_cpu,
                   "sample": f"noise_grad3d_bake_rgba8({cs}, 4, 2.0, 0.5, {cperiod}, 0xc14d): same lattice step per voxel, "
                             f"{dt:.1f} s on one core of {os.cpu_count()}"}

    line = {
        "metric": metric, "value": units / (ms_per_step * 1e-3) / 1e6, "unit": unit, "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "l2": "a 512 MiB buffer (> L2) is rewritten between timed steps"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(args.workload), "kernel": kernel, "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_update": bytes_per_unit, "peak_source": peak_src,
                     "note": "HBM-bound: one coalesced pass over the map and the output buffers" if kind == "mesh" else
                             "nominally output-bound (HBM), in practice ALU/SFU-bound: see DESIGN.md section 4"},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import clap_b200
    from clap_b200 import _lib
    from clap_b200.rules import ca3d_rule

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    clap_b200.init(local)

    if args.workload in CA2D_WORKLOADS:
        if rank == 0:
            run_ca2d(args, torch, clap_b200, dev, local)        # fits L2 of one GPU: replicas only
        return
    if args.workload in FIELD_WORKLOADS:
        if rank == 0:
            run_fields(args, torch, clap_b200, dev, local)      # independent per point: replicas only
        return
    d0, d1, d2, gens, rule_index = WORKLOADS[args.workload]
    rule = ca3d_rule(rule_index)
    engine = {"auto": 0, "wavefront": 1, "bitplane": 2}[args.engine]

    if world > 1:
        from clap_b200.slab import run_sharded_bench
        return run_sharded_bench(args, WORKLOADS[args.workload], synth_planes)

    cells = d0 * d1 * d2
    updates = cells * gens
    seed_dev = synth_planes(torch, d0, d1, 0, d2, dev)
    torch.cuda.synchronize()
    grid = clap_b200.Grid(d0, d1, d2)

    def step():
        grid.upload(seed_dev.data_ptr())            # device-to-device reset of the state (not timed)
        pop = grid.run3d(rule, gens, engine=engine)
        return pop, grid.stats()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    t_wall0 = time.perf_counter()
    tot_ms, ker_ms, launches, pop, st = 0.0, 0.0, 0, 0, None
    for _ in range(args.steps):
        pop, st = step()
        tot_ms += st["total_ms"]
        ker_ms += st["kernel_ms"]
        launches += st["launches"]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()

    ms_per_step = tot_ms / args.steps
    value = updates / (ms_per_step * 1e-3) / 1e9
    kernel_ms = ker_ms / args.steps
    peak, peak_src = measured_peak()
    achieved = updates * 2.0 / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload), "kernel": "ca3d_sweep_kernel (all generations fused)",
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_update": 2.0, "peak_source": peak_src}

    # ---- end to end through the C ABI with pinned host buffers ------------------------------------
    # clapca_grid_run3d_streamed: host -> device -> host as one pipeline (H2D chunks, pack / sweep / unpack items of
    # ONE launch, D2H chunks); the result is checked against the device-resident run's.
    e2e = None
    if not args.no_e2e:
        host_in = torch.empty(cells, dtype=torch.uint8, pin_memory=True)
        host_out = torch.empty(cells, dtype=torch.uint8, pin_memory=True)
        host_in.copy_(seed_dev.reshape(-1))
        torch.cuda.synchronize()
        n_e2e = max(1, min(args.steps, 3))
        streamed = False
        for i in range(1 + n_e2e):
            if i == 1:
                t0 = time.perf_counter()
            pop_e = grid.run3d_streamed(rule, gens, host_in.data_ptr(), host_out.data_ptr(), max_value=5)
            streamed = grid.stats()["streamed"]
        dt = (time.perf_counter() - t0) / n_e2e
        assert pop_e == pop, "end-to-end population differs from the device-resident run"
        # the device-resident result, downloaded over the (no longer needed) input buffer, must equal the streamed one
        grid.upload(seed_dev.data_ptr())
        grid.run3d(rule, gens, engine=engine)
        grid.download(host_in.data_ptr())
        assert torch.equal(host_in, host_out), "streamed result differs from the device-resident run"
        e2e = {"value": updates / dt / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": cells,
               "d2h_bytes_per_step": cells + 8, "ms_per_step": dt * 1e3, "steps": n_e2e,
               "pipeline": "streamed: H2D chunks | pack+sweep+unpack in one launch | D2H chunks" if streamed
                           else "upload, run, download one after the other",
               "verified_equal_to_resident_run": True}
        del host_in, host_out

    cpu = None
    if not args.no_cpu:
        cpu = cpu_reference_sample(np, rule_index, gens, args.cpu_seconds)

    line = {
        "metric": "ca3d cell-updates/s", "value": value, "unit": "GCUPS", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: ca3d_run {d0}x{d1}x{d2} uint8, {gens} generations, rule {rule.name}, "
                               f"seed P(alive)=1/4 values 1..5",
                   "engine": st["engine"], "planes": st["planes"], "workers": st["workers"],
                   "layout": "pack / unpack fused into the sweep launch" if st["launches"] == 2 else
                             "separate pack / unpack kernels",
                   "l2": "input volume (%.1f GiB) is larger than L2; state is reset from a pristine device copy "
                         "before every step" % (cells / 2 ** 30),
                   "population": pop, "wall_s_timed_region": wall},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
