"""Regenerates the golden fixtures in this directory from oracle/_ref/libclapref.so, i.e. from the
UNMODIFIED reference sources compiled in the build container (make -C oracle ref).

    python tests/golden/make_golden.py

The reference's own tests pin no values on this path (core/test.c:616-670 only check "population != 0"),
so these vectors -- outputs of the reference itself -- are what pins the oracle and the GPU results.
Everything is seeded; rerunning reproduces the files bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402

CA_TEST = dict(born=3 << 2, surv=3 << 7, nr=4, decay=1, neigh=oracle_lib.NEIGH_M1)     # core/terrain.c:391-398


def main():
    ref = oracle_lib.ref()
    port = oracle_lib.port()
    if ref is None:
        raise SystemExit("oracle/_ref/libclapref.so missing: make -C oracle ref (needs /root/reference)")

    # ---- ca2d ---------------------------------------------------------------------------------
    out = {}
    # BASELINE config 1: ca_test, 256 x 256, 5 generations, srand48(1234)
    out["cfg1_final"] = ref.ca2d_generate(side=256, steps=5, seed=1234, **CA_TEST)
    out["cfg1_seed"] = ref.ca2d_generate(side=256, steps=0, seed=1234, **CA_TEST)
    # every neighbourhood, with and without decay, non-power-of-two side, partial sweep
    cases = []
    rules = [(3 << 2, 3 << 7, 4, 1), (0x1e, 0xff, 20, 0), (0x1e0, 0x1f0, 1, 1), (0x0c, 0x1c, 3, 1),
             (0xffffff, 0xffffff, 21, 0)]
    for neigh in range(4):
        for ri, (born, surv, nr, decay) in enumerate(rules):
            seed = 100 + 10 * neigh + ri
            side = 45 + neigh
            start = ref.ca2d_generate(born, surv, nr, decay, neigh, side, 0, seed)
            if ri == 1:     # give the value-comparing neighbourhoods a spread of values
                rng = np.random.default_rng(seed)
                start = (rng.integers(0, 6, start.shape) * (rng.random(start.shape) < 0.6)).astype(np.uint8)
            final = ref.ca2d_step(start.copy(), born, surv, nr, decay, neigh, steps=6)
            part = ref.ca2d_step(start.copy(), born, surv, nr, decay, neigh, side=side - 7, steps=3)
            cases.append((neigh, born, surv, nr, decay, side))
            out[f"case{len(cases) - 1}_start"] = start
            out[f"case{len(cases) - 1}_final6"] = final
            out[f"case{len(cases) - 1}_partial3"] = part
    out["cases"] = np.array(cases, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "ca2d.npz"), **out)

    # ---- ca3d ---------------------------------------------------------------------------------
    out = {}
    seed_small = ref.ca3d_make(16, 8, 4, seed=42)                # the shape of core/test.c:628
    out["make_16_8_4"] = seed_small
    fin = seed_small.copy()
    out["make_16_8_4_coral4_pop"] = np.int64(ref.ca3d_run(fin, 7, 4))
    out["make_16_8_4_coral4"] = fin
    seed_mid = ref.ca3d_make(37, 23, 32, seed=43)
    dense = seed_mid.copy()
    st = port.srand48(44)
    flat = dense.reshape(-1)
    for i in range(flat.size):                                    # SURVEY 8(d) "seed B": denser start
        if flat[i] == 0 and port.lrand48(st) % 4 == 0:
            flat[i] = 1 + port.lrand48(st) % 5
    out["seedA_37_23_32"] = seed_mid
    out["seedB_37_23_32"] = dense
    pops = []
    for nca in range(9):
        for tag, start in (("A", seed_mid), ("B", dense)):
            fin = start.copy()
            pops.append(ref.ca3d_run(fin, nca, 5))
            out[f"rule{nca}_seed{tag}_5gen"] = fin
    out["pops_37_23_32"] = np.array(pops, dtype=np.int64)
    # BASELINE config 2: 128^3, 10 generations, coral, srand48(42) ca3d_make seed -> fingerprints only
    ref.ca3d_make(16, 8, 4, seed=42)
    big = ref.ca3d_make(128, 128, 128, seed=42)
    out["cfg2_seed_hash"] = np.uint64(port.fnv(big))
    out["cfg2_seed_hist"] = np.bincount(big.ravel(), minlength=256).astype(np.int64)
    out["cfg2_pop"] = np.int64(ref.ca3d_run(big, 7, 10))
    out["cfg2_final_hash"] = np.uint64(port.fnv(big))
    out["cfg2_final_hist"] = np.bincount(big.ravel(), minlength=256).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "ca3d.npz"), **out)

    # ---- noise --------------------------------------------------------------------------------
    out = {}
    out["bake_16_p5"] = ref.noise_bake(16, 4, 2.0, 0.5, 5.0, 0xC14D)          # non-integer step 5/16
    out["bake_24_p37_o3"] = ref.noise_bake(24, 3, 2.3, 0.45, 37.0, 99)
    default = ref.noise_bake(64, 4, 2.0, 0.5, 64.0, 0xC14D)                   # engine default, noise.c:309-317
    out["bake_64_default_hash"] = np.uint64(port.fnv(default))
    out["bake_64_default_first"] = default.reshape(-1)[:64].copy()
    rng = np.random.default_rng(5)
    pts = (rng.random((4096, 3)) * 90 - 10).astype(np.float32)
    out["fbm_points"] = pts
    out["fbm_values"] = ref.fbm3(pts, 4, 2.0, 0.5, 37, 0xC14D)
    out["hash31_1_2_3_7"] = np.float32(ref.lib.ref_hash31(1, 2, 3, 7))
    np.savez_compressed(os.path.join(HERE, "noise.npz"), **out)

    # ---- terrain ------------------------------------------------------------------------------
    out = {}
    for seed in (12345, -99, (1 << 40) + 17):
        out[f"map0_64_seed{seed}"] = ref.terrain_map0(seed, 64)
    nr_v = 128
    map0 = ref.terrain_map0(12345, nr_v)
    maze = ref.ca2d_generate(side=nr_v // 8, steps=4, seed=7, **CA_TEST)
    out["maze_16"] = maze
    out["field_128"] = ref.terrain_field(12345, map0, 0.0, 1.0, 4)
    out["field_128_y3_amp2_o3"] = ref.terrain_field(12345, map0, 3.0, 2.5, 3)
    out["heightmap_128"] = ref.terrain_heightmap(12345, map0, 0.0, maze)
    m1024 = ref.terrain_map0(12345, 1024)
    f1024 = ref.terrain_field(12345, m1024, 0.0, 1.0, 4)
    out["survey_map0_7"] = np.float32(m1024.reshape(-1)[7])                   # SURVEY.md 8(c) spot values
    out["survey_map_12345"] = np.float32(f1024.reshape(-1)[12345])
    out["survey_map_sum"] = np.float64(f1024.astype(np.float64).sum())
    np.savez_compressed(os.path.join(HERE, "terrain.npz"), **out)

    # ---- terrain mesh stage: the reference's calc_normal() over the golden heightmap -------
    mesh = {"normals_128": ref.terrain_normals(out["heightmap_128"])}
    rough = (np.random.default_rng(3).random((37, 37)) * 40 - 20).astype(np.float32)
    mesh["rough_37"] = rough
    mesh["normals_rough_37"] = ref.terrain_normals(rough)
    # terrain_height() (terrain.c:336-379) of the reference at seeded points, inside and outside the square
    pts = (np.random.default_rng(4).random((400, 2)) * 340 - 20).astype(np.float32) + np.float32([10.0, 5.0])
    mesh["height_pts"] = pts
    mesh["heights_128_x10_z5_side300"] = ref.terrain_height(out["heightmap_128"], 10.0, 5.0, 300, pts)
    np.savez_compressed(os.path.join(HERE, "terrain_mesh.npz"), **mesh)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
