"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same seeded inputs, against
the committed golden vectors, and -- at sizes the oracle cannot reach in seconds -- through engine cross-checks
and size-independent properties.  CA results must be bit-exact; float fields carry their tolerance here."""
import os

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WAVEFRONT, BITPLANE, DIAGONAL = 1, 2, 3
ENGINES = [pytest.param(WAVEFRONT, id="wavefront"), pytest.param(BITPLANE, id="bitplane")]


def synth(rng, shape, p_alive=0.25, vmax=5, with255=False):
    vol = (rng.integers(1, vmax + 1, shape) * (rng.random(shape) < p_alive)).astype(np.uint8)
    if with255:
        vol[rng.random(shape) < 0.003] = 255
    return vol


# ---- ca3d -----------------------------------------------------------------------------------------

@pytest.mark.parametrize("engine", ENGINES)
def test_ca3d_golden_all_rules(gpu, engine):
    g = np.load(os.path.join(G, "ca3d.npz"))
    pops = g["pops_37_23_32"].tolist()
    k = 0
    for nca in range(9):
        for tag in "AB":
            vol = g[f"seed{tag}_37_23_32"].copy()
            pop = gpu.ca3d_run(vol, nca, 5, engine=engine)
            assert pop == pops[k], (nca, tag)
            assert np.array_equal(vol, g[f"rule{nca}_seed{tag}_5gen"]), (nca, tag)
            k += 1


@pytest.mark.parametrize("engine", ENGINES)
def test_ca3d_test0_shape(gpu, engine):
    """The reference's own test shape: ca3d_make(16, 8, 4) then 4 steps of ca_coral (core/test.c:628-633)."""
    g = np.load(os.path.join(G, "ca3d.npz"))
    vol = g["make_16_8_4"].copy()
    pop = gpu.ca3d_run(vol, 7, 4, engine=engine)
    assert pop == int(g["make_16_8_4_coral4_pop"]) and pop != 0
    assert np.array_equal(vol, g["make_16_8_4_coral4"])


@pytest.mark.parametrize("engine", ENGINES)
def test_ca3d_cfg2_128cube_10gen(gpu, oracle, engine):
    """BASELINE config 2: 128^3, 10 generations, coral, on the srand48(42) ca3d_make seed."""
    g = np.load(os.path.join(G, "ca3d.npz"))
    oracle.ca3d_make(16, 8, 4, 42)
    vol = oracle.ca3d_make(128, 128, 128, 42)
    assert oracle.fnv(vol) == int(g["cfg2_seed_hash"])
    pop = gpu.ca3d_run(vol, 7, 10, engine=engine)
    assert pop == int(g["cfg2_pop"]) == 272840
    assert oracle.fnv(vol) == int(g["cfg2_final_hash"])


def _seed_b(oracle, vol, seed):
    """SURVEY 8(d) cfg 2 'seed B' (denser): every zero cell, in memory order, becomes 1 + lrand48() % 5 with
    probability 1/4 -- two draws of one lrand48 stream per hit, one per miss"""
    from clap_b200.ca import Rand48
    out = vol.copy().reshape(-1)
    zero = np.flatnonzero(out == 0)
    # draw k decides zero cell k; a hit consumes one more draw: resolve the data-dependent stream in blocks
    rng = Rand48(seed)
    draws = rng.lrand48_block(2 * len(zero) + 2)
    i = k = 0
    while k < len(zero):
        if draws[i] % 4 == 0:
            out[zero[k]] = 1 + draws[i + 1] % 5
            i += 2
        else:
            i += 1
        k += 1
    return out.reshape(vol.shape)


_CFG2_SEEDS = {}


def _cfg2_seed(oracle, variant):
    if not _CFG2_SEEDS:
        a = oracle.ca3d_make(128, 128, 128, 42)
        _CFG2_SEEDS["A255"] = a.copy()
        _CFG2_SEEDS["A255"][np.random.default_rng(255).random(a.shape) < 0.002] = 255
        _CFG2_SEEDS["A"] = np.minimum(a, 5)     # seed A without ca3d_prune's 255 marks (those are the A255 case)
        _CFG2_SEEDS["B"] = _seed_b(oracle, _CFG2_SEEDS["A"], 43)
    return _CFG2_SEEDS[variant].copy()


@pytest.mark.parametrize("variant", ["A", "B", "A255"])
@pytest.mark.parametrize("nca", range(9))
def test_ca3d_cfg2_128cube_10gen_every_rule_and_seed(gpu, oracle, nca, variant):
    """SURVEY 7.2 exit criterion / BASELINE config 2 in full: 128^3 x 10 generations for ALL nine rules, on seed A
    (srand48(42); ca3d_make), the denser seed B, and seed A with 255s injected -- against the oracle port."""
    vol = _cfg2_seed(oracle, variant)
    want = vol.copy()
    s, b, n = oracle.ca3d_rule(nca)
    wpop = oracle.ca3d_run(want, s, b, n, 10)
    pop = gpu.ca3d_run(vol, nca, 10, engine=BITPLANE)
    assert pop == wpop and np.array_equal(vol, want), (nca, variant)


def test_ca3d_test0_loop_over_all_nine_rules(gpu, oracle):
    """core/test.c:625-637 loops CA3D_MAX times over ca3d_make(16, 8, 4) + ca3d_run(4 steps) and asks for a non-empty
    result; here every nca of that loop (and the wrap-around indices ca3d.c:126 allows) against the oracle"""
    for nca in list(range(9)) + [9, 16, -1]:
        vol = oracle.ca3d_make(16, 8, 4, 7 + (nca % 9))
        want = vol.copy()
        s, b, n = oracle.ca3d_rule(nca)
        wpop = oracle.ca3d_run(want, s, b, n, 4)
        for engine in (WAVEFRONT, BITPLANE):
            got = vol.copy()
            assert gpu.ca3d_run(got, nca, 4, engine=engine) == wpop
            assert np.array_equal(got, want), (nca, engine)
    assert gpu.ca3d_rule(-1).name == "ca_spiky_growth"        # (size_t)-1 % 9 == 6, as in the reference


@pytest.mark.slow
def test_ca3d_cfg4_chain_1024cube_2_generations_vs_port(gpu, oracle):
    """Link (i) of the config-4 parity chain at its upper end: 1024^3 = 2^30 cells -- the largest cube the
    reference's 32-bit index can address -- for 2 generations against the oracle port (~40 s of CPU)."""
    import torch
    from clap_b200 import synth as seedgen
    side = 1024
    vol = seedgen.synth_torch(torch, side, side, 0, side, "cuda:0").cpu().numpy()
    want = vol.copy()
    s, b, n = oracle.ca3d_rule(7)
    wpop = oracle.ca3d_run(want, s, b, n, 2)
    pop = gpu.ca3d_run(vol, 7, 2, engine=BITPLANE)
    assert pop == wpop
    assert np.array_equal(vol, want)


def test_ca3d_benched_volume_bottom_planes_equal_the_unmodified_reference(gpu):
    """Links (ii)/(iii) anchor: planes 0..7 of the BENCHED 2048^3 volume after 50 generations are determined by its
    bottom 58 planes (plane z of generation g sees z+1 of g-1 and nothing above).  The unmodified reference computed
    them once (tests/golden/make_golden_cfg4_planes.py); the GPU runs the same 58-plane slab here, bench.py checks the
    same fingerprints on the full volume at every N."""
    import json
    import torch
    from clap_b200 import synth as seedgen
    from clap_b200.ca import hash_planes
    with open(os.path.join(G, "cfg4_planes_2048.json")) as f:
        a = json.load(f)
    side, k, gens = a["side"], a["planes"], a["generations"]
    seed = seedgen.synth_torch(torch, side, side, 0, k + gens, "cuda:0")
    grid = gpu.Grid(side, side, k + gens)
    assert ["%016x" % int(h) for h in hash_planes(seed.data_ptr(), side * side, k)] == a["seed_plane_hashes"]
    grid.upload(seed.data_ptr())
    grid.run3d(a["nca"], gens)
    got = ["%016x" % int(h) for h in hash_planes(grid.device_ptr(), side * side, k)]
    grid.close()
    assert got == a["plane_hashes"]


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("shape", [(1, 1, 1), (1, 7, 3), (5, 1, 9), (9, 5, 1), (31, 6, 5), (32, 6, 5), (33, 6, 5),
                                   (100, 17, 11), (130, 9, 40)])
def test_ca3d_ragged_shapes_vs_oracle(gpu, oracle, engine, shape):
    d0, d1, d2 = shape
    rng = np.random.default_rng(d0 * 1000 + d1 * 10 + d2)
    for nca in (0, 3, 6, 7, 8):
        vol = synth(rng, (d2, d1, d0), 0.4, 6, with255=True)
        want = vol.copy()
        s, b, n = oracle.ca3d_rule(nca)
        wpop = oracle.ca3d_run(want, s, b, n, 6)
        pop = gpu.ca3d_run(vol, nca, 6, engine=engine)
        assert pop == wpop and np.array_equal(vol, want), (shape, nca)


def test_ca3d_custom_masks_run_time_rule(gpu, oracle):
    """Masks outside cas[] go through the run-time-table kernel."""
    rng = np.random.default_rng(11)
    for _ in range(4):
        rule = gpu.CellAutomaton("custom", born_mask=int(rng.integers(0, 1 << 27)) & int(rng.integers(0, 1 << 27)),
                                 surv_mask=int(rng.integers(0, 1 << 27)), nr_states=int(rng.integers(1, 200)))
        vol = synth(rng, (9, 14, 50), 0.5, 9)
        want = vol.copy()
        wpop = oracle.ca3d_run(want, rule.surv_mask, rule.born_mask, rule.nr_states, 5)
        assert gpu.ca3d_run(vol, rule, 5, engine=BITPLANE) == wpop
        assert np.array_equal(vol, want)


@pytest.mark.parametrize("flip", [None, (7, 14)])
def test_ca3d_chain_rules_long_dependency_runs(gpu, oracle, flip):
    """Rules whose table flips between every K and K+1 on a binary volume: every cell depends on its in-row
    predecessor, so a row is one dependency chain (or a few, with two table entries flipped back) -- the scan's
    rare paths (runs of >= 8 dependent cells in a word, lanes that never break the chain) carry the whole result."""
    surv, born = 0x5555555, 0x2aaaaaa
    if flip:
        surv ^= 1 << flip[0]
        born ^= 1 << flip[1]
    rule = gpu.CellAutomaton("chain", born_mask=born, surv_mask=surv, nr_states=2)
    rng = np.random.default_rng(77)
    for shape in ((6, 9, 300), (4, 7, 1030), (3, 6, 2048), (5, 5, 45)):
        vol = (rng.random(shape) < 0.5).astype(np.uint8)
        want = vol.copy()
        wpop = oracle.ca3d_run(want, surv, born, 2, 4)
        assert gpu.ca3d_run(vol, rule, 4, engine=BITPLANE) == wpop, shape
        assert np.array_equal(vol, want), shape


def test_ca3d_wide_rows_thin_slab_vs_oracle(gpu, oracle):
    """Rows of the BASELINE config-4 width (2048 cells = 2 words per lane) on a thin slab the oracle finishes."""
    rng = np.random.default_rng(12)
    for d0 in (1024, 1500, 2048, 4096):
        vol = synth(rng, (6, 40, d0))
        want = vol.copy()
        s, b, n = oracle.ca3d_rule(7)
        wpop = oracle.ca3d_run(want, s, b, n, 4)
        assert gpu.ca3d_run(vol, 7, 4, engine=BITPLANE) == wpop
        assert np.array_equal(vol, want), d0


# ---- ca3d: layout items of the sweep launch (fused pack / unpack, streamed host -> device -> host run) ----

def _pinned(shape):
    import torch
    t = torch.empty(shape, dtype=torch.uint8, pin_memory=True)
    return t, t.numpy()


@pytest.mark.parametrize("shape,nca,steps,chunk_planes", [
    ((45, 20, 12), 7, 5, 1), ((64, 33, 7), 0, 9, 2), ((33, 6, 5), 3, 6, 0), ((100, 17, 11), 6, 6, 3),
    ((1500, 12, 6), 7, 4, 2), ((2048, 40, 18), 7, 4, 4), ((4096, 8, 5), 8, 3, 1), ((1, 1, 1), 7, 3, 0),
])
def test_ca3d_streamed_vs_oracle(gpu, oracle, monkeypatch, shape, nca, steps, chunk_planes):
    """clapca_grid_run3d_streamed with page-locked buffers: H2D chunks, pack / sweep / unpack items and D2H chunks
    overlap inside one launch; the result must be the oracle's, for one-plane chunks as well as for one big one."""
    if chunk_planes:
        monkeypatch.setenv("CLAPCA_IO_CHUNK_PLANES", str(chunk_planes))
    d0, d1, d2 = shape
    rng = np.random.default_rng(d0 + 7 * d1 + 31 * d2)
    keep_in, host_in = _pinned((d2, d1, d0))
    keep_out, host_out = _pinned((d2, d1, d0))
    host_in[...] = synth(rng, (d2, d1, d0), 0.4, 6)
    host_out[...] = 0xEE
    want = host_in.copy()
    s, b, n = oracle.ca3d_rule(nca)
    wpop = oracle.ca3d_run(want, s, b, n, steps)
    grid = gpu.Grid(d0, d1, d2)
    for _ in range(2):                              # twice: the epoch / counters of a grid are reused
        pop = grid.run3d_streamed(nca, steps, host_in, host_out, max_value=6)
        assert grid.stats()["streamed"]
        assert pop == wpop and np.array_equal(host_out, want), (shape, nca)
        host_out[...] = 0xEE
    grid.close()


def test_ca3d_streamed_in_place_255_and_bound_violation(gpu, oracle, monkeypatch):
    monkeypatch.setenv("CLAPCA_IO_CHUNK_PLANES", "2")
    oracle.ca3d_make(16, 8, 4, 42)
    seed = oracle.ca3d_make(48, 12, 10, 42)         # ca3d_prune leaves 255s here: 8 state planes
    assert int((seed == 255).sum()) > 0
    keep, host = _pinned(seed.shape)
    host[...] = seed
    want = seed.copy()
    s, b, n = oracle.ca3d_rule(7)
    wpop = oracle.ca3d_run(want, s, b, n, 6)
    grid = gpu.Grid(48, 12, 10)
    assert grid.run3d_streamed(7, 6, host, host, max_value=255) == wpop        # in place
    assert np.array_equal(host, want)
    host[...] = seed
    with pytest.raises(gpu.ClapcaError) as ei:      # a bound below the data: detected by the pack items, never wrong
        grid.run3d_streamed(7, 6, host, host, max_value=5)
    assert ei.value.status == 2
    host[...] = seed                                # the grid stays usable
    assert grid.run3d_streamed(7, 6, host, host, max_value=255) == wpop
    assert np.array_equal(host, want)
    grid.close()


def test_ca3d_streamed_pageable_buffers_take_the_sequential_path(gpu, oracle):
    rng = np.random.default_rng(77)
    vol = synth(rng, (9, 14, 50))
    want = vol.copy()
    s, b, n = oracle.ca3d_rule(7)
    wpop = oracle.ca3d_run(want, s, b, n, 5)
    out = np.zeros_like(vol)
    grid = gpu.Grid(50, 14, 9)
    assert grid.run3d_streamed(7, 5, vol, out, max_value=5) == wpop
    assert not grid.stats()["streamed"]
    assert np.array_equal(out, want)
    grid.close()


@pytest.mark.parametrize("team", ["16", "0"])
def test_ca3d_streamed_matches_resident_run_medium(gpu, monkeypatch, team):
    """512 x 512 x 192, 20 generations: beyond the oracle's reach in seconds -- the streamed pipeline (many chunks,
    all SMs busy, team and one-warp-per-sweep kernels) against the device-resident run with separate layout kernels."""
    monkeypatch.setenv("CLAPCA_TEAM", team)
    monkeypatch.setenv("CLAPCA_IO_CHUNK_PLANES", "8")
    rng = np.random.default_rng(5)
    shape = (192, 512, 512)
    keep_in, host_in = _pinned(shape)
    keep_out, host_out = _pinned(shape)
    host_in[...] = synth(rng, shape)
    ref = host_in.copy()
    rpop = gpu.ca3d_run(ref, 7, 20, engine=BITPLANE)
    grid = gpu.Grid(512, 512, 192)
    assert grid.run3d_streamed(7, 20, host_in, host_out, max_value=5) == rpop
    assert np.array_equal(host_out, ref)
    grid.close()


@pytest.mark.parametrize("team", ["16", "0"])
def test_ca3d_fused_layout_resident_vs_oracle(gpu, oracle, monkeypatch, team):
    """CLAPCA_FUSED_LAYOUT=1: the device-resident run converts the layout inside the sweep launch too."""
    monkeypatch.setenv("CLAPCA_FUSED_LAYOUT", "1")
    monkeypatch.setenv("CLAPCA_TEAM", team)
    rng = np.random.default_rng(21)
    for shape, nca in (((100, 17, 11), 7), ((64, 33, 40), 2), ((2048, 12, 6), 7)):
        d0, d1, d2 = shape
        vol = synth(rng, (d2, d1, d0), 0.4, 6, with255=(nca == 2))
        want = vol.copy()
        s, b, n = oracle.ca3d_rule(nca)
        wpop = oracle.ca3d_run(want, s, b, n, 6)
        assert gpu.ca3d_run(vol, nca, 6, engine=BITPLANE) == wpop
        assert np.array_equal(vol, want), shape


@pytest.mark.parametrize("dims", [(48, 24, 20), (5, 5, 5), (4, 4, 4), (3, 3, 3), (2, 2, 2), (1, 1, 1), (3, 1, 7),
                                  (33, 17, 9), (64, 64, 64), (9, 40, 12)])
def test_ca3d_make_on_the_device_equals_the_reference_builder(gpu, oracle, dims):
    """clapca_grid_make3d: faces + random walk + prune of ca3d_make() (core/ca3d.c:41-99, 144-169) with the volume on
    the device -- the walk runs on the host against a sparse picture, the prune is a fixed-point of a parallel kernel
    (an enclosed EMPTY cell becomes 255 and counts for the cells after it only, like in the reference's sweep)."""
    from ctypes import c_uint64
    from clap_b200.ca import Rand48
    d0, d1, d2 = dims
    for seed in (42, 7, 20260101):
        want = oracle.ca3d_make(d0, d1, d2, seed)
        st = oracle.srand48(seed)
        chk = np.zeros((d2, d1, d0), np.uint8)
        oracle.lib.ora_ca3d_make(chk.ctypes.data, d0, d1, d2, __import__("ctypes").byref(st))
        assert np.array_equal(chk, want)
        g = gpu.Grid(d0, d1, d2)
        rng = Rand48(seed)
        pop = g.make3d(rng)
        got = np.empty((d2, d1, d0), np.uint8)
        g.download(got)
        g.close()
        assert np.array_equal(got, want), (dims, seed, int((got != want).sum()))
        assert pop == int((want != 0).sum())
        assert rng.x == int(st.value), "the lrand48 stream continues where the reference's walk leaves it"


def test_ca3d_make_prune_marks_enclosed_empty_cells_for_their_successors_only(gpu, oracle):
    """the order-dependent corner of ca3d_prune(): an EMPTY cell with six occupied neighbours becomes 255 and counts as
    occupied for the cells the sweep visits after it.  At these sizes a quarter of all seeds enclose such a cell (checked
    against a re-run of the walk when this test was written), chain reactions included: the device prune's fixed point
    must equal the oracle's sequential sweep for every one of them."""
    from clap_b200.ca import Rand48
    flips = 0
    for dims in ((6, 6, 6), (8, 8, 8), (7, 9, 8), (10, 10, 10)):
        g = gpu.Grid(*dims)
        got = np.empty(dims[::-1], np.uint8)
        for seed in range(1, 41):
            want = oracle.ca3d_make(*dims, seed)
            pop = g.make3d(Rand48(seed))
            g.download(got)
            assert np.array_equal(got, want), (dims, seed)
            assert pop == int((want != 0).sum())
            flips += int((want[1:-1, 1:-1, 1:-1] == 255).sum())
        g.close()
    assert flips > 100


def test_ca3d_zero_steps_and_population(gpu):
    rng = np.random.default_rng(13)
    vol = synth(rng, (8, 9, 10))
    keep = vol.copy()
    assert gpu.ca3d_run(vol, 7, 0) == int(np.count_nonzero(keep))
    assert np.array_equal(vol, keep)


def test_ca3d_engines_agree_256cube(gpu):
    """No CPU in the loop: the two independent GPU engines must agree bit for bit at 256^3 x 6."""
    rng = np.random.default_rng(14)
    vol = synth(rng, (256, 256, 256))
    a, b = vol.copy(), vol.copy()
    pa = gpu.ca3d_run(a, 7, 6, engine=WAVEFRONT)
    pb = gpu.ca3d_run(b, 7, 6, engine=BITPLANE)
    assert pa == pb == int(np.count_nonzero(a))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("team,tile_gens", [(1, 1), (3, 1), (16, 1), (16, 2), (16, 4), (6, 2), (4, 2)])
def test_ca3d_tile_mode_vs_oracle(gpu, oracle, monkeypatch, team, tile_gens):
    """Tile mode (CLAPCA_TEAM compute warps per CTA, CLAPCA_TILE_GENS generations per tile: one CTA sweeps nz planes
    x ng generations, warps follow each other through shared-memory row counters, the service warp publishes the
    gpu-scope counters) against the oracle on ragged shapes, 255-valued cells and every plane count."""
    monkeypatch.setenv("CLAPCA_TEAM", str(team))
    monkeypatch.setenv("CLAPCA_TILE_GENS", str(tile_gens))
    rng = np.random.default_rng(400 + team)
    for shape, nca in (((1, 1, 1), 7), ((33, 6, 5), 0), ((100, 17, 11), 3), ((130, 9, 40), 7), ((70, 40, 37), 6),
                       ((1500, 24, 9), 7), ((2048, 30, 20), 8), ((4096, 12, 5), 7), ((64, 20, 35), 2)):
        d0, d1, d2 = shape
        vol = synth(rng, (d2, d1, d0), 0.4, 6, with255=(nca == 6))
        want = vol.copy()
        s, b, n = oracle.ca3d_rule(nca)
        wpop = oracle.ca3d_run(want, s, b, n, 7)
        pop = gpu.ca3d_run(vol, nca, 7, engine=BITPLANE)
        assert pop == wpop and np.array_equal(vol, want), (team, shape, nca)


def test_ca3d_team_mode_agrees_with_sweep_mode_512(gpu, monkeypatch):
    """No CPU in the loop: team mode == one-warp-per-sweep mode at 2048 x 512 x 300, 12 generations."""
    rng = np.random.default_rng(16)
    vol = synth(rng, (300, 512, 2048))
    a, b = vol.copy(), vol.copy()
    monkeypatch.setenv("CLAPCA_TEAM", "0")
    pa = gpu.ca3d_run(a, 7, 12, engine=BITPLANE)
    monkeypatch.setenv("CLAPCA_TEAM", "16")
    monkeypatch.setenv("CLAPCA_TILE_GENS", "1")
    pb = gpu.ca3d_run(b, 7, 12, engine=BITPLANE)
    assert pa == pb == int(np.count_nonzero(a))
    assert np.array_equal(a, b)
    for tg in ("2", "4"):                  # 8 x 2 and 4 x 4 tiles: generations of a row a few rows apart
        c = vol.copy()
        monkeypatch.setenv("CLAPCA_TILE_GENS", tg)
        assert gpu.ca3d_run(c, 7, 12, engine=BITPLANE) == pa
        assert np.array_equal(a, c), tg


def test_ca3d_generations_compose_large(gpu):
    """Size-independent property at 1024 x 1024 x 96: G1 then G2 generations == G1 + G2 in one fused run."""
    rng = np.random.default_rng(15)
    vol = synth(rng, (96, 1024, 1024))
    grid = gpu.Grid(1024, 1024, 96)
    out1 = np.empty_like(vol)
    out2 = np.empty_like(vol)
    grid.upload(vol)
    p1 = grid.run3d(7, 9, engine=BITPLANE)
    grid.download(out1)
    grid.upload(vol)
    grid.run3d(7, 4, engine=BITPLANE)
    p2 = grid.run3d(7, 5, engine=BITPLANE)
    grid.download(out2)
    grid.close()
    assert p1 == p2 == int(np.count_nonzero(out1))
    assert np.array_equal(out1, out2)


def test_ca3d_device_resident_grid_stats(gpu, oracle):
    rng = np.random.default_rng(16)
    vol = synth(rng, (20, 30, 70))
    want = vol.copy()
    s, b, n = oracle.ca3d_rule(0)
    wpop = oracle.ca3d_run(want, s, b, n, 3)
    grid = gpu.Grid(70, 30, 20)
    grid.upload(vol)
    assert grid.run3d(0, 3) == wpop == grid.count()
    st = grid.stats()
    assert st["engine"] == "bitplane" and st["planes"] == 3 and st["kernel_ms"] > 0 and st["launches"] >= 3
    got = np.empty_like(vol)
    grid.download(got)
    grid.close()
    assert np.array_equal(got, want)


# ---- ca2d -----------------------------------------------------------------------------------------

def test_ca2d_golden_cases(gpu):
    g = np.load(os.path.join(G, "ca2d.npz"))
    for i, (neigh, born, surv, nr, decay, side) in enumerate(g["cases"].tolist()):
        ca = gpu.CellAutomaton("case", born, surv, nr, bool(decay), neigh)
        arr = g[f"case{i}_start"].copy()
        gpu.ca2d_step(ca, arr, steps=6)
        assert np.array_equal(arr, g[f"case{i}_final6"]), (i, neigh)
        arr = g[f"case{i}_start"].copy()
        gpu.ca2d_step(ca, arr, side=side - 7, steps=3)
        assert np.array_equal(arr, g[f"case{i}_partial3"]), (i, neigh)


def test_ca2d_cfg1_cave_generation(gpu):
    """BASELINE config 1: ca2d_generate(&ca_test, 256, 5) after srand48(1234)."""
    from clap_b200.ca import Rand48
    g = np.load(os.path.join(G, "ca2d.npz"))
    arr = gpu.ca2d_generate(gpu.CA_TEST, 256, 5, Rand48(1234))
    assert np.array_equal(arr, g["cfg1_final"])


def test_ca2d_instantiator_rules_and_steps_compose(gpu, oracle):
    rng = np.random.default_rng(21)
    maze = oracle.ca2d_run(oracle.ca2d_seed(128, 4, 7), 3 << 2, 3 << 7, 4, 1, oracle_lib.NEIGH_M1, 4)
    want = maze.copy()
    for ca in gpu.CA_INSTORS:           # core/terrain.c:473-477: one step of each instantiator rule
        gpu.ca2d_step(ca, maze)
        oracle.ca2d_run(want, ca.born_mask, ca.surv_mask, ca.nr_states, ca.decay, ca.neigh, 1)
    assert np.array_equal(maze, want)
    arr = synth(rng, (300, 300), 0.6, 4)
    a, b = arr.copy(), arr.copy()
    gpu.ca2d_step(gpu.CA_TEST, a, steps=7)
    gpu.ca2d_step(gpu.CA_TEST, b, steps=3)
    gpu.ca2d_step(gpu.CA_TEST, b, steps=4)
    assert np.array_equal(a, b)
    assert np.array_equal(a, oracle.ca2d_run(arr.copy(), 3 << 2, 3 << 7, 4, 1, oracle_lib.NEIGH_M1, 7))


def test_ca2d_rectangular_and_empty(gpu, oracle):
    rng = np.random.default_rng(22)
    arr = synth(rng, (37, 91), 0.5, 4)           # h = 37, w = 91; side sweeps min(side, extent)
    want = oracle.ca2d_run(arr.copy(), 3 << 2, 3 << 7, 4, 1, oracle_lib.NEIGH_MV, 3, side=60)
    ca = gpu.CellAutomaton("r", 3 << 2, 3 << 7, 4, True, oracle_lib.NEIGH_MV)
    gpu.ca2d_step(ca, arr, side=60, steps=3)
    assert np.array_equal(arr, want)
    z = np.zeros((16, 16), np.uint8)
    gpu.ca2d_step(gpu.CA_TEST, z, steps=5)
    assert not z.any()


# ---- ca2d bit-plane engine (transposed bit planes, one CTA per generation) ---------------------------------

CAVE_BIN = dict(born=0x1E0, surv=0x1F0, nr=1, decay=True)          # SURVEY 8(d) cfg 3: binary cave smoothing


def _ca(gpu, born, surv, nr, decay, neigh):
    return gpu.CellAutomaton("t", born, surv, nr, bool(decay), neigh)


@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (9, 1), (31, 33), (64, 64), (100, 37), (257, 130), (1025, 40),
                                   (2100, 24), (4097, 9)])
def test_ca2d_bitplane_vs_oracle_shapes(gpu, oracle, shape):
    """Every alive-bit rule family on ragged shapes (h = cells per engine row, crossing lane and warp spans)."""
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    rules = [
        (3 << 2, 3 << 7, 4, 1, oracle_lib.NEIGH_M1, 4),                     # ca_test, 3 planes
        (0x1E0, 0x1F0, 1, 1, oracle_lib.NEIGH_M1, 1),                       # binary, 1 plane
        (0x6, 0x1C, 2, 1, oracle_lib.NEIGH_VN1, 3),                         # von Neumann
        (0x1E, 0xFF, 20, 0, oracle_lib.NEIGH_MV, 20),                       # ca_instors[0]: mv without decay
        (0xFFFFFF, 0xFFFFFF, 21, 0, oracle_lib.NEIGH_VNV, 9),               # vnv without decay
        (0x0C, 0x03, 300, 1, oracle_lib.NEIGH_M1, 255),                     # nr_states wraps to 44; 255-valued cells
    ]
    for born, surv, nr, decay, neigh, vmax in rules:
        arr = synth(rng, shape, 0.45, vmax)
        want = oracle.ca2d_run(arr.copy(), born, surv, nr, decay, neigh, 5, side=max(shape))
        gpu.ca2d_step(_ca(gpu, born, surv, nr, decay, neigh), arr, side=max(shape), steps=5, engine=BITPLANE)
        assert np.array_equal(arr, want), (shape, born, surv, nr, neigh)


def test_ca2d_bitplane_many_generations_more_than_resident_ctas(gpu, oracle):
    """700 generations on a small grid: more generations than CTAs the device keeps resident."""
    rng = np.random.default_rng(5)
    arr = synth(rng, (96, 80), 0.5, 1)
    want = oracle.ca2d_run(arr.copy(), 0x8, 0xC, 1, 1, oracle_lib.NEIGH_M1, 700, side=96)      # Life: B3/S23
    gpu.ca2d_step(_ca(gpu, 0x8, 0xC, 1, 1, oracle_lib.NEIGH_M1), arr, side=96, steps=700, engine=BITPLANE)
    assert np.array_equal(arr, want)
    assert arr.any()


def test_ca2d_engines_agree_1024(gpu):
    rng = np.random.default_rng(6)
    arr = synth(rng, (1024, 1024), 0.55, 4)
    a, b = arr.copy(), arr.copy()
    gpu.ca2d_step(gpu.CA_TEST, a, steps=12, engine=BITPLANE)
    gpu.ca2d_step(gpu.CA_TEST, b, steps=12, engine=WAVEFRONT)
    assert np.array_equal(a, b)


def test_ca2d_bitplane_rejects_what_it_cannot_run(gpu):
    arr = np.ones((64, 64), np.uint8)
    with pytest.raises(gpu.ClapcaError):          # partial sweep
        gpu.ca2d_step(gpu.CA_TEST, arr, side=40, steps=1, engine=BITPLANE)
    with pytest.raises(gpu.ClapcaError):          # value-comparing neighbourhood with decay
        gpu.ca2d_step(_ca(gpu, 0xC, 0x180, 4, True, oracle_lib.NEIGH_MV), arr, steps=1, engine=BITPLANE)


def test_ca2d_cfg3_full_width_strip_vs_oracle(gpu, oracle):
    """BASELINE config 3 geometry in one dimension: rows of 16384 cells (16 warps per CTA), 64 of them, 100
    generations of the binary cave rule, bit-exact against the oracle."""
    rng = np.random.default_rng(7)
    arr = synth(rng, (16384, 64), 0.45, 1)
    want = oracle.ca2d_run(arr.copy(), CAVE_BIN["born"], CAVE_BIN["surv"], 1, 1, oracle_lib.NEIGH_M1, 100, side=16384)
    gpu.ca2d_step(_ca(gpu, 0x1E0, 0x1F0, 1, True, oracle_lib.NEIGH_M1), arr, side=16384, steps=100, engine=BITPLANE)
    assert np.array_equal(arr, want)
    assert arr.any() and not arr.all()

# ---- the diagonal 2D engine (ca2d_skew.cuh): binary grids stored along 2x + y, no in-row chain ------------------

DIAG_RULES = [
    # born, surv, nr, decay, neigh
    (0x1E0, 0x1F0, 1, 1, oracle_lib.NEIGH_M1),          # BASELINE config 3: cave smoothing (compile-time rule, folded count)
    (0x00C, 0x180, 1, 1, oracle_lib.NEIGH_M1),          # ca_test's masks (compile-time rule, two tables per case)
    (0x008, 0x00C, 1, 1, oracle_lib.NEIGH_M1),          # Life (run-time tables), not monotone
    (0x006, 0x01C, 1, 1, oracle_lib.NEIGH_VN1),         # von Neumann
    (0x01E, 0x0FF, 1, 0, oracle_lib.NEIGH_MV),          # mv without decay reduces to the alive-bit count
    (0x00C, 0x180, 256, 1, oracle_lib.NEIGH_M1),        # nr_states wraps to 0: nothing is ever born
]


@pytest.mark.parametrize("wpl", ["1", "2"])
@pytest.mark.parametrize("shape", [(1, 1), (1, 9), (9, 1), (31, 33), (100, 37), (257, 130), (1025, 40), (40, 1025),
                                   (2100, 24), (24, 2100), (4097, 9), (9, 4097), (3000, 2500)])
def test_ca2d_diagonal_vs_oracle_shapes(gpu, oracle, monkeypatch, shape, wpl):
    """Every rule family on ragged shapes: diagonals shorter than a word, crossing lane and warp spans, warps that
    join and leave with the band of valid cells, both words-per-lane builds."""
    monkeypatch.setenv("CLAPCA_2D_SKEW_WPL", wpl)
    rng = np.random.default_rng(shape[0] * 977 + shape[1])
    for born, surv, nr, decay, neigh in DIAG_RULES:
        arr = synth(rng, shape, 0.5, 1)
        want = oracle.ca2d_run(arr.copy(), born, surv, nr, decay, neigh, 5, side=max(shape))
        gpu.ca2d_step(_ca(gpu, born, surv, nr, decay, neigh), arr, side=max(shape), steps=5, engine=DIAGONAL)
        assert np.array_equal(arr, want), (shape, born, surv, nr, neigh)


def test_ca2d_diagonal_many_generations_more_than_resident_ctas(gpu, oracle):
    rng = np.random.default_rng(5)
    arr = synth(rng, (96, 80), 0.5, 1)
    want = oracle.ca2d_run(arr.copy(), 0x8, 0xC, 1, 1, oracle_lib.NEIGH_M1, 700, side=96)      # Life: B3/S23
    gpu.ca2d_step(_ca(gpu, 0x8, 0xC, 1, 1, oracle_lib.NEIGH_M1), arr, side=96, steps=700, engine=DIAGONAL)
    assert np.array_equal(arr, want)
    assert arr.any()


def test_ca2d_diagonal_rejects_what_it_cannot_run(gpu):
    ca = _ca(gpu, 0x1E0, 0x1F0, 1, True, oracle_lib.NEIGH_M1)
    with pytest.raises(gpu.ClapcaError):          # partial sweep
        gpu.ca2d_step(ca, np.ones((64, 64), np.uint8), side=40, steps=1, engine=DIAGONAL)
    with pytest.raises(gpu.ClapcaError):          # more than one state plane
        gpu.ca2d_step(gpu.CA_TEST, np.full((64, 64), 3, np.uint8), steps=1, engine=DIAGONAL)


@pytest.mark.parametrize("wpl", ["1", "2"])
def test_ca2d_diagonal_cfg3_full_width_strips_vs_oracle(gpu, oracle, monkeypatch, wpl):
    """BASELINE config 3's geometry one dimension at a time, 100 generations of the binary cave rule against the
    oracle: 16384 columns (every warp of the CTA, the whole mailbox chain) x 64 rows, and 64 columns x 16384 rows
    (one warp, 16 k steady steps)."""
    monkeypatch.setenv("CLAPCA_2D_SKEW_WPL", wpl)
    for k, shape in enumerate([(16384, 64), (64, 16384)]):
        rng = np.random.default_rng(70 + k)
        arr = synth(rng, shape, 0.45, 1)
        want = oracle.ca2d_run(arr.copy(), 0x1E0, 0x1F0, 1, 1, oracle_lib.NEIGH_M1, 100, side=16384)
        gpu.ca2d_step(_ca(gpu, 0x1E0, 0x1F0, 1, True, oracle_lib.NEIGH_M1), arr, side=16384, steps=100, engine=DIAGONAL)
        assert np.array_equal(arr, want), shape
        assert arr.any() and not arr.all()


def test_ca2d_engines_agree_4096(gpu):
    """row engine == diagonal engine == cell wavefront on a 4096^2 binary grid, 40 generations (three different orders
    of the same in-place sweep)"""
    rng = np.random.default_rng(61)
    arr = synth(rng, (4096, 4096), 0.47, 1)
    ca = _ca(gpu, 0x1E0, 0x1F0, 1, True, oracle_lib.NEIGH_M1)
    outs = []
    for engine, env in ((DIAGONAL, None), (BITPLANE, "0"), (WAVEFRONT, None)):
        a = arr.copy()
        if env is not None:
            os.environ["CLAPCA_2D_SKEW"] = env
        try:
            gpu.ca2d_step(ca, a, steps=40, engine=engine)
        finally:
            os.environ.pop("CLAPCA_2D_SKEW", None)
        outs.append(a)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert outs[0].any() and not outs[0].all()

def test_ca2d_auto_measures_both_engines_and_stays_exact(gpu, oracle, monkeypatch):
    """AUTO on a large binary grid measures the row and the diagonal engine on the first run of a shape class
    (clapca_api.cu:run2d_tune) and takes the faster one afterwards: every run, measuring or not, equals the oracle."""
    monkeypatch.delenv("CLAPCA_2D_SKEW", raising=False)
    monkeypatch.delenv("CLAPCA_2D_TUNE", raising=False)
    rng = np.random.default_rng(62)
    arr = synth(rng, (2048, 2304), 0.47, 1)         # 4.7 M cells: above the measuring threshold
    ca = _ca(gpu, 0x1E0, 0x1F0, 1, True, oracle_lib.NEIGH_M1)
    want = oracle.ca2d_run(arr.copy(), 0x1E0, 0x1F0, 1, 1, oracle_lib.NEIGH_M1, 20, side=2304)
    for _ in range(3):
        a = arr.copy()
        gpu.ca2d_step(ca, a, side=2304, steps=20)
        assert np.array_equal(a, want)



@pytest.mark.parametrize("wpl", ["1", "2"])
def test_ca2d_diagonal_cfg3_full_size_reference_fingerprint(gpu, oracle, monkeypatch, wpl):
    """BASELINE config 3 at FULL size (16384^2 x 100) on the diagonal engine against the unmodified reference's
    fingerprints (golden/cfg3_16384.json), both words-per-lane builds; and the run stats name the engine."""
    import json
    from clap_b200.ca import Rand48
    monkeypatch.setenv("CLAPCA_2D_SKEW", "1")
    monkeypatch.setenv("CLAPCA_2D_SKEW_WPL", wpl)
    with open(os.path.join(G, "cfg3_16384.json")) as f:
        c = json.load(f)["cave_bin_16384_x100_seed1"]
    r = c["rule"]
    ca = gpu.CellAutomaton("cave", born_mask=r["born"], surv_mask=r["surv"], nr_states=r["nr"], decay=bool(r["decay"]),
                           neigh=r["neigh"])
    got = gpu.ca2d_generate(ca, c["side"], c["steps"], Rand48(c["srand48"]))
    want = c["final_grid"]
    assert int(np.count_nonzero(got)) == want["population"]
    for row, h in want["rows"].items():
        assert "%016x" % oracle.fnv(got[int(row)]) == h, row
    assert "%016x" % oracle.fnv(got) == want["fnv1a64"]
    grid = gpu.Grid(c["side"], c["side"], 1)
    grid.upload(got)
    grid.run2d(ca, 2)
    assert grid.stats()["engine"] == "diagonal"



def test_ca2d_cfg3_16384_generations_compose(gpu):
    """BASELINE config 3 at full size (16384^2, 100 generations), too large for the oracle in seconds: the fused
    100-generation run must equal 60 then 40 generations (different pipeline depths over the same data), and the
    on-device population count must match the downloaded grid."""
    rng = np.random.default_rng(8)
    side = 16384
    arr = (rng.random((side, side), dtype=np.float32) < 0.45).astype(np.uint8)
    ca = _ca(gpu, 0x1E0, 0x1F0, 1, True, oracle_lib.NEIGH_M1)
    grid = gpu.Grid(side, side, 1)
    grid.upload(arr)
    grid.run2d(ca, 100)
    st = grid.stats()
    assert st["engine"] in ("bitplane", "diagonal") and st["planes"] == 1
    a = np.empty_like(arr)
    grid.download(a)
    grid.upload(arr)
    grid.run2d(ca, 60)
    grid.run2d(ca, 40)
    b = np.empty_like(arr)
    grid.download(b)
    assert np.array_equal(a, b)
    assert grid.count() == int(np.count_nonzero(a)) and 0 < grid.count() < side * side


# ---- noise ----------------------------------------------------------------------------------------

def test_noise_bake_golden_exact(gpu, oracle):
    g = np.load(os.path.join(G, "noise.npz"))
    a = gpu.noise_grad3d_bake_rgba8(16, 4, 2.0, 0.5, 5.0, 0xC14D)
    assert np.array_equal(a, g["bake_16_p5"])
    a = gpu.noise_grad3d_bake_rgba8(24, 3, 2.3, 0.45, 37.0, 99)
    assert np.array_equal(a, g["bake_24_p37_o3"])
    d = gpu.noise_grad3d_bake_rgba8(64)          # engine defaults (noise.c:309-317)
    assert oracle.fnv(d) == int(g["bake_64_default_hash"])
    assert d.reshape(-1)[:4].tolist() == [122, 5, 93, 0]


@pytest.mark.parametrize("case", [
    # size octaves lacunarity gain period seed
    (16, 4, 2.0, 0.5, 5.0, 0xC14D),        # dyadic step: the lattice path (fBm once per lattice point)
    (24, 3, 2.3, 0.45, 37.0, 99),          # 37 / 24 is not a float-exact lattice: the six-samples-per-voxel kernel
    (48, 4, 2.0, 0.5, 7.0, 5),             # 7 / 48: not exact either
    (40, 2, 2.0, 0.5, 10.0, 1),            # 10 / 40 = 0.25
    (33, 4, 2.0, 0.5, 33.0, 7),            # step 1: every sample sits on a lattice corner of octave 0
])
def test_noise_bake_lattice_path_equals_direct_path_and_oracle(gpu, oracle, monkeypatch, case):
    """The bake evaluates the fBm once per lattice point when the +- eps samples are bit for bit the neighbours'
    centres (checked on the host in the reference's float arithmetic) and six times per voxel otherwise; both must
    give the reference's bytes."""
    size, octv, lac, gain, period, seed = case
    a = gpu.noise_grad3d_bake_rgba8(size, octv, lac, gain, period, seed)
    monkeypatch.setenv("CLAPCA_NOISE_DIRECT", "1")
    b = gpu.noise_grad3d_bake_rgba8(size, octv, lac, gain, period, seed)
    assert np.array_equal(a, b)
    assert np.array_equal(a, oracle.noise_bake(size, octv, lac, gain, period, seed))


def test_noise_fbm_float_field(gpu):
    """north_star tolerance: 1e-5 relative; the kernel mirrors the reference's double promotions and is
    expected (and checked) to be exact to the last bit."""
    g = np.load(os.path.join(G, "noise.npz"))
    got = gpu.noise_fbm3(g["fbm_points"], 4, 2.0, 0.5, 37, 0xC14D)
    want = g["fbm_values"]
    assert np.allclose(got, want, rtol=1e-5, atol=0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_noise_bake_256_against_oracle_slices(gpu, oracle):
    """SURVEY 8(d) N1 case (size 256, period 37): exact-match on the z slices the oracle computes in seconds."""
    got = gpu.noise_grad3d_bake_rgba8(256, 4, 2.0, 0.5, 37.0, 0xC14D)
    for z in (0, 101, 255):
        want = oracle.noise_bake(256, 4, 2.0, 0.5, 37.0, 0xC14D, z0=z, z1=z + 1)
        assert np.array_equal(got[z], want[z]), z
    assert not got[..., 3].any()


def test_noise_bake_into_cuda_array_equals_linear_bake_and_golden(gpu, oracle):
    """the bake written straight into a 3D CUDA array through a surface (the object a renderer's
    TEX_3D / RGBA8 texture maps under interop) reads back identical to the linear bake and to the golden vectors;
    odd sizes exercise the array's pitched layout."""
    g = np.load(os.path.join(G, "noise.npz"))
    for size, args, key in ((16, (4, 2.0, 0.5, 5.0, 0xC14D), "bake_16_p5"), (24, (3, 2.3, 0.45, 37.0, 99), "bake_24_p37_o3")):
        tex = gpu.NoiseTexture3D(size, *args)
        assert tex.array
        got = tex.download()
        tex.close()
        assert np.array_equal(got, g[key]), key
    for size in (1, 7, 33, 64):
        tex = gpu.NoiseTexture3D(size, 4, 2.0, 0.5, 9.0, 7)
        got = tex.download()
        tex.close()
        assert np.array_equal(got, gpu.noise_grad3d_bake_rgba8(size, 4, 2.0, 0.5, 9.0, 7)), size
    tex = gpu.NoiseTexture3D(64)                # engine defaults (noise.c:309-317)
    assert oracle.fnv(tex.download()) == int(g["bake_64_default_hash"])
    tex.close()


@pytest.mark.parametrize("seed", [1, 20260101])
def test_blue_noise2d_pixels_vs_port(gpu, oracle, seed):
    """blue_noise2d_tex (core/noise.c:96-169): the RGBA32F pixels, against the port (double-precision DFT; the
    reference's kissfft is not vendored, so float rounding of the transform is the agreed tolerance: 2e-5 of the
    [0, 1] range), and the drand48 stream is left exactly where the reference leaves it."""
    from clap_b200.ca import Rand48
    rng = Rand48(seed)
    st0 = rng.x
    got = gpu.blue_noise2d(rng)
    want, st1 = oracle.blue_noise2d(st0)
    assert rng.x == st1
    assert got.shape == (64, 64, 4) and np.all(got[:, :, 3] == 1.0)
    assert np.abs(got - want).max() < 2e-5
    assert got[:, :, :3].min() == 0.0 and got[:, :, :3].max() == 1.0
    with pytest.raises(gpu.ClapcaError):
        gpu.blue_noise2d(Rand48(1), size=32)        # the reference's arrays are 64 x 64 whatever size it is given


def test_noise_lattice_wrap_far_outside_the_period(gpu, oracle):
    """fk_wrap() takes a conditional add / subtract within one period of [0, period) and the reference's double
    modulo beyond: points many periods away (both signs) must still match the oracle bit for bit"""
    rng = np.random.default_rng(9)
    pts = np.concatenate([rng.uniform(-1e4, 1e4, (4000, 3)), rng.uniform(-40, 80, (4000, 3)),
                          np.array([[-37.0, 37.0, 74.0], [-0.5, 36.999, 37.0], [-74.0001, 73.9999, 0.0]])]).astype(np.float32)
    got = gpu.noise_fbm3(pts, 4, 2.0, 0.5, 37, 0xC14D)
    want = oracle.fbm3(pts, 4, 2.0, 0.5, 37, 0xC14D)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


# ---- terrain --------------------------------------------------------------------------------------

def _field_close(got, want):
    """|delta| <= 1e-5 * max(|ref|, field RMS) (cosf/powf differ from glibc by <= 2 ulp)."""
    rms = float(np.sqrt(np.mean(want.astype(np.float64) ** 2)))
    tol = 1e-5 * np.maximum(np.abs(want), rms)
    bad = np.abs(got.astype(np.float64) - want) > tol
    assert not bad.any(), (int(bad.sum()), float(np.abs(got - want).max()))


def test_terrain_map0_exact(gpu):
    g = np.load(os.path.join(G, "terrain.npz"))
    for seed in (12345, -99, (1 << 40) + 17):
        m = gpu.terrain_map0(seed, 64)
        assert np.array_equal(m.view(np.uint32), g[f"map0_64_seed{seed}"].view(np.uint32))


def test_terrain_heightmap_golden(gpu):
    g = np.load(os.path.join(G, "terrain.npz"))
    _field_close(gpu.terrain_heightmap(12345, 128, 0.0, None, 1.0, 4), g["field_128"])
    _field_close(gpu.terrain_heightmap(12345, 128, 3.0, None, 2.5, 3), g["field_128_y3_amp2_o3"])
    _field_close(gpu.terrain_heightmap(12345, 128, 0.0, g["maze_16"]), g["heightmap_128"])


def test_terrain_smooth_table_is_bit_identical_to_direct_evaluation(gpu, monkeypatch):
    """The get_avg_height() table (terrain_smooth_kernel) and the tabulated cosine blend factors
    (terrain_heightmap_tab_kernel) must not change a single bit of the map: same float operations in the same order
    as evaluating the 3x3 kernel and cosf() at every use (terrain.c:35-71, interp.h:35-42)."""
    g = np.load(os.path.join(G, "terrain.npz"))
    for nr_v, maze, octv in ((128, g["maze_16"], 4), (131, None, 4), (1024, None, 4), (200, None, 1), (77, None, 3),
                             (96, None, 6)):        # 6 octaves: 32 distinct fractions, no factor table
        a = gpu.terrain_heightmap(12345, nr_v, 0.5, maze, 1.25, octv)
        monkeypatch.setenv("CLAPCA_TERRAIN_DIRECT", "1")
        b = gpu.terrain_heightmap(12345, nr_v, 0.5, maze, 1.25, octv)
        monkeypatch.delenv("CLAPCA_TERRAIN_DIRECT")
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (nr_v, octv)


def test_terrain_survey_spot_values_1024(gpu):
    g = np.load(os.path.join(G, "terrain.npz"))
    m = gpu.terrain_map0(12345, 1024)
    assert m.reshape(-1)[7] == g["survey_map0_7"]
    f = gpu.terrain_heightmap(12345, 1024, 0.0, None, 1.0, 4)
    assert abs(float(f.reshape(-1)[12345]) - float(g["survey_map_12345"])) < 1e-5
    assert abs(float(f.astype(np.float64).sum()) - float(g["survey_map_sum"])) < 0.05


def test_terrain_cfg5_rows_vs_oracle(gpu, oracle):
    """BASELINE config 5 shape (8192^2 with the 1024^2 ca_test maze): full GPU map, oracle on sampled rows."""
    from clap_b200.ca import Rand48
    nr_v = 8192
    maze = gpu.ca2d_generate(gpu.CA_TEST, nr_v // 8, 4, Rand48(7))
    want_maze = oracle.ca2d_run(oracle.ca2d_seed(nr_v // 8, 4, 7), 3 << 2, 3 << 7, 4, 1, oracle_lib.NEIGH_M1, 4)
    assert np.array_equal(maze, want_maze)
    got = gpu.terrain_heightmap(12345, nr_v, 0.0, maze)
    map0 = oracle.terrain_map0(12345, nr_v)
    assert np.array_equal(gpu.terrain_map0(12345, nr_v).view(np.uint32), map0.view(np.uint32))
    # 256 rows spread over the map: both edges, maze-cell boundaries (multiples of 8) and everything in between
    rows = sorted(set([0, 1, 2, 3, 8188, 8189, 8190, 8191] + list(range(5, nr_v, 33))))
    assert len(rows) >= 256
    for i0 in rows:
        want = oracle.terrain_heightmap(map0, 0.0, maze, i0=i0, i1=i0 + 1)
        _field_close(got[i0:i0 + 1], want[i0:i0 + 1])


# ---- terrain mesh stage: core/terrain.c:93-110 (calc_normal), :479-516 (vertex / normal / uv / index buffers) ----

MESH_RTOL = 1e-5        # north_star tolerance for float fields; the kernels mirror the float ops, so 0 ulp is expected


def _mesh_check(got, want, what):
    for name, a, b in zip(("vx", "norm", "tx"), got[:3], want[:3]):
        assert a.shape == b.shape
        # nr_v == 1 divides by nr_v - 1 == 0 in the reference too: NaN must meet NaN
        assert np.allclose(a, b, rtol=MESH_RTOL, atol=1e-7, equal_nan=True), (what, name, float(np.abs(a - b).max()))
        exact = float(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).mean())
        assert exact > 0.999, (what, name, exact)
    assert np.array_equal(got[3], want[3]), (what, "idx")


def test_terrain_mesh_golden_normals(gpu):
    t = np.load(os.path.join(G, "terrain.npz"))
    m = np.load(os.path.join(G, "terrain_mesh.npz"))
    _, norm, _, _ = gpu.terrain_mesh(t["heightmap_128"], 10.0, -2.0, 5.0, 300.0)
    assert np.allclose(norm, m["normals_128"], rtol=MESH_RTOL, atol=1e-7)
    _, norm, _, _ = gpu.terrain_mesh(m["rough_37"])
    assert np.allclose(norm, m["normals_rough_37"], rtol=MESH_RTOL, atol=1e-7)


@pytest.mark.parametrize("nr_v", [1, 2, 3, 31, 32, 33, 100, 300, 1000])
def test_terrain_mesh_vs_oracle(gpu, oracle, nr_v):
    """Ragged tile edges, a single vertex, and nr_v > 256 where the reference's unsigned-short indices wrap."""
    rng = np.random.default_rng(nr_v)
    hmap = (rng.random((nr_v, nr_v)) * 30 - 10).astype(np.float32)
    got = gpu.terrain_mesh(hmap, 3.5, -1.25, 100.0, 777.0)
    want = oracle.terrain_mesh(hmap, 3.5, -1.25, 100.0, 777.0)
    _mesh_check(got, want, nr_v)


def test_terrain_mesh_from_the_generated_heightmap_4096(gpu, oracle):
    """The pipeline the engine runs: cave maze -> heightmap -> mesh, at a size the oracle still does in seconds."""
    from clap_b200.ca import Rand48
    nr_v = 4096
    maze = gpu.ca2d_generate(gpu.CA_TEST, nr_v // 8, 4, Rand48(7))
    hmap = gpu.terrain_heightmap(12345, nr_v, 0.0, maze)
    got = gpu.terrain_mesh(hmap, 0.0, 0.0, 0.0, 2048.0)
    want = oracle.terrain_mesh(hmap, 0.0, 0.0, 0.0, 2048.0)
    _mesh_check(got, want, nr_v)


# ---- device-side seeding of ca2d_generate(): core/ca2d.c:86-90 with the lrand48() stream jumped per cell ----

@pytest.mark.parametrize("side,nr,seed", [(1, 4, 7), (31, 4, 1234), (32, 1, 1), (33, 7, 5), (256, 4, 1234), (300, 20, 9),
                                          (1000, 0, 3), (1024, 4, 7)])
def test_ca2d_device_seeding_equals_host_loop(gpu, oracle, side, nr, seed):
    from clap_b200.ca import Rand48
    ca = gpu.CellAutomaton("seed", born_mask=3 << 2, surv_mask=3 << 7, nr_states=nr, decay=True, neigh=oracle_lib.NEIGH_M1)
    rng = Rand48(seed)
    got = gpu.ca2d_generate(ca, side, 0, rng)
    assert np.array_equal(got, oracle.ca2d_seed(side, nr, seed)), (side, nr)
    host = Rand48(seed)
    assert np.array_equal(got, gpu.ca2d_seed(ca, side, host))
    assert rng.x == host.x                          # the stream is left where side * side draws leave it
    st = oracle.srand48(seed)
    for _ in range(min(side * side, 2000)):
        oracle.lrand48(st)
    if side * side <= 2000:
        assert rng.lrand48() == oracle.lrand48(st)


def test_ca2d_device_seeding_mid_stream_and_grid(gpu, oracle):
    """A stream that has already been drawn from, a resident grid larger than the seeded square, then generations."""
    from clap_b200.ca import Rand48
    rng, host = Rand48(99), Rand48(99)
    for _ in range(1234):
        rng.lrand48(); host.lrand48()
    grid = gpu.Grid(80, 72, 1)
    grid.seed2d(gpu.CA_TEST, rng, side=64)
    got = np.full((72, 80), 0xEE, np.uint8)
    grid.download(got)
    want = np.zeros((72, 80), np.uint8)
    want[:64, :64] = gpu.ca2d_seed(gpu.CA_TEST, 64, host)
    assert np.array_equal(got, want) and rng.x == host.x
    grid.run2d(gpu.CA_TEST, 3, side=64)
    grid.download(got)
    oracle.ca2d_run(want, 3 << 2, 3 << 7, 4, 1, oracle_lib.NEIGH_M1, 3, side=64)
    assert np.array_equal(got, want)
    grid.close()


def test_ca2d_generate_16384_seed_row_samples(gpu, oracle):
    """BASELINE config 3 size: 2^28 draws jumped on the device; the oracle walks the stream to a few columns."""
    from clap_b200.ca import Rand48
    side = 16384
    ca = gpu.CellAutomaton("cave", born_mask=0x1E0, surv_mask=0x1F0, nr_states=1, decay=True, neigh=oracle_lib.NEIGH_M1)
    rng = Rand48(1)
    got = gpu.ca2d_generate(ca, side, 0, rng)
    assert abs(float((got != 0).mean()) - 2 / 8) < 1e-3     # v in {0, 1} of 0..7 is alive
    vals = Rand48(1).lrand48_block(6 * side) % np.uint64(8)
    for x in range(6):
        col = np.where(vals[x * side:(x + 1) * side] <= 1, 1, 0).astype(np.uint8)
        assert np.array_equal(got[:, x], col), x


# ---- instantiator extraction: core/terrain.c:555-570 with terrain_height() (:336-379) ----

def _instor_check(got, want):
    assert got.shape == want.shape
    assert np.array_equal(got["kind"], want["kind"])
    if len(want) == 0:
        return
    for f in ("dx", "dy", "dz"):
        assert np.allclose(got[f], want[f], rtol=MESH_RTOL, atol=1e-6), f
        assert float((got[f].view(np.uint32) == want[f].view(np.uint32)).mean()) > 0.999, f


@pytest.mark.parametrize("nr_v,density", [(8, 1.0), (16, 0.5), (64, 0.0), (64, 1.0), (128, 0.1), (264, 0.02), (1000, 0.3)])
def test_terrain_instantiators_vs_oracle(gpu, oracle, nr_v, density):
    rng = np.random.default_rng(nr_v)
    mside = nr_v // 8
    maze = rng.integers(0, 20, (mside, mside)).astype(np.uint8)
    hit = rng.random((mside, mside)) < density
    maze[hit] = rng.integers(20, 22, int(hit.sum())).astype(np.uint8)
    hmap = (rng.random((nr_v, nr_v)) * 30 - 10).astype(np.float32)
    got = gpu.terrain_instantiators(maze, (20, 21), hmap, -12.5, 40.0, 500.0)
    want = oracle.terrain_instantiators(maze, (20, 21), hmap, -12.5, 40.0, 500.0)
    assert len(want) == int(hit.sum())
    _instor_check(got, want)


def test_terrain_instantiators_full_pipeline_2048(gpu, oracle):
    """terrain.c:434-570 end to end: cave maze, heightmap, the two ca_instors passes, then the extraction."""
    from clap_b200.ca import Rand48
    nr_v = 2048
    maze = gpu.ca2d_generate(gpu.CA_TEST, nr_v // 8, 4, Rand48(7))
    hmap = gpu.terrain_heightmap(12345, nr_v, 0.0, maze)
    for ca in gpu.CA_INSTORS:
        gpu.ca2d_step(ca, maze, nr_v // 8)
    kinds = tuple(ca.nr_states for ca in gpu.CA_INSTORS)
    got = gpu.terrain_instantiators(maze, kinds, hmap, 0.0, 0.0, 1024.0)
    want = oracle.terrain_instantiators(maze, kinds, hmap, 0.0, 0.0, 1024.0)
    assert len(want) > 0
    _instor_check(got, want)


def test_ca2d_cfg3_full_size_reference_fingerprint(gpu, oracle):
    """BASELINE config 3 at FULL size against the unmodified reference: srand48(1); ca2d_generate(cave rule, 16384,
    100) took the reference a quarter of an hour on one core (tests/golden/make_golden_cfg3.py); its fingerprints
    are committed in golden/cfg3_16384.json.  Seeding and all 100 generations run on the device here."""
    import json
    from clap_b200.ca import Rand48
    with open(os.path.join(G, "cfg3_16384.json")) as f:
        cfg = json.load(f)
    assert "cave_bin_16384_x100_seed1" in cfg
    for name, c in cfg.items():
        r = c["rule"]
        ca = gpu.CellAutomaton(name, born_mask=r["born"], surv_mask=r["surv"], nr_states=r["nr"], decay=bool(r["decay"]),
                               neigh=r["neigh"])
        for steps, key in ((0, "seed_grid"), (c["steps"], "final_grid")):
            got = gpu.ca2d_generate(ca, c["side"], steps, Rand48(c["srand48"]))
            want = c[key]
            assert int(np.count_nonzero(got)) == want["population"], (name, key)
            for row, h in want["rows"].items():
                assert "%016x" % oracle.fnv(got[int(row)]) == h, (name, key, row)
            assert "%016x" % oracle.fnv(got) == want["fnv1a64"], (name, key)


def _chain_seed(side=256, seed=2048):
    rng = np.random.default_rng(seed)
    return (rng.integers(1, 6, (side, side, side)) * (rng.random((side, side, side)) < 0.25)).astype(np.uint8)


@pytest.mark.parametrize("name", ["ca_coral", "ca_445m"])
def test_ca3d_cfg4_chain_256cube_full_50_generations_reference_fingerprint(gpu, oracle, name):
    """Link (i) of the config-4 parity chain: 256^3, the FULL 50 generations, against the unmodified
    reference's fingerprint (tests/golden/make_golden_cfg4_chain.py, golden/cfg4_chain_256.json); resident run,
    then the streamed host -> host run on the same input."""
    import json
    with open(os.path.join(G, "cfg4_chain_256.json")) as f:
        cfg = json.load(f)
    vol0 = _chain_seed()
    assert "%016x" % oracle.fnv(vol0) == cfg["seed_fnv1a64"]
    c = cfg[name]
    vol = vol0.copy()
    assert gpu.ca3d_run(vol, c["nca"], c["generations"], engine=BITPLANE) == c["population"]
    assert "%016x" % oracle.fnv(vol) == c["fnv1a64"]
    keep_in, host_in = _pinned(vol0.shape)
    keep_out, host_out = _pinned(vol0.shape)
    host_in[...] = vol0
    grid = gpu.Grid(256, 256, 256)
    assert grid.run3d_streamed(c["nca"], c["generations"], host_in, host_out, max_value=5) == c["population"]
    assert np.array_equal(host_out, vol)
    grid.close()


def test_ca3d_cfg4_full_plane_thin_slab_vs_oracle(gpu, oracle):
    """Link (ii) of the config-4 parity chain: planes of the full 2048 x 2048 size -- 2048 rows per
    sweep, one team of 16 planes -- on a slab thin enough for the 64-bit-index oracle (a few seconds of CPU)."""
    rng = np.random.default_rng(4096)
    shape = (16, 2048, 2048)
    vol = (rng.integers(1, 6, shape) * (rng.random(shape) < 0.25)).astype(np.uint8)
    want = vol.copy()
    s, b, n = oracle.ca3d_rule(7)
    wpop = oracle.ca3d_run(want, s, b, n, 3)
    assert gpu.ca3d_run(vol, 7, 3, engine=BITPLANE) == wpop
    assert np.array_equal(vol, want)
