/*
 * emu_ca3d.cpp -- runs the bit-plane ca3d kernels (the real kernel source,
 * compiled with -DCLAPCA_EMU) on the host warp emulator and compares the final
 * volume and the population with the oracle restatement of ca3d_run().
 * TEST ONLY: built and executed by tests/test_emu_kernels.py.
 *
 * usage: emu_ca3d W H Z G nca P WPL seedkind rngseed [warps [ranks [block [seg [flagrows [genbatch [pubworkers [team [tilegens [layout [chunk [gskew]]]]]]]]]]]]
 *   team     > 0: tile mode, compute warps per CTA (every CTA gets one more warp, the service warp); `warps` / team
 *            CTAs are launched per rank.  tilegens = wanted generations per tile (the planner lowers it when the
 *            launch has too few CTAs for the forward dependency).  Several ranks need tile mode: the service warp
 *            carries the halo rows.
 *   layout   1 = layout items (pack / unpack run as work items of the sweep launch, cells resident),
 *            2 = layout items fed by a "copy engine" thread that delivers the cells chunk by chunk and raises
 *                in_ready, while a second thread drains finished chunks as their planes' out_done words show the epoch
 *            (single rank, time-key or team order)
 *   nca      0..8 = compile-time rule of cas[], 9 = run-time rule (coral masks),
 *            10   = run-time rule with random masks, 11 / 12 = run-time "chain" rules (rows are one / a few
 *                   dependency chains: the long-run paths of the in-row scan)
 *   seedkind 0 = sparse values 0..5, 1 = dense 0..min(2^P-1,255), 2 = ca3d_make seed (has 255s), 3 = binary (0 / 1)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <thread>
#include <chrono>
#include <algorithm>
#include "emu_runtime.h"
#include "../../clap_b200/csrc/ca3d_bitplane.cuh"
#include "../../clap_b200/csrc/ca3d_layout.cuh"
#include "../../clap_b200/csrc/bp_plan.h"
extern "C" {
#include "../../oracle/port/oracle_port.h"
}

using namespace clapca;

#define B(n) (1u << (n))
#define RANGE(s, e) (((1u << ((e) - (s))) - 1u) << (s))

typedef Rule3Const<B(4), B(4), 5> R0;
typedef Rule3Const<B(6) | B(7) | B(8), B(6) | B(7) | B(8), 3> R1;
typedef Rule3Const<B(4) | B(5) | B(6) | B(7), B(6) | B(7) | B(8), 10> R2;
typedef Rule3Const<RANGE(9, 26), B(5) | B(6) | B(7) | B(12) | B(13) | B(15), 5> R3;
typedef Rule3Const<B(2) | B(6) | B(9), B(4) | B(6) | B(8) | B(9), 10> R4;
typedef Rule3Const<B(1) | B(4) | B(8) | B(11) | RANGE(13, 26), RANGE(13, 26), 5> R5;
typedef Rule3Const<RANGE(0, 3) | RANGE(7, 9) | RANGE(11, 13) | B(18) | B(21) | B(22) | B(24) | B(26),
                   B(4) | B(13) | B(17) | RANGE(20, 24) | B(26), 4> R6;
typedef Rule3Const<RANGE(5, 8), RANGE(6, 7) | B(9) | B(12), 4> R7;
typedef Rule3Const<RANGE(0, 6), B(1) | B(3), 2> R8;

template <int P, int WPL, class Rule>
static void launch_sweep(const Bp3Params &p, int warps)
{
    if (p.team > 0) {
        /* tile mode: CTAs of `team` compute warps + the service warp, each CTA sweeps a tile of planes x generations */
        const int blocks = (warps + p.team - 1) / p.team;
        emu_launch(blocks, (p.team + 1) * 32, [&]() { Sweep3<P, WPL, Rule>::tile_loop(p); });
        return;
    }
    if (p.pub_workers > 0) {
        /* publisher mode: CTAs of pub_workers worker warps + 1 publisher warp */
        const int blocks = (warps + p.pub_workers - 1) / p.pub_workers;
        emu_launch(blocks, (p.pub_workers + 1) * 32, [&]() { Sweep3<P, WPL, Rule>::kernel_body(p); });
        return;
    }
    emu_launch(1, warps * 32, [&]() { Sweep3<P, WPL, Rule>::kernel_body(p); });
}

template <int P, int WPL>
static void dispatch_rule(int nca, const Bp3Params &p, int warps)
{
    switch (nca) {
    case 0: launch_sweep<P, WPL, R0>(p, warps); break;
    case 1: launch_sweep<P, WPL, R1>(p, warps); break;
    case 2: launch_sweep<P, WPL, R2>(p, warps); break;
    case 3: launch_sweep<P, WPL, R3>(p, warps); break;
    case 4: launch_sweep<P, WPL, R4>(p, warps); break;
    case 5: launch_sweep<P, WPL, R5>(p, warps); break;
    case 6: launch_sweep<P, WPL, R6>(p, warps); break;
    case 7: launch_sweep<P, WPL, R7>(p, warps); break;
    case 8: launch_sweep<P, WPL, R8>(p, warps); break;
    default: launch_sweep<P, WPL, Rule3Dyn>(p, warps); break;
    }
}

template <int P>
static void dispatch_wpl(int WPL, int nca, const Bp3Params &p, int warps)
{
    switch (WPL) {
    case 1: dispatch_rule<P, 1>(nca, p, warps); break;
    case 2: dispatch_rule<P, 2>(nca, p, warps); break;
    case 4: dispatch_rule<P, 4>(nca, p, warps); break;
    default: fprintf(stderr, "bad WPL\n"); exit(2);
    }
}

static uint64_t rng_state = 88172645463325252ULL;
static uint32_t rnd()
{
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 11);
}

int main(int argc, char **argv)
{
    if (argc < 10) {
        fprintf(stderr, "usage: emu_ca3d W H Z G nca P WPL seedkind rngseed [warps [ranks [block [seg [flagrows]]]]]\n");
        return 2;
    }
    int W = atoi(argv[1]), H = atoi(argv[2]), Z = atoi(argv[3]), G = atoi(argv[4]);
    int nca = atoi(argv[5]), P = atoi(argv[6]), WPL = atoi(argv[7]), seedkind = atoi(argv[8]);
    rng_state ^= (uint64_t)atoll(argv[9]) * 0x9E3779B97F4A7C15ULL;
    int warps = argc > 10 ? atoi(argv[10]) : 6;
    int ranks = argc > 11 ? atoi(argv[11]) : 1;      /* emulated GPUs (z-block slab decomposition) */
    int blockB = argc > 12 ? atoi(argv[12]) : 0;     /* planes per z-block, 0 = contiguous slabs */
    int segL = argc > 13 ? atoi(argv[13]) : 0;       /* rows per work item, 0 = planner default */
    int flagRows = argc > 14 ? atoi(argv[14]) : 2;   /* rows per progress-counter update */
    int genBatch = argc > 15 ? atoi(argv[15]) : 0;   /* > 0: generation-batched diagonal order, -1: time-key order */
    int pubWorkers = argc > 16 ? atoi(argv[16]) : 0; /* > 0: publisher mode, worker warps per CTA */
    int team = argc > 17 ? atoi(argv[17]) : 0;       /* > 0: team mode, warps (= planes of a group) per CTA */
    int tileGens = argc > 18 ? atoi(argv[18]) : 1;   /* tile mode: wanted generations per tile */
    int layout = argc > 19 ? atoi(argv[19]) : 0;
    int chunk = argc > 20 ? atoi(argv[20]) : 2;
    int gskew = argc > 21 ? atoi(argv[21]) : 0;      /* tile mode: key distance between generation groups (0 = Tz + 1) */
    if (ranks > 1 && team <= 0) {
        fprintf(stderr, "several ranks need tile mode (team > 0)\n");
        return 2;
    }
    if (layout && (team <= 0 && genBatch >= 0)) {
        fprintf(stderr, "layout items: time-key (genbatch -1) or tile order\n");
        return 2;
    }
    if (layout == 2 && ranks != 1) {
        fprintf(stderr, "layout 2 (copy-engine threads): single rank\n");
        return 2;
    }

    unsigned surv, born, nr;
    if (nca <= 9) {
        ora_ca3d_rule(nca == 9 ? 7 : nca, &surv, &born, &nr);
    } else if (nca == 11 || nca == 12) {
        /* "chain" rules: the table flips between every K and K + 1, so (almost) every cell depends on its in-row
           predecessor -- rows are single dependency chains (11) or chains broken only where K is 6 or 13 (12): the
           long-run paths of the in-row scan (bitslice.cuh: bs_scan_word_e_hi, the warp stage) */
        surv = 0x5555555u;
        born = 0x2aaaaaau;
        if (nca == 12) { surv ^= 1u << 7; born ^= 1u << 14; }
        nr = 2;
    } else {
        surv = rnd() & 0x7ffffff;
        born = rnd() & rnd() & 0x7ffffff;
        nr = 1 + rnd() % ((1u << (P > 7 ? 8 : P)) - 1);
    }
    unsigned bornval = (nr - 1) & 0xff;
    if (bornval >> P) {
        fprintf(stderr, "rule needs more than %d planes\n", P);
        return 2;
    }

    size_t n = (size_t)W * H * Z;
    std::vector<uint8_t> cells(n), want(n);
    unsigned vmax = P >= 8 ? 255u : (1u << P) - 1u;
    if (seedkind == 2) {
        uint64_t st;
        ora_srand48(&st, (long)atoll(argv[9]));
        ora_ca3d_make(cells.data(), W, H, Z, &st);
        if (P < 8)
            for (auto &c : cells) if (c > vmax) c = (uint8_t)vmax;
    } else {
        for (auto &c : cells) {
            if (seedkind == 3)
                c = (uint8_t)(rnd() & 1u);
            else if (seedkind == 0)
                c = (rnd() % 4 == 0) ? (uint8_t)(1 + rnd() % (vmax < 5 ? vmax : 5)) : 0;
            else
                c = (rnd() % 5 < 2) ? 0 : (uint8_t)(rnd() % (vmax + 1));
        }
    }
    want = cells;
    int64_t want_pop = ora_ca3d_run(want.data(), W, H, Z, surv, born, nr, G);

    /* device-side flow per rank: pack -> (halo init) -> sweep -> unpack; ranks run concurrently */
    int RWP = 32 * WPL;
    if (W > 32 * RWP) {
        fprintf(stderr, "row too wide for WPL\n");
        return 2;
    }
    const int NP = P + 2;
    const int Gcap = G > 0 ? G : 1;
    struct Rank {
        SlabGeom geo;
        HaloLayout hl;
        std::vector<uint8_t> cells;         /* local planes, reference layout */
        std::vector<uint32_t> rows, halo;
        std::vector<int> prog;
        std::vector<Bp3Plane> planes;
        std::vector<int4> order;
        std::vector<int> out_done;
        int in_ready = 0;
        unsigned ticket = 0;
        unsigned long long pop = 0;
        Bp3Params p;
    };
    std::vector<Rank> rk(ranks);
    int err = 0;
    for (int r = 0; r < ranks; r++) {
        Rank &k = rk[r];
        k.geo = SlabGeom{ Z, ranks, r, blockB > 0 ? blockB : (Z + ranks - 1) / ranks };
        k.hl = slab_halo_layout(k.geo, H, RWP, NP, Gcap);
        const int Zl = k.geo.local_planes();
        k.cells.resize((size_t)W * H * (Zl ? Zl : 1));
        for (int lb = 0; lb < k.geo.local_blocks(); lb++) {
            int jb = k.geo.global_block(lb);
            memcpy(k.cells.data() + (size_t)k.geo.local_z0(lb) * W * H,
                   cells.data() + (size_t)k.geo.block_z0(jb) * W * H, (size_t)k.geo.block_len(jb) * W * H);
        }
        k.rows.assign((size_t)(Zl ? Zl : 1) * H * NP * RWP, 0xdeadbeefu);
        k.halo.assign(k.hl.total_words, 0u);
        k.prog.assign((size_t)(Gcap + 1) * (Zl ? Zl : 1), 0);       /* row 0 = "generation -1" of the layout items */
    }
    for (int r = 0; r < ranks; r++) {
        Rank &k = rk[r];
        const int Zl = k.geo.local_planes();
        SlabPtrs ptr = { k.rows.data(), k.prog.data() + (Zl ? Zl : 1), k.halo.data(), rk[(r + 1) % ranks].halo.data(),
                         rk[(r + ranks - 1) % ranks].halo.data() };
        bp3_build_planes(k.geo, ptr, k.hl, H, RWP, NP, k.planes);
        std::vector<WorkItem> items;
        if (team > 0) {
            const int ctas = (warps + team - 1) / team;
            int Tz = team, Tg = 1;
            if (ranks > 1) Tg = bp3_tile_shape_all_ranks(k.geo, H, G, team, tileGens, ctas, &Tz, layout != 0, gskew);
            else Tg = bp3_tile_shape(k.planes, H, G, team, tileGens, ctas, layout != 0, &Tz, gskew);
            if (r == 0) fprintf(stderr, "tile shape %d planes x %d generations\n", Tz, Tg);
            bp3_make_items_tile(k.planes, H, G, Tz, Tg, items, layout != 0, nullptr, gskew);
        }
        else if (genBatch > 0)
            bp3_make_items_batched(k.planes, Z, H, G, genBatch, items);
        else if (genBatch < 0)
            bp3_make_items_timekey(k.planes, H, G, items, layout != 0);
        else
            bp3_make_items(k.planes, Z, H, G, segL > 0 ? segL : bp3_segment_rows(Z, H, G, warps), items);
        k.order.resize(items.size());
        for (size_t i = 0; i < items.size(); i++) k.order[i] = make_int4(items[i].z, items[i].g, items[i].y0, items[i].y1);
        if (Zl && !layout) {
            Bp3Layout L = { k.cells.data(), k.rows.data(), W, H, Zl, P, RWP, &k.pop };
            if (P <= 3) emu_launch(2, 64, [&]() { ca3d_pack_rows_kernel<3>(L); });
            else if (P == 4) emu_launch(2, 64, [&]() { ca3d_pack_rows_kernel<4>(L); });
            else emu_launch(2, 64, [&]() { ca3d_pack_rows_kernel<8>(L); });
        }
        memset(&k.p, 0, sizeof(k.p));
        k.p.rows = k.rows.data();
        k.p.planes = k.planes.data();
        k.p.W = W; k.p.H = H; k.p.Z = Zl; k.p.G = G; k.p.RWP = RWP;
        k.p.prog = k.prog.data() + (Zl ? Zl : 1);
        k.p.order = k.order.data();
        k.p.nsweeps = (int)k.order.size();
        k.p.flag_rows = flagRows;
        k.p.pub_workers = pubWorkers;
        k.p.team = team;
        static const uint32_t zero_row[1024] = { 0 };
        k.p.zeros = zero_row;
        /* halo rows: the product's default (ld / st batches) unless CLAPCA_HALO_LDST=0 asks for the bulk-copy path */
        k.p.halo_ldst = getenv("CLAPCA_HALO_LDST") ? atoi(getenv("CLAPCA_HALO_LDST")) != 0 : 1;
        k.p.ticket = &k.ticket;
        k.p.err = &err;
        k.p.surv = surv; k.p.born = born; k.p.bornval = bornval;
        k.p.spin_limit = 20LL * 1000 * 1000 * 1000;      /* 20 s of emulator wall clock */
        if (layout) {
            k.out_done.assign(Zl + 1, 0);
            k.p.layout_items = 1;
            k.p.io_cells = k.cells.data();
            k.p.in_ready = layout == 2 ? &k.in_ready : nullptr;
            k.p.out_done = k.out_done.data();
            k.p.io_chunk = chunk;
            k.p.io_epoch = 3;
            k.p.population = &k.pop;
        }
    }
    /* halo init: the first plane of every block but the first seeds the ghost plane above the previous block */
    for (int r = 0; r < ranks && !layout; r++) {    /* layout items: the pack items' service warps seed the ghost planes */
        for (size_t l = 0; l < rk[r].planes.size(); l++) {
            const Bp3Plane &pl = rk[r].planes[l];
            if (!pl.push_dn_rows) continue;
            const uint32_t *src = rk[r].rows.data() + (size_t)l * H * NP * RWP;
            uint32_t *dst = pl.push_dn_rows;
            emu_launch(1, 64, [&]() { halo_seed_kernel(dst, src, H, RWP, NP); });
        }
    }
    std::vector<uint8_t> host_in, host_out;
    if (layout == 2) {
        /* the cells only arrive while the kernel runs: start from garbage */
        host_in = rk[0].cells;
        host_out.assign(n, 0xEE);
        memset(rk[0].cells.data(), 0xA5, rk[0].cells.size());
    }
    if (G > 0) {
        std::vector<std::thread> ths;
        if (layout == 2) {
            Rank &k = rk[0];
            const size_t plane = (size_t)W * H;
            const int nchunks = (Z + chunk - 1) / chunk;
            ths.emplace_back([&, plane, nchunks]() {            /* "H2D copy stream": memcpy, then the 32-bit write */
                for (int c = 0; c < nchunks; c++) {
                    std::this_thread::sleep_for(std::chrono::microseconds(300));
                    const int z0 = c * chunk, z1 = std::min(Z, z0 + chunk);
                    memcpy(k.cells.data() + z0 * plane, host_in.data() + z0 * plane, (z1 - z0) * plane);
                    __atomic_store_n(&k.in_ready, c + 1, __ATOMIC_SEQ_CST);
                }
            });
            ths.emplace_back([&, plane, nchunks]() {            /* "D2H copy stream": 32-bit wait, then memcpy */
                for (int c = 0; c < nchunks; c++) {
                    const int z0 = c * chunk, z1 = std::min(Z, z0 + chunk);
                    for (int z = z0; z < z1; z++)
                        while (__atomic_load_n(&k.out_done[z], __ATOMIC_SEQ_CST) != 3 && !__atomic_load_n(&err, __ATOMIC_SEQ_CST))
                            std::this_thread::yield();
                    memcpy(host_out.data() + z0 * plane, k.cells.data() + z0 * plane, (z1 - z0) * plane);
                }
            });
        }
        for (int r = 0; r < ranks; r++)
            ths.emplace_back([&, r]() {
                const Bp3Params &p = rk[r].p;
                if (!p.nsweeps) return;
                switch (P) {
                case 3: dispatch_wpl<3>(WPL, nca, p, warps); break;
                case 4: dispatch_wpl<4>(WPL, nca, p, warps); break;
                case 8: dispatch_wpl<8>(WPL, nca, p, warps); break;
                default: fprintf(stderr, "bad P\n"); exit(2);
                }
            });
        for (auto &t : ths) t.join();
    }
    if (err) {
        printf("FAIL watchdog err=%d\n", err);
        return 1;
    }
    std::vector<uint8_t> got(n, 0xEE);
    unsigned long long pop = 0;
    if (layout) {
        for (int r = 0; r < ranks; r++) {
            Rank &k = rk[r];
            const int Zl = k.geo.local_planes();
            for (int z = 0; z < Zl; z++)
                if (k.out_done[z] != 3) {
                    printf("FAIL rank %d out_done[%d]=%d\n", r, z, k.out_done[z]);
                    return 1;
                }
            pop += k.pop;
            if (layout == 2) {
                got = host_out;
                continue;
            }
            for (int lb = 0; lb < k.geo.local_blocks(); lb++) {
                int jb = k.geo.global_block(lb);
                memcpy(got.data() + (size_t)k.geo.block_z0(jb) * W * H, k.cells.data() + (size_t)k.geo.local_z0(lb) * W * H,
                       (size_t)k.geo.block_len(jb) * W * H);
            }
        }
    }
    for (int r = 0; r < ranks && !layout; r++) {
        Rank &k = rk[r];
        const int Zl = k.geo.local_planes();
        if (!Zl) continue;
        std::vector<uint8_t> out(k.cells.size(), 0xEE);
        k.pop = 0;
        Bp3Layout L = { out.data(), k.rows.data(), W, H, Zl, P, RWP, &k.pop };
        if (P <= 3) emu_launch(2, 64, [&]() { ca3d_unpack_rows_kernel<3>(L); });
        else if (P == 4) emu_launch(2, 64, [&]() { ca3d_unpack_rows_kernel<4>(L); });
        else emu_launch(2, 64, [&]() { ca3d_unpack_rows_kernel<8>(L); });
        pop += k.pop;
        for (int lb = 0; lb < k.geo.local_blocks(); lb++) {
            int jb = k.geo.global_block(lb);
            memcpy(got.data() + (size_t)k.geo.block_z0(jb) * W * H, out.data() + (size_t)k.geo.local_z0(lb) * W * H,
                   (size_t)k.geo.block_len(jb) * W * H);
        }
    }

    size_t diff = 0, first = n;
    for (size_t i = 0; i < n; i++)
        if (got[i] != want[i]) { if (!diff) first = i; diff++; }
    if (diff || (int64_t)pop != want_pop) {
        printf("FAIL diff=%zu first=%zu (x=%zu y=%zu z=%zu got=%d want=%d) pop=%llu want_pop=%lld\n", diff, first,
               first % W, (first / W) % H, first / ((size_t)W * H), first < n ? got[first] : -1,
               first < n ? want[first] : -1, pop, (long long)want_pop);
        return 1;
    }
    printf("OK W=%d H=%d Z=%d G=%d nca=%d P=%d WPL=%d ranks=%d B=%d pop=%llu\n", W, H, Z, G, nca, P, WPL, ranks, blockB, pop);
    return 0;
}
