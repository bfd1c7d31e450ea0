/*
 * emu_ca2d_skew.cpp -- runs the diagonal 2D sweep (ca2d_skew.cuh: the real kernel source, compiled with
 * -DCLAPCA_EMU) and its layout kernels on the host warp emulator and compares the final grid and population
 * with the oracle restatement of ca2d_step().  TEST ONLY: built and executed by tests/test_emu_kernels.py.
 *
 * usage: emu_ca2d_skew W H G born surv nr_states decay moore WPL rngseed [ctas [forcedyn [density]]]
 *   W, H     grid extent in x / y (index y*W + x); the whole grid is swept
 *   density  cells are alive with probability density / 8 (default 5: the reference's seeding)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "emu_runtime.h"
#include "../../clap_b200/csrc/ca2d_skew_layout.cuh"
extern "C" {
#include "../../oracle/port/oracle_port.h"
}

using namespace clapca;

static uint64_t rng_state = 0x9E3779B97F4A7C15ULL;
static uint32_t rnd()
{
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return (uint32_t)((rng_state * 0x2545F4914F6CDD1DULL) >> 32);
}

static int force_dyn = 0;

template <int WPL, bool MOORE>
static void dispatch_rule(const Sk2Params &p, int ctas, int warps)
{
    const int rule = force_dyn ? BP2_RULE_DYN : bp2_rule_for(p.born, p.surv, p.nrval);
    fprintf(stderr, "rule instantiation %d\n", rule);
    if (rule == BP2_RULE_CAVE)
        emu_launch(ctas, (warps + 1) * 32, [&]() { ca2d_skew_kernel<WPL, MOORE, Sk2RuleCave>(p); });
    else if (rule == BP2_RULE_TEST)
        emu_launch(ctas, (warps + 1) * 32, [&]() { ca2d_skew_kernel<WPL, MOORE, Sk2RuleTest>(p); });
    else
        emu_launch(ctas, (warps + 1) * 32, [&]() { ca2d_skew_kernel<WPL, MOORE, Sk2RuleDyn>(p); });
}

template <int WPL>
static void dispatch(bool moore, const Sk2Params &p, int ctas, int warps)
{
    if (moore)
        dispatch_rule<WPL, true>(p, ctas, warps);
    else
        dispatch_rule<WPL, false>(p, ctas, warps);
}

int main(int argc, char **argv)
{
    if (argc < 11) {
        fprintf(stderr, "usage: emu_ca2d_skew W H G born surv nr decay moore WPL rngseed [ctas [forcedyn [density]]]\n");
        return 2;
    }
    int W = atoi(argv[1]), H = atoi(argv[2]), G = atoi(argv[3]);
    unsigned born = (unsigned)strtoul(argv[4], 0, 0), surv = (unsigned)strtoul(argv[5], 0, 0);
    unsigned nr = (unsigned)atoi(argv[6]);
    int decay = atoi(argv[7]), moore = atoi(argv[8]), WPL = atoi(argv[9]);
    rng_state ^= (uint64_t)atoll(argv[10]) * 0x9E3779B97F4A7C15ULL;
    int ctas = argc > 11 ? atoi(argv[11]) : 3;
    force_dyn = argc > 12 ? atoi(argv[12]) : 0;
    unsigned density = argc > 13 ? (unsigned)atoi(argv[13]) : 5u;

    const unsigned nrval = nr & 0xffu;
    if (nrval > 1u) { fprintf(stderr, "one state plane: nr_states must be 0 or 1 (mod 256)\n"); return 2; }
    const int warps = sk2_warps_for(W, WPL);
    if (warps * WPL > SK2_MAX_WARPS) { fprintf(stderr, "grid too wide\n"); return 2; }
    const int RS = SK2_RS, TR = sk2_rows_alloc(W, H), T = sk2_diagonals(W, H);

    size_t n = (size_t)W * H;
    std::vector<uint8_t> cells(n), want;
    for (auto &c : cells) c = (rnd() % 8 < density) ? 1 : 0;
    want = cells;
    const int side = W > H ? W : H;
    ora_ca2d_run(want.data(), W, H, side, born, surv, nr, decay, moore ? ORA_NEIGH_M1 : ORA_NEIGH_VN1, G);
    int64_t want_pop = ora_count(want.data(), (int64_t)n);

    std::vector<uint32_t> rows((size_t)TR * RS, 0u);
    std::vector<int> prog(G > 0 ? G : 1, 0);
    unsigned ticket = 0;
    int err = 0;
    unsigned long long pop = 0;
    Sk2Layout L = { cells.data(), rows.data(), W, H, &pop };
    emu_launch(3, 64, [&]() { ca2d_skew_pack_kernel(L); });

    Sk2Params p;
    memset(&p, 0, sizeof(p));
    p.rows = rows.data();
    p.W = W; p.H = H; p.G = G; p.T = T;
    p.prog = prog.data();
    p.ticket = &ticket;
    p.err = &err;
    p.born = born & 0x1ffu;
    p.surv = decay ? (surv & 0x1ffu) : 0x1ffu;
    p.nrval = nrval;
    p.spin_limit = 20LL * 1000 * 1000 * 1000;
    if (G > 0) {
        switch (WPL) {
        case 1: dispatch<1>(moore, p, ctas, warps); break;
        case 2: dispatch<2>(moore, p, ctas, warps); break;
        default: fprintf(stderr, "bad WPL\n"); return 2;
        }
    }
    if (err) { printf("FAIL watchdog err=%d\n", err); return 1; }

    /* nothing may leak outside the band of valid cells or into the pad words */
    size_t leaks = 0;
    for (int t = 0; t < TR; t++)
        for (int wd = 0; wd < RS; wd++) {
            uint32_t v = rows[(size_t)t * RS + wd];
            for (int b = 0; v && b < 32; b++)
                if ((v >> b) & 1u) {
                    long long x = 32LL * wd + b, y = (long long)t - 2 * x;
                    if (x >= W || y < 0 || y >= H) leaks++;
                }
        }
    if (leaks) { printf("FAIL %zu bits outside the grid\n", leaks); return 1; }

    std::vector<uint8_t> got(n, 0xEE);
    Sk2Layout U = { got.data(), rows.data(), W, H, &pop };
    emu_launch(3, 64, [&]() { ca2d_skew_unpack_kernel(U); });

    size_t diff = 0, first = n;
    for (size_t i = 0; i < n; i++)
        if (got[i] != want[i]) { if (!diff) first = i; diff++; }
    if (diff || (int64_t)pop != want_pop) {
        printf("FAIL diff=%zu first=%zu (x=%zu y=%zu got=%d want=%d) pop=%llu want_pop=%lld\n", diff, first,
               first % W, first / W, first < n ? got[first] : -1, first < n ? want[first] : -1, pop, (long long)want_pop);
        return 1;
    }
    printf("OK W=%d H=%d G=%d WPL=%d warps=%d ctas=%d pop=%llu\n", W, H, G, WPL, warps, ctas, pop);
    return 0;
}
