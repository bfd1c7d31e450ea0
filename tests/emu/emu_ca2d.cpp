/*
 * emu_ca2d.cpp -- runs the 2D bit-plane kernels (the real kernel source, compiled with -DCLAPCA_EMU)
 * on the host warp emulator and compares the final grid and population with the oracle restatement of
 * ca2d_step().  TEST ONLY: built and executed by tests/test_emu_kernels.py.
 *
 * usage: emu_ca2d W H G born surv nr_states decay moore P WPL warps seedkind rngseed [ctas [flagrows [forcedyn]]]
 *   W, H     grid extent in x / y (index y*W + x); the whole grid is swept (side >= max(W,H))
 *   seedkind 0 = reference seeding density (cells are 0 or nr_states), 1 = dense random 0..2^P-1
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "emu_runtime.h"
#include "../../clap_b200/csrc/ca2d_bitplane.cuh"
#include "../../clap_b200/csrc/ca2d_layout.cuh"
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

static int force_dyn = 0;       /* argv[16]: 1 = run the run-time-mask instantiation even for a rule that has its own */

template <int P, int WPL, bool MOORE>
static void dispatch_rule(const Bp2Params &p, int ctas, int warps)
{
    const int rule = force_dyn ? BP2_RULE_DYN : bp2_rule_for(p.born, p.surv, p.nrval);
    fprintf(stderr, "rule instantiation %d\n", rule);
    if (rule == BP2_RULE_CAVE)
        emu_launch(ctas, (warps + 1) * 32, [&]() { ca2d_sweep_kernel<P, WPL, MOORE, Rule2Cave>(p); });
    else if (rule == BP2_RULE_TEST)
        emu_launch(ctas, (warps + 1) * 32, [&]() { ca2d_sweep_kernel<P, WPL, MOORE, Rule2Test>(p); });
    else
        emu_launch(ctas, (warps + 1) * 32, [&]() { ca2d_sweep_kernel<P, WPL, MOORE, Rule2Dyn>(p); });
}

template <int P, int WPL>
static void dispatch(bool moore, const Bp2Params &p, int ctas, int warps)
{
    if (moore)
        dispatch_rule<P, WPL, true>(p, ctas, warps);
    else
        dispatch_rule<P, WPL, false>(p, ctas, warps);
}

template <int P>
static void dispatch_wpl(int WPL, bool moore, const Bp2Params &p, int ctas, int warps)
{
    switch (WPL) {
    case 1: dispatch<P, 1>(moore, p, ctas, warps); break;
    case 2: dispatch<P, 2>(moore, p, ctas, warps); break;
    case 4: dispatch<P, 4>(moore, p, ctas, warps); break;
    default: fprintf(stderr, "bad WPL\n"); exit(2);
    }
}

int main(int argc, char **argv)
{
    if (argc < 14) {
        fprintf(stderr, "usage: emu_ca2d W H G born surv nr decay moore P WPL warps seedkind rngseed [ctas [flagrows]]\n");
        return 2;
    }
    int W = atoi(argv[1]), H = atoi(argv[2]), G = atoi(argv[3]);
    unsigned born = (unsigned)strtoul(argv[4], 0, 0), surv = (unsigned)strtoul(argv[5], 0, 0);
    unsigned nr = (unsigned)atoi(argv[6]);
    int decay = atoi(argv[7]), moore = atoi(argv[8]), P = atoi(argv[9]), WPL = atoi(argv[10]), warps = atoi(argv[11]);
    int seedkind = atoi(argv[12]);
    rng_state ^= (uint64_t)atoll(argv[13]) * 0x9E3779B97F4A7C15ULL;
    int ctas = argc > 14 ? atoi(argv[14]) : 3;
    int flagRows = argc > 15 ? atoi(argv[15]) : 2;
    force_dyn = argc > 16 ? atoi(argv[16]) : 0;

    const unsigned nrval = nr & 0xffu;
    const unsigned vmax = P >= 8 ? 255u : (1u << P) - 1u;
    if (nrval > vmax) { fprintf(stderr, "rule needs more planes\n"); return 2; }
    const int RW = (H + 31) / 32;
    if (RW > warps * 32 * WPL) { fprintf(stderr, "row too wide for warps*WPL\n"); return 2; }
    const int RWS = warps * 32 * WPL;

    size_t n = (size_t)W * H;
    std::vector<uint8_t> cells(n), want;
    for (auto &c : cells) {
        if (seedkind == 0) c = (rnd() % 8 <= 4) ? (uint8_t)nrval : 0;
        else c = (rnd() % 5 < 2) ? 0 : (uint8_t)(rnd() % (vmax + 1));
    }
    want = cells;
    const int side = W > H ? W : H;
    ora_ca2d_run(want.data(), W, H, side, born, surv, nr, decay, moore ? ORA_NEIGH_M1 : ORA_NEIGH_VN1, G);
    int64_t want_pop = ora_count(want.data(), (int64_t)n);

    std::vector<uint32_t> rows((size_t)W * P * RWS, 0u);
    std::vector<int> prog(G > 0 ? G : 1, 0);
    unsigned ticket = 0;
    int err = 0;
    unsigned long long pop = 0;
    Bp2Layout L = { cells.data(), rows.data(), W, H, P, RWS, &pop };
    emu_launch(2, 64, [&]() { ca2d_pack_kernel(L); });

    Bp2Params p;
    memset(&p, 0, sizeof(p));
    p.rows = rows.data();
    p.N = H; p.M = W; p.G = G; p.RWS = RWS;
    p.prog = prog.data();
    p.ticket = &ticket;
    p.err = &err;
    p.born = born & 0x1ffu;
    p.surv = decay ? (surv & 0x1ffu) : 0x1ffu;
    p.nrval = nrval;
    p.flag_rows = flagRows;
    p.spin_limit = 20LL * 1000 * 1000 * 1000;
    if (G > 0) {
        switch (P) {
        case 1: dispatch_wpl<1>(WPL, moore, p, ctas, warps); break;
        case 3: dispatch_wpl<3>(WPL, moore, p, ctas, warps); break;
        case 4: dispatch_wpl<4>(WPL, moore, p, ctas, warps); break;
        case 8: dispatch_wpl<8>(WPL, moore, p, ctas, warps); break;
        default: fprintf(stderr, "bad P\n"); return 2;
        }
    }
    if (err) { printf("FAIL watchdog err=%d\n", err); return 1; }

    std::vector<uint8_t> got(n, 0xEE);
    Bp2Layout U = { got.data(), rows.data(), W, H, P, RWS, &pop };
    emu_launch(2, 64, [&]() { ca2d_unpack_kernel(U); });

    size_t diff = 0, first = n;
    for (size_t i = 0; i < n; i++)
        if (got[i] != want[i]) { if (!diff) first = i; diff++; }
    if (diff || (int64_t)pop != want_pop) {
        printf("FAIL diff=%zu first=%zu (x=%zu y=%zu got=%d want=%d) pop=%llu want_pop=%lld\n", diff, first,
               first % W, first / W, first < n ? got[first] : -1, first < n ? want[first] : -1, pop, (long long)want_pop);
        return 1;
    }
    printf("OK W=%d H=%d G=%d P=%d WPL=%d warps=%d ctas=%d pop=%llu\n", W, H, G, P, WPL, warps, ctas, pop);
    return 0;
}
