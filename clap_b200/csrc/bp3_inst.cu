/*
 * bp3_inst.cu -- instantiates ca3d_sweep_kernel for ONE rule (-DBP3_RULE=0..9)
 * and all (P, WPL) variants, and provides its cooperative launcher.
 */
#include <stdlib.h>
#include <algorithm>
#include "bp3_launch.h"

#ifndef BP3_RULE
#error "compile with -DBP3_RULE=<0..9>"
#endif

namespace clapca {

#define B(n) (1u << (n))
#define RANGE(s, e) (((1u << ((e) - (s))) - 1u) << (s))     /* core/ca3d.h:34 */

/* cas[]: core/ca3d.c:110-122 -- <surv, born, nr_states> */
#if BP3_RULE == 0
typedef Rule3Const<B(4), B(4), 5> TheRule;                                          /* ca_445m */
#elif BP3_RULE == 1
typedef Rule3Const<B(6) | B(7) | B(8), B(6) | B(7) | B(8), 3> TheRule;              /* ca_678_678_3m */
#elif BP3_RULE == 2
typedef Rule3Const<B(4) | B(5) | B(6) | B(7), B(6) | B(7) | B(8), 10> TheRule;      /* ca_pyroclastic */
#elif BP3_RULE == 3
typedef Rule3Const<RANGE(9, 26), B(5) | B(6) | B(7) | B(12) | B(13) | B(15), 5> TheRule;   /* ca_amoeba */
#elif BP3_RULE == 4
typedef Rule3Const<B(2) | B(6) | B(9), B(4) | B(6) | B(8) | B(9), 10> TheRule;      /* ca_builder */
#elif BP3_RULE == 5
typedef Rule3Const<B(1) | B(4) | B(8) | B(11) | RANGE(13, 26), RANGE(13, 26), 5> TheRule;  /* ca_slow_decay */
#elif BP3_RULE == 6
typedef Rule3Const<RANGE(0, 3) | RANGE(7, 9) | RANGE(11, 13) | B(18) | B(21) | B(22) | B(24) | B(26),
                   B(4) | B(13) | B(17) | RANGE(20, 24) | B(26), 4> TheRule;        /* ca_spiky_growth */
#elif BP3_RULE == 7
typedef Rule3Const<RANGE(5, 8), RANGE(6, 7) | B(9) | B(12), 4> TheRule;             /* ca_coral */
#elif BP3_RULE == 8
typedef Rule3Const<RANGE(0, 6), B(1) | B(3), 2> TheRule;                            /* ca_crystal_1 */
#else
typedef Rule3Dyn TheRule;                                                           /* run-time masks */
#endif

template <int P, int WPL>
static cudaError_t launch_one(const Bp3Params &p, int sms, cudaStream_t stream, Bp3LaunchInfo *info)
{
    auto kern = p.team > 0 ? ca3d_team_kernel<P, WPL, TheRule> : ca3d_sweep_kernel<P, WPL, TheRule>;
    int team_cap = bp3_team_cap(P, WPL);
    if constexpr (P == 3 && WPL <= 2) {
        /* teams beyond the default kernel's launch bound: the 96-register build (bp3_team_cap_wide) */
        if (p.team > team_cap) {
            kern = ca3d_team_wide_kernel<P, WPL, TheRule>;
            team_cap = bp3_team_cap_wide(P, WPL);
        }
    }
    /* small CTAs: co-residency is bounded by registers, 128-thread granularity wastes the least of the file */
    int threads = 128, wpc = 4;             /* threads and WORKER warps per CTA */
    if (const char *e = getenv("CLAPCA_CTA_THREADS")) { int v = atoi(e); if (v == 32 || v == 64 || v == 128) threads = v; }
    wpc = threads / 32;
    if (p.pub_workers > 0 && p.team <= 0) {
        /* publisher mode: pub_workers worker warps + the publisher warp per CTA */
        wpc = std::min(std::min(p.pub_workers, (int)BP3_MAX_PUB_WORKERS), Bp3Bounds<P, WPL>::kMaxThreads / 32 - 1);
        threads = 32 * (wpc + 1);
    }
    if (p.team > 0) {
        /* tile mode: one CTA = `team` compute warps sweeping a tile + the service warp; an item occupies a whole CTA */
        wpc = std::min(p.team, team_cap);
        threads = 32 * (wpc + 1);
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    /* fewer resident warps run faster each: worth it when the dependency DAG, not the SM count, bounds the parallelism */
    if (p.max_ctas_per_sm > 0 && per_sm > p.max_ctas_per_sm) per_sm = p.max_ctas_per_sm;
    int blocks = per_sm * sms;
    /* several ranks sharing one device (tests): each launch keeps to its share of the SMs */
    if (p.max_ctas > 0 && blocks > p.max_ctas) blocks = p.max_ctas;
    if (p.nsweeps >= 0) {
        const int need = p.team > 0 ? p.nsweeps : (p.nsweeps + wpc - 1) / wpc;
        if (blocks > need) blocks = need;
    }
    if (blocks < 1) blocks = 1;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    if (info) {
        info->blocks = blocks;
        info->threads = threads;
        info->workers = blocks * wpc;
        info->regs = fa.numRegs;
    }
    if (p.nsweeps < 0)
        return cudaSuccess;             /* occupancy query (bp3_max_workers) */
    Bp3Params pp = p;
    pp.pub_workers = (p.pub_workers > 0 && p.team <= 0) ? wpc : 0;
    pp.team = p.team > 0 ? wpc : 0;
    void *args[] = { &pp };
    /*
     * Several ranks sharing one device (p.max_ctas > 0: LocalRanks, tests): the ranks' kernels feed each other, so
     * they must run CONCURRENTLY -- cooperative launches of different streams are not guaranteed to (the driver may
     * hold a cooperative grid back until the device is free of others).  Every rank keeps to its share of the SMs
     * (the sum of all ranks' CTAs does not exceed the SM count), so plain launches are co-resident as well.
     */
    if (p.max_ctas > 0) {
        kern<<<dim3(blocks), dim3(threads), 0, stream>>>(pp);
        return cudaGetLastError();
    }
    /* cooperative launch: fails instead of silently running a non-co-resident grid */
    return cudaLaunchCooperativeKernel((void *)kern, dim3(blocks), dim3(threads), args, 0, stream);
}

#define BP3_CONCAT2(a, b) a##b
#define BP3_CONCAT(a, b) BP3_CONCAT2(a, b)

cudaError_t BP3_CONCAT(bp3_launch_rule, BP3_RULE)(int P, int WPL, const Bp3Params &p, int sms, cudaStream_t stream,
                                                  Bp3LaunchInfo *info)
{
#define BP3_CASE(PP, WW) if (P == PP && WPL == WW) return launch_one<PP, WW>(p, sms, stream, info);
    BP3_CASE(3, 1) BP3_CASE(3, 2) BP3_CASE(3, 4)
    BP3_CASE(4, 1) BP3_CASE(4, 2) BP3_CASE(4, 4)
    BP3_CASE(8, 1) BP3_CASE(8, 2) BP3_CASE(8, 4)
    return cudaErrorInvalidValue;
}

} // namespace clapca
