/*
 * bp2_inst.cu -- instantiates ca2d_sweep_kernel for ONE rule (-DBP2_RULE=0..2: run-time masks, cave smoothing,
 * ca_test) and every (P, WPL, neighbourhood) variant, and provides its cooperative launcher.
 */
#include "bp2_launch.h"

#ifndef BP2_RULE
#error "compile with -DBP2_RULE=<0..2>"
#endif

namespace clapca {

#if BP2_RULE == 1
typedef Rule2Cave TheRule2;
#elif BP2_RULE == 2
typedef Rule2Test TheRule2;
#else
typedef Rule2Dyn TheRule2;
#endif

template <int P, int WPL, bool MOORE>
static cudaError_t launch_one(int warps, const Bp2Params &p, int sms, cudaStream_t stream, Bp2LaunchInfo *info)
{
    auto kern = ca2d_sweep_kernel<P, WPL, MOORE, TheRule2>;
    const int threads = (warps + 1) * 32;       /* the compute warps + the publisher warp */
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    int blocks = per_sm * sms;
    if (p.G >= 0 && blocks > p.G) blocks = p.G;
    if (blocks < 1) blocks = 1;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    if (info) {
        info->blocks = blocks;
        info->threads = threads;
        info->regs = fa.numRegs;
    }
    if (p.G < 0)
        return cudaSuccess;
    Bp2Params pp = p;
    void *args[] = { &pp };
    /* cooperative launch: the generation pipeline needs every claimed generation resident */
    return cudaLaunchCooperativeKernel((void *)kern, dim3(blocks), dim3(threads), args, 0, stream);
}

#define BP2_CONCAT2(a, b) a##b
#define BP2_CONCAT(a, b) BP2_CONCAT2(a, b)

cudaError_t BP2_CONCAT(bp2_launch_rule, BP2_RULE)(int P, int WPL, bool moore, int warps, const Bp2Params &p, int sms,
                                                  cudaStream_t stream, Bp2LaunchInfo *info)
{
    if (warps < 1 || warps > BP2_MAX_WARPS)
        return cudaErrorInvalidValue;
#define BP2_CASE(PP, WW) \
    if (P == PP && WPL == WW) \
        return moore ? launch_one<PP, WW, true>(warps, p, sms, stream, info) \
                     : launch_one<PP, WW, false>(warps, p, sms, stream, info);
    BP2_CASE(1, 1) BP2_CASE(1, 2) BP2_CASE(1, 4)
    BP2_CASE(3, 1) BP2_CASE(3, 2) BP2_CASE(3, 4)
    BP2_CASE(4, 1) BP2_CASE(4, 2) BP2_CASE(4, 4)
    BP2_CASE(8, 1) BP2_CASE(8, 2)
    return cudaErrorInvalidValue;
}

} // namespace clapca
