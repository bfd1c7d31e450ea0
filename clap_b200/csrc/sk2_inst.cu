/*
 * sk2_inst.cu -- instantiates ca2d_skew_kernel for ONE rule (-DSK2_RULE=0..2: run-time masks, cave smoothing,
 * ca_test) and every (WPL, neighbourhood) variant, and provides its cooperative launcher.
 */
#include "sk2_launch.h"

#ifndef SK2_RULE
#error "compile with -DSK2_RULE=<0..2>"
#endif

namespace clapca {

#if SK2_RULE == 1
typedef Sk2RuleCave TheRuleSk;
#elif SK2_RULE == 2
typedef Sk2RuleTest TheRuleSk;
#else
typedef Sk2RuleDyn TheRuleSk;
#endif

template <int WPL, bool MOORE>
static cudaError_t launch_one(int warps, const Sk2Params &p, int sms, cudaStream_t stream, Bp2LaunchInfo *info)
{
    auto kern = ca2d_skew_kernel<WPL, MOORE, TheRuleSk>;
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
    Sk2Params pp = p;
    void *args[] = { &pp };
    /* cooperative launch: the generation pipeline needs every claimed generation resident */
    return cudaLaunchCooperativeKernel((void *)kern, dim3(blocks), dim3(threads), args, 0, stream);
}

#define SK2_CONCAT2(a, b) a##b
#define SK2_CONCAT(a, b) SK2_CONCAT2(a, b)

cudaError_t SK2_CONCAT(sk2_launch_rule, SK2_RULE)(int WPL, bool moore, int warps, const Sk2Params &p, int sms,
                                                  cudaStream_t stream, Bp2LaunchInfo *info)
{
    if (warps < 1 || warps * WPL > SK2_MAX_WARPS)
        return cudaErrorInvalidValue;
    if (WPL == 1)
        return moore ? launch_one<1, true>(warps, p, sms, stream, info) : launch_one<1, false>(warps, p, sms, stream, info);
    if (WPL == 2)
        return moore ? launch_one<2, true>(warps, p, sms, stream, info) : launch_one<2, false>(warps, p, sms, stream, info);
    return cudaErrorInvalidValue;
}

} // namespace clapca
