/*
 * bp2_launch.h -- host-side launcher interface of the 2D bit-plane sweep kernels (ca2d_bitplane.cuh),
 * instantiated in bp2_inst.cu for P in {1,3,4,8} x WPL in {1,2,4} x {von Neumann, Moore}.
 */
#ifndef CLAPCA_BP2_LAUNCH_H
#define CLAPCA_BP2_LAUNCH_H
#include <cuda_runtime.h>
#include "ca2d_bitplane.cuh"

namespace clapca {

struct Bp2LaunchInfo {
    int blocks, threads, regs;
};

/*
 * Cooperative launch with `warps` compute warps per CTA (the row is split across them; one more warp publishes the
 * progress counters) and at most p.G CTAs.  The rule instantiation follows from p.born / p.surv / p.nrval
 * (bp2_rule_for).  p.G < 0: occupancy query only (info->blocks = CTAs the device can keep resident).
 */
cudaError_t bp2_launch(int P, int WPL, bool moore, int warps, const Bp2Params &p, int sms, cudaStream_t stream,
                       Bp2LaunchInfo *info);

/* one translation unit per rule (bp2_inst.cu with -DBP2_RULE=n): the builds run in parallel */
#define BP2_DECLARE_RULE(n) \
    cudaError_t bp2_launch_rule##n(int P, int WPL, bool moore, int warps, const Bp2Params &p, int sms, \
                                   cudaStream_t stream, Bp2LaunchInfo *info);
BP2_DECLARE_RULE(0) BP2_DECLARE_RULE(1) BP2_DECLARE_RULE(2)

/* state planes for values up to maxval: {1,3,4,8} */
inline int bp2_planes_for(unsigned maxval)
{
    if (maxval < 2) return 1;
    if (maxval < 8) return 3;
    if (maxval < 16) return 4;
    return 8;
}

/*
 * Words per lane and warps per CTA for a row of N cells; false = too wide.  Measured on B200
 * (profiles/r01_knobs_ca2d.txt): the per-warp cost of the warp- and CTA-level scan stages favours at most
 * 8 warps per row (16384 cells: 2 words per lane x 8 warps, 12.2 ms vs 14.5 ms for 1 x 16).
 */
inline bool bp2_shape_for(long long N, int P, int *WPL, int *warps)
{
    const long long rw = (N + 31) / 32;
    const int max_wpl = P >= 8 ? 2 : 4;        /* register budget of the state pipeline */
    for (int cap = 8; cap <= 16; cap *= 2)
        for (int wpl = 1; wpl <= max_wpl; wpl *= 2) {
            const long long nw = (rw + 32 * wpl - 1) / (32 * wpl);
            if (nw <= cap) {
                *WPL = wpl;
                *warps = (int)(nw < 1 ? 1 : nw);
                return true;
            }
        }
    return false;
}

} // namespace clapca
#endif
