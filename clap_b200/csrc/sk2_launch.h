/*
 * sk2_launch.h -- host-side launcher interface of the diagonal 2D sweep (ca2d_skew.cuh), instantiated in
 * sk2_inst.cu for WPL in {1,2} x {von Neumann, Moore} and one rule per translation unit.
 */
#ifndef CLAPCA_SK2_LAUNCH_H
#define CLAPCA_SK2_LAUNCH_H
#include <cuda_runtime.h>
#include "ca2d_skew.cuh"
#include "bp2_launch.h"

namespace clapca {

/*
 * Cooperative launch with `warps` compute warps per CTA (a diagonal is cut across them by x; one more warp publishes
 * the progress counters) and at most p.G CTAs.  The rule instantiation follows from p.born / p.surv / p.nrval
 * (bp2_rule_for).  p.G < 0: occupancy query only.
 */
cudaError_t sk2_launch(int WPL, bool moore, int warps, const Sk2Params &p, int sms, cudaStream_t stream,
                       Bp2LaunchInfo *info);

#define SK2_DECLARE_RULE(n) \
    cudaError_t sk2_launch_rule##n(int WPL, bool moore, int warps, const Sk2Params &p, int sms, cudaStream_t stream, \
                                   Bp2LaunchInfo *info);
SK2_DECLARE_RULE(0) SK2_DECLARE_RULE(1) SK2_DECLARE_RULE(2)

/*
 * Words per lane and warps per CTA for diagonals of W cells (x extent); false = too wide (more than 16384 cells).
 * Two words per lane halve the per-step overhead (mailbox, counters, loop) per cell; one word per lane only where
 * the grid is too narrow to give every scheduler of the SM a warp otherwise.
 */
inline bool sk2_shape_for(long long W, int *WPL, int *warps)
{
    if (W < 1 || W > 16384)
        return false;
    *WPL = W <= 4096 ? 1 : 2;
    *warps = sk2_warps_for((int)W, *WPL);
    return *warps * *WPL <= SK2_MAX_WARPS;
}

} // namespace clapca
#endif
