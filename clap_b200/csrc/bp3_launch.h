/*
 * bp3_launch.h -- host-side launcher interface of the bit-plane ca3d sweep
 * kernels.  Each of the nine cas[] rules (core/ca3d.c:110-122) plus the
 * run-time-mask rule is compiled in its own translation unit
 * (bp3_inst.cu with -DBP3_RULE=n) so the builds run in parallel.
 */
#ifndef CLAPCA_BP3_LAUNCH_H
#define CLAPCA_BP3_LAUNCH_H
#include <cuda_runtime.h>
#include "ca3d_bitplane.cuh"

namespace clapca {

enum { BP3_RULE_DYN = 9, BP3_NRULES = 10 };

struct Bp3LaunchInfo {
    int blocks, threads, workers, regs;
};

/* cooperative launch of the sweep kernel for (rule, P in {3,4,8}, WPL in {1,2,4}) */
cudaError_t bp3_launch(int rule, int P, int WPL, const Bp3Params &p, int sms, cudaStream_t stream,
                       Bp3LaunchInfo *info);

/* number of persistent (compute) warps a cooperative launch of this variant can keep resident; team > 0: tile mode */
int bp3_max_workers(int rule, int P, int WPL, int sms, int team = 0, int max_ctas = 0);

#define BP3_DECLARE_RULE(n) \
    cudaError_t bp3_launch_rule##n(int P, int WPL, const Bp3Params &p, int sms, cudaStream_t stream, \
                                   Bp3LaunchInfo *info);
BP3_DECLARE_RULE(0) BP3_DECLARE_RULE(1) BP3_DECLARE_RULE(2) BP3_DECLARE_RULE(3) BP3_DECLARE_RULE(4)
BP3_DECLARE_RULE(5) BP3_DECLARE_RULE(6) BP3_DECLARE_RULE(7) BP3_DECLARE_RULE(8) BP3_DECLARE_RULE(9)

} // namespace clapca
#endif
