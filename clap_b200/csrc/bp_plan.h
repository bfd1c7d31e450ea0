/*
 * bp_plan.h -- host-side planning shared by the C-ABI library and the kernel
 * emulator tests: sweep claim order and bit-plane count for the bit-plane
 * engines.  Plain C++ (no CUDA).
 */
#ifndef CLAPCA_BP_PLAN_H
#define CLAPCA_BP_PLAN_H
#include <stdint.h>
#include <vector>
#include <algorithm>

namespace clapca {

struct SweepId { int z, g; };

/*
 * Claim order of the (z,g) sweeps.  Sweep (z,g) consumes (z-1,g), (z+1,g-1)
 * and (z,g-1); any order that is monotone in key = 2z + 4g extends that
 * partial order (the three producers have keys key-2, key-2, key-4), and it
 * keeps all generations of a plane a few rows apart (L2-resident window).
 */
inline void bp3_make_order(int Z, int G, std::vector<SweepId> &order)
{
    order.clear();
    order.reserve((size_t)Z * G);
    const long long kmax = 2LL * (Z - 1) + 4LL * (G - 1);
    for (long long key = 0; key <= kmax; key += 2)
        for (int g = 0; g < G; g++) {
            long long z2 = key - 4LL * g;
            if (z2 < 0)
                break;
            if (z2 / 2 < Z)
                order.push_back(SweepId{ (int)(z2 / 2), g });
        }
}

/* bit planes needed so that every value that can ever occur fits: {3,4,8} are instantiated */
inline int bp_planes_for(unsigned maxval)
{
    if (maxval < 8) return 3;
    if (maxval < 16) return 4;
    return 8;
}

/* words per lane for a row of W cells handled by one warp: {1,2,4}, 0 = too wide */
inline int bp_wpl_for(int W)
{
    if (W <= 1024) return 1;
    if (W <= 2048) return 2;
    if (W <= 4096) return 4;
    return 0;
}

} // namespace clapca
#endif
