/*
 * bp3_types.h -- plain-data types shared by the bit-plane ca3d kernels
 * (ca3d_bitplane.cuh) and the host-side planner (bp_plan.h).
 */
#ifndef CLAPCA_BP3_TYPES_H
#define CLAPCA_BP3_TYPES_H
#include <stdint.h>

namespace clapca {

/*
 * Where one plane of this device's slab gets its neighbours from and whom it
 * feeds.  Built on the host (bp_plan.h).
 *
 * Local sources.  Row r of a neighbouring local plane lives at rows + r * stride
 * (words): H0 at +0, H1 at +RWP; its progress counter for generation g at
 * flag[g * gstride].
 *
 * Ghost sources (multi-GPU).  The plane below the first / above the last plane
 * of a z-block belongs to the neighbouring GPU, which stores the H rows of its
 * edge plane straight into a "ghost plane" in this GPU's memory (peer stores
 * over NVLink).  Ghost rows carry no counters and need no fences: every 32-bit
 * data word travels in one aligned 8-byte store together with a tag
 *         tag = epoch << 16 | (generation + 1)        (0 = the seed state)
 * so a consumer that reads {word, tag} atomically knows which version it got
 * and simply re-reads until the expected tag shows up (the LL protocol of
 * collective libraries).  A lane's 2*WPL words of a ghost row are the pairs
 * {H0[0..WPL), H1[0..WPL)}, two pairs per 16-byte vector; vector i of lane l sits
 * at 16-byte index r * stride/4 + 32 * i + l, so every warp-wide store of a row is
 * ONE contiguous 512-byte burst on NVLink (16-byte scattered peer stores are what
 * the link is worst at).
 */
struct Bp3Plane {
    const uint32_t *dn_rows, *up_rows;      /* NULL: outside the volume (reads 0) */
    const int *dn_flag, *up_flag;           /* NULL for ghost (tagged) sources */
    uint32_t *push_dn_rows, *push_up_rows;  /* peer ghost planes fed by this plane (NULL: none) */
    uint32_t dn_stride, up_stride, dn_gstride, up_gstride;
    uint32_t push_dn_stride, push_up_stride;
    uint32_t ghost_mask;                    /* bit 0: dn is a ghost plane, bit 1: up is a ghost plane */
    int zglobal;                            /* global z of this plane */
};

/* arguments of the layout kernels (ca3d_layout.cuh): reference-layout cells <-> row records */
struct Bp3Layout {
    uint8_t *cells;         /* reference layout, z*W*H + y*W + x */
    uint32_t *rows;         /* row records */
    int W, H, Z, P, RWP;
    unsigned long long *population;     /* unpack: number of non-zero cells */
};

} // namespace clapca
#endif
