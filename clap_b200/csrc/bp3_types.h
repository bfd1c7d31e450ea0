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
 * feeds.  Built on the host (bp_plan.h).  On a single GPU every source is the
 * adjacent local plane; in a multi-GPU slab decomposition the plane below the
 * first / above the last plane of a z-block is a "ghost" plane that the
 * neighbouring GPU fills over NVLink (peer stores), and the block's edge
 * planes push their freshly computed H rows into the neighbour's ghost plane.
 *
 * Row r of a source lives at rows + r * stride (words): H0 at +0, H1 at +RWP.
 * The progress counter of a source for generation g lives at flag[g * gstride].
 */
struct Bp3Plane {
    const uint32_t *dn_rows, *up_rows;      /* NULL: outside the volume (reads 0) */
    const int *dn_flag, *up_flag;           /* NULL with rows != NULL never happens */
    uint32_t *push_dn_rows, *push_up_rows;  /* peer ghost planes fed by this plane (NULL: none) */
    int *push_dn_flag, *push_up_flag;
    uint32_t dn_stride, up_stride, dn_gstride, up_gstride;
    uint32_t push_dn_stride, push_up_stride, push_dn_gstride, push_up_gstride;
    uint32_t remote_mask;                   /* bit 0: dn_flag is written by a peer GPU, bit 1: up_flag */
    int zglobal;                            /* global z of this plane (diagnostics) */
};

} // namespace clapca
#endif
