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
 * Sources.  Row r of a neighbouring plane lives at rows + r * RECW (words, RECW =
 * NP * RWP = the row-record stride of the launched variant): H0 at +0, H1 at
 * +RWP; its progress counter for generation g at flag[g * gstride].
 *
 * Ghost sources (multi-GPU).  The plane below the first / above the last plane
 * of a z-block belongs to the neighbouring GPU.  Its H rows are mirrored into a
 * "ghost plane" in THIS GPU's memory -- same record stride as a local plane, only
 * the H0 | H1 part of a record is ever written -- together with one progress
 * counter per generation (gstride = 1).  The owner's service warp (tile mode,
 * ca3d_bitplane.cuh) copies finished rows over NVLink with plain peer stores,
 * then fence.acq_rel.sys, then the counter, so a ghost plane is an ordinary
 * source for its consumer: the row loop has no multi-GPU code.  ghost_mask only
 * tells the consumer to poll that counter at system scope.  The counters exist
 * in two banks used by alternating runs (the host picks the bank when it builds
 * the descriptors), so a slow peer still finishing run k can never disturb the
 * freshly cleared counters of run k+1.
 */
struct Bp3Plane {
    const uint32_t *dn_rows, *up_rows;      /* NULL: outside the volume (reads 0) */
    const int *dn_flag, *up_flag;           /* progress counters of the source, generation g at [g * gstride] */
    uint32_t *push_dn_rows, *push_up_rows;  /* peer ghost planes fed by this plane (NULL: none) */
    int *push_dn_flag, *push_up_flag;       /* ... and their counters, generation g at [g] */
    uint32_t dn_gstride, up_gstride;
    uint32_t ghost_mask;                    /* bit 0: dn is a ghost plane, bit 1: up is a ghost plane */
    int zglobal;                            /* global z of this plane */
};

/* arguments of the layout kernels (ca3d_layout.cuh): reference-layout cells <-> row records */
struct Bp3Layout {
    uint8_t *cells;         /* reference layout, z*W*H + y*W + x */
    uint32_t *rows;         /* row records */
    int W, H, Z, P, RWP;
    unsigned long long *population;     /* unpack: number of non-zero cells */
    unsigned *over;         /* pack, optional: |= 1 when a cell value does not fit the P state planes */
};

} // namespace clapca
#endif
