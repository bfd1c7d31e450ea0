/*
 * api_internal.h -- what the translation units behind the C ABI (clapca_api.cu: context, grids, ca3d, ca2d;
 * api_slab.cu: multi-GPU slabs; api_fields.cu: noise / terrain fields) share: the device context, the error
 * channel, small host helpers and the launch planning of the bit-plane sweep.  Not installed, not part of the ABI.
 */
#ifndef CLAPCA_API_INTERNAL_H
#define CLAPCA_API_INTERNAL_H

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <new>

#include "../../include/clapca.h"
#include "bp3_launch.h"
#include "bp_plan.h"

namespace clapca {
namespace api {

/* records the message for clapca_last_error() (per thread) and returns `code` */
int fail(int code, const char *fmt, ...);

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return ::clapca::api::fail(e__ == cudaErrorMemoryAllocation ? CLAPCA_ERR_NOMEM : CLAPCA_ERR_CUDA, \
                        "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

/* generations fused per launch: bounds the progress-counter table, not the result */
static const int kMaxFusedGenerations = 4096;

/* control words of a sweep launch: [0] ticket, [1] err, [4..11] four 64-bit diagnostic cycle counters */
static const int kTicketWords = 16;

struct Ctx {
    int device = -1;
    int sms = 0;
    size_t mem = 0;
    int coop = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_in = nullptr, stream_out = nullptr;     /* copy streams of the streamed ca3d run */
    /* small device scalars shared by the helpers */
    unsigned long long *d_count = nullptr;
    unsigned *d_max = nullptr;
    /* terrain: table of get_avg_height() (field_kernels.cuh), kept between calls */
    void *d_smooth = nullptr;
    size_t smooth_bytes = 0;
    /* grow-only device staging of the one-shot field calls (cudaMalloc / cudaFree of GBs per call costs more than the kernels) */
    void *scratch[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
    void *oneshot_grid = nullptr;   /* the grid of the last one-shot call (clapca_ca3d_run & co.), kept for the next one of the same shape */
    uint32_t *d_zeros = nullptr;    /* 1 KiB of zeros: the H row of a plane outside the volume (Bp3Params::zeros) */
    size_t scratch_bytes[5] = { 0, 0, 0, 0, 0 };
};
extern Ctx g_ctx;

int need_init();
int grid_blocks_for(size_t work_items, int threads, int per_sm_cap = 8);
/* grow-only device buffer */
int ensure_bytes(void **ptr, size_t *have, size_t want);
/* synchronise the context's stream, then the elapsed time between two events recorded on it */
int timed_sync(cudaEvent_t a, cudaEvent_t b, float *ms);

/* the nine cas[] entries: core/ca3d.c:110-122 (surv, born, nr_states) */
extern const uint32_t kCas[9][3];

bool diag_enabled();
void diag_report(const char *what, const unsigned *d_ticket, cudaStream_t stream, int rank);

/* ---- launch planning of the bit-plane sweep (clapca_api.cu), shared with the slab path ---- */

/* work-item claim order of the bit-plane sweep (see bp_plan.h) */
struct OrderCfg {
    int mode;           /* 0 time-key, 1 skewed row segments, 2 generation-batched diagonals, 3 tiles (planes x generations) */
    int seg_rows;       /* mode 1 */
    int gen_batch;      /* mode 2 */
    int team;           /* mode 3: compute warps per CTA = tile_z * tile_g */
    int tile_z, tile_g; /* mode 3: planes / generations per tile */
    int ctas;           /* mode 3: CTAs the launch keeps resident (bounds tile_g, see bp_plan.h) */
    int tile_skew;      /* mode 3: key distance between generation groups (0 = tile_z + 1, see bp3_make_items_tile) */
    int key() const
    {
        return mode * 100000 + (mode == 1 ? seg_rows : (mode == 2 ? gen_batch : (mode == 3 ? tile_z * 256 + tile_g : 0)))
             + (mode == 3 ? (tile_skew % 4000) * 500000 : 0);
    }
};
int team_config(int P, int WPL, bool single_gpu = false);
/* wanted generations per tile for a team of `team` compute warps (CLAPCA_TILE_GENS overrides) */
int tile_gens_config(int team);
void sweep_knobs(Bp3Params &p, int team, bool single_gpu = false);
OrderCfg order_config(int Z, int H, int G, int max_workers, int team);
void make_items(const OrderCfg &oc, const std::vector<Bp3Plane> &planes, int Zg, int H, int G,
                std::vector<WorkItem> &items, bool layout_items);

/* host side of a streamed run: see stream_volume() in clapca_api.cu */
struct StreamPlan {
    uint8_t *d_cells;           /* device cells of the (local) planes, reference layout */
    const uint8_t *host_in;     /* pinned */
    uint8_t *host_out;          /* pinned */
    size_t plane_bytes;
    int Z, chunk;               /* planes, planes per H2D chunk */
    int *d_in_ready;            /* device word: chunks landed */
    int *h_io;                  /* pinned + mapped: [0, Z) per-plane done flags, then the chunk ordinals 1, 2, ... */
    int epoch;                  /* value the unpack items store into the done flags */
    cudaStream_t s_kernel, s_in, s_out;
};
int stream_volume(const StreamPlan &sp, bool *stuck);
/* planes per H2D chunk of a streamed run */
int io_chunk_planes(size_t plane_bytes, int Z);

/* the layout kernels live in ONE translation unit (clapca_api.cu); these launch them */
cudaError_t launch_ca3d_pack(const Bp3Layout &L, cudaStream_t stream);
cudaError_t launch_ca3d_unpack(const Bp3Layout &L, cudaStream_t stream);
cudaError_t preload_ca3d_layout(int P);
cudaError_t launch_max_u8(const uint8_t *cells, size_t n, unsigned *d_max, cudaStream_t stream);
cudaError_t launch_halo_seed(uint32_t *dst, const uint32_t *src, int H, int RWP, int NP, cudaStream_t stream);

} // namespace api
} // namespace clapca

#endif /* CLAPCA_API_INTERNAL_H */
