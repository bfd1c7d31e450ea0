/*
 * api_slab.cu -- multi-GPU part of the C ABI (include/clapca.h, clapca_slab_*): z-block slabs of one ca3d
 * volume, one process per GPU (or, for tests, several slabs of one process sharing a device), halo rows exchanged
 * by peer stores inside the sweep kernel: the service warp of a tile that touches a z-block edge copies finished H
 * rows into the neighbour's ghost plane and raises the neighbour's progress counter (ca3d_bitplane.cuh).
 */
#include "api_internal.h"

using namespace clapca;
using namespace clapca::api;

extern "C" {
#pragma GCC visibility push(default)

/* ---- multi-GPU z-block slabs ------------------------------------------------------ */

struct clapca_slab {
    SlabGeom geo;
    HaloLayout hl;
    int W, H, P, WPL, RWP, NP, Gcap, Zl;
    uint8_t *cells = nullptr;           /* local planes, reference layout, local order */
    uint32_t *rows = nullptr;
    uint32_t *halo = nullptr;           /* exported to the neighbours */
    uint32_t *halo_next = nullptr, *halo_prev = nullptr;
    bool opened_next = false, opened_prev = false;
    int *prog = nullptr;
    Bp3Plane *planes = nullptr;
    std::vector<Bp3Plane> h_planes;
    int4 *order = nullptr;
    size_t order_bytes = 0;
    int n_items = 0, order_G = -1, team = -1, order_key = -1;
    unsigned *ticket = nullptr;
    unsigned long long *d_pop = nullptr;
    unsigned *d_max = nullptr;
    cudaStream_t stream = nullptr;      /* the slab's own stream: slabs of one process run concurrently */
    int max_ctas = 0;                   /* > 0: CTAs of the sweep launch (ranks sharing one device) */
    unsigned max_value = 0;
    /* streamed runs (layout items): copy streams, chunk-arrival word, per-plane done flags in host-mapped memory */
    cudaStream_t stream_in = nullptr, stream_out = nullptr;
    int *d_in_ready = nullptr;
    int *h_io = nullptr;
    size_t h_io_count = 0;
    int io_epoch = 0, io_chunk = 1;
    cudaEvent_t ev_io = nullptr;
    bool streamed_order = false;        /* the claim order on the device holds pack / unpack items */
    cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
    uint32_t surv = 0, born = 0, nr_states = 0;
    int G = 0, rule = BP3_RULE_DYN;
    bool prepared = false;
    uint32_t epoch = 0;             /* run number: its parity selects the bank of ghost counters */
    clapca_run_stats stats;
};

int clapca_slab_create(clapca_slab **out, int64_t d0, int64_t d1, int64_t d2_global, int rank, int nranks,
                       int block_planes, int max_generations, unsigned max_value)
{
    if (int rc = need_init()) return rc;
    if (!out || d0 < 1 || d1 < 1 || d2_global < 1 || nranks < 1 || rank < 0 || rank >= nranks || block_planes < 1 ||
        max_generations < 1 || max_generations > kMaxFusedGenerations)
        return fail(CLAPCA_ERR_ARG, "slab_create: bad arguments");
    const int WPL = bp_wpl_for((int)d0);
    if (!WPL || d1 >= (1 << 30) || d2_global >= (1 << 30))
        return fail(CLAPCA_ERR_UNSUPPORTED, "slab_create: rows of at most 4096 cells are supported (d0 = %lld)",
                    (long long)d0);
    clapca_slab *s = new (std::nothrow) clapca_slab();
    if (!s) return fail(CLAPCA_ERR_NOMEM, "slab_create: host allocation failed");
    s->geo = SlabGeom{ (int)d2_global, nranks, rank, nranks == 1 ? (int)d2_global : block_planes };
    s->W = (int)d0; s->H = (int)d1;
    s->P = bp_planes_for(max_value);
    s->WPL = WPL; s->RWP = 32 * WPL; s->NP = s->P + 2;
    s->Gcap = max_generations;
    s->Zl = s->geo.local_planes();
    s->hl = slab_halo_layout(s->geo, s->H, s->RWP, s->NP, s->Gcap);
    s->max_value = max_value;
    memset(&s->stats, 0, sizeof(s->stats));
    const size_t zl = s->Zl ? s->Zl : 1;
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&s->cells, zl * s->W * s->H);
    if (e == cudaSuccess) e = cudaMalloc(&s->rows, zl * s->H * s->NP * s->RWP * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&s->halo, s->hl.total_words * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(s->halo, 0, s->hl.total_words * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&s->prog, (size_t)(s->Gcap + 1) * zl * sizeof(int));   /* row 0: "generation -1" (pack items) */
    if (e == cudaSuccess) e = cudaMalloc(&s->planes, zl * sizeof(Bp3Plane));
    if (e == cudaSuccess) e = cudaMalloc(&s->ticket, kTicketWords * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_pop, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_max, sizeof(unsigned));
    for (int i = 0; i < 5 && e == cudaSuccess; i++)
        e = cudaEventCreate(&s->ev[i]);
    if (e != cudaSuccess) {
        clapca_slab_destroy(s);
        return fail(e == cudaErrorMemoryAllocation ? CLAPCA_ERR_NOMEM : CLAPCA_ERR_CUDA, "slab_create: %s",
                    cudaGetErrorString(e));
    }
    *out = s;
    return CLAPCA_OK;
}

int clapca_slab_destroy(clapca_slab *s)
{
    if (!s) return CLAPCA_OK;
    if (s->opened_next && s->halo_next) cudaIpcCloseMemHandle(s->halo_next);
    if (s->opened_prev && s->halo_prev) cudaIpcCloseMemHandle(s->halo_prev);
    if (s->stream) cudaStreamSynchronize(s->stream);
    void *bufs[] = { s->cells, s->rows, s->halo, s->prog, s->planes, s->order, s->ticket, s->d_pop, s->d_max };
    for (void *b : bufs)
        if (b) cudaFree(b);
    if (s->d_in_ready) cudaFree(s->d_in_ready);
    if (s->h_io) cudaFreeHost(s->h_io);
    if (s->ev_io) cudaEventDestroy(s->ev_io);
    if (s->stream_in) cudaStreamDestroy(s->stream_in);
    if (s->stream_out) cudaStreamDestroy(s->stream_out);
    if (s->stream) cudaStreamDestroy(s->stream);
    for (int i = 0; i < 5; i++)
        if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    delete s;
    return CLAPCA_OK;
}

int clapca_slab_local_planes(clapca_slab *s, int *n)
{
    if (!s || !n) return fail(CLAPCA_ERR_ARG, "slab_local_planes: NULL argument");
    *n = s->Zl;
    return CLAPCA_OK;
}

int clapca_slab_plane_map(clapca_slab *s, int64_t *zglobal)
{
    if (!s || !zglobal) return fail(CLAPCA_ERR_ARG, "slab_plane_map: NULL argument");
    for (int lb = 0; lb < s->geo.local_blocks(); lb++) {
        const int j = s->geo.global_block(lb), l0 = s->geo.local_z0(lb);
        for (int i = 0; i < s->geo.block_len(j); i++)
            zglobal[l0 + i] = s->geo.block_z0(j) + i;
    }
    return CLAPCA_OK;
}

void *clapca_slab_device_ptr(clapca_slab *s) { return s ? s->cells : nullptr; }

int clapca_slab_ipc_handle(clapca_slab *s, void *handle64)
{
    if (!s || !handle64) return fail(CLAPCA_ERR_ARG, "slab_ipc_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->halo));
    memcpy(handle64, &h, sizeof(h));
    return CLAPCA_OK;
}

/* this rank's plane descriptors for the bank of ghost counters the next run uses */
static int slab_build_planes(clapca_slab *s, int bank)
{
    SlabPtrs ptr = { s->rows, s->prog + (s->Zl ? s->Zl : 1), s->halo, s->halo_next, s->halo_prev };
    bp3_build_planes(s->geo, ptr, s->hl, s->H, s->RWP, s->NP, s->h_planes, bank);
    if (s->Zl)
        CU(cudaMemcpyAsync(s->planes, s->h_planes.data(), s->h_planes.size() * sizeof(Bp3Plane), cudaMemcpyHostToDevice,
                           s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return CLAPCA_OK;
}

int clapca_slab_connect(clapca_slab *s, const void *handle_next, const void *handle_prev)
{
    if (!s) return fail(CLAPCA_ERR_ARG, "slab_connect: NULL slab");
    const int R = s->geo.R;
    if (R == 1) {
        s->halo_next = s->halo_prev = s->halo;
    } else {
        if (!handle_next || !handle_prev) return fail(CLAPCA_ERR_ARG, "slab_connect: NULL handle");
        cudaIpcMemHandle_t hn, hp;
        memcpy(&hn, handle_next, sizeof(hn));
        memcpy(&hp, handle_prev, sizeof(hp));
        void *pn = nullptr, *pp = nullptr;
        CU(cudaIpcOpenMemHandle(&pn, hn, cudaIpcMemLazyEnablePeerAccess));
        s->halo_next = (uint32_t *)pn;
        s->opened_next = true;
        if (R == 2) {
            s->halo_prev = s->halo_next;        /* both neighbours are the same rank: one mapping */
        } else {
            CU(cudaIpcOpenMemHandle(&pp, hp, cudaIpcMemLazyEnablePeerAccess));
            s->halo_prev = (uint32_t *)pp;
            s->opened_prev = true;
        }
    }
    return slab_build_planes(s, 0);
}

void *clapca_slab_halo_ptr(clapca_slab *s) { return s ? s->halo : nullptr; }

int clapca_slab_connect_local(clapca_slab *s, void *halo_next_rank, void *halo_prev_rank, int max_ctas)
{
    if (!s) return fail(CLAPCA_ERR_ARG, "slab_connect_local: NULL slab");
    if (s->geo.R > 1 && (!halo_next_rank || !halo_prev_rank)) return fail(CLAPCA_ERR_ARG, "slab_connect_local: NULL halo");
    s->halo_next = s->geo.R > 1 ? (uint32_t *)halo_next_rank : s->halo;
    s->halo_prev = s->geo.R > 1 ? (uint32_t *)halo_prev_rank : s->halo;
    s->max_ctas = max_ctas > 0 ? max_ctas : 0;
    s->order_G = -1;                    /* the tile shape depends on the number of CTAs */
    return slab_build_planes(s, 0);
}

int clapca_slab_upload(clapca_slab *s, const uint8_t *src)
{
    if (!s || !src) return fail(CLAPCA_ERR_ARG, "slab_upload: NULL argument");
    if (s->Zl) {
        CU(cudaMemcpyAsync(s->cells, src, (size_t)s->Zl * s->W * s->H, cudaMemcpyDefault, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    return CLAPCA_OK;
}

int clapca_slab_download(clapca_slab *s, uint8_t *dst)
{
    if (!s || !dst) return fail(CLAPCA_ERR_ARG, "slab_download: NULL argument");
    if (s->Zl) {
        CU(cudaMemcpyAsync(dst, s->cells, (size_t)s->Zl * s->W * s->H, cudaMemcpyDefault, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    return CLAPCA_OK;
}

/*
 * Step 1 of a sharded run (all ranks, then a barrier).  Resident runs: lay the local planes out as bit-plane row
 * records, seed the neighbour's ghost planes with the H rows of every block's first plane (the "old plane above" of
 * generation 0).  Streamed runs: nothing is on the device yet -- the pack items of the launch do both.  Either way:
 * the claim order for `steps` generations, this run's bank of ghost counters cleared, the progress table cleared.
 */
static int slab_prepare(clapca_slab *s, uint32_t surv, uint32_t born, uint32_t nr_states, int steps, bool streamed)
{
    if (int rc = need_init()) return rc;
    if (!s) return fail(CLAPCA_ERR_ARG, "slab_prepare: NULL slab");
    if (s->h_planes.empty() && s->Zl) return fail(CLAPCA_ERR_STATE, "slab_prepare: call clapca_slab_connect first");
    if (steps < 1 || steps > s->Gcap) return fail(CLAPCA_ERR_ARG, "slab_prepare: steps %d outside 1..%d", steps, s->Gcap);
    const uint32_t bornval = (nr_states - 1u) & 0xffu;
    if (born && (bornval >> s->P))
        return fail(CLAPCA_ERR_ARG, "slab_prepare: rule needs more than the %d state planes of this slab", s->P);
    s->surv = surv; s->born = born; s->nr_states = nr_states; s->G = steps;
    s->rule = BP3_RULE_DYN;
    for (int i = 0; i < 9; i++)
        if (kCas[i][0] == surv && kCas[i][1] == born && kCas[i][2] == nr_states) { s->rule = i; break; }

    /* the cells must fit the state planes chosen at create time: the pack kernel would silently drop the upper bits
       (streamed runs: the pack items check while they convert, err = 4) */
    if (s->Zl && !streamed) {
        unsigned maxv = 0;
        CU(cudaMemsetAsync(s->d_max, 0, sizeof(unsigned), s->stream));
        CU(launch_max_u8(s->cells, (size_t)s->Zl * s->W * s->H, s->d_max, s->stream));
        CU(cudaMemcpyAsync(&maxv, s->d_max, sizeof(maxv), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
        if (s->P < 8 && (maxv >> s->P))
            return fail(CLAPCA_ERR_ARG, "slab_prepare: a cell value of %u exceeds the max_value %u given to slab_create "
                        "(%d state planes)", maxv, s->max_value, s->P);
    }
    /*
     * Every kernel a rank launches while a neighbour's sweep may already be spinning on this rank's rows must be
     * LOADED by now: with lazy module loading (the CUDA 12 default) the first launch of a kernel loads it, loading
     * synchronises with the running work of the context, and a rank of the same process (LocalRanks, or a caller's
     * threads) would wait for a sweep that waits for it -- measured: the first P = 8 run of a process sat out both
     * watchdogs in exactly that way, behind the unpack kernel its neighbour launches right after its sweep.  The sweep
     * kernel itself is loaded by the occupancy query of order_config() below.
     */
    CU(preload_ca3d_layout(s->P));
    s->epoch++;                                 /* every rank prepares successfully the same number of times */
    const int bank = (int)(s->epoch & 1u);

    /* multi-GPU runs are always tiled: the service warp of a tile carries the halo rows (every variant has a team size) */
    /* the wide kernel where it exists (3-plane variants): it wins at 2, 4 and 8 GPUs as well, with z-blocks that are a
       multiple of its 6-plane tiles (profiles/r02_knobs_multi_wide_n{2,4,8}.txt) */
    int team = team_config(s->P, s->WPL, true);
    if (team <= 0) team = bp3_team_cap(s->P, s->WPL);
    Bp3Params knobs;
    memset(&knobs, 0, sizeof(knobs));
    sweep_knobs(knobs, team);
    const int max_ctas = s->max_ctas > 0 ? s->max_ctas : knobs.max_ctas;
    OrderCfg oc = order_config(s->geo.Zg, s->H, steps, bp3_max_workers(s->rule, s->P, s->WPL, g_ctx.sms, team, max_ctas), team);
    /* sharded volumes keep generation group b + 1 right behind group b: the z pipeline across the ranks needs the
       other groups to fill its bubbles (CLAPCA_SLAB_TILE_SKEW for experiments; every rank must use the same value) */
    /* ... but not glued together either: four tiles between consecutive groups of a plane group (tz + 1 = all groups
       interleaved: N = 4 / 8 sweep 28.2 / 16.9 ms; 24 planes: 26.2 / 16.4 ms; 48: 16.5; 100: 27.7 ms at N = 4;
       profiles/r02_knobs_multi_slab_skew.txt) */
    oc.tile_skew = 4 * std::max(1, team / std::max(1, oc.tile_g));
    if (streamed) oc.tile_skew = 0;     /* layout items: the output of a plane follows its input one group distance per
                                           generation group behind -- keep the groups glued (the measured e2e setting) */
    if (const char *e = getenv("CLAPCA_SLAB_TILE_SKEW")) oc.tile_skew = std::max(0, atoi(e));
    /* every rank must settle on the same tile shape: decide it from all ranks' plane lists */
    oc.tile_g = bp3_tile_shape_all_ranks(s->geo, s->H, steps, team, oc.tile_g, oc.ctas, &oc.tile_z, streamed, oc.tile_skew);
    if (s->order_G != steps || s->team != team || s->order_key != oc.key() || s->streamed_order != streamed) {
        std::vector<WorkItem> items;
        s->team = team;
        make_items(oc, s->h_planes, s->geo.Zg, s->H, steps, items, streamed);
        void *p = s->order;
        if (int rc = ensure_bytes(&p, &s->order_bytes, (items.size() ? items.size() : 1) * sizeof(int4))) {
            s->order = nullptr;
            return rc;
        }
        s->order = (int4 *)p;
        if (!items.empty())
            CU(cudaMemcpy(s->order, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice));
        s->n_items = (int)items.size();
        s->order_G = steps;
        s->order_key = oc.key();
        s->streamed_order = streamed;
    }
    /* ghost counters of this run's bank: cleared here, raised by the neighbours only after the barrier that follows */
    if (int rc = slab_build_planes(s, bank)) return rc;
    if (s->geo.R > 1)
        CU(cudaMemsetAsync(s->halo + s->hl.flags + (size_t)bank * s->hl.bank_words(), 0, s->hl.bank_words() * sizeof(int),
                           s->stream));
    CU(cudaEventRecord(s->ev[0], s->stream));
    if (s->Zl && !streamed) {
        Bp3Layout L = { s->cells, s->rows, s->W, s->H, s->Zl, s->P, s->RWP, s->d_pop };
        CU(launch_ca3d_pack(L, s->stream));
        /* halo seed: H rows of each block's first plane -> ghost plane above the previous block (generation 0's "old plane above") */
        for (size_t l = 0; l < s->h_planes.size(); l++) {
            const Bp3Plane &pl = s->h_planes[l];
            if (!pl.push_dn_rows) continue;
            CU(launch_halo_seed(pl.push_dn_rows, s->rows + l * (size_t)s->H * s->NP * s->RWP, s->H, s->RWP, s->NP, s->stream));
        }
    }
    if (streamed) {
        /* chunk-arrival word, per-plane done flags (host-mapped), copy streams */
        if (!s->stream_in) CU(cudaStreamCreateWithFlags(&s->stream_in, cudaStreamNonBlocking));
        if (!s->stream_out) CU(cudaStreamCreateWithFlags(&s->stream_out, cudaStreamNonBlocking));
        if (!s->d_in_ready) CU(cudaMalloc(&s->d_in_ready, 64));
        if (!s->ev_io) CU(cudaEventCreateWithFlags(&s->ev_io, cudaEventDisableTiming));
        const int Zl = s->Zl ? s->Zl : 1;
        s->io_chunk = io_chunk_planes((size_t)s->W * s->H, Zl);
        const int nchunks = (Zl + s->io_chunk - 1) / s->io_chunk;
        const size_t want = (size_t)Zl + nchunks;
        if (s->h_io_count < want) {
            if (s->h_io) cudaFreeHost(s->h_io);
            s->h_io = nullptr;
            s->h_io_count = 0;
            CU(cudaHostAlloc((void **)&s->h_io, want * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
            memset(s->h_io, 0, want * sizeof(int));
            s->h_io_count = want;
            s->io_epoch = 0;
        }
        for (int c = 0; c < nchunks; c++) s->h_io[Zl + c] = c + 1;
        s->io_epoch++;
        CU(cudaMemsetAsync(s->d_in_ready, 0, sizeof(int), s->stream));
        CU(cudaMemsetAsync(s->d_pop, 0, sizeof(unsigned long long), s->stream));
    }
    CU(cudaMemsetAsync(s->prog, 0, (size_t)(s->Gcap + 1) * (s->Zl ? s->Zl : 1) * sizeof(int), s->stream));
    CU(cudaMemsetAsync(s->ticket, 0, kTicketWords * sizeof(unsigned), s->stream));
    CU(cudaEventRecord(s->ev[1], s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->prepared = true;
    return CLAPCA_OK;
}

int clapca_slab_prepare(clapca_slab *s, uint32_t surv, uint32_t born, uint32_t nr_states, int steps)
{
    return slab_prepare(s, surv, born, nr_states, steps, false);
}

int clapca_slab_prepare_streamed(clapca_slab *s, uint32_t surv, uint32_t born, uint32_t nr_states, int steps)
{
    return slab_prepare(s, surv, born, nr_states, steps, true);
}

/* CLAPCA_DUMP=1: after a watchdog, the progress table and the ghost counters of this rank (what was everybody waiting for?) */
static void slab_post_mortem(clapca_slab *s)
{
    const char *e = getenv("CLAPCA_DUMP");
    if (!e || !atoi(e))
        return;
    const int Zl = s->Zl ? s->Zl : 1;
    std::vector<int> prog((size_t)(s->G + 1) * Zl);
    if (cudaMemcpy(prog.data(), s->prog, prog.size() * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
        return;
    fprintf(stderr, "clapca post-mortem rank %d: H %d, local planes %d, generations %d; prog[g][z] (g = -1 first):\n",
            s->geo.rank, s->H, s->Zl, s->G);
    for (int g = 0; g <= s->G; g++) {
        fprintf(stderr, "  g %2d:", g - 1);
        for (int z = 0; z < Zl; z++) fprintf(stderr, " %d", prog[(size_t)g * Zl + z]);
        fprintf(stderr, "\n");
    }
    if (s->geo.R > 1) {
        const int bank = (int)(s->epoch & 1u);
        std::vector<int> fl(s->hl.bank_words());
        if (cudaMemcpy(fl.data(), s->halo + s->hl.flags + (size_t)bank * s->hl.bank_words(), fl.size() * sizeof(int),
                       cudaMemcpyDeviceToHost) != cudaSuccess)
            return;
        for (int up = 0; up < 2; up++)
            for (int lb = 0; lb < s->hl.nlb_max; lb++) {
                fprintf(stderr, "  ghost %s of local block %d, counters g = -1 ..:", up ? "above" : "below", lb);
                for (int g = 0; g <= s->hl.Gcap; g++)
                    fprintf(stderr, " %d", fl[((size_t)up * s->hl.nlb_max + lb) * (s->hl.Gcap + 1) + g]);
                fprintf(stderr, "\n");
            }
    }
}

/* the sweep launch of a prepared slab; streamed: with layout items (pack / unpack inside the launch) */
static int slab_launch(clapca_slab *s, bool streamed, int *workers)
{
    *workers = 0;
    if (!s->n_items)
        return CLAPCA_OK;
    Bp3Params p;
    memset(&p, 0, sizeof(p));
    p.rows = s->rows;
    p.planes = s->planes;
    p.W = s->W; p.H = s->H; p.Z = s->Zl; p.G = s->G; p.RWP = s->RWP;
    p.prog = s->prog + (s->Zl ? s->Zl : 1);
    p.order = s->order;
    p.nsweeps = s->n_items;
    p.zeros = g_ctx.d_zeros;
    sweep_knobs(p, s->team);
    if (s->max_ctas > 0) p.max_ctas = s->max_ctas;
    p.ticket = s->ticket;
    p.err = (int *)(s->ticket + 1);
    p.diag = diag_enabled() ? (unsigned long long *)(s->ticket + 4) : nullptr;
    p.surv = s->surv; p.born = s->born; p.bornval = (s->nr_states - 1u) & 0xffu;
    p.spin_limit = 20000000000LL;       /* ~10 s: ranks enter the kernel at slightly different times */
    if (streamed) {
        int *h_io_dev = nullptr;
        CU(cudaHostGetDevicePointer((void **)&h_io_dev, s->h_io, 0));
        p.layout_items = 1;
        p.io_cells = s->cells;
        p.io_chunk = s->io_chunk;
        p.population = s->d_pop;
        p.in_ready = s->d_in_ready;
        p.out_done = h_io_dev;
        p.io_epoch = s->io_epoch;
    }
    Bp3LaunchInfo info;
    CU(bp3_launch(s->rule, s->P, s->WPL, p, g_ctx.sms, s->stream, &info));
    *workers = info.workers;
    return CLAPCA_OK;
}

/*
 * Step 2 (after the barrier): the fused sweep kernel -- halo rows travel as peer stores inside it --
 * then the layout conversion back and the local population count.
 */
int clapca_slab_run(clapca_slab *s, int64_t *local_population)
{
    if (int rc = need_init()) return rc;
    if (!s || !s->prepared || s->streamed_order) return fail(CLAPCA_ERR_STATE, "slab_run: slab is not prepared (for a resident run)");
    s->prepared = false;
    memset(&s->stats, 0, sizeof(s->stats));
    CU(cudaEventRecord(s->ev[2], s->stream));
    int workers = 0;
    if (int rc = slab_launch(s, false, &workers)) return rc;
    CU(cudaEventRecord(s->ev[3], s->stream));
    diag_report("slab", s->ticket, s->stream, s->geo.rank);
    CU(cudaMemsetAsync(s->d_pop, 0, sizeof(unsigned long long), s->stream));
    if (s->Zl) {
        Bp3Layout L = { s->cells, s->rows, s->W, s->H, s->Zl, s->P, s->RWP, s->d_pop };
        CU(launch_ca3d_unpack(L, s->stream));
    }
    CU(cudaEventRecord(s->ev[4], s->stream));
    unsigned long long pop = 0;
    int err = 0;
    CU(cudaMemcpyAsync(&pop, s->d_pop, sizeof(pop), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(&err, s->ticket + 1, sizeof(err), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (err) {
        slab_post_mortem(s);
        return fail(CLAPCA_ERR_TIMEOUT, "slab_run: dataflow watchdog fired on rank %d (err=%d)", s->geo.rank, err);
    }
    if (local_population) *local_population = (int64_t)pop;
    float prep = 0, sweep = 0, tail = 0;
    CU(cudaEventElapsedTime(&prep, s->ev[0], s->ev[1]));
    CU(cudaEventElapsedTime(&sweep, s->ev[2], s->ev[3]));
    CU(cudaEventElapsedTime(&tail, s->ev[3], s->ev[4]));
    s->stats.total_ms = prep + sweep + tail;
    s->stats.kernel_ms = sweep;
    s->stats.launches = (s->Zl ? 2 : 0) + (s->n_items ? 1 : 0);
    s->stats.engine = CLAPCA_ENGINE_BITPLANE;
    s->stats.planes = s->P;
    s->stats.workers = workers;
    return CLAPCA_OK;
}

/*
 * Step 2 of a streamed run (after the barrier): host -> device -> host as ONE pipeline per rank.  host_in / host_out
 * hold this rank's planes in local order (pinned memory); the launch's pack items follow the H2D chunks, every
 * generation follows a few planes behind, the unpack items flag finished planes and this thread releases their D2H
 * copies -- while the z-block edges trade halo rows with the neighbouring ranks as in a resident run.
 */
int clapca_slab_run_streamed(clapca_slab *s, const uint8_t *host_in, uint8_t *host_out, int64_t *local_population)
{
    if (int rc = need_init()) return rc;
    if (!s || !s->prepared || !s->streamed_order)
        return fail(CLAPCA_ERR_STATE, "slab_run_streamed: call clapca_slab_prepare_streamed first");
    if (s->Zl && (!host_in || !host_out)) return fail(CLAPCA_ERR_ARG, "slab_run_streamed: NULL buffer");
    s->prepared = false;
    memset(&s->stats, 0, sizeof(s->stats));
    CU(cudaEventRecord(s->ev_io, s->stream));
    CU(cudaStreamWaitEvent(s->stream_in, s->ev_io, 0));         /* the chunk-arrival word is cleared before the first chunk lands */
    CU(cudaEventRecord(s->ev[2], s->stream));
    int workers = 0;
    if (int rc = slab_launch(s, true, &workers)) return rc;
    CU(cudaEventRecord(s->ev[3], s->stream));
    bool stuck = false;
    if (s->Zl) {
        StreamPlan sp = { s->cells, host_in, host_out, (size_t)s->W * s->H, s->Zl, s->io_chunk, s->d_in_ready, s->h_io,
                          s->io_epoch, s->stream, s->stream_in, s->stream_out };
        if (int rc = stream_volume(sp, &stuck)) return rc;
    }
    unsigned long long pop = 0;
    int err = 0;
    CU(cudaMemcpyAsync(&pop, s->d_pop, sizeof(pop), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(&err, s->ticket + 1, sizeof(err), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    diag_report("slab (streamed)", s->ticket, s->stream, s->geo.rank);
    if (err == 4)
        return fail(CLAPCA_ERR_ARG, "slab_run_streamed: a cell value exceeds the max_value %u given to slab_create", s->max_value);
    if (err || stuck)
        return fail(CLAPCA_ERR_TIMEOUT, "slab_run_streamed: dataflow watchdog fired on rank %d (err=%d)", s->geo.rank, err);
    if (local_population) *local_population = (int64_t)pop;
    float sweep = 0;
    CU(cudaEventElapsedTime(&sweep, s->ev[2], s->ev[3]));
    s->stats.total_ms = s->stats.kernel_ms = sweep;
    s->stats.launches = s->n_items ? 1 : 0;
    s->stats.engine = CLAPCA_ENGINE_BITPLANE;
    s->stats.planes = s->P;
    s->stats.workers = workers;
    s->stats.streamed = 1;
    return CLAPCA_OK;
}

int clapca_slab_last_stats(clapca_slab *s, clapca_run_stats *st)
{
    if (!s || !st) return fail(CLAPCA_ERR_ARG, "slab_last_stats: NULL argument");
    *st = s->stats;
    return CLAPCA_OK;
}

#pragma GCC visibility pop
} /* extern "C" */
