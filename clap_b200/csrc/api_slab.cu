/*
 * api_slab.cu -- multi-GPU part of the C ABI (include/clapca.h, clapca_slab_*): z-block slabs of one ca3d
 * volume, one process per GPU, halo rows exchanged by peer stores inside the sweep kernel.
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
    int n_items = 0, order_G = -1, team = -1;
    unsigned *ticket = nullptr;
    unsigned long long *d_pop = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
    uint32_t surv = 0, born = 0, nr_states = 0;
    int G = 0, rule = BP3_RULE_DYN;
    bool prepared = false;
    uint32_t epoch = 0;             /* run number: upper half of the ghost-row tags */
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
    s->hl = slab_halo_layout(s->geo, s->H, s->RWP);
    s->stream = g_ctx.stream;
    memset(&s->stats, 0, sizeof(s->stats));
    const size_t zl = s->Zl ? s->Zl : 1;
    cudaError_t e = cudaMalloc(&s->cells, zl * s->W * s->H);
    if (e == cudaSuccess) e = cudaMalloc(&s->rows, zl * s->H * s->NP * s->RWP * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&s->halo, s->hl.total_words * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(s->halo, 0, s->hl.total_words * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&s->prog, (size_t)s->Gcap * zl * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&s->planes, zl * sizeof(Bp3Plane));
    if (e == cudaSuccess) e = cudaMalloc(&s->ticket, kTicketWords * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_pop, sizeof(unsigned long long));
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
    void *bufs[] = { s->cells, s->rows, s->halo, s->prog, s->planes, s->order, s->ticket, s->d_pop };
    for (void *b : bufs)
        if (b) cudaFree(b);
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
    SlabPtrs ptr = { s->rows, s->prog, s->halo, s->halo_next, s->halo_prev };
    bp3_build_planes(s->geo, ptr, s->hl, s->H, s->RWP, s->NP, s->h_planes);
    /* progress counters of a generation are Zl apart only when all Gcap generations share one table */
    if (s->Zl)
        CU(cudaMemcpy(s->planes, s->h_planes.data(), s->h_planes.size() * sizeof(Bp3Plane), cudaMemcpyHostToDevice));
    return CLAPCA_OK;
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
 * Step 1 of a sharded run (all ranks, then a barrier): lay the local planes out as bit-plane row
 * records, seed the neighbour's ghost planes with the H rows of every block's first plane (the
 * "old plane above" of generation 0) and clear the progress counters.
 */
int clapca_slab_prepare(clapca_slab *s, uint32_t surv, uint32_t born, uint32_t nr_states, int steps)
{
    if (int rc = need_init()) return rc;
    if (!s) return fail(CLAPCA_ERR_ARG, "slab_prepare: NULL slab");
    if (s->h_planes.empty() && s->Zl) return fail(CLAPCA_ERR_STATE, "slab_prepare: call clapca_slab_connect first");
    if (steps < 1 || steps > s->Gcap) return fail(CLAPCA_ERR_ARG, "slab_prepare: steps %d outside 1..%d", steps, s->Gcap);
    const uint32_t bornval = (nr_states - 1u) & 0xffu;
    if (born && (bornval >> s->P))
        return fail(CLAPCA_ERR_ARG, "slab_prepare: rule needs more than the %d state planes of this slab", s->P);
    s->surv = surv; s->born = born; s->nr_states = nr_states; s->G = steps;
    s->epoch = (s->epoch + 1) & 0xffffu;        /* every rank prepares the same number of times */
    if (!s->epoch) s->epoch = 1;
    s->rule = BP3_RULE_DYN;
    for (int i = 0; i < 9; i++)
        if (kCas[i][0] == surv && kCas[i][1] == born && kCas[i][2] == nr_states) { s->rule = i; break; }

    const int team = team_config(s->P, s->WPL);
    if (s->order_G != steps || s->team != team) {
        std::vector<WorkItem> items;
        const OrderCfg oc = order_config(s->geo.Zg, s->H, steps, bp3_max_workers(s->rule, s->P, s->WPL, g_ctx.sms), team);
        s->team = team;
        make_items(oc, s->h_planes, s->geo.Zg, s->H, steps, items, false);
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
    }
    CU(cudaEventRecord(s->ev[0], s->stream));
    if (s->Zl) {
        Bp3Layout L = { s->cells, s->rows, s->W, s->H, s->Zl, s->P, s->RWP, s->d_pop };
        CU(launch_ca3d_pack(L, s->stream));
        /* halo seed: H rows of each block's first plane -> ghost plane above the previous block, tag = seed state */
        for (size_t l = 0; l < s->h_planes.size(); l++) {
            const Bp3Plane &pl = s->h_planes[l];
            if (!pl.push_dn_rows) continue;
            CU(launch_halo_seed(pl.push_dn_rows, s->rows + l * (size_t)s->H * s->NP * s->RWP, s->H, s->RWP, s->NP, s->WPL,
                                s->epoch << 16, s->stream));
        }
    }
    CU(cudaMemsetAsync(s->prog, 0, (size_t)s->Gcap * (s->Zl ? s->Zl : 1) * sizeof(int), s->stream));
    CU(cudaMemsetAsync(s->ticket, 0, kTicketWords * sizeof(unsigned), s->stream));
    CU(cudaEventRecord(s->ev[1], s->stream));
    CU(cudaStreamSynchronize(s->stream));
    s->prepared = true;
    return CLAPCA_OK;
}

/*
 * Step 2 (after the barrier): the fused sweep kernel -- halo rows travel as peer stores inside it --
 * then the layout conversion back and the local population count.
 */
int clapca_slab_run(clapca_slab *s, int64_t *local_population)
{
    if (int rc = need_init()) return rc;
    if (!s || !s->prepared) return fail(CLAPCA_ERR_STATE, "slab_run: slab is not prepared");
    s->prepared = false;
    memset(&s->stats, 0, sizeof(s->stats));
    CU(cudaEventRecord(s->ev[2], s->stream));
    int workers = 0;
    if (s->n_items) {
        Bp3Params p;
        memset(&p, 0, sizeof(p));
        p.rows = s->rows;
        p.planes = s->planes;
        p.W = s->W; p.H = s->H; p.Z = s->Zl; p.G = s->G; p.RWP = s->RWP;
        p.prog = s->prog;
        p.order = s->order;
        p.nsweeps = s->n_items;
        p.epoch = s->epoch;
        sweep_knobs(p, s->team);
        p.ticket = s->ticket;
        p.err = (int *)(s->ticket + 1);
        p.diag = diag_enabled() ? (unsigned long long *)(s->ticket + 4) : nullptr;
        p.surv = s->surv; p.born = s->born; p.bornval = (s->nr_states - 1u) & 0xffu;
        p.spin_limit = 20000000000LL;       /* ~10 s: ranks enter the kernel at slightly different times */
        for (const Bp3Plane &pl : s->h_planes)
            if (pl.ghost_mask || pl.push_dn_rows || pl.push_up_rows) { p.edge_loop = 1; break; }
#if CLAPCA_EDGE_DEFER
        if (const char *e = getenv("CLAPCA_GHOST_DEFER")) if (p.edge_loop && atoi(e) != 0) p.edge_loop = 2;
#endif
        Bp3LaunchInfo info;
        CU(bp3_launch(s->rule, s->P, s->WPL, p, g_ctx.sms, s->stream, &info));
        workers = info.workers;
    }
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
    if (err)
        return fail(CLAPCA_ERR_TIMEOUT, "slab_run: dataflow watchdog fired on rank %d (err=%d)", s->geo.rank, err);
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

int clapca_slab_last_stats(clapca_slab *s, clapca_run_stats *st)
{
    if (!s || !st) return fail(CLAPCA_ERR_ARG, "slab_last_stats: NULL argument");
    *st = s->stats;
    return CLAPCA_OK;
}

#pragma GCC visibility pop
} /* extern "C" */
