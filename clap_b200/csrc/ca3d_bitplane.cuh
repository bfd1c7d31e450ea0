/*
 * ca3d_bitplane.cuh -- bit-plane engine for ca3d_run() (core/ca3d.c:124-142).
 *
 * The reference updates the volume IN PLACE in z/y/x order, so a cell sees
 * the new generation of the 13 neighbours that precede it in the sweep and the
 * old generation of the other 13.  This engine reproduces that order exactly,
 * for all generations at once:
 *
 * Layout.  The uint8 volume is re-laid as bit planes.  For every grid row
 * (y,z) there is one "row record" of NP = 2 + P plane-rows, each RWP 32-bit
 * words (bit i of word w = cell x = 32 w + i):
 *     [ H0 | H1 | S0 ... S(P-1) ]
 * S_p is bit p of the cell state (P = bits needed for the largest value that
 * can ever occur), H1:H0 is the horizontal alive 3-sum a(x-1)+a(x)+a(x+1) of
 * the row, published so consumers never recompute it.  Padding bits of S are
 * always 0.
 *
 * Row step.  For row (y,z) of generation g every input except the new alive
 * bit of cell x-1 is known before the row starts:
 *     K = V(z-1,new)[y] + V(z+1,old)[y] + H(y-1,new) + H(y+1,old) + a_old(x+1)
 * with V = H(y-1)+H(y)+H(y+1) of one plane.  f0/f1 = new alive bit for
 * n = K / K+1 come from the rule tables, and the in-row chain is resolved by
 * the GF(2) affine scan of bitslice.cuh.  One warp updates a whole row.
 *
 * Sweeps and dataflow.  A "sweep" (z,g) = rows y = 0..H-1 of plane z at
 * generation g, run by one persistent warp.  Row y of sweep (z,g) needs rows
 * <= y+1 of (z-1,g) [new plane below], of (z+1,g-1) [old plane above] and of
 * (z,g-1) [own old state]; each sweep publishes a progress counter
 * prog[g][z] = rows completed (store after __threadfence), consumers poll it.
 * Sweeps are claimed from a ticket counter in an order that extends the
 * dependency order (key 2z+4g), so with all warps co-resident (cooperative
 * launch) the earliest unfinished sweep can always advance: no deadlock.  All
 * generations of a plane follow each other a few rows apart, so the working
 * set lives in L2 and HBM sees one read and one write of the volume for the
 * whole run (temporal blocking over ALL generations).  The same in-place
 * argument as the reference's makes single buffering safe: a row is only
 * overwritten after every reader of its previous version has passed it.
 */
#ifndef CLAPCA_CA3D_BITPLANE_CUH
#define CLAPCA_CA3D_BITPLANE_CUH

#include "bitslice.cuh"
#include "bp3_types.h"

namespace clapca {

struct Bp3Params {
    uint32_t *rows;         /* local row records, [(z*H + y) * NP + plane] * RWP */
    const Bp3Plane *planes; /* [Z] */
    int W, H, Z, G;         /* cells per row, rows per plane, LOCAL planes, generations */
    int RWP;                /* words per plane-row = 32 * WPL */
    int *prog;              /* [G][Z] rows completed by local sweep (z,g) */
    const int4 *order;      /* work items in claim order: (local z, g, first row, end row) */
    int nsweeps;            /* number of work items */
    uint32_t epoch;         /* run number, upper half of the ghost-row tags */
    int flag_rows;          /* progress counters are raised every flag_rows rows (and at segment ends) */
    int prefetch_rows;      /* > 0: prefetch.L2 the own record this many rows ahead (HBM -> L2 latency) */
    unsigned *ticket;       /* next sweep to claim */
    int *err;               /* != 0: watchdog fired, everybody bails out */
    unsigned long long *diag;   /* optional: [0] cycles waiting on counters, [1] on ghost tags, [2] in work items */
    uint32_t surv, born;    /* rule masks (run-time rule only) */
    uint32_t bornval;       /* (nr_states - 1) & 0xff */
    long long spin_limit;   /* watchdog budget in clock ticks per wait */
    int max_ctas_per_sm;    /* host side only: > 0 caps the resident CTAs per SM of the launch */
    int pub_workers;        /* > 0: every CTA = pub_workers worker warps + ONE publisher warp (see PubSlot) */
    int team;               /* > 0: team mode -- a CTA of `team` warps sweeps `team` consecutive planes (see below) */
    int edge_flag_rows;     /* team mode: counter period of a team's LAST plane (it feeds the next team); 0 = flag_rows */
    int edge_loop;          /* != 0: some plane of this launch has a ghost source / feeds a peer (multi-GPU): see run_segment;
                               2 = the deferred-tag-check loop (only in builds with CLAPCA_EDGE_DEFER) */
    /*
     * Layout items (optional, see "Layout items" below): the conversion between the reference's uint8 cells and
     * the row records runs INSIDE the sweep launch, so a volume can stream host -> device -> host through it.
     */
    int layout_items;       /* != 0: items with g == -1 pack a plane, items with g == G unpack it; prog has a row -1 */
    uint8_t *io_cells;      /* reference-layout cells of the local planes, (z*H + y)*W + x */
    const int *in_ready;    /* chunks of io_chunk planes that have landed in io_cells (raised by the copy stream) */
    int *out_done;          /* [Z] = io_epoch once the plane's final cells are back in io_cells (host-mapped memory) */
    int io_chunk;           /* planes per H2D chunk */
    int io_epoch;           /* run number written into out_done */
    unsigned long long *population;     /* += non-zero cells of every unpacked plane */
};

/*
 * Layout items.  With layout_items set the claim order holds two more item kinds around the G sweep
 * generations of every plane: "generation -1" PACKS the plane (uint8 cells -> row record, publishing
 * prog[-1][z] so that sweep (z,0) and (z-1,0) wait for it exactly as they wait for a previous generation),
 * and "generation G" UNPACKS it (row record -> uint8 cells + population) as soon as sweep (z,G-1) is done.
 * A pack item first waits until the copy engine has delivered the plane's chunk (in_ready, raised by a
 * stream-ordered 4-byte copy after each chunk's H2D copy); an unpack item stores the run's epoch into
 * out_done[z] -- host-mapped memory -- after a system-scope fence, and the host thread inside the C-ABI call
 * issues a chunk's D2H copy as soon as all of its planes carry the epoch.  Upload, all
 * generations and download of a volume then overlap inside ONE launch: the sweep follows the H2D front a
 * few planes behind and the D2H follows the last generation.  Cells whose value does not fit the P state
 * planes of the launched variant raise err = 4 (the host picks P from the caller's bound and verifies here).
 */

/*
 * Team mode.  With one warp per sweep and gpu-scope counters, plane z+1 trails plane z by flag_rows + ~3 rows
 * (the counter period plus the MEMBAR.GPU / poll latency), and that hop -- paid Z times in a row -- bounds how
 * many sweeps the dependency DAG lets run at once: G * H / hop.  One GPU has fewer warps than that; eight
 * GPUs sharing one volume do not.  In team mode a work item is a GROUP of up to `team` consecutive planes of
 * one generation, swept by the warps of ONE CTA: warp w takes plane z0 + w and follows warp w - 1 through a
 * shared-memory row counter raised after EVERY row (CTA-scope release: MEMBAR.CTA + STS, tens of cycles), so
 * inside a team the hop is the 3 rows the data dependency itself asks for.  The row data still travels
 * through L2 (st.cg / ld.cg by warps of the same SM, ordered by the CTA-scope release/acquire pair).  Only
 * the team's last plane hands over to another CTA (or GPU) through the gpu-scope counter / ghost tags, and
 * every plane keeps raising its gpu-scope counter every flag_rows rows for the next generation's readers.
 */
enum { BP3_MAX_TEAM = 24 };

/*
 * -DCLAPCA_EDGE_DEFER=1 additionally builds a third instantiation of the row loop (edge_loop == 2) in which the
 * tags of a ghost row are checked one row step AFTER its loads were issued (see fetch_h_tagged); off by default:
 * not yet measured on hardware, and every instantiation costs build time.  The emulator tests build it.
 */
#ifndef CLAPCA_EDGE_DEFER
#define CLAPCA_EDGE_DEFER 0
#endif

/*
 * Publisher mode.  Raising a progress counter needs a gpu-scope release, and MEMBAR.GPU on a two-die B200
 * costs a few microseconds -- the time of two row steps.  A worker that fences itself can therefore only
 * afford a counter update every ~8 rows, which makes every consumer trail its producer by ~10 rows and
 * bounds the number of sweeps the dependency DAG lets run concurrently (the limit that matters once a
 * rank owns only G/N sweeps per dependency level).  In publisher mode the workers never fence at gpu
 * scope: after the stores of a row they bump `done` in a shared-memory mailbox (CTA-scope release: cheap),
 * and one extra warp per CTA loops { read all mailboxes; ONE fence.acq_rel.gpu; store the counters that
 * moved }.  Release cumulativity (worker stores -> CTA-scope release/acquire -> gpu-scope fence -> counter)
 * makes the rows visible before the counter, exactly like bar.sync + thread 0 fencing in a grid barrier.
 */
struct PubSlot {
    unsigned long long flag;    /* global counter of the sweep the worker is on (int *) */
    int done;                   /* rows completed (worker-written) */
    int pub;                    /* rows published (publisher-written; worker resets it between items) */
};
enum { BP3_MAX_PUB_WORKERS = 23 };

/* ---- rules ---------------------------------------------------------------- */

/* compile-time rule: the nine entries of cas[] (core/ca3d.c:110-122) */
template <uint32_t SURV, uint32_t BORN, uint32_t NR>
struct Rule3Const {
    static constexpr uint32_t kBornVal = (NR - 1u) & 0xffu;
    CA_MDEV uint32_t bornval(const Bp3Params &) { return kBornVal; }
    /* tables at n = K and n = K + 1 */
    CA_MDEV void eval(const Bp3Params &, const uint32_t k[5],
                            uint32_t &s0, uint32_t &s1, uint32_t &b0, uint32_t &b1)
    {
        s0 = bs_tab5<SURV>(k);
        s1 = bs_tab5<(SURV >> 1)>(k);
        if (kBornVal) {
            b0 = bs_tab5<BORN>(k);
            b1 = bs_tab5<(BORN >> 1)>(k);
        } else {
            b0 = b1 = 0u;
        }
    }
};

/* run-time rule (arbitrary masks through the C ABI) */
struct Rule3Dyn {
    CA_MDEV uint32_t bornval(const Bp3Params &p) { return p.bornval; }
    CA_MDEV void eval(const Bp3Params &p, const uint32_t k[5],
                            uint32_t &s0, uint32_t &s1, uint32_t &b0, uint32_t &b1)
    {
        s0 = bs_tab_dyn(p.surv, k, 5);
        s1 = bs_tab_dyn(p.surv >> 1, k, 5);
        if (p.bornval) {
            b0 = bs_tab_dyn(p.born, k, 5);
            b1 = bs_tab_dyn(p.born >> 1, k, 5);
        } else {
            b0 = b1 = 0u;
        }
    }
};

/* ---- vector access to a lane's WPL consecutive words ------------------------ */

template <int WPL> struct LaneVec;
template <> struct LaneVec<1> {
    CA_MDEV void ld(const uint32_t *p, uint32_t v[1]) { v[0] = dp_ld_cg(p); }
    CA_MDEV void st(uint32_t *p, const uint32_t v[1]) { dp_st_cg(p, v[0]); }
};
template <> struct LaneVec<2> {
    CA_MDEV void ld(const uint32_t *p, uint32_t v[2])
    {
        uint2 t = dp_ld_cg(reinterpret_cast<const uint2 *>(p));
        v[0] = t.x; v[1] = t.y;
    }
    CA_MDEV void st(uint32_t *p, const uint32_t v[2])
    {
        dp_st_cg(reinterpret_cast<uint2 *>(p), make_uint2(v[0], v[1]));
    }
};
template <> struct LaneVec<4> {
    CA_MDEV void ld(const uint32_t *p, uint32_t v[4])
    {
        uint4 t = dp_ld_cg(reinterpret_cast<const uint4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    CA_MDEV void st(uint32_t *p, const uint32_t v[4])
    {
        dp_st_cg(reinterpret_cast<uint4 *>(p), make_uint4(v[0], v[1], v[2], v[3]));
    }
};

/* horizontal 3-sum a(x-1)+a(x)+a(x+1) of a lane-distributed row -> 2 bit planes */
template <int WPL>
CA_DEV void bp_hsum(const uint32_t a[WPL], uint32_t h0[WPL], uint32_t h1[WPL])
{
    const int lane = dp_lane();
    uint32_t prev = dp_shfl_up(a[WPL - 1], 1);      /* last word of the lane on the left  */
    uint32_t next = dp_shfl_down(a[0], 1);          /* first word of the lane on the right */
    if (lane == 0)  prev = 0u;
    if (lane == 31) next = 0u;
#pragma unroll
    for (int j = 0; j < WPL; j++) {
        uint32_t left  = j ? a[j - 1] : prev;
        uint32_t right = (j + 1 < WPL) ? a[j + 1] : next;
        uint32_t l = dp_funnel_l(left, a[j], 1);    /* bit i = a(x-1) */
        uint32_t r = dp_funnel_r(a[j], right, 1);   /* bit i = a(x+1) */
        h0[j] = bs_xor3(l, a[j], r);
        h1[j] = bs_maj3(l, a[j], r);
    }
}

/* valid-cell mask of word w of a row of W cells */
CA_DEV uint32_t bp_valid_mask(int w, int W)
{
    int rem = W - 32 * w;
    return rem >= 32 ? ~0u : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}

/* ---- the persistent sweep kernel -------------------------------------------- */

/*
 * Register budget and instruction count decide the speed of this kernel (ncu, round 1: it is issue-bound,
 * not HBM-bound), so the row loop is written for a short instruction stream:
 *   - all sliding windows live in register arrays whose slot is (row - y0) % 3 and the loop body is
 *     instantiated three times (step<0>, step<1>, step<2>): no register rotation moves;
 *   - every source is walked with a running, lane-adjusted pointer (one 64-bit add per row, plane offsets
 *     are compile-time immediates because the record stride NP * 32 * WPL is a template constant);
 *   - the producers' counters are cached as ONE number (`have` = min over the producers); the fast path of
 *     a dependency check is one compare, the slow path polls with ld.acquire in lanes 0..2 and reduces
 *     with redux.min;
 *   - counters are published with st.release every flag_rows rows.
 */
template <int P, int WPL, class Rule>
struct Sweep3 {
    static constexpr int NP = P + 2;
    static constexpr int RWP = 32 * WPL;            /* words per plane-row */
    static constexpr int RECW = NP * RWP;           /* words per row record */
    static constexpr int GHW = 4 * RWP;             /* words per ghost row ({word, tag} pairs of H0|H1) */
    enum { SRC_NONE = 0, SRC_LOCAL = 1, SRC_GHOST = 2 };

    struct St {
        /* slot of row r = (r - y0 + 3) % 3 */
        uint32_t hd[3][2][WPL], hu[3][2][WPL];      /* H rows of the plane below (new) / above (old) */
        uint32_t so[3][P][WPL];                     /* own state rows y, y+1, y+2 */
        uint32_t ho[2][WPL];                        /* H of own old row y+1 */
        uint32_t hn[2][WPL];                        /* H of own new row y-1 */
        uint32_t vmask[WPL];
        const uint32_t *dn, *up;                    /* lane-adjusted, at the next row to load */
        uint32_t *rec;                              /* lane-adjusted own record of the current row */
        const uint32_t *pf;                         /* lanes < NP*WPL: one 128-byte line of the record prefetch_rows ahead */
        uint32_t *push_dn, *push_up;                /* lane-adjusted peer ghost rows of the current row */
        PubSlot *slot;                              /* publisher mode: this worker's mailbox (else nullptr) */
        const int *sdn;                             /* team mode: shared-memory row counter of the plane below (else nullptr) */
        int *sown;                                  /* team mode: own shared-memory row counter (nullptr: nobody follows) */
        int have_s;                                 /* rows the plane below has published through sdn */
        int flag_period;                            /* rows between two raises of the own gpu-scope counter */
        const int *flagp;                           /* the producer counter this lane polls (lanes 0..2) */
        int have;                                   /* min over the producers' published row counts */
        int next_raise;                             /* next row count at which the own counter is raised */
        int dn_mode, up_mode;
        uint32_t tag_dn, tag_up, tag_out;
        uint32_t bad_dn, bad_up;                    /* deferred loop: != 0 in some lane = the ghost row fetched a step ago was stale */
        long long waited_flag, waited_tag, waited_team;     /* diagnostics: cycles spent in the slow paths */
    };

    /* false = watchdog fired / abort requested */
    CA_MDEV bool wait_rows(const Bp3Params &p, St &st, int need)
    {
        if (st.sdn && st.have_s < need && !wait_team(p, st, need))
            return false;
        if (st.have >= need)
            return true;
        long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            int v = st.flagp ? dp_ld_acquire(st.flagp) : 0x7fffffff;
            st.have = dp_reduce_min(v);
            if (st.have >= need)
                break;
            if (spins == 0) t0 = dp_clock();
            dp_nanosleep(32);
            if ((spins & 127u) == 127u) {
                bool bad = dp_ld_flag(p.err) != 0 || (dp_clock() - t0) > p.spin_limit;
                if (!dp_all(!bad)) {
                    if (dp_lane() == 0)
                        dp_set_error(p.err, 1);
                    return false;
                }
            }
        }
        if (t0) st.waited_flag += dp_clock() - t0;
        dp_syncwarp();          /* the polling lanes acquired; the warp barrier extends it to every lane */
        return true;
    }

    /*
     * Team mode: rows of the plane below, swept by the previous warp of this CTA.  Every lane reads the same
     * shared-memory word (a broadcast), the vote keeps the warp converged, and every lane acquires for itself.
     */
    CA_MDEV bool wait_team(const Bp3Params &p, St &st, int need)
    {
        long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            const int v = dp_ld_volatile(st.sdn);
            if (dp_all(v >= need)) {
                st.have_s = dp_reduce_min(v);
                break;
            }
            if (spins == 0) t0 = dp_clock();
            if ((spins & 1023u) == 1023u) {
                bool bad = dp_ld_flag(p.err) != 0 || (dp_clock() - t0) > p.spin_limit;
                if (!dp_all(!bad)) {
                    if (dp_lane() == 0)
                        dp_set_error(p.err, 3);
                    return false;
                }
            }
            dp_team_pause();
        }
        if (t0) st.waited_team += dp_clock() - t0;
        dp_fence_cta();
        return true;
    }

    CA_MDEV void zero2(uint32_t h[2][WPL])
    {
#pragma unroll
        for (int j = 0; j < WPL; j++) h[0][j] = h[1][j] = 0u;
    }

    /*
     * Ghost rows (see bp3_types.h): {word, tag} pairs written by the neighbouring GPU.  Re-read until every
     * pair of this lane -- and of the whole warp -- carries the expected tag.  false = watchdog / abort.
     */
    CA_MDEV bool load_h_tagged(const Bp3Params &p, St &st, const uint32_t *src, uint32_t expect, uint32_t h[2][WPL])
    {
        long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < WPL; i++) {             /* pairs 2i, 2i+1 of this lane */
                uint4 v = dp_ld_cg(reinterpret_cast<const uint4 *>(src) + 32 * i);   /* vector i of every lane is contiguous */
                const int q0 = 2 * i, q1 = 2 * i + 1;   /* pair index -> (plane, word) = (q / WPL, q % WPL) */
                h[q0 / WPL][q0 % WPL] = v.x;
                h[q1 / WPL][q1 % WPL] = v.z;
                ok = ok && v.y == expect && v.w == expect;
            }
            if (dp_all(ok)) {
                if (t0) st.waited_tag += dp_clock() - t0;
                return true;
            }
            if (spins == 0) t0 = dp_clock();
            dp_nanosleep(100);
            if ((spins & 63u) == 63u) {
                bool bad = dp_ld_flag(p.err) != 0 || (dp_clock() - t0) > p.spin_limit;
                if (!dp_all(!bad)) {
                    if (dp_lane() == 0)
                        dp_set_error(p.err, 2);
                    return false;
                }
            }
        }
    }

    CA_MDEV void store_h_tagged(const Bp3Params &p, uint32_t *dst, uint32_t tag, const uint32_t h0[WPL], const uint32_t h1[WPL])
    {
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            const int q0 = 2 * i, q1 = 2 * i + 1;
            const uint32_t a = (q0 / WPL) ? h1[q0 % WPL] : h0[q0 % WPL];
            const uint32_t b = (q1 / WPL) ? h1[q1 % WPL] : h0[q1 % WPL];
            dp_st_cg(reinterpret_cast<uint4 *>(dst) + 32 * i, make_uint4(a, tag, b, tag));   /* one 512-byte burst per warp */
        }
    }

    /*
     * Deferred loop (EDGE == 2).  The ghost row fetched inside a row step is not needed before the NEXT step, so
     * the loads are only issued: the data words go straight into the window registers, the tags are folded into one
     * word per lane, and settle_side() looks at that word a whole row step later -- the L2 latency of a remotely
     * written line leaves the critical path of the edge plane (which in team mode paces its 15 team-mates).  The
     * re-read of a stale row is a cold, out-of-line function that returns the row by value: the hot loop only
     * grows by the fold, one vote and one branch per side (inlining the polling loop at every site made the loop
     * 16 % longer and the whole kernel 30 % slower).
     */
    struct GhostRow {
        uint32_t w[2 * WPL];
        int ok;
    };

    CA_MDEV void fetch_h_tagged(const uint32_t *src, uint32_t expect, uint32_t h[2][WPL], uint32_t &bad)
    {
        uint32_t acc = 0u;
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            uint4 v = dp_ld_cg(reinterpret_cast<const uint4 *>(src) + 32 * i);
            const int q0 = 2 * i, q1 = 2 * i + 1;
            h[q0 / WPL][q0 % WPL] = v.x;
            h[q1 / WPL][q1 % WPL] = v.z;
            acc |= (v.y ^ expect) | (v.w ^ expect);
        }
        bad = acc;
    }

    CA_MCOLD GhostRow repoll_ghost(const uint32_t *src, uint32_t expect, int *err, long long spin_limit)
    {
        GhostRow r;
        long long t0 = dp_clock();
        for (unsigned spins = 0;; spins++) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < WPL; i++) {
                uint4 v = dp_ld_cg(reinterpret_cast<const uint4 *>(src) + 32 * i);
                r.w[2 * i] = v.x;
                r.w[2 * i + 1] = v.z;
                ok = ok && v.y == expect && v.w == expect;
            }
            if (dp_all(ok)) {
                r.ok = 1;
                return r;
            }
            dp_nanosleep(100);
            if ((spins & 63u) == 63u) {
                bool bad = dp_ld_flag(err) != 0 || (dp_clock() - t0) > spin_limit;
                if (!dp_all(!bad)) {
                    if (dp_lane() == 0)
                        dp_set_error(err, 2);
                    r.ok = 0;
                    return r;
                }
            }
        }
    }

    /* the row fetched by fetch_h_tagged() one step ago is about to be used */
    CA_MDEV bool settle_side(const Bp3Params &p, const uint32_t *next, uint32_t tag, uint32_t &bad, uint32_t h[2][WPL])
    {
        if (dp_all(bad == 0u))
            return true;
        bad = 0u;
        const GhostRow r = repoll_ghost(next - GHW, tag, p.err, p.spin_limit);
#pragma unroll
        for (int q = 0; q < 2 * WPL; q++)
            h[q / WPL][q % WPL] = r.w[q];
        return r.ok != 0;
    }

    /* next H row of the plane below / above into h; advances the running pointer.  false = aborted */
    template <int EDGE>
    CA_MDEV bool load_side(const Bp3Params &p, St &st, const uint32_t *&src, int mode, uint32_t tag, uint32_t h[2][WPL])
    {
        if (mode == SRC_LOCAL) {
            LaneVec<WPL>::ld(src, h[0]);
            LaneVec<WPL>::ld(src + RWP, h[1]);
            src += RECW;
        } else if (EDGE == 0 || mode == SRC_NONE) {
            zero2(h);
        } else {
            if (!load_h_tagged(p, st, src, tag, h)) return false;
            src += GHW;
        }
        return true;
    }

    /* own record at st.rec + D rows: state planes, and optionally the H planes */
    template <int D>
    CA_MDEV void load_own_s(const St &st, uint32_t s[P][WPL])
    {
#pragma unroll
        for (int q = 0; q < P; q++)
            LaneVec<WPL>::ld(st.rec + D * RECW + (2 + q) * RWP, s[q]);
    }
    template <int D>
    CA_MDEV void load_own_h(const St &st, uint32_t h[2][WPL])
    {
        LaneVec<WPL>::ld(st.rec + D * RECW, h[0]);
        LaneVec<WPL>::ld(st.rec + D * RECW + RWP, h[1]);
    }
    CA_MDEV void zero_s(uint32_t s[P][WPL])
    {
#pragma unroll
        for (int q = 0; q < P; q++)
#pragma unroll
            for (int j = 0; j < WPL; j++) s[q][j] = 0u;
    }

    /*
     * One row step.  M = (y - y0) % 3 selects the register slots: rows y-1 / y / y+1 of the side planes
     * sit in slots (M+2)%3 / M / (M+1)%3, own state rows y / y+1 in slots M / (M+1)%3; row y+2 is
     * prefetched into slot (M+2)%3 once row y-1 has been consumed.
     */
    template <int M, int EDGE>
    CA_MDEV bool step(const Bp3Params &p, St &st, int y, int y1, int *myprog)
    {
        constexpr int A = (M + 2) % 3, B = M, C = (M + 1) % 3;
        const int lane = dp_lane();
        const int H = p.H;
        const uint32_t bornval = Rule::bornval(p);
        uint32_t k[WPL][5], ao[WPL], ge2[WPL];

        if (EDGE == 2) {        /* ghost rows y+1 were only fetched during the previous step */
            if (st.dn_mode == SRC_GHOST && !settle_side(p, st.dn, st.tag_dn, st.bad_dn, st.hd[C])) return false;
            if (st.up_mode == SRC_GHOST && !settle_side(p, st.up, st.tag_up, st.bad_up, st.hu[C])) return false;
        }


        /* ---- neighbour count K (everything but the in-row predecessor) ---- */
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            uint32_t a = st.so[B][0][j], hi = 0u;
#pragma unroll
            for (int q = 1; q < P; q++) hi |= st.so[B][q][j];
            ao[j] = a | hi;
            ge2[j] = hi;
        }
        uint32_t nxt = dp_shfl_down(ao[0], 1);
        if (lane == 31) nxt = 0u;
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            uint32_t a_[2] = { st.hd[A][0][j], st.hd[A][1][j] }, b_[2] = { st.hd[B][0][j], st.hd[B][1][j] };
            uint32_t c_[2] = { st.hd[C][0][j], st.hd[C][1][j] };
            uint32_t vd[4], vu[4];
            bs_add3x2(a_, b_, c_, vd);
            uint32_t e_[2] = { st.hu[A][0][j], st.hu[A][1][j] }, f_[2] = { st.hu[B][0][j], st.hu[B][1][j] };
            uint32_t g_[2] = { st.hu[C][0][j], st.hu[C][1][j] };
            bs_add3x2(e_, f_, g_, vu);
            uint32_t n_[2] = { st.hn[0][j], st.hn[1][j] }, o_[2] = { st.ho[0][j], st.ho[1][j] };
            uint32_t right = (j + 1 < WPL) ? ao[j + 1] : nxt;
            uint32_t r = dp_funnel_r(ao[j], right, 1);      /* old alive bit of x+1 */
            bs_count3d(vd, vu, n_, o_, r, k[j]);
        }

        /* ---- prefetch row y+2 while the rule / scan / update below run ---- */
        if (st.pf) {
            if (y + p.prefetch_rows < H)
                dp_prefetch_l2(st.pf);
            st.pf += RECW;
        }
        if (y + 2 <= y1) {          /* row y1 is still needed (as "row y+1" of the last step), y1+1 is not */
            if (y + 2 < H) {
                if (!wait_rows(p, st, y + 3 < H ? y + 3 : H))
                    return false;
                load_own_h<2>(st, st.ho);
                load_own_s<2>(st, st.so[A]);
                if (EDGE == 2 && st.dn_mode == SRC_GHOST) {
                    fetch_h_tagged(st.dn, st.tag_dn, st.hd[A], st.bad_dn);
                    st.dn += GHW;
                } else if (!load_side<EDGE>(p, st, st.dn, st.dn_mode, st.tag_dn, st.hd[A])) {
                    return false;
                }
                if (EDGE == 2 && st.up_mode == SRC_GHOST) {
                    fetch_h_tagged(st.up, st.tag_up, st.hu[A], st.bad_up);
                    st.up += GHW;
                } else if (!load_side<EDGE>(p, st, st.up, st.up_mode, st.tag_up, st.hu[A])) {
                    return false;
                }
            } else {
                zero2(st.ho); zero2(st.hd[A]); zero2(st.hu[A]);
                zero_s(st.so[A]);
            }
        }

        /* ---- rule tables, in-row scan ---- */
        uint32_t s0[WPL], s1[WPL], b0[WPL], b1[WPL], D[WPL], Cc[WPL];
        uint32_t dl = 1u, cl = 0u;          /* lane map: identity so far */
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            Rule::eval(p, k[j], s0[j], s1[j], b0[j], b1[j]);
            /* new alive bit if the predecessor's new alive bit is 0 / 1 */
            uint32_t f0 = bs_mux(ao[j], s0[j] | ge2[j], b0[j]) & st.vmask[j];
            uint32_t f1 = bs_mux(ao[j], s1[j] | ge2[j], b1[j]) & st.vmask[j];
            D[j] = f0 ^ f1;
            Cc[j] = f0;
            bs_scan_word(D[j], Cc[j]);
            /* compose into the lane map: carry-out = c ^ (d & carry-in) */
            uint32_t d = D[j] >> 31, c = Cc[j] >> 31;
            cl = c ^ (d & cl);
            dl = d & dl;
        }
        uint32_t cin = bs_scan_warp(dl, cl, 0u);

        /* ---- apply: new alive bits, state planes ---- */
        uint32_t an[WPL];
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            uint32_t cm = 0u - cin;
            an[j] = Cc[j] ^ (D[j] & cm);
            uint32_t pred = (an[j] << 1) | cin;              /* new alive bit of x-1 */
            cin = an[j] >> 31;
            uint32_t sv = bs_mux(pred, s1[j], s0[j]);
            uint32_t bn = bs_mux(pred, b1[j], b0[j]);
            uint32_t dec = ao[j] & ~sv;                      /* alive, not surviving: state - 1 */
            uint32_t brn = ~ao[j] & bn & st.vmask[j];        /* dead, born: state = nr_states - 1 */
            uint32_t borrow = dec;
#pragma unroll
            for (int q = 0; q < P; q++) {
                uint32_t t = st.so[B][q][j];
                st.so[B][q][j] = t ^ borrow;
                borrow &= ~t;
                if ((bornval >> q) & 1u) st.so[B][q][j] |= brn;
            }
        }
        bp_hsum<WPL>(an, st.hn[0], st.hn[1]);

        /* ---- publish row y ---- */
        LaneVec<WPL>::st(st.rec, st.hn[0]);
        LaneVec<WPL>::st(st.rec + RWP, st.hn[1]);
#pragma unroll
        for (int q = 0; q < P; q++)
            LaneVec<WPL>::st(st.rec + (2 + q) * RWP, st.so[B][q]);
        st.rec += RECW;
        /*
         * Local consumers: warp barrier (orders every lane's row stores before lane 0), then ONE
         * release store of the counter (MEMBAR.GPU + store).  Raised every flag_rows rows.
         */
        const bool at_mark = (y + 1 == st.next_raise);
        if (at_mark)
            st.next_raise += st.flag_period;
        if (at_mark || y + 1 == y1 || st.sown) {
            dp_syncwarp();
            if (lane == 0) {
                if (st.sown) {                      /* team mode: the next warp of this CTA follows row by row */
                    dp_fence_cta();
                    dp_st_volatile(st.sown, y + 1);
                }
                if (at_mark || y + 1 == y1) {
                    if (st.slot) {
                        dp_fence_cta();
                        dp_st_volatile(&st.slot->done, y + 1);
                    } else {
                        dp_st_release(myprog, y + 1);
                    }
                }
            }
        }
        /*
         * A z-block's edge plane also feeds the neighbouring GPU's ghost plane: tagged peer stores over
         * NVLink, fire and forget -- no fence, no counter (the tag travels with every word).  They are issued AFTER the
         * counter's release so that its MEMBAR never waits for this row's NVLink round trip.
         */
        if (EDGE != 0 && st.push_dn) {
            store_h_tagged(p, st.push_dn, st.tag_out, st.hn[0], st.hn[1]);
            st.push_dn += GHW;
        }
        if (EDGE != 0 && st.push_up) {
            store_h_tagged(p, st.push_up, st.tag_out, st.hn[0], st.hn[1]);
            st.push_up += GHW;
        }
        return true;
    }

    /* one work item: rows [y0, y1) of plane z at generation g.  false = aborted */
    /*
     * The row loop exists twice.  A launch in which some plane has a ghost source or feeds a peer (multi-GPU)
     * runs the EDGE instantiation for EVERY plane; a single-GPU launch runs a loop without a single ghost
     * instruction in it.  The kernel is bound by instruction issue and fetch: measured on B200 at 2048^3 x 50,
     * the plain loop takes 119.3 ms where the loop carrying the (never executed) ghost paths takes 121-122 ms,
     * more inlined tag-polling code in the same loop cost 30 %, and mixing the two instantiations inside one CTA
     * (edge planes EDGE, their team-mates plain) was the slowest of all -- one hot loop per launch it is.
     */
    CA_MDEV bool run_segment(const Bp3Params &p, int z, int g, int y0, int y1, PubSlot *slot,
                             const int *sdn = nullptr, int *sown = nullptr, bool team_edge = false)
    {
        if constexpr (CLAPCA_EDGE_DEFER != 0) {
            if (p.edge_loop == 2)
                return run_rows<2>(p, z, g, y0, y1, slot, sdn, sown, team_edge);
        }
        if (p.edge_loop)
            return run_rows<1>(p, z, g, y0, y1, slot, sdn, sown, team_edge);
        return run_rows<0>(p, z, g, y0, y1, slot, sdn, sown, team_edge);
    }

    template <int EDGE>
    CA_MDEV bool run_rows(const Bp3Params &p, int z, int g, int y0, int y1, PubSlot *slot,
                          const int *sdn, int *sown, bool team_edge)
    {
        const int lane = dp_lane();
        const int H = p.H, Z = p.Z;
        const Bp3Plane pl = p.planes[z];
        int *myprog = p.prog + (size_t)g * Z + z;
        St st;

        /* ---- sources ---- */
        const int first = y0 > 0 ? y0 - 1 : 0;      /* first row loaded from the side planes */
        st.dn_mode = !pl.dn_rows ? SRC_NONE : ((pl.ghost_mask & 1u) ? SRC_GHOST : SRC_LOCAL);
        st.up_mode = !pl.up_rows ? SRC_NONE : ((pl.ghost_mask & 2u) ? SRC_GHOST : SRC_LOCAL);
        st.dn = pl.dn_rows ? pl.dn_rows + (size_t)first * pl.dn_stride + lane * (st.dn_mode == SRC_GHOST ? 4 : WPL)
                           : nullptr;
        st.up = pl.up_rows ? pl.up_rows + (size_t)first * pl.up_stride + lane * (st.up_mode == SRC_GHOST ? 4 : WPL)
                           : nullptr;
        st.rec = p.rows + ((size_t)z * H + y0) * RECW + lane * WPL;
        st.pf = (p.prefetch_rows > 0 && lane < NP * WPL)
              ? p.rows + ((size_t)z * H + y0 + p.prefetch_rows) * RECW + lane * 32 : nullptr;
        st.push_dn = pl.push_dn_rows ? pl.push_dn_rows + (size_t)y0 * GHW + lane * 4 : nullptr;
        st.push_up = pl.push_up_rows ? pl.push_up_rows + (size_t)y0 * GHW + lane * 4 : nullptr;
        st.tag_dn = (p.epoch << 16) | (uint32_t)(g + 1);    /* plane below: already generation g */
        st.tag_up = (p.epoch << 16) | (uint32_t)g;          /* plane above: still generation g-1 */
        st.tag_out = (p.epoch << 16) | (uint32_t)(g + 1);

        /* ---- producers (ghost sources are synchronised by their row tags, not by counters) ---- */
        {
            /* team mode: the plane below is followed through shared memory, not through its gpu-scope counter */
            const int *fdn = (st.dn_mode == SRC_LOCAL && !sdn) ? pl.dn_flag + (size_t)g * pl.dn_gstride : nullptr;
            /* layout items: "generation -1" is the pack item of a plane, counted in row -1 of the table */
            const bool prev = g > 0 || p.layout_items;
            const int *fup = (prev && st.up_mode == SRC_LOCAL) ? pl.up_flag + ((long long)g - 1) * (long long)pl.up_gstride : nullptr;
            const int *fown = prev ? p.prog + ((long long)g - 1) * Z + z : nullptr;
            st.flagp = lane == 0 ? fdn : (lane == 1 ? fup : (lane == 2 ? fown : nullptr));
            st.have = (fdn || fup || fown) ? 0 : 0x7fffffff;
        }
        st.sdn = sdn;
        st.sown = sown;
        st.have_s = 0;
        st.flag_period = (team_edge && p.edge_flag_rows > 0) ? p.edge_flag_rows : p.flag_rows;
        st.next_raise = (y0 / st.flag_period + 1) * st.flag_period;
        st.waited_flag = st.waited_tag = st.waited_team = 0;
        st.bad_dn = st.bad_up = 0u;
        const long long t_item = dp_clock();
#pragma unroll
        for (int j = 0; j < WPL; j++) st.vmask[j] = bp_valid_mask(lane * WPL + j, p.W);

        /* rows < y0 of this sweep belong to earlier segments (skewed-segment order only) */
        if (y0 > 0) {
            long long t0 = dp_clock();
            for (unsigned spins = 1;; spins++) {
                int v = dp_reduce_min(lane == 0 ? dp_ld_acquire(myprog) : 0x7fffffff);
                if (v >= y0)
                    break;
                dp_nanosleep(32);
                if ((spins & 127u) == 0u) {
                    bool bad = dp_ld_flag(p.err) != 0 || (dp_clock() - t0) > p.spin_limit;
                    if (!dp_all(!bad)) {
                        if (lane == 0)
                            dp_set_error(p.err, 1);
                        return false;
                    }
                }
            }
            dp_syncwarp();
        }
        /*
         * Publisher mode: hand the mailbox over to this sweep.  The previous item ended with pub == done, so
         * the publisher is idle on this slot; it reads pub, then done, then flag -- written here in the
         * opposite order -- so it can never pair an old row count with the new counter address.
         */
        st.slot = slot;
        if (slot && lane == 0) {
            dp_st_volatile64(&slot->flag, (unsigned long long)(size_t)myprog);
            dp_fence_cta();
            dp_st_volatile(&slot->done, y0);
            dp_fence_cta();
            dp_st_volatile(&slot->pub, y0);
        }
        if (!wait_rows(p, st, y0 + 2 < H ? y0 + 2 : H))
            return false;

        /* ---- fill the windows: rows y0-1 (slot 2), y0 (slot 0), y0+1 (slot 1) ---- */
        if (y0 > 0) {
            if (!load_side<EDGE>(p, st, st.dn, st.dn_mode, st.tag_dn, st.hd[2])) return false;
            if (!load_side<EDGE>(p, st, st.up, st.up_mode, st.tag_up, st.hu[2])) return false;
            load_own_h<-1>(st, st.hn);              /* row y0-1 of this plane is already generation g */
        } else {
            zero2(st.hd[2]); zero2(st.hu[2]); zero2(st.hn);
        }
        if (!load_side<EDGE>(p, st, st.dn, st.dn_mode, st.tag_dn, st.hd[0])) return false;
        if (!load_side<EDGE>(p, st, st.up, st.up_mode, st.tag_up, st.hu[0])) return false;
        load_own_s<0>(st, st.so[0]);
        if (y0 + 1 < H) {
            if (!load_side<EDGE>(p, st, st.dn, st.dn_mode, st.tag_dn, st.hd[1])) return false;
            if (!load_side<EDGE>(p, st, st.up, st.up_mode, st.tag_up, st.hu[1])) return false;
            load_own_h<1>(st, st.ho);
            load_own_s<1>(st, st.so[1]);
        } else {
            zero2(st.hd[1]); zero2(st.hu[1]); zero2(st.ho);
            zero_s(st.so[1]);
        }
        zero_s(st.so[2]);

        int y = y0;
        for (; y + 3 <= y1; y += 3) {
            if (!step<0, EDGE>(p, st, y, y1, myprog)) return false;
            if (!step<1, EDGE>(p, st, y + 1, y1, myprog)) return false;
            if (!step<2, EDGE>(p, st, y + 2, y1, myprog)) return false;
        }
        if (y < y1) {
            if (!step<0, EDGE>(p, st, y, y1, myprog)) return false;
            y++;
        }
        if (y < y1) {
            if (!step<1, EDGE>(p, st, y, y1, myprog)) return false;
        }
        /* publisher mode: the mailbox is reused by the next item only once the last rows are out */
        if (slot) {
            if (lane == 0)
                while (dp_ld_volatile(&slot->pub) < y1)
                    dp_nanosleep(20);
            dp_syncwarp();
        }
        if (p.diag && lane == 0) {
            dp_atomic_add64(p.diag + 0, (unsigned long long)st.waited_flag);
            dp_atomic_add64(p.diag + 1, (unsigned long long)st.waited_tag);
            dp_atomic_add64(p.diag + 2, (unsigned long long)(dp_clock() - t_item));
            dp_atomic_add64(p.diag + 3, (unsigned long long)st.waited_team);
        }
        return true;
    }

    /* ---- layout items: one warp converts one plane -------------------------------------------------------- */

    /* poll *flag (written by the copy stream or by other SMs) until it reaches `need`; false = watchdog / abort */
    CA_MDEV bool wait_word(const Bp3Params &p, const int *flag, int need, bool sys, int code)
    {
        const int lane = dp_lane();
        long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            int v = 0x7fffffff;
            if (lane == 0)
                v = sys ? dp_ld_flag_sys(flag) : dp_ld_acquire(flag);
            if (dp_reduce_min(v) >= need)
                break;
            if (spins == 0) t0 = dp_clock();
            dp_nanosleep(200);
            if ((spins & 127u) == 127u) {
                /* the copy engine may take its time: 16 x the budget of a dataflow wait */
                bool bad = dp_ld_flag(p.err) != 0 || (dp_clock() - t0) > 16 * p.spin_limit;
                if (!dp_all(!bad)) {
                    if (lane == 0)
                        dp_set_error(p.err, code);
                    return false;
                }
            }
        }
        if (sys)
            dp_fence_sys();
        dp_syncwarp();
        return true;
    }

    /* uint8 cells of plane z -> row records; publishes prog[-1][z] */
    CA_MDEV bool pack_plane(const Bp3Params &p, int z)
    {
        const int lane = dp_lane();
        const int H = p.H, W = p.W;
        if (p.in_ready && !wait_word(p, p.in_ready, z / p.io_chunk + 1, true, 5))
            return false;
        int *myprog = p.prog - p.Z + z;
        const uint8_t *src = p.io_cells + (size_t)z * H * W;
        uint32_t *rec = p.rows + (size_t)z * H * RECW + lane * WPL;
        const bool vec = (W & 15) == 0;
        constexpr uint32_t kOver = (P >= 8) ? 0u : (((0xffu << P) & 0xffu) * 0x01010101u);
        uint32_t over = 0u;
        for (int y = 0; y < H; y++, src += W, rec += RECW) {
            uint32_t s[P][WPL], a[WPL], h0[WPL], h1[WPL];
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                const int x0 = 32 * (lane * WPL + j);
#pragma unroll
                for (int q = 0; q < P; q++) s[q][j] = 0u;
                if (vec && x0 + 32 <= W) {
                    const uint4 lo = dp_ld_cg(reinterpret_cast<const uint4 *>(src + x0));
                    const uint4 hi = dp_ld_cg(reinterpret_cast<const uint4 *>(src + x0 + 16));
                    const uint32_t r[8] = { lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w };
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        over |= r[k];
#pragma unroll
                        for (int q = 0; q < P; q++)
                            s[q][j] |= ((((r[k] >> q) & 0x01010101u) * 0x01020408u) >> 24) << (4 * k);
                    }
                } else {
                    for (int i = 0; i < 32 && x0 + i < W; i++) {
                        const uint32_t v = dp_ld_cg(src + x0 + i);
                        over |= v;
#pragma unroll
                        for (int q = 0; q < P; q++) s[q][j] |= ((v >> q) & 1u) << i;
                    }
                }
                a[j] = 0u;
#pragma unroll
                for (int q = 0; q < P; q++) a[j] |= s[q][j];
            }
            bp_hsum<WPL>(a, h0, h1);
            LaneVec<WPL>::st(rec, h0);
            LaneVec<WPL>::st(rec + RWP, h1);
#pragma unroll
            for (int q = 0; q < P; q++)
                LaneVec<WPL>::st(rec + (2 + q) * RWP, s[q]);
            if (((y + 1) & 31) == 0 || y + 1 == H) {
                dp_syncwarp();
                if (lane == 0)
                    dp_st_release(myprog, y + 1);
            }
        }
        if (!dp_all((over & kOver) == 0u)) {
            if (lane == 0)
                dp_set_error(p.err, 4);
            return false;
        }
        return true;
    }

    /* row records of plane z after the last generation -> uint8 cells, population, out_done[z] */
    CA_MDEV bool unpack_plane(const Bp3Params &p, int z)
    {
        const int lane = dp_lane();
        const int H = p.H, W = p.W;
        if (p.G > 0 && !wait_word(p, p.prog + (size_t)(p.G - 1) * p.Z + z, H, false, 6))
            return false;
        if (p.G <= 0 && !wait_word(p, p.prog - p.Z + z, H, false, 6))
            return false;
        uint8_t *dst = p.io_cells + (size_t)z * H * W;
        const uint32_t *rec = p.rows + (size_t)z * H * RECW + lane * WPL;
        const bool vec = (W & 15) == 0;
        unsigned pop = 0u;
        for (int y = 0; y < H; y++, dst += W, rec += RECW) {
            uint32_t s[P][WPL];
#pragma unroll
            for (int q = 0; q < P; q++)
                LaneVec<WPL>::ld(rec + (2 + q) * RWP, s[q]);
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                const int x0 = 32 * (lane * WPL + j);
                uint32_t alive = 0u;
#pragma unroll
                for (int q = 0; q < P; q++) alive |= s[q][j];
                pop += (unsigned)dp_popc(alive);
                if (vec && x0 + 32 <= W) {
                    uint32_t r[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        uint32_t v = 0u;
#pragma unroll
                        for (int q = 0; q < P; q++)
                            v |= ((((s[q][j] >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u) << q;
                        r[k] = v;
                    }
                    dp_st_cg(reinterpret_cast<uint4 *>(dst + x0), make_uint4(r[0], r[1], r[2], r[3]));
                    dp_st_cg(reinterpret_cast<uint4 *>(dst + x0 + 16), make_uint4(r[4], r[5], r[6], r[7]));
                } else {
                    for (int i = 0; i < 32 && x0 + i < W; i++) {
                        uint32_t v = 0u;
#pragma unroll
                        for (int q = 0; q < P; q++) v |= ((s[q][j] >> i) & 1u) << q;
                        dst[x0 + i] = (uint8_t)v;
                    }
                }
            }
        }
        for (int o = 16; o; o >>= 1)
            pop += dp_shfl_down(pop, o);
        dp_syncwarp();
        if (lane == 0) {
            if (p.population)
                dp_atomic_add64(p.population, (unsigned long long)pop);
            if (p.out_done) {
                dp_fence_sys();             /* the cells must be visible to the copy engine before the word moves */
                dp_st_flag_sys(p.out_done + z, p.io_epoch);
            }
        }
        return true;
    }

    /* dispatch on the item kind: g == -1 pack, g == G unpack (layout items), else a sweep segment */
    CA_MDEV bool run_item(const Bp3Params &p, int z, int g, int y0, int y1, PubSlot *slot,
                          const int *sdn = nullptr, int *sown = nullptr, bool team_edge = false)
    {
        if (p.layout_items) {
            if (g < 0)
                return pack_plane(p, z);
            if (g >= p.G)
                return unpack_plane(p, z);
        }
        return run_segment(p, z, g, y0, y1, slot, sdn, sown, team_edge);
    }

    /* the worker loop: claim items in dependency order until the list is exhausted */
    CA_MDEV void work_loop(const Bp3Params &p, PubSlot *slot)
    {
        const int lane = dp_lane();
        for (;;) {
            unsigned t = 0;
            if (lane == 0) {
                t = dp_atomic_inc(p.ticket);
                if (dp_ld_flag(p.err) != 0)
                    t = 0xffffffffu;
            }
            t = dp_shfl(t, 0);
            if (t >= (unsigned)p.nsweeps)
                break;
            int4 it = p.order[t];
            if (!run_item(p, it.x, it.y, it.z, it.w, slot))
                break;
        }
    }

    /* publisher warp: lane i serves the mailbox of worker warp i; one gpu-scope fence per pass for all of them */
    CA_MDEV void publish_loop(PubSlot *slots, int *nexit, int nw)
    {
        const int lane = dp_lane();
        PubSlot *my = slots + (lane < nw ? lane : 0);
        for (;;) {
            int pb = 0, d = 0;
            if (lane < nw) {
                pb = dp_ld_volatile(&my->pub);
                dp_fence_cta();
                d = dp_ld_volatile(&my->done);
            }
            const bool need = d > pb;
            if (dp_any(need)) {
                unsigned long long f = 0;
                if (need) {
                    dp_fence_cta();
                    f = dp_ld_volatile64(&my->flag);
                }
                dp_fence_release();                     /* fence.acq_rel.gpu: the ONE expensive instruction */
                if (need) {
                    dp_st_flag((int *)(size_t)f, d);
                    dp_st_volatile(&my->pub, d);
                }
            } else {
                if (dp_all(dp_ld_volatile(nexit) >= nw))        /* warp-uniform: the lanes read at different times */
                    break;
                dp_nanosleep(40);
            }
        }
    }

    /*
     * Team mode: the CTA claims one group of planes at a time (item = first local plane, generation, number of
     * planes); warp w sweeps plane w of the group, warps beyond the group's size sit the item out.
     */
    CA_MDEV void team_loop(const Bp3Params &p)
    {
        CA_SHARED(int, sm, BP3_MAX_TEAM + 1);       /* row counters of the team's warps, then the claimed ticket */
        const int w = dp_warp_in_block();
        for (;;) {
            if (dp_thread() == 0) {
                unsigned t = dp_atomic_inc(p.ticket);
                if (dp_ld_flag(p.err) != 0)
                    t = 0xffffffffu;
                sm[BP3_MAX_TEAM] = (int)t;
            }
            if (dp_thread() < BP3_MAX_TEAM)
                sm[dp_thread()] = 0;
            dp_syncblock();
            const unsigned t = (unsigned)dp_ld_volatile(&sm[BP3_MAX_TEAM]);
            if (t >= (unsigned)p.nsweeps)
                break;
            const int4 it = p.order[t];
            if (w < it.z)
                run_item(p, it.x + w, it.y, 0, p.H, nullptr, w > 0 ? sm + (w - 1) : nullptr,
                         w + 1 < it.z ? sm + w : nullptr, w + 1 == it.z);
            dp_syncblock();                         /* the counters are cleared for the next item */
        }
    }

    CA_MDEV void kernel_body(const Bp3Params &p)
    {
        if (p.pub_workers <= 0) {
            work_loop(p, nullptr);
            return;
        }
        CA_SHARED(PubSlot, slots, BP3_MAX_PUB_WORKERS + 1);     /* last entry: exit counter */
        int *nexit = &slots[BP3_MAX_PUB_WORKERS].done;
        const int nw = dp_block_threads() / 32 - 1;
        const int w = dp_warp_in_block();
        if (dp_thread() <= BP3_MAX_PUB_WORKERS) {
            slots[dp_thread()].flag = 0ull;
            slots[dp_thread()].done = 0;
            slots[dp_thread()].pub = 0;
        }
        dp_syncblock();
        if (w < nw) {
            work_loop(p, slots + w);
            dp_syncwarp();
            if (dp_lane() == 0)
                dp_atomic_add_cta(nexit, 1);
        } else {
            publish_loop(slots, nexit, nw);
        }
    }
};

/*
 * Self-publishing mode: 128-thread CTAs, 5 (2 words per lane) / 6 (1 word per lane) of them per SM.
 * Publisher mode: one CTA of up to 640 / 768 threads per SM (19 / 23 workers + the publisher).  The register
 * cap is the same either way: 65536 / 640 = 102, 65536 / 768 = 85.
 */
template <int P, int WPL>
struct Bp3Bounds {
    static constexpr int kMaxThreads = (P == 3 && WPL == 2) ? 640 : ((P <= 4 && WPL == 1) ? 768 : 160);
};

/*
 * Team mode is a kernel of its own (one CTA per SM): 512 threads leave every thread 128 registers -- the row
 * loop of the common variants then needs no spills -- and the widest variants run 256-thread teams.
 */
template <int P, int WPL>
struct Bp3TeamBounds {
    static constexpr int kMaxThreads = (P <= 4 && WPL <= 2) ? 512 : 256;
};

/* largest team (warps per CTA) of a variant */
inline int bp3_team_cap(int P, int WPL)
{
    const int t = ((P <= 4 && WPL <= 2) ? 512 : 256) / 32;
    return t < BP3_MAX_TEAM ? t : (int)BP3_MAX_TEAM;
}

template <int P, int WPL, class Rule>
CA_GLOBAL void __launch_bounds__(Bp3Bounds<P, WPL>::kMaxThreads, 1) ca3d_sweep_kernel(Bp3Params p)
{
    Sweep3<P, WPL, Rule>::kernel_body(p);
}

template <int P, int WPL, class Rule>
CA_GLOBAL void __launch_bounds__(Bp3TeamBounds<P, WPL>::kMaxThreads, 1) ca3d_team_kernel(Bp3Params p)
{
    Sweep3<P, WPL, Rule>::team_loop(p);
}

} // namespace clapca
#endif
