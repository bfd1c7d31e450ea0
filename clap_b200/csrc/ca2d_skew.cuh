/*
 * ca2d_skew.cuh -- the 2D sweep on DIAGONALS: ca2d_step() (core/ca2d.c:61-77) for one-plane (binary) grids with
 * the alive-bit neighbourhoods (core/ca2d.c:11-33), without any in-row dependency.
 *
 * The reference sweeps x outer / y inner in place, so cell (x,y) sees NEW values at (x-1,y-1), (x-1,y), (x-1,y+1),
 * (x,y-1) and OLD values at (x,y+1), (x+1,y-1), (x+1,y), (x+1,y+1).  With tau = 2x + y those eight neighbours sit
 * on the diagonals tau-3, tau-2, tau-1, tau-1 (new) and tau+1, tau+1, tau+2, tau+3 (old): cells of EQUAL tau are
 * mutually independent.  The row engine (ca2d_bitplane.cuh) keeps rows of equal x and pays a scan + a CTA barrier
 * per row to resolve the (x,y-1) -> (x,y) chain: 16384 dependent row steps of ~0.54 us at BASELINE config 3.  Here
 * the grid is stored SKEWED -- row t of the record set holds diagonal t, bit x of it is cell (x, t - 2x) -- and a
 * step is pure LOP3 work:
 *
 *   new part   l1 = new(t-1) << 1   cell (x-1,y+1)      n1 = new(t-1)        cell (x,y-1)
 *              l2 = new(t-2) << 1   cell (x-1,y)        l3 = new(t-3) << 1   cell (x-1,y-1)
 *   old part   o1 = old(t+1)        cell (x,y+1)        r1 = old(t+1) >> 1   cell (x+1,y-1)
 *              r2 = old(t+2) >> 1   cell (x+1,y)        r3 = old(t+3) >> 1   cell (x+1,y+1)
 *
 *   K = l2 + l3 + o1 + r1 + r2 + r3 is known one step early; the tables t_j = rule(K + j), j = 0..2, too; the only
 *   work behind the previous step's result is  new(t) = l1 ? (n1 ? t2 : t1) : (n1 ? t1 : t0)  -- one funnel shift
 *   and three LOP3, 2 (W-1) + H steps instead of W row steps of ~150 dependent instructions.
 *
 * Parallelism.  One CTA per generation (claimed by ticket, all generations in flight in one cooperative launch,
 * generation g+1 two step groups behind g, progress counters as in the row engine).  A diagonal is cut by x: warp w
 * owns 31 lanes x WPL consecutive words; its lane 31 is a HALO lane that only loads the words right of the span (they
 * are lane 0's of the next warp) and hands them to lane 30 by shuffle -- one load per lane and step, no special case.
 * Only x - 1 feeds a cell from the NEW side, so the warps form a chain, not a ring: lanes hand bit 31 to the right by
 * shuffle, a warp hands it to the next warp through a tagged mailbox ring in shared memory (one word = step tag + bit:
 * no fence, no barrier), four posts fetched per look so that the steps themselves hold no branch, and the next warp
 * simply runs a few steps behind -- no CTA barrier anywhere in the sweep.  A warp only works while the band of valid
 * cells (0 <= t - 2x < H) crosses its words; groups of steps wholly inside the band run without masks.
 *
 * Old rows come from L2 (the previous generation wrote them on another SM) through a register window SLOTS steps deep.
 * What ncu taught about that window on B200 (profiles/r02_ncu_full_ca2d_diagonal_first.txt): a load must land in a
 * register that is DEAD (the slot of the row used one step earlier), a two-word slot must stay ONE 64-bit value until
 * its first use, and there must be one load per step -- otherwise ptxas parks the result in a shared temporary and
 * MOVes it out a few instructions later, and every step waits a full L2 latency for a row it needs eleven steps on
 * (25 ms instead of 13 at BASELINE config 3).  mbarrier-armed bulk copies were considered for this stream and dropped
 * on the guide's numbers: a try_wait costs 60-90 cycles against a step budget of ~100.
 *
 * Measured (profiles/r02_ca2d_diagonal_ab.txt): bit-identical to the row engine at config 3, 13.1 ms against its 8.9 ms
 * -- a step is bound by ONE warp's dependent instruction stream.  Hence CLAPCA_ENGINE_AUTO measures both engines on the
 * first large run of a shape class instead of assuming one (clapca_api.cu:run2d_tune).
 */
#ifndef CLAPCA_CA2D_SKEW_CUH
#define CLAPCA_CA2D_SKEW_CUH

#include "bitslice.cuh"
#include "ca3d_bitplane.cuh"        /* LaneVec */
#include "ca2d_bitplane.cuh"        /* bp2_rule_for, Rule2Dyn's table walk */

namespace clapca {

enum { SK2_SLOTS = 12, SK2_MAX_WARPS = 18, SK2_RING = 48, SK2_PAD_WORDS = 6 };
/* lanes 0..30 of a warp own words; lane 31 only LOADS the words right of the span (they are lane 0's of the next warp)
   and hands them to lane 30 by shuffle: one load per lane and step, no lane-31 special case in the step */
enum { SK2_OWN_LANES = 31 };
/* words between diagonals: fixed (18 warps x 31 lanes + pad, a multiple of 4), so that every row address of a group of
   steps is the group's pointer + a constant */
enum { SK2_RS = SK2_MAX_WARPS * SK2_OWN_LANES + SK2_PAD_WORDS };
CA_HOSTDEV int sk2_warp_cells(int WPL)              { return SK2_OWN_LANES * 32 * WPL; }
enum { SK2_SM_DONE = 0, SK2_SM_RING = SK2_MAX_WARPS, SK2_SM_TICKET = SK2_SM_RING + SK2_MAX_WARPS * SK2_RING,
       SK2_SMEM_WORDS = SK2_SM_TICKET + 1 };
#define SK2_INF 0x7fffffff

struct Sk2Params {
    uint32_t *rows;         /* [TR][SK2_RS]: bit x of row t = cell (x, t - 2x); rows >= T and the pad words stay zero */
    int W, H, G;            /* x extent, y extent (reference index y*W + x), generations */
    int T;                  /* diagonals that hold cells: 2 (W-1) + H */
    int *prog;              /* [G] diagonals completed by generation g */
    unsigned *ticket;       /* next generation to claim */
    int *err;               /* != 0: watchdog fired, everybody bails out */
    uint32_t born, surv;    /* 9-bit masks; surv is all ones when the rule does not decay */
    uint32_t nrval;         /* (uint8_t)nr_states: 1, or 0 (then nothing is ever born) */
    long long spin_limit;
};

/* host + device: geometry of the skewed record set */
CA_HOSTDEV int sk2_diagonals(int W, int H)          { return 2 * (W - 1) + H; }
CA_HOSTDEV int sk2_rows_alloc(int W, int H)
{
    const int T = sk2_diagonals(W, H);
    /* the sweep's last group prefetches up to 2 SLOTS rows past T; unpack reads 62 rows past the last diagonal of a
       ragged last word column */
    return (T + SK2_SLOTS - 1) / SK2_SLOTS * SK2_SLOTS + 2 * SK2_SLOTS + 64;
}
/* at most SK2_MAX_WARPS / WPL warps (the kernels' launch bounds): 17856 cells per diagonal either way */
CA_HOSTDEV int sk2_warps_for(int W, int WPL)        { return (W + sk2_warp_cells(WPL) - 1) / sk2_warp_cells(WPL); }

/* bits 0 .. k-1 (none for k <= 0, all for k >= 32): one clamped funnel shift */
CA_DEV uint32_t sk2_ones_below(int k)               { return dp_funnel_l_clamp(~0u, 0u, (uint32_t)(k > 0 ? k : 0)); }

/*
 * Rules: the new alive bit of a cell with K counted neighbours (K = 0..6, bits k0..k2) plus j = 0..2 more, current
 * alive bit a.  kFold: surv[n] == born[n+1] for every n (cave smoothing: "alive next iff n + a >= 5") -- then the
 * alive bit goes into the count as a carry-in and one LOP3 per table is left.
 */
template <uint32_t BORN, uint32_t SURV>
struct Sk2RuleConst {
    static constexpr bool kFold = ((SURV & 0xFFu) == ((BORN >> 1) & 0xFFu));
    static constexpr uint32_t E = (BORN & 0x1FFu) | (((SURV >> 8) & 1u) << 9);   /* table over n + a = 0..9 */
    struct Tabs { };
    CA_MDEV void setup(const Sk2Params &, Tabs &) { }
    CA_MDEV void eval(const Tabs &, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t a, uint32_t t[3])
    {
        if constexpr (kFold) {      /* k = K + a */
            t[0] = bs_tab3<E & 0xFFu>(k0, k1, k2);
            t[1] = bs_tab3<(E >> 1) & 0xFFu>(k0, k1, k2);
            t[2] = bs_tab3<(E >> 2) & 0xFFu>(k0, k1, k2);
        } else {
            t[0] = bs_mux(a, bs_tab3<SURV & 0xFFu>(k0, k1, k2), bs_tab3<BORN & 0xFFu>(k0, k1, k2));
            t[1] = bs_mux(a, bs_tab3<(SURV >> 1) & 0xFFu>(k0, k1, k2), bs_tab3<(BORN >> 1) & 0xFFu>(k0, k1, k2));
            t[2] = bs_mux(a, bs_tab3<(SURV >> 2) & 0xFFu>(k0, k1, k2), bs_tab3<(BORN >> 2) & 0xFFu>(k0, k1, k2));
        }
    }
};

struct Sk2RuleDyn {
    static constexpr bool kFold = false;
    struct Tabs { uint32_t s[10], b[10]; };      /* mask bits broadcast to words, once per sweep */
    CA_MDEV void setup(const Sk2Params &p, Tabs &tb)
    {
#pragma unroll
        for (int n = 0; n < 10; n++) {
            tb.s[n] = n < 9 ? bs_bit(p.surv, n) : 0u;
            tb.b[n] = (n < 9 && p.nrval) ? bs_bit(p.born, n) : 0u;   /* nr_states == 0: a "born" cell takes the value 0 */
        }
    }
    CA_MDEV uint32_t walk(const uint32_t *m, uint32_t k0, uint32_t k1, uint32_t k2)
    {
        uint32_t a = bs_mux(k0, m[1], m[0]), b = bs_mux(k0, m[3], m[2]);
        uint32_t c = bs_mux(k0, m[5], m[4]), d = bs_mux(k0, m[7], m[6]);
        return bs_mux(k2, bs_mux(k1, d, c), bs_mux(k1, b, a));
    }
    CA_MDEV void eval(const Tabs &tb, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t a, uint32_t t[3])
    {
#pragma unroll
        for (int j = 0; j < 3; j++)
            t[j] = bs_mux(a, walk(tb.s + j, k0, k1, k2), walk(tb.b + j, k0, k1, k2));
    }
};

typedef Sk2RuleConst<0x1E0u, 0x1F0u> Sk2RuleCave;
typedef Sk2RuleConst<0x00Cu, 0x180u> Sk2RuleTest;

/*
 * A window slot: the lane's words of one old row.  Two words per lane are ONE 64-bit value from the load to the use:
 * as two 32-bit values ptxas gave a few slots unpaired home registers, loaded those rows into a scratch pair and
 * MOVed them out five instructions later -- a full L2 latency inside the step.  The halves are taken where the row is
 * first used (three steps before its diagonal), not where ptxas would like to (next to the load).
 */
template <int WPL> struct Sk2Win;
template <> struct Sk2Win<1> {
    uint32_t v;
    CA_MEMBER void load(const uint32_t *p) { v = dp_ld_cg(p); }
    CA_MEMBER void unpack(uint32_t w[1]) const { w[0] = v; }
};
template <> struct Sk2Win<2> {
    unsigned long long v;
    CA_MEMBER void load(const uint32_t *p) { v = dp_ld_cg(reinterpret_cast<const unsigned long long *>(p)); }
    CA_MEMBER void unpack(uint32_t w[2]) const { dp_unpack64_here(v, w[0], w[1]); }
};

template <int WPL, bool MOORE, class Rule>
struct Skew2 {
    enum { SLOTS = SK2_SLOTS };
    static constexpr bool kSteadyBuild = false;
    static_assert(SLOTS % 3 == 0 && SK2_RING % SLOTS == 0, "the shifted rows rotate with period 3; a group never wraps the ring");

    struct St {
        Sk2Win<WPL> win[SLOTS];         /* old rows t+4 .. t+SLOTS-1 as loaded, slot = row % SLOTS */
        uint32_t uw[SLOTS][WPL];        /* old rows t .. t+3 (von Neumann: t+2), unpacked */
        uint32_t rsh[3][WPL];           /* old rows t+1 .. t+3 shifted to x+1, slot = row % 3 */
        uint32_t lsh[3][WPL];           /* new rows t-1 .. t-3 shifted to x-1, slot = row % 3 */
        uint32_t n1[WPL];               /* new row t-1 */
        uint32_t wmask[WPL];            /* x < W */
        uint32_t *row;                  /* the lane's words of the current group's first row */
        const int *flagp;               /* lane 0: counter of the previous generation, else null */
        int have;
        int xb;                         /* x of bit 0 of the lane's first word */
        const uint32_t *in_ring;        /* posts of the warp on the left: bit 31 of its last word, steps in_first .. in_last */
        uint32_t *out_ring;             /* this warp's posts, steps out_first .. out_last */
        const uint32_t *in_prev;        /* slot of the post of the step before the current group */
        int rb;                         /* ring slot of the current group's first step */
        uint32_t mailbits;              /* bit k: the left warp's carry into lane 0 at step k of the current block of posts */
        bool has_left, feeds;
        int in_first, in_span, out_first, out_last;
        const int *cons_done;           /* the consumer's step counter */
        int cons_start, cons_ok;        /* posts up to step cons_ok fit the ring without asking */
        typename Rule::Tabs tabs;
    };

    /* rows < need of the previous generation are complete */
    CA_MDEV void wait_rows(const Sk2Params &p, St &st, int need)
    {
        if (st.have >= need)
            return;
        long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            int v = st.flagp ? dp_ld_acquire(st.flagp) : SK2_INF;
            st.have = dp_reduce_min(v);
            if (st.have >= need)
                break;
            if (spins == 0) t0 = dp_clock();
            /* a warp whose band is far away sleeps long: its polls would take issue slots and L2 bandwidth from the
               warps that are working (a step is ~30 ns) */
            dp_nanosleep(need - st.have > 256 ? 2000 : (need - st.have > 48 ? 300 : 20));
            if ((spins & 127u) == 127u) {
                bool bad = dp_ld_flag(p.err) != 0 || (dp_clock() - t0) > p.spin_limit;
                if (!dp_all(!bad)) {
                    if (dp_lane() == 0)
                        dp_atomic_max(p.err, 1);
                    st.have = SK2_INF;      /* keep going to the end of the sweep; the claim loop exits on err */
                    break;
                }
            }
        }
        dp_syncwarp();
    }

    /* the ring has room for this warp's posts up to step `upto` */
    CA_MDEV void wait_ring(const Sk2Params &p, St &st, int upto)
    {
        long long t0 = 0;
        for (unsigned spins = 0; st.cons_ok < upto; spins++) {
            int d = (int)dp_shfl(dp_lane() == 0 ? (uint32_t)dp_ld_volatile(st.cons_done) : 0u, 0);   /* one look, everybody follows */
            if (d < st.cons_start) d = st.cons_start;
            st.cons_ok = d >= SK2_INF - SK2_RING ? SK2_INF : d + SK2_RING - 2;
            if (st.cons_ok >= upto)
                break;
            if (spins == 0) t0 = dp_clock();
            dp_team_pause();
            if ((spins & 127u) == 127u) {
                bool bad = dp_ld_flag(p.err) != 0 || (dp_clock() - t0) > p.spin_limit;
                if (!dp_all(!bad)) {
                    if (dp_lane() == 0)
                        dp_atomic_max(p.err, 4);
                    st.cons_ok = SK2_INF;
                    break;
                }
            }
        }
        dp_fence_cta();     /* the consumer's reads of the slots precede its counter; this warp's stores follow */
    }

    /*
     * The left warp's posts of the MB steps before steps t .. t+MB-1 (post s: slot s % RING, tag s + 1, bit 0 = bit 31
     * of its last owned word), fetched TOGETHER so that the steps themselves hold no branch: a warp runs at least MB
     * steps behind its left neighbour.  Result: bit k of st.mailbits = the carry into lane 0 at step t + k.
     * S = t % SLOTS.  Every decision is taken on a vote or on one lane's broadcast load: lanes that polled on their own
     * could disagree and leave the loop apart -- into different warp collectives.
     */
    enum { MB = 4 };
    template <int S, bool STEADY>
    CA_MDEV void fetch_posts(const Sk2Params &p, St &st, int t)
    {
        static_assert(S % MB == 0 && SLOTS % MB == 0, "blocks of posts are aligned to the step groups");
        /* the left warp posts at every step from its first to its last (in_first + in_span), which lies behind this
           warp's first step by construction: only the END of its posts has to be looked at */
        const int in_last = st.in_first + st.in_span;
        if (!st.has_left || t - 1 > in_last) {
            st.mailbits = 0u;
            return;
        }
        const bool all = STEADY || t + MB - 2 <= in_last;
        long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            uint32_t bits = 0u;
            bool ok = true;
#pragma unroll
            for (int k = 0; k < MB; k++) {
                const uint32_t *slot = (S + k) ? st.in_ring + st.rb + (S + k - 1) : st.in_prev;
                const bool want = all || t + k - 1 <= in_last;
                const uint32_t m = dp_ld_volatile_u32(slot);
                ok = ok && (!want || (m >> 1) == (uint32_t)(t + k));
                bits |= (want ? (m & 1u) : 0u) << k;
            }
            if (dp_all(ok)) {
                st.mailbits = bits;
                return;
            }
            /* right behind the producer: poll fast; a warp waiting for its band to arrive backs off to 2 us */
            if (spins == 0) t0 = dp_clock();
            dp_nanosleep(spins < 8 ? 20 : (spins < 64 ? 200 : 2000));
            if ((spins & 255u) == 255u || spins == 8) {    /* spins == 8: a watchdog that fired elsewhere is noticed early */
                bool bad = dp_ld_flag(p.err) != 0 || (dp_clock() - t0) > p.spin_limit;
                if (!dp_all(!bad)) {
                    if (dp_lane() == 0)
                        dp_atomic_max(p.err, 5);
                    st.mailbits = 0u;   /* the watchdog has fired: run on to the end, the claim loop exits on err */
                    return;
                }
            }
        }
    }

    /* row (group base + D) of the lane's words into window slot S */
    template <int D, int S>
    CA_MDEV void load_row(St &st)
    {
        st.win[S].load(st.row + D * SK2_RS);
    }

    /*
     * One diagonal.  S = t % SLOTS; st.row points at the lane's words of row t - S (the group's first row): every
     * address of the group is that pointer + a constant.  STEADY: every cell of the warp is inside the grid on all
     * diagonals of the group (no masks) and every step reads a post (warps > 0) and writes one (warps that feed a
     * neighbour); otherwise masks and the two post windows are looked at step by step.
     */
    template <int S, bool STEADY>
    CA_MDEV void step(const Sk2Params &p, St &st, int t)
    {
        constexpr int S1 = (S + 1) % SLOTS, S3 = (S + 3) % SLOTS;
        constexpr int I0 = S % 3, I1 = (S + 1) % 3, I2 = (S + 2) % 3;      /* t % 3, (t+1) % 3, (t+2) % 3 */
        const int lane = dp_lane();

        /* ---- this step's rows out of the window; the row three steps ahead shifted to x+1 ---- */
        uint32_t cur[WPL], o1[WPL];
        {
            constexpr int SR = MOORE ? S3 : (S + 2) % SLOTS;       /* von Neumann: only (x+1,y) = row t+2 from the right */
            constexpr int IR = MOORE ? I0 : I2;
            st.win[SR].unpack(st.uw[SR]);
            const uint32_t next = dp_shfl_down(st.uw[SR][0], 1);   /* lane 31 (the halo lane) gets its own word: unused */
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                cur[j] = st.uw[S][j];
                o1[j] = st.uw[S1][j];
                st.rsh[IR][j] = dp_funnel_r(st.uw[SR][j], j + 1 < WPL ? st.uw[SR][j + 1] : next, 1);
            }
        }
        /* row t + SLOTS - 1 takes the slot of row t - 1, which the previous step read for the last time: the load has a
           dead register to land in (into the slot of row t itself it landed in a temporary and was MOVed over right
           away -- a full L2 latency inside every step, ncu: 1000 cycles per step) */
        load_row<S + SLOTS - 1, (S + SLOTS - 1) % SLOTS>(st);

        /* ---- everything that does not depend on the previous step ---- */
        uint32_t tb[WPL][3];
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            uint32_t k0, k1, k2;
            const uint32_t cin = Rule::kFold ? cur[j] : 0u;
            if (MOORE) {
                /* old part o1 + r1 + r2 + r3 (0..4), new part l2 + l3 (0..2), [+ alive bit] */
                const uint32_t r1 = st.rsh[I1][j], r2 = st.rsh[I2][j], r3 = st.rsh[I0][j];
                const uint32_t l2 = st.lsh[I1][j], l3 = st.lsh[I0][j];      /* rows t-2, t-3 */
                const uint32_t s = bs_xor3(r1, r2, r3), c = bs_maj3(r1, r2, r3);
                const uint32_t a0 = s ^ o1[j], ca = s & o1[j];
                const uint32_t a1 = c ^ ca, a2 = c & ca;
                const uint32_t b0 = l2 ^ l3, b1 = l2 & l3;
                k0 = bs_xor3(a0, b0, cin);
                const uint32_t c0 = bs_maj3(a0, b0, cin);
                k1 = bs_xor3(a1, b1, c0);
                k2 = a2 | bs_maj3(a1, b1, c0);          /* K + a <= 7 */
            } else {
                /* K = (x-1,y) + (x+1,y) + (x,y+1) [+ alive bit] = 0..4 */
                const uint32_t l2 = st.lsh[I1][j], r2 = st.rsh[I2][j];
                const uint32_t s = bs_xor3(l2, r2, o1[j]), c = bs_maj3(l2, r2, o1[j]);
                k0 = s ^ cin;
                const uint32_t c0 = s & cin;
                k1 = c ^ c0;
                k2 = c & c0;
            }
            Rule::eval(st.tabs, k0, k1, k2, cur[j], tb[j]);
        }
        uint32_t vm[WPL];
        if (!STEADY) {
            const int hi = (t >> 1) + 1, lo = (t - p.H + 2) >> 1;      /* cells lo <= x < hi are in the grid on this diagonal */
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                const int b = st.xb + 32 * j;
                vm[j] = sk2_ones_below(hi - b) & ~sk2_ones_below(lo - b) & st.wmask[j];
            }
        }

        /* ---- behind the previous step: the carry from the left, one shift, three LOP3 ---- */
        uint32_t left = dp_shfl_up(st.n1[WPL - 1], 1);
        if (lane == 0) left = st.mailbits << (31 - S % MB);        /* bit 31 = the left warp's carry of step t-1 */
        uint32_t nw[WPL];
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            const uint32_t n1 = st.n1[j];
            if (MOORE) {
                const uint32_t l1 = dp_funnel_l(j ? st.n1[j - 1] : left, n1, 1);
                st.lsh[I2][j] = l1;                                    /* row t-1 */
                uint32_t A0 = bs_mux(n1, tb[j][1], tb[j][0]), A1 = bs_mux(n1, tb[j][2], tb[j][1]);
                if (!STEADY) {      /* the mask on the two candidates: beside the carry's way through the shuffle, not behind it */
                    A0 &= vm[j];
                    A1 &= vm[j];
                }
                nw[j] = bs_mux(l1, A1, A0);
            } else {
                st.lsh[I2][j] = dp_funnel_l(j ? st.n1[j - 1] : left, n1, 1);
                nw[j] = bs_mux(n1, tb[j][1], tb[j][0]);
                if (!STEADY)
                    nw[j] &= vm[j];
            }
        }
        if (lane < SK2_OWN_LANES)
            LaneVec<WPL>::st(st.row + S * SK2_RS, nw);
        /* this warp's post of step t: slot t % RING, tag t + 1 */
        /* posted at EVERY step of a warp that feeds a neighbour: before the first cell of the neighbour's window the bit
           is 0 and nobody waits for it; the ring's back-pressure (sweep) only covers the steps that are read */
        if (lane == SK2_OWN_LANES - 1 && st.feeds)
            dp_st_volatile((int *)st.out_ring + st.rb + S, (int)((nw[WPL - 1] >> 31) + (((uint32_t)t + 1u) << 1)));
#pragma unroll
        for (int j = 0; j < WPL; j++) st.n1[j] = nw[j];
    }

    template <int S, bool STEADY>
    CA_MDEV void steps_from(const Sk2Params &p, St &st, int t)
    {
        if constexpr (S < SLOTS) {
            if constexpr (S % MB == 0)
                fetch_posts<S, STEADY>(p, st, t);
            step<S, STEADY>(p, st, t);
            steps_from<S + 1, STEADY>(p, st, t + 1);
        }
    }

    /* all diagonals of generation g that cross this warp's words */
    CA_MDEV void sweep(const Sk2Params &p, int g, uint32_t *smem)
    {
        const int lane = dp_lane(), warp = dp_warp_in_block(), nw = (dp_block_threads() >> 5) - 1;
        const int H = p.H, W = p.W;
        const int X0 = warp * sk2_warp_cells(WPL), X1 = X0 + sk2_warp_cells(WPL);
        const int Xe = X1 < W ? X1 : W;                         /* cells X0 <= x < Xe exist */
        const int first = X0 ? 2 * X0 - 1 : 0;                  /* one step before the warp's first cell (X0, 0) */
        const int start = first / SLOTS * SLOTS;
        const int last = 2 * (Xe - 1) + H - 1;                  /* diagonal of the warp's last cell (Xe-1, H-1) */
        int *done = (int *)smem + SK2_SM_DONE + warp;
        St st;
        st.xb = X0 + 32 * WPL * lane;
        st.row = p.rows + (size_t)start * SK2_RS + (size_t)(warp * SK2_OWN_LANES + lane) * WPL;
        st.flagp = (g > 0 && lane == 0) ? p.prog + (g - 1) : nullptr;
        st.have = g > 0 ? 0 : SK2_INF;
        /* posts: the last cell of the warp on the left is (X0-1, y), on diagonals 2 X0 - 2 .. 2 X0 - 2 + H - 1 */
        st.has_left = warp > 0;
        st.in_ring = smem + SK2_SM_RING + (warp ? warp - 1 : 0) * SK2_RING;
        st.in_first = warp ? 2 * X0 - 2 : 0x3fffffff;           /* warp 0: no step is ever inside the (unsigned) span */
        st.in_span = warp ? H - 1 : 0;
        st.out_ring = smem + SK2_SM_RING + warp * SK2_RING;
        st.feeds = warp + 1 < nw && X1 < W;
        st.out_first = st.feeds ? 2 * X1 - 2 : SK2_INF;
        st.out_last = st.feeds ? 2 * X1 - 2 + H - 1 : -1;
        st.cons_done = (const int *)smem + SK2_SM_DONE + (st.feeds ? warp + 1 : warp);
        st.cons_start = st.feeds ? (2 * X1 - 1) / SLOTS * SLOTS : 0;
        st.cons_ok = st.feeds ? st.cons_start + SK2_RING - 2 : SK2_INF;
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            st.wmask[j] = lane < SK2_OWN_LANES ? bp_valid_mask((st.xb >> 5) + j, W) : 0u;
            st.n1[j] = 0u;
#pragma unroll
            for (int i = 0; i < 3; i++) st.lsh[i][j] = st.rsh[i][j] = 0u;
        }
        Rule::setup(p, st.tabs);

        /* the window: rows start .. start + SLOTS - 2 (the last slot is loaded by the first step), and rows start+1,
           start+2 shifted to x+1 */
        wait_rows(p, st, start + SLOTS < p.T ? start + SLOTS : p.T);
        fill_window<0>(st);
        {
            st.win[0].unpack(st.uw[0]);
            st.win[1].unpack(st.uw[1]);
            st.win[2].unpack(st.uw[2]);
            const uint32_t next1 = dp_shfl_down(st.uw[1][0], 1), next2 = dp_shfl_down(st.uw[2][0], 1);
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                st.rsh[1][j] = dp_funnel_r(st.uw[1][j], j + 1 < WPL ? st.uw[1][j + 1] : next1, 1);
                st.rsh[2][j] = dp_funnel_r(st.uw[2][j], j + 1 < WPL ? st.uw[2][j + 1] : next2, 1);
            }
        }

        /* groups whose steps all see every cell of the warp inside the grid (and, warps > 0, a post of the left warp) */
        const int full_lo = 2 * X1 - 2, full_hi = 2 * X0 + H - 2;
        st.rb = start % SK2_RING;
        for (int t = start; t <= last; t += SLOTS) {
            wait_rows(p, st, t + 2 * SLOTS < p.T ? t + 2 * SLOTS : p.T);
            if (t + SLOTS - 1 >= st.out_first && t <= st.out_last && st.cons_ok < t + SLOTS - 1)
                wait_ring(p, st, t + SLOTS - 1);
            st.in_prev = st.in_ring + (st.rb ? st.rb - 1 : SK2_RING - 1);
            /*
             * The mask-free STEADY build of the steps exists (kSteadyBuild) but is not instantiated: at any moment one
             * warp of the chain is entering the band and one is leaving it, the chain moves at THEIR pace, and a second
             * copy of the unrolled group doubles the code the SM's instruction caches have to hold (ncu, first version:
             * 0.36 cycles per issued instruction without an instruction to issue).
             */
            if (kSteadyBuild && X1 <= W && t >= full_lo && t + SLOTS - 1 <= full_hi)
                steps_from<0, kSteadyBuild>(p, st, t);
            else
                steps_from<0, false>(p, st, t);
            st.row += SLOTS * SK2_RS;
            st.rb = st.rb + SLOTS == SK2_RING ? 0 : st.rb + SLOTS;
            /* every lane's stores of this group, then the warp's step counter (CTA scope; the publisher warp carries it on) */
            dp_fence_cta();
            dp_syncwarp();
            if (lane == 0) {
                dp_fence_cta();     /* release in the storing thread, after it has synchronised with the other lanes */
                dp_st_volatile(done, t + SLOTS);
            }
        }
        dp_syncwarp();
        if (lane == 0)
            dp_st_volatile(done, SK2_INF);
    }

    template <int S>
    CA_MDEV void fill_window(St &st)
    {
        if constexpr (S < SLOTS - 1) {
            load_row<S, S>(st);
            fill_window<S + 1>(st);
        }
    }

    /* the publisher warp: min over the compute warps' counters -> prog[g]; the only gpu-scope fences of the CTA */
    CA_MDEV void publish(const Sk2Params &p, int g, const uint32_t *smem)
    {
        const int lane = dp_lane(), nw = (dp_block_threads() >> 5) - 1;
        const int X0 = lane * sk2_warp_cells(WPL);
        const int first = X0 ? 2 * X0 - 1 : 0;
        const int start = first / SLOTS * SLOTS;        /* rows before a warp's first group are complete as far as it goes */
        int pub = 0;
        long long t_idle = dp_clock();
        for (unsigned spins = 0; pub < p.T; spins++) {
            int d = SK2_INF;
            if (lane < nw) {
                d = dp_ld_volatile((const int *)smem + SK2_SM_DONE + lane);
                if (d < start) d = start;
            }
            int m = dp_reduce_min(d);
            if (m > p.T) m = p.T;
            if (m > pub) {
                dp_fence_cta();
                dp_fence_release();
                if (lane == 0)
                    dp_st_flag(p.prog + g, m);
                pub = m;
                t_idle = dp_clock();
            } else {
                dp_nanosleep(40);
                if ((spins & 255u) == 255u) {
                    bool bad = dp_ld_flag(p.err) != 0 || (dp_clock() - t_idle) > 4 * p.spin_limit;
                    if (!dp_all(!bad)) {
                        if (lane == 0)
                            dp_atomic_max(p.err, 3);
                        break;
                    }
                }
            }
        }
    }

    CA_MDEV void kernel_body(const Sk2Params &p, uint32_t *smem)
    {
        const int warp = dp_warp_in_block(), nw = (dp_block_threads() >> 5) - 1;
        for (;;) {
            if (dp_thread() == 0) {
                unsigned t = dp_atomic_inc(p.ticket);
                if (dp_ld_flag(p.err) != 0)
                    t = 0xffffffffu;
                smem[SK2_SM_TICKET] = t;
            }
            for (int i = dp_thread(); i < SK2_SM_TICKET; i += dp_block_threads())
                smem[i] = 0u;                       /* step counters and mailbox tags restart with every generation */
            dp_syncblock();
            const unsigned g = smem[SK2_SM_TICKET];
            if (g >= (unsigned)p.G)
                break;
            if (warp < nw)
                sweep(p, (int)g, smem);
            else
                publish(p, (int)g, smem);
            dp_syncblock();         /* everything of this generation is out before the shared words are reused */
        }
    }
};

template <int WPL, bool MOORE, class Rule>
CA_GLOBAL void __launch_bounds__(32 * (SK2_MAX_WARPS / WPL + 1), 1) ca2d_skew_kernel(Sk2Params p)
{
    CA_SHARED(uint32_t, smem, SK2_SMEM_WORDS);
    Skew2<WPL, MOORE, Rule>::kernel_body(p, smem);
}

} // namespace clapca
#endif
