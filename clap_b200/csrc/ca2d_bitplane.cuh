/*
 * ca2d_bitplane.cuh -- bit-plane engine for ca2d_step() (core/ca2d.c:61-77) with the alive-bit
 * neighbourhoods ca2d_neigh_m1 / ca2d_neigh_vn1 (core/ca2d.c:11-33).  The value-comparing
 * neighbourhoods (vnv / mv, :35-59) reduce to these when the rule does not decay -- a non-zero cell then
 * never changes and a zero cell counts neighbours > 0 -- which the C-ABI layer exploits; with decay they
 * run on the cell-wavefront engine (ca_wavefront.cuh).
 *
 * The reference sweeps x outer / y inner over an array indexed y*side + x and updates in place, so cell
 * (x,y) sees the NEW row x-1 (all three of y-1, y, y+1), the new cell (x,y-1), the old cell (x,y+1) and
 * the OLD row x+1.  Transposing the grid (pack kernel, ca2d_layout.cuh) turns "row" into the reference's
 * x and the in-row position into y; then, exactly like the 3D engine (ca3d_bitplane.cuh):
 *
 *   K      = H3(new row x-1) + H3(old row x+1) + a_old(y+1)          [Moore; H3 = a(y-1)+a(y)+a(y+1)]
 *   K      = a_new(x-1,y) + a_old(x+1,y) + a_old(x,y+1)              [von Neumann]
 *   n      = K + a_new(x,y-1)                   -- the only serial dependency inside a row
 *   a'(y)  = f0(y) ^ (a'(y-1) & (f0(y) ^ f1(y)))                     -- GF(2) affine scan (bitslice.cuh)
 *
 * Layout: row record x = [S0 .. S(P-1)] plane-rows of RWS 32-bit words, bit i of word w = cell y = 32w+i,
 * padding bits always 0.  P = 1 for binary rules (BASELINE config 3: 16384^2 cells = 32 MiB).
 *
 * Parallelism.  One CTA sweeps one generation; warp w / lane l owns WPL consecutive words of every row, and the
 * in-row chain is resolved word -> warp (ballots) -> CTA (one packed word per warp in shared memory, ONE barrier
 * per row over the compute warps).  The row-to-row chain (row x+1 needs the new row x) is the critical path of the
 * whole run -- side + 2 G row steps -- so a row step is built for a short DEPENDENT instruction stream (round 1:
 * 326 instructions per warp and row, 0.79 us per row, ncu: stalled on fixed-latency dependencies, not on memory):
 *
 *   - the rule is a template parameter where it matters (one LOP3 per table instead of a 7-deep mux tree over
 *     run-time mask bits; the run-time rule broadcasts its mask bits to words once per sweep);
 *   - monotone rules (cave smoothing) resolve the in-row chain with integer adds -- the carry chain of f0 + f1 --
 *     at every level (word, warp ballots, CTA) instead of three 5-step Kogge-Stone scans;
 *   - rows are prefetched three rows ahead into a 4-slot register window: the L2 latency of the previous
 *     generation's rows is off the chain;
 *   - no gpu-scope fence in the compute warps: thread 0 raises a row counter in shared memory every few rows and ONE
 *     extra publisher warp per CTA (it does not take part in the row barrier) carries it to the global counter
 *     prog[g] that generation g+1 polls.
 *   (A barrier-free variant -- warps skewed, coupled to their neighbours by shared-memory mailboxes -- was measured
 *   on B200 and lost 3x: two polling hops per row cost more than one barrier; profiles/r02_ca2d_variants.txt.)
 *
 * Generation g follows generation g-1 a few rows behind (row x needs rows <= x+1 of the previous generation, the
 * prefetch asks for x+3), so all G generations are in flight in ONE launch.  Storage is single-buffered, like the
 * reference's: a row is overwritten only after its last reader has passed.  The whole working set (a few rows per
 * generation) lives in L2.
 */
#ifndef CLAPCA_CA2D_BITPLANE_CUH
#define CLAPCA_CA2D_BITPLANE_CUH

#include "bitslice.cuh"
#include "ca3d_bitplane.cuh"        /* LaneVec, bp_valid_mask */

namespace clapca {

struct Bp2Params {
    uint32_t *rows;         /* [M][P][RWS] */
    int N, M, G;            /* cells per row (reference y extent), rows (reference x extent), generations */
    int RWS;                /* words per plane-row = compute warps per CTA * 32 * WPL */
    int *prog;              /* [G] rows completed by generation g */
    unsigned *ticket;       /* next generation to claim */
    int *err;               /* != 0: watchdog fired, everybody bails out */
    uint32_t born, surv;    /* 9-bit masks; surv is all ones when the rule does not decay */
    uint32_t nrval;         /* (uint8_t)nr_states: the value a born cell takes (core/ca2d.c:71) */
    int flag_rows;          /* the shared-memory row counter is raised every flag_rows rows (the publisher warp takes it from there) */
    long long spin_limit;
};

enum { BP2_MAX_WARPS = 16 };
/*
 * shared memory of a CTA (words): four rotating slots of two words each for the warps' carry maps, the row counter, the
 * claimed generation.  A slot packs one bit per warp: word 0 = [generate | propagate << 16] (general rules: the warp
 * map's [D | C << 16]), word 1 = the same two bits of the warp's FIRST cell -- lane 0 of every warp ORs its bits in
 * before the row barrier, everybody reads the two words after it: the CTA-level chain starts from plain masks (no
 * ballots), and the slot of row x+2 is cleared by thread 0 right after the barrier of row x (its last readers, row
 * x-2's, are past barrier x-1; its next writers, row x+2's, are behind barrier x+1, which thread 0 has yet to reach).
 */
enum { BP2_SM_SLOT = 0, BP2_SM_DONE = 2 * 32, BP2_SM_TICKET = 2 * 32 + 1, BP2_SMEM_WORDS = 2 * 32 + 2 };

/*
 * Rules.  K has 3 bits (0..7), the tables are needed at n = K and n = K + 1.  A compile-time rule costs ONE LOP3 per
 * table; the run-time rule walks a mux tree over mask bits that are broadcast to words once per sweep (Tabs).
 * kMono: both masks are upward closed in n (a cell that is alive with n neighbours is alive with n+1) -- then the
 * new alive bit of the in-row predecessor can only turn a dead cell alive, the chain a'(y) = f0 | (f1 & a'(y-1)) is
 * a CARRY chain, and one integer add resolves a whole word (see step()).  The cave-smoothing rules are of that kind.
 */
constexpr bool bp2_upward_closed(uint32_t mask)
{
    for (int n = 0; n < 8; n++)
        if (((mask >> n) & 1u) && !((mask >> (n + 1)) & 1u))
            return false;
    return true;
}

template <uint32_t BORN, uint32_t SURV>
struct Rule2Const {
    static constexpr bool kMono = bp2_upward_closed(BORN) && bp2_upward_closed(SURV);
    struct Tabs { };
    CA_MDEV void setup(const Bp2Params &, Tabs &) { }
    CA_MDEV void eval(const Bp2Params &, const Tabs &, const uint32_t k[3], int,
                      uint32_t &s0, uint32_t &s1, uint32_t &b0, uint32_t &b1)
    {
        s0 = bs_tab3<SURV & 0xFFu>(k[0], k[1], k[2]);
        s1 = bs_tab3<(SURV >> 1) & 0xFFu>(k[0], k[1], k[2]);
        b0 = bs_tab3<BORN & 0xFFu>(k[0], k[1], k[2]);
        b1 = bs_tab3<(BORN >> 1) & 0xFFu>(k[0], k[1], k[2]);
    }
};

struct Rule2Dyn {
    static constexpr bool kMono = false;
    struct Tabs { uint32_t t[4][8]; };
    CA_MDEV void setup(const Bp2Params &p, Tabs &tb)
    {
        const uint32_t m[4] = { p.surv, p.surv >> 1, p.nrval ? p.born : 0u, p.nrval ? p.born >> 1 : 0u };
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int n = 0; n < 8; n++) tb.t[i][n] = bs_bit(m[i], n);
    }
    CA_MDEV uint32_t tab(const uint32_t t[8], const uint32_t k[3], int nb)
    {
        uint32_t a = bs_mux(k[0], t[1], t[0]), b = bs_mux(k[0], t[3], t[2]);
        uint32_t lo = bs_mux(k[1], b, a);
        if (nb < 3)
            return lo;
        uint32_t c = bs_mux(k[0], t[5], t[4]), d = bs_mux(k[0], t[7], t[6]);
        return bs_mux(k[2], bs_mux(k[1], d, c), lo);
    }
    CA_MDEV void eval(const Bp2Params &, const Tabs &tb, const uint32_t k[3], int nb,
                      uint32_t &s0, uint32_t &s1, uint32_t &b0, uint32_t &b1)
    {
        s0 = tab(tb.t[0], k, nb);
        s1 = tab(tb.t[1], k, nb);
        b0 = tab(tb.t[2], k, nb);       /* nr_states == 0: a "born" cell takes the value 0 -- the tables are all zero */
        b1 = tab(tb.t[3], k, nb);
    }
};

template <int P, int WPL, bool MOORE, class Rule>
struct Sweep2 {
    /* register window of state rows x .. x+AHEAD; row x+AHEAD is loaded during step x.  The wide variants keep the
       three-row window of round 1 (the fourth row would spill) */
    enum { SLOTS = (P * WPL >= 12) ? 3 : 4, AHEAD = SLOTS - 1 };

    struct St {
        uint32_t so[SLOTS][P][WPL]; /* state rows, slot = row % 4 */
        uint32_t xo[SLOTS];         /* alive word beyond this warp's span (lane 0: left word, lane 31: right word) */
        uint32_t hn[2][WPL];        /* Moore: H3 of the new row x-1; von Neumann: hn[0] = its alive bits */
        uint32_t vmask[WPL];
        uint32_t *rec;              /* lane-adjusted record of the current row */
        const uint32_t *xrec;       /* lane 0 / 31: the word just outside the warp's span (row x), else null */
        const int *flagp;           /* lane 0: counter of the previous generation, else null */
        int have;
        int next_raise;             /* next row count at which thread 0 raises the shared-memory row counter */
        typename Rule::Tabs tabs;   /* run-time rule: the mask bits, broadcast to words once per sweep */
    };

    CA_MDEV void wait_rows(const Bp2Params &p, St &st, int need)
    {
        if (st.have >= need)
            return;
        long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            int v = st.flagp ? dp_ld_acquire(st.flagp) : 0x7fffffff;
            st.have = dp_reduce_min(v);
            if (st.have >= need)
                break;
            if (spins == 0) t0 = dp_clock();
            dp_nanosleep(20);
            if ((spins & 127u) == 127u) {
                bool bad = dp_ld_flag(p.err) != 0 || (dp_clock() - t0) > p.spin_limit;
                if (!dp_all(!bad)) {
                    if (dp_lane() == 0)
                        dp_atomic_max(p.err, 1);
                    /* keep going: every warp of the CTA must reach the row barriers; the claim loop exits on err */
                    st.have = 0x7fffffff;
                    break;
                }
            }
        }
        dp_syncwarp();
    }

    /* state planes (and the outside word) of the row D rows after the current one into slot S */
    template <int D, int S>
    CA_MDEV void load_row(St &st, int RWS)
    {
        const size_t recw = (size_t)P * RWS;
#pragma unroll
        for (int q = 0; q < P; q++)
            LaneVec<WPL>::ld(st.rec + D * recw + (size_t)q * RWS, st.so[S][q]);
        uint32_t x = 0u;
        if (st.xrec) {
#pragma unroll
            for (int q = 0; q < P; q++)
                x |= dp_ld_cg(st.xrec + D * recw + (size_t)q * RWS);
        }
        st.xo[S] = x;
    }
    template <int S>
    CA_MDEV void zero_row(St &st)
    {
#pragma unroll
        for (int q = 0; q < P; q++)
#pragma unroll
            for (int j = 0; j < WPL; j++) st.so[S][q][j] = 0u;
        st.xo[S] = 0u;
    }

    template <int S>
    CA_MDEV void alive(const St &st, uint32_t a[WPL])
    {
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            uint32_t v = st.so[S][0][j];
#pragma unroll
            for (int q = 1; q < P; q++) v |= st.so[S][q][j];
            a[j] = v;
        }
    }

    /* words to the left / right of each of the lane's words: lanes via shuffle, warp edges via xo */
    CA_MDEV void neighbours(const uint32_t a[WPL], uint32_t xo, uint32_t &prev, uint32_t &next)
    {
        const int lane = dp_lane();
        prev = dp_shfl_up(a[WPL - 1], 1);
        next = dp_shfl_down(a[0], 1);
        if (lane == 0)  prev = xo;
        if (lane == 31) next = xo;
    }

    /* M = x % SLOTS (slot of the current row) */
    template <int M>
    CA_MDEV void step(const Bp2Params &p, St &st, int x, uint32_t *smem)
    {
        constexpr int B = M, C = (M + 1) % SLOTS, A = (M + AHEAD) % SLOTS;
        const int lane = dp_lane(), warp = dp_warp_in_block(), nw = (dp_block_threads() >> 5) - 1;
        const int RWS = p.RWS;

        /* ---- neighbour count K (everything but the in-row predecessor) ---- */
        uint32_t ao[WPL], a1[WPL], ge2[WPL], k[WPL][3];
        alive<B>(st, ao);
        alive<C>(st, a1);
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            uint32_t hi = 0u;
#pragma unroll
            for (int q = 1; q < P; q++) hi |= st.so[B][q][j];
            ge2[j] = hi;
        }
        uint32_t prev0, next0, prev1, next1;
        neighbours(ao, st.xo[B], prev0, next0);
        if (MOORE)
            neighbours(a1, st.xo[C], prev1, next1);
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            uint32_t right0 = (j + 1 < WPL) ? ao[j + 1] : next0;
            uint32_t r = dp_funnel_r(ao[j], right0, 1);             /* old alive bit of y+1, this row */
            if (MOORE) {
                uint32_t left1 = j ? a1[j - 1] : prev1, right1 = (j + 1 < WPL) ? a1[j + 1] : next1;
                uint32_t l = dp_funnel_l(left1, a1[j], 1), rr = dp_funnel_r(a1[j], right1, 1);
                uint32_t ho0 = bs_xor3(l, a1[j], rr), ho1 = bs_maj3(l, a1[j], rr);     /* H3 of old row x+1 */
                k[j][0] = bs_xor3(st.hn[0][j], ho0, r);
                uint32_t c0 = bs_maj3(st.hn[0][j], ho0, r);
                k[j][1] = bs_xor3(st.hn[1][j], ho1, c0);
                k[j][2] = bs_maj3(st.hn[1][j], ho1, c0);
            } else {
                k[j][0] = bs_xor3(st.hn[0][j], a1[j], r);
                k[j][1] = bs_maj3(st.hn[0][j], a1[j], r);
                k[j][2] = 0u;
            }
        }

        /* ---- prefetch row x+3 into the slot row x-1 used to occupy ---- */
        if (x + AHEAD < p.M) {
            wait_rows(p, st, x + AHEAD + 1 < p.M ? x + AHEAD + 1 : p.M);
            load_row<AHEAD, A>(st, RWS);
        } else {
            zero_row<A>(st);
        }

        /* ---- rule tables; the map from the carry INTO this warp's span to every cell ---- */
        uint32_t s0[WPL], s1[WPL], b0[WPL], b1[WPL], f0[WPL], f1[WPL];
        const int nb = MOORE ? 3 : 2;
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            Rule::eval(p, st.tabs, k[j], nb, s0[j], s1[j], b0[j], b1[j]);
            /* new alive bit if the predecessor's new alive bit is 0 / 1 */
            f0[j] = bs_mux(ao[j], s0[j] | ge2[j], b0[j]) & st.vmask[j];
            f1[j] = bs_mux(ao[j], s1[j] | ge2[j], b1[j]) & st.vmask[j];
        }
        uint32_t an[WPL], pred[WPL];    /* new alive bits of y and of y-1 */
        uint32_t cin_lane, first_next;
        uint32_t *slot = smem + BP2_SM_SLOT + (x & 3) * 2;
        if constexpr (Rule::kMono) {
            /*
             * Monotone rule: f0 is a subset of f1, a'(y) = f0 | (f1 & a'(y-1)) -- generate / propagate, exactly the
             * carry chain of f0 + f1.  Lane level: carry out of the lane's words for carry-in 0 (g) and whether
             * every cell propagates (pr); warp level: the same add on the ballots; CTA level: once more on the
             * warps' (generate, propagate) masks.
             */
            uint32_t c = 0u, pr = ~0u;
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                const unsigned long long t = (unsigned long long)f0[j] + f1[j] + c;
                c = (uint32_t)(t >> 32);
                pr &= f0[j] ^ f1[j];
            }
            const uint32_t BG = dp_ballot(c != 0u), BP = dp_ballot(pr == ~0u);
            const unsigned long long tw = (unsigned long long)BG + (BG | BP);
            if (lane == 0) {
                dp_atomic_or_cta(slot, ((uint32_t)(tw >> 32) << warp) | ((uint32_t)(BP == ~0u) << (16 + warp)));
                dp_atomic_or_cta(slot + 1, ((f0[0] & 1u) << warp) | ((f1[0] & 1u) << (16 + warp)));
            }
            dp_syncblock_named(1, nw * 32);
            after_barrier(p, st, x, smem);
            const uint32_t w0 = dp_ld_volatile_u32(slot), w1 = dp_ld_volatile_u32(slot + 1);
            const uint32_t WG = w0 & 0xffffu, WP = w0 >> 16;        /* warps beyond the last: kill */
            const uint32_t cw = (WG + (WG | WP)) ^ WG ^ (WG | WP);  /* bit w = carry into warp w (nw <= 16: no overflow) */
            const uint32_t cin_w = (cw >> warp) & 1u, cout_w = (cw >> (warp + 1)) & 1u;
            /* new alive bit of the first cell of the next warp (its bits are 0 beyond the row) */
            first_next = ((((w1 & 0xffffu) >> (warp + 1)) | (((w1 >> 16) >> (warp + 1)) & cout_w)) & 1u);
            const uint32_t cb = (uint32_t)(tw + cin_w) ^ BG ^ (BG | BP);   /* bit l = carry into lane l */
            c = cin_lane = (cb >> lane) & 1u;
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                const unsigned long long t = (unsigned long long)f0[j] + f1[j] + c;
                pred[j] = (uint32_t)t ^ f0[j] ^ f1[j];          /* bit i = carry into bit i = a'(y-1) */
                c = (uint32_t)(t >> 32);
                an[j] = (pred[j] >> 1) | (c << 31);
            }
        } else {
            /* general rule: GF(2) affine maps, Kogge-Stone inside a word, on the warp's ballots, on the warps' maps */
            uint32_t D[WPL], Cc[WPL];
            uint32_t dl = 1u, cl = 0u;
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                D[j] = f0[j] ^ f1[j];
                Cc[j] = f0[j];
                bs_scan_word(D[j], Cc[j]);
                uint32_t d = D[j] >> 31, c = Cc[j] >> 31;
                cl = c ^ (d & cl);
                dl = d & dl;
            }
            uint32_t BD = dp_ballot(dl != 0), BC = dp_ballot(cl != 0);
            bs_scan_word(BD, BC);
            /* this lane's carry-in is c_in0 ^ (p_in & warp carry-in) */
            const uint32_t c_in0 = lane ? (BC >> (lane - 1)) & 1u : 0u;
            const uint32_t p_in = lane ? (BD >> (lane - 1)) & 1u : 1u;
            if (lane == 0) {    /* word 0: the warp's map, word 1: the map of its first cell */
                dp_atomic_or_cta(slot, ((BD >> 31) << warp) | ((BC >> 31) << (16 + warp)));
                dp_atomic_or_cta(slot + 1, ((D[0] & 1u) << warp) | ((Cc[0] & 1u) << (16 + warp)));
            }
            dp_syncblock_named(1, nw * 32);
            after_barrier(p, st, x, smem);
            const uint32_t w0 = dp_ld_volatile_u32(slot), w1 = dp_ld_volatile_u32(slot + 1);
            uint32_t WD = (w0 & 0xffffu) | (~0u << nw), WC = w0 >> 16;     /* identity map beyond the last warp */
            bs_scan_word(WD, WC);
            const uint32_t cin_w = warp ? (WC >> (warp - 1)) & 1u : 0u;
            const uint32_t cout_w = (WC >> warp) & 1u;
            /* new alive bit of the first cell of the next warp (0 beyond the row) */
            first_next = ((((w1 >> 16) >> (warp + 1)) ^ (((w1 & 0xffffu) >> (warp + 1)) & cout_w)) & 1u);
            uint32_t cin = c_in0 ^ (p_in & cin_w);
            cin_lane = cin;
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                uint32_t cm = 0u - cin;
                an[j] = Cc[j] ^ (D[j] & cm);
                pred[j] = (an[j] << 1) | cin;
                cin = an[j] >> 31;
            }
        }

        /* ---- apply: state planes ---- */
        if constexpr (P == 1) {
            /* one state plane: values are 0 / 1 (nr_states is 1 -- or 0 / 256, and then nothing is ever born), so the
               new state IS the new alive bit */
#pragma unroll
            for (int j = 0; j < WPL; j++) st.so[B][0][j] = an[j];
        } else {
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                uint32_t sv = bs_mux(pred[j], s1[j], s0[j]);
                uint32_t bn = bs_mux(pred[j], b1[j], b0[j]);
                uint32_t dec = ao[j] & ~sv;                      /* alive, neither surviving nor exempt: value - 1 */
                uint32_t brn = ~ao[j] & bn & st.vmask[j];        /* dead, born: value = nr_states */
                uint32_t borrow = dec;
#pragma unroll
                for (int q = 0; q < P; q++) {
                    uint32_t t = st.so[B][q][j];
                    st.so[B][q][j] = t ^ borrow;
                    borrow &= ~t;
                    if ((p.nrval >> q) & 1u) st.so[B][q][j] |= brn;
                }
            }
        }

        /* ---- store row x ---- */
#pragma unroll
        for (int q = 0; q < P; q++)
            LaneVec<WPL>::st(st.rec + (size_t)q * RWS, st.so[B][q]);
        st.rec += (size_t)P * RWS;
        if (st.xrec) st.xrec += (size_t)P * RWS;

        /* ---- what the next row needs from this one ---- */
        if (MOORE) {
            uint32_t nx = dp_shfl_down(an[0], 1);
            if (lane == 31) nx = first_next;
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                /* bit i of l = a'(y-1): the predecessor chain again, word by word */
                uint32_t lcarry = j ? an[j - 1] >> 31 : cin_lane;
                uint32_t l = (an[j] << 1) | lcarry;
                uint32_t right = (j + 1 < WPL) ? an[j + 1] : nx;
                uint32_t r = dp_funnel_r(an[j], right, 1);
                st.hn[0][j] = bs_xor3(l, an[j], r);
                st.hn[1][j] = bs_maj3(l, an[j], r);
            }
        } else {
#pragma unroll
            for (int j = 0; j < WPL; j++) st.hn[0][j] = an[j];
        }
    }

    /*
     * Right after the row barrier of row x: every warp's stores of rows < x precede it.  Thread 0 clears the slot of
     * row x+2 and, every flag_rows rows, says so in shared memory (CTA-scope release) for the publisher warp.  The row
     * marks are counted, not computed: a run-time x % flag_rows is an integer division, ~25 dependent instructions in
     * ONE warp of the CTA, and that warp was late for every row barrier.
     */
    CA_MDEV void after_barrier(const Bp2Params &p, St &st, int x, uint32_t *smem)
    {
        const bool mark = x == st.next_raise;
        if (mark)
            st.next_raise += p.flag_rows;
        if (dp_thread() == 0) {
            uint32_t *clr = smem + BP2_SM_SLOT + ((x + 2) & 3) * 2;
            dp_st_volatile((int *)clr, 0);
            dp_st_volatile((int *)clr + 1, 0);
            if (mark) {
                dp_fence_cta();
                dp_st_volatile((int *)smem + BP2_SM_DONE, x);
            }
        }
    }

    /* all rows of generation g, this warp's span */
    CA_MDEV void sweep(const Bp2Params &p, int g, uint32_t *smem)
    {
        const int lane = dp_lane(), warp = dp_warp_in_block(), nw = (dp_block_threads() >> 5) - 1;
        const int RWS = p.RWS, M = p.M;
        const int word0 = (warp * 32 + lane) * WPL;
        St st;
        st.rec = p.rows + word0;
        st.xrec = (lane == 0 && warp > 0) ? p.rows + word0 - 1
                : ((lane == 31 && warp + 1 < nw) ? p.rows + word0 + WPL : nullptr);
        st.flagp = (g > 0 && lane == 0) ? p.prog + (g - 1) : nullptr;
        st.have = g > 0 ? 0 : 0x7fffffff;
        st.next_raise = p.flag_rows;
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            st.vmask[j] = bp_valid_mask(word0 + j, p.N);
            st.hn[0][j] = st.hn[1][j] = 0u;
        }
        Rule::setup(p, st.tabs);
        wait_rows(p, st, AHEAD < M ? AHEAD : M);
        load_row<0, 0>(st, RWS);
        if (M > 1) load_row<1, 1>(st, RWS); else zero_row<1>(st);
        int x = 0;
        if constexpr (SLOTS == 4) {
            if (M > 2) load_row<2, 2>(st, RWS); else zero_row<2>(st);
            zero_row<3>(st);
            for (; x + 4 <= M; x += 4) {
                step<0>(p, st, x, smem);
                step<1>(p, st, x + 1, smem);
                step<2>(p, st, x + 2, smem);
                step<3>(p, st, x + 3, smem);
            }
            if (x < M) { step<0>(p, st, x, smem); x++; }
            if (x < M) { step<1>(p, st, x, smem); x++; }
            if (x < M) { step<2>(p, st, x, smem); x++; }
        } else {
            zero_row<2>(st);
            for (; x + 3 <= M; x += 3) {
                step<0>(p, st, x, smem);
                step<1>(p, st, x + 1, smem);
                step<2>(p, st, x + 2, smem);
            }
            if (x < M) { step<0>(p, st, x, smem); x++; }
            if (x < M) { step<1>(p, st, x, smem); x++; }
        }
        dp_syncblock_named(1, nw * 32);     /* the last row's stores of every warp */
        if (dp_thread() == 0) {
            dp_fence_cta();
            dp_st_volatile((int *)smem + BP2_SM_DONE, M);
        }
    }

    /* the publisher warp: carries the CTA's row counter to prog[g]; the only gpu-scope fences of the CTA */
    CA_MDEV void publish(const Bp2Params &p, int g, const uint32_t *smem)
    {
        const int lane = dp_lane();
        int pub = 0;
        long long t_idle = dp_clock();
        for (unsigned spins = 0; pub < p.M; spins++) {
            /* one lane looks, everybody follows: the lanes must agree on when the loop ends */
            const int d = (int)dp_shfl(lane == 0 ? (uint32_t)dp_ld_volatile((const int *)smem + BP2_SM_DONE) : 0u, 0);
            if (d > pub) {
                dp_fence_cta();
                dp_fence_release();         /* fence.acq_rel.gpu: rows before the counter (cumulative over the CTA-scope release) */
                if (lane == 0)
                    dp_st_flag(p.prog + g, d);
                pub = d;
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

    CA_MDEV void kernel_body(const Bp2Params &p, uint32_t *smem)
    {
        const int warp = dp_warp_in_block(), nw = (dp_block_threads() >> 5) - 1;
        for (;;) {
            if (dp_thread() == 0) {
                unsigned t = dp_atomic_inc(p.ticket);
                if (dp_ld_flag(p.err) != 0)
                    t = 0xffffffffu;
                smem[BP2_SM_TICKET] = t;
            }
            for (int i = dp_thread(); i < BP2_SM_TICKET; i += dp_block_threads())
                smem[i] = 0u;                       /* carry maps and the row counter restart with every generation */
            dp_syncblock();
            const unsigned g = smem[BP2_SM_TICKET];
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

/* the compile-time rules: BASELINE config 3's cave smoothing (born 5..8, survive 4..8) and ca_test (terrain.c:391-398) */
typedef Rule2Const<0x1E0u, 0x1F0u> Rule2Cave;
typedef Rule2Const<0x00Cu, 0x180u> Rule2Test;
enum { BP2_RULE_DYN = 0, BP2_RULE_CAVE = 1, BP2_RULE_TEST = 2, BP2_NRULES = 3 };

/* which instantiation serves (born, surv, nrval) -- surv is already all ones for a rule that does not decay */
inline int bp2_rule_for(uint32_t born, uint32_t surv, uint32_t nrval)
{
    if (!nrval) return BP2_RULE_DYN;            /* births are no-ops: only the run-time tables model that */
    if ((born & 0x1ffu) == 0x1E0u && (surv & 0x1ffu) == 0x1F0u) return BP2_RULE_CAVE;
    if ((born & 0x1ffu) == 0x00Cu && (surv & 0x1ffu) == 0x180u) return BP2_RULE_TEST;
    return BP2_RULE_DYN;
}

template <int P, int WPL, bool MOORE, class Rule>
CA_GLOBAL void __launch_bounds__(32 * (BP2_MAX_WARPS + 1), 1) ca2d_sweep_kernel(Bp2Params p)
{
    CA_SHARED(uint32_t, smem, BP2_SMEM_WORDS);
    Sweep2<P, WPL, MOORE, Rule>::kernel_body(p, smem);
}

} // namespace clapca
#endif
