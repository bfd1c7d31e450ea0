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
 * Parallelism.  One CTA sweeps one generation: warp w / lane l owns WPL consecutive words of the row; the
 * scan is resolved word -> warp (ballots) -> CTA (one packed word per warp in shared memory, one
 * __syncthreads per row).  Generation g follows generation g-1 two rows behind (row x needs rows <= x+1
 * of the previous generation), so all G generations are in flight in ONE launch, each publishing a
 * progress counter prog[g] = rows completed (st.release after the CTA barrier) that the next generation
 * polls (ld.acquire).  Storage is single-buffered, like the reference's: a row is overwritten only after
 * its last reader has passed.  The whole working set (a few rows per generation) lives in L2.
 */
#ifndef CLAPCA_CA2D_BITPLANE_CUH
#define CLAPCA_CA2D_BITPLANE_CUH

#include "bitslice.cuh"
#include "ca3d_bitplane.cuh"        /* LaneVec, bp_valid_mask */

namespace clapca {

struct Bp2Params {
    uint32_t *rows;         /* [M][P][RWS] */
    int N, M, G;            /* cells per row (reference y extent), rows (reference x extent), generations */
    int RWS;                /* words per plane-row = warps per CTA * 32 * WPL */
    int *prog;              /* [G] rows completed by generation g */
    unsigned *ticket;       /* next generation to claim */
    int *err;               /* != 0: watchdog fired, everybody bails out */
    uint32_t born, surv;    /* 9-bit masks; surv is all ones when the rule does not decay */
    uint32_t nrval;         /* (uint8_t)nr_states: the value a born cell takes (core/ca2d.c:71) */
    int flag_rows;          /* the progress counter is raised every flag_rows rows */
    long long spin_limit;
};

enum { BP2_SMEM_WORDS = 2 * 32 + 2 };

template <int P, int WPL, bool MOORE>
struct Sweep2 {
    struct St {
        uint32_t so[3][P][WPL];     /* state rows x, x+1, x+2; slot = row % 3 */
        uint32_t xo[3];             /* alive word beyond this warp's span (lane 0: left word, lane 31: right word) */
        uint32_t hn[2][WPL];        /* Moore: H3 of the new row x-1; von Neumann: hn[0] = its alive bits */
        uint32_t vmask[WPL];
        uint32_t *rec;              /* lane-adjusted record of the current row */
        const uint32_t *xrec;       /* lane 0 / 31: the word just outside the warp's span (row x), else null */
        int have;
    };

    CA_MDEV bool wait_rows(const Bp2Params &p, St &st, const int *flag, int need)
    {
        if (st.have >= need)
            return true;
        long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            int v = dp_lane() == 0 ? dp_ld_acquire(flag) : 0x7fffffff;
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
                    /* keep going: every warp of the CTA must reach the barriers; the claim loop exits on err */
                    st.have = 0x7fffffff;
                    break;
                }
            }
        }
        dp_syncwarp();
        return true;
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

    /* M = x % 3 (slot of the current row) */
    template <int M>
    CA_MDEV void step(const Bp2Params &p, St &st, int x, const int *flag_prev, int *myprog, uint32_t *smem,
                      int &next_raise)
    {
        constexpr int B = M, C = (M + 1) % 3, A = (M + 2) % 3;
        const int lane = dp_lane(), warp = dp_warp_in_block(), nw = dp_block_threads() >> 5;
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

        /* ---- prefetch row x+2 into the slot row x-1 used to occupy ---- */
        if (x + 2 < p.M) {
            if (flag_prev)
                wait_rows(p, st, flag_prev, x + 3 < p.M ? x + 3 : p.M);
            load_row<2, A>(st, RWS);
        } else {
            zero_row<A>(st);
        }

        /* ---- rule tables, word scan ---- */
        uint32_t s0[WPL], s1[WPL], b0[WPL], b1[WPL], D[WPL], Cc[WPL];
        uint32_t dl = 1u, cl = 0u;
        const int nb = MOORE ? 3 : 2;
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            s0[j] = bs_tab_dyn(p.surv, k[j], nb);
            s1[j] = bs_tab_dyn(p.surv >> 1, k[j], nb);
            if (p.nrval) {
                b0[j] = bs_tab_dyn(p.born, k[j], nb);
                b1[j] = bs_tab_dyn(p.born >> 1, k[j], nb);
            } else {
                b0[j] = b1[j] = 0u;     /* a "born" cell takes the value 0: nothing happens */
            }
            uint32_t f0 = bs_mux(ao[j], s0[j] | ge2[j], b0[j]) & st.vmask[j];
            uint32_t f1 = bs_mux(ao[j], s1[j] | ge2[j], b1[j]) & st.vmask[j];
            D[j] = f0 ^ f1;
            Cc[j] = f0;
            bs_scan_word(D[j], Cc[j]);
            uint32_t d = D[j] >> 31, c = Cc[j] >> 31;
            cl = c ^ (d & cl);
            dl = d & dl;
        }

        /* ---- warp scan, then the CTA scan through shared memory ---- */
        uint32_t BD = dp_ballot(dl != 0), BC = dp_ballot(cl != 0);
        bs_scan_word(BD, BC);
        /* this lane's carry-in is c_in0 ^ (p_in & warp carry-in) */
        const uint32_t c_in0 = lane ? (BC >> (lane - 1)) & 1u : 0u;
        const uint32_t p_in = lane ? (BD >> (lane - 1)) & 1u : 1u;
        uint32_t *slot = smem + (x & 1) * 32;
        if (lane == 0)      /* bit 0/1: the warp's map, bit 2/3: the map of its first cell */
            slot[warp] = (BD >> 31) | ((BC >> 31) << 1) | ((D[0] & 1u) << 2) | ((Cc[0] & 1u) << 3);
        dp_syncblock();
        /* every warp's stores of row x-1 precede this barrier: the counter may now say x rows are done */
        if (x == next_raise) {
            next_raise += p.flag_rows;
            if (dp_thread() == 0)
                dp_st_release(myprog, x);
        }
        const uint32_t mine = lane < nw ? slot[lane] : 1u;      /* identity map beyond the last warp */
        uint32_t WD = dp_ballot((mine & 1u) != 0), WC = dp_ballot((mine & 2u) != 0);
        bs_scan_word(WD, WC);
        const uint32_t cin_w = warp ? (WC >> (warp - 1)) & 1u : 0u;
        const uint32_t cout_w = (WC >> warp) & 1u;
        /* new alive bit of the first cell of the next warp (0 beyond the row) */
        const uint32_t nxt = dp_shfl(mine, warp + 1 < 32 ? warp + 1 : 31);
        const uint32_t first_next = (warp + 1 < nw) ? (((nxt >> 3) ^ ((nxt >> 2) & cout_w)) & 1u) : 0u;
        uint32_t cin = c_in0 ^ (p_in & cin_w);
        const uint32_t cin_lane = cin;

        /* ---- apply: new alive bits, state planes ---- */
        uint32_t an[WPL];
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            uint32_t cm = 0u - cin;
            an[j] = Cc[j] ^ (D[j] & cm);
            uint32_t pred = (an[j] << 1) | cin;              /* new alive bit of y-1 */
            cin = an[j] >> 31;
            uint32_t sv = bs_mux(pred, s1[j], s0[j]);
            uint32_t bn = bs_mux(pred, b1[j], b0[j]);
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

        /* ---- store row x ---- */
#pragma unroll
        for (int q = 0; q < P; q++)
            LaneVec<WPL>::st(st.rec + (size_t)q * RWS, st.so[B][q]);
        st.rec += (size_t)P * RWS;
        if (st.xrec) st.xrec += (size_t)P * RWS;
    }

    /* all rows of generation g */
    CA_MDEV void sweep(const Bp2Params &p, int g, uint32_t *smem)
    {
        const int lane = dp_lane(), warp = dp_warp_in_block(), nw = dp_block_threads() >> 5;
        const int RWS = p.RWS, M = p.M;
        const int word0 = (warp * 32 + lane) * WPL;
        int *myprog = p.prog + g;
        const int *flag_prev = g > 0 ? p.prog + g - 1 : nullptr;
        St st;
        st.rec = p.rows + word0;
        st.xrec = (lane == 0 && warp > 0) ? p.rows + word0 - 1
                : ((lane == 31 && warp + 1 < nw) ? p.rows + word0 + WPL : nullptr);
        st.have = flag_prev ? 0 : 0x7fffffff;
#pragma unroll
        for (int j = 0; j < WPL; j++) {
            st.vmask[j] = bp_valid_mask(word0 + j, p.N);
            st.hn[0][j] = st.hn[1][j] = 0u;
        }
        if (flag_prev)
            wait_rows(p, st, flag_prev, 2 < M ? 2 : M);
        load_row<0, 0>(st, RWS);
        if (M > 1) load_row<1, 1>(st, RWS);
        else zero_row<1>(st);
        zero_row<2>(st);

        int next_raise = p.flag_rows;
        int x = 0;
        for (; x + 3 <= M; x += 3) {
            step<0>(p, st, x, flag_prev, myprog, smem, next_raise);
            step<1>(p, st, x + 1, flag_prev, myprog, smem, next_raise);
            step<2>(p, st, x + 2, flag_prev, myprog, smem, next_raise);
        }
        if (x < M) { step<0>(p, st, x, flag_prev, myprog, smem, next_raise); x++; }
        if (x < M) { step<1>(p, st, x, flag_prev, myprog, smem, next_raise); x++; }
        dp_syncblock();             /* the last row's stores of every warp */
        if (dp_thread() == 0)
            dp_st_release(myprog, M);
    }

    CA_MDEV void kernel_body(const Bp2Params &p, uint32_t *smem)
    {
        for (;;) {
            if (dp_thread() == 0) {
                unsigned t = dp_atomic_inc(p.ticket);
                if (dp_ld_flag(p.err) != 0)
                    t = 0xffffffffu;
                smem[64] = t;
            }
            dp_syncblock();
            const unsigned g = smem[64];
            dp_syncblock();         /* smem[64] may be rewritten by the next claim */
            if (g >= (unsigned)p.G)
                break;
            sweep(p, (int)g, smem);
        }
    }
};

template <int P, int WPL, bool MOORE>
CA_GLOBAL void __launch_bounds__(512) ca2d_sweep_kernel(Bp2Params p)
{
    CA_SHARED(uint32_t, smem, BP2_SMEM_WORDS);
    Sweep2<P, WPL, MOORE>::kernel_body(p, smem);
}

} // namespace clapca
#endif
