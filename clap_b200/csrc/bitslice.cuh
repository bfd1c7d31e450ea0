/*
 * bitslice.cuh -- bit-sliced building blocks shared by the 2D and 3D bit-plane
 * engines: one 32-bit word carries the same bit of 32 neighbouring cells, so a
 * single LOP3 acts on 32 cells at once.
 *
 *   - full/half adders and the small adder trees that turn neighbour alive
 *     bits into a binary neighbour count K (one word per count bit);
 *   - rule-table lookup T[K] on a bit-sliced K (compile-time masks fold into
 *     at most 7 LOP3 per table, run-time masks use a mux tree);
 *   - the in-row dependency solver: the reference sweeps a row in place, so
 *     cell x sees the *new* alive bit of cell x-1 (core/ca3d.c:129-140,
 *     core/ca2d.c:61-77).  With f0/f1 = new alive bit of x given that bit is
 *     0/1, a'(x) = f0(x) ^ (a'(x-1) & (f0(x)^f1(x))) is an affine recurrence
 *     over GF(2); affine maps compose associatively, so a Kogge-Stone scan
 *     (5 shift steps per word, then the same 5 steps on the warp ballots)
 *     resolves 1024*WPL cells in O(log) depth, bit-exactly.
 */
#ifndef CLAPCA_BITSLICE_CUH
#define CLAPCA_BITSLICE_CUH

#include "devport.h"

namespace clapca {

CA_DEV uint32_t bs_xor3(uint32_t a, uint32_t b, uint32_t c) { return dp_lop3<0x96>(a, b, c); }
CA_DEV uint32_t bs_maj3(uint32_t a, uint32_t b, uint32_t c) { return dp_lop3<0xE8>(a, b, c); }
/* s ? a : b, bitwise */
CA_DEV uint32_t bs_mux(uint32_t s, uint32_t a, uint32_t b)  { return dp_lop3<0xCA>(s, a, b); }

/* sum of three 2-bit numbers (each 0..3) -> 4-bit number (0..9) */
CA_DEV void bs_add3x2(const uint32_t a[2], const uint32_t b[2], const uint32_t c[2], uint32_t v[4])
{
    uint32_t s0 = bs_xor3(a[0], b[0], c[0]), c0 = bs_maj3(a[0], b[0], c[0]);
    uint32_t s1 = bs_xor3(a[1], b[1], c[1]), c1 = bs_maj3(a[1], b[1], c[1]);
    uint32_t u = c0 & s1;
    v[0] = s0;
    v[1] = c0 ^ s1;
    v[2] = c1 ^ u;
    v[3] = c1 & u;
}

/*
 * 3D neighbour count without the in-row predecessor:
 *   K = Vdn(0..9) + Vup(0..9) + Hnew(0..3) + Hold(0..3) + r(0..1)   (max 25)
 */
CA_DEV void bs_count3d(const uint32_t d[4], const uint32_t u[4], const uint32_t n[2], const uint32_t o[2],
                       uint32_t r, uint32_t k[5])
{
    /* weight 1: d0 u0 n0 o0 r */
    uint32_t sa = bs_xor3(d[0], u[0], n[0]), ca = bs_maj3(d[0], u[0], n[0]);
    k[0] = bs_xor3(sa, o[0], r);
    uint32_t cb = bs_maj3(sa, o[0], r);
    /* weight 2: d1 u1 n1 o1 ca cb */
    uint32_t sc = bs_xor3(d[1], u[1], n[1]), cc = bs_maj3(d[1], u[1], n[1]);
    uint32_t sd = bs_xor3(o[1], ca, cb),     cd = bs_maj3(o[1], ca, cb);
    k[1] = sc ^ sd;
    uint32_t ce = sc & sd;
    /* weight 4: d2 u2 cc cd ce */
    uint32_t sf = bs_xor3(d[2], u[2], cc), cf = bs_maj3(d[2], u[2], cc);
    k[2] = bs_xor3(sf, cd, ce);
    uint32_t cg = bs_maj3(sf, cd, ce);
    /* weight 8: d3 u3 cf cg */
    uint32_t sh = bs_xor3(d[3], u[3], cf), ch = bs_maj3(d[3], u[3], cf);
    k[3] = sh ^ cg;
    /* weight 16: at most one of the two carries can be set (K <= 25) */
    k[4] = ch | (sh & cg);
}

/* T = a + b of two 2-bit numbers (each 0..3) -> 3-bit number (0..6): 4 LOP3 */
CA_DEV void bs_add2x2(const uint32_t a[2], const uint32_t b[2], uint32_t t[3])
{
    uint32_t c0 = a[0] & b[0];
    t[0] = a[0] ^ b[0];
    t[1] = bs_xor3(a[1], b[1], c0);
    t[2] = bs_maj3(a[1], b[1], c0);
}

/*
 * 3D neighbour count without the in-row predecessor, from the per-row pair sums T(r) = Hdn(r) + Hup(r) (0..6) of rows
 * y-1, y, y+1 -- each T is formed once (bs_add2x2) and serves three consecutive rows:
 *   K = T(y-1) + T(y) + T(y+1) + Hnew(0..3) + Hold(0..3) + r(0..1)   (max 25)
 * as a carry-save tree of 11 full / 2 half adders: 22 LOP3 (+ 4 for the new T) against 34 for two vertical 3-sums
 * and their sum.
 */
CA_DEV void bs_count3d_t(const uint32_t ta[3], const uint32_t tb[3], const uint32_t tc[3], const uint32_t n[2],
                         const uint32_t o[2], uint32_t r, uint32_t k[5])
{
    /* weight 1: ta0 tb0 tc0 n0 o0 r */
    uint32_t s1 = bs_xor3(ta[0], tb[0], tc[0]), c1 = bs_maj3(ta[0], tb[0], tc[0]);
    uint32_t s2 = bs_xor3(n[0], o[0], r),       c2 = bs_maj3(n[0], o[0], r);
    k[0] = s1 ^ s2;
    uint32_t c3 = s1 & s2;
    /* weight 2: ta1 tb1 tc1 n1 o1 c1 c2 c3 */
    uint32_t s4 = bs_xor3(ta[1], tb[1], tc[1]), d1 = bs_maj3(ta[1], tb[1], tc[1]);
    uint32_t s5 = bs_xor3(n[1], o[1], c1),      d2 = bs_maj3(n[1], o[1], c1);
    uint32_t s6 = bs_xor3(s4, c2, c3),          d3 = bs_maj3(s4, c2, c3);
    k[1] = s5 ^ s6;
    uint32_t d4 = s5 & s6;
    /* weight 4: ta2 tb2 tc2 d1 d2 d3 d4 */
    uint32_t s7 = bs_xor3(ta[2], tb[2], tc[2]), e1 = bs_maj3(ta[2], tb[2], tc[2]);
    uint32_t s8 = bs_xor3(d1, d2, d3),          e2 = bs_maj3(d1, d2, d3);
    k[2] = bs_xor3(s7, s8, d4);
    uint32_t e3 = bs_maj3(s7, s8, d4);
    /* weight 8 / 16 */
    k[3] = bs_xor3(e1, e2, e3);
    k[4] = bs_maj3(e1, e2, e3);
}

/* ---- rule tables on a bit-sliced count ----------------------------------- */

/* compile-time 8-entry table over (k2,k1,k0) */
template <unsigned LUT>
CA_DEV uint32_t bs_tab3(uint32_t k0, uint32_t k1, uint32_t k2)
{
    if (LUT == 0x00) return 0u;
    if (LUT == 0xFF) return ~0u;
    return dp_lop3<LUT>(k2, k1, k0);
}

/* compile-time table T[K], K = k0 + 2 k1 + 4 k2 + 8 k3 + 16 k4 (K < 32) */
template <uint32_t MASK>
CA_DEV uint32_t bs_tab5(const uint32_t k[5])
{
    constexpr unsigned L0 = MASK & 0xFF, L1 = (MASK >> 8) & 0xFF, L2 = (MASK >> 16) & 0xFF, L3 = (MASK >> 24) & 0xFF;
    uint32_t lo, hi;
    if (L0 == L1) lo = bs_tab3<L0>(k[0], k[1], k[2]);
    else          lo = bs_mux(k[3], bs_tab3<L1>(k[0], k[1], k[2]), bs_tab3<L0>(k[0], k[1], k[2]));
    if (L0 == L2 && L1 == L3) return lo;
    if (L2 == L3) hi = bs_tab3<L2>(k[0], k[1], k[2]);
    else          hi = bs_mux(k[3], bs_tab3<L3>(k[0], k[1], k[2]), bs_tab3<L2>(k[0], k[1], k[2]));
    return bs_mux(k[4], hi, lo);
}

/* run-time table: mux tree over broadcast mask bits */
CA_DEV uint32_t bs_bit(uint32_t mask, int n) { return 0u - ((mask >> n) & 1u); }

CA_DEV uint32_t bs_tab_dyn(uint32_t mask, const uint32_t *k, int nbits)
{
    /* level 0 folds k[0]; higher levels fold k[1..nbits-1] */
    uint32_t t[16];
    const int n0 = 1 << (nbits - 1);
#pragma unroll
    for (int i = 0; i < 16; i++)
        if (i < n0)
            t[i] = bs_mux(k[0], bs_bit(mask, 2 * i + 1), bs_bit(mask, 2 * i));
#pragma unroll
    for (int lvl = 1; lvl < 5; lvl++)
        if (lvl < nbits) {
            const int n = 1 << (nbits - 1 - lvl);
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (i < n)
                    t[i] = bs_mux(k[lvl], t[2 * i + 1], t[2 * i]);
        }
    return t[0];
}

/* ---- in-row dependency scan ---------------------------------------------- */

/*
 * Word-local Kogge-Stone.  In: per-bit maps a'(x) = C(x) ^ (D(x) & a'(x-1)).
 * Out: per-bit maps from the word's carry-in: a'(x) = C(x) ^ (D(x) & cin).
 */
CA_DEV void bs_scan_word(uint32_t &D, uint32_t &C)
{
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        C ^= D & (C << s);
        D &= (D << s) | ((1u << s) - 1u);
    }
}

/*
 * The same scan on E = ~D ("the chain breaks here"): C ^= ~E & (C << s); E |= E << s.  Both shifts are plain left
 * shifts (IMAD.SHL on the FMA pipe) where the D form needs (D << s) | low-ones -- an LEA on the ALU pipe, the pipe
 * that bounds the sweep kernels: 2 instead of 3 ALU-pipe instructions per step.
 */
CA_DEV void bs_scan_word_e(uint32_t &E, uint32_t &C)
{
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        C = dp_lop3<0xD2>(C, E, C << s);        /* C ^ (~E & (C << s)) */
        E |= E << s;
    }
}

/*
 * Warp-level stage: every lane contributes the map (dl, cl) from its carry-in
 * to its last cell; returns the lane's carry-in bit (0 or 1).  `cut_after`
 * is a lane mask of lanes after which the chain restarts with 0 (row ends).
 */
CA_DEV uint32_t bs_scan_warp(uint32_t dl, uint32_t cl, uint32_t cut_after)
{
    uint32_t BD = dp_ballot(dl != 0) & ~cut_after;
    uint32_t BC = dp_ballot(cl != 0) & ~cut_after;
    bs_scan_word(BD, BC);           /* carry-in of lane 0 is 0: result bit i = BC bit i */
    int lane = dp_lane();
    return lane ? (BC >> (lane - 1)) & 1u : 0u;
}

/*
 * The scan in two parts.  After the shifts by 1, 2, 4 bit i of E is clear only if cells i-7 .. i ALL pass their
 * predecessor's bit on; unless such a run of 8 exists somewhere (bits 8..31 of ~E), the shifts by 8 and 16 change
 * nothing and the caller skips them for the whole warp (one vote).  Exact either way -- the fast path is just the
 * common case: a run of 8 dependent cells is rare in any rule's steady state.
 */
CA_DEV void bs_scan_word_e_lo(uint32_t &E, uint32_t &C)
{
#pragma unroll
    for (int s = 1; s < 8; s <<= 1) {
        C = dp_lop3<0xD2>(C, E, C << s);
        E |= E << s;
    }
}
CA_DEV void bs_scan_word_e_hi(uint32_t &E, uint32_t &C)
{
#pragma unroll
    for (int s = 8; s < 32; s <<= 1) {
        C = dp_lop3<0xD2>(C, E, C << s);
        E |= E << s;
    }
}

/*
 * The warp-level stage on E = ~D; el / cl: this lane's map breaks the chain / emits 1 for carry-in 0.  When every
 * lane breaks the chain (the common case: a lane holds 32 * WPL cells) the carry into lane i is simply cl of lane i-1.
 */
CA_DEV uint32_t bs_scan_warp_e(bool el, bool cl)
{
    uint32_t BE = dp_ballot(el);
    uint32_t BC = dp_ballot(cl);
    if (BE != 0xffffffffu)
        bs_scan_word_e(BE, BC);     /* carry-in of lane 0 is 0: result bit i = BC bit i */
    return ((BC + BC) >> dp_lane()) & 1u;
}

} // namespace clapca
#endif
