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
    int no_fence;           /* experiments only: skip the release fence (NOT a valid configuration) */
    int flag_rows;          /* progress counters are raised every flag_rows rows (and at segment ends) */
    unsigned *ticket;       /* next sweep to claim */
    int *err;               /* != 0: watchdog fired, everybody bails out */
    uint32_t surv, born;    /* rule masks (run-time rule only) */
    uint32_t bornval;       /* (nr_states - 1) & 0xff */
    long long spin_limit;   /* watchdog budget in clock ticks per wait */
};

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

template <int P, int WPL, class Rule>
struct Sweep3 {
    static constexpr int NP = P + 2;

    struct Flags {          /* cached progress of the three producers of a sweep */
        const int *dn, *up, *own, *self;
        int vdn, vup, vown, vself;
    };

    /*
     * Wait until the three producers have completed `need` rows and this sweep's own
     * earlier segments `need_self` rows; false = watchdog fired / abort requested.
     */
    CA_MDEV bool wait_rows(const Bp3Params &p, Flags &f, int need, int need_self)
    {
        if (f.vdn >= need && f.vup >= need && f.vown >= need && f.vself >= need_self)
            return true;
        long long t0 = dp_clock();
        unsigned spins = 0;
        for (;;) {
            /* lanes 0..3 poll one counter each; the values are made warp-uniform by shuffle */
            const int lane = dp_lane();
            const int *src = lane == 0 ? f.dn : (lane == 1 ? f.up : (lane == 2 ? f.own : (lane == 3 ? f.self : nullptr)));
            uint32_t v = src ? (uint32_t)dp_ld_flag(src) : 0x7fffffffu;
            f.vdn   = (int)dp_shfl(v, 0);
            f.vup   = (int)dp_shfl(v, 1);
            f.vown  = (int)dp_shfl(v, 2);
            f.vself = (int)dp_shfl(v, 3);
            if (f.vdn >= need && f.vup >= need && f.vown >= need && f.vself >= need_self)
                break;
            dp_nanosleep(40);
            if ((++spins & 63u) == 0u) {
                bool bad = dp_ld_flag(p.err) != 0 || (dp_clock() - t0) > p.spin_limit;
                if (!dp_all(!bad)) {
                    if (dp_lane() == 0)
                        dp_atomic_max(p.err, 1);
                    return false;
                }
            }
        }
        /* acquire: the polling lanes fence after their relaxed load, the warp barrier extends it to all lanes */
        dp_fence_acquire();
        dp_syncwarp();
        return true;
    }

    struct RowIn {          /* what one row step pulls from memory for row r */
        uint32_t hd[2][WPL], hu[2][WPL], ho[2][WPL];
    };

    CA_MDEV void load_h(const uint32_t *rec, int RWP, int lane, uint32_t h[2][WPL])
    {
        LaneVec<WPL>::ld(rec + 0 * RWP + lane * WPL, h[0]);
        LaneVec<WPL>::ld(rec + 1 * RWP + lane * WPL, h[1]);
    }
    CA_MDEV void zero_h(uint32_t h[2][WPL])
    {
#pragma unroll
        for (int j = 0; j < WPL; j++) h[0][j] = h[1][j] = 0u;
    }
    CA_MDEV void load_s(const uint32_t *rec, int RWP, int lane, uint32_t s[P][WPL])
    {
#pragma unroll
        for (int q = 0; q < P; q++)
            LaneVec<WPL>::ld(rec + (2 + q) * RWP + lane * WPL, s[q]);
    }

    /*
     * Ghost rows (see bp3_types.h): {word, tag} pairs written by the neighbouring GPU.  Re-read until every
     * pair of this lane -- and of the whole warp -- carries the expected tag.  false = watchdog / abort.
     */
    CA_MDEV bool load_h_tagged(const Bp3Params &p, const uint32_t *row, int lane, uint32_t expect, uint32_t h[2][WPL])
    {
        const uint32_t *src = row + (size_t)lane * 4 * WPL;
        long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < WPL; i++) {             /* pairs 2i, 2i+1 of this lane */
                uint4 v = dp_ld_cg(reinterpret_cast<const uint4 *>(src) + i);
                const int q0 = 2 * i, q1 = 2 * i + 1;   /* pair index -> (plane, word) = (q / WPL, q % WPL) */
                h[q0 / WPL][q0 % WPL] = v.x;
                h[q1 / WPL][q1 % WPL] = v.z;
                ok = ok && v.y == expect && v.w == expect;
            }
            if (dp_all(ok))
                return true;
            if (spins == 0) t0 = dp_clock();
            dp_nanosleep(100);
            if ((spins & 63u) == 63u) {
                bool bad = dp_ld_flag(p.err) != 0 || (dp_clock() - t0) > p.spin_limit;
                if (!dp_all(!bad)) {
                    if (dp_lane() == 0)
                        dp_atomic_max(p.err, 2);
                    return false;
                }
            }
        }
    }

    CA_MDEV void store_h_tagged(uint32_t *row, int lane, uint32_t tag, const uint32_t h0[WPL], const uint32_t h1[WPL])
    {
        uint32_t *dst = row + (size_t)lane * 4 * WPL;
#pragma unroll
        for (int i = 0; i < WPL; i++) {
            const int q0 = 2 * i, q1 = 2 * i + 1;
            const uint32_t a = (q0 / WPL) ? h1[q0 % WPL] : h0[q0 % WPL];
            const uint32_t b = (q1 / WPL) ? h1[q1 % WPL] : h0[q1 % WPL];
            dp_st_cg(reinterpret_cast<uint4 *>(dst) + i, make_uint4(a, tag, b, tag));
        }
    }

    /* rows r of the three sources of a sweep: H of dn/up/own-old and S of own-old.  false = aborted */
    CA_MDEV bool load_row(const Bp3Params &p, const Bp3Plane &pl, int z, int g, int r, int lane, RowIn &in,
                          uint32_t so[P][WPL])
    {
        const size_t recw = (size_t)NP * p.RWP;
        if (r < p.H) {
            const uint32_t *own = p.rows + ((size_t)z * p.H + r) * recw;
            load_h(own, p.RWP, lane, in.ho);
            load_s(own, p.RWP, lane, so);
            if (!pl.dn_rows) zero_h(in.hd);
            else if (!(pl.ghost_mask & 1u)) load_h(pl.dn_rows + (size_t)r * pl.dn_stride, p.RWP, lane, in.hd);
            else if (!load_h_tagged(p, pl.dn_rows + (size_t)r * pl.dn_stride, lane,
                                    (p.epoch << 16) | (uint32_t)(g + 1), in.hd)) return false;
            if (!pl.up_rows) zero_h(in.hu);
            else if (!(pl.ghost_mask & 2u)) load_h(pl.up_rows + (size_t)r * pl.up_stride, p.RWP, lane, in.hu);
            else if (!load_h_tagged(p, pl.up_rows + (size_t)r * pl.up_stride, lane,
                                    (p.epoch << 16) | (uint32_t)g, in.hu)) return false;
        } else {
            zero_h(in.hd); zero_h(in.hu); zero_h(in.ho);
#pragma unroll
            for (int q = 0; q < P; q++)
#pragma unroll
                for (int j = 0; j < WPL; j++) so[q][j] = 0u;
        }
        return true;
    }

    /* one work item: rows [y0, y1) of plane z at generation g.  false = aborted */
    CA_MDEV bool run_segment(const Bp3Params &p, int z, int g, int y0, int y1)
    {
        const int lane = dp_lane();
        const int H = p.H, Z = p.Z;
        const Bp3Plane pl = p.planes[z];
        const uint32_t bornval = Rule::bornval(p);
        const size_t recw = (size_t)NP * p.RWP;
        int *myprog = p.prog + (size_t)g * Z + z;

        Flags f;
        /* ghost sources are synchronised by their row tags, not by counters */
        f.dn  = (pl.dn_rows && !(pl.ghost_mask & 1u)) ? pl.dn_flag + (size_t)g * pl.dn_gstride : nullptr;
        f.up  = (g > 0 && pl.up_rows && !(pl.ghost_mask & 2u)) ? pl.up_flag + (size_t)(g - 1) * pl.up_gstride : nullptr;
        f.own = g > 0 ? p.prog + (size_t)(g - 1) * Z + z : nullptr;
        f.vdn  = f.dn  ? 0 : 0x7fffffff;
        f.vup  = f.up  ? 0 : 0x7fffffff;
        f.vown = f.own ? 0 : 0x7fffffff;
        f.self = y0 > 0 ? myprog : nullptr;         /* rows < y0 belong to earlier segments of this sweep */
        f.vself = f.self ? 0 : 0x7fffffff;

        uint32_t vmask[WPL];
#pragma unroll
        for (int j = 0; j < WPL; j++) vmask[j] = bp_valid_mask(lane * WPL + j, p.W);

        /* sliding windows (see file header) */
        uint32_t hdA[2][WPL], hdB[2][WPL], huA[2][WPL], huB[2][WPL], hn[2][WPL];
        uint32_t so[P][WPL], so1[P][WPL], so2[P][WPL];
        RowIn in;

        if (!wait_rows(p, f, y0 + 2 < H ? y0 + 2 : H, y0))
            return false;
        {
            /* (re)build the sliding windows: rows y0-1, y0 and y0+1 of the sources */
            RowIn r0;
            if (y0 > 0) {
                if (!load_row(p, pl, z, g, y0 - 1, lane, r0, so)) return false;
#pragma unroll
                for (int b = 0; b < 2; b++)
#pragma unroll
                    for (int j = 0; j < WPL; j++) {
                        hdA[b][j] = r0.hd[b][j]; huA[b][j] = r0.hu[b][j];
                        hn[b][j] = r0.ho[b][j];     /* row y0-1 of this plane is already generation g */
                    }
            } else {
                zero_h(hdA); zero_h(huA); zero_h(hn);
            }
            if (!load_row(p, pl, z, g, y0, lane, r0, so)) return false;
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int j = 0; j < WPL; j++) { hdB[b][j] = r0.hd[b][j]; huB[b][j] = r0.hu[b][j]; }
            if (!load_row(p, pl, z, g, y0 + 1, lane, in, so1)) return false;
        }

        for (int y = y0; y < y1; y++) {
            uint32_t k[WPL][5], ao[WPL], ge2[WPL];

            /* ---- neighbour count K (everything but the in-row predecessor) ---- */
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                uint32_t a = so[0][j], hi = 0u;
#pragma unroll
                for (int q = 1; q < P; q++) hi |= so[q][j];
                ao[j] = a | hi;
                ge2[j] = hi;
            }
            uint32_t nxt = dp_shfl_down(ao[0], 1);
            if (lane == 31) nxt = 0u;
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                uint32_t a_[2] = { hdA[0][j], hdA[1][j] }, b_[2] = { hdB[0][j], hdB[1][j] };
                uint32_t c_[2] = { in.hd[0][j], in.hd[1][j] };
                uint32_t vd[4], vu[4];
                bs_add3x2(a_, b_, c_, vd);
                uint32_t e_[2] = { huA[0][j], huA[1][j] }, f_[2] = { huB[0][j], huB[1][j] };
                uint32_t g_[2] = { in.hu[0][j], in.hu[1][j] };
                bs_add3x2(e_, f_, g_, vu);
                uint32_t n_[2] = { hn[0][j], hn[1][j] }, o_[2] = { in.ho[0][j], in.ho[1][j] };
                uint32_t right = (j + 1 < WPL) ? ao[j + 1] : nxt;
                uint32_t r = dp_funnel_r(ao[j], right, 1);      /* old alive bit of x+1 */
                bs_count3d(vd, vu, n_, o_, r, k[j]);
            }
            /* rotate the H windows: `in` is free from here on */
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int j = 0; j < WPL; j++) {
                    hdA[b][j] = hdB[b][j]; hdB[b][j] = in.hd[b][j];
                    huA[b][j] = huB[b][j]; huB[b][j] = in.hu[b][j];
                }

            /* ---- prefetch row y+2 while the rule / scan / update below run ---- */
            if (y + 2 <= y1) {          /* row y1 is still needed (as "row y+1" of the last step), y1+1 is not */
                int need = y + 3 < H ? y + 3 : H;
                if (!wait_rows(p, f, need, 0))
                    return false;
                if (!load_row(p, pl, z, g, y + 2, lane, in, so2)) return false;
            }

            /* ---- rule tables, in-row scan ---- */
            uint32_t s0[WPL], s1[WPL], b0[WPL], b1[WPL], D[WPL], C[WPL];
            uint32_t dl = 1u, cl = 0u;          /* lane map: identity so far */
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                Rule::eval(p, k[j], s0[j], s1[j], b0[j], b1[j]);
                /* new alive bit if the predecessor's new alive bit is 0 / 1 */
                uint32_t f0 = bs_mux(ao[j], s0[j] | ge2[j], b0[j]) & vmask[j];
                uint32_t f1 = bs_mux(ao[j], s1[j] | ge2[j], b1[j]) & vmask[j];
                D[j] = f0 ^ f1;
                C[j] = f0;
                bs_scan_word(D[j], C[j]);
                /* compose into the lane map: carry-out = c ^ (d & carry-in) */
                uint32_t d = D[j] >> 31, c = C[j] >> 31;
                cl = c ^ (d & cl);
                dl = d & dl;
            }
            uint32_t cin = bs_scan_warp(dl, cl, 0u);

            /* ---- apply: new alive bits, state planes ---- */
            uint32_t an[WPL];
#pragma unroll
            for (int j = 0; j < WPL; j++) {
                uint32_t cm = 0u - cin;
                an[j] = C[j] ^ (D[j] & cm);
                uint32_t pred = (an[j] << 1) | cin;              /* new alive bit of x-1 */
                cin = an[j] >> 31;
                uint32_t sv = bs_mux(pred, s1[j], s0[j]);
                uint32_t bn = bs_mux(pred, b1[j], b0[j]);
                uint32_t dec = ao[j] & ~sv;                      /* alive, not surviving: state - 1 */
                uint32_t brn = ~ao[j] & bn & vmask[j];           /* dead, born: state = nr_states - 1 */
                uint32_t borrow = dec;
#pragma unroll
                for (int q = 0; q < P; q++) {
                    uint32_t t = so[q][j];
                    so[q][j] = t ^ borrow;
                    borrow &= ~t;
                    if ((bornval >> q) & 1u) so[q][j] |= brn;
                }
            }
            bp_hsum<WPL>(an, hn[0], hn[1]);

            /* ---- publish row y ---- */
            {
                uint32_t *rec = p.rows + ((size_t)z * H + y) * recw;
                LaneVec<WPL>::st(rec + 0 * p.RWP + lane * WPL, hn[0]);
                LaneVec<WPL>::st(rec + 1 * p.RWP + lane * WPL, hn[1]);
#pragma unroll
                for (int q = 0; q < P; q++)
                    LaneVec<WPL>::st(rec + (2 + q) * p.RWP + lane * WPL, so[q]);
                /*
                 * A z-block's edge plane also feeds the neighbouring GPU's ghost plane: tagged peer stores over
                 * NVLink, fire and forget -- no fence, no counter (the tag travels with every word).
                 */
                const uint32_t tag = (p.epoch << 16) | (uint32_t)(g + 1);
                if (pl.push_dn_rows)
                    store_h_tagged(pl.push_dn_rows + (size_t)y * pl.push_dn_stride, lane, tag, hn[0], hn[1]);
                if (pl.push_up_rows)
                    store_h_tagged(pl.push_up_rows + (size_t)y * pl.push_up_stride, lane, tag, hn[0], hn[1]);
                /*
                 * Local consumers: warp barrier (orders every lane's row stores before lane 0), then ONE
                 * cumulative gpu-scope fence and the relaxed counter store -- the same shape as
                 * cooperative-groups' grid barrier arrive.  Counters are raised every flag_rows rows.
                 */
                const bool raise = (y + 1 == y1) || ((y + 1) % p.flag_rows) == 0;
                if (raise)
                    dp_syncwarp();
                if (raise && lane == 0) {
                    if (!p.no_fence) dp_fence_release();
                    dp_st_flag(myprog, y + 1);
                }
            }

            /* ---- rotate the state pipeline ---- */
#pragma unroll
            for (int q = 0; q < P; q++)
#pragma unroll
                for (int j = 0; j < WPL; j++) { so[q][j] = so1[q][j]; so1[q][j] = so2[q][j]; }
        }
        return true;
    }

    CA_MDEV void kernel_body(const Bp3Params &p)
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
            if (!run_segment(p, it.x, it.y, it.z, it.w))
                break;
        }
    }
};

template <int P, int WPL, class Rule>
CA_GLOBAL void __launch_bounds__(256) ca3d_sweep_kernel(Bp3Params p)
{
    Sweep3<P, WPL, Rule>::kernel_body(p);
}

} // namespace clapca
#endif
