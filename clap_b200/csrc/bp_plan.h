/*
 * bp_plan.h -- host-side planning shared by the C-ABI library and the kernel
 * emulator tests: bit-plane count, sweep claim order, and the z-block slab
 * decomposition (who owns which planes, where ghost planes live, whom an
 * edge plane feeds).  Plain C++ (no CUDA).
 */
#ifndef CLAPCA_BP_PLAN_H
#define CLAPCA_BP_PLAN_H
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include <algorithm>
#include "bp3_types.h"

namespace clapca {

/* bit planes needed so that every value that can ever occur fits: {3,4,8} are instantiated */
inline int bp_planes_for(unsigned maxval)
{
    if (maxval < 8) return 3;
    if (maxval < 16) return 4;
    return 8;
}

/* words per lane for a row of W cells handled by one warp: {1,2,4}, 0 = too wide */
inline int bp_wpl_for(int W)
{
    if (W <= 1024) return 1;
    if (W <= 2048) return 2;
    if (W <= 4096) return 4;
    return 0;
}

/*
 * Slab decomposition of the z axis (the outermost sweep axis, core/ca3d.c:130)
 * over R ranks: the Zg planes are cut into blocks of B planes, block j belongs
 * to rank j % R (block-cyclic).  Each rank stores its blocks back to back.
 * Contiguous slabs are the special case B = ceil(Zg / R).  Small blocks keep
 * the pipeline fill short: rank r can start 3*r*B row steps after rank 0
 * instead of 3*r*Zg/R.
 */
struct SlabGeom {
    int Zg, R, rank, B;
    int nblocks() const { return (Zg + B - 1) / B; }
    int blocks_per_rank_max() const { return (nblocks() + R - 1) / R; }
    int local_blocks() const { int n = nblocks(); return n > rank ? (n - rank + R - 1) / R : 0; }
    int block_z0(int j) const { return j * B; }
    int block_len(int j) const { return std::min(B, Zg - j * B); }
    int global_block(int lb) const { return lb * R + rank; }
    int local_planes() const
    {
        int n = 0;
        for (int lb = 0; lb < local_blocks(); lb++) n += block_len(global_block(lb));
        return n;
    }
    /* local index of the first plane of local block lb */
    int local_z0(int lb) const
    {
        int n = 0;
        for (int i = 0; i < lb; i++) n += block_len(global_block(i));
        return n;
    }
};

/*
 * Layout of a rank's halo region (one device allocation, exported to the two neighbouring ranks), offsets in
 * 32-bit words: the ghost planes below / above each local block -- a ghost plane has the record stride of a local
 * plane (NP * RWP words per row, only H0 | H1 are ever written), so the sweep reads it like any other plane --
 * followed by their progress counters, [bank][below / above][local block][generation].  Every rank uses the same
 * layout (nlb_max is the largest number of blocks any rank owns), so a peer's offsets are known without asking.
 */
struct HaloLayout {
    size_t gp;                      /* words per ghost plane */
    size_t ghost_dn, ghost_up;      /* first ghost plane below / above; local block lb at + gp * lb */
    size_t flags;                   /* counters: bank b at + b * bank_words() ints; Gcap + 1 per ghost plane, the first one
                                       is "generation -1" (layout items: the neighbour's pack item) */
    size_t total_words;
    int nlb_max, Gcap;
    size_t flag_index(int bank, int up, int lb) const
    {
        return flags + (((size_t)bank * 2 + (size_t)up) * nlb_max + (size_t)lb) * (Gcap + 1) + 1;    /* generation 0 */
    }
    size_t bank_words() const { return (size_t)2 * nlb_max * (Gcap + 1); }
};

inline HaloLayout slab_halo_layout(const SlabGeom &geo, int H, int RWP, int NP, int Gcap)
{
    HaloLayout h;
    h.nlb_max = geo.blocks_per_rank_max();
    h.Gcap = Gcap > 0 ? Gcap : 1;
    h.gp = (size_t)H * NP * RWP;
    h.ghost_dn = 0;
    h.ghost_up = h.ghost_dn + h.gp * h.nlb_max;
    h.flags = h.ghost_up + h.gp * h.nlb_max;
    h.total_words = h.flags + 2 * h.bank_words();
    if (geo.R == 1) {               /* one rank: no ghosts at all */
        h.gp = 0;
        h.ghost_dn = h.ghost_up = h.flags = 0;
        h.total_words = 4;
    }
    return h;
}

/* base pointers of one rank as seen from its own process */
struct SlabPtrs {
    uint32_t *rows;         /* local row records */
    int *prog;              /* local progress counters [G][Zl] */
    uint32_t *halo;         /* own halo region */
    uint32_t *halo_next;    /* halo region of rank (rank + 1) % R, peer-mapped */
    uint32_t *halo_prev;    /* halo region of rank (rank - 1 + R) % R, peer-mapped */
};

/* plane descriptors for every local plane of `geo.rank`; `bank` selects the ghost counters of this run */
inline void bp3_build_planes(const SlabGeom &geo, const SlabPtrs &ptr, const HaloLayout &hl, int H, int RWP, int NP,
                             std::vector<Bp3Plane> &out, int bank = 0)
{
    const size_t recw = (size_t)NP * RWP;
    const size_t planew = (size_t)H * recw;
    const int Zl = geo.local_planes();
    const int nb = geo.nblocks();
    out.assign(Zl, Bp3Plane());
    for (int lb = 0; lb < geo.local_blocks(); lb++) {
        const int j = geo.global_block(lb);
        const int l0 = geo.local_z0(lb), len = geo.block_len(j);
        for (int i = 0; i < len; i++) {
            const int l = l0 + i;
            Bp3Plane &p = out[l];
            p = Bp3Plane();
            p.zglobal = geo.block_z0(j) + i;
            if (i > 0) {
                p.dn_rows = ptr.rows + (size_t)(l - 1) * planew;
                p.dn_flag = ptr.prog + (l - 1);
                p.dn_gstride = (uint32_t)Zl;
            } else if (j > 0) {
                p.dn_rows = ptr.halo + hl.ghost_dn + hl.gp * lb;
                p.dn_flag = (const int *)(ptr.halo + hl.flag_index(bank, 0, lb));
                p.dn_gstride = 1u;
                p.ghost_mask |= 1u;
                /* ... and this plane feeds the ghost plane ABOVE the previous block */
                const int lbp = (j - 1) / geo.R;
                p.push_dn_rows = ptr.halo_prev + hl.ghost_up + hl.gp * lbp;
                p.push_dn_flag = (int *)(ptr.halo_prev + hl.flag_index(bank, 1, lbp));
            }
            if (i + 1 < len) {
                p.up_rows = ptr.rows + (size_t)(l + 1) * planew;
                p.up_flag = ptr.prog + (l + 1);
                p.up_gstride = (uint32_t)Zl;
            } else if (j + 1 < nb) {
                p.up_rows = ptr.halo + hl.ghost_up + hl.gp * lb;
                p.up_flag = (const int *)(ptr.halo + hl.flag_index(bank, 1, lb));
                p.up_gstride = 1u;
                p.ghost_mask |= 2u;
                /* ... and this plane feeds the ghost plane BELOW the next block */
                const int lbn = (j + 1) / geo.R;
                p.push_up_rows = ptr.halo_next + hl.ghost_dn + hl.gp * lbn;
                p.push_up_flag = (int *)(ptr.halo_next + hl.flag_index(bank, 0, lbn));
            }
        }
    }
}

struct WorkItem { int z, g, y0, y1; };      /* local plane, generation, rows [y0, y1) */

/*
 * Work items and their claim order.
 *
 * Dependencies (global z): rows <= y+2 of (z-1,g) [new plane below], of (z+1,g-1) [old plane
 * above] and of (z,g-1) [own old state] must be complete before row y of (z,g).  Cutting every
 * sweep into segments of L rows whose boundaries are skewed by 3 rows per plane and 6 rows per
 * generation,
 *         segment s of (z,g) = rows [s L - 3 z - 6 g, (s+1) L - 3 z - 6 g)  clipped to [0,H),
 * makes segment s of (z,g) depend only on segment s (and earlier) of its three producers and on
 * its own segment s-1.  Items are claimed band by band (s), inside a band along anti-diagonals
 * z + g (then g): this extends the dependency order -- (z-1,g,s), (z,g-1,s) lie on the previous
 * anti-diagonal, (z+1,g-1,s) earlier on the same one -- and it makes consecutive generations of a
 * plane follow each other a handful of rows apart, so a band's working set stays in L2 and HBM
 * sees roughly one read and one write of the volume for ALL generations.  Every rank enumerates
 * the same global order and keeps its own planes, so the globally first unfinished item is always
 * running or next in line on its rank: no deadlock.
 */
inline void bp3_make_items(const std::vector<Bp3Plane> &planes, int Zg, int H, int G, int L,
                           std::vector<WorkItem> &items)
{
    items.clear();
    if (G <= 0 || planes.empty())
        return;
    std::vector<int> local_of(Zg, -1);
    for (size_t l = 0; l < planes.size(); l++)
        local_of[planes[l].zglobal] = (int)l;
    const long long span = (long long)H + 3LL * (Zg - 1) + 6LL * (G - 1);
    const int nbands = (int)((span + L - 1) / L);
    for (int s = 0; s < nbands; s++)
        for (int d = 0; d <= (Zg - 1) + (G - 1); d++)
            for (int g = std::max(0, d - (Zg - 1)); g <= std::min(G - 1, d); g++) {
                const int z = d - g;
                const int l = local_of[z];
                if (l < 0)
                    continue;
                long long y0 = (long long)s * L - 3LL * z - 6LL * g;
                long long y1 = y0 + L;
                if (y0 < 0) y0 = 0;
                if (y1 > H) y1 = H;
                if (y0 < y1)
                    items.push_back(WorkItem{ l, g, (int)y0, (int)y1 });
            }
}

/*
 * Alternative claim order: whole-plane sweeps (one item per (z,g)) sorted by the time key 2 z + 4 g.
 * Items with equal key are mutually independent and their producers sit 50..100 tickets earlier,
 * so claimed items rarely stall on each other; the price is that consecutive generations of a plane
 * run far apart and the volume streams through HBM once per generation.
 */
inline void bp3_make_items_timekey(const std::vector<Bp3Plane> &planes, int H, int G, std::vector<WorkItem> &items,
                                   bool layout_items = false)
{
    std::vector<std::pair<long long, WorkItem>> tmp;
    tmp.reserve(planes.size() * (size_t)(G + 2));
    /*
     * layout_items (ca3d_bitplane.cuh): every plane gets a pack item at "generation -1" and an unpack item at
     * "generation G".  The same key orders them: pack (z,-1) and (z+1,-1) precede sweep (z,0), unpack (z,G)
     * follows sweep (z,G-1).
     */
    for (int g = layout_items ? -1 : 0; g < (layout_items ? G + 1 : G); g++)
        for (size_t l = 0; l < planes.size(); l++)
            tmp.push_back({ (2LL * planes[l].zglobal + 4LL * g) * 65536 + g, WorkItem{ (int)l, g, 0, H } });
    std::sort(tmp.begin(), tmp.end(),
              [](const std::pair<long long, WorkItem> &a, const std::pair<long long, WorkItem> &b) {
                  return a.first < b.first;
              });
    items.clear();
    for (auto &t : tmp) items.push_back(t.second);
}

/*
 * Generation-batched diagonal order (whole-plane sweeps): the G generations are cut into batches of Gb;
 * inside a batch items are claimed along anti-diagonals d = z + (g - g0), generation ascending.  This
 * extends the dependency order -- (z-1,g) lies on the previous diagonal, (z+1,g-1) just before (z,g) on the
 * same one, earlier batches come first -- and it puts (z,g+1) only Gb tickets after (z,g): consecutive
 * generations of a plane chase each other a few rows apart, so a row written by generation g is still in
 * L2 when generation g+1 reads and overwrites it.  HBM then sees one read and one write of the volume per
 * BATCH instead of per generation.  (In time-key order every key holds one item per generation, G tickets
 * separate (z,g) from (z,g+1), and with ~2400 resident warps that is ~40 rows x 2400 rows of other traffic:
 * far more than L2.)  Small Gb = short reuse distance but fewer independent chains per diagonal.
 */
inline void bp3_make_items_batched(const std::vector<Bp3Plane> &planes, int Zg, int H, int G, int Gb,
                                   std::vector<WorkItem> &items)
{
    items.clear();
    if (G <= 0 || planes.empty())
        return;
    if (Gb < 1) Gb = 1;
    std::vector<int> local_of(Zg, -1);
    for (size_t l = 0; l < planes.size(); l++)
        local_of[planes[l].zglobal] = (int)l;
    /* equal-sized batches: ceil(G / Gb) of them */
    const int nb = (G + Gb - 1) / Gb;
    for (int b = 0; b < nb; b++) {
        const int g0 = (int)((long long)G * b / nb), g1 = (int)((long long)G * (b + 1) / nb);
        const int n = g1 - g0;
        for (int d = 0; d <= (Zg - 1) + (n - 1); d++)
            for (int k = std::max(0, d - (Zg - 1)); k <= std::min(n - 1, d); k++) {
                const int l = local_of[d - k];
                if (l >= 0)
                    items.push_back(WorkItem{ l, g0 + k, 0, H });
            }
    }
}

/*
 * Tile mode (ca3d_bitplane.cuh): an item is a TILE of up to Tz consecutive planes of one z-block x up to Tg
 * consecutive generations, WorkItem{ first local plane, first generation, nz | ng << 8, H }, swept by one CTA.
 *
 * Row-level dependencies: (z,g) needs (z-1,g), (z+1,g-1) and (z,g-1).  For a tile X = planes z0..z0+nz-1 x
 * generation group b that means the tiles holding z0-1 in group b, X's planes in group b-1 -- and, for ng > 1, the
 * tile holding z0+nz in the SAME group (its first plane at generation g0+j-1 is the "old plane above" of our last
 * plane at g0+j), which in turn depends on X: mutually dependent tiles advance together row by row and must be
 * co-resident.  The claim key is z0 + (Tz+1) b.  Everything a row of X can transitively depend on is some
 * (z', g') with z' <= z0+nz-1+k and g' <= g0+ng-1-k for a k >= 0; the tile holding it has a key of at most
 * key(X) + Tz + Tg - 2 provided Tg <= Tz + 1 (k more planes cost k keys, every Tg generations back give Tz+1 keys
 * back).  The key is a function of global coordinates, so every rank's list is a sub-sequence of ONE global order.
 * Progress: let X be the unfinished tile with the smallest key.  Whatever X waits for is finished or lies in the
 * key window [key(X), key(X) + Tz + Tg]; every rank claims in key order and nothing below key(X) is unfinished, so
 * as long as every rank has more CTAs than items inside any such window, all of X's partners get claimed, and then
 * the row-level DAG lets the earliest unfinished row run.  bp3_tile_shape() picks the largest Tg that satisfies
 * this for the given number of CTAs (Tg = 1 has no forward dependency at all: the plane groups of round 1).
 */
inline void bp3_make_items_tile(const std::vector<Bp3Plane> &planes, int H, int G, int Tz, int Tg,
                                std::vector<WorkItem> &items, bool layout_items = false,
                                std::vector<long long> *keys_out = nullptr, int gskew = 0)
{
    /*
     * gskew = key distance between consecutive generation groups of a plane group (> Tz, so that every predecessor
     * of a tile has a smaller key).  Tz + 1 keeps generation group b + 1 right behind group b (reuse through L2, and
     * every rank of a sharded volume busy); a large value runs the generation groups one after the other, which puts
     * a tile's FORWARD partner (next tile in z, same group) on the very next ticket -- the partner then starts ~one
     * plane-hop after the tile instead of ~one claim interval per active generation group after it, and the tile's
     * warps finish together instead of waiting for each other at the end-of-item barrier (measured: 9.7 % of all warp
     * cycles at 2048^3 x 50 with Tz + 1).
     */
    if (gskew <= Tz) gskew = Tz + 1;
    items.clear();
    if (keys_out) keys_out->clear();
    if (G <= 0 || planes.empty())
        return;
    if (Tz < 1) Tz = 1;
    if (Tz > 255) Tz = 255;
    if (Tg < 1) Tg = 1;
    if (Tg > 255) Tg = 255;
    struct Group { int l0, n, z0; };
    std::vector<Group> groups;
    for (size_t l = 0; l < planes.size();) {
        size_t e = l + 1;
        /* extend while the next local plane is the next global plane of the same z-block (its "below" is local) */
        while (e < planes.size() && (int)(e - l) < Tz && planes[e].zglobal == planes[e - 1].zglobal + 1 &&
               planes[e].dn_rows && !(planes[e].ghost_mask & 1u))
            e++;
        groups.push_back(Group{ (int)l, (int)(e - l), planes[l].zglobal });
        l = e;
    }
    const int nb = (G + Tg - 1) / Tg;
    std::vector<std::pair<long long, WorkItem>> tmp;
    tmp.reserve(groups.size() * (size_t)(nb + 2));
    /* layout_items: pack groups one "generation group" before the first, unpack groups one after the last */
    for (int b = layout_items ? -1 : 0; b < (layout_items ? nb + 1 : nb); b++)
        for (const Group &gr : groups) {
            const int g0 = b < 0 ? -1 : (b >= nb ? G : b * Tg);
            const int ng = (b < 0 || b >= nb) ? 1 : std::min(Tg, G - g0);
            tmp.push_back({ ((long long)gr.z0 + (long long)gskew * b) * 65536 + (b + 1),
                            WorkItem{ gr.l0, g0, gr.n | (ng << 8), H } });
        }
    std::sort(tmp.begin(), tmp.end(),
              [](const std::pair<long long, WorkItem> &a, const std::pair<long long, WorkItem> &b) {
                  return a.first < b.first;
              });
    for (auto &t : tmp) {
        items.push_back(t.second);
        if (keys_out) keys_out->push_back(t.first >> 16);
    }
}

/* the plane list of `geo.rank` with just enough filled in for the grouping of bp3_make_items_tile (no pointers to follow) */
inline void bp3_topology_planes(const SlabGeom &geo, std::vector<Bp3Plane> &out)
{
    static const uint32_t marker = 0;
    out.assign(geo.local_planes(), Bp3Plane());
    for (int lb = 0; lb < geo.local_blocks(); lb++) {
        const int j = geo.global_block(lb);
        const int l0 = geo.local_z0(lb), len = geo.block_len(j);
        for (int i = 0; i < len; i++) {
            Bp3Plane &p = out[l0 + i];
            p.zglobal = geo.block_z0(j) + i;
            if (i > 0) {
                p.dn_rows = &marker;
            } else if (j > 0) {
                p.dn_rows = &marker;
                p.ghost_mask |= 1u;
            }
        }
    }
}

/* most items of one rank inside any key window [k, k + width] of the tile order */
inline int bp3_tile_window_items(const std::vector<long long> &keys, long long width)
{
    int best = 0;
    size_t lo = 0;
    for (size_t hi = 0; hi < keys.size(); hi++) {
        while (keys[hi] - keys[lo] > width) lo++;
        best = std::max(best, (int)(hi - lo + 1));
    }
    return best;
}

/*
 * Shape of the tiles for a team of `team` compute warps: Tz * Tg = team with the largest Tg <= min(want, Tz) for which
 * `ctas` co-resident CTAs exceed the number of this rank's items inside any window of Tz + Tg keys (see above);
 * Tg = 1 (plane groups, no forward dependency) always qualifies.  Returns Tg, stores Tz.
 */
inline int bp3_tile_shape(const std::vector<Bp3Plane> &planes, int H, int G, int team, int want, int ctas,
                          bool layout_items, int *Tz_out, int gskew = 0)
{
    if (team < 1) team = 1;
    std::vector<WorkItem> items;
    std::vector<long long> keys;
    for (int Tg = std::max(1, std::min(want, G)); Tg > 1; Tg--) {
        if (team % Tg) continue;
        const int Tz = team / Tg;
        if (Tg > Tz) continue;
        bp3_make_items_tile(planes, H, G, Tz, Tg, items, layout_items, &keys, gskew);
        if (bp3_tile_window_items(keys, (long long)Tz + Tg) < ctas) {
            *Tz_out = Tz;
            return Tg;
        }
    }
    *Tz_out = team;
    return 1;
}

/* the same for a sharded volume: every rank must run the same shape, so take the smallest Tg over the ranks' plane lists */
inline int bp3_tile_shape_all_ranks(const SlabGeom &geo, int H, int G, int team, int want, int ctas, int *Tz_out,
                                    bool layout_items = false, int gskew = 0)
{
    int Tg = std::max(1, want), Tz = team;
    std::vector<Bp3Plane> planes;
    for (int r = 0; r < geo.R; r++) {
        SlabGeom g = geo;
        g.rank = r;
        bp3_topology_planes(g, planes);
        Tg = std::min(Tg, bp3_tile_shape(planes, H, G, team, Tg, ctas, layout_items, &Tz, gskew));
    }
    /* Tg only ever shrinks along the loop and a shape that passed for a larger Tg was not necessarily re-checked
       for this one on the earlier ranks: confirm, falling back to plane groups */
    if (Tg > 1) {
        for (int r = 0; r < geo.R; r++) {
            SlabGeom g = geo;
            g.rank = r;
            bp3_topology_planes(g, planes);
            int tz = team;
            if (bp3_tile_shape(planes, H, G, team, Tg, ctas, layout_items, &tz, gskew) != Tg) { Tg = 1; break; }
        }
    }
    *Tz_out = Tg > 1 ? team / Tg : team;
    return Tg;
}

/* segment length: long enough that a band offers ~2x more independent items than there are workers */
inline int bp3_segment_rows(int Zg, int H, int G, int workers)
{
    const double chain = 3.0 * Zg + 6.0 * G;            /* dependency chain through one band, in rows */
    int L = 32;
    while (L < 256 && (double)Zg * G * L / (chain + L) < 2.0 * workers)
        L *= 2;
    if (L > H) L = std::max(H, 1);
    return L;
}

} // namespace clapca
#endif
