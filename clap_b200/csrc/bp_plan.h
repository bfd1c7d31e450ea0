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

/* size (in 32-bit words) of one ghost plane: H rows of {word, tag} pairs for [H0 | H1] */
inline size_t slab_ghost_plane_words(int H, int RWP) { return (size_t)H * 4 * RWP; }

/*
 * Layout of a rank's halo region (one device allocation, exported to the two
 * neighbouring ranks): the ghost planes below / above each local block.
 * Offsets in 32-bit words.
 */
struct HaloLayout {
    size_t ghost_dn, ghost_up, total_words;
    int nlb_max;
};

inline HaloLayout slab_halo_layout(const SlabGeom &geo, int H, int RWP)
{
    HaloLayout h;
    h.nlb_max = geo.blocks_per_rank_max();
    size_t gp = slab_ghost_plane_words(H, RWP);
    h.ghost_dn = 0;
    h.ghost_up = h.ghost_dn + gp * h.nlb_max;
    h.total_words = h.ghost_up + gp * h.nlb_max;
    if (h.total_words == 0) h.total_words = 4;
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

/* plane descriptors for every local plane of `geo.rank` */
inline void bp3_build_planes(const SlabGeom &geo, const SlabPtrs &ptr, const HaloLayout &hl, int H, int RWP, int NP,
                             std::vector<Bp3Plane> &out)
{
    const size_t recw = (size_t)NP * RWP;
    const size_t planew = (size_t)H * recw;
    const size_t gp = slab_ghost_plane_words(H, RWP);
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
                p.dn_stride = (uint32_t)recw;
                p.dn_flag = ptr.prog + (l - 1);
                p.dn_gstride = (uint32_t)Zl;
            } else if (j > 0) {
                p.dn_rows = ptr.halo + hl.ghost_dn + gp * lb;
                p.dn_stride = 4u * RWP;
                p.ghost_mask |= 1u;
                /* ... and this plane feeds the ghost plane ABOVE the previous block */
                const int lbp = (j - 1) / geo.R;
                p.push_dn_rows = ptr.halo_prev + hl.ghost_up + gp * lbp;
                p.push_dn_stride = 4u * RWP;
            }
            if (i + 1 < len) {
                p.up_rows = ptr.rows + (size_t)(l + 1) * planew;
                p.up_stride = (uint32_t)recw;
                p.up_flag = ptr.prog + (l + 1);
                p.up_gstride = (uint32_t)Zl;
            } else if (j + 1 < nb) {
                p.up_rows = ptr.halo + hl.ghost_up + gp * lb;
                p.up_stride = 4u * RWP;
                p.ghost_mask |= 2u;
                /* ... and this plane feeds the ghost plane BELOW the next block */
                const int lbn = (j + 1) / geo.R;
                p.push_up_rows = ptr.halo_next + hl.ghost_dn + gp * lbn;
                p.push_up_stride = 4u * RWP;
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
 * Team mode (ca3d_bitplane.cuh): an item is a GROUP of up to T consecutive planes of one z-block at one
 * generation, WorkItem{ first local plane, g, number of planes, H }.  Group (z0..z0+n-1, g) needs the group
 * holding z0-1 at g, the group holding z0+n at g-1 and itself at g-1; the key z0 + (T+1) g orders all three
 * before it for any grouping with n <= T (z0' >= z0 - T; z0 + n + (T+1)(g-1) < z0 + (T+1) g), and it is a
 * function of global coordinates, so every rank's list is a sub-sequence of ONE global linear extension of
 * the dependency order: the globally first unfinished group is always running or next in line on its rank.
 */
inline void bp3_make_items_team(const std::vector<Bp3Plane> &planes, int H, int G, int T,
                                std::vector<WorkItem> &items, bool layout_items = false)
{
    items.clear();
    if (G <= 0 || planes.empty())
        return;
    if (T < 1) T = 1;
    struct Group { int l0, n, z0; };
    std::vector<Group> groups;
    for (size_t l = 0; l < planes.size();) {
        size_t e = l + 1;
        /* extend while the next local plane is the next global plane of the same z-block (its "below" is local) */
        while (e < planes.size() && (int)(e - l) < T && planes[e].zglobal == planes[e - 1].zglobal + 1 &&
               planes[e].dn_rows && !(planes[e].ghost_mask & 1u))
            e++;
        groups.push_back(Group{ (int)l, (int)(e - l), planes[l].zglobal });
        l = e;
    }
    std::vector<std::pair<long long, WorkItem>> tmp;
    tmp.reserve(groups.size() * (size_t)(G + 2));
    /* layout_items: pack groups at "generation -1", unpack groups at "generation G"; the key orders them as well */
    for (int g = layout_items ? -1 : 0; g < (layout_items ? G + 1 : G); g++)
        for (const Group &gr : groups)
            tmp.push_back({ ((long long)gr.z0 + (long long)(T + 1) * g) * 65536 + g, WorkItem{ gr.l0, g, gr.n, H } });
    std::sort(tmp.begin(), tmp.end(),
              [](const std::pair<long long, WorkItem> &a, const std::pair<long long, WorkItem> &b) {
                  return a.first < b.first;
              });
    for (auto &t : tmp) items.push_back(t.second);
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
