/*
 * ca3d_layout.cuh -- conversion between the reference layout (uint8 cells,
 * z*d0*d1 + y*d0 + x, core/xyarray.c:43) and the row records of the bit-plane
 * engine (see ca3d_bitplane.cuh), plus the population count of the result
 * (xyzarray_count(), core/xyarray.c:68-78) fused into the way back.
 */
#ifndef CLAPCA_CA3D_LAYOUT_CUH
#define CLAPCA_CA3D_LAYOUT_CUH

#include "devport.h"
#include "bp3_types.h"

namespace clapca {

/* ---- layout conversion ------------------------------------------------------- */

/* struct Bp3Layout: bp3_types.h */


/* gather bit q of each of 4 packed cells (bytes) into a nibble: cell k -> bit k */
CA_DEV uint32_t lay_gather4(uint32_t bytes_lsb)
{
    /* bytes_lsb has the wanted bit at bit 0 of every byte; 0x01020408 moves byte k's bit to bit 24+k */
    return ((bytes_lsb & 0x01010101u) * 0x01020408u) >> 24;
}

/* spread a nibble to the lsb of 4 bytes: bit k -> byte k */
CA_DEV uint32_t lay_spread4(uint32_t nibble)
{
    return ((nibble & 0xfu) * 0x00204081u) & 0x01010101u;
}

/* 0x01 in every byte of v that is non-zero */
CA_DEV uint32_t lay_nonzero4(uint32_t v)
{
    v |= v >> 4;
    v |= v >> 2;
    v |= v >> 1;
    return v & 0x01010101u;
}

/*
 * uint8 volume -> row records.  One warp per 32 consecutive words of a row (a row is 32 * WPL words): lane l packs the
 * 32 cells of word 32 * group + l with two 16-byte loads and one multiply-gather per plane and register; the alive
 * bits of the neighbouring words come from the neighbouring lanes, so H0 | H1 cost two shuffles (the first / last
 * lane of a group that is not at the row's end looks at one more byte).  P is a template parameter and the item ->
 * (row, word) mapping is shifts: the first version of this kernel spent more instructions on two 64-bit divisions
 * and a run-time plane loop than on the cells (4.7 ms at 2048^3 against 2.1 ms for its 14 GB of HBM traffic).
 *
 * alive = OR of the P state planes -- exact when every value fits P planes.  The caller either knows that (it scanned
 * the volume, or a caller's bound was verified), or it passes L.over: the kernel ORs in a 1 when a cell does not fit,
 * and the caller packs again with the number of planes a max scan asks for (volumes seeded with 255s, core/ca3d.c:41-59,
 * take that path; a volume whose values stay below nr_states saves the scan: 1.4 ms at 2048^3).
 */
template <int P>
CA_GLOBAL void __launch_bounds__(256) ca3d_pack_rows_kernel(Bp3Layout L)
{
    constexpr int NP = P + 2;
    constexpr uint32_t kOver = P >= 8 ? 0u : (((0xffu << P) & 0xffu) * 0x01010101u);
    const int lane = dp_lane();
    const int gshift = L.RWP == 32 ? 0 : (L.RWP == 64 ? 1 : 2);        /* RWP = 32 * WPL, WPL = 1, 2, 4 */
    const size_t nitems = ((size_t)L.Z * L.H) << gshift;
    const size_t wstride = (size_t)dp_grid_blocks() * (dp_block_threads() >> 5);
    const bool vec = (L.W & 15) == 0;
    uint32_t over = 0u;
    for (size_t it = (size_t)dp_block() * (dp_block_threads() >> 5) + dp_warp_in_block(); it < nitems; it += wstride) {
        const size_t row = it >> gshift;
        const int w = ((int)(it & ((1u << gshift) - 1u)) << 5) + lane;
        const uint8_t *src = L.cells + row * L.W;
        uint32_t *rec = L.rows + row * NP * L.RWP + w;
        const int x0 = 32 * w;
        int n = L.W - x0;
        n = n > 32 ? 32 : (n < 0 ? 0 : n);
        uint32_t s[P];
#pragma unroll
        for (int q = 0; q < P; q++) s[q] = 0u;
        if (n == 32 && vec) {
            const uint4 lo = *reinterpret_cast<const uint4 *>(src + x0);
            const uint4 hi = *reinterpret_cast<const uint4 *>(src + x0 + 16);
            const uint32_t r[8] = { lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w };
#pragma unroll
            for (int j = 0; j < 8; j++) {
                over |= r[j];
#pragma unroll
                for (int q = 0; q < P; q++)
                    s[q] += lay_gather4(r[j] >> q) << (4 * j);      /* disjoint bits: + is |, and an IMAD */
            }
        } else {
            for (int i2 = 0; i2 < n; i2++) {
                const uint32_t v = src[x0 + i2];
                over |= v;
#pragma unroll
                for (int q = 0; q < P; q++) s[q] |= ((v >> q) & 1u) << i2;
            }
        }
        uint32_t alive = 0u;
#pragma unroll
        for (int q = 0; q < P; q++) alive |= s[q];
        uint32_t left = dp_shfl_up0(alive) >> 31, right = dp_shfl_down0(alive) & 1u;
        if (lane == 0 && x0 > 0 && n > 0) left = (uint32_t)(src[x0 - 1] != 0);
        if (lane == 31 && x0 + 32 < L.W) right = (uint32_t)(src[x0 + 32] != 0);
        const uint32_t l = (alive << 1) | left, r = (alive >> 1) | (right << 31);
        rec[0] = l ^ alive ^ r;
        rec[(size_t)L.RWP] = (l & alive) | (l & r) | (alive & r);
#pragma unroll
        for (int q = 0; q < P; q++)
            rec[(size_t)(2 + q) * L.RWP] = s[q];
    }
    if (L.over && dp_any((over & kOver) != 0u) && lane == 0)
        dp_atomic_or_global(L.over, 1u);
}

/* row records -> uint8 volume, and the population count of the result; same work split as the pack kernel */
template <int P>
CA_GLOBAL void __launch_bounds__(256) ca3d_unpack_rows_kernel(Bp3Layout L)
{
    constexpr int NP = P + 2;
    const int lane = dp_lane();
    const int gshift = L.RWP == 32 ? 0 : (L.RWP == 64 ? 1 : 2);
    const size_t nitems = ((size_t)L.Z * L.H) << gshift;
    const size_t wstride = (size_t)dp_grid_blocks() * (dp_block_threads() >> 5);
    const bool vec = (L.W & 15) == 0;
    unsigned pop = 0u;
    unsigned long long total = 0ull;
    for (size_t it = (size_t)dp_block() * (dp_block_threads() >> 5) + dp_warp_in_block(); it < nitems; it += wstride) {
        const size_t row = it >> gshift;
        const int w = ((int)(it & ((1u << gshift) - 1u)) << 5) + lane;
        const int x0 = 32 * w;
        int n = L.W - x0;
        n = n > 32 ? 32 : (n < 0 ? 0 : n);
        if (n > 0) {
            const uint32_t *rec = L.rows + row * NP * L.RWP + w;
            uint8_t *dst = L.cells + row * L.W;
            uint32_t s[P], alive = 0u;
#pragma unroll
            for (int q = 0; q < P; q++) {
                s[q] = rec[(size_t)(2 + q) * L.RWP];
                alive |= s[q];
            }
            pop += (unsigned)dp_popc(alive);
            if (n == 32 && vec) {
                uint32_t r[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    uint32_t v = 0u;
#pragma unroll
                    for (int q = 0; q < P; q++)
                        v += lay_spread4(s[q] >> (4 * j)) << q;
                    r[j] = v;
                }
                *reinterpret_cast<uint4 *>(dst + x0) = make_uint4(r[0], r[1], r[2], r[3]);
                *reinterpret_cast<uint4 *>(dst + x0 + 16) = make_uint4(r[4], r[5], r[6], r[7]);
            } else {
                for (int i2 = 0; i2 < n; i2++) {
                    uint32_t v = 0u;
#pragma unroll
                    for (int q = 0; q < P; q++) v |= ((s[q] >> i2) & 1u) << q;
                    dst[x0 + i2] = (uint8_t)v;
                }
            }
        }
        if (pop >= 0x40000000u) { total += pop; pop = 0u; }
    }
    total += pop;
    /* warp sum, one atomic per warp */
    uint32_t lo = (uint32_t)total, hi = (uint32_t)(total >> 32);
    for (int o = 16; o; o >>= 1) {
        const uint32_t lo2 = dp_shfl_down(lo, o), hi2 = dp_shfl_down(hi, o);
        const unsigned long long a = ((unsigned long long)hi << 32 | lo) + ((unsigned long long)hi2 << 32 | lo2);
        lo = (uint32_t)a; hi = (uint32_t)(a >> 32);
    }
    if (lane == 0 && (lo | hi))
        dp_atomic_add64(L.population, (unsigned long long)hi << 32 | lo);
}

/* ---- ca3d_make() on a device-resident volume: core/ca3d.c:41-59, 144-169 ------------------------------------- */

/* the six faces of the (zeroed) volume get the value 5: one thread per row (y, z) */
CA_GLOBAL void make3d_faces_kernel(uint8_t *cells, int d0, int d1, int d2)
{
    const size_t rows = (size_t)d1 * d2;
    const size_t stride = (size_t)dp_grid_blocks() * dp_block_threads();
    for (size_t r = (size_t)dp_block() * dp_block_threads() + dp_thread(); r < rows; r += stride) {
        const int y = (int)(r % d1), z = (int)(r / d1);
        uint8_t *row = cells + r * d0;
        if (y == 0 || y == d1 - 1 || z == 0 || z == d2 - 1) {
            for (int x = 0; x < d0; x++) row[x] = 5;
        } else {
            row[0] = 5;
            row[d0 - 1] = 5;
        }
    }
}

/* the cells the random walk visited */
CA_GLOBAL void make3d_scatter_kernel(uint8_t *cells, const unsigned long long *idx, size_t n)
{
    const size_t stride = (size_t)dp_grid_blocks() * dp_block_threads();
    for (size_t i = (size_t)dp_block() * dp_block_threads() + dp_thread(); i < n; i += stride)
        cells[idx[i]] = 5;
}

/*
 * ca3d_prune(), first loop (core/ca3d.c:45-49): in sweep order, a cell whose six von Neumann neighbours are all
 * occupied becomes (unsigned char)-1 = 255 -- an EMPTY enclosed cell too, and from then on it counts as occupied for
 * the cells after it (+x, +y, +z), not for the cells before it.  The second loop never finds an int equal to -1
 * (SURVEY F5), so the marks stay.  In parallel: a cell sees the NEW state of its three predecessors and the OLD state
 * of its three successors; the system is triangular, so iterating "mark what has six" to a fixed point gives the
 * sweep's result.  An empty cell that turns occupied is held as 254 until the fixed point (occupied for the cells
 * after it, still empty for the cells before it); occupied cells go to 255 directly (nobody's count depends on it).
 * *changed counts the empty cells newly marked in this pass.
 */
CA_GLOBAL void make3d_prune_kernel(uint8_t *cells, int d0, int d1, int d2, unsigned *changed)
{
    const size_t n = (size_t)d0 * d1 * d2, plane = (size_t)d0 * d1;
    const size_t stride = (size_t)dp_grid_blocks() * dp_block_threads();
    for (size_t i = (size_t)dp_block() * dp_block_threads() + dp_thread(); i < n; i += stride) {
        const int x = (int)(i % d0), y = (int)((i / d0) % d1), z = (int)(i / plane);
        if (x == 0 || y == 0 || z == 0 || x == d0 - 1 || y == d1 - 1 || z == d2 - 1)
            continue;                                   /* a neighbour outside the volume reads 0: never six */
        const volatile uint8_t *c = cells + i;
        const uint8_t v = c[0];
        if (v == 255 || v == 254)
            continue;
        /* predecessors: new state (254 counts); successors: old state (254 is still empty) */
        const uint8_t sx = c[1], sy = c[d0], sz = c[plane];
        const int cnt = (c[-1] != 0) + (c[-(ptrdiff_t)d0] != 0) + (c[-(ptrdiff_t)plane] != 0) +
                        (sx != 0 && sx != 254) + (sy != 0 && sy != 254) + (sz != 0 && sz != 254);
        if (cnt == 6) {
            if (v == 0) {
                cells[i] = 254;
                dp_atomic_or_global(changed, 1u);
            } else {
                cells[i] = 255;
            }
        }
    }
}

/* after the fixed point: the held marks become 255; counts the occupied cells (xyzarray_count, core/ca3d.c:98) */
CA_GLOBAL void make3d_finish_kernel(uint8_t *cells, size_t n, unsigned long long *population)
{
    const size_t stride = (size_t)dp_grid_blocks() * dp_block_threads();
    unsigned long long pop = 0;
    for (size_t i = (size_t)dp_block() * dp_block_threads() + dp_thread(); i < n; i += stride) {
        uint8_t v = cells[i];
        if (v == 254)
            cells[i] = v = 255;
        pop += v != 0;
    }
    if (pop)
        dp_atomic_add64(population, pop);
}

/*
 * Seed a neighbour's ghost plane with the H rows of one local plane (the "old plane above" of generation 0):
 * src = first row record of the plane, dst = ghost plane (peer memory), both with the record stride NP * RWP.
 */
CA_GLOBAL void halo_seed_kernel(uint32_t *dst, const uint32_t *src, int H, int RWP, int NP)
{
    const size_t n = (size_t)H * 2 * RWP;
    const size_t stride = (size_t)dp_grid_blocks() * dp_block_threads();
    for (size_t i = (size_t)dp_block() * dp_block_threads() + dp_thread(); i < n; i += stride) {
        const size_t y = i / (2 * RWP), q = i % (2 * RWP);
        dst[y * NP * RWP + q] = src[y * NP * RWP + q];
    }
}

} // namespace clapca
#endif
