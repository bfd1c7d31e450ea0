/*
 * ca2d_skew_layout.cuh -- conversion between the reference's 2D grid (uint8, index y*w + x, core/xyarray.c:43 with
 * d2 == 1) and the diagonal rows of the skewed 2D engine (ca2d_skew.cuh): bit x of row t = cell (x, y = t - 2x).
 */
#ifndef CLAPCA_CA2D_SKEW_LAYOUT_CUH
#define CLAPCA_CA2D_SKEW_LAYOUT_CUH

#include "ca2d_skew.cuh"

namespace clapca {

/* ---- layout: uint8 grid (index y*w + x) <-> diagonal rows ------------------------------------------------- */

struct Sk2Layout {
    uint8_t *cells;
    uint32_t *rows;         /* [TR][SK2_RS] */
    int w, h;
    unsigned long long *population;     /* unpack: number of non-zero cells */
};

enum { SK2_TILE_T = 450, SK2_TILE_Y = SK2_TILE_T + 62 };    /* diagonals per pack tile / grid rows it touches */

/*
 * pack: one CTA = one word column j (cells x = 32 j .. 32 j + 31) x SK2_TILE_T diagonals.  The grid rows those
 * words touch (y = t - 2x: SK2_TILE_T + 62 rows of 32 bytes) go through shared memory so that global reads run along
 * x; then one ballot per word.  Tiles outside the band of valid cells are skipped (the record set is zeroed first).
 */
CA_GLOBAL void ca2d_skew_pack_kernel(Sk2Layout L)
{
    CA_SHARED(uint32_t, tile, SK2_TILE_Y * 8);
    const int lane = dp_lane(), warp = dp_warp_in_block(), nwarp = dp_block_threads() >> 5;
    const int cols = (L.w + 31) / 32;
    const int T = sk2_diagonals(L.w, L.h);
    const int tiles_t = (T + SK2_TILE_T - 1) / SK2_TILE_T;
    const long long ntiles = (long long)cols * tiles_t;
    for (long long tl = dp_block(); tl < ntiles; tl += dp_grid_blocks()) {
        const int j = (int)(tl % cols), t0 = (int)(tl / cols) * SK2_TILE_T;
        const int t1 = t0 + SK2_TILE_T < T ? t0 + SK2_TILE_T : T;
        const int x0 = 32 * j;
        const int ylo = t0 - 2 * (x0 + 31), yhi = t1 - 1 - 2 * x0;     /* grid rows the tile's words read */
        if (yhi < 0 || ylo >= L.h)
            continue;                                                   /* uniform over the CTA */
        const bool vec = (L.w & 3) == 0 && x0 + 32 <= L.w && ((size_t)L.cells & 3) == 0;
        for (int i = dp_thread(); i < SK2_TILE_Y * 8; i += dp_block_threads()) {
            const int y = ylo + (i >> 3), c = (i & 7) * 4;
            uint32_t v = 0u;
            if (y >= 0 && y < L.h && y <= yhi) {
                const uint8_t *row = L.cells + (size_t)y * L.w + x0 + c;
                if (vec) {
                    v = *reinterpret_cast<const uint32_t *>(row);
                } else {
                    for (int q = 0; q < 4; q++)
                        if (x0 + c + q < L.w) v |= (uint32_t)row[q] << (8 * q);
                }
            }
            tile[i] = v;
        }
        dp_syncblock();
        for (int t = t0 + warp; t < t1; t += nwarp) {
            const int y = t - 2 * (x0 + lane);
            uint32_t v = 0u;
            if (y >= ylo && y <= yhi)
                v = (tile[(y - ylo) * 8 + (lane >> 2)] >> (8 * (lane & 3))) & 0xffu;
            const uint32_t word = dp_ballot(v != 0u);
            if (lane == 0)
                L.rows[(size_t)t * SK2_RS + j] = word;
        }
        dp_syncblock();
    }
}

/*
 * unpack: one CTA = 32 columns x SK2_TILE_T grid rows; the words (t, j), t = 2x + y, it needs (SK2_TILE_T + 62 of one
 * word column) are staged in shared memory, every thread writes 4 consecutive cells of a grid row.
 */
CA_GLOBAL void ca2d_skew_unpack_kernel(Sk2Layout L)
{
    CA_SHARED(uint32_t, words, SK2_TILE_Y);
    const int lane = dp_lane();
    const int cols = (L.w + 31) / 32;
    const int tiles_y = (L.h + SK2_TILE_T - 1) / SK2_TILE_T;
    const long long ntiles = (long long)cols * tiles_y;
    unsigned long long pop = 0;
    for (long long tl = dp_block(); tl < ntiles; tl += dp_grid_blocks()) {
        const int j = (int)(tl % cols), y0 = (int)(tl / cols) * SK2_TILE_T;
        const int y1 = y0 + SK2_TILE_T < L.h ? y0 + SK2_TILE_T : L.h;
        const int x0 = 32 * j;
        const int tlo = 2 * x0 + y0;                                    /* .. 2 (x0 + 31) + y1 - 1 */
        const int nt = 62 + (y1 - y0);
        for (int i = dp_thread(); i < nt; i += dp_block_threads())
            words[i] = L.rows[(size_t)(tlo + i) * SK2_RS + j];
        dp_syncblock();
        const bool vec = (L.w & 3) == 0 && x0 + 32 <= L.w && ((size_t)L.cells & 3) == 0;
        for (int i = dp_thread(); i < (y1 - y0) * 8; i += dp_block_threads()) {
            const int y = y0 + (i >> 3), c = (i & 7) * 4;
            uint32_t v = 0u;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int xl = c + q;                                   /* x - x0 */
                const uint32_t bit = (words[2 * xl + (y - y0)] >> xl) & 1u;
                v |= bit << (8 * q);
            }
            uint8_t *row = L.cells + (size_t)y * L.w + x0 + c;
            if (vec) {
                *reinterpret_cast<uint32_t *>(row) = v;
                pop += (unsigned)dp_popc(v);
            } else {
                for (int q = 0; q < 4; q++)
                    if (x0 + c + q < L.w) {
                        row[q] = (uint8_t)(v >> (8 * q));
                        pop += (v >> (8 * q)) & 1u;
                    }
            }
        }
        dp_syncblock();
    }
    for (int o = 16; o; o >>= 1)
        pop += ((unsigned long long)dp_shfl_down((uint32_t)(pop >> 32), o) << 32) | dp_shfl_down((uint32_t)pop, o);
    if (lane == 0 && pop)
        dp_atomic_add64(L.population, pop);
}

} // namespace clapca
#endif
