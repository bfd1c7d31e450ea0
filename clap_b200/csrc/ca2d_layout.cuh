/*
 * ca2d_layout.cuh -- conversion between the reference's 2D grid (uint8, index y*w + x,
 * core/xyarray.c:43 with d2 == 1) and the transposed bit-plane row records of the 2D engine
 * (ca2d_bitplane.cuh): record x = [S0 .. S(P-1)], bit i of word wy = bit q of cell (x, y = 32 wy + i).
 * The transposition is what makes the reference's x-outer / y-inner sweep (core/ca2d.c:65-66) run along
 * memory.  One warp converts a tile of 128 x-columns by 32 y-rows: lanes read/write 4 consecutive cells
 * of a grid row (coalesced 128-byte rows) and own the words of their 4 columns.
 */
#ifndef CLAPCA_CA2D_LAYOUT_CUH
#define CLAPCA_CA2D_LAYOUT_CUH

#include "devport.h"

namespace clapca {

struct Bp2Layout {
    uint8_t *cells;         /* reference layout, y*w + x */
    uint32_t *rows;         /* [w][P][RWS] */
    int w, h, P, RWS;
    unsigned long long *population;     /* unpack: number of non-zero cells */
};

CA_DEV uint32_t lay2_load4(const uint8_t *row, int x, int w, bool vec)
{
    if (vec)
        return *reinterpret_cast<const uint32_t *>(row + x);
    uint32_t v = 0u;
    for (int c = 0; c < 4; c++)
        if (x + c < w) v |= (uint32_t)row[x + c] << (8 * c);
    return v;
}

CA_GLOBAL void ca2d_pack_kernel(Bp2Layout L)
{
    const int lane = dp_lane();
    const int tiles_x = (L.w + 127) / 128, tiles_y = (L.h + 31) / 32;
    const long long ntiles = (long long)tiles_x * tiles_y;
    const long long nwarps = (long long)dp_grid_blocks() * (dp_block_threads() >> 5);
    const bool vec = (L.w & 3) == 0;
    for (long long t = (long long)dp_block() * (dp_block_threads() >> 5) + dp_warp_in_block(); t < ntiles; t += nwarps) {
        const int ty = (int)(t / tiles_x), tx = (int)(t % tiles_x);
        const int x = tx * 128 + lane * 4;
        uint32_t acc[4][8];
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int q = 0; q < 8; q++) acc[c][q] = 0u;
        if (x < L.w) {
#pragma unroll 4
            for (int i = 0; i < 32; i++) {
                const int y = ty * 32 + i;
                if (y >= L.h) break;
                const uint32_t v = lay2_load4(L.cells + (size_t)y * L.w, x, L.w, vec);
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int q = 0; q < 8; q++)
                        if (q < L.P) acc[c][q] |= ((v >> (8 * c + q)) & 1u) << i;
            }
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (x + c < L.w)
#pragma unroll
                    for (int q = 0; q < 8; q++)
                        if (q < L.P) L.rows[((size_t)(x + c) * L.P + q) * L.RWS + ty] = acc[c][q];
        }
    }
}

/* row records -> uint8 grid, and the population count of the result */
CA_GLOBAL void ca2d_unpack_kernel(Bp2Layout L)
{
    const int lane = dp_lane();
    const int tiles_x = (L.w + 127) / 128, tiles_y = (L.h + 31) / 32;
    const long long ntiles = (long long)tiles_x * tiles_y;
    const long long nwarps = (long long)dp_grid_blocks() * (dp_block_threads() >> 5);
    const bool vec = (L.w & 3) == 0;
    unsigned long long pop = 0;
    for (long long t = (long long)dp_block() * (dp_block_threads() >> 5) + dp_warp_in_block(); t < ntiles; t += nwarps) {
        const int ty = (int)(t / tiles_x), tx = (int)(t % tiles_x);
        const int x = tx * 128 + lane * 4;
        if (x >= L.w) continue;
        uint32_t acc[4][8];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t any = 0u;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                acc[c][q] = (q < L.P && x + c < L.w) ? L.rows[((size_t)(x + c) * L.P + q) * L.RWS + ty] : 0u;
                any |= acc[c][q];
            }
            pop += (unsigned)dp_popc(any);
        }
#pragma unroll 4
        for (int i = 0; i < 32; i++) {
            const int y = ty * 32 + i;
            if (y >= L.h) break;
            uint32_t v = 0u;
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (q < L.P) v |= ((acc[c][q] >> i) & 1u) << (8 * c + q);
            uint8_t *row = L.cells + (size_t)y * L.w;
            if (vec) {
                *reinterpret_cast<uint32_t *>(row + x) = v;
            } else {
                for (int c = 0; c < 4; c++)
                    if (x + c < L.w) row[x + c] = (uint8_t)(v >> (8 * c));
            }
        }
    }
    for (int o = 16; o; o >>= 1)
        pop += ((unsigned long long)dp_shfl_down((uint32_t)(pop >> 32), o) << 32) | dp_shfl_down((uint32_t)pop, o);
    if (lane == 0 && pop)
        dp_atomic_add64(L.population, pop);
}

/* ---- device-side seeding: ca2d_generate()'s fill loop, core/ca2d.c:86-90 ------------------------------- */

/*
 * glibc lrand48(): X' = (a X + c) mod 2^48 with a = 0x5DEECE66D, c = 0xB, result X' >> 17.  The affine map
 * composes, so the state n steps ahead is G X + C with (G, C) from O(log n) squarings (Brown's arbitrary-stride
 * jump): every cell can start from the same 48-bit state the host's stream is in.
 */
struct Rand48Jump { unsigned long long G, C; };

CA_HOSTDEV Rand48Jump r48_jump(unsigned long long n)
{
    const unsigned long long M = (1ull << 48) - 1;
    unsigned long long h = 0x5DEECE66Dull, f = 0xBull, G = 1ull, C = 0ull;
    while (n) {
        if (n & 1ull) {
            G = (G * h) & M;
            C = (C * h + f) & M;
        }
        f = (f * (h + 1ull)) & M;
        h = (h * h) & M;
        n >>= 1;
    }
    Rand48Jump j = { G, C };
    return j;
}

CA_HOSTDEV unsigned long long r48_advance(unsigned long long x, unsigned long long n)
{
    const Rand48Jump j = r48_jump(n);
    return (j.G * x + j.C) & ((1ull << 48) - 1);
}

#ifndef CLAPCA_EMU
/*
 * The reference draws one lrand48() % 8 per cell in x-OUTER / y-inner order and stores cell (x,y) at
 * arr[y*w + x]: draw number k = x*side + y.  A thread jumps to the first draw of a run of 32 consecutive y of one
 * x and iterates; the CTA's 32 (x) x 256 (y) tile goes through shared memory so the stores run along x.
 */
__global__ void __launch_bounds__(256) ca2d_seed_kernel(uint8_t *cells, int w, int side, uint32_t nr_states,
                                                        unsigned long long x0)
{
    __shared__ __align__(16) uint8_t tile[256][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xb = blockIdx.x * 32, yb = blockIdx.y * 256;
    const int x = xb + lane, ys = yb + 32 * warp;
    const unsigned long long M = (1ull << 48) - 1;
    if (x < side && ys < side) {
        unsigned long long st = r48_advance(x0, (unsigned long long)x * side + ys);
        for (int s = 0; s < 32; s++) {
            st = (st * 0x5DEECE66Dull + 0xBull) & M;
            const uint32_t v = (uint32_t)(st >> 17) & 7u;               /* lrand48() % 8 */
            tile[32 * warp + s][lane] = v <= nr_states ? (uint8_t)nr_states : (uint8_t)0;
        }
    }
    __syncthreads();
    const int y = yb + threadIdx.x;
    if (y >= side)
        return;
    uint8_t *dst = cells + (size_t)y * w + xb;
    if (xb + 32 <= side && (reinterpret_cast<size_t>(dst) & 15) == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(tile[threadIdx.x]);
        reinterpret_cast<uint4 *>(dst)[0] = src[0];
        reinterpret_cast<uint4 *>(dst)[1] = src[1];
    } else {
        for (int i = 0; i < 32 && xb + i < side; i++)
            dst[i] = tile[threadIdx.x][i];
    }
}
#endif

} // namespace clapca
#endif
