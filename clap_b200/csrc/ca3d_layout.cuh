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
 * 32 cells held in r[8] (4 per register, x ascending) -> P state words + alive word.  P covers every value that
 * occurs (the caller derives it from the maximum of the volume), so alive = OR of the state planes: no separate
 * "byte != 0" reduction -- the kernel is bound by the integer pipe, and that reduction was 40 % of its work.
 */
CA_DEV void lay_pack32(const uint32_t r[8], int P, uint32_t s[8], uint32_t &alive)
{
#pragma unroll
    for (int q = 0; q < 8; q++) s[q] = 0u;
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (q < P)
                s[q] |= lay_gather4(r[j] >> q) << (4 * j);
    }
    alive = 0u;
#pragma unroll
    for (int q = 0; q < 8; q++)
        if (q < P)
            alive |= s[q];
}

/* uint8 volume -> row records (one thread per word of a plane-row) */
CA_GLOBAL void ca3d_pack_kernel(Bp3Layout L)
{
    const int NP = L.P + 2;
    const size_t nwords = (size_t)L.Z * L.H * L.RWP;
    const size_t stride = (size_t)dp_grid_blocks() * dp_block_threads();
    for (size_t i = (size_t)dp_block() * dp_block_threads() + dp_thread(); i < nwords; i += stride) {
        const int w = (int)(i % L.RWP);
        const size_t row = i / L.RWP;
        uint32_t *rec = L.rows + row * NP * L.RWP + w;
        const uint8_t *src = L.cells + row * L.W;
        const int x0 = 32 * w;
        uint32_t s[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        uint32_t alive = 0u;
        int n = L.W - x0;
        n = n > 32 ? 32 : n;
        if (n == 32 && (L.W & 15) == 0) {
            /* whole word, 16-byte aligned rows: two 128-bit loads, multiply-gather per plane */
            const uint4 lo = *reinterpret_cast<const uint4 *>(src + x0);
            const uint4 hi = *reinterpret_cast<const uint4 *>(src + x0 + 16);
            const uint32_t r[8] = { lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w };
            lay_pack32(r, L.P, s, alive);
        } else {
            for (int i2 = 0; i2 < n; i2++) {
                uint32_t v = src[x0 + i2];
#pragma unroll
                for (int q = 0; q < 8; q++) s[q] |= ((v >> q) & 1u) << i2;
                alive |= (uint32_t)(v != 0) << i2;
            }
        }
        uint32_t left  = (n > 0 && x0 > 0) ? (uint32_t)(src[x0 - 1] != 0) : 0u;
        uint32_t right = (n > 0 && x0 + 32 < L.W) ? (uint32_t)(src[x0 + 32] != 0) : 0u;
        uint32_t l = (alive << 1) | left, r = (alive >> 1) | (right << 31);
        rec[0] = l ^ alive ^ r;
        rec[(size_t)L.RWP] = (l & alive) | (l & r) | (alive & r);
        for (int q = 0; q < L.P; q++)
            rec[(size_t)(2 + q) * L.RWP] = s[q];
    }
}

/* row records -> uint8 volume, and the population count of the result */
CA_GLOBAL void ca3d_unpack_kernel(Bp3Layout L)
{
    const int NP = L.P + 2;
    const int RW = (L.W + 31) / 32;
    const size_t nwords = (size_t)L.Z * L.H * RW;
    const size_t stride = (size_t)dp_grid_blocks() * dp_block_threads();
    unsigned long long pop = 0;
    for (size_t i = (size_t)dp_block() * dp_block_threads() + dp_thread(); i < nwords; i += stride) {
        const int w = (int)(i % RW);
        const size_t row = i / RW;
        const uint32_t *rec = L.rows + row * NP * L.RWP + w;
        uint8_t *dst = L.cells + row * L.W;
        const int x0 = 32 * w;
        uint32_t s[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        uint32_t alive = 0u;
        for (int q = 0; q < L.P; q++) {
            s[q] = rec[(size_t)(2 + q) * L.RWP];
            alive |= s[q];
        }
        pop += (unsigned)dp_popc(alive);
        int n = L.W - x0;
        n = n > 32 ? 32 : n;
        if (n == 32 && (L.W & 15) == 0) {
            uint32_t r[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t v = 0u;
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (q < L.P)
                        v |= lay_spread4(s[q] >> (4 * j)) << q;
                r[j] = v;
            }
            *reinterpret_cast<uint4 *>(dst + x0) = make_uint4(r[0], r[1], r[2], r[3]);
            *reinterpret_cast<uint4 *>(dst + x0 + 16) = make_uint4(r[4], r[5], r[6], r[7]);
        } else {
            for (int i2 = 0; i2 < n; i2++) {
                uint32_t v = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) v |= ((s[q] >> i2) & 1u) << q;
                dst[x0 + i2] = (uint8_t)v;
            }
        }
    }
    if (pop)
        dp_atomic_add64(L.population, pop);
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
