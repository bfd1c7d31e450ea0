/*
 * ca_wavefront.cuh -- uint8 "cell wavefront" engines: the simplest schedule
 * that is provably identical to the reference's in-place sweeps, for ANY rule,
 * neighbourhood and shape.  Used for the value-comparing 2D neighbourhoods
 * (ca2d_neigh_vnv / ca2d_neigh_mv with decay), for partial sweeps (side < w),
 * for shapes the bit-plane engines do not cover, and as an independent
 * on-device cross-check of the bit-plane engines.
 *
 * Schedule.  Give every cell update a time stamp
 *     3D: tau = x + 2y + 4z + 8g        (core/ca3d.c:129-140: z, y, x order)
 *     2D: tau = 2x + y + 4g             (core/ca2d.c:65-66:   x outer, y inner)
 * All neighbours the reference reads in their NEW state have a smaller tau
 * within the same generation, all neighbours read in their OLD state were
 * last written at a smaller tau (generation g-1) and are next written at a
 * larger one, and no two cells with the same tau are neighbours.  So the
 * update can run in place, one grid-wide barrier per tau, several
 * generations in flight at once.
 */
#ifndef CLAPCA_CA_WAVEFRONT_CUH
#define CLAPCA_CA_WAVEFRONT_CUH

#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

namespace clapca {

namespace cg = cooperative_groups;

struct Wf3Params {
    uint8_t *a;
    long long d0, d1, d2;       /* x fastest */
    uint32_t surv, born, bornval;
    int G;
};

__device__ __forceinline__ int wf_cell(const uint8_t *a, long long idx)
{
    return __ldcg(a + idx);     /* written by other SMs within this launch */
}

/* ca3d_run(): core/ca3d.c:124-142 with ca3d_neighbors_m1() (:29-39) */
__global__ void __launch_bounds__(256) ca3d_wavefront_kernel(Wf3Params p)
{
    cg::grid_group grid = cg::this_grid();
    const long long d0 = p.d0, d1 = p.d1, d2 = p.d2;
    const long long pairs = d1 * d2;
    const long long tmax = (d0 - 1) + 2 * (d1 - 1) + 4 * (d2 - 1) + 8LL * (p.G - 1);
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;

    for (long long tau = 0; tau <= tmax; tau++) {
        for (long long i = tid; i < pairs; i += nthreads) {
            const long long y = i % d1, z = i / d1;
            const long long r = tau - 2 * y - 4 * z;            /* = x + 8g */
            if (r < 0)
                continue;
            long long glo = r - (d0 - 1) > 0 ? (r - (d0 - 1) + 7) / 8 : 0;
            long long ghi = r / 8;
            if (ghi > p.G - 1) ghi = p.G - 1;
            for (long long g = glo; g <= ghi; g++) {
                const long long x = r - 8 * g;
                const long long c = (z * d1 + y) * d0 + x;
                int n = 0;
                for (int dz = -1; dz <= 1; dz++) {
                    const long long zz = z + dz;
                    if (zz < 0 || zz >= d2) continue;
                    for (int dy = -1; dy <= 1; dy++) {
                        const long long yy = y + dy;
                        if (yy < 0 || yy >= d1) continue;
                        const long long rowi = (zz * d1 + yy) * d0;
                        if (x > 0)      n += wf_cell(p.a, rowi + x - 1) != 0;
                        if (dz | dy)    n += wf_cell(p.a, rowi + x) != 0;
                        if (x + 1 < d0) n += wf_cell(p.a, rowi + x + 1) != 0;
                    }
                }
                const int s = wf_cell(p.a, c);
                if (s && !((p.surv >> n) & 1u))
                    p.a[c] = (uint8_t)(s - 1);
                else if (!s && ((p.born >> n) & 1u))
                    p.a[c] = (uint8_t)p.bornval;
            }
        }
        grid.sync();
    }
}

struct Wf2Params {
    uint8_t *a;
    long long w, h;             /* array extent, index y*w + x */
    long long sx, sy;           /* swept region: x < sx, y < sy (min(side, extent)) */
    uint32_t born, surv, nrval; /* nrval = (uint8_t)nr_states */
    int decay, neigh, G;
};

/* ca2d_step(): core/ca2d.c:61-77 with the neighbourhoods of :11-59 */
__global__ void __launch_bounds__(256) ca2d_wavefront_kernel(Wf2Params p)
{
    cg::grid_group grid = cg::this_grid();
    const long long w = p.w, h = p.h, sx = p.sx, sy = p.sy;
    const long long tmax = 2 * (sx - 1) + (sy - 1) + 4LL * (p.G - 1);
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool moore = (p.neigh == 1 || p.neigh == 3);
    const bool by_value = (p.neigh == 2 || p.neigh == 3);

    for (long long tau = 0; tau <= tmax; tau++) {
        for (long long x = tid; x < sx; x += nthreads) {
            const long long r = tau - 2 * x;                    /* = y + 4g */
            if (r < 0)
                continue;
            long long glo = r - (sy - 1) > 0 ? (r - (sy - 1) + 3) / 4 : 0;
            long long ghi = r / 4;
            if (ghi > p.G - 1) ghi = p.G - 1;
            for (long long g = glo; g <= ghi; g++) {
                const long long y = r - 4 * g;
                const long long c = y * w + x;
                const int v = wf_cell(p.a, c);
                const int thr = by_value ? v : 0;               /* neighbour counts if value > thr */
                int n = 0;
                for (int dy = -1; dy <= 1; dy++) {
                    const long long yy = y + dy;
                    if (yy < 0 || yy >= h) continue;
                    for (int dx = -1; dx <= 1; dx++) {
                        const long long xx = x + dx;
                        if (xx < 0 || xx >= w) continue;
                        if (!(dx | dy)) continue;
                        if (!moore && dx && dy) continue;
                        n += wf_cell(p.a, yy * w + xx) > thr;
                    }
                }
                if (!v && ((p.born >> n) & 1u))
                    p.a[c] = (uint8_t)p.nrval;
                else if (v && ((p.surv >> n) & 1u))
                    ;
                else if (v && p.decay)
                    p.a[c] = (uint8_t)(v - 1);
            }
        }
        grid.sync();
    }
}

/* xyzarray_count(): core/xyarray.c:68-78 */
__global__ void __launch_bounds__(256) count_nonzero_kernel(const uint8_t *a, size_t n, unsigned long long *out)
{
    unsigned long long c = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n16 = ((uintptr_t)a % 16 == 0) ? n / 16 : 0;
    const uint4 *a4 = reinterpret_cast<const uint4 *>(a);
    for (size_t i = tid; i < n16; i += stride) {
        uint4 v = a4[i];
        uint32_t wds[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int k = 0; k < 4; k++) {
            /* count non-zero bytes: OR-fold each byte onto its lsb */
            uint32_t t = wds[k];
            t |= t >> 4; t |= t >> 2; t |= t >> 1;
            c += __popc(t & 0x01010101u);
        }
    }
    for (size_t i = n16 * 16 + tid; i < n; i += stride)
        c += a[i] != 0;
    for (int o = 16; o; o >>= 1)
        c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c)
        atomicAdd(out, c);
}

/*
 * Plane fingerprint: h = sum over the plane's 8-byte little-endian words w_k (zero-padded tail) of
 * splitmix64(w_k ^ (k + 1) * 0x9E3779B97F4A7C15) mod 2^64.  Order-independent, so it reduces in parallel; the
 * position enters every term, so permuted or shifted cells change it.  clap_b200/synth.py holds the numpy twin.
 */
__device__ __forceinline__ unsigned long long plane_hash_mix(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) plane_hash_kernel(const uint8_t *a, size_t plane_bytes, size_t nplanes,
                                                         unsigned long long *out)
{
    const size_t words = (plane_bytes + 7) / 8, full = plane_bytes / 8;
    for (size_t pl = blockIdx.y; pl < nplanes; pl += gridDim.y) {
        const uint8_t *base = a + pl * plane_bytes;
        const bool aligned = ((uintptr_t)base & 7) == 0;
        unsigned long long h = 0;
        for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < words; k += (size_t)gridDim.x * blockDim.x) {
            unsigned long long w = 0;
            if (k < full && aligned) {
                w = *reinterpret_cast<const unsigned long long *>(base + 8 * k);
            } else {
                for (int i = 0; i < 8; i++)
                    if (8 * k + i < plane_bytes)
                        w |= (unsigned long long)base[8 * k + i] << (8 * i);
            }
            h += plane_hash_mix(w ^ ((unsigned long long)(k + 1) * 0x9E3779B97F4A7C15ull));
        }
        for (int o = 16; o; o >>= 1)
            h += __shfl_down_sync(0xffffffffu, h, o);
        if ((threadIdx.x & 31) == 0 && h)
            atomicAdd(out + pl, h);
    }
}

/* largest cell value (chooses the number of bit planes) */
__global__ void __launch_bounds__(256) max_u8_kernel(const uint8_t *a, size_t n, unsigned *out)
{
    unsigned m = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n16 = ((uintptr_t)a % 16 == 0) ? n / 16 : 0;
    const uint4 *a4 = reinterpret_cast<const uint4 *>(a);
    for (size_t i = tid; i < n16; i += stride) {
        uint4 v = a4[i];
        uint32_t t = __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w));
        t = __vmaxu4(t, t >> 16);
        t = __vmaxu4(t, t >> 8);
        m = max(m, t & 0xffu);
    }
    for (size_t i = n16 * 16 + tid; i < n; i += stride)
        m = max(m, (unsigned)a[i]);
    for (int o = 16; o; o >>= 1)
        m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m)
        atomicMax(out, m);
}

} // namespace clapca
#endif
