/*
 * field_kernels.cuh -- per-lattice-point field evaluation:
 *   N1: noise.c fBm-gradient bake   (core/noise.h:9-17, core/noise.c:171-270)
 *   N2: terrain.c heightmap chain   (core/terrain.c:15-91, :447-467)
 *
 * Numerics.  The reference's interp.h helpers mix float arguments with double
 * literals, so parts of every lerp / smoothstep are evaluated in double and
 * rounded back to float.  The device functions below keep the
 * same expression trees, and this translation unit is compiled with
 * -fmad=false so nothing is contracted into an FMA: hash31, the value noise,
 * the fBm sum, the gradient and the RGBA8 packing are then bit-identical to
 * the x86-64 build of the reference (floorf, lrintf, sqrtf and IEEE division
 * are exactly rounded on both sides).  The heightmap differs only through
 * cosf()/powf(), whose CUDA implementations are within 2 ulp of glibc's.
 *
 * Both kernels are embarrassingly parallel and ALU-bound: one thread per
 * output element, consecutive threads on consecutive addresses so stores
 * coalesce into full 128-byte lines.
 */
#ifndef CLAPCA_FIELD_KERNELS_CUH
#define CLAPCA_FIELD_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

namespace clapca {

#define CLAPCA_PI 3.14159265358979323846   /* M_PI */

/* smoothf(): core/interp.h:11-14 */
__device__ __forceinline__ float fk_smoothf(float x)
{
    float xx = x * x;
    return (float)((double)xx * (3.0 - 2.0 * (double)x));
}

/* linf_interp(): core/interp.h:25-29 */
__device__ __forceinline__ float fk_linf(float a, float b, float blend)
{
    float bb = b * blend;
    return (float)((double)a * (1.0 - (double)blend) + (double)bb);
}

/* cosf_interp(): core/interp.h:35-42, split into the blend factor and the mix so the factor can be tabulated */
__device__ __forceinline__ float fk_cos_factor(float blend)
{
    float theta = (float)((double)blend * CLAPCA_PI);
    return (float)((1.0 - (double)cosf(theta)) / 2.0);
}

__device__ __forceinline__ float fk_cos_mix(float a, float b, float f)
{
    float bf = b * f;
    return (float)((double)a * (1.0 - (double)f) + (double)bf);
}

__device__ __forceinline__ float fk_cosf_interp(float a, float b, float blend)
{
    return fk_cos_mix(a, b, fk_cos_factor(blend));
}

/* hash31(): core/noise.h:9-17 */
__device__ __forceinline__ float fk_hash31(int x, int y, int z, uint32_t seed)
{
    uint32_t h = (uint32_t)x * 374761393u + (uint32_t)y * 668265263u +
                 (uint32_t)z * 362437u + seed * 2246822519u;
    h = (h ^ (h >> 13)) * 1274126177u;
    return (float)(h ^ (h >> 16)) * (1.0f / 4294967296.0f);
}

/*
 * ((v % period) + period) % period of the reference (core/noise.c:177-185) without the two signed divisions: lattice
 * coordinates of a periodic field lie within one period of [0, period) (the bake samples [-eps, period_units + eps]),
 * where one conditional add / subtract gives the same number; anything further out takes the division.
 */
__device__ __forceinline__ int fk_wrap(int v, int period)
{
    if (v >= 0 && v < period)
        return v;
    if (v < 0 && v >= -period)
        return v + period;
    if (v >= period && v - period < period)
        return v - period;
    return (v % period + period) % period;
}

/* value_noise3d_periodic(): core/noise.c:171-202 */
__device__ __forceinline__ float fk_value_noise3d(float x, float y, float z, int period, uint32_t seed)
{
    int x0 = (int)floorf(x), y0 = (int)floorf(y), z0 = (int)floorf(z);
    float xf = x - (float)x0, yf = y - (float)y0, zf = z - (float)z0;
    /* wrap(v + 1) == wrap(wrap(v) + 1): the upper corner follows from the wrapped lower one */
    x0 = fk_wrap(x0, period);
    y0 = fk_wrap(y0, period);
    z0 = fk_wrap(z0, period);
    const int x1 = x0 + 1 == period ? 0 : x0 + 1, y1 = y0 + 1 == period ? 0 : y0 + 1, z1 = z0 + 1 == period ? 0 : z0 + 1;

    float ux = fk_smoothf(xf), uy = fk_smoothf(yf), uz = fk_smoothf(zf);
    float lo0 = fk_linf(fk_hash31(x0, y0, z0, seed), fk_hash31(x1, y0, z0, seed), ux);
    float lo1 = fk_linf(fk_hash31(x0, y1, z0, seed), fk_hash31(x1, y1, z0, seed), ux);
    float hi0 = fk_linf(fk_hash31(x0, y0, z1, seed), fk_hash31(x1, y0, z1, seed), ux);
    float hi1 = fk_linf(fk_hash31(x0, y1, z1, seed), fk_hash31(x1, y1, z1, seed), ux);
    return fk_linf(fk_linf(lo0, lo1, uy), fk_linf(hi0, hi1, uy), uz);
}

/* fbm3_periodic(): core/noise.c:204-220 */
__device__ __forceinline__ float fk_fbm3(float x, float y, float z, int octaves, float lacunarity, float gain,
                                         int period, uint32_t seed)
{
    float amp = 0.5f, sum = 0.0f;
    for (int i = 0; i < octaves; i++) {
        sum += fk_value_noise3d(x, y, z, period, seed + (uint32_t)i) * amp;
        x *= lacunarity;
        y *= lacunarity;
        z *= lacunarity;
        period = (int)lrintf((float)period * lacunarity);
        amp *= gain;
    }
    return sum;
}

__device__ __forceinline__ uint32_t fk_unorm8(float g)
{
    return (uint32_t)(uint8_t)lrintf((g * 0.5f + 0.5f) * 255.0f);
}

struct NoiseBakeParams {
    uint32_t *out;          /* RGBA8 texels, x fastest (linear bake); unused by the array bake */
    unsigned size;
    int octaves;
    float lacunarity, gain, period_units;
    uint32_t seed;
    cudaSurfaceObject_t surf;   /* array bake: a 3D CUDA array of uchar4 bound as a surface */
    const float *field;     /* lattice bake: fBm at the (size + 2)^3 lattice points -1 .. size, x fastest (else nullptr) */
};

/* one voxel of noise_grad3d_bake_rgba8(): core/noise.c:236-264 -- central differences of the fBm, normalised, packed */
__device__ __forceinline__ uint32_t fk_bake_voxel(const NoiseBakeParams &p, unsigned x, unsigned y, unsigned z)
{
    const float step = p.period_units / (float)p.size;
    const float eps = step;
    const int period = (int)p.period_units;
    const float scale = 0.5f / eps;
    const float px = (float)x * step, py = (float)y * step, pz = (float)z * step;

    float gx = (fk_fbm3(px + eps, py, pz, p.octaves, p.lacunarity, p.gain, period, p.seed) -
                fk_fbm3(px - eps, py, pz, p.octaves, p.lacunarity, p.gain, period, p.seed)) * scale;
    float gy = (fk_fbm3(px, py + eps, pz, p.octaves, p.lacunarity, p.gain, period, p.seed) -
                fk_fbm3(px, py - eps, pz, p.octaves, p.lacunarity, p.gain, period, p.seed)) * scale;
    float gz = (fk_fbm3(px, py, pz + eps, p.octaves, p.lacunarity, p.gain, period, p.seed) -
                fk_fbm3(px, py, pz - eps, p.octaves, p.lacunarity, p.gain, period, p.seed)) * scale;
    float len2 = gx * gx + gy * gy + gz * gz;
    float inv = 1.0f / sqrtf(len2 > FLT_MIN ? len2 : FLT_MIN);

    return fk_unorm8(gx * inv) | (fk_unorm8(gy * inv) << 8) | (fk_unorm8(gz * inv) << 16);
}

/*
 * The lattice bake.  A voxel's six fBm samples sit at p +- eps along one axis with eps = step -- the positions of its
 * NEIGHBOURS' centres whenever float arithmetic makes fl(fl(x step) + step) == fl((x + 1) step) and
 * fl(fl(x step) - step) == fl((x - 1) step) for every x (true for the engine's own bake: 37 / 256 is a dyadic
 * rational; the host checks all `size` values before choosing this path).  Then every sample point is a lattice point
 * k in [-1, size]^3, the fBm is evaluated ONCE per lattice point instead of six times per voxel, and the voxel is the
 * central difference of stored neighbours: bit-identical output, a sixth of the work (the bake is bound by the
 * float <-> double conversions of interp.h's promoted arithmetic: ncu, mio_throttle 7 per issue).
 */
__global__ void __launch_bounds__(256) noise_field_kernel(NoiseBakeParams p, float *field)
{
    const unsigned S2 = p.size + 2u;
    const size_t points = (size_t)S2 * S2 * S2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float step = p.period_units / (float)p.size;
    const int period = (int)p.period_units;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < points; i += stride) {
        const int kx = (int)(i % S2) - 1, ky = (int)((i / S2) % S2) - 1, kz = (int)(i / ((size_t)S2 * S2)) - 1;
        /* only the axis neighbours of voxels are ever read: skip lattice points outside [0, size) in two axes */
        const int out = (kx < 0 || kx >= (int)p.size) + (ky < 0 || ky >= (int)p.size) + (kz < 0 || kz >= (int)p.size);
        if (out > 1)
            continue;
        field[i] = fk_fbm3((float)kx * step, (float)ky * step, (float)kz * step, p.octaves, p.lacunarity, p.gain, period,
                           p.seed);
    }
}

/* one voxel from the stored lattice: core/noise.c:246-264 */
__device__ __forceinline__ uint32_t fk_bake_voxel_lattice(const NoiseBakeParams &p, unsigned x, unsigned y, unsigned z)
{
    const size_t S2 = p.size + 2u;
    const float step = p.period_units / (float)p.size;
    const float scale = 0.5f / step;
    const float *c = p.field + ((size_t)(z + 1u) * S2 + (y + 1u)) * S2 + (x + 1u);
    float gx = (c[1] - c[-1]) * scale;
    float gy = (c[S2] - c[-(ptrdiff_t)S2]) * scale;
    float gz = (c[S2 * S2] - c[-(ptrdiff_t)(S2 * S2)]) * scale;
    float len2 = gx * gx + gy * gy + gz * gz;
    float inv = 1.0f / sqrtf(len2 > FLT_MIN ? len2 : FLT_MIN);

    return fk_unorm8(gx * inv) | (fk_unorm8(gy * inv) << 8) | (fk_unorm8(gz * inv) << 16);
}

/* noise_grad3d_bake_rgba8(): core/noise.c:222-270, one thread per voxel, into a linear buffer */
__global__ void __launch_bounds__(256) noise_bake_kernel(NoiseBakeParams p)
{
    const size_t voxels = (size_t)p.size * p.size * p.size;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < voxels; i += stride) {
        const unsigned x = (unsigned)(i % p.size), y = (unsigned)((i / p.size) % p.size);
        const unsigned z = (unsigned)(i / ((size_t)p.size * p.size));
        p.out[i] = p.field ? fk_bake_voxel_lattice(p, x, y, z) : fk_bake_voxel(p, x, y, z);
    }
}

/*
 * The same bake straight into a 3D CUDA array through a surface: the array is what the renderer's 3D
 * texture maps under CUDA / GL or Vulkan interop (noise_grad3d_bake_rgba8_tex, core/noise.c:272-294, uploads the host
 * copy with texture_load) -- the texels never visit the host.
 */
__global__ void __launch_bounds__(256) noise_bake_surface_kernel(NoiseBakeParams p)
{
    const size_t voxels = (size_t)p.size * p.size * p.size;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < voxels; i += stride) {
        const unsigned x = (unsigned)(i % p.size), y = (unsigned)((i / p.size) % p.size);
        const unsigned z = (unsigned)(i / ((size_t)p.size * p.size));
        const uint32_t t = p.field ? fk_bake_voxel_lattice(p, x, y, z) : fk_bake_voxel(p, x, y, z);
        surf3Dwrite(make_uchar4((unsigned char)t, (unsigned char)(t >> 8), (unsigned char)(t >> 16), 0), p.surf,
                    (int)(x * sizeof(uchar4)), (int)y, (int)z);
    }
}

/* fbm3_periodic() at arbitrary points (parity tests of the float field) */
__global__ void noise_fbm3_kernel(float *out, const float *xyz, size_t n, int octaves, float lacunarity,
                                  float gain, int period, uint32_t seed)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = fk_fbm3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], octaves, lacunarity, gain, period, seed);
}

/* ---- terrain ------------------------------------------------------------------ */

/*
 * get_rand_height(): core/terrain.c:15-19 = glibc srand48(seed ^ (x + z*43210))
 * followed by drand48(): X0 = low32(s) << 16 | 0x330E, X1 = (a X0 + c) mod 2^48,
 * value = X1 / 2^48 (exact in double), then *2 - 1 and a float rounding.
 */
__device__ __forceinline__ float fk_rand_height(long long seed, int x, int z)
{
    long long s = seed ^ (long long)(x + z * 43210);
    unsigned long long X = (((unsigned long long)s & 0xffffffffULL) << 16) | 0x330EULL;
    X = (X * 0x5DEECE66DULL + 0xBULL) & 0xFFFFFFFFFFFFULL;
    double d = (double)X * (1.0 / 281474976710656.0);       /* 2^-48, exact */
    return (float)(d * 2 - 1);
}

/* lattice fill: core/terrain.c:447-450, map0[x*nr_v + z] */
__global__ void __launch_bounds__(256) terrain_map0_kernel(float *map0, long long seed, unsigned nr_v)
{
    const size_t n = (size_t)nr_v * nr_v;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        map0[i] = fk_rand_height(seed, (int)(i / nr_v), (int)(i % nr_v));
}

struct TerrainParams {
    float *map;
    const float *map0;
    const uint8_t *maze;        /* mside x mside, index y*mside + x; NULL = plain octave field */
    unsigned nr_v, mside;
    float ty, amp;
    int oct;
    const float *smooth;        /* (nr_v+1)^2 table of get_avg_height(), or NULL: evaluate it per use */
};

/* get_mapped_rand_height(): core/terrain.c:21-33 */
__device__ __forceinline__ float fk_lattice(const TerrainParams &p, int x, int z)
{
    const int nr = (int)p.nr_v;
    if (x < 0) x = nr - 1; else if (x >= nr) x = 0;
    if (z < 0) z = nr - 1; else if (z >= nr) z = 0;
    return __ldg(p.map0 + (size_t)x * nr + z);
}

/* get_avg_height(): core/terrain.c:35-54 */
__device__ __forceinline__ float fk_smooth3x3(const TerrainParams &p, int x, int z)
{
    float corners, sides, self;
    corners  = fk_lattice(p, x - 1, z - 1);
    corners += fk_lattice(p, x + 1, z - 1);
    corners += fk_lattice(p, x - 1, z + 1);
    corners += fk_lattice(p, x + 1, z + 1);
    corners /= 16.f;
    sides  = fk_lattice(p, x - 1, z);
    sides += fk_lattice(p, x + 1, z);
    sides += fk_lattice(p, x, z - 1);
    sides += fk_lattice(p, x, z + 1);
    sides /= 8.f;
    self = fk_lattice(p, x, z) / 4.f;
    return corners + sides + self;
}

/*
 * get_avg_height() is a pure function of the integer lattice point, and get_height() only ever asks for
 * points 0..nr_v (x * freq with freq <= 1, plus one): tabulating it once -- same float operations in the
 * same order, so the same bits -- turns the 36 lattice reads of an interpolation into 4 table reads.
 */
__global__ void __launch_bounds__(256) terrain_smooth_kernel(TerrainParams p, float *smooth)
{
    const unsigned side = p.nr_v + 1;
    const size_t n = (size_t)side * side;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        smooth[i] = fk_smooth3x3(p, (int)(i / side), (int)(i % side));
}

__device__ __forceinline__ float fk_avg_height(const TerrainParams &p, int x, int z)
{
    if (p.smooth)
        return __ldg(p.smooth + (size_t)x * (p.nr_v + 1) + z);
    return fk_smooth3x3(p, x, z);
}

/* get_interp_height(): core/terrain.c:56-71 */
__device__ __forceinline__ float fk_interp_height(const TerrainParams &p, float x, float z)
{
    int ix = (int)floor((double)x), iz = (int)floor((double)z);
    float fx = x - (float)ix, fz = z - (float)iz;
    float v1 = fk_avg_height(p, ix, iz), v2 = fk_avg_height(p, ix + 1, iz);
    float v3 = fk_avg_height(p, ix, iz + 1), v4 = fk_avg_height(p, ix + 1, iz + 1);
    return fk_cosf_interp(fk_cosf_interp(v1, v2, fx), fk_cosf_interp(v3, v4, fx), fz);
}

/* get_height(): core/terrain.c:77-91; pow(2,i), pow(0.5f,i) are exact powers of two */
__device__ __forceinline__ float fk_octave_height(const TerrainParams &p, int x, int z, float amp0, int oct)
{
    float total = 0;
    float d = (float)ldexp(1.0, oct - 1);
    for (int i = 0; i < oct; i++) {
        float freq = (float)(ldexp(1.0, i) / (double)d);
        float amp = (float)(ldexp(1.0, -i) * (double)amp0);
        total += fk_interp_height(p, (float)x * freq, (float)z * freq) * amp;
    }
    return p.ty + total;
}

__device__ __forceinline__ int fk_maze(const TerrainParams &p, int x, int y)
{
    if (x < 0 || y < 0 || x >= (int)p.mside || y >= (int)p.mside)
        return 0;                                       /* xyarray_get(): OOB reads 0 */
    return p.maze[(size_t)y * p.mside + x];
}

/*
 * Map fill: core/terrain.c:451-467 (MAZE_FAC 8, OCTAVES 4).  Thread per vertex,
 * j (the contiguous index of map[i*nr_v + j]) fastest across the warp.
 */
__global__ void __launch_bounds__(256) terrain_heightmap_kernel(TerrainParams p)
{
    const size_t n = (size_t)p.nr_v * p.nr_v;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
        const int i = (int)(idx / p.nr_v), j = (int)(idx % p.nr_v);
        float h;
        if (p.maze) {
            float fi = fmodf((float)i, 8.f) / 8, fj = fmodf((float)j, 8.f) / 8;
            int mi = i / 8, mj = j / 8;
            int cn = fk_maze(p, mi, mj);
            int xn = fk_maze(p, (double)fi >= 0.5 ? mi + 1 : mi - 1, mj);
            int yn = fk_maze(p, mi, (double)fj >= 0.5 ? mj + 1 : mj - 1);
            float xa = cn > xn ? (float)cn : fk_cosf_interp((float)cn, (float)xn, 2 * fi - 1);
            float ya = cn > yn ? (float)cn : fk_cosf_interp((float)cn, (float)yn, 2 * fj - 1);
            float avg = fk_cosf_interp(xa, ya, fabsf(fi - fj));
            h = fk_octave_height(p, i, j, powf(1.5f, avg), 4) + avg;
        } else {
            h = fk_octave_height(p, i, j, p.amp, p.oct);
        }
        p.map[idx] = h;
    }
}

/*
 * Table-driven map fill (the default): same arithmetic, far fewer instructions.
 *   - get_avg_height() comes from the smooth table (terrain_smooth_kernel);
 *   - get_height() samples octave i at x * 2^i / 2^(oct-1): the fractional parts are multiples of 1 / 2^(oct-1), and
 *     the maze blend fractions are multiples of 1/8 (MAZE_FAC), so every cosf() of the fill has one of a handful of
 *     arguments.  The blend factors are tabulated per CTA in shared memory by the SAME device function the direct
 *     kernel calls (fk_cos_factor), hence the same bits; only powf(1.5, avg) is still evaluated per vertex.
 * D = 2^(oct-1) <= 8 (the reference fixes OCTAVES = 4, terrain.c:73).
 */
__global__ void __launch_bounds__(256) terrain_heightmap_tab_kernel(TerrainParams p)
{
    __shared__ float f_oct[8];          /* factor(k / D) */
    __shared__ float f_frac[8];         /* factor(k / 8): |fi - fj| */
    __shared__ float f_edge[8];         /* factor(2 * (k / 8) - 1): blend toward the x / y maze neighbour */
    const int oct = p.maze ? 4 : p.oct;
    const int D = 1 << (oct > 0 ? oct - 1 : 0);
    if (threadIdx.x < 8) {
        const int k = threadIdx.x;
        f_oct[k] = fk_cos_factor((float)k / (float)D);
        float fr = fmodf((float)k, 8.f) / 8;
        f_frac[k] = fk_cos_factor(fr);
        f_edge[k] = fk_cos_factor(2 * fr - 1);
    }
    __syncthreads();

    const size_t n = (size_t)p.nr_v * p.nr_v;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float d = (float)ldexp(1.0, oct - 1);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
        const int i = (int)(idx / p.nr_v), j = (int)(idx % p.nr_v);
        float amp0 = p.amp, avg = 0.f;
        if (p.maze) {
            const int ki = i & 7, kj = j & 7;                   /* fmodf(i, 8) for i >= 0 */
            const int mi = i >> 3, mj = j >> 3;
            int cn = fk_maze(p, mi, mj);
            int xn = fk_maze(p, ki >= 4 ? mi + 1 : mi - 1, mj);
            int yn = fk_maze(p, mi, kj >= 4 ? mj + 1 : mj - 1);
            float xa = cn > xn ? (float)cn : fk_cos_mix((float)cn, (float)xn, f_edge[ki]);
            float ya = cn > yn ? (float)cn : fk_cos_mix((float)cn, (float)yn, f_edge[kj]);
            avg = fk_cos_mix(xa, ya, f_frac[ki > kj ? ki - kj : kj - ki]);
            amp0 = powf(1.5f, avg);
        }
        float total = 0;
        for (int o = 0; o < oct; o++) {
            float freq = (float)(ldexp(1.0, o) / (double)d);
            float amp = (float)(ldexp(1.0, -o) * (double)amp0);
            float x = (float)i * freq, z = (float)j * freq;
            int ix = (int)floorf(x), iz = (int)floorf(z);       /* x >= 0: floor() of the promoted value is the same */
            float fx = x - (float)ix, fz = z - (float)iz;
            float gx = f_oct[(int)(fx * (float)D)], gz = f_oct[(int)(fz * (float)D)];
            const float *row0 = p.smooth + (size_t)ix * (p.nr_v + 1) + iz;
            const float *row1 = row0 + (p.nr_v + 1);
            float v1 = __ldg(row0), v2 = __ldg(row1), v3 = __ldg(row0 + 1), v4 = __ldg(row1 + 1);
            total += fk_cos_mix(fk_cos_mix(v1, v2, gx), fk_cos_mix(v3, v4, gx), gz) * amp;
        }
        float h = p.ty + total;
        p.map[idx] = p.maze ? h + avg : h;
    }
}

/* ---- terrain mesh: core/terrain.c:93-110 (calc_normal), :479-519 (vertex / normal / uv / index buffers) -------- */

struct TerrainMeshParams {
    const float *map;           /* t->map, nr_v * nr_v, map[a*nr_v + b] */
    unsigned nr_v;
    float x, y, z, side;        /* terrain origin and edge length (terrain.c:441-445) */
    float *vx, *norm, *tx;      /* 3, 3, 2 floats per vertex (any may be NULL) */
    unsigned short *idx;        /* 6 per quad */
};

/*
 * Vertex it = i*nr_v + j reads the map TRANSPOSED -- map[j*nr_v + i], and calc_normal(t, n, j, i) its four
 * neighbours (terrain.c:497-500) -- so a CTA stages a 32 x 32 map tile plus a one-cell halo in shared memory with
 * coalesced row reads and then walks it column-wise: consecutive threads produce consecutive vertices.
 * Arithmetic is the reference's, operation by operation (no FMA contraction in this translation unit):
 * calc_normal() zeroes the neighbour beyond an edge (its torus indices are computed but never used),
 * vec3_norm() is k = (float)(1.0 / (double)sqrtf(n.n)) then three float products (linmath.h:48-62).
 */
/*
 * One warp = 32 consecutive vertices of one output row; it stages each attribute in shared memory and writes it
 * back as whole 16-byte vectors (96 floats = 24 float4 for vx / norm, 16 for tx), so the 12-byte vertex stride
 * never reaches the memory system as strided 4-byte stores.
 */
__device__ __forceinline__ void fk_flush_row(float *dst, const float *st, int nfloats, bool vec, int lane)
{
    __syncwarp();
    if (vec && (reinterpret_cast<size_t>(dst) & 15) == 0) {
        if (lane < nfloats / 4)
            reinterpret_cast<float4 *>(dst)[lane] = reinterpret_cast<const float4 *>(st)[lane];
    } else {
        for (int k = lane; k < nfloats; k += 32)
            dst[k] = st[k];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(256) terrain_mesh_vertex_kernel(TerrainMeshParams p)
{
    __shared__ float tile[34][35];
    __shared__ __align__(16) float stage[8][96];
    const int nr = (int)p.nr_v;
    const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;      /* map row (= vertex j) / map column (= vertex i) origin */
    const int tx = threadIdx.x, ty = threadIdx.y;               /* 32 x 8 threads: a warp per output row, 4 rows each */
    for (int r = ty; r < 34; r += 8)
        for (int c = tx; c < 34; c += 32) {
            const int a = a0 + r - 1, b = b0 + c - 1;
            tile[r][c] = (a >= 0 && a < nr && b >= 0 && b < nr) ? p.map[(size_t)a * nr + b] : 0.f;
        }
    __syncthreads();
    const int j = a0 + tx;                                      /* j fastest: it = i*nr_v + j is contiguous per warp */
    const bool live = j < nr;
    const int count = nr - a0 < 32 ? nr - a0 : 32;              /* vertices of a warp's row */
    const bool vec = count == 32 && (nr & 3) == 0;              /* 16-byte aligned rows of 96 / 64 floats */
    const float den = (float)p.nr_v - 1;
    float *st = stage[ty];
    for (int ci = ty; ci < 32; ci += 8) {
        const int i = b0 + ci;
        if (i >= nr)
            break;                                              /* the whole warp */
        const size_t it0 = (size_t)i * nr + a0;
        if (p.vx) {
            if (live) {
                st[tx * 3 + 0] = p.x + (float)j / den * p.side;
                st[tx * 3 + 1] = p.y + tile[tx + 1][ci + 1];
                st[tx * 3 + 2] = p.z + (float)i / den * p.side;
            }
            fk_flush_row(p.vx + it0 * 3, st, count * 3, vec, tx);
        }
        if (p.norm) {
            if (live) {
                /* calc_normal(t, n, x = j, z = i): hl/hr along the first map index, hd/hu along the second */
                const float hl = tile[tx][ci + 1], hr = tile[tx + 2][ci + 1];
                const float hd = tile[tx + 1][ci], hu = tile[tx + 1][ci + 2];
                const float n0 = hl - hr, n1 = 2.f, n2 = hd - hu;
                float dot = 0.f;
                dot += n0 * n0;
                dot += n1 * n1;
                dot += n2 * n2;
                const float k = (float)(1.0 / (double)sqrtf(dot));
                st[tx * 3 + 0] = n0 * k;
                st[tx * 3 + 1] = n1 * k;
                st[tx * 3 + 2] = n2 * k;
            }
            fk_flush_row(p.norm + it0 * 3, st, count * 3, vec, tx);
        }
        if (p.tx) {
            if (live) {
                st[tx * 2 + 0] = (float)j * 32 / den;
                st[tx * 2 + 1] = (float)i * 32 / den;
            }
            fk_flush_row(p.tx + it0 * 2, st, count * 2, vec, tx);
        }
    }
}

/* two triangles per quad, indices truncated to unsigned short exactly like the reference's buffer (terrain.c:503-516) */
__global__ void __launch_bounds__(256) terrain_mesh_index_kernel(TerrainMeshParams p)
{
    const size_t q = (size_t)p.nr_v - 1;
    const size_t n = q * q;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < n; it += stride) {
        const unsigned i = (unsigned)(it / q), j = (unsigned)(it % q);
        const unsigned tl = i * p.nr_v + j, tr = tl + 1, bl = (i + 1) * p.nr_v + j, br = bl + 1;
        uint32_t *out = reinterpret_cast<uint32_t *>(p.idx + it * 6);       /* 12 bytes per quad: 4-byte aligned */
        out[0] = (tl & 0xffffu) | (bl << 16);
        out[1] = (tr & 0xffffu) | (tr << 16);
        out[2] = (bl & 0xffffu) | (br << 16);
    }
}

/* ---- instantiator extraction: core/terrain.c:555-570 with terrain_height() (:336-379), barrycentric() (interp.h:49-56) ---- */

struct InstorParams {
    const uint8_t *maze;        /* mside x mside xyarray payload after the ca_instors[] passes (terrain.c:473-477) */
    unsigned mside;
    uint32_t kinds[4];          /* ca_instors[k].nr_states: a cell equal to it spawns an instantiator of kind k */
    int nkinds;
    const float *map;           /* t->map */
    unsigned nr_v;
    float x, z, side;           /* arguments of terrain_init_square_landscape() */
    unsigned tside;             /* t->side = (unsigned)side (terrain.h:19, terrain.c:443) */
    unsigned *counts;           /* [mside + 1]: matches per maze column i, then their exclusive scan (last = total) */
    int4 *out;                  /* { kind, dx, dy, dz } as bit patterns */
    unsigned long long cap;
};

/* interp.h:49-56, operation by operation */
__device__ __forceinline__ float fk_barrycentric(const float p1[3], const float p2[3], const float p3[3], float px, float pz)
{
    float det = (p2[2] - p3[2]) * (p1[0] - p3[0]) + (p3[0] - p2[0]) * (p1[2] - p3[2]);
    float l1 = ((p2[2] - p3[2]) * (px - p3[0]) + (p3[0] - p2[0]) * (pz - p3[2])) / det;
    float l2 = ((p3[2] - p1[2]) * (px - p3[0]) + (p1[0] - p3[0]) * (pz - p3[2])) / det;
    float l3 = 1.0f - l1 - l2;
    return l1 * p1[1] + l2 * p2[1] + l3 * p3[1];
}

/* terrain_height(): core/terrain.c:336-379 (t->y plays no part; t->side is the truncated unsigned) */
__device__ __forceinline__ float fk_terrain_height(const InstorParams &p, float x, float z)
{
    const int nr = (int)p.nr_v;
    float square = (float)p.tside / (float)(p.nr_v - 1u);
    float tx = x - p.x, tz = z - p.z;
    int gridx = (int)floorf(tx / square), gridz = (int)floorf(tz / square);
    float xoff = (tx - square * (float)gridx) / square;
    float zoff = (tz - square * (float)gridz) / square;
    if (x < p.x || x > p.x + (float)p.tside || z < p.z || z > p.z + (float)p.tside)
        return 0.f;
    const float h00 = p.map[(size_t)gridx * nr + gridz], h10 = p.map[(size_t)(gridx + 1) * nr + gridz];
    const float h01 = p.map[(size_t)gridx * nr + gridz + 1];
    if (xoff <= 1 - zoff) {
        const float p1[3] = { 0, h00, 0 }, p2[3] = { 1, h10, 0 }, p3[3] = { 0, h01, 1 };
        return fk_barrycentric(p1, p2, p3, xoff, zoff);
    }
    const float h11 = p.map[(size_t)(gridx + 1) * nr + gridz + 1];
    const float p1[3] = { 1, h10, 0 }, p2[3] = { 1, h11, 1 }, p3[3] = { 0, h01, 1 };
    return fk_barrycentric(p1, p2, p3, xoff, zoff);
}

__device__ __forceinline__ unsigned fk_instor_matches(const InstorParams &p, unsigned cell)
{
    unsigned m = 0;
    for (int k = 0; k < p.nkinds; k++)
        m |= (unsigned)(cell == p.kinds[k]) << k;
    return m;
}

/* pass 1: one warp per maze column i (the reference's OUTER loop index): number of instantiators it spawns */
__global__ void __launch_bounds__(256) instor_count_kernel(InstorParams p)
{
    const unsigned i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= p.mside)
        return;
    unsigned n = 0;
    for (unsigned j = lane; j < p.mside; j += 32)
        n += __popc(fk_instor_matches(p, p.maze[(size_t)j * p.mside + i]));     /* xyarray_get(maze, i, j) */
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0)
        p.counts[i] = n;
}

/* pass 2: exclusive scan of the per-column counts in place, total in counts[mside] (one CTA: mside <= 5792) */
__global__ void __launch_bounds__(1024) instor_scan_kernel(unsigned *counts, unsigned n)
{
    __shared__ unsigned part[1024];
    const unsigned per = (n + 1023) / 1024, t = threadIdx.x;
    const unsigned lo = t * per < n ? t * per : n, hi = lo + per < n ? lo + per : n;
    unsigned sum = 0;
    for (unsigned k = lo; k < hi; k++) sum += counts[k];
    part[t] = sum;
    __syncthreads();
    for (unsigned d = 1; d < 1024; d <<= 1) {
        unsigned v = t >= d ? part[t - d] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned run = part[t] - sum;                       /* exclusive prefix of this thread's slice */
    for (unsigned k = lo; k < hi; k++) {
        unsigned c = counts[k];
        counts[k] = run;
        run += c;
    }
    if (t == 1023)
        counts[n] = part[1023];
}

/* pass 3: the records, in the reference's order: i outer, j inner, kind innermost */
__global__ void __launch_bounds__(256) instor_emit_kernel(InstorParams p)
{
    const unsigned i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= p.mside)
        return;
    unsigned long long base = p.counts[i];
    if (p.counts[i + 1] == p.counts[i])
        return;
    const float den = (float)(p.nr_v - 1u);
    for (unsigned j0 = 0; j0 < p.mside; j0 += 32) {
        const unsigned j = j0 + lane;
        const unsigned m = j < p.mside ? fk_instor_matches(p, p.maze[(size_t)j * p.mside + i]) : 0u;
        unsigned n = __popc(m), incl = n;
        for (int d = 1; d < 32; d <<= 1) {
            unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += v;
        }
        unsigned long long at = base + incl - n;
        if (m) {
            /* terrain.c:563-565 */
            const float dx = p.x + (float)((double)i + 0.5) * 8 * p.side / den;
            const float dz = p.z + (float)((double)j + 0.5) * 8 * p.side / den;
            const float dy = fk_terrain_height(p, dx, dz);
            for (int k = 0; k < p.nkinds; k++)
                if ((m >> k) & 1u) {
                    if (at < p.cap)
                        p.out[at] = make_int4(k, __float_as_int(dx), __float_as_int(dy), __float_as_int(dz));
                    at++;
                }
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

/* ---- blue_noise2d_tex(): core/noise.c:17-169 (the 64 x 64 film-grain texture, up to the upload) --------------- */

enum { FK_GRAIN = 64 };     /* FILM_GRAIN_SIZE, core/shader_constants.h:13: the reference's spectrum arrays have this size */

struct BlueNoiseParams {
    float *out;                     /* FK_GRAIN^2 RGBA32F pixels */
    unsigned long long state;       /* the caller's drand48() stream: 48-bit state before the first draw */
};

/* glibc rand48 step (the same generator as ca2d_layout.cuh; repeated here: this TU is built with -fmad=false) */
__device__ __forceinline__ unsigned long long fk_r48_step(unsigned long long x)
{
    return (x * 0x5DEECE66Dull + 0xBull) & ((1ull << 48) - 1ull);
}
__device__ inline unsigned long long fk_r48_advance(unsigned long long x, unsigned long long n)
{
    const unsigned long long M = (1ull << 48) - 1ull;
    unsigned long long h = 0x5DEECE66Dull, f = 0xBull, G = 1ull, C = 0ull;
    while (n) {
        if (n & 1ull) { G = (G * h) & M; C = (C * h + f) & M; }
        f = (f * (h + 1ull)) & M;
        h = (h * h) & M;
        n >>= 1;
    }
    return (G * x + C) & M;
}

/*
 * 64 independent 64-point FFTs over the lines of X (line l, element e at X[l * ls + e * es]): bit reversal, then six
 * radix-2 stages of 64 x 32 butterflies spread over the CTA.  sign = -1: forward (kiss_fft inverse = 0), +1: the
 * unnormalised inverse.  tw[m] = exp(-2 pi i m / 64), m < 32.
 */
__device__ inline void fk_fft64_lines(float2 *X, int ls, int es, const float2 *tw, float sign)
{
    for (int idx = threadIdx.x; idx < FK_GRAIN * FK_GRAIN; idx += blockDim.x) {
        const int l = idx >> 6, e = idx & 63;
        const int r = (int)(__brev((unsigned)e) >> 26);
        if (e < r) {
            float2 a = X[l * ls + e * es], b = X[l * ls + r * es];
            X[l * ls + e * es] = b;
            X[l * ls + r * es] = a;
        }
    }
    __syncthreads();
    for (int s = 0; s < 6; s++) {
        const int half = 1 << s;
        for (int idx = threadIdx.x; idx < FK_GRAIN * 32; idx += blockDim.x) {
            const int l = idx >> 5, k = idx & 31;
            const int j = k & (half - 1);
            const int i0 = ((k >> s) << (s + 1)) + j, i1 = i0 + half;
            float2 w = tw[j << (5 - s)];
            w.y *= -sign;                               /* tw holds the forward (negative-angle) twiddles */
            const float2 a = X[l * ls + i0 * es], b = X[l * ls + i1 * es];
            const float2 t = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
            X[l * ls + i0 * es] = make_float2(a.x + t.x, a.y + t.y);
            X[l * ls + i1 * es] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncthreads();
    }
}

/*
 * One CTA: the three colour channels one after the other through a 64 x 64 complex tile in shared memory (white noise
 * from the caller's drand48 stream, rows + columns forward, radial gain r / r_max, rows + columns back, / 64^2), then
 * the joint min / max of the three channels and the normalisation to [0, 1]; alpha = 1.
 */
__global__ void __launch_bounds__(256) blue_noise2d_kernel(BlueNoiseParams p)
{
    __shared__ float2 X[FK_GRAIN * FK_GRAIN];
    __shared__ float2 tw[32];
    __shared__ float red[2][8];
    const int t = threadIdx.x, nt = blockDim.x;
    if (t < 32) {
        float sn, cs;
        sincospif((float)t / 32.0f, &sn, &cs);
        tw[t] = make_float2(cs, -sn);
    }
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int c = 0; c < 3; c++) {
        __syncthreads();
        /* noise.c:106-115: pixel q = x + y * size takes draws 3 q, 3 q + 1, 3 q + 2 of the stream */
        const int per = (FK_GRAIN * FK_GRAIN + nt - 1) / nt;
        const int q0 = t * per, q1 = min(FK_GRAIN * FK_GRAIN, q0 + per);
        if (q0 < q1) {
            unsigned long long st = fk_r48_advance(p.state, 3ull * (unsigned long long)q0 + (unsigned long long)c);
            for (int q = q0; q < q1; q++) {
                st = fk_r48_step(st);
                const double d = (double)st * (1.0 / 281474976710656.0);        /* drand48(): X / 2^48 */
                const double wgt = c == 0 ? 0.299 : (c == 1 ? 0.587 : 0.114);
                X[q] = make_float2((float)(((d * 4.0 - 1.0) / 3.0) * wgt), 0.0f);
                st = fk_r48_step(fk_r48_step(st));
            }
        }
        __syncthreads();
        fk_fft64_lines(X, FK_GRAIN, 1, tw, -1.0f);      /* rows */
        fk_fft64_lines(X, 1, FK_GRAIN, tw, -1.0f);      /* columns */
        /* blue_noise2d_gain(): noise.c:75-92 */
        const float maxr = sqrtf((float)((FK_GRAIN / 2) * (FK_GRAIN / 2) + (FK_GRAIN / 2) * (FK_GRAIN / 2)));
        for (int idx = t; idx < FK_GRAIN * FK_GRAIN; idx += nt) {
            const int y = idx >> 6, x = idx & 63;
            const int fy = (y <= FK_GRAIN / 2) ? y : y - FK_GRAIN, fx = (x <= FK_GRAIN / 2) ? x : x - FK_GRAIN;
            const float r = (float)sqrt((double)(fx * fx + fy * fy));
            const float gain = r / maxr;
            X[idx].x *= gain;
            X[idx].y *= gain;
        }
        __syncthreads();
        fk_fft64_lines(X, FK_GRAIN, 1, tw, 1.0f);
        fk_fft64_lines(X, 1, FK_GRAIN, tw, 1.0f);
        for (int idx = t; idx < FK_GRAIN * FK_GRAIN; idx += nt) {
            const float v = X[idx].x / (float)(FK_GRAIN * FK_GRAIN);
            p.out[idx * 4 + c] = v;
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
    }
    /* noise.c:141-152: joint range of the three channels */
    for (int o = 16; o; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((t & 31) == 0) { red[0][t >> 5] = mn; red[1][t >> 5] = mx; }
    __syncthreads();
    mn = red[0][0]; mx = red[1][0];
    for (int w = 1; w < (nt >> 5); w++) { mn = fminf(mn, red[0][w]); mx = fmaxf(mx, red[1][w]); }
    for (int idx = t; idx < FK_GRAIN * FK_GRAIN; idx += nt) {
        for (int c = 0; c < 3; c++)
            p.out[idx * 4 + c] = (p.out[idx * 4 + c] - mn) / (mx - mn);
        p.out[idx * 4 + 3] = 1.0f;
    }
}

} // namespace clapca
#endif
