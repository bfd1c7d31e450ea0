/*
 * api_fields.cu -- field part of the C ABI (include/clapca.h): the noise.c gradient bake, the terrain.c
 * heightmap chain, the terrain mesh buffers and the instantiator extraction.  Compiled with -fmad=false:
 * the kernels mirror the reference's unfused float / double arithmetic operation by operation.
 */
#include "api_internal.h"
#include "field_kernels.cuh"

using namespace clapca;
using namespace clapca::api;

extern "C" {
#pragma GCC visibility push(default)

/*
 * Can the bake take the lattice path (field_kernels.cuh, noise_field_kernel)?  Yes when the +- eps samples of every
 * voxel are bit for bit the coordinates of its neighbours' centres, in the float arithmetic the reference uses
 * (core/noise.c:236-254: step = period_units / size, p = (float)x * step, samples p + eps and p - eps with eps = step),
 * and the lattice fits a bounded scratch buffer.  CLAPCA_NOISE_DIRECT=1 forces the six-samples-per-voxel kernel.
 */
static bool noise_lattice_ok(size_t size, float period_units)
{
    if (const char *e = getenv("CLAPCA_NOISE_DIRECT"))
        if (atoi(e)) return false;
    if ((size + 2) * (size + 2) * (size + 2) * sizeof(float) > ((size_t)2 << 30))
        return false;
    volatile float step = period_units / (float)size;
    for (size_t x = 0; x < size; x++) {
        volatile float p = (float)x * step;
        volatile float plus = p + step, minus = p - step;
        volatile float next = (float)(x + 1) * step, prev = (float)((double)x - 1.0) * step;
        if (plus != next || minus != prev)
            return false;
    }
    return true;
}

/* fills p.field (lattice path) when the precondition holds; the caller launches its voxel kernel afterwards */
static int noise_prepare_lattice(NoiseBakeParams &p)
{
    p.field = nullptr;
    if (!noise_lattice_ok(p.size, p.period_units))
        return CLAPCA_OK;
    const size_t S2 = (size_t)p.size + 2, points = S2 * S2 * S2;
    if (int rc = ensure_bytes(&g_ctx.scratch[0], &g_ctx.scratch_bytes[0], points * sizeof(float))) return rc;
    float *field = (float *)g_ctx.scratch[0];
    noise_field_kernel<<<grid_blocks_for(points, 256, 8), 256, 0, g_ctx.stream>>>(p, field);
    CU(cudaGetLastError());
    p.field = field;
    return CLAPCA_OK;
}

int clapca_noise_bake_device(void *d_out, size_t size, int octaves, float lacunarity, float gain,
                             float period_units, uint32_t seed, float *kernel_ms)
{
    if (int rc = need_init()) return rc;
    if (!d_out || size < 1 || size > 4096 || octaves < 0 || (int)period_units < 1)
        return fail(CLAPCA_ERR_ARG, "noise bake: bad arguments (size %zu, octaves %d, period %g)", size, octaves,
                    (double)period_units);
    NoiseBakeParams p = { (uint32_t *)d_out, (unsigned)size, octaves, lacunarity, gain, period_units, seed, 0, nullptr };
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    size_t voxels = size * size * size;
    CU(cudaEventRecord(a, g_ctx.stream));
    if (int rc = noise_prepare_lattice(p)) return rc;
    noise_bake_kernel<<<grid_blocks_for(voxels, 256, 8), 256, 0, g_ctx.stream>>>(p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(b, g_ctx.stream));
    int rc = timed_sync(a, b, kernel_ms);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return rc;
}

int clapca_noise_grad3d_bake_rgba8(uint8_t *out, size_t size, int octaves, float lacunarity, float gain,
                                   float period_units, uint32_t seed)
{
    if (int rc = need_init()) return rc;
    if (!out) return fail(CLAPCA_ERR_ARG, "noise bake: NULL output");
    size_t bytes = size * size * size * 4;
    /* grow-only staging (a cudaMalloc / cudaFree pair per call costs more than the bake: 2.8 .. 17 ms measured) */
    if (int rc = ensure_bytes(&g_ctx.scratch[1], &g_ctx.scratch_bytes[1], bytes ? bytes : 4)) return rc;
    void *d = g_ctx.scratch[1];
    int rc = clapca_noise_bake_device(d, size, octaves, lacunarity, gain, period_units, seed, nullptr);
    if (!rc) rc = clapca_memcpy_d2h(out, d, bytes);
    return rc;
}

/* ---- blue_noise2d_tex(): core/noise.c:96-169, the pixels it uploads ------------------------------------------ */

int clapca_noise_blue2d_device(void *d_out, int size, uint64_t rand48_state, uint64_t *state_out, float *kernel_ms)
{
    if (int rc = need_init()) return rc;
    if (!d_out) return fail(CLAPCA_ERR_ARG, "blue noise: NULL output");
    if (size != FK_GRAIN)
        return fail(CLAPCA_ERR_UNSUPPORTED, "blue noise: size must be FILM_GRAIN_SIZE = %d (the reference's spectrum "
                    "arrays have that size whatever it is called with, core/noise.c:17-73)", (int)FK_GRAIN);
    BlueNoiseParams p = { (float *)d_out, rand48_state & ((1ull << 48) - 1ull) };
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    CU(cudaEventRecord(a, g_ctx.stream));
    blue_noise2d_kernel<<<1, 256, 0, g_ctx.stream>>>(p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(b, g_ctx.stream));
    int rc = timed_sync(a, b, kernel_ms);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    /* the stream after the 3 * size^2 draws of noise.c:106-115 */
    if (!rc && state_out) {
        const unsigned long long M = (1ull << 48) - 1ull;
        unsigned long long h = 0x5DEECE66Dull, f = 0xBull, G = 1ull, C = 0ull, n = 3ull * size * size;
        while (n) {
            if (n & 1ull) { G = (G * h) & M; C = (C * h + f) & M; }
            f = (f * (h + 1ull)) & M;
            h = (h * h) & M;
            n >>= 1;
        }
        *state_out = (G * p.state + C) & M;
    }
    return rc;
}

int clapca_noise_blue2d_rgba32f(float *out, int size, uint64_t rand48_state, uint64_t *state_out)
{
    if (int rc = need_init()) return rc;
    if (!out) return fail(CLAPCA_ERR_ARG, "blue noise: NULL output");
    const size_t bytes = (size_t)FK_GRAIN * FK_GRAIN * 4 * sizeof(float);
    void *d = nullptr;
    CU(cudaMalloc(&d, bytes));
    int rc = clapca_noise_blue2d_device(d, size, rand48_state, state_out, nullptr);
    if (!rc) rc = clapca_memcpy_d2h(out, d, bytes);
    cudaFree(d);
    return rc;
}

/* ---- the bake as a device-resident 3D texture object ---------------------------------------- */

struct clapca_tex3d {
    cudaArray_t array = nullptr;
    cudaSurfaceObject_t surf = 0;
    size_t size = 0;
};

int clapca_tex3d_destroy(clapca_tex3d *t)
{
    if (!t) return CLAPCA_OK;
    if (t->surf) cudaDestroySurfaceObject(t->surf);
    if (t->array) cudaFreeArray(t->array);
    delete t;
    return CLAPCA_OK;
}

int clapca_noise_bake_array(clapca_tex3d **out, size_t size, int octaves, float lacunarity, float gain,
                            float period_units, uint32_t seed, float *kernel_ms)
{
    if (int rc = need_init()) return rc;
    if (!out || size < 1 || size > 2048 || octaves < 0 || (int)period_units < 1)
        return fail(CLAPCA_ERR_ARG, "noise bake (array): bad arguments (size %zu, octaves %d, period %g)", size, octaves,
                    (double)period_units);
    clapca_tex3d *t = new (std::nothrow) clapca_tex3d();
    if (!t) return fail(CLAPCA_ERR_NOMEM, "noise bake (array): host allocation failed");
    t->size = size;
    const cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();     /* TEX_FMT_RGBA8 */
    cudaError_t e = cudaMalloc3DArray(&t->array, &desc, make_cudaExtent(size, size, size), cudaArraySurfaceLoadStore);
    if (e == cudaSuccess) {
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = t->array;
        e = cudaCreateSurfaceObject(&t->surf, &rd);
    }
    if (e != cudaSuccess) {
        clapca_tex3d_destroy(t);
        return fail(e == cudaErrorMemoryAllocation ? CLAPCA_ERR_NOMEM : CLAPCA_ERR_CUDA, "noise bake (array): %s",
                    cudaGetErrorString(e));
    }
    NoiseBakeParams p = { nullptr, (unsigned)size, octaves, lacunarity, gain, period_units, seed, t->surf, nullptr };
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    CU(cudaEventRecord(a, g_ctx.stream));
    if (int rc = noise_prepare_lattice(p)) { clapca_tex3d_destroy(t); return rc; }
    noise_bake_surface_kernel<<<grid_blocks_for(size * size * size, 256, 8), 256, 0, g_ctx.stream>>>(p);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventRecord(b, g_ctx.stream);
    int rc = e == cudaSuccess ? timed_sync(a, b, kernel_ms) : fail(CLAPCA_ERR_CUDA, "noise bake (array): %s", cudaGetErrorString(e));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    if (rc) {
        clapca_tex3d_destroy(t);
        return rc;
    }
    *out = t;
    return CLAPCA_OK;
}

void *clapca_tex3d_array(clapca_tex3d *t) { return t ? (void *)t->array : nullptr; }

int clapca_tex3d_download(clapca_tex3d *t, uint8_t *host_rgba8)
{
    if (int rc = need_init()) return rc;
    if (!t || !host_rgba8) return fail(CLAPCA_ERR_ARG, "tex3d_download: NULL argument");
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof(cp));
    cp.srcArray = t->array;
    cp.dstPtr = make_cudaPitchedPtr(host_rgba8, t->size * 4, t->size, t->size);
    cp.extent = make_cudaExtent(t->size, t->size, t->size);
    cp.kind = cudaMemcpyDeviceToHost;
    CU(cudaMemcpy3DAsync(&cp, g_ctx.stream));
    CU(cudaStreamSynchronize(g_ctx.stream));
    return CLAPCA_OK;
}

int clapca_noise_fbm3(float *out, const float *xyz, size_t n, int octaves, float lacunarity, float gain,
                      int period, uint32_t seed)
{
    if (int rc = need_init()) return rc;
    if (!out || !xyz || period < 1) return fail(CLAPCA_ERR_ARG, "noise_fbm3: bad arguments");
    if (!n) return CLAPCA_OK;
    float *d_in = nullptr, *d_out = nullptr;
    CU(cudaMalloc(&d_in, n * 3 * sizeof(float)));
    cudaError_t e = cudaMalloc(&d_out, n * sizeof(float));
    if (e != cudaSuccess) { cudaFree(d_in); return fail(CLAPCA_ERR_NOMEM, "noise_fbm3: %s", cudaGetErrorString(e)); }
    int rc = clapca_memcpy_h2d(d_in, xyz, n * 3 * sizeof(float));
    if (!rc) {
        noise_fbm3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, g_ctx.stream>>>(d_out, d_in, n, octaves, lacunarity,
                                                                                 gain, period, seed);
        if (cudaGetLastError() != cudaSuccess) rc = fail(CLAPCA_ERR_CUDA, "noise_fbm3 launch failed");
    }
    if (!rc) rc = clapca_memcpy_d2h(out, d_out, n * sizeof(float));
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}

int clapca_terrain_heightmap_device(void *d_map, void *d_map0, long seed, unsigned nr_v, float ty,
                                    const void *d_maze, unsigned mside, float amp, int oct,
                                    float *map0_ms, float *map_ms)
{
    if (int rc = need_init()) return rc;
    if (!d_map0 || nr_v < 1 || nr_v > 46340)
        return fail(CLAPCA_ERR_ARG, "terrain: bad arguments (nr_v %u)", nr_v);
    if (d_maze && mside < 1) return fail(CLAPCA_ERR_ARG, "terrain: maze without a side length");
    cudaEvent_t e0, e1, e2;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventCreate(&e2));
    size_t n = (size_t)nr_v * nr_v;
    CU(cudaEventRecord(e0, g_ctx.stream));
    terrain_map0_kernel<<<grid_blocks_for(n, 256, 8), 256, 0, g_ctx.stream>>>((float *)d_map0, (long long)seed, nr_v);
    CU(cudaGetLastError());
    CU(cudaEventRecord(e1, g_ctx.stream));
    if (d_map) {
        TerrainParams p = { (float *)d_map, (const float *)d_map0, (const uint8_t *)d_maze, nr_v, mside, ty, amp, oct,
                            nullptr };
        const char *direct = getenv("CLAPCA_TERRAIN_DIRECT");       /* diagnostics: evaluate the 3x3 kernel per use */
        if (!(direct && atoi(direct) > 0)) {
            const size_t sn = ((size_t)nr_v + 1) * ((size_t)nr_v + 1);
            if (int rc = ensure_bytes(&g_ctx.d_smooth, &g_ctx.smooth_bytes, sn * sizeof(float))) return rc;
            terrain_smooth_kernel<<<grid_blocks_for(sn, 256, 8), 256, 0, g_ctx.stream>>>(p, (float *)g_ctx.d_smooth);
            CU(cudaGetLastError());
            p.smooth = (const float *)g_ctx.d_smooth;
        }
        /* tabulated blend factors need 2^(oct-1) <= 8 distinct fractions (the reference fixes OCTAVES = 4) */
        if (p.smooth && (d_maze || (oct >= 0 && oct <= 4)))
            terrain_heightmap_tab_kernel<<<grid_blocks_for(n, 256, 8), 256, 0, g_ctx.stream>>>(p);
        else
            terrain_heightmap_kernel<<<grid_blocks_for(n, 256, 8), 256, 0, g_ctx.stream>>>(p);
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(e2, g_ctx.stream));
    int rc = timed_sync(e0, e1, map0_ms);
    if (!rc && map_ms) {
        cudaError_t e = cudaEventElapsedTime(map_ms, e1, e2);
        if (e != cudaSuccess) rc = fail(CLAPCA_ERR_CUDA, "terrain: %s", cudaGetErrorString(e));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    return rc;
}

int clapca_terrain_map0(float *map0, long seed, unsigned nr_v)
{
    if (int rc = need_init()) return rc;
    if (!map0) return fail(CLAPCA_ERR_ARG, "terrain_map0: NULL output");
    size_t bytes = (size_t)nr_v * nr_v * sizeof(float);
    void *d = nullptr;
    CU(cudaMalloc(&d, bytes ? bytes : 4));
    int rc = clapca_terrain_heightmap_device(nullptr, d, seed, nr_v, 0.f, nullptr, 0, 0.f, 0, nullptr, nullptr);
    if (!rc) rc = clapca_memcpy_d2h(map0, d, bytes);
    cudaFree(d);
    return rc;
}

int clapca_terrain_heightmap(float *map, long seed, unsigned nr_v, float ty, const uint8_t *maze, unsigned mside,
                             float amp, int oct)
{
    if (int rc = need_init()) return rc;
    if (!map) return fail(CLAPCA_ERR_ARG, "terrain_heightmap: NULL output");
    size_t bytes = (size_t)nr_v * nr_v * sizeof(float);
    void *d_map = nullptr, *d_map0 = nullptr, *d_maze = nullptr;
    if (int rc = ensure_bytes(&g_ctx.scratch[0], &g_ctx.scratch_bytes[0], bytes ? bytes : 4)) return rc;
    if (int rc = ensure_bytes(&g_ctx.scratch[1], &g_ctx.scratch_bytes[1], bytes ? bytes : 4)) return rc;
    d_map = g_ctx.scratch[0];
    d_map0 = g_ctx.scratch[1];
    int rc = CLAPCA_OK;
    if (maze) {
        if (int rc2 = ensure_bytes(&g_ctx.scratch[2], &g_ctx.scratch_bytes[2], (size_t)mside * mside)) return rc2;
        d_maze = g_ctx.scratch[2];
        rc = clapca_memcpy_h2d(d_maze, maze, (size_t)mside * mside);
    }
    if (!rc) rc = clapca_terrain_heightmap_device(d_map, d_map0, seed, nr_v, ty, d_maze, mside, amp, oct, nullptr, nullptr);
    if (!rc) rc = clapca_memcpy_d2h(map, d_map, bytes);
    return rc;
}

/* ---- terrain mesh ------------------------------------------------------------ */

int clapca_terrain_mesh_device(const void *d_map, unsigned nr_v, float x, float y, float z, float side,
                               void *d_vx, void *d_norm, void *d_tx, void *d_idx, float *kernel_ms)
{
    if (int rc = need_init()) return rc;
    if (!d_map || nr_v < 1 || nr_v > 46340)
        return fail(CLAPCA_ERR_ARG, "terrain_mesh: bad arguments (nr_v %u)", nr_v);
    TerrainMeshParams p = { (const float *)d_map, nr_v, x, y, z, side, (float *)d_vx, (float *)d_norm, (float *)d_tx,
                            (unsigned short *)d_idx };
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, g_ctx.stream));
    if (d_vx || d_norm || d_tx) {
        const unsigned tiles = (nr_v + 31) / 32;
        terrain_mesh_vertex_kernel<<<dim3(tiles, tiles), dim3(32, 8), 0, g_ctx.stream>>>(p);
        CU(cudaGetLastError());
    }
    if (d_idx && nr_v > 1) {
        const size_t quads = (size_t)(nr_v - 1) * (nr_v - 1);
        terrain_mesh_index_kernel<<<grid_blocks_for(quads, 256, 8), 256, 0, g_ctx.stream>>>(p);
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(e1, g_ctx.stream));
    int rc = timed_sync(e0, e1, kernel_ms);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

int clapca_terrain_mesh(const float *map, unsigned nr_v, float x, float y, float z, float side,
                        float *vx, float *norm, float *tx, unsigned short *idx)
{
    if (int rc = need_init()) return rc;
    if (!map || nr_v < 1 || nr_v > 46340)
        return fail(CLAPCA_ERR_ARG, "terrain_mesh: bad arguments (nr_v %u)", nr_v);
    const size_t nv = (size_t)nr_v * nr_v, nq = (size_t)(nr_v - 1) * (nr_v - 1);
    const size_t bytes[5] = { nv * 4, vx ? nv * 12 : 0, norm ? nv * 12 : 0, tx ? nv * 8 : 0, idx ? nq * 12 : 0 };
    void *d[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
    for (int i = 0; i < 5; i++)
        if (bytes[i]) {
            if (int rc = ensure_bytes(&g_ctx.scratch[i], &g_ctx.scratch_bytes[i], bytes[i])) return rc;
            d[i] = g_ctx.scratch[i];
        }
    int rc = clapca_memcpy_h2d(d[0], map, bytes[0]);
    if (!rc) rc = clapca_terrain_mesh_device(d[0], nr_v, x, y, z, side, d[1], d[2], d[3], d[4], nullptr);
    void *host[5] = { nullptr, vx, norm, tx, idx };
    for (int i = 1; i < 5 && !rc; i++)
        if (bytes[i]) rc = clapca_memcpy_d2h(host[i], d[i], bytes[i]);
    return rc;
}

/* ---- instantiator extraction ----------------------------------------------------- */

int clapca_terrain_instantiators_device(const void *d_maze, unsigned mside, const uint32_t *nr_states, int nkinds,
                                        const void *d_map, unsigned nr_v, float x, float z, float side,
                                        void *d_out, size_t cap, size_t *count, float *kernel_ms)
{
    if (int rc = need_init()) return rc;
    if (!d_maze || !d_map || !nr_states || !count || mside < 1 || mside > 8192 || nkinds < 1 || nkinds > 4 || nr_v < 2 ||
        nr_v > 46340 || (cap && !d_out))
        return fail(CLAPCA_ERR_ARG, "terrain_instantiators: bad arguments (mside %u, kinds %d, nr_v %u)", mside, nkinds,
                    nr_v);
    if ((unsigned long long)mside * 8ull > nr_v)
        return fail(CLAPCA_ERR_ARG, "terrain_instantiators: maze of %u cells of 8 vertices does not fit %u vertices",
                    mside, nr_v);
    InstorParams p;
    memset(&p, 0, sizeof(p));
    p.maze = (const uint8_t *)d_maze;
    p.mside = mside;
    for (int k = 0; k < nkinds; k++) p.kinds[k] = nr_states[k];
    p.nkinds = nkinds;
    p.map = (const float *)d_map;
    p.nr_v = nr_v;
    p.x = x; p.z = z; p.side = side;
    p.tside = (unsigned)side;
    if (int rc = ensure_bytes(&g_ctx.scratch[4], &g_ctx.scratch_bytes[4], ((size_t)mside + 1) * sizeof(unsigned))) return rc;
    p.counts = (unsigned *)g_ctx.scratch[4];
    p.out = (int4 *)d_out;
    p.cap = cap;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, g_ctx.stream));
    const unsigned blocks = (mside + 7) / 8;
    instor_count_kernel<<<blocks, 256, 0, g_ctx.stream>>>(p);
    instor_scan_kernel<<<1, 1024, 0, g_ctx.stream>>>(p.counts, mside);
    instor_emit_kernel<<<blocks, 256, 0, g_ctx.stream>>>(p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(e1, g_ctx.stream));
    unsigned total = 0;
    CU(cudaMemcpyAsync(&total, p.counts + mside, sizeof(total), cudaMemcpyDeviceToHost, g_ctx.stream));
    int rc = timed_sync(e0, e1, kernel_ms);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *count = total;
    return rc;
}

int clapca_terrain_instantiators(const uint8_t *maze, unsigned mside, const uint32_t *nr_states, int nkinds,
                                 const float *map, unsigned nr_v, float x, float z, float side,
                                 clapca_instor *out, size_t cap, size_t *count)
{
    if (int rc = need_init()) return rc;
    if (!maze || !map || !count || (cap && !out))
        return fail(CLAPCA_ERR_ARG, "terrain_instantiators: NULL argument");
    const size_t mbytes = (size_t)mside * mside, hbytes = (size_t)nr_v * nr_v * sizeof(float);
    if (int rc = ensure_bytes(&g_ctx.scratch[0], &g_ctx.scratch_bytes[0], hbytes ? hbytes : 4)) return rc;
    if (int rc = ensure_bytes(&g_ctx.scratch[2], &g_ctx.scratch_bytes[2], mbytes ? mbytes : 4)) return rc;
    if (int rc = ensure_bytes(&g_ctx.scratch[3], &g_ctx.scratch_bytes[3], cap ? cap * sizeof(clapca_instor) : 16)) return rc;
    int rc = clapca_memcpy_h2d(g_ctx.scratch[0], map, hbytes);
    if (!rc) rc = clapca_memcpy_h2d(g_ctx.scratch[2], maze, mbytes);
    if (!rc) rc = clapca_terrain_instantiators_device(g_ctx.scratch[2], mside, nr_states, nkinds, g_ctx.scratch[0], nr_v, x, z,
                                                      side, g_ctx.scratch[3], cap, count, nullptr);
    if (!rc && cap && *count)
        rc = clapca_memcpy_d2h(out, g_ctx.scratch[3], std::min(cap, *count) * sizeof(clapca_instor));
    return rc;
}

#pragma GCC visibility pop
} /* extern "C" */
