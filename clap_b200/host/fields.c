/*
 * fields.c -- host wrappers for the two field evaluators
 * (include/clap/noise_bake.h, include/clap/terrain_field.h): allocate the
 * caller-owned result buffer, run the GPU kernels, hand the buffer back.
 */
#include <stdlib.h>
#include <stddef.h>
#include "noise_bake.h"
#include "terrain_field.h"
#include "xyarray.h"
#include "clapca.h"
#include "shim_common.h"

unsigned char *clap_noise_grad3d_bake_rgba8(size_t size, int octaves, float lacunarity, float gain,
                                            float period_units, uint32_t seed)
{
    size_t bytes = size * size * size * 4;
    unsigned char *out = malloc(bytes > 0 ? bytes : 1);
    int rc;

    if (!out)
        return NULL;                    /* the reference reports CERR_NOMEM here (noise.c:231-232) */
    shim_require_gpu();
    rc = clapca_noise_grad3d_bake_rgba8(out, size, octaves, lacunarity, gain, period_units, seed);
    if (rc != CLAPCA_OK)
        shim_fatal("clapca_noise_grad3d_bake_rgba8", rc);
    return out;
}

float *clap_blue_noise2d_rgba32f(int size)
{
    float *out = malloc((size_t)(size > 0 ? size : 1) * (size_t)(size > 0 ? size : 1) * 4 * sizeof(float));
    uint64_t after = 0;
    int rc;

    if (!out)
        return NULL;                    /* the reference reports CERR_NOMEM here (noise.c:101-103) */
    shim_require_gpu();
    /* the three drand48() draws per pixel come out of the process-wide stream, exactly like noise.c:106-115 */
    rc = clapca_noise_blue2d_rgba32f(out, size, shim_rand48_peek(), &after);
    if (rc != CLAPCA_OK)
        shim_fatal("clapca_noise_blue2d_rgba32f", rc);
    shim_rand48_poke(after);
    return out;
}

float *clap_terrain_map0(long seed, unsigned int nr_v)
{
    float *map0 = shim_alloc_zeroed((size_t)nr_v * nr_v * sizeof(float));
    int rc;

    shim_require_gpu();
    rc = clapca_terrain_map0(map0, seed, nr_v);
    if (rc != CLAPCA_OK)
        shim_fatal("clapca_terrain_map0", rc);
    return map0;
}

float *clap_terrain_heightmap(long seed, unsigned int nr_v, float y, unsigned char *maze)
{
    float *map = shim_alloc_zeroed((size_t)nr_v * nr_v * sizeof(float));
    unsigned mside = 0;
    int rc;

    if (maze) {
        struct xyzarray *g = (struct xyzarray *)((char *)maze - offsetof(struct xyzarray, arr));
        mside = (unsigned)g->dim[0];
    }
    shim_require_gpu();
    rc = clapca_terrain_heightmap(map, seed, nr_v, y, maze, mside, 1.0f, 4);
    if (rc != CLAPCA_OK)
        shim_fatal("clapca_terrain_heightmap", rc);
    return map;
}

void clap_terrain_mesh(const float *map, unsigned int nr_v, float x, float y, float z, float side,
                       float **vx, float **norm, float **tx, unsigned short **idx)
{
    size_t total = (size_t)nr_v * nr_v;
    size_t quads = nr_v ? (size_t)(nr_v - 1) * (nr_v - 1) : 0;
    int rc;

    /* the reference returns NULL when one of these allocations fails (terrain.c:487-488); here that is fatal */
    *vx = shim_alloc_zeroed(total * 3 * sizeof(float));
    *norm = shim_alloc_zeroed(total * 3 * sizeof(float));
    *tx = shim_alloc_zeroed(total * 2 * sizeof(float));
    *idx = shim_alloc_zeroed((quads ? quads : 1) * 6 * sizeof(unsigned short));
    shim_require_gpu();
    rc = clapca_terrain_mesh(map, nr_v, x, y, z, side, *vx, *norm, *tx, *idx);
    if (rc != CLAPCA_OK)
        shim_fatal("clapca_terrain_mesh", rc);
}
