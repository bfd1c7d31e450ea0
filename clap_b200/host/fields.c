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
