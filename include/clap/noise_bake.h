/*
 * noise_bake.h -- host entry point for the fBm-gradient RGBA8 bake that
 * core/noise.c:222-270 (noise_grad3d_bake_rgba8) performs on one CPU core.
 * The reference wraps its result in the engine's cresp(void) error union
 * (core/error.h); outside the engine tree the same contract is a malloc()ed
 * buffer or NULL.  INTEGRATION.md shows the three-line body that turns this
 * into the reference's own signature.
 */
#ifndef CLAPCA_COMPAT_NOISE_BAKE_H
#define CLAPCA_COMPAT_NOISE_BAKE_H

#include <stddef.h>
#include <stdint.h>

/* size^3 * 4 bytes (x fastest; R,G,B = packed unit gradient, A = 0); caller frees with free() */
unsigned char *clap_noise_grad3d_bake_rgba8(size_t size, int octaves, float lacunarity, float gain,
                                            float period_units, uint32_t seed);

/*
 * The pixels blue_noise2d_tex() (core/noise.c:96-169) uploads as the film-grain texture: size x size RGBA32F, drawn
 * from the process-wide drand48() stream (3 draws per pixel) like the reference; size must be FILM_GRAIN_SIZE (64).
 * Caller frees with free().
 */
float *clap_blue_noise2d_rgba32f(int size);

#endif
