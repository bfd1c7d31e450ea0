/*
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Pulls the reference's
 * core/noise.c into this translation unit *by inclusion from where it lies
 * under /root/reference* so that its file-static field functions become
 * callable from the parity tests.  No reference source is copied here.
 */
#include "noise.c"      /* resolved through -I$(REF)/core */

float ref_hash31(int x, int y, int z, uint32_t seed) { return hash31(x, y, z, seed); }

float ref_value_noise3d_periodic(float x, float y, float z, int period, uint32_t seed)
{
    return value_noise3d_periodic(x, y, z, period, seed);
}

float ref_fbm3_periodic(float x, float y, float z, int octaves, float lacunarity, float gain,
                        int period, uint32_t seed)
{
    return fbm3_periodic(x, y, z, octaves, lacunarity, gain, period, seed);
}

/* cresp(void) unwrapped: returns the mem_alloc()ed RGBA8 buffer or NULL */
void *ref_noise_grad3d_bake_rgba8(size_t size, int octaves, float lacunarity, float gain,
                                  float period_units, uint32_t seed)
{
    cresp(void) res = noise_grad3d_bake_rgba8(size, octaves, lacunarity, gain, period_units, seed);
    if (IS_CERR(res))
        return NULL;
    return res.val;
}

void ref_mem_free(void *p) { mem_free(p); }
