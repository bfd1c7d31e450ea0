/* shim_common.c -- see shim_common.h */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "clapca.h"
#include "shim_common.h"

void shim_fatal(const char *what, int status)
{
    fprintf(stderr, "clapca: %s failed (status %d): %s\n", what, status, clapca_last_error());
    abort();
}

void shim_require_gpu(void)
{
    static int bound;
    const char *dev;
    int rc;

    if (bound)
        return;
    dev = getenv("CLAPCA_DEVICE");
    rc = clapca_init(dev ? atoi(dev) : 0);
    if (rc != CLAPCA_OK)
        shim_fatal("clapca_init (no CPU fallback exists for this path)", rc);
    bound = 1;
}

void *shim_alloc_zeroed(size_t bytes)
{
    void *p = calloc(1, bytes ? bytes : 1);

    if (!p) {
        fprintf(stderr, "clapca: out of host memory (%zu bytes)\n", bytes);
        abort();
    }
    return p;
}

uint64_t shim_rand48_peek(void)
{
    unsigned short probe[3] = { 0, 0, 0 }, cur[3];

    memcpy(cur, seed48(probe), sizeof(cur));
    seed48(cur);
    return (uint64_t)cur[0] | (uint64_t)cur[1] << 16 | (uint64_t)cur[2] << 32;
}

void shim_rand48_poke(uint64_t x)
{
    unsigned short v[3] = { (unsigned short)x, (unsigned short)(x >> 16), (unsigned short)(x >> 32) };

    seed48(v);
}
