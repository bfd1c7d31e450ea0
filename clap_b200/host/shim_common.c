/* shim_common.c -- see shim_common.h */
#include <stdio.h>
#include <stdlib.h>
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
