/*
 * ca3d.c -- host side of the 3D automaton, API of the reference's
 * core/ca3d.c.  ca3d_run() (:124-142) executes on the GPU; the seed builder
 * ca3d_make()/walk/prune (:41-99, :144-169) stays on the host because it is a
 * short random walk driven by the process-wide lrand48() stream.  Its quirks
 * are kept on purpose: prune marks enclosed cells with
 * (unsigned char)-1 == 255 and then never finds an int equal to -1, so the
 * marks stay; callers therefore see 255-valued cells and the GPU engines
 * handle the full uint8 state range.
 */
#include <stdlib.h>
#include <string.h>
#include "ca3d.h"
#include "xyarray.h"
#include "clapca.h"
#include "shim_common.h"

int ca3d_neighbors_vn1(struct xyzarray *xyz, int x, int y, int z)
{
    static const signed char axis[6][3] = {
        { -1, 0, 0 }, { 1, 0, 0 }, { 0, -1, 0 }, { 0, 1, 0 }, { 0, 0, -1 }, { 0, 0, 1 },
    };
    int n = 0;

    for (int i = 0; i < 6; i++)
        n += !!xyzarray_get(xyz, (ivec3){ x + axis[i][0], y + axis[i][1], z + axis[i][2] });
    return n;
}

int ca3d_prune(struct xyzarray *xyz)
{
    int removed = 0;

    for (int z = 0; z < xyz->dim[2]; z++)
        for (int y = 0; y < xyz->dim[1]; y++)
            for (int x = 0; x < xyz->dim[0]; x++)
                if (ca3d_neighbors_vn1(xyz, x, y, z) == 6)
                    xyzarray_set(xyz, (ivec3){ x, y, z }, -1);          /* stored as 255 */
    for (int z = 0; z < xyz->dim[2]; z++)
        for (int y = 0; y < xyz->dim[1]; y++)
            for (int x = 0; x < xyz->dim[0]; x++)
                if (xyzarray_get(xyz, (ivec3){ x, y, z }) == -1) {      /* 0..255 never equals -1 */
                    xyzarray_set(xyz, (ivec3){ x, y, z }, 0);
                    removed++;
                }
    return removed;
}

/* random walk from the centre writing `val`, back-tracking through a bounded history */
static int carve_walk(struct xyzarray *xyz, int steps, int val)
{
    enum { HISTORY = 128, ATTEMPTS = 12 };
    ivec3 trail[HISTORY], at = { xyz->dim[0] / 2, xyz->dim[1] / 2, xyz->dim[2] / 2 };
    int depth = 0;

    for (int s = 0; s < steps; s++) {
        ivec3 to;
        int k;

        xyzarray_set(xyz, at, val);
        for (k = 0; k < ATTEMPTS; k++) {
            int axis, delta;

            memcpy(to, at, sizeof(to));
            axis = (int)(lrand48() % 3);
            delta = (lrand48() & 1) ? 1 : -1;
            to[axis] += delta;
            if (xyzarray_valid(xyz, to) && !xyzarray_get(xyz, to))
                break;
        }
        if (k == ATTEMPTS) {
            /* dead end: step back (the reference reads trail[-1] when depth is 0; we stay put) */
            if (depth > 0)
                memcpy(at, trail[--depth], sizeof(at));
            continue;
        }
        if (depth == HISTORY)
            continue;
        memcpy(trail[depth++], to, sizeof(to));
        memcpy(at, to, sizeof(to));
    }
    ca3d_prune(xyz);
    return xyzarray_count(xyz);
}

struct xyzarray *ca3d_make(int d0, int d1, int d2)
{
    struct xyzarray *xyz = xyzarray_new((ivec3){ d0, d1, d2 });
    int a = d0 * d1, b = d1 * d2, c = d0 * d2;
    int steps = a < b ? (a < c ? a : c) : (b < c ? b : c);

    /* the six faces start alive (value 5) */
    for (int x = 0; x < d0; x++)
        for (int y = 0; y < d1; y++) {
            xyzarray_set(xyz, (ivec3){ x, y, 0 }, 5);
            xyzarray_set(xyz, (ivec3){ x, y, d2 - 1 }, 5);
        }
    for (int x = 0; x < d0; x++)
        for (int z = 0; z < d2; z++) {
            xyzarray_set(xyz, (ivec3){ x, 0, z }, 5);
            xyzarray_set(xyz, (ivec3){ x, d1 - 1, z }, 5);
        }
    for (int y = 0; y < d1; y++)
        for (int z = 0; z < d2; z++) {
            xyzarray_set(xyz, (ivec3){ 0, y, z }, 5);
            xyzarray_set(xyz, (ivec3){ d0 - 1, y, z }, 5);
        }
    carve_walk(xyz, steps, 5);
    return xyz;
}

int ca3d_run(struct xyzarray *xyz, int nca, int steps)
{
    const int64_t dim[3] = { xyz->dim[0], xyz->dim[1], xyz->dim[2] };
    uint32_t surv, born, nr_states;
    int64_t population = 0;
    int rc;

    shim_require_gpu();
    rc = clapca_ca3d_rule(nca, &surv, &born, &nr_states);      /* nca % 9, core/ca3d.c:126 */
    if (rc == CLAPCA_OK)
        rc = clapca_ca3d_run(xyz->arr, dim, surv, born, nr_states, steps < 0 ? 0 : steps,
                             CLAPCA_ENGINE_AUTO, &population);
    if (rc != CLAPCA_OK)
        shim_fatal("clapca_ca3d_run", rc);
    return (int)population;
}
