/*
 * xyarray.c -- host container for CA grids, behaviour of the reference's
 * core/xyarray.c: zeroed allocation (:8-19), bounds-checked access where
 * out-of-range reads return 0 and writes are ignored (:21-51), population
 * count (:68-78) and the 2D facade that hands out the payload pointer of a
 * {w, w, 1} volume (:80-109).  Pure host memory management: the generation
 * work on these grids happens in ca2d.c / ca3d.c via libclapca_cuda.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#include "xyarray.h"
#include "shim_common.h"

#define GRID_OF(_payload) ((struct xyzarray *)((char *)(_payload) - offsetof(struct xyzarray, arr)))

static inline size_t cell_index(const struct xyzarray *g, const int *p)
{
    return ((size_t)p[2] * g->dim[1] + p[1]) * g->dim[0] + p[0];
}

struct xyzarray *xyzarray_new(ivec3 dim)
{
    size_t cells = (size_t)dim[0] * dim[1] * dim[2];
    struct xyzarray *g = shim_alloc_zeroed(offsetof(struct xyzarray, arr) + cells);

    for (int i = 0; i < 3; i++)
        g->dim[i] = dim[i];
    return g;
}

bool xyzarray_valid(struct xyzarray *g, ivec3 p)
{
    return p[0] >= 0 && p[0] < g->dim[0] &&
           p[1] >= 0 && p[1] < g->dim[1] &&
           p[2] >= 0 && p[2] < g->dim[2];
}

/* same (quirky) predicate as core/xyarray.c:30-37: coordinate 1 or dim-1 on any axis */
bool xyzarray_edgemost(struct xyzarray *g, ivec3 p)
{
    for (int i = 0; i < 3; i++)
        if (p[i] == 1 || p[i] + 1 == g->dim[i])
            return true;
    return false;
}

int xyzarray_get(struct xyzarray *g, ivec3 p)
{
    return xyzarray_valid(g, p) ? g->arr[cell_index(g, p)] : 0;
}

void xyzarray_set(struct xyzarray *g, ivec3 p, int val)
{
    if (xyzarray_valid(g, p))
        g->arr[cell_index(g, p)] = (unsigned char)val;
}

int xyzarray_count(struct xyzarray *g)
{
    size_t cells = (size_t)g->dim[0] * g->dim[1] * g->dim[2];
    int alive = 0;

    for (size_t i = 0; i < cells; i++)
        alive += g->arr[i] != 0;
    return alive;
}

void xyzarray_print(struct xyzarray *g)
{
    char *line = shim_alloc_zeroed((size_t)g->dim[0] + 1);

    for (int z = 0; z < g->dim[2]; z++)
        for (int y = 0; y < g->dim[1]; y++) {
            for (int x = 0; x < g->dim[0]; x++)
                line[x] = xyzarray_get(g, (ivec3){ x, y, z }) ? '#' : ' ';
            fprintf(stderr, " #%d# |%s|\n", z, line);
        }
    free(line);
}

unsigned char *xyarray_new(int width)
{
    return xyzarray_new((ivec3){ width, width, 1 })->arr;
}

void xyarray_free(unsigned char *arr)
{
    free(GRID_OF(arr));
}

unsigned char xyarray_get(unsigned char *arr, int x, int y)
{
    return (unsigned char)xyzarray_get(GRID_OF(arr), (ivec3){ x, y, 0 });
}

void xyarray_set(unsigned char *arr, int x, int y, unsigned char v)
{
    xyzarray_set(GRID_OF(arr), (ivec3){ x, y, 0 }, v);
}

void xyarray_print(unsigned char *arr)
{
    static const char glyph[] = " .+oO############_^tTF";
    struct xyzarray *g = GRID_OF(arr);

    for (int y = 0; y < g->dim[1]; y++) {
        fprintf(stderr, "arr[%02d]: ", y);
        for (int x = 0; x < g->dim[0]; x++) {
            unsigned v = xyarray_get(arr, x, y);
            fprintf(stderr, "%c ", v < sizeof(glyph) - 1 ? glyph[v] : '?');
        }
        fputc('\n', stderr);
    }
}
