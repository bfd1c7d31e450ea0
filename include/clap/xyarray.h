/*
 * xyarray.h -- byte-per-cell grid container, source-compatible with the
 * reference's core/xyarray.h:5-24.  A 2D grid is the payload pointer of a
 * {side, side, 1} xyzarray (core/xyarray.c:80-88); cells live at
 * z*d0*d1 + y*d0 + x (core/xyarray.c:43).  Out-of-range reads give 0,
 * out-of-range writes are dropped.
 */
#ifndef CLAPCA_COMPAT_XYARRAY_H
#define CLAPCA_COMPAT_XYARRAY_H

#include <stdbool.h>

typedef int ivec3[3];

struct xyzarray {
    ivec3           dim;
    unsigned char   arr[0];
};

struct xyzarray *xyzarray_new(ivec3 dim);
bool xyzarray_valid(struct xyzarray *xyz, ivec3 pos);
bool xyzarray_edgemost(struct xyzarray *xyz, ivec3 pos);
int  xyzarray_get(struct xyzarray *xyz, ivec3 pos);
void xyzarray_set(struct xyzarray *xyz, ivec3 pos, int val);
void xyzarray_print(struct xyzarray *xyz);
int  xyzarray_count(struct xyzarray *xyz);

unsigned char *xyarray_new(int width);
void xyarray_free(unsigned char *arr);
unsigned char xyarray_get(unsigned char *arr, int x, int y);
void xyarray_set(unsigned char *arr, int x, int y, unsigned char v);
void xyarray_print(unsigned char *arr);

#endif
