/*
 * ca2d.c -- host side of the 2D automaton, API of the reference's
 * core/ca2d.c.  The generation sweep (ca2d_step, :61-77) is executed by
 * libclapca_cuda; the neighbour counters (:11-59) remain ordinary host
 * functions because (a) callers may use them directly and (b) their
 * addresses are how a rule names its neighbourhood (ca-common.h neigh_2d).
 */
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include "ca2d.h"
#include "xyarray.h"
#include "clapca.h"
#include "shim_common.h"

static const signed char ring[8][2] = {
    { 1, 0 }, { -1, 0 }, { 0, 1 }, { 0, -1 },           /* von Neumann */
    { 1, 1 }, { -1, 1 }, { 1, -1 }, { -1, -1 },         /* diagonals */
};

static int count_ring(unsigned char *arr, int x, int y, int first, int last, int threshold)
{
    int n = 0;

    for (int i = first; i < last; i++)
        n += xyarray_get(arr, x + ring[i][0], y + ring[i][1]) > threshold;
    return n;
}

int ca2d_neigh_vn1(unsigned char *arr, int x, int y) { return count_ring(arr, x, y, 0, 4, 0); }
int ca2d_neigh_m1(unsigned char *arr, int x, int y)  { return count_ring(arr, x, y, 0, 8, 0); }
int ca2d_neigh_vnv(unsigned char *arr, int x, int y) { return count_ring(arr, x, y, 0, 4, xyarray_get(arr, x, y)); }
int ca2d_neigh_mv(unsigned char *arr, int x, int y)  { return count_ring(arr, x, y, 0, 8, xyarray_get(arr, x, y)); }

/* the rule's neighbourhood arrives as a host function pointer: translate it */
static int neighbourhood_id(const struct cell_automaton *ca)
{
    if (ca->neigh_2d == ca2d_neigh_vn1) return CLAPCA_NEIGH_VN1;
    if (ca->neigh_2d == ca2d_neigh_m1)  return CLAPCA_NEIGH_M1;
    if (ca->neigh_2d == ca2d_neigh_vnv) return CLAPCA_NEIGH_VNV;
    if (ca->neigh_2d == ca2d_neigh_mv)  return CLAPCA_NEIGH_MV;
    fprintf(stderr, "clapca: rule '%s' uses a neighbour function the GPU library does not know; "
                    "there is no CPU fallback\n", ca->name ? ca->name : "?");
    abort();
}

static void run_steps(const struct cell_automaton *ca, unsigned char *arr, int side, int steps)
{
    struct xyzarray *g = (struct xyzarray *)((char *)arr - offsetof(struct xyzarray, arr));
    int rc;

    shim_require_gpu();
    rc = clapca_ca2d_run(arr, g->dim[0], g->dim[1], side, ca->born_mask, ca->surv_mask, ca->nr_states,
                         ca->decay, neighbourhood_id(ca), steps, CLAPCA_ENGINE_AUTO);
    if (rc != CLAPCA_OK)
        shim_fatal("clapca_ca2d_run", rc);
}

void ca2d_step(const struct cell_automaton *ca, unsigned char *arr, int side)
{
    run_steps(ca, arr, side, 1);
}

unsigned char *ca2d_generate(const struct cell_automaton *ca, int side, int steps)
{
    unsigned char *arr = xyarray_new(side);
    uint64_t after = 0;
    int rc;

    /*
     * Seeding consumes the process-wide lrand48() stream in x-outer / y-inner order exactly like
     * core/ca2d.c:86-90 -- on the device: every cell jumps ahead to its own draw from the stream's current
     * state, and the stream is left where side * side host calls would have left it, so a caller's
     * srand48() gives the same initial grid, and the same later draws, as with the reference.
     */
    if (side <= 0)
        return arr;
    shim_require_gpu();
    rc = clapca_ca2d_generate(arr, side, ca->born_mask, ca->surv_mask, ca->nr_states, ca->decay,
                              neighbourhood_id(ca), steps > 0 ? steps : 0, CLAPCA_ENGINE_AUTO, shim_rand48_peek(), &after);
    if (rc != CLAPCA_OK)
        shim_fatal("clapca_ca2d_generate", rc);
    shim_rand48_poke(after);
    return arr;
}
