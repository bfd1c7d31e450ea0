/*
 * ca2d.h -- 2D cellular automaton entry points, source-compatible with the
 * reference's core/ca2d.h:8-13.  ca2d_step()/ca2d_generate() run on the GPU
 * through libclapca_cuda (clapca_ca2d_run); the four neighbour counters stay
 * callable on the host and double as the rule's neighbourhood selector.
 */
#ifndef CLAPCA_COMPAT_CA2D_H
#define CLAPCA_COMPAT_CA2D_H

#include <stdbool.h>
#include "ca-common.h"

int ca2d_neigh_vn1(unsigned char *arr, int x, int y);
int ca2d_neigh_m1(unsigned char *arr, int x, int y);
int ca2d_neigh_vnv(unsigned char *arr, int x, int y);
int ca2d_neigh_mv(unsigned char *arr, int x, int y);
void ca2d_step(const struct cell_automaton *ca, unsigned char *arr, int side);
unsigned char *ca2d_generate(const struct cell_automaton *ca, int side, int steps);

#endif
