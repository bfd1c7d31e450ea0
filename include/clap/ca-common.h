/*
 * ca-common.h -- rule descriptor shared by the 2D and 3D automata.
 * Source-compatible with the reference's core/ca-common.h:10-32: same member
 * names, types and order (32 bytes on x86-64), so rule tables written for the
 * reference compile and run unchanged on top of libclapca_cuda.
 */
#ifndef CLAPCA_COMPAT_CA_COMMON_H
#define CLAPCA_COMPAT_CA_COMMON_H

#include <stdbool.h>

struct xyzarray;

struct cell_automaton {
    const char      *name;
    unsigned int    born_mask;  /* bit n set: a dead cell with n counted neighbours is born */
    unsigned int    surv_mask;  /* bit n set: a live cell with n counted neighbours keeps its value */
    unsigned int    nr_states;  /* value given to a newborn cell (3D: nr_states - 1) */
    bool            decay;      /* 2D only: non-surviving live cells lose 1; 3D always decays */
    union {                     /* neighbour counter; the 2D array is an xyzarray payload */
        int         (*neigh_2d)(unsigned char *arr, int x, int y);
        int         (*neigh_3d)(struct xyzarray *xyz, int x, int y, int z);
    };
};

#endif
