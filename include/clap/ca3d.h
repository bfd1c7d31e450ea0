/*
 * ca3d.h -- 3D cellular automaton entry points, source-compatible with the
 * reference's core/ca3d.h:7-52: neighbour-count mask bits CA_0..CA_26,
 * CA_RANGE(a, b) = bits a..b-1, the rule enumeration and the four functions.
 * ca3d_run() executes on the GPU through libclapca_cuda (clapca_ca3d_run).
 */
#ifndef CLAPCA_COMPAT_CA3D_H
#define CLAPCA_COMPAT_CA3D_H

#include "ca-common.h"

#define CA_BIT(_n)              (1 << (_n))
#define CA_0  CA_BIT(0)
#define CA_1  CA_BIT(1)
#define CA_2  CA_BIT(2)
#define CA_3  CA_BIT(3)
#define CA_4  CA_BIT(4)
#define CA_5  CA_BIT(5)
#define CA_6  CA_BIT(6)
#define CA_7  CA_BIT(7)
#define CA_8  CA_BIT(8)
#define CA_9  CA_BIT(9)
#define CA_10 CA_BIT(10)
#define CA_11 CA_BIT(11)
#define CA_12 CA_BIT(12)
#define CA_13 CA_BIT(13)
#define CA_14 CA_BIT(14)
#define CA_15 CA_BIT(15)
#define CA_16 CA_BIT(16)
#define CA_17 CA_BIT(17)
#define CA_18 CA_BIT(18)
#define CA_19 CA_BIT(19)
#define CA_20 CA_BIT(20)
#define CA_21 CA_BIT(21)
#define CA_22 CA_BIT(22)
#define CA_23 CA_BIT(23)
#define CA_24 CA_BIT(24)
#define CA_25 CA_BIT(25)
#define CA_26 CA_BIT(26)
/* neighbour counts _start .. _end - 1 */
#define CA_RANGE(_start, _end)  (((1 << ((_end) - (_start))) - 1) << (_start))

/* index into the built-in rule table, as taken by ca3d_run() */
enum {
    ca_445m = 0,
    ca_678_678_3m,
    ca_pyroclastic,
    ca_amoeba,
    ca_builder,
    ca_slow_decay,
    ca_spiky_growth,
    ca_coral,
    ca_crystal_1,
    CA3D_MAX
};

int ca3d_neighbors_vn1(struct xyzarray *xyz, int x, int y, int z);
int ca3d_prune(struct xyzarray *xyz);
int ca3d_run(struct xyzarray *xyz, int nca, int steps);
struct xyzarray *ca3d_make(int d0, int d1, int d2);

#endif
