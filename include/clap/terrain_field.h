/*
 * terrain_field.h -- the heightmap field of core/terrain.c as callable
 * functions.  In the reference these are file-static helpers
 * (get_rand_height .. get_height, terrain.c:15-91) and the two fill loops in
 * the middle of terrain_init_square_landscape() (terrain.c:447-467); the GPU
 * library computes the whole lattice and the whole map in one call each.
 */
#ifndef CLAPCA_COMPAT_TERRAIN_FIELD_H
#define CLAPCA_COMPAT_TERRAIN_FIELD_H

/* t->map0: nr_v * nr_v floats, map0[x*nr_v + z] (terrain.c:447-450); caller frees with free() */
float *clap_terrain_map0(long seed, unsigned int nr_v);

/*
 * t->map: nr_v * nr_v floats (terrain.c:451-467).  `maze` is the xyarray
 * returned by ca2d_generate(&ca_test, nr_v / 8, 4); `y` is t->y.
 */
float *clap_terrain_heightmap(long seed, unsigned int nr_v, float y, unsigned char *maze);

/*
 * Mesh buffers (terrain.c:479-516, calc_normal :93-110) from the finished map: *vx / *norm get 3 floats per
 * vertex, *tx 2, *idx 6 unsigned shorts per quad, each a malloc()ed buffer the caller frees -- what the
 * reference hands to mesh_attr_add().  `x,y,z,side` are the arguments of terrain_init_square_landscape().
 */
void clap_terrain_mesh(const float *map, unsigned int nr_v, float x, float y, float z, float side,
                       float **vx, float **norm, float **tx, unsigned short **idx);

#endif
