/*
 * oracle_port.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's procedural-generation hot path
 * (virtuoso/clap core/ca2d.c, core/ca3d.c, core/xyarray.c, core/noise.c,
 * core/terrain.c:15-91,447-467), with 64-bit sizes/indices so shapes beyond
 * the reference's 32-bit `int` limit can be checked.
 *
 * Parity status: PINNED -- every function here is compared bit-for-bit (CA,
 * lattice, RGBA8) or to 0 ulp (noise floats) against oracle/_ref/libclapref.so,
 * i.e. the unmodified reference sources compiled in the build container, by
 * tests/test_oracle_vs_reference.py, and against the committed fixtures under
 * tests/golden/ (generated from that same library by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path never does.
 */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include <stdint.h>
#include <stddef.h>

enum { ORA_NEIGH_VN1 = 0, ORA_NEIGH_M1 = 1, ORA_NEIGH_VNV = 2, ORA_NEIGH_MV = 3 };

/* glibc rand48 family (srand48/lrand48/drand48), explicit state */
void     ora_srand48(uint64_t *state, long seed);
long     ora_lrand48(uint64_t *state);
double   ora_drand48(uint64_t *state);

/* core/ca2d.c */
void ora_ca2d_seed(uint8_t *arr, int64_t side, unsigned nr_states, uint64_t *rng);
void ora_ca2d_step(uint8_t *arr, int64_t w, int64_t h, int64_t side,
                   unsigned born, unsigned surv, unsigned nr_states, int decay, int neigh);
void ora_ca2d_run(uint8_t *arr, int64_t w, int64_t h, int64_t side,
                  unsigned born, unsigned surv, unsigned nr_states, int decay, int neigh, int steps);

/* core/ca3d.c */
int     ora_ca3d_rule(int nca, unsigned *surv, unsigned *born, unsigned *nr_states);
int64_t ora_ca3d_run(uint8_t *arr, int64_t d0, int64_t d1, int64_t d2,
                     unsigned surv, unsigned born, unsigned nr_states, int steps);
int64_t ora_count(const uint8_t *arr, int64_t n);
void    ora_ca3d_make(uint8_t *arr, int d0, int d1, int d2, uint64_t *rng);

/* core/noise.h, core/noise.c */
float ora_hash31(int x, int y, int z, uint32_t seed);
float ora_value_noise3d_periodic(float x, float y, float z, int period, uint32_t seed);
float ora_fbm3_periodic(float x, float y, float z, int octaves, float lacunarity, float gain,
                        int period, uint32_t seed);
void  ora_noise_grad3d_bake_rgba8(uint8_t *out, size_t size, size_t z0, size_t z1, int octaves,
                                  float lacunarity, float gain, float period_units, uint32_t seed);

/*
 * blue_noise2d_tex(), core/noise.c:96-169, up to the upload: 64 x 64 RGBA32F pixels from the drand48 stream *rng.
 * PARITY UNPINNED for the transform: the reference calls kissfft (deps/bootstrap.json: mborgerding/kissfft@7bce4153,
 * not vendored, absent here), so this restates its documented contract -- forward DFT sum x[n] exp(-2 pi i k n / N),
 * unnormalised inverse -- as a plain double-precision DFT; everything around it (draw order, weights, gain,
 * normalisation) follows the reference line by line.
 */
void  ora_blue_noise2d(float *rgba, uint64_t *rng);

/* core/terrain.c */
float ora_get_rand_height(long seed, int x, int z);
void  ora_terrain_map0(long seed, unsigned nr_v, float *map0);
void  ora_terrain_field(unsigned nr_v, const float *map0, float ty, float amp, int oct,
                        unsigned i0, unsigned i1, float *map);
void  ora_terrain_heightmap(unsigned nr_v, const float *map0, float ty, const uint8_t *maze,
                            unsigned mside, unsigned i0, unsigned i1, float *map);

/* mesh buffers, core/terrain.c:93-110,479-516; any output may be NULL */
void  ora_terrain_mesh(const float *map, unsigned nr_v, float x, float y, float z, float side,
                       unsigned i0, unsigned i1, float *vx, float *norm, float *tx, unsigned short *idx);

/* terrain_height() core/terrain.c:336-379 and the instantiator loop :555-570 */
float  ora_terrain_height(const float *map, unsigned nr_vert, float t_x, float t_z, unsigned t_side, float x, float z);
size_t ora_terrain_instantiators(const uint8_t *maze, unsigned mside, const unsigned *nr_states, int nkinds,
                                 const float *map, unsigned nr_v, float x, float z, float side,
                                 void *out, size_t cap);

uint64_t ora_fnv1a64(const void *buf, size_t n);
#endif
