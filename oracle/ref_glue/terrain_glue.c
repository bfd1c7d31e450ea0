/*
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Includes the reference's
 * core/terrain.c from where it lies so the file-static heightmap chain
 * (get_rand_height .. get_height, terrain.c:15-91) is reachable.  The map
 * fill below drives those reference functions in the order of the loop at
 * terrain.c:447-467, which cannot be called on its own because it lives in
 * the middle of terrain_init_square_landscape() (needs a scene + renderer).
 */
#include "terrain.c"    /* resolved through -I$(REF)/core */

float ref_get_rand_height(long seed, int x, int z)
{
    struct terrain t = { .seed = seed };
    return get_rand_height(&t, x, z);
}

void ref_terrain_map0(long seed, unsigned int nr_v, float *map0)
{
    struct terrain t = { .seed = seed, .nr_vert = nr_v };
    for (unsigned int i = 0; i < nr_v; i++)
        for (unsigned int j = 0; j < nr_v; j++)
            map0[(size_t)i * nr_v + j] = get_rand_height(&t, i, j);
}

float ref_get_height(long seed, unsigned int nr_v, float *map0, float ty, int x, int z, float amp, int oct)
{
    struct terrain t = { .seed = seed, .nr_vert = nr_v, .map0 = map0, .y = ty };
    return get_height(&t, x, z, amp, oct);
}

/* plain octave field: map[i*nr_v+j] = get_height(i, j, amp, oct) for rows [i0,i1) */
void ref_terrain_field(long seed, unsigned int nr_v, float *map0, float ty, float amp, int oct,
                       unsigned int i0, unsigned int i1, float *map)
{
    struct terrain t = { .seed = seed, .nr_vert = nr_v, .map0 = map0, .y = ty };
    for (unsigned int i = i0; i < i1; i++)
        for (unsigned int j = 0; j < nr_v; j++)
            map[(size_t)i * nr_v + j] = get_height(&t, i, j, amp, oct);
}

/*
 * Maze-modulated fill for vertex rows [i0,i1).  Drives the reference's own
 * statics/inlines (get_height, cosf_interp, xyarray_get, MAZE_FAC, OCTAVES) with
 * the operand types and evaluation order of the loop body at terrain.c:459-467.
 */
static float glue_blend_toward(unsigned char self, unsigned char other, float frac)
{
    /* a higher cell is kept flat, otherwise cosine-blend toward the neighbour */
    return self > other ? self : cosf_interp(self, other, 2 * frac - 1);
}

void ref_terrain_heightmap(long seed, unsigned int nr_v, float *map0, float ty, unsigned char *maze,
                           unsigned int i0, unsigned int i1, float *map)
{
    struct terrain t = { .seed = seed, .nr_vert = nr_v, .map0 = map0, .y = ty };
    int i, j;

    for (i = i0; i < (int)i1; i++) {
        float fi = fmodf(i, MAZE_FAC) / MAZE_FAC;
        int mi = i / MAZE_FAC;

        for (j = 0; j < (int)nr_v; j++) {
            float fj = fmodf(j, MAZE_FAC) / MAZE_FAC;
            int mj = j / MAZE_FAC;
            unsigned char here  = xyarray_get(maze, mi, mj);
            unsigned char along = xyarray_get(maze, fi >= 0.5 ? mi + 1 : mi - 1, mj);
            unsigned char cross = xyarray_get(maze, mi, fj >= 0.5 ? mj + 1 : mj - 1);
            float a = glue_blend_toward(here, along, fi);
            float b = glue_blend_toward(here, cross, fj);
            float avg = cosf_interp(a, b, fabsf(fi - fj));

            map[(size_t)i * nr_v + j] = get_height(&t, i, j, powf(1.5, avg), OCTAVES) + avg;
        }
    }
}

/*
 * Normals of the mesh stage: drives the reference's own calc_normal() (terrain.c:93-110) in the order and
 * with the argument swap of the vertex loop at terrain.c:493-500 (vertex it = i*nr_v + j uses calc_normal(t, n, j, i)).
 */
void ref_terrain_normals(unsigned int nr_v, float *map, float *norm)
{
    struct terrain t = { .nr_vert = nr_v, .map = map };
    size_t it = 0;

    for (unsigned int i = 0; i < nr_v; i++)
        for (unsigned int j = 0; j < nr_v; j++, it++) {
            vec3 n;
            calc_normal(&t, n, j, i);
            norm[it * 3 + 0] = n[0];
            norm[it * 3 + 1] = n[1];
            norm[it * 3 + 2] = n[2];
        }
}

/* the reference's own terrain_height() (terrain.c:336-379) on a caller-supplied map */
float ref_terrain_height(float *map, unsigned int nr_v, float t_x, float t_z, unsigned int t_side, float x, float z)
{
    struct terrain t = { .nr_vert = nr_v, .map = map, .x = t_x, .z = t_z, .side = t_side };
    return terrain_height(&t, x, z);
}

const struct cell_automaton *ref_ca_test(void) { return &ca_test; }
const struct cell_automaton *ref_ca_instor(int i) { return &ca_instors[i]; }
