/*
 * oracle_port.c -- TEST INFRASTRUCTURE ONLY (see oracle_port.h for the rules).
 *
 * Plain-C restatement of the reference hot path, written from the behaviour of
 * the reference sources; every function names the reference file:line it
 * follows.  Index arithmetic is 64-bit, everything else keeps the reference's
 * types so that results are bit-identical wherever the reference can run.
 *
 * Build with -ffp-contract=off: the reference is built for baseline x86-64
 * (no FMA), so no multiply-add is ever fused there.
 */
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdlib.h>
#include "oracle_port.h"

/* ------------------------------------------------------------------------- *
 * glibc rand48: X(n+1) = (a X(n) + c) mod 2^48, a = 0x5DEECE66D, c = 0xB.
 * srand48(s): X = (low32(s) << 16) | 0x330E.  lrand48() = X >> 17.
 * drand48() = X * 2^-48.  (glibc stdlib/srand48_r.c, drand48-iter.c,
 * lrand48_r.c / nrand48_r.c, erand48_r.c; used by core/ca2d.c:88,
 * core/ca3d.c:80-81, core/terrain.c:17-18.)
 * ------------------------------------------------------------------------- */
#define R48_A    0x5DEECE66DULL
#define R48_C    0xBULL
#define R48_MASK 0xFFFFFFFFFFFFULL

void ora_srand48(uint64_t *state, long seed)
{
    *state = ((((uint64_t)seed) & 0xFFFFFFFFULL) << 16) | 0x330EULL;
}

static inline uint64_t r48_next(uint64_t *state)
{
    *state = (*state * R48_A + R48_C) & R48_MASK;
    return *state;
}

long ora_lrand48(uint64_t *state)
{
    return (long)(r48_next(state) >> 17);
}

double ora_drand48(uint64_t *state)
{
    return ldexp((double)r48_next(state), -48);
}

/* ------------------------------------------------------------------------- *
 * Grid access: core/xyarray.c:21-28 (valid), :39-44 (get, OOB -> 0),
 * :46-51 (set, OOB dropped).  Layout z*d0*d1 + y*d0 + x (xyarray.c:43).
 * ------------------------------------------------------------------------- */
typedef struct { uint8_t *a; int64_t d0, d1, d2; } grid_t;

static inline int g_get(const grid_t *g, int64_t x, int64_t y, int64_t z)
{
    if (x < 0 || x >= g->d0 || y < 0 || y >= g->d1 || z < 0 || z >= g->d2)
        return 0;
    return g->a[(z * g->d1 + y) * g->d0 + x];
}

static inline void g_set(const grid_t *g, int64_t x, int64_t y, int64_t z, int v)
{
    if (x < 0 || x >= g->d0 || y < 0 || y >= g->d1 || z < 0 || z >= g->d2)
        return;
    g->a[(z * g->d1 + y) * g->d0 + x] = (uint8_t)v;
}

/* xyzarray_count(): core/xyarray.c:68-78 */
int64_t ora_count(const uint8_t *arr, int64_t n)
{
    int64_t c = 0;
    for (int64_t i = 0; i < n; i++)
        c += arr[i] != 0;
    return c;
}

/* ------------------------------------------------------------------------- *
 * 2D automaton: core/ca2d.c
 * ------------------------------------------------------------------------- */

/* ca2d_neigh_vn1 :11-21, _m1 :23-33, _vnv :35-46, _mv :48-59 */
static int neigh2d(const grid_t *g, int64_t x, int64_t y, int kind)
{
    static const int off[8][2] = {
        { 1, 0 }, { -1, 0 }, { 0, 1 }, { 0, -1 },       /* von Neumann */
        { 1, 1 }, { -1, 1 }, { 1, -1 }, { -1, -1 },     /* + diagonals = Moore */
    };
    int cnt = (kind == ORA_NEIGH_VN1 || kind == ORA_NEIGH_VNV) ? 4 : 8;
    int by_value = (kind == ORA_NEIGH_VNV || kind == ORA_NEIGH_MV);
    int v = g_get(g, x, y, 0), n = 0;

    for (int i = 0; i < cnt; i++) {
        int c = g_get(g, x + off[i][0], y + off[i][1], 0);
        n += by_value ? (c > v) : (c != 0);
    }
    return n;
}

/*
 * ca2d_step(): core/ca2d.c:61-77.  The reference reads the true array extent
 * through container_of (w,h here) and uses `side` only as the loop bound, with
 * x as the OUTER loop although memory is y*w + x.
 */
void ora_ca2d_step(uint8_t *arr, int64_t w, int64_t h, int64_t side,
                   unsigned born, unsigned surv, unsigned nr_states, int decay, int neigh)
{
    grid_t g = { arr, w, h, 1 };

    for (int64_t x = 0; x < side; x++)
        for (int64_t y = 0; y < side; y++) {
            int n = neigh2d(&g, x, y, neigh);
            int v = g_get(&g, x, y, 0);

            if (!v && (born & (1u << n)))
                g_set(&g, x, y, 0, (uint8_t)nr_states);
            else if (v && (surv & (1u << n)))
                ;
            else if (v && decay)
                g_set(&g, x, y, 0, v - 1);
        }
}

void ora_ca2d_run(uint8_t *arr, int64_t w, int64_t h, int64_t side,
                  unsigned born, unsigned surv, unsigned nr_states, int decay, int neigh, int steps)
{
    for (int s = 0; s < steps; s++)
        ora_ca2d_step(arr, w, h, side, born, surv, nr_states, decay, neigh);
}

/* seeding loop of ca2d_generate(): core/ca2d.c:86-90 (x outer, y inner) */
void ora_ca2d_seed(uint8_t *arr, int64_t side, unsigned nr_states, uint64_t *rng)
{
    for (int64_t x = 0; x < side; x++)
        for (int64_t y = 0; y < side; y++) {
            int v = (int)(ora_lrand48(rng) % 8);
            /* `v <= ca->nr_states` compares int with unsigned: v is never negative */
            arr[y * side + x] = (unsigned)v <= nr_states ? (uint8_t)nr_states : 0;
        }
}

/* ------------------------------------------------------------------------- *
 * 3D automaton: core/ca3d.c
 * ------------------------------------------------------------------------- */
#define B(n) (1u << (n))
#define RANGE(s, e) (((1u << ((e) - (s))) - 1) << (s))     /* core/ca3d.h:34: bits s..e-1 */

/* rule table cas[]: core/ca3d.c:110-122 -- (surv, born, nr_states) */
int ora_ca3d_rule(int nca, unsigned *surv, unsigned *born, unsigned *nr_states)
{
    static const unsigned tab[9][3] = {
        /* 445m         */ { B(4), B(4), 5 },
        /* 678_678_3m   */ { B(6) | B(7) | B(8), B(6) | B(7) | B(8), 3 },
        /* pyroclastic  */ { B(4) | B(5) | B(6) | B(7), B(6) | B(7) | B(8), 10 },
        /* amoeba       */ { RANGE(9, 26), B(5) | B(6) | B(7) | B(12) | B(13) | B(15), 5 },
        /* builder      */ { B(2) | B(6) | B(9), B(4) | B(6) | B(8) | B(9), 10 },
        /* slow_decay   */ { B(1) | B(4) | B(8) | B(11) | RANGE(13, 26), RANGE(13, 26), 5 },
        /* spiky_growth */ { RANGE(0, 3) | RANGE(7, 9) | RANGE(11, 13) | B(18) | B(21) | B(22) | B(24) | B(26),
                             B(4) | B(13) | B(17) | RANGE(20, 24) | B(26), 4 },
        /* coral        */ { RANGE(5, 8), RANGE(6, 7) | B(9) | B(12), 4 },
        /* crystal_1    */ { RANGE(0, 6), B(1) | B(3), 2 },
    };
    int i = (int)((size_t)(long)nca % 9u);              /* ca3d.c:126: the modulo is taken in size_t */
    *surv = tab[i][0];
    *born = tab[i][1];
    *nr_states = tab[i][2];
    return i;
}

/* ca3d_neighbors_m1(): core/ca3d.c:29-39 (26-cell Moore, alive bits) */
static int moore3d(const grid_t *g, int64_t x, int64_t y, int64_t z)
{
    int n = 0;
    for (int64_t cz = z - 1; cz < z + 2; cz++)
        for (int64_t cy = y - 1; cy < y + 2; cy++)
            for (int64_t cx = x - 1; cx < x + 2; cx++)
                n += !!g_get(g, cx, cy, cz);
    return n - !!g_get(g, x, y, z);
}

/*
 * ca3d_run(): core/ca3d.c:124-142.  In place, z/y/x order, always Moore
 * (the rule's neigh_3d pointer is not consulted), decay unconditional.
 */
int64_t ora_ca3d_run(uint8_t *arr, int64_t d0, int64_t d1, int64_t d2,
                     unsigned surv, unsigned born, unsigned nr_states, int steps)
{
    grid_t g = { arr, d0, d1, d2 };

    for (; steps; steps--)
        for (int64_t z = 0; z < d2; z++)
            for (int64_t y = 0; y < d1; y++)
                for (int64_t x = 0; x < d0; x++) {
                    int n = moore3d(&g, x, y, z);
                    int s = g_get(&g, x, y, z);

                    if (s && !(surv & (1u << n)))
                        g_set(&g, x, y, z, s - 1);
                    else if (!s && (born & (1u << n)))
                        g_set(&g, x, y, z, (int)(nr_states - 1));
                }
    return ora_count(arr, d0 * d1 * d2);
}

/* ca3d_neighbors_vn1(): core/ca3d.c:15-27 */
static int vn3d(const grid_t *g, int64_t x, int64_t y, int64_t z)
{
    return !!g_get(g, x - 1, y, z) + !!g_get(g, x + 1, y, z) +
           !!g_get(g, x, y - 1, z) + !!g_get(g, x, y + 1, z) +
           !!g_get(g, x, y, z - 1) + !!g_get(g, x, y, z + 1);
}

/*
 * ca3d_make(): core/ca3d.c:144-169, ca3d_walk(): :63-99, ca3d_prune(): :41-59.
 * Faithful to the quirks: prune stores (unsigned char)-1 = 255
 * and its second pass compares an int 255 with -1, so nothing is cleared.
 * The reference reads history[-1] when it has to roll back with an empty
 * history (undefined behaviour); here that case keeps the current position.
 */
void ora_ca3d_make(uint8_t *arr, int d0, int d1, int d2, uint64_t *rng)
{
    enum { HIST = 128, TRIES = 12 };
    grid_t g = { arr, d0, d1, d2 };
    int a = d0 * d1, b = d1 * d2, c = d0 * d2;
    int steps = a < b ? (a < c ? a : c) : (b < c ? b : c);
    int hist[HIST][3], histp = 0, cur[3] = { d0 / 2, d1 / 2, d2 / 2 };

    memset(arr, 0, (size_t)d0 * d1 * d2);
    for (int x = 0; x < d0; x++)
        for (int y = 0; y < d1; y++) {
            g_set(&g, x, y, 0, 5);
            g_set(&g, x, y, d2 - 1, 5);
        }
    for (int x = 0; x < d0; x++)
        for (int z = 0; z < d2; z++) {
            g_set(&g, x, 0, z, 5);
            g_set(&g, x, d1 - 1, z, 5);
        }
    for (int y = 0; y < d1; y++)
        for (int z = 0; z < d2; z++) {
            g_set(&g, 0, y, z, 5);
            g_set(&g, d0 - 1, y, z, 5);
        }

    for (int step = 0; step < steps; step++) {
        int next[3], try;

        g_set(&g, cur[0], cur[1], cur[2], 5);
        for (try = 0; try < TRIES; try++) {
            memcpy(next, cur, sizeof(cur));
            int dir = (int)(ora_lrand48(rng) % 3);
            next[dir] += (ora_lrand48(rng) & 1) ? 1 : -1;
            if (next[0] >= 0 && next[0] < d0 && next[1] >= 0 && next[1] < d1 &&
                next[2] >= 0 && next[2] < d2 && !g_get(&g, next[0], next[1], next[2]))
                break;
        }
        if (try == TRIES) {             /* roll back */
            if (histp > 0)
                memcpy(cur, hist[--histp], sizeof(cur));
            continue;
        }
        if (histp == HIST)
            continue;
        memcpy(hist[histp++], next, sizeof(next));
        memcpy(cur, next, sizeof(next));
    }

    /* ca3d_prune(), first pass only has an effect */
    for (int z = 0; z < d2; z++)
        for (int y = 0; y < d1; y++)
            for (int x = 0; x < d0; x++)
                if (vn3d(&g, x, y, z) == 6)
                    g_set(&g, x, y, z, -1);
}

/* ------------------------------------------------------------------------- *
 * interp.h numerics: float arguments meet double literals, so the arithmetic
 * is carried out in double and rounded to float on return.
 * ------------------------------------------------------------------------- */

/* smoothf(): core/interp.h:11-14 -- x*x is a float product, the rest double */
static inline float p_smoothf(float x)
{
    float xx = x * x;
    return (float)((double)xx * (3.0 - 2.0 * (double)x));
}

/* linf_interp(): core/interp.h:25-29 -- b*blend is a float product */
static inline float p_linf(float a, float b, float blend)
{
    float bb = b * blend;
    return (float)((double)a * (1.0 - (double)blend) + (double)bb);
}

/* cosf_interp(): core/interp.h:35-42 */
static inline float p_cosf_interp(float a, float b, float blend)
{
    float theta = (float)((double)blend * M_PI);
    float f = (float)((1.0 - (double)cosf(theta)) / 2.0);
    float bf = b * f;
    return (float)((double)a * (1.0 - (double)f) + (double)bf);
}

/* ------------------------------------------------------------------------- *
 * noise.c field
 * ------------------------------------------------------------------------- */

/* hash31(): core/noise.h:9-17 */
float ora_hash31(int x, int y, int z, uint32_t seed)
{
    uint32_t h = (uint32_t)x * 374761393u + (uint32_t)y * 668265263u +
                 (uint32_t)z * 362437u + seed * 2246822519u;
    h = (h ^ (h >> 13)) * 1274126177u;
    return (float)(h ^ (h >> 16)) * (1.0f / 4294967296.0f);
}

static inline int wrap(int v, int period)
{
    return (v % period + period) % period;
}

/* value_noise3d_periodic(): core/noise.c:171-202 */
float ora_value_noise3d_periodic(float x, float y, float z, int period, uint32_t seed)
{
    int x0 = (int)floorf(x), y0 = (int)floorf(y), z0 = (int)floorf(z);
    float xf = x - x0, yf = y - y0, zf = z - z0;
    int x1 = wrap(x0 + 1, period), y1 = wrap(y0 + 1, period), z1 = wrap(z0 + 1, period);

    x0 = wrap(x0, period);
    y0 = wrap(y0, period);
    z0 = wrap(z0, period);

    float ux = p_smoothf(xf), uy = p_smoothf(yf), uz = p_smoothf(zf);
    float lo0 = p_linf(ora_hash31(x0, y0, z0, seed), ora_hash31(x1, y0, z0, seed), ux);
    float lo1 = p_linf(ora_hash31(x0, y1, z0, seed), ora_hash31(x1, y1, z0, seed), ux);
    float hi0 = p_linf(ora_hash31(x0, y0, z1, seed), ora_hash31(x1, y0, z1, seed), ux);
    float hi1 = p_linf(ora_hash31(x0, y1, z1, seed), ora_hash31(x1, y1, z1, seed), ux);

    return p_linf(p_linf(lo0, lo1, uy), p_linf(hi0, hi1, uy), uz);
}

/* fbm3_periodic(): core/noise.c:204-220 */
float ora_fbm3_periodic(float x, float y, float z, int octaves, float lacunarity, float gain,
                        int period, uint32_t seed)
{
    float amp = 0.5f, sum = 0.0f;

    for (int i = 0; i < octaves; i++) {
        sum += ora_value_noise3d_periodic(x, y, z, period, seed + (uint32_t)i) * amp;
        x *= lacunarity;
        y *= lacunarity;
        z *= lacunarity;
        period = (int)lrintf((float)period * lacunarity);
        amp *= gain;
    }
    return sum;
}

static inline uint8_t pack_unorm8(float g)
{
    return (uint8_t)lrintf((g * 0.5f + 0.5f) * 255.0f);
}

/*
 * noise_grad3d_bake_rgba8(): core/noise.c:222-270, z-slices [z0,z1) written at
 * their final offsets in `out` (size^3 * 4 bytes, x fastest).
 */
void ora_noise_grad3d_bake_rgba8(uint8_t *out, size_t size, size_t z0, size_t z1, int octaves,
                                 float lacunarity, float gain, float period_units, uint32_t seed)
{
    const float step = period_units / (float)size;
    const float eps = step;
    const int period = (int)period_units;
    const float scale = 0.5f / eps;

    for (size_t z = z0; z < z1; z++) {
        float pz = z * step;
        for (size_t y = 0; y < size; y++) {
            float py = y * step;
            uint8_t *o = out + ((z * size + y) * size) * 4;
            for (size_t x = 0; x < size; x++, o += 4) {
                float px = x * step;
                float gx = (ora_fbm3_periodic(px + eps, py, pz, octaves, lacunarity, gain, period, seed) -
                            ora_fbm3_periodic(px - eps, py, pz, octaves, lacunarity, gain, period, seed)) * scale;
                float gy = (ora_fbm3_periodic(px, py + eps, pz, octaves, lacunarity, gain, period, seed) -
                            ora_fbm3_periodic(px, py - eps, pz, octaves, lacunarity, gain, period, seed)) * scale;
                float gz = (ora_fbm3_periodic(px, py, pz + eps, octaves, lacunarity, gain, period, seed) -
                            ora_fbm3_periodic(px, py, pz - eps, octaves, lacunarity, gain, period, seed)) * scale;
                float len2 = gx * gx + gy * gy + gz * gz;
                float inv = 1.0f / sqrtf(len2 > FLT_MIN ? len2 : FLT_MIN);

                o[0] = pack_unorm8(gx * inv);
                o[1] = pack_unorm8(gy * inv);
                o[2] = pack_unorm8(gz * inv);
                o[3] = 0;
            }
        }
    }
}

/* ------------------------------------------------------------------------- *
 * terrain.c heightmap chain
 * ------------------------------------------------------------------------- */

/* get_rand_height(): core/terrain.c:15-19 -- int arithmetic inside, long xor */
float ora_get_rand_height(long seed, int x, int z)
{
    uint64_t st;
    ora_srand48(&st, seed ^ (long)(x + z * 43210));
    return (float)(ora_drand48(&st) * 2 - 1);
}

void ora_terrain_map0(long seed, unsigned nr_v, float *map0)
{
    for (unsigned i = 0; i < nr_v; i++)                  /* terrain.c:447-450 */
        for (unsigned j = 0; j < nr_v; j++)
            map0[(size_t)i * nr_v + j] = ora_get_rand_height(seed, (int)i, (int)j);
}

typedef struct { const float *map0; int nr; float y; } terr_t;

/* get_mapped_rand_height(): core/terrain.c:21-33 ("torus": one-step clamp-wrap) */
static inline float lattice(const terr_t *t, int x, int z)
{
    if (x < 0) x = t->nr - 1; else if (x >= t->nr) x = 0;
    if (z < 0) z = t->nr - 1; else if (z >= t->nr) z = 0;
    return t->map0[(size_t)x * t->nr + z];
}

/* get_avg_height(): core/terrain.c:35-54 */
static float smooth3x3(const terr_t *t, int x, int z)
{
    float corners, sides, self;

    corners  = lattice(t, x - 1, z - 1);
    corners += lattice(t, x + 1, z - 1);
    corners += lattice(t, x - 1, z + 1);
    corners += lattice(t, x + 1, z + 1);
    corners /= 16.f;
    sides  = lattice(t, x - 1, z);
    sides += lattice(t, x + 1, z);
    sides += lattice(t, x, z - 1);
    sides += lattice(t, x, z + 1);
    sides /= 8.f;
    self = lattice(t, x, z) / 4.f;
    return corners + sides + self;
}

/* get_interp_height(): core/terrain.c:56-71 */
static float interp_height(const terr_t *t, float x, float z)
{
    int ix = (int)floor(x), iz = (int)floor(z);
    float fx = x - ix, fz = z - iz;
    float v1 = smooth3x3(t, ix, iz), v2 = smooth3x3(t, ix + 1, iz);
    float v3 = smooth3x3(t, ix, iz + 1), v4 = smooth3x3(t, ix + 1, iz + 1);

    return p_cosf_interp(p_cosf_interp(v1, v2, fx), p_cosf_interp(v3, v4, fx), fz);
}

/* get_height(): core/terrain.c:77-91 (ROUGHNESS 0.5f) */
static float octave_height(const terr_t *t, int x, int z, float amp0, int oct)
{
    float total = 0;
    float d = (float)pow(2, oct - 1);

    for (int i = 0; i < oct; i++) {
        float freq = (float)(pow(2, i) / d);
        float amp = (float)(pow(0.5f, i) * amp0);
        total += interp_height(t, x * freq, z * freq) * amp;
    }
    return t->y + total;
}

void ora_terrain_field(unsigned nr_v, const float *map0, float ty, float amp, int oct,
                       unsigned i0, unsigned i1, float *map)
{
    terr_t t = { map0, (int)nr_v, ty };
    for (unsigned i = i0; i < i1; i++)
        for (unsigned j = 0; j < nr_v; j++)
            map[(size_t)i * nr_v + j] = octave_height(&t, (int)i, (int)j, amp, oct);
}

/*
 * Map fill of terrain_init_square_landscape(): core/terrain.c:451-467
 * (MAZE_FAC 8, OCTAVES 4).  `maze` is an mside x mside xyarray payload.
 */
void ora_terrain_heightmap(unsigned nr_v, const float *map0, float ty, const uint8_t *maze,
                           unsigned mside, unsigned i0, unsigned i1, float *map)
{
    terr_t t = { map0, (int)nr_v, ty };
    grid_t mz = { (uint8_t *)maze, mside, mside, 1 };

    for (int i = (int)i0; i < (int)i1; i++)
        for (int j = 0; j < (int)nr_v; j++) {
            float fi = fmodf(i, 8) / 8, fj = fmodf(j, 8) / 8;
            int mi = i / 8, mj = j / 8;
            uint8_t cn = (uint8_t)g_get(&mz, mi, mj, 0);
            uint8_t xn = (uint8_t)g_get(&mz, fi >= 0.5 ? mi + 1 : mi - 1, mj, 0);
            uint8_t yn = (uint8_t)g_get(&mz, mi, fj >= 0.5 ? mj + 1 : mj - 1, 0);
            float xa = cn > xn ? (float)cn : p_cosf_interp(cn, xn, 2 * fi - 1);
            float ya = cn > yn ? (float)cn : p_cosf_interp(cn, yn, 2 * fj - 1);
            float avg = p_cosf_interp(xa, ya, fabsf(fi - fj));

            map[(size_t)i * nr_v + j] = octave_height(&t, i, j, powf(1.5, avg), 4) + avg;
        }
}

/*
 * calc_normal(): core/terrain.c:93-110.  The torus indices left/right/up/down are used only
 * away from the edges; at an edge the neighbour height counts as 0.  vec3_norm() is
 * linmath.h:48-62: p = 0; p += v[i]*v[i]; k = 1.0 / sqrtf(p) (double division, stored to float);
 * r[i] = v[i] * k.
 */
static void p_calc_normal(const float *map, int nr, float n[3], int x, int z)
{
    float hl = x == 0 ? 0 : map[(size_t)(x - 1) * nr + z];
    float hr = x == nr - 1 ? 0 : map[(size_t)(x + 1) * nr + z];
    float hd = z == 0 ? 0 : map[(size_t)x * nr + (z - 1)];
    float hu = z == nr - 1 ? 0 : map[(size_t)x * nr + (z + 1)];
    float v[3] = { hl - hr, 2.f, hd - hu };
    float p = 0.f;
    for (int i = 0; i < 3; i++)
        p += v[i] * v[i];
    float k = 1.0 / sqrtf(p);
    for (int i = 0; i < 3; i++)
        n[i] = v[i] * k;
}

/*
 * Mesh buffers of terrain_init_square_landscape(): core/terrain.c:479-516 (vertex rows [i0,i1),
 * quad rows [i0, min(i1, nr_v-1))).  Outputs are indexed from the start of the full buffers.
 */
void ora_terrain_mesh(const float *map, unsigned nr_v, float x, float y, float z, float side,
                      unsigned i0, unsigned i1, float *vx, float *norm, float *tx, unsigned short *idx)
{
    const int nr = (int)nr_v;
    for (int i = (int)i0; i < (int)i1; i++)
        for (int j = 0; j < nr; j++) {
            size_t it = (size_t)i * nr + j;
            float n[3];
            if (vx) {
                vx[it * 3 + 0] = x + (float)j / ((float)nr_v - 1) * side;
                vx[it * 3 + 1] = y + map[(size_t)j * nr + i];
                vx[it * 3 + 2] = z + (float)i / ((float)nr_v - 1) * side;
            }
            if (norm) {
                p_calc_normal(map, nr, n, j, i);
                norm[it * 3 + 0] = n[0];
                norm[it * 3 + 1] = n[1];
                norm[it * 3 + 2] = n[2];
            }
            if (tx) {
                tx[it * 2 + 0] = (float)j * 32 / ((float)nr_v - 1);
                tx[it * 2 + 1] = (float)i * 32 / ((float)nr_v - 1);
            }
        }
    if (!idx)
        return;
    for (int i = (int)i0; i < (int)i1 && i < nr - 1; i++)
        for (int j = 0; j < nr - 1; j++) {
            size_t it = ((size_t)i * (nr - 1) + j) * 6;
            int top_left = i * nr + j, top_right = top_left + 1;
            int bottom_left = (i + 1) * nr + j, bottom_right = bottom_left + 1;
            idx[it + 0] = (unsigned short)top_left;
            idx[it + 1] = (unsigned short)bottom_left;
            idx[it + 2] = (unsigned short)top_right;
            idx[it + 3] = (unsigned short)top_right;
            idx[it + 4] = (unsigned short)bottom_left;
            idx[it + 5] = (unsigned short)bottom_right;
        }
}

/* barrycentric(): core/interp.h:49-56 */
static float p_barrycentric(const float p1[3], const float p2[3], const float p3[3], const float pos[2])
{
    float det = (p2[2] - p3[2]) * (p1[0] - p3[0]) + (p3[0] - p2[0]) * (p1[2] - p3[2]);
    float l1  = ((p2[2] - p3[2]) * (pos[0] - p3[0]) + (p3[0] - p2[0]) * (pos[1] - p3[2])) / det;
    float l2  = ((p3[2] - p1[2]) * (pos[0] - p3[0]) + (p1[0] - p3[0]) * (pos[1] - p3[2])) / det;
    float l3  = 1.0f - l1 - l2;
    return l1 * p1[1] + l2 * p2[1] + l3 * p3[1];
}

/* terrain_height(): core/terrain.c:336-379; t->side is an unsigned int (terrain.h:19) */
float ora_terrain_height(const float *map, unsigned nr_vert, float t_x, float t_z, unsigned t_side, float x, float z)
{
    float square = (float)t_side / (nr_vert - 1);
    float tx = x - t_x;
    float tz = z - t_z;
    int gridx = floorf(tx / square);
    int gridz = floorf(tz / square);
    float xoff = (tx - square * gridx) / square;
    float zoff = (tz - square * gridz) / square;
    float pos[2] = { xoff, zoff };

    if (!map)
        return 0;
    if (x < t_x || x > t_x + t_side || z < t_z || z > t_z + t_side)
        return 0;
    if (xoff <= 1 - zoff) {
        float p1[3] = { 0, map[(size_t)gridx * nr_vert + gridz], 0 };
        float p2[3] = { 1, map[(size_t)(gridx + 1) * nr_vert + gridz], 0 };
        float p3[3] = { 0, map[(size_t)gridx * nr_vert + gridz + 1], 1 };
        return p_barrycentric(p1, p2, p3, pos);
    } else {
        float p1[3] = { 1, map[(size_t)(gridx + 1) * nr_vert + gridz], 0 };
        float p2[3] = { 1, map[(size_t)(gridx + 1) * nr_vert + gridz + 1], 1 };
        float p3[3] = { 0, map[(size_t)gridx * nr_vert + gridz + 1], 1 };
        return p_barrycentric(p1, p2, p3, pos);
    }
}

/*
 * Instantiator loop of terrain_init_square_landscape(): core/terrain.c:555-570.  out = records of
 * { kind, dx, dy, dz } (4 x 32 bit); returns the number found, writes at most cap.
 */
size_t ora_terrain_instantiators(const uint8_t *maze, unsigned mside, const unsigned *nr_states, int nkinds,
                                 const float *map, unsigned nr_v, float x, float z, float side,
                                 void *out, size_t cap)
{
    grid_t mz = { (uint8_t *)maze, mside, mside, 1 };
    struct rec { int32_t kind; float dx, dy, dz; } *o = out;
    size_t n = 0;

    for (int i = 0; i < (int)mside; i++)
        for (int j = 0; j < (int)mside; j++)
            for (int ca = 0; ca < nkinds; ca++)
                if ((unsigned)g_get(&mz, i, j, 0) == nr_states[ca]) {
                    float dx = x + (float)(i + 0.5) * 8 * side / (nr_v - 1);
                    float dz = z + (float)(j + 0.5) * 8 * side / (nr_v - 1);
                    if (n < cap) {
                        o[n].kind = ca;
                        o[n].dx = dx;
                        o[n].dz = dz;
                        o[n].dy = ora_terrain_height(map, nr_v, x, z, (unsigned)side, dx, dz);
                    }
                    n++;
                }
    return n;
}

/* FNV-1a 64 over a byte buffer (fixture fingerprints) */
/* ------------------------------------------------------------------------- *
 * blue_noise2d_tex(): core/noise.c:17-169.  kissfft is not in the tree (see the
 * header): the two transforms are plain O(N^2) DFTs per line in double precision,
 * rounded to float where the reference stores floats (kiss_fft_scalar = float).
 * ------------------------------------------------------------------------- */
enum { ORA_GRAIN = 64 };        /* FILM_GRAIN_SIZE, core/shader_constants.h:13 */

static void ora_dft64_lines(float (*re)[ORA_GRAIN], float (*im)[ORA_GRAIN], int columns, int inverse)
{
    static double cs[ORA_GRAIN], sn[ORA_GRAIN];
    static int have;
    if (!have) {
        for (int m = 0; m < ORA_GRAIN; m++) {
            cs[m] = cos(2.0 * M_PI * m / ORA_GRAIN);
            sn[m] = sin(2.0 * M_PI * m / ORA_GRAIN);
        }
        have = 1;
    }
    for (int l = 0; l < ORA_GRAIN; l++) {
        double xr[ORA_GRAIN], xi[ORA_GRAIN];
        for (int e = 0; e < ORA_GRAIN; e++) {
            xr[e] = columns ? re[e][l] : re[l][e];
            xi[e] = columns ? im[e][l] : im[l][e];
        }
        for (int k = 0; k < ORA_GRAIN; k++) {
            double ar = 0.0, ai = 0.0;
            for (int n = 0; n < ORA_GRAIN; n++) {
                const int m = (k * n) % ORA_GRAIN;
                const double c = cs[m], s = inverse ? sn[m] : -sn[m];
                ar += xr[n] * c - xi[n] * s;
                ai += xr[n] * s + xi[n] * c;
            }
            if (columns) { re[k][l] = (float)ar; im[k][l] = (float)ai; }
            else         { re[l][k] = (float)ar; im[l][k] = (float)ai; }
        }
    }
}

void ora_blue_noise2d(float *rgba, uint64_t *rng)
{
    const int size = ORA_GRAIN;
    /* noise.c:106-115 */
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++) {
            float r = ((ora_drand48(rng) * 4.0 - 1.0) / 3.0) * 0.299;
            float g = ((ora_drand48(rng) * 4.0 - 1.0) / 3.0) * 0.587;
            float b = ((ora_drand48(rng) * 4.0 - 1.0) / 3.0) * 0.114;
            rgba[(x + y * size) * 4 + 0] = r;
            rgba[(x + y * size) * 4 + 1] = g;
            rgba[(x + y * size) * 4 + 2] = b;
            rgba[(x + y * size) * 4 + 3] = 1.0;
        }
    static float re[ORA_GRAIN][ORA_GRAIN], im[ORA_GRAIN][ORA_GRAIN];
    for (int c = 0; c < 3; c++) {
        for (int i = 0; i < size * size; i++) {
            re[i / size][i % size] = rgba[i * 4 + c];
            im[i / size][i % size] = 0.0f;
        }
        ora_dft64_lines(re, im, 0, 0);              /* fft2d_fwd, noise.c:17-45: rows, then columns */
        ora_dft64_lines(re, im, 1, 0);
        /* blue_noise2d_gain, noise.c:75-92 */
        float maxr = sqrtf((ORA_GRAIN / 2) * (ORA_GRAIN / 2) + (ORA_GRAIN / 2) * (ORA_GRAIN / 2));
        for (int y = 0; y < ORA_GRAIN; y++) {
            int fy = (y <= ORA_GRAIN / 2) ? y : y - ORA_GRAIN;
            for (int x = 0; x < ORA_GRAIN; x++) {
                int fx = (x <= ORA_GRAIN / 2) ? x : x - ORA_GRAIN;
                float r = sqrt(fx * fx + fy * fy);
                float gain = r / maxr;
                re[y][x] *= gain;
                im[y][x] *= gain;
            }
        }
        ora_dft64_lines(re, im, 0, 1);              /* fft2d_inv, noise.c:47-73 */
        ora_dft64_lines(re, im, 1, 1);
        for (int i = 0; i < size * size; i++)
            rgba[i * 4 + c] = re[i / size][i % size] / (ORA_GRAIN * ORA_GRAIN);
    }
    /* noise.c:141-152 */
    float minv = INFINITY, maxv = -INFINITY;
    for (int i = 0; i < size * size * 4; i++) {
        if ((i & 3) == 3) continue;
        if (rgba[i] < minv) minv = rgba[i];
        if (rgba[i] > maxv) maxv = rgba[i];
    }
    for (int i = 0; i < size * size * 4; i++)
        if ((i & 3) != 3)
            rgba[i] = (rgba[i] - minv) / (maxv - minv);
}

uint64_t ora_fnv1a64(const void *buf, size_t n)
{
    const uint8_t *p = buf;
    uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < n; i++) {
        h ^= p[i];
        h *= 0x100000001b3ULL;
    }
    return h;
}
