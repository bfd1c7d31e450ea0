/*
 * clapca.h -- C ABI of libclapca_cuda, the B200 (sm_100a) implementation of the
 * virtuoso/clap procedural-generation hot path.
 *
 * Every entry point is extern "C", takes plain pointers / sizes / scalars and
 * returns an int status (CLAPCA_OK == 0).  The library owns all device
 * buffers, streams and events; callers hand in host memory laid out exactly
 * like the reference's containers (struct xyzarray payload: uint8,
 * index = z*d0*d1 + y*d0 + x, core/xyarray.c:43) and get results back in the
 * same layout.  There is no CPU fallback anywhere behind this interface: if
 * no CUDA device is usable the calls fail with CLAPCA_ERR_CUDA.
 *
 * Each function cites the reference interface (file:line under
 * /root/reference) whose work it takes over.  The source-compatible C headers
 * a reference caller would keep including (ca-common.h, ca2d.h, ca3d.h,
 * xyarray.h, noise_bake.h, terrain_field.h) live in include/clap/ and are
 * implemented by the host shim clap_b200/host/ on top of this ABI
 * (see INTEGRATION.md).
 */
#ifndef CLAPCA_H
#define CLAPCA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------ */
enum {
    CLAPCA_OK              = 0,
    CLAPCA_ERR_CUDA        = 1,   /* a CUDA runtime/driver call failed (no device, launch error, ...) */
    CLAPCA_ERR_ARG         = 2,   /* bad argument (NULL, non-positive size, unknown enum) */
    CLAPCA_ERR_NOMEM       = 3,   /* device or pinned-host allocation failed */
    CLAPCA_ERR_TIMEOUT     = 4,   /* in-kernel dataflow watchdog fired (would have dead-locked) */
    CLAPCA_ERR_UNSUPPORTED = 5,   /* shape/rule outside what the requested engine handles */
    CLAPCA_ERR_STATE       = 6,   /* library not initialised / handle in the wrong state */
};

/* 2D neighbourhood selectors: the four functions of core/ca2d.c:11-59 */
enum {
    CLAPCA_NEIGH_VN1 = 0,         /* ca2d_neigh_vn1: alive count, 4 neighbours */
    CLAPCA_NEIGH_M1  = 1,         /* ca2d_neigh_m1 : alive count, 8 neighbours */
    CLAPCA_NEIGH_VNV = 2,         /* ca2d_neigh_vnv: neighbours with value > own, 4 */
    CLAPCA_NEIGH_MV  = 3,         /* ca2d_neigh_mv : neighbours with value > own, 8 */
};

/* kernel family used by the *_run calls */
enum {
    CLAPCA_ENGINE_AUTO      = 0,  /* fastest engine that supports the shape/rule */
    CLAPCA_ENGINE_WAVEFRONT = 1,  /* uint8 cells, skewed cell wavefront + grid barrier (any rule/shape) */
    CLAPCA_ENGINE_BITPLANE  = 2,  /* bit-sliced planes, row scan, flag dataflow, all generations fused */
    CLAPCA_ENGINE_DIAGONAL  = 3,  /* 2D, one-plane (binary) grids of at most 16384 columns: bit rows stored along the
                                     diagonals 2x + y, no in-row chain (ca2d_skew.cuh).  Runs when asked for (or with
                                     CLAPCA_2D_SKEW=1 under AUTO / BITPLANE) and is named in the run stats; the row
                                     engine is faster at BASELINE config 3 (8.9 vs 13.1 ms) and stays the default */
};

/* ---- library lifetime --------------------------------------------------- */

/* Number of visible CUDA devices (0 if none / driver missing). */
int clapca_device_count(void);

/*
 * Bind this process to `device` (one process per GPU), create the streams and
 * scratch the engines need.  Idempotent for the same device.
 *
 * Threading: the context (its stream, its population / maximum accumulators, its grow-only scratch buffers and the
 * grid it keeps for one-shot calls) is ONE per process and not locked -- call the grid, field and one-shot entry points
 * from one thread at a time, and do not re-initialise for another device while grids exist.  Slabs are the exception:
 * each owns its stream and accumulators, and clapca_slab_prepare / run / run_streamed of DIFFERENT slabs may run on
 * different threads (several ranks of one process, see clapca_slab_connect_local).  Only the error text is per thread.
 */
int clapca_init(int device);
void clapca_shutdown(void);

/* Human-readable description of the last non-OK status on this thread. */
const char *clapca_last_error(void);

/* SM count and HBM bytes of the bound device (0 before clapca_init). */
int    clapca_sm_count(void);
size_t clapca_device_mem_bytes(void);

/* ---- one-shot calls on host buffers (what the C shim binds) ------------- */

/*
 * ca3d_run(): core/ca3d.c:124-142.  `steps` in-place generations of the 3D
 * rule (surv_mask, born_mask, nr_states) in z/y/x sweep order with the 26-cell
 * Moore neighbourhood, on the uint8 volume `arr` (dim = {d0,d1,d2}, d0
 * fastest).  *population (optional) receives xyzarray_count() of the result
 * (core/xyarray.c:68-78).  Host -> device -> host copies are inside the call.
 */
int clapca_ca3d_run(uint8_t *arr, const int64_t dim[3],
                    uint32_t surv_mask, uint32_t born_mask, uint32_t nr_states,
                    int steps, int engine, int64_t *population);

/* Rule table cas[] of core/ca3d.c:110-122, indexed like ca3d_run (nca % 9). */
int clapca_ca3d_rule(int nca, uint32_t *surv_mask, uint32_t *born_mask, uint32_t *nr_states);

/*
 * ca2d_step() x steps: core/ca2d.c:61-77.  `arr` is a w*h uint8 grid
 * (index y*w + x); cells x,y < side are swept x-outer / y-inner in place.
 * The reference passes w == h == side.
 */
int clapca_ca2d_run(uint8_t *arr, int64_t w, int64_t h, int64_t side,
                    uint32_t born_mask, uint32_t surv_mask, uint32_t nr_states,
                    int decay, int neigh, int steps, int engine);

/*
 * ca2d_generate(): core/ca2d.c:79-98 with the seeding loop (:86-90) on the device as well.  The
 * reference draws one lrand48() % 8 per cell from the process-wide stream in x-outer / y-inner order;
 * `rand48_state` is that stream's current 48-bit state X (what seed48() reports), every cell jumps
 * ahead to its own draw (the LCG's affine map composes), and *rand48_state_after is the state after the
 * side*side draws so the caller can put the stream where the reference would have left it.  Then
 * `steps` generations; `arr` receives the side x side grid (index y*side + x).
 */
int clapca_ca2d_generate(uint8_t *arr, int64_t side, uint32_t born_mask, uint32_t surv_mask,
                         uint32_t nr_states, int decay, int neigh, int steps, int engine,
                         uint64_t rand48_state, uint64_t *rand48_state_after);

/*
 * noise_grad3d_bake_rgba8(): core/noise.c:222-270.  Fills out[size^3 * 4]
 * (x fastest, RGBA8, A = 0) with the normalised central-difference gradient
 * of the periodic fBm field (hash31 / value_noise3d_periodic / fbm3_periodic,
 * core/noise.h:9-17, core/noise.c:171-220).
 */
int clapca_noise_grad3d_bake_rgba8(uint8_t *out, size_t size, int octaves, float lacunarity,
                                   float gain, float period_units, uint32_t seed);

/*
 * blue_noise2d_tex() (core/noise.c:96-169) up to the texture upload: the size x size RGBA32F pixels of the film-grain
 * texture -- white noise from the caller's drand48() stream (3 draws per pixel, r g b, pixel x + y * size), forward
 * 2-D FFT per channel, radial gain r / r_max, inverse FFT, joint min / max normalisation, alpha = 1.  `size` must be
 * FILM_GRAIN_SIZE = 64 (core/shader_constants.h:13): the reference's spectrum arrays have that size whatever it is
 * called with.  rand48_state = the stream's 48-bit state before the first draw (what seed48() hands back);
 * *state_out = the state after the 3 * size^2 draws.  The reference transforms with kissfft (un-vendored, float
 * radix-4 butterflies), so its pixels are matched to float rounding (tests: 2e-5 of the [0, 1] range), not bit for bit.
 */
int clapca_noise_blue2d_rgba32f(float *out, int size, uint64_t rand48_state, uint64_t *state_out);
/* the same into device memory (size^2 * 4 floats) */
int clapca_noise_blue2d_device(void *d_out, int size, uint64_t rand48_state, uint64_t *state_out, float *kernel_ms);

/* fbm3_periodic() at n sample points (xyz interleaved), for parity tests of the float field. */
int clapca_noise_fbm3(float *out, const float *xyz, size_t n, int octaves, float lacunarity,
                      float gain, int period, uint32_t seed);

/*
 * Lattice of terrain_init_square_landscape(): core/terrain.c:447-450 with
 * get_rand_height() (:15-19): map0[x*nr_v + z] = drand48-after-srand48(seed ^ (x + z*43210))*2-1.
 */
int clapca_terrain_map0(float *map0, long seed, unsigned nr_v);

/*
 * Heightmap fill: core/terrain.c:451-467 on top of get_height() (:21-91).
 * map[i*nr_v + j] = get_height(i, j, powf(1.5, avg), 4) + avg, where avg is the
 * cosine blend of the `maze` cell (mside x mside xyarray payload, mside =
 * nr_v/8) and its neighbours.  The lattice is generated on the device from
 * `seed`.  With maze == NULL the plain octave field get_height(i, j, amp, oct)
 * is written instead (amp/oct are ignored when a maze is given: the reference
 * fixes OCTAVES = 4 there).
 */
int clapca_terrain_heightmap(float *map, long seed, unsigned nr_v, float ty,
                             const uint8_t *maze, unsigned mside, float amp, int oct);

/*
 * Terrain mesh buffers: core/terrain.c:479-519 with calc_normal() (:93-110).  From the nr_v x nr_v
 * heightmap t->map: vx[it*3..] = { x + j/(nr_v-1)*side, y + map[j*nr_v+i], z + i/(nr_v-1)*side },
 * norm[it*3..] = normalised { hl-hr, 2, hd-hu } (neighbours beyond an edge count as 0),
 * tx[it*2..] = { 32 j/(nr_v-1), 32 i/(nr_v-1) } for it = i*nr_v + j, and idx = two triangles per quad
 * (6 * (nr_v-1)^2 entries, truncated to unsigned short like the reference's buffer, terrain.h:13).
 * Any output pointer may be NULL.
 */
int clapca_terrain_mesh(const float *map, unsigned nr_v, float x, float y, float z, float side,
                        float *vx, float *norm, float *tx, unsigned short *idx);

/*
 * Instantiator extraction: core/terrain.c:555-570.  After the ca_instors[] passes over the maze
 * (terrain.c:473-477) every maze cell (i outer, j inner) whose value equals nr_states[k] spawns an
 * instantiator of kind k at dx = x + (i+0.5)*8*side/(nr_v-1), dz likewise with j, dy =
 * terrain_height(t, dx, dz) (terrain.c:336-379: barycentric interpolation of t->map, interp.h:49-56).
 * Records come out in the reference's list order.  *count receives the number found; at most `cap`
 * are written (call with cap == 0 to size the buffer).
 */
typedef struct clapca_instor {
    int32_t kind;                 /* index into the caller's rule table (ca_instors[], terrain.c:400-415) */
    float   dx, dy, dz;
} clapca_instor;

int clapca_terrain_instantiators(const uint8_t *maze, unsigned mside, const uint32_t *nr_states, int nkinds,
                                 const float *map, unsigned nr_v, float x, float z, float side,
                                 clapca_instor *out, size_t cap, size_t *count);

/* ---- device-resident grids (benchmarks, pipelines, multi-GPU slabs) ------ */

typedef struct clapca_grid clapca_grid;

/* A d0 x d1 x d2 uint8 grid in device memory (2D grids: d2 == 1). */
int clapca_grid_create(clapca_grid **out, int64_t d0, int64_t d1, int64_t d2);
int clapca_grid_destroy(clapca_grid *g);
/* host <-> device in the reference layout (async on the grid's stream + sync) */
int clapca_grid_upload(clapca_grid *g, const uint8_t *host);
int clapca_grid_download(clapca_grid *g, uint8_t *host);
/* raw device pointer of the uint8 cells (for zero-copy interop with torch tensors) */
void *clapca_grid_device_ptr(clapca_grid *g);
/* the CUDA stream (cudaStream_t) the grid's kernels are launched on */
void *clapca_grid_stream(clapca_grid *g);

int clapca_grid_run3d(clapca_grid *g, uint32_t surv_mask, uint32_t born_mask, uint32_t nr_states,
                      int steps, int engine, int64_t *population);
/* CLAPCA_ENGINE_AUTO / _BITPLANE on a binary grid of >= 4 M cells and >= 16 generations: the FIRST run of a (shape,
 * rule) class also runs the row and the diagonal engine on a scratch copy of the input (at most 32 generations each,
 * a few tens of milliseconds once per process) and later runs of the class take the one that was faster -- results are
 * identical either way.  CLAPCA_2D_TUNE=0 skips the measurement (row engine). */
int clapca_grid_run2d(clapca_grid *g, int64_t side, uint32_t born_mask, uint32_t surv_mask,
                      uint32_t nr_states, int decay, int neigh, int steps, int engine);
/* xyzarray_count(): core/xyarray.c:68-78 */
int clapca_grid_count(clapca_grid *g, int64_t *population);
/*
 * 64-bit fingerprint of every plane of a uint8 volume in DEVICE memory (nplanes planes of plane_bytes bytes each;
 * clapca_grid_device_ptr / clapca_slab_device_ptr): the sum over the plane's 8-byte little-endian words w_k of
 * splitmix64(w_k ^ (k + 1) * 0x9E3779B97F4A7C15), mod 2^64.  Lets a sharded run be compared plane by plane with a
 * single-GPU run (xyzarray contents, not just xyzarray_count()) without moving 8 GiB through the host.
 */
int clapca_hash_planes(const void *d_cells, size_t plane_bytes, size_t nplanes, uint64_t *hashes);
/* the seeding loop of ca2d_generate() (core/ca2d.c:86-90) into a device-resident 2D grid; see clapca_ca2d_generate */
int clapca_grid_seed2d(clapca_grid *g, int64_t side, uint32_t nr_states, uint64_t rand48_state,
                       uint64_t *rand48_state_after);

/*
 * ca3d_make() (core/ca3d.c:144-169) into a device-resident grid: faces of value 5, the random walk of ca3d_walk()
 * (:63-99) from the centre, ca3d_prune() (:41-59, marks of 255 included).  The walk is a serial chain of lrand48() draws:
 * it runs on the host against a sparse picture of the volume (faces + its own cells); faces, scatter and prune are
 * kernels.  rand48_state in / out as for clapca_ca2d_generate; *population = xyzarray_count() of the result.
 */
int clapca_grid_make3d(clapca_grid *g, uint64_t rand48_state, uint64_t *rand48_state_after, int64_t *population);

/*
 * ca3d_run() (core/ca3d.c:124-142) from host memory to host memory as ONE pipeline: the volume is copied in
 * chunk by chunk while the sweep kernel already runs, the kernel itself converts the layout plane by plane
 * (no separate pack / unpack pass), every generation follows the upload front a few planes behind, and
 * finished planes are copied out while later planes are still being swept -- H2D, all generations and D2H of
 * one volume overlap instead of adding up.  host_in / host_out are d0*d1*d2 uint8 cells in the reference
 * layout (core/xyarray.c:43) and may be the same buffer.  `max_value` is an upper bound on the input cells
 * (it fixes the number of state bit planes before the data has arrived); the kernel verifies it and the call
 * fails with CLAPCA_ERR_ARG -- output undefined -- if a cell exceeds it; 255 is always safe.  Buffers that
 * are not page-locked (cudaHostAlloc / cudaHostRegister), and shapes the bit-plane engine does not take, run
 * the same three steps one after the other; the result is identical either way.
 */
int clapca_grid_run3d_streamed(clapca_grid *g, const uint8_t *host_in, uint8_t *host_out, unsigned max_value,
                               uint32_t surv_mask, uint32_t born_mask, uint32_t nr_states, int steps,
                               int64_t *population);

/*
 * Timing of the last *_run on this grid, measured with CUDA events on the
 * grid's stream: total device milliseconds, the share spent in the
 * generation kernel(s) alone (excludes layout pack/unpack and the count), the
 * number of kernel launches, and the engine that actually ran.
 */
typedef struct clapca_run_stats {
    float   total_ms;
    float   kernel_ms;
    int     launches;
    int     engine;
    int     planes;          /* bit planes used by the bit-plane engine (0 otherwise) */
    int     workers;         /* persistent warps (bit-plane) / threads per wavefront step */
    int     streamed;        /* 1: the run overlapped H2D / sweep / D2H (clapca_grid_run3d_streamed) */
} clapca_run_stats;
int clapca_grid_last_stats(clapca_grid *g, clapca_run_stats *st);

/* ---- multi-GPU: z-block slabs of one ca3d volume, one process per GPU ------------------ */

/*
 * The d2_global planes of the volume are cut into blocks of `block_planes` planes, block j lives on
 * rank j % nranks (contiguous slabs: block_planes = ceil(d2_global / nranks)).  The in-place sweep
 * order of ca3d_run() (core/ca3d.c:129-140) makes plane z of generation g depend on plane z-1 of
 * generation g and plane z+1 of generation g-1, so neighbouring ranks exchange one H-row per plane
 * row and generation.  That exchange is fused into the sweep kernel: edge planes store their rows
 * straight into the neighbour GPU's ghost plane over NVLink (CUDA IPC peer mapping) and raise its
 * progress counter; there is no separate halo kernel and no host round trip per generation.
 *
 * Call order on every rank: create -> ipc_handle -> (exchange handles) -> connect -> { upload ->
 * prepare -> (barrier across ranks) -> run -> download } repeated.  The barrier after prepare is the only one:
 * the ghost-plane counters alternate between two banks, so a rank may upload / prepare run k+1 while a slower
 * neighbour is still finishing run k.  `max_value` must be the largest cell value over ALL ranks (it fixes the
 * number of state planes; prepare verifies the uploaded cells against it and fails with CLAPCA_ERR_ARG),
 * `max_generations` bounds `steps`.  Every rank must call prepare the same number of times.
 */
typedef struct clapca_slab clapca_slab;

int clapca_slab_create(clapca_slab **out, int64_t d0, int64_t d1, int64_t d2_global, int rank, int nranks,
                       int block_planes, int max_generations, unsigned max_value);
int clapca_slab_destroy(clapca_slab *s);
/* number of planes this rank owns, and the global z of each of them (local storage order) */
int clapca_slab_local_planes(clapca_slab *s, int *n);
int clapca_slab_plane_map(clapca_slab *s, int64_t *zglobal);
/* device pointer of the local planes (uint8, d0*d1 bytes per plane, local order) */
void *clapca_slab_device_ptr(clapca_slab *s);
/* 64-byte CUDA IPC handle of this rank's halo region; connect() maps the two neighbours' regions */
int clapca_slab_ipc_handle(clapca_slab *s, void *handle64);
int clapca_slab_connect(clapca_slab *s, const void *handle_next_rank, const void *handle_prev_rank);
/*
 * Several slabs of ONE process on one device (tests, single-GPU boxes): instead of IPC handles the slabs exchange
 * their halo pointers directly.  The sweep launches of all ranks must be co-resident, so each keeps to
 * `max_ctas` CTAs (0 = no limit); every rank runs prepare / run on its own thread -- a slab owns its stream.
 */
void *clapca_slab_halo_ptr(clapca_slab *s);
int clapca_slab_connect_local(clapca_slab *s, void *halo_next_rank, void *halo_prev_rank, int max_ctas);
/* local planes from/to host or device memory, local order */
int clapca_slab_upload(clapca_slab *s, const uint8_t *src);
int clapca_slab_download(clapca_slab *s, uint8_t *dst);
int clapca_slab_prepare(clapca_slab *s, uint32_t surv_mask, uint32_t born_mask, uint32_t nr_states, int steps);
int clapca_slab_run(clapca_slab *s, int64_t *local_population);
/*
 * The same run from host memory to host memory as one pipeline per rank (cf. clapca_grid_run3d_streamed): no upload /
 * download calls -- host_in / host_out hold this rank's planes in local order in page-locked memory, the launch packs
 * the planes as their H2D chunks land, seeds the neighbours' ghost planes itself, and hands finished planes back
 * while later ones are still being swept.  Call order: prepare_streamed -> (barrier across ranks) -> run_streamed.
 * Cells above the max_value given to slab_create make the run fail with CLAPCA_ERR_ARG.
 */
int clapca_slab_prepare_streamed(clapca_slab *s, uint32_t surv_mask, uint32_t born_mask, uint32_t nr_states, int steps);
int clapca_slab_run_streamed(clapca_slab *s, const uint8_t *host_in, uint8_t *host_out, int64_t *local_population);
int clapca_slab_last_stats(clapca_slab *s, clapca_run_stats *st);

/*
 * SURVEY 8(f).3 -- the gradient-noise bake as a device-resident 3D texture: noise_grad3d_bake_rgba8_tex()
 * (core/noise.c:272-294) bakes on the host and uploads with texture_load(); here the kernel writes the RGBA8 texels
 * straight into a 3D CUDA array through a surface.  clapca_tex3d_array() returns the cudaArray_t -- the object a
 * renderer's TEX_3D / TEX_FMT_RGBA8 texture is mapped to under CUDA / GL or Vulkan interop, so the texels never visit
 * the host; clapca_tex3d_download() reads it back (size^3 * 4 bytes, x fastest) for whoever wants the host copy.
 */
typedef struct clapca_tex3d clapca_tex3d;
int clapca_noise_bake_array(clapca_tex3d **out, size_t size, int octaves, float lacunarity, float gain,
                            float period_units, uint32_t seed, float *kernel_ms);
void *clapca_tex3d_array(clapca_tex3d *t);
int clapca_tex3d_download(clapca_tex3d *t, uint8_t *host_rgba8);
int clapca_tex3d_destroy(clapca_tex3d *t);

/* device-resident field evaluation for benchmarks: results stay in device memory */
int clapca_noise_bake_device(void *d_out, size_t size, int octaves, float lacunarity, float gain,
                             float period_units, uint32_t seed, float *kernel_ms);
int clapca_terrain_heightmap_device(void *d_map, void *d_map0, long seed, unsigned nr_v, float ty,
                                    const void *d_maze, unsigned mside, float amp, int oct,
                                    float *map0_ms, float *map_ms);
/* the mesh buffers straight from a device-resident heightmap (all pointers are device memory; outputs may be NULL) */
int clapca_terrain_mesh_device(const void *d_map, unsigned nr_v, float x, float y, float z, float side,
                               void *d_vx, void *d_norm, void *d_tx, void *d_idx, float *kernel_ms);
/* the same from a device-resident maze and heightmap (d_out: device array of clapca_instor) */
int clapca_terrain_instantiators_device(const void *d_maze, unsigned mside, const uint32_t *nr_states, int nkinds,
                                        const void *d_map, unsigned nr_v, float x, float z, float side,
                                        void *d_out, size_t cap, size_t *count, float *kernel_ms);
void *clapca_device_alloc(size_t bytes);
int   clapca_device_free(void *p);
int   clapca_memcpy_h2d(void *dst, const void *src, size_t bytes);
int   clapca_memcpy_d2h(void *dst, const void *src, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* CLAPCA_H */
