/*
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build): link-time stand-ins for the
 * engine symbols that the reference's hot-path objects pull in but never call
 * on the CA / noise-bake / heightmap path (message bus subscription from
 * logger.c, texture upload + kissfft from noise.c's out-of-scope wrappers).
 * Nothing here computes anything on the path.
 */
#include <stddef.h>
#define MODNAME "oracle_stub"
#include "error.h"

struct clap_context;
cerr subscribe(struct clap_context *ctx, int type, void *fn, void *data) { return CERR_OK; }

/* noise.c: texture wrappers (render.h) -- never reached by noise_grad3d_bake_rgba8() */
cerr _texture_init(void *tex, const void *opts) { return CERR_OK; }
cerr texture_load(void *tex, int format, unsigned int w, unsigned int h, void *buf) { return CERR_OK; }
void texture_deinit(void *tex) {}
/* noise.c: blue_noise2d_tex() only (out of scope) */
void *kiss_fft_alloc(int nfft, int inverse, void *mem, size_t *lenmem) { return NULL; }
void kiss_fft(void *cfg, const void *in, void *out) {}

/*
 * terrain.c: the mesh/entity/physics half of terrain_init_square_landscape()
 * (out of scope) references engine refclasses and scene/model/pipeline calls.
 * They are never reached from the heightmap glue; dummy storage / aborting
 * bodies only satisfy the dynamic loader.
 */
#include <stdlib.h>
char ref_class_entity3d[512], ref_class_mesh[512], ref_class_model3d[512], ref_class_model3dtx[512];
#define ORACLE_UNREACHABLE(_name) void _name(void) { abort(); }
ORACLE_UNREACHABLE(_ref_drop)
ORACLE_UNREACHABLE(clap_get_phys)
ORACLE_UNREACHABLE(clap_get_pipeline)
ORACLE_UNREACHABLE(entity3d_add_physics)
ORACLE_UNREACHABLE(entity3d_reset)
ORACLE_UNREACHABLE(entity3d_set)
ORACLE_UNREACHABLE(model3d_ref)
ORACLE_UNREACHABLE(pipeline_shader_find_get)
ORACLE_UNREACHABLE(ref_class_add)
ORACLE_UNREACHABLE(ref_class_unuse)
ORACLE_UNREACHABLE(scene_add_model)
ORACLE_UNREACHABLE(shader_prog_ref)
