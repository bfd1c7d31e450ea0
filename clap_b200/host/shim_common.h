/*
 * shim_common.h -- internals of the C host shim (libclapca_host): lazy
 * binding to the GPU library and the reference's "log and abort" error
 * convention (the CA API has no error channel: core/util.h:228-234 CHECK,
 * core/xyarray.c:13 .fatal_fail).
 */
#ifndef CLAPCA_SHIM_COMMON_H
#define CLAPCA_SHIM_COMMON_H

#include <stddef.h>
#include <stdint.h>

/* bind to the device named by $CLAPCA_DEVICE (default 0) on first use */
void shim_require_gpu(void);
/* print "clapca: <what>: <library error>" and abort() */
void shim_fatal(const char *what, int status) __attribute__((noreturn));
void *shim_alloc_zeroed(size_t bytes);
/*
 * The process-wide lrand48() / drand48() stream: seed48() hands back the 48-bit state it replaces, so reading it is
 * "swap in anything, copy the old value, swap it back".  (seed48() also restores the default multiplier; the
 * reference never calls lcong48().)
 */
uint64_t shim_rand48_peek(void);
void shim_rand48_poke(uint64_t state);

#endif
