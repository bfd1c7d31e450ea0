/*
 * shim_common.h -- internals of the C host shim (libclapca_host): lazy
 * binding to the GPU library and the reference's "log and abort" error
 * convention (the CA API has no error channel: core/util.h:228-234 CHECK,
 * core/xyarray.c:13 .fatal_fail).
 */
#ifndef CLAPCA_SHIM_COMMON_H
#define CLAPCA_SHIM_COMMON_H

#include <stddef.h>

/* bind to the device named by $CLAPCA_DEVICE (default 0) on first use */
void shim_require_gpu(void);
/* print "clapca: <what>: <library error>" and abort() */
void shim_fatal(const char *what, int status) __attribute__((noreturn));
void *shim_alloc_zeroed(size_t bytes);

#endif
