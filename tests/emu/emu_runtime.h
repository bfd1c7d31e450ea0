/*
 * emu_runtime.h -- host-side warp emulator for the CA kernels (TEST ONLY).
 *
 * Runs kernel source compiled with -DCLAPCA_EMU: each warp is a ring of 32
 * ucontext fibers scheduled round-robin on one OS thread; warp collectives
 * (shuffle, ballot, syncwarp) are rendez-vous points of the ring; all warps of
 * a launch run concurrently on their own OS threads, which is what the
 * persistent dataflow kernels need.
 */
#ifndef CLAPCA_EMU_RUNTIME_H
#define CLAPCA_EMU_RUNTIME_H
#include <functional>

namespace clapca {
/* run `body` once per emulated thread of a <<<blocks, threads>>> launch */
void emu_launch(int blocks, int threads, const std::function<void()> &body);
}
#endif
