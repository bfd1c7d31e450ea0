/* emu_runtime.cpp -- see emu_runtime.h (TEST ONLY) */
#include <ucontext.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <atomic>
#include <thread>
#include <vector>
#include "emu_runtime.h"

namespace clapca {

namespace {

constexpr int kLanes = 32;
constexpr size_t kStack = 256 * 1024;

struct NamedBarrier {
    std::atomic<int> arrived{0};
    std::atomic<unsigned> generation{0};
};

struct Block {
    std::atomic<int> arrived{0};
    std::atomic<unsigned> generation{0};
    NamedBarrier named[16];
    int warps = 0;
    std::vector<char> smem;
};

struct Warp {
    Block *blk;
    ucontext_t main_ctx;
    ucontext_t ctx[kLanes];
    std::vector<char> stacks;
    bool done[kLanes];
    int live;
    int cur;
    uint32_t buf[2][kLanes];
    unsigned seq[kLanes];
    int block, warp_in_block, grid_blocks, block_threads;
    const std::function<void()> *body;
};

thread_local Warp *tw = nullptr;

void switch_to_next(Warp *w)
{
    int from = w->cur;
    for (int i = 1; i <= kLanes; i++) {
        int n = (from + i) % kLanes;
        if (!w->done[n]) {
            if (n == from)
                return;
            w->cur = n;
            swapcontext(&w->ctx[from], &w->ctx[n]);
            return;
        }
    }
}

void lane_entry()
{
    Warp *w = tw;
    (*w->body)();
    int me = w->cur;
    w->done[me] = true;
    w->live--;
    if (w->live == 0) {
        setcontext(&w->main_ctx);
    } else {
        for (int i = 1; i <= kLanes; i++) {
            int n = (me + i) % kLanes;
            if (!w->done[n]) {
                w->cur = n;
                setcontext(&w->ctx[n]);
            }
        }
    }
    abort();
}

void run_warp(Warp *w)
{
    tw = w;
    w->stacks.resize(kStack * kLanes);
    for (int l = 0; l < kLanes; l++) {
        w->done[l] = false;
        w->seq[l] = 0;
        getcontext(&w->ctx[l]);
        w->ctx[l].uc_stack.ss_sp = w->stacks.data() + kStack * l;
        w->ctx[l].uc_stack.ss_size = kStack;
        w->ctx[l].uc_link = nullptr;
        makecontext(&w->ctx[l], (void (*)())lane_entry, 0);
    }
    w->live = kLanes;
    w->cur = 0;
    swapcontext(&w->main_ctx, &w->ctx[0]);
    tw = nullptr;
}

/* wait until every lane of the warp has deposited collective number s */
void rendezvous(Warp *w, unsigned s)
{
    unsigned idle = 0;
    for (;;) {
        bool all = true;
        for (int l = 0; l < kLanes; l++) {
            if ((int)(w->seq[l] - s) <= 0) {
                if (w->done[l]) {
                    fprintf(stderr, "emu: lane %d exited before collective %u of its warp\n", l, s);
                    abort();
                }
                all = false;
                break;
            }
        }
        if (all)
            return;
        switch_to_next(w);
        if (++idle > 64) { sched_yield(); idle = 0; }
    }
}

} // namespace

int emu_lane()          { return tw->cur; }
int emu_warp_in_block() { return tw->warp_in_block; }
int emu_block()         { return tw->block; }
int emu_grid_blocks()   { return tw->grid_blocks; }
int emu_block_threads() { return tw->block_threads; }

uint32_t emu_exchange(uint32_t v, int src_lane)
{
    Warp *w = tw;
    int l = w->cur;
    unsigned s = w->seq[l];
    w->buf[s & 1][l] = v;
    w->seq[l] = s + 1;
    rendezvous(w, s);
    return w->buf[s & 1][src_lane];
}

uint32_t emu_ballot(bool p)
{
    Warp *w = tw;
    int l = w->cur;
    unsigned s = w->seq[l];
    w->buf[s & 1][l] = p ? 1u : 0u;
    w->seq[l] = s + 1;
    rendezvous(w, s);
    uint32_t r = 0;
    for (int i = 0; i < kLanes; i++)
        r |= (w->buf[s & 1][i] & 1u) << i;
    return r;
}

void emu_yield()
{
    /* a lane that polls memory: let the other fibers and the other warps run */
    switch_to_next(tw);
    sched_yield();
}

void emu_syncblock()
{
    Warp *w = tw;
    (void)emu_ballot(true);                 /* the whole warp has arrived */
    if (w->cur == 0) {
        Block *b = w->blk;
        unsigned g = b->generation.load(std::memory_order_acquire);
        if (b->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == b->warps) {
            b->arrived.store(0, std::memory_order_relaxed);
            b->generation.fetch_add(1, std::memory_order_acq_rel);
        } else {
            while (b->generation.load(std::memory_order_acquire) == g)
                sched_yield();
        }
    }
    (void)emu_ballot(true);                 /* lane 0 is through: release the other lanes */
}

/* bar.sync id, nthreads: the first nthreads / 32 arriving warps of the block form the barrier */
void emu_syncblock_named(int id, int nthreads)
{
    Warp *w = tw;
    (void)emu_ballot(true);
    if (w->cur == 0) {
        NamedBarrier *b = &w->blk->named[id & 15];
        const int nwarps = nthreads / kLanes;
        unsigned g = b->generation.load(std::memory_order_acquire);
        if (b->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == nwarps) {
            b->arrived.store(0, std::memory_order_relaxed);
            b->generation.fetch_add(1, std::memory_order_acq_rel);
        } else {
            while (b->generation.load(std::memory_order_acquire) == g)
                sched_yield();
        }
    }
    (void)emu_ballot(true);
}

void *emu_block_shared(size_t bytes)
{
    Block *b = tw->blk;
    if (b->smem.size() < bytes) {
        fprintf(stderr, "emu: block shared memory request %zu exceeds the launch's %zu bytes\n", bytes, b->smem.size());
        abort();
    }
    return b->smem.data();
}

long long emu_clock()
{
    using namespace std::chrono;
    return duration_cast<nanoseconds>(steady_clock::now().time_since_epoch()).count();
}

void emu_launch(int blocks, int threads, const std::function<void()> &body)
{
    if (threads % kLanes) {
        fprintf(stderr, "emu: block size must be a multiple of 32\n");
        abort();
    }
    int wpb = threads / kLanes;
    std::vector<Warp *> warps;
    std::vector<Block> blks(blocks);
    for (int b = 0; b < blocks; b++) {
        blks[b].warps = wpb;
        blks[b].smem.assign(64 * 1024, 0);
    }
    for (int b = 0; b < blocks; b++)
        for (int wi = 0; wi < wpb; wi++) {
            Warp *w = new Warp();
            w->blk = &blks[b];
            w->block = b;
            w->warp_in_block = wi;
            w->grid_blocks = blocks;
            w->block_threads = threads;
            w->body = &body;
            warps.push_back(w);
        }
    std::vector<std::thread> ths;
    for (Warp *w : warps)
        ths.emplace_back(run_warp, w);
    for (auto &t : ths)
        t.join();
    for (Warp *w : warps)
        delete w;
}

} // namespace clapca
