/*
 * emu_bitslice.cpp -- TEST ONLY: the bit-sliced building blocks of clap_b200/csrc/bitslice.cuh against plain integer
 * arithmetic, cell by cell: adders, the 3D neighbour-count tree, compile-time and run-time rule tables (the nine cas[]
 * entries of core/ca3d.c:110-122 and random masks), and the in-row dependency scan in both forms (D and E = ~D, split
 * into its common and its rare part) against the serial recurrence a'(x) = C(x) ^ (D(x) & a'(x-1)).
 * The warp-level pieces are covered by emu_ca3d / emu_ca2d; nothing here needs the fiber runtime.
 */
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>

#ifndef CLAPCA_EMU
#define CLAPCA_EMU
#endif
#include "../../clap_b200/csrc/bitslice.cuh"

namespace clapca {
/* the runtime hooks devport.h declares: unused by the word-level functions under test */
int emu_lane() { return 0; }
int emu_warp_in_block() { return 0; }
int emu_block() { return 0; }
int emu_grid_blocks() { return 1; }
int emu_block_threads() { return 32; }
uint32_t emu_exchange(uint32_t v, int) { return v; }
uint32_t emu_ballot(bool p) { return p ? 1u : 0u; }
void emu_yield() { }
long long emu_clock() { return 0; }
void emu_syncblock() { }
void emu_syncblock_named(int, int) { }
void *emu_block_shared(size_t) { return nullptr; }
}
using namespace clapca;

static uint64_t rs = 0x9E3779B97F4A7C15ull;
static uint32_t rnd()
{
    rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17;
    return (uint32_t)(rs >> 16);
}
static int bad;
#define CHECK(c, ...) do { if (!(c)) { if (bad < 20) { printf("FAIL %s:%d ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } bad++; } } while (0)

static unsigned bit(uint32_t w, int i) { return (w >> i) & 1u; }

template <uint32_t MASK>
static void check_tab5(const char *name)
{
    for (int rep = 0; rep < 64; rep++) {
        uint32_t k[5];
        for (int b = 0; b < 5; b++) k[b] = rnd();
        const uint32_t t = bs_tab5<MASK>(k), d = bs_tab_dyn(MASK, k, 5);
        for (int i = 0; i < 32; i++) {
            const unsigned K = bit(k[0], i) | bit(k[1], i) << 1 | bit(k[2], i) << 2 | bit(k[3], i) << 3 | bit(k[4], i) << 4;
            CHECK(bit(t, i) == ((MASK >> K) & 1u), "tab5 %s K=%u", name, K);
            CHECK(bit(d, i) == ((MASK >> K) & 1u), "tab_dyn %s K=%u", name, K);
        }
    }
}

#define RANGE(a, b) ((((1u << ((b) - (a) + 1)) - 1u)) << (a))
#define B(n) (1u << (n))

int main()
{
    /* adders */
    for (int rep = 0; rep < 2000; rep++) {
        uint32_t a[2] = { rnd(), rnd() }, b[2] = { rnd(), rnd() }, c[2] = { rnd(), rnd() }, t[3], v[4];
        bs_add2x2(a, b, t);
        bs_add3x2(a, b, c, v);
        for (int i = 0; i < 32; i++) {
            const unsigned va = bit(a[0], i) + 2 * bit(a[1], i), vb = bit(b[0], i) + 2 * bit(b[1], i), vc = bit(c[0], i) + 2 * bit(c[1], i);
            CHECK(bit(t[0], i) + 2 * bit(t[1], i) + 4 * bit(t[2], i) == va + vb, "add2x2");
            CHECK(bit(v[0], i) + 2 * bit(v[1], i) + 4 * bit(v[2], i) + 8 * bit(v[3], i) == va + vb + vc, "add3x2");
        }
    }
    /* the 3D count: K = T(y-1) + T(y) + T(y+1) + Hnew + Hold + r with T = Hdn + Hup (each H in 0..3), both trees */
    for (int rep = 0; rep < 4000; rep++) {
        uint32_t hd[3][2], hu[3][2], n[2] = { rnd(), rnd() }, o[2] = { rnd(), rnd() }, r = rnd(), t[3][3], k[5], k2[5], vd[4], vu[4];
        for (int q = 0; q < 3; q++) {
            for (int b = 0; b < 2; b++) { hd[q][b] = rnd(); hu[q][b] = rnd(); }
            if (rep & 1) { hd[q][0] = hd[q][1] = ~0u; hu[q][0] = hu[q][1] = ~0u; }     /* the maximum: 25 */
            bs_add2x2(hd[q], hu[q], t[q]);
        }
        if (rep & 1) { n[0] = n[1] = o[0] = o[1] = ~0u; r = ~0u; }
        bs_count3d_t(t[0], t[1], t[2], n, o, r, k);
        bs_add3x2(hd[0], hd[1], hd[2], vd);
        bs_add3x2(hu[0], hu[1], hu[2], vu);
        bs_count3d(vd, vu, n, o, r, k2);
        for (int i = 0; i < 32; i++) {
            unsigned want = bit(n[0], i) + 2 * bit(n[1], i) + bit(o[0], i) + 2 * bit(o[1], i) + bit(r, i);
            for (int q = 0; q < 3; q++)
                want += bit(hd[q][0], i) + 2 * bit(hd[q][1], i) + bit(hu[q][0], i) + 2 * bit(hu[q][1], i);
            unsigned got = 0, got2 = 0;
            for (int b = 0; b < 5; b++) { got |= bit(k[b], i) << b; got2 |= bit(k2[b], i) << b; }
            CHECK(got == want, "count3d_t %u != %u", got, want);
            CHECK(got2 == want, "count3d %u != %u", got2, want);
        }
    }
    /* rule tables: cas[] of core/ca3d.c:110-122 (and the shifted masks the kernels use for n = K + 1) */
    check_tab5<B(4)>("445m");
    check_tab5<B(6) | B(7) | B(8)>("678");
    check_tab5<RANGE(4, 7)>("pyro surv");
    check_tab5<RANGE(9, 26)>("amoeba surv");
    check_tab5<B(5) | B(6) | B(7) | B(12) | B(13) | B(15)>("amoeba born");
    check_tab5<B(2) | B(6) | B(9)>("builder surv");
    check_tab5<B(4) | B(6) | B(8) | B(9)>("builder born");
    check_tab5<B(1) | B(4) | B(8) | B(11) | RANGE(13, 26)>("slow decay surv");
    check_tab5<RANGE(0, 3) | RANGE(7, 9) | RANGE(11, 13) | B(18) | B(21) | B(22) | B(24) | B(26)>("spiky surv");
    check_tab5<B(4) | B(13) | B(17) | RANGE(20, 24) | B(26)>("spiky born");
    check_tab5<RANGE(5, 8)>("coral surv");
    check_tab5<(RANGE(5, 8) >> 1)>("coral surv + 1");
    check_tab5<RANGE(6, 7) | B(9) | B(12)>("coral born");
    check_tab5<((RANGE(6, 7) | B(9) | B(12)) >> 1)>("coral born + 1");
    check_tab5<RANGE(0, 6)>("crystal surv");
    check_tab5<B(1) | B(3)>("crystal born");
    check_tab5<0u>("zero");
    check_tab5<0x03ffffffu>("ones");
    /* the scan: serial recurrence vs D form vs E form (common part alone where it is complete, then the rare part) */
    for (int rep = 0; rep < 20000; rep++) {
        uint32_t D = rnd(), C = rnd();
        switch (rep % 5) {              /* densities: random, sparse D, dense D, all-propagate, long runs */
        case 1: D &= rnd() & rnd(); break;
        case 2: D |= rnd() | rnd(); break;
        case 3: D = ~0u; break;
        case 4: D = ~(1u << (rnd() & 31)) & ~(rnd() & rnd() & rnd() & rnd()); break;
        }
        for (unsigned cin = 0; cin < 2; cin++) {
            uint32_t want = 0, prev = cin;
            for (int i = 0; i < 32; i++) {
                prev = bit(C, i) ^ (bit(D, i) & prev);
                want |= prev << i;
            }
            uint32_t d1 = D, c1 = C;
            bs_scan_word(d1, c1);
            CHECK((c1 ^ (d1 & (0u - cin))) == want, "scan_word D=%08x C=%08x", D, C);
            uint32_t e2 = ~D, c2 = C;
            bs_scan_word_e(e2, c2);
            CHECK((c2 ^ (~e2 & (0u - cin))) == want, "scan_word_e D=%08x C=%08x", D, C);
            uint32_t e3 = ~D, c3 = C;
            bs_scan_word_e_lo(e3, c3);
            const bool more = (~e3 & 0xffffff00u) != 0u;
            if (!more)
                CHECK((c3 ^ (~e3 & (0u - cin))) == want, "scan_word_e_lo alone D=%08x C=%08x", D, C);
            bs_scan_word_e_hi(e3, c3);
            CHECK((c3 ^ (~e3 & (0u - cin))) == want, "scan_word_e lo+hi D=%08x C=%08x", D, C);
            CHECK(e3 == e2 && c3 == c2, "split scan differs from the whole one");
        }
    }
    if (bad) { printf("FAIL %d checks\n", bad); return 1; }
    printf("OK bitslice\n");
    return 0;
}
