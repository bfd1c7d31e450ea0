#!/usr/bin/env python
"""Coarse timing model of the sharded tile schedule (no GPU): which z-block size / tile shape / segment length keeps 8
ranks busy?  Every rank claims its tiles in key order with `ctas` CTAs; a tile advances at `rate` rows per microsecond
but never past what its producers allow (rows ahead of it by `lag` row steps, more across a GPU boundary).  Structural
effects only -- pipeline fill, the tail of the last items, load imbalance -- which is what bounded the 8-GPU run.

usage: sim_tiles.py [ranks] [block] [Tz] [Tg] [seg_rows]
"""
import sys
import numpy as np


def simulate(Z=2048, H=2048, G=50, R=8, B=40, Tz=5, Tg=3, ctas=148, rate=0.877, lag_local=19.0, lag_remote=29.0,
             dt=4.0, seg=0):
    # tile columns (never straddle a z-block), generation groups
    cols = []
    for j, z0 in enumerate(range(0, Z, B)):
        for c0 in range(z0, min(Z, z0 + B), Tz):
            cols.append((c0, min(Tz, min(Z, z0 + B) - c0), j % R))
    nb = -(-G // Tg)
    nseg = 1 if seg <= 0 else -(-H // seg)
    L = H if seg <= 0 else seg
    items = []      # (key, col index, b, s)
    for ci, (c0, n, r) in enumerate(cols):
        for b in range(nb):
            for s in range(nseg):
                items.append(((c0 + (Tz + 1) * b) * 4096 + s * 0 + b, ci, b, s))
    # claim order: key, but segments of a tile in order and spaced so that later segments come later: key + s * big
    # (a later segment only becomes useful ~L rows of sweep later: interleave by the row-time estimate)
    def tkey(it):
        key, ci, b, s = it
        c0 = cols[ci][0]
        return (3.7 * c0 + 22.0 * b + s * L, b)
    items.sort(key=tkey)
    n = len(items)
    col = np.array([it[1] for it in items]); grp = np.array([it[2] for it in items]); sg = np.array([it[3] for it in items])
    rank = np.array([cols[c][2] for c in col])
    y0 = sg * L
    y1 = np.minimum(H, y0 + L)
    index = {(items[i][1], items[i][2], items[i][3]): i for i in range(n)}
    ncols = len(cols)

    def find(ci, b, y):         # item holding row y of tile (ci, b), or -1 when that row does not exist
        if ci < 0 or ci >= ncols or b < 0:
            return -1
        s = 0 if seg <= 0 else min(nseg - 1, int(y // L))
        return index[(ci, b, s)]

    pos = y0.astype(float).copy()       # next row to do
    started = np.zeros(n, bool)
    done = np.zeros(n, bool)
    nextp = np.zeros(R, int)            # per-rank pointer into its list
    lists = [np.flatnonzero(rank == r) for r in range(R)]
    running = [[] for _ in range(R)]
    # producers of an item: (col-1, b) same segment rows; (col+1, b-1); (col, b-1); own previous segment
    prod = []
    for i in range(n):
        ci, b, s = col[i], grp[i], sg[i]
        p = []
        if ci > 0:
            p.append((ci - 1, b, lag_remote if cols[ci - 1][2] != cols[ci][2] else lag_local))
        if b > 0:
            p.append((ci, b - 1, lag_local))
            if ci + 1 < ncols:
                p.append((ci + 1, b - 1, lag_remote if cols[ci + 1][2] != cols[ci][2] else lag_local))
        prod.append(p)
    rowpos = {}
    t = 0.0
    busy = np.zeros(R)
    finish = np.zeros(R)
    # progress of tile (ci,b) in absolute rows = max over its segments' pos where started/done
    prog = np.zeros((ncols, nb))
    while not done.all():
        for r in range(R):
            run = running[r]
            while len(run) < ctas and nextp[r] < len(lists[r]):
                i = lists[r][nextp[r]]; nextp[r] += 1
                run.append(i); started[i] = True
            new = []
            for i in run:
                ci, b = col[i], grp[i]
                limit = y1[i]
                # own previous segment must be complete
                if sg[i] > 0 and prog[ci, b] < y0[i]:
                    limit = y0[i]
                for (pc, pb, lag) in prod[i]:
                    limit = min(limit, max(y0[i], prog[pc, pb] - lag) if prog[pc, pb] < H else y1[i])
                adv = min(rate * dt, max(0.0, limit - pos[i]))
                pos[i] += adv
                if adv > 0:
                    busy[r] += adv / rate
                if pos[i] >= y1[i] - 1e-9:
                    done[i] = True
                    finish[r] = t + dt
                else:
                    new.append(i)
            running[r] = new
        for r in range(R):
            pass
        # publish progress (after all ranks moved: one dt of latency, fine)
        for i in np.flatnonzero(started):
            ci, b = col[i], grp[i]
            if pos[i] > prog[ci, b] and (sg[i] == 0 or prog[ci, b] >= y0[i] - 1e-9):
                prog[ci, b] = pos[i]
        t += dt
        if t > 5e5:
            print("stuck"); break
    planes = [sum(c[1] for c in cols if c[2] == r) for r in range(R)]
    ideal = Z * G * H / rate / (R * ctas * Tz * Tg)
    return t, ideal, planes, busy / ctas


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    R = a[0] if len(a) > 0 else 8
    B = a[1] if len(a) > 1 else 40
    Tz = a[2] if len(a) > 2 else 5
    Tg = a[3] if len(a) > 3 else 3
    seg = a[4] if len(a) > 4 else 0
    t, ideal, planes, busy = simulate(R=R, B=B, Tz=Tz, Tg=Tg, seg=seg)
    print(f"R={R} B={B} tile {Tz}x{Tg} seg {seg}: {t / 1000:.2f} ms (ideal {ideal / 1000:.2f} ms, eff {ideal / t:.3f}); planes/rank max {max(planes)}")
