"""Synthetic seed volume of BASELINE config 4: a libc-free, position-keyed generator so that ANY
sub-volume -- a rank's z-blocks, a thin slab for the 64-bit-index oracle, the whole 2048^3 on one GPU -- holds
exactly the cells a single sequential fill would give it, in numpy (oracle side) and in torch (on the device).

    i     = (z * d1 + y) * d0 + x                              linear index in the reference layout (xyarray.c:43)
    s     = ((i + 1) * 0x9E3779B97F4A7C15) ^ (seed * 0xD1342543DE82EF95)   (mod 2^64; 0 -> 1)
    s    ^= s >> 12;  s ^= s << 25;  s ^= s >> 27;  r = s * 0x2545F4914F6CDD1D                   (one xorshift64* step)
    cell  = 1 + ((r >> 32) & 0xFFFFFF) % 5   if (r >> 62) == 0   else 0       P(alive) = 1/4, values uniform 1..5

Test / bench infrastructure: nothing in the product path imports this module.
"""

SEED = 0xC1A9
_GOLD = 0x9E3779B97F4A7C15
_SMIX = 0xD1342543DE82EF95
_STAR = 0x2545F4914F6CDD1D
_M64 = (1 << 64) - 1


def _i64(v):
    """two's-complement int64 view of a uint64 constant (torch has no uint64 arithmetic)"""
    v &= _M64
    return v - (1 << 64) if v >> 63 else v


def synth_numpy(np, d0, d1, z0, z1, seed=SEED):
    """planes z0..z1-1 of the d0 x d1 x * volume as a (z1-z0, d1, d0) uint8 array"""
    out = np.empty((z1 - z0, d1, d0), np.uint8)
    plane = d0 * d1
    k = np.arange(plane, dtype=np.uint64)
    smix = np.uint64((seed * _SMIX) & _M64)
    with np.errstate(over="ignore"):
        for z in range(z0, z1):
            s = (k + np.uint64(z * plane + 1)) * np.uint64(_GOLD) ^ smix
            s[s == 0] = 1
            s ^= s >> np.uint64(12)
            s ^= s << np.uint64(25)
            s ^= s >> np.uint64(27)
            r = s * np.uint64(_STAR)
            val = (np.uint64(1) + ((r >> np.uint64(32)) & np.uint64(0xFFFFFF)) % np.uint64(5)).astype(np.uint8)
            out[z - z0] = np.where((r >> np.uint64(62)) == 0, val, 0).reshape(d1, d0)
    return out


def synth_torch(torch, d0, d1, z0, z1, device, seed=SEED, chunk_planes=8):
    """the same cells generated on `device` (int64 arithmetic wraps mod 2^64; right shifts are masked to be logical)"""
    out = torch.empty((z1 - z0, d1, d0), dtype=torch.uint8, device=device)
    plane = d0 * d1
    smix = _i64(seed * _SMIX)
    for c0 in range(z0, z1, chunk_planes):
        c1 = min(z1, c0 + chunk_planes)
        i = torch.arange(c0 * plane + 1, c1 * plane + 1, dtype=torch.int64, device=device)
        s = (i * _i64(_GOLD)) ^ smix
        s = torch.where(s == 0, torch.ones_like(s), s)
        s = s ^ ((s >> 12) & ((1 << 52) - 1))
        s = s ^ (s << 25)
        s = s ^ ((s >> 27) & ((1 << 37) - 1))
        r = s * _i64(_STAR)
        val = 1 + ((r >> 32) & 0xFFFFFF) % 5
        alive = ((r >> 62) & 3) == 0
        out[c0 - z0:c1 - z0] = torch.where(alive, val, torch.zeros_like(val)).to(torch.uint8).reshape(c1 - c0, d1, d0)
        del i, s, r, val, alive
    return out


def plane_hashes_numpy(np, vol):
    """numpy twin of clapca_hash_planes(): one 64-bit fingerprint per plane of a (planes, d1, d0) uint8 array"""
    vol = np.ascontiguousarray(vol, np.uint8)
    n = vol.shape[0]
    flat = vol.reshape(n, -1)
    pad = (-flat.shape[1]) % 8
    if pad:
        flat = np.concatenate([flat, np.zeros((n, pad), np.uint8)], axis=1)
    w = flat.view("<u8")
    k = np.arange(1, w.shape[1] + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (w ^ (k * np.uint64(_GOLD))) + np.uint64(_GOLD)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
        return x.sum(axis=1, dtype=np.uint64)
