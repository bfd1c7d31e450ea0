"""Cellular-automaton entry points over numpy arrays and device-resident grids.

Function names and argument meaning follow the reference's C API
(core/ca2d.h:8-13, core/ca3d.h:49-52); arrays use the reference layout
(``arr[z, y, x]`` C-contiguous uint8 == index z*d0*d1 + y*d0 + x, core/xyarray.c:43).
"""
import ctypes
from ctypes import byref, c_int64, c_void_p

import numpy as np

from . import _lib
from ._lib import ENGINE_AUTO, RunStats, check
from .rules import CellAutomaton, ca3d_rule

_R48_A, _R48_C, _R48_MASK = 0x5DEECE66D, 0xB, (1 << 48) - 1


def _as_u8(arr, ndim):
    if not isinstance(arr, np.ndarray) or arr.dtype != np.uint8 or arr.ndim != ndim:
        raise TypeError(f"expected a {ndim}-D uint8 numpy array")
    if not arr.flags.c_contiguous or not arr.flags.writeable:
        raise ValueError("array must be C-contiguous and writeable (results are written in place)")
    return arr


def ca3d_run(xyz, nca, steps, engine=ENGINE_AUTO):
    """ca3d_run(): core/ca3d.c:124-142.  ``xyz[z, y, x]`` is updated in place; ``nca`` is a rule
    index (mod 9) or a :class:`CellAutomaton`.  Returns the population (xyzarray_count)."""
    lib = _lib.lib()
    xyz = _as_u8(xyz, 3)
    rule = nca if isinstance(nca, CellAutomaton) else ca3d_rule(int(nca))
    dim = (c_int64 * 3)(xyz.shape[2], xyz.shape[1], xyz.shape[0])
    pop = c_int64(0)
    check(lib, lib.clapca_ca3d_run(xyz.ctypes.data_as(c_void_p), dim, rule.surv_mask, rule.born_mask,
                                   rule.nr_states, int(steps), engine, byref(pop)))
    return pop.value


def ca2d_step(ca, arr, side=None, steps=1, engine=ENGINE_AUTO):
    """ca2d_step(): core/ca2d.c:61-77, ``steps`` times.  ``arr[y, x]`` (index y*w + x) is updated in
    place; cells x, y < side are swept x-outer / y-inner (side defaults to the array width)."""
    lib = _lib.lib()
    arr = _as_u8(arr, 2)
    h, w = arr.shape
    side = w if side is None else int(side)
    check(lib, lib.clapca_ca2d_run(arr.ctypes.data_as(c_void_p), w, h, side, ca.born_mask, ca.surv_mask,
                                   ca.nr_states, int(ca.decay), ca.neigh, int(steps), engine))
    return arr


class Rand48:
    """glibc srand48()/lrand48() stream (X' = a X + c mod 2^48, result X' >> 17), so that host-side
    seeding reproduces what a C caller of the reference gets after ``srand48(seed)``."""

    def __init__(self, seed):
        self.x = ((int(seed) & 0xFFFFFFFF) << 16) | 0x330E

    def lrand48(self):
        self.x = (self.x * _R48_A + _R48_C) & _R48_MASK
        return self.x >> 17

    def lrand48_block(self, n):
        """n consecutive lrand48() values as a numpy array (vectorised LCG jump-ahead)."""
        out = np.empty(n, dtype=np.uint64)
        # powers of the affine map x -> a x + c, applied in blocks
        block = 1 << 12
        a_pows = np.empty(block, dtype=object)
        c_pows = np.empty(block, dtype=object)
        a, c = 1, 0
        for i in range(block):
            a, c = (a * _R48_A) & _R48_MASK, (c * _R48_A + _R48_C) & _R48_MASK
            a_pows[i], c_pows[i] = a, c
        a_lo = np.array([int(v) & 0xFFFFFF for v in a_pows], dtype=np.uint64)
        a_hi = np.array([int(v) >> 24 for v in a_pows], dtype=np.uint64)
        c_arr = np.array([int(v) for v in c_pows], dtype=np.uint64)
        mask = np.uint64(_R48_MASK)
        pos = 0
        while pos < n:
            m = min(block, n - pos)
            x = self.x
            x_lo, x_hi = np.uint64(x & 0xFFFFFF), np.uint64(x >> 24)
            # (a * x) mod 2^48 with 24-bit limbs: lo*lo + ((lo*hi + hi*lo) << 24)
            prod = a_lo[:m] * x_lo + (((a_lo[:m] * x_hi + a_hi[:m] * x_lo) & np.uint64(0xFFFFFF)) << np.uint64(24))
            vals = (prod + c_arr[:m]) & mask
            out[pos:pos + m] = vals >> np.uint64(17)
            self.x = int(vals[m - 1])
            pos += m
        return out


def ca2d_seed(ca, side, rng):
    """Seeding loop of ca2d_generate(): core/ca2d.c:86-90 -- x outer, y inner, one lrand48() % 8 per
    cell, cell = nr_states if v <= nr_states else 0.  ``rng`` is a :class:`Rand48`."""
    v = (rng.lrand48_block(side * side) % np.uint64(8)).astype(np.int64).reshape(side, side)   # [x, y]
    cells = np.where(v <= ca.nr_states, ca.nr_states & 0xFF, 0).astype(np.uint8)
    return np.ascontiguousarray(cells.T)                                                        # [y, x]


def ca2d_generate(ca, side, steps, rng, engine=ENGINE_AUTO):
    """ca2d_generate(): core/ca2d.c:79-98 with an explicit rand48 stream instead of libc's global.  Seeding and
    the generations both run on the device (clapca_ca2d_generate); ``rng`` is left where side*side lrand48()
    calls would have left it.  (:func:`ca2d_seed` is the same seeding loop on the host.)"""
    lib = _lib.lib()
    arr = np.empty((side, side), dtype=np.uint8)
    after = ctypes.c_uint64(0)
    check(lib, lib.clapca_ca2d_generate(arr.ctypes.data_as(c_void_p), side, ca.born_mask, ca.surv_mask, ca.nr_states,
                                        int(ca.decay), ca.neigh, int(steps), engine, rng.x, byref(after)))
    rng.x = int(after.value)
    return arr


def hash_planes(device_ptr, plane_bytes, nplanes):
    """clapca_hash_planes(): one 64-bit fingerprint per plane of a uint8 volume in device memory (numpy uint64)."""
    lib = _lib.lib()
    out = np.zeros(max(1, int(nplanes)), np.uint64)
    if nplanes:
        check(lib, lib.clapca_hash_planes(c_void_p(int(device_ptr)), int(plane_bytes), int(nplanes),
                                          out.ctypes.data_as(c_void_p)))
    return out[:int(nplanes)]


class Grid:
    """A uint8 grid resident in device memory (clapca_grid_*): upload once, run many times."""

    def __init__(self, d0, d1, d2=1):
        self._lib = _lib.lib()
        self.dims = (int(d0), int(d1), int(d2))
        h = c_void_p()
        check(self._lib, self._lib.clapca_grid_create(byref(h), *self.dims))
        self._h = h

    @property
    def cells(self):
        return self.dims[0] * self.dims[1] * self.dims[2]

    def close(self):
        if self._h:
            self._lib.clapca_grid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, host):
        """host: numpy uint8 array or a raw address (e.g. a pinned torch tensor's data_ptr())."""
        ptr = host.ctypes.data if isinstance(host, np.ndarray) else int(host)
        check(self._lib, self._lib.clapca_grid_upload(self._h, c_void_p(ptr)))

    def download(self, host):
        ptr = host.ctypes.data if isinstance(host, np.ndarray) else int(host)
        check(self._lib, self._lib.clapca_grid_download(self._h, c_void_p(ptr)))

    def device_ptr(self):
        return self._lib.clapca_grid_device_ptr(self._h)

    def stream(self):
        return self._lib.clapca_grid_stream(self._h)

    def run3d(self, rule, steps, engine=ENGINE_AUTO):
        rule = rule if isinstance(rule, CellAutomaton) else ca3d_rule(int(rule))
        pop = c_int64(0)
        check(self._lib, self._lib.clapca_grid_run3d(self._h, rule.surv_mask, rule.born_mask, rule.nr_states,
                                                     int(steps), engine, byref(pop)))
        return pop.value

    def make3d(self, rng):
        """clapca_grid_make3d(): ca3d_make() (core/ca3d.c:144-169) into this device-resident grid -- faces, the random
        walk (host, on a sparse picture of the volume) and the prune (255 marks included).  `rng`: the caller's Rand48
        stream, advanced by the walk's draws.  Returns the population."""
        from ctypes import c_uint64
        pop, after = c_int64(0), c_uint64(0)
        check(self._lib, self._lib.clapca_grid_make3d(self._h, rng.x, byref(after), byref(pop)))
        rng.x = int(after.value)
        return pop.value

    def run3d_streamed(self, rule, steps, host_in, host_out, max_value=255):
        """clapca_grid_run3d_streamed(): host -> device -> host as one pipeline (upload, all generations and download
        overlap inside one sweep launch when the buffers are page-locked).  host_in / host_out: numpy uint8 arrays
        or raw addresses (pinned torch tensors' data_ptr()); max_value bounds the input cells (255 is always safe)."""
        rule = rule if isinstance(rule, CellAutomaton) else ca3d_rule(int(rule))
        pin = host_in.ctypes.data if isinstance(host_in, np.ndarray) else int(host_in)
        pout = host_out.ctypes.data if isinstance(host_out, np.ndarray) else int(host_out)
        pop = c_int64(0)
        check(self._lib, self._lib.clapca_grid_run3d_streamed(self._h, c_void_p(pin), c_void_p(pout), int(max_value),
                                                              rule.surv_mask, rule.born_mask, rule.nr_states,
                                                              int(steps), byref(pop)))
        return pop.value

    def seed2d(self, ca, rng, side=None):
        """clapca_grid_seed2d(): the seeding loop of ca2d_generate() into this device-resident grid."""
        side = self.dims[0] if side is None else int(side)
        after = ctypes.c_uint64(0)
        check(self._lib, self._lib.clapca_grid_seed2d(self._h, side, ca.nr_states, rng.x, byref(after)))
        rng.x = int(after.value)

    def run2d(self, ca, steps, side=None, engine=ENGINE_AUTO):
        side = self.dims[0] if side is None else int(side)
        check(self._lib, self._lib.clapca_grid_run2d(self._h, side, ca.born_mask, ca.surv_mask, ca.nr_states,
                                                     int(ca.decay), ca.neigh, int(steps), engine))

    def count(self):
        pop = c_int64(0)
        check(self._lib, self._lib.clapca_grid_count(self._h, byref(pop)))
        return pop.value

    def stats(self):
        st = RunStats()
        check(self._lib, self._lib.clapca_grid_last_stats(self._h, byref(st)))
        return {"total_ms": st.total_ms, "kernel_ms": st.kernel_ms, "launches": st.launches,
                "engine": _lib.ENGINE_NAMES.get(st.engine, str(st.engine)), "planes": st.planes,
                "workers": st.workers, "streamed": bool(st.streamed)}
