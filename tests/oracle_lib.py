"""ctypes access to the oracle libraries for tests / smoke / bench baselines ONLY.

``port()``  -> oracle/liboracle_port.so (our C restatement, built on demand with gcc)
``ref()``   -> oracle/_ref/libclapref.so (unmodified reference sources) or None
Nothing under clap_b200/ may import this module.
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, byref, c_bool, c_char_p, c_float, c_int, c_int64, c_long, c_size_t, c_uint, c_uint32, \
    c_uint64, c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "liboracle_port.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libclapref.so")

NEIGH_VN1, NEIGH_M1, NEIGH_VNV, NEIGH_MV = range(4)


def _vp(a):
    return a.ctypes.data_as(c_void_p)


class Port:
    def __init__(self, lib):
        self.lib = lib
        lib.ora_lrand48.restype = c_long
        lib.ora_ca3d_run.restype = c_int64
        lib.ora_count.restype = c_int64
        lib.ora_hash31.restype = c_float
        lib.ora_fbm3_periodic.restype = c_float
        lib.ora_fbm3_periodic.argtypes = [c_float] * 3 + [c_int, c_float, c_float, c_int, c_uint32]
        lib.ora_get_rand_height.restype = c_float
        lib.ora_get_rand_height.argtypes = [c_long, c_int, c_int]
        lib.ora_noise_grad3d_bake_rgba8.argtypes = [c_void_p, c_size_t, c_size_t, c_size_t, c_int, c_float, c_float,
                                                    c_float, c_uint32]
        lib.ora_blue_noise2d.argtypes = [c_void_p, c_void_p]
        lib.ora_terrain_map0.argtypes = [c_long, c_uint, c_void_p]
        lib.ora_terrain_field.argtypes = [c_uint, c_void_p, c_float, c_float, c_int, c_uint, c_uint, c_void_p]
        lib.ora_terrain_heightmap.argtypes = [c_uint, c_void_p, c_float, c_void_p, c_uint, c_uint, c_uint, c_void_p]
        lib.ora_terrain_mesh.argtypes = [c_void_p, c_uint, c_float, c_float, c_float, c_float, c_uint, c_uint,
                                         c_void_p, c_void_p, c_void_p, c_void_p]
        lib.ora_terrain_height.restype = c_float
        lib.ora_terrain_height.argtypes = [c_void_p, c_uint, c_float, c_float, c_uint, c_float, c_float]
        lib.ora_terrain_instantiators.restype = c_size_t
        lib.ora_terrain_instantiators.argtypes = [c_void_p, c_uint, c_void_p, c_int, c_void_p, c_uint, c_float, c_float,
                                                  c_float, c_void_p, c_size_t]
        lib.ora_fnv1a64.restype = c_uint64
        lib.ora_fnv1a64.argtypes = [c_void_p, c_size_t]

    # --- rand48 ---
    def srand48(self, seed):
        st = c_uint64()
        self.lib.ora_srand48(byref(st), c_long(seed))
        return st

    def lrand48(self, st):
        return self.lib.ora_lrand48(byref(st))

    # --- CA ---
    def ca3d_rule(self, nca):
        s, b, n = c_uint(), c_uint(), c_uint()
        self.lib.ora_ca3d_rule(nca, byref(s), byref(b), byref(n))
        return s.value, b.value, n.value

    def ca3d_run(self, xyz, surv, born, nr_states, steps):
        """xyz[z, y, x] uint8, in place; returns population."""
        assert xyz.dtype == np.uint8 and xyz.flags.c_contiguous
        d2, d1, d0 = xyz.shape
        return self.lib.ora_ca3d_run(_vp(xyz), c_int64(d0), c_int64(d1), c_int64(d2), surv, born, nr_states, steps)

    def ca3d_make(self, d0, d1, d2, seed):
        out = np.zeros((d2, d1, d0), np.uint8)
        st = self.srand48(seed)
        self.lib.ora_ca3d_make(_vp(out), d0, d1, d2, byref(st))
        return out

    def ca2d_seed(self, side, nr_states, seed):
        out = np.zeros((side, side), np.uint8)
        st = self.srand48(seed)
        self.lib.ora_ca2d_seed(_vp(out), c_int64(side), nr_states, byref(st))
        return out

    def ca2d_run(self, arr, born, surv, nr_states, decay, neigh, steps, side=None):
        assert arr.dtype == np.uint8 and arr.flags.c_contiguous
        h, w = arr.shape
        side = w if side is None else side
        self.lib.ora_ca2d_run(_vp(arr), c_int64(w), c_int64(h), c_int64(side), born, surv, nr_states, int(decay),
                              neigh, steps)
        return arr

    # --- fields ---
    def fbm3(self, xyz, octaves, lac, gain, period, seed):
        return np.array([self.lib.ora_fbm3_periodic(float(p[0]), float(p[1]), float(p[2]), octaves, lac, gain,
                                                    period, seed) for p in xyz], dtype=np.float32)

    def noise_bake(self, size, octaves, lac, gain, period_units, seed, z0=0, z1=None):
        out = np.zeros((size, size, size, 4), np.uint8)
        self.lib.ora_noise_grad3d_bake_rgba8(_vp(out), size, z0, size if z1 is None else z1, octaves, lac, gain,
                                             period_units, seed)
        return out

    def blue_noise2d(self, state):
        """(pixels float32[64, 64, 4], rand48 state after the draws) from the 48-bit state `state`"""
        out = np.zeros((64, 64, 4), np.float32)
        st = np.array([state], np.uint64)
        self.lib.ora_blue_noise2d(_vp(out), _vp(st))
        return out, int(st[0])

    def terrain_map0(self, seed, nr_v):
        out = np.zeros((nr_v, nr_v), np.float32)
        self.lib.ora_terrain_map0(seed, nr_v, _vp(out))
        return out

    def terrain_field(self, map0, ty, amp, octv, i0=0, i1=None):
        nr_v = map0.shape[0]
        out = np.zeros((nr_v, nr_v), np.float32)
        self.lib.ora_terrain_field(nr_v, _vp(map0), ty, amp, octv, i0, nr_v if i1 is None else i1, _vp(out))
        return out

    def terrain_heightmap(self, map0, ty, maze, i0=0, i1=None):
        nr_v = map0.shape[0]
        out = np.zeros((nr_v, nr_v), np.float32)
        maze = np.ascontiguousarray(maze, np.uint8)
        self.lib.ora_terrain_heightmap(nr_v, _vp(map0), ty, _vp(maze), maze.shape[0], i0,
                                       nr_v if i1 is None else i1, _vp(out))
        return out

    def terrain_mesh(self, hmap, x, y, z, side):
        """(vx[n,3], norm[n,3], tx[n,2], idx[6*(nr_v-1)^2]) of core/terrain.c:479-516"""
        hmap = np.ascontiguousarray(hmap, np.float32)
        nr_v = hmap.shape[0]
        n = nr_v * nr_v
        vx, norm, tx = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 2), np.float32)
        idx = np.zeros(6 * (nr_v - 1) * (nr_v - 1), np.uint16)
        self.lib.ora_terrain_mesh(_vp(hmap), nr_v, x, y, z, side, 0, nr_v, _vp(vx), _vp(norm), _vp(tx), _vp(idx))
        return vx, norm, tx, idx

    def terrain_height(self, hmap, t_x, t_z, t_side, pts):
        hmap = np.ascontiguousarray(hmap, np.float32)
        return np.array([self.lib.ora_terrain_height(_vp(hmap), hmap.shape[0], t_x, t_z, int(t_side), float(x), float(z))
                         for x, z in pts], np.float32)

    def terrain_instantiators(self, maze, nr_states, hmap, x, z, side):
        """structured array (kind, dx, dy, dz) of core/terrain.c:555-570"""
        maze = np.ascontiguousarray(maze, np.uint8)
        hmap = np.ascontiguousarray(hmap, np.float32)
        kinds = np.array(nr_states, np.uint32)
        dt = np.dtype([("kind", np.int32), ("dx", np.float32), ("dy", np.float32), ("dz", np.float32)])
        args = (_vp(maze), maze.shape[0], _vp(kinds), len(kinds), _vp(hmap), hmap.shape[0], x, z, side)
        n = self.lib.ora_terrain_instantiators(*args, None, 0)
        out = np.zeros(n, dt)
        self.lib.ora_terrain_instantiators(*args, _vp(out), n)
        return out

    def fnv(self, arr):
        arr = np.ascontiguousarray(arr)
        return self.lib.ora_fnv1a64(_vp(arr), arr.nbytes)


class _CA(ctypes.Structure):     # struct cell_automaton, core/ca-common.h:10-32
    _fields_ = [("name", c_char_p), ("born", c_uint), ("surv", c_uint), ("nr", c_uint), ("decay", c_bool),
                ("neigh", c_void_p)]


class Ref:
    """The reference's own functions (oracle/_ref).  Grids are passed as numpy arrays with the
    12-byte xyzarray header in front, exactly like the reference's containers."""

    def __init__(self, lib):
        self.lib = lib
        self.libc = ctypes.CDLL(None)
        lib.ca3d_make.restype = c_void_p
        lib.ca3d_run.argtypes = [c_void_p, c_int, c_int]
        lib.ca2d_generate.restype = c_void_p
        lib.ca2d_step.argtypes = [c_void_p, c_void_p, c_int]
        lib.ref_fbm3_periodic.restype = c_float
        lib.ref_fbm3_periodic.argtypes = [c_float] * 3 + [c_int, c_float, c_float, c_int, c_uint32]
        lib.ref_hash31.restype = c_float
        lib.ref_noise_grad3d_bake_rgba8.restype = c_void_p
        lib.ref_noise_grad3d_bake_rgba8.argtypes = [c_size_t, c_int, c_float, c_float, c_float, c_uint32]
        lib.ref_mem_free.argtypes = [c_void_p]
        lib.ref_terrain_map0.argtypes = [c_long, c_uint, c_void_p]
        lib.ref_terrain_field.argtypes = [c_long, c_uint, c_void_p, c_float, c_float, c_int, c_uint, c_uint, c_void_p]
        lib.ref_terrain_heightmap.argtypes = [c_long, c_uint, c_void_p, c_float, c_void_p, c_uint, c_uint, c_void_p]
        self._neigh = [lib.ca2d_neigh_vn1, lib.ca2d_neigh_m1, lib.ca2d_neigh_vnv, lib.ca2d_neigh_mv]

    @staticmethod
    def _boxed(cells):
        """Copy `cells` ([z,y,x] or [y,x]) behind an xyzarray header; returns (buffer, payload view)."""
        c3 = cells if cells.ndim == 3 else cells[None]
        d2, d1, d0 = c3.shape
        buf = np.zeros(12 + c3.size, np.uint8)
        buf[:12].view(np.int32)[:] = (d0, d1, d2)
        buf[12:] = c3.ravel()
        return buf, buf[12:]

    def ca3d_make(self, d0, d1, d2, seed):
        self.libc.srand48(c_long(seed))
        p = self.lib.ca3d_make(d0, d1, d2)
        arr = np.ctypeslib.as_array(ctypes.cast(p + 12, POINTER(ctypes.c_ubyte)), shape=(d2, d1, d0)).copy()
        self.lib.ref_mem_free(p)
        return arr

    def ca3d_run(self, xyz, nca, steps):
        buf, payload = self._boxed(xyz)
        pop = self.lib.ca3d_run(buf.ctypes.data, nca, steps)
        xyz[...] = payload.reshape(xyz.shape)
        return pop

    def _ca(self, born, surv, nr, decay, neigh):
        return _CA(b"t", born, surv, nr, bool(decay), ctypes.cast(self._neigh[neigh], c_void_p))

    def ca2d_generate(self, born, surv, nr, decay, neigh, side, steps, seed):
        self.libc.srand48(c_long(seed))
        ca = self._ca(born, surv, nr, decay, neigh)
        p = self.lib.ca2d_generate(byref(ca), side, steps)
        arr = np.ctypeslib.as_array(ctypes.cast(p, POINTER(ctypes.c_ubyte)), shape=(side, side)).copy()
        self.lib.ref_mem_free(p - 12)
        return arr

    def ca2d_step(self, arr, born, surv, nr, decay, neigh, side=None, steps=1):
        buf, payload = self._boxed(arr)
        ca = self._ca(born, surv, nr, decay, neigh)
        for _ in range(steps):
            self.lib.ca2d_step(byref(ca), buf.ctypes.data + 12, arr.shape[1] if side is None else side)
        arr[...] = payload.reshape(arr.shape)
        return arr

    def fbm3(self, xyz, octaves, lac, gain, period, seed):
        return np.array([self.lib.ref_fbm3_periodic(float(p[0]), float(p[1]), float(p[2]), octaves, lac, gain,
                                                    period, seed) for p in xyz], dtype=np.float32)

    def noise_bake(self, size, octaves, lac, gain, period_units, seed):
        p = self.lib.ref_noise_grad3d_bake_rgba8(size, octaves, lac, gain, period_units, seed)
        arr = np.ctypeslib.as_array(ctypes.cast(p, POINTER(ctypes.c_ubyte)), shape=(size, size, size, 4)).copy()
        self.lib.ref_mem_free(p)
        return arr

    def terrain_map0(self, seed, nr_v):
        out = np.zeros((nr_v, nr_v), np.float32)
        self.lib.ref_terrain_map0(seed, nr_v, _vp(out))
        return out

    def terrain_field(self, seed, map0, ty, amp, octv):
        nr_v = map0.shape[0]
        out = np.zeros((nr_v, nr_v), np.float32)
        self.lib.ref_terrain_field(seed, nr_v, _vp(map0), ty, amp, octv, 0, nr_v, _vp(out))
        return out

    def terrain_height(self, hmap, t_x, t_z, t_side, pts):
        """the reference's terrain_height() (terrain.c:336-379) at the points pts[n, 2] = (x, z)"""
        hmap = np.ascontiguousarray(hmap, np.float32)
        self.lib.ref_terrain_height.restype = c_float
        self.lib.ref_terrain_height.argtypes = [c_void_p, c_uint, c_float, c_float, c_uint, c_float, c_float]
        return np.array([self.lib.ref_terrain_height(_vp(hmap), hmap.shape[0], t_x, t_z, int(t_side), float(x), float(z))
                         for x, z in pts], np.float32)

    def terrain_normals(self, hmap):
        """the reference's calc_normal() (terrain.c:93-110) for every vertex, in mesh order"""
        hmap = np.ascontiguousarray(hmap, np.float32)
        nr_v = hmap.shape[0]
        out = np.zeros((nr_v * nr_v, 3), np.float32)
        self.lib.ref_terrain_normals(c_uint(nr_v), _vp(hmap), _vp(out))
        return out

    def terrain_heightmap(self, seed, map0, ty, maze):
        nr_v = map0.shape[0]
        out = np.zeros((nr_v, nr_v), np.float32)
        buf, _ = self._boxed(np.ascontiguousarray(maze, np.uint8))
        self.lib.ref_terrain_heightmap(seed, nr_v, _vp(map0), ty, buf.ctypes.data + 12, 0, nr_v, _vp(out))
        return out


_port = None
_ref = False


def build_port():
    src = os.path.join(ORACLE_DIR, "port", "oracle_port.c")
    if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR, "port"], check=True, capture_output=True)
    return PORT_SO


def port():
    global _port
    if _port is None:
        _port = Port(ctypes.CDLL(build_port()))
    return _port


def ref():
    global _ref
    if _ref is False:
        _ref = Ref(ctypes.CDLL(REF_SO)) if os.path.exists(REF_SO) else None
    return _ref
