"""Field evaluation: the noise.c fBm-gradient bake and the terrain.c heightmap chain."""
from ctypes import POINTER, byref, c_size_t, c_uint32, c_void_p

import numpy as np

from . import _lib
from ._lib import check


def noise_grad3d_bake_rgba8(size, octaves=4, lacunarity=2.0, gain=0.5, period_units=64.0, seed=0xC14D):
    """noise_grad3d_bake_rgba8(): core/noise.c:222-270 (defaults of noise3d_make, :309-317).
    Returns uint8[size, size, size, 4] indexed [z, y, x, rgba]."""
    lib = _lib.lib()
    out = np.empty((size, size, size, 4), dtype=np.uint8)
    check(lib, lib.clapca_noise_grad3d_bake_rgba8(out.ctypes.data_as(c_void_p), size, octaves, lacunarity, gain,
                                                  period_units, seed))
    return out


class NoiseTexture3D:
    """The bake as a device-resident 3D RGBA8 texture (clapca_noise_bake_array): a CUDA array written
    through a surface -- what noise_grad3d_bake_rgba8_tex() (core/noise.c:272-294) hands the renderer, minus the
    host round trip.  `array` is the cudaArray_t to register with the graphics API; download() reads it back."""

    def __init__(self, size, octaves=4, lacunarity=2.0, gain=0.5, period_units=64.0, seed=0xC14D):
        from ctypes import c_float
        self._lib = _lib.lib()
        self.size = int(size)
        h, ms = c_void_p(), c_float()
        check(self._lib, self._lib.clapca_noise_bake_array(byref(h), size, octaves, lacunarity, gain, period_units, seed,
                                                           byref(ms)))
        self._h, self.kernel_ms = h, ms.value

    @property
    def array(self):
        return self._lib.clapca_tex3d_array(self._h)

    def download(self):
        out = np.empty((self.size, self.size, self.size, 4), dtype=np.uint8)
        check(self._lib, self._lib.clapca_tex3d_download(self._h, out.ctypes.data_as(c_void_p)))
        return out

    def close(self):
        if self._h:
            self._lib.clapca_tex3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def blue_noise2d(rng, size=64):
    """blue_noise2d_tex(): core/noise.c:96-169 up to the upload -- the (size, size, 4) float32 pixels of the film-grain
    texture.  `rng` is the caller's rand48 stream (clap_b200.ca.Rand48); it is advanced by the 3 * size^2 draws."""
    from ctypes import c_uint64
    lib = _lib.lib()
    out = np.empty((size, size, 4), np.float32)
    after = c_uint64(0)
    check(lib, lib.clapca_noise_blue2d_rgba32f(out.ctypes.data_as(c_void_p), size, rng.x, byref(after)))
    rng.x = int(after.value)
    return out


def noise_fbm3(xyz, octaves, lacunarity, gain, period, seed):
    """fbm3_periodic(): core/noise.c:204-220 at the points xyz[n, 3] (float32)."""
    lib = _lib.lib()
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    out = np.empty(xyz.shape[0], dtype=np.float32)
    check(lib, lib.clapca_noise_fbm3(out.ctypes.data_as(c_void_p), xyz.ctypes.data_as(c_void_p), xyz.shape[0],
                                     octaves, lacunarity, gain, period, seed))
    return out


def terrain_map0(seed, nr_v):
    """Lattice t->map0 of terrain_init_square_landscape(): core/terrain.c:447-450 (float32[nr_v, nr_v],
    indexed [x, z])."""
    lib = _lib.lib()
    out = np.empty((nr_v, nr_v), dtype=np.float32)
    check(lib, lib.clapca_terrain_map0(out.ctypes.data_as(c_void_p), seed, nr_v))
    return out


def terrain_heightmap(seed, nr_v, y=0.0, maze=None, amp=1.0, octaves=4):
    """Heightmap t->map: core/terrain.c:451-467.  With ``maze`` (uint8[mside, mside], the ca2d cave grid,
    mside = nr_v // 8) the maze-modulated map; without it the plain field get_height(i, j, amp, octaves)."""
    lib = _lib.lib()
    out = np.empty((nr_v, nr_v), dtype=np.float32)
    if maze is not None:
        maze = np.ascontiguousarray(maze, dtype=np.uint8)
        if maze.ndim != 2 or maze.shape[0] != maze.shape[1]:
            raise ValueError("maze must be a square 2-D uint8 array")
        mptr, mside = maze.ctypes.data_as(c_void_p), maze.shape[0]
    else:
        mptr, mside = None, 0
    check(lib, lib.clapca_terrain_heightmap(out.ctypes.data_as(c_void_p), seed, nr_v, y, mptr, mside, amp, octaves))
    return out


def terrain_mesh(hmap, x=0.0, y=0.0, z=0.0, side=1.0, indices=True):
    """Mesh buffers of terrain_init_square_landscape(): core/terrain.c:479-516 with calc_normal() (:93-110).
    ``hmap`` is t->map (float32[nr_v, nr_v]).  Returns (vx[n, 3], norm[n, 3], tx[n, 2], idx) with n = nr_v^2 and
    idx = uint16[6 * (nr_v - 1)^2] (None when ``indices`` is false) -- truncated to 16 bits like the reference's."""
    lib = _lib.lib()
    hmap = np.ascontiguousarray(hmap, dtype=np.float32)
    if hmap.ndim != 2 or hmap.shape[0] != hmap.shape[1]:
        raise ValueError("hmap must be a square 2-D float32 array")
    nr_v = hmap.shape[0]
    n = nr_v * nr_v
    vx, norm = np.empty((n, 3), dtype=np.float32), np.empty((n, 3), dtype=np.float32)
    tx = np.empty((n, 2), dtype=np.float32)
    idx = np.empty(6 * (nr_v - 1) * (nr_v - 1), dtype=np.uint16) if indices else None
    check(lib, lib.clapca_terrain_mesh(hmap.ctypes.data_as(c_void_p), nr_v, x, y, z, side, vx.ctypes.data_as(c_void_p),
                                       norm.ctypes.data_as(c_void_p), tx.ctypes.data_as(c_void_p),
                                       idx.ctypes.data_as(c_void_p) if indices else None))
    return vx, norm, tx, idx


INSTOR_DTYPE = np.dtype([("kind", np.int32), ("dx", np.float32), ("dy", np.float32), ("dz", np.float32)])


def terrain_instantiators(maze, nr_states, hmap, x=0.0, z=0.0, side=1.0):
    """Instantiator extraction of terrain_init_square_landscape(): core/terrain.c:555-570 with terrain_height()
    (:336-379).  ``maze`` is the cave grid after the ca_instors passes, ``nr_states`` the values that spawn an
    instantiator (one per kind, e.g. (20, 21) for ca_instors[]), ``hmap`` is t->map.  Returns a structured array
    (kind, dx, dy, dz) in the reference's list order."""
    lib = _lib.lib()
    maze = np.ascontiguousarray(maze, dtype=np.uint8)
    hmap = np.ascontiguousarray(hmap, dtype=np.float32)
    kinds = (c_uint32 * len(nr_states))(*[int(v) for v in nr_states])
    n = c_size_t(0)
    args = (maze.ctypes.data_as(c_void_p), maze.shape[0], kinds, len(nr_states), hmap.ctypes.data_as(c_void_p),
            hmap.shape[0], x, z, side)
    check(lib, lib.clapca_terrain_instantiators(*args, None, 0, byref(n)))
    out = np.zeros(n.value, dtype=INSTOR_DTYPE)
    if n.value:
        check(lib, lib.clapca_terrain_instantiators(*args, out.ctypes.data_as(c_void_p), n.value, byref(n)))
    return out
