"""ctypes binding of libclapca_cuda / libclapca_host (the C ABI in include/clapca.h).

The libraries are built in-tree (``clap_b200/lib/``) by ``__graft_entry__.build()`` or
``make -C clap_b200/csrc``.  There is no fallback: if the CUDA library is missing or no
GPU is usable, importing the binding works but the first call raises ``ClapcaError``.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_float, c_int, c_int64, c_long, c_size_t,
                    c_uint, c_uint32, c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
# CLAPCA_LIB_DIR: another in-tree build of the same sources (A/B measurements of compile-time variants)
LIB_DIR = os.environ.get("CLAPCA_LIB_DIR") or os.path.join(_HERE, "lib")
CUDA_LIB = os.path.join(LIB_DIR, "libclapca_cuda.so")
HOST_LIB = os.path.join(LIB_DIR, "libclapca_host.so")

OK, ERR_CUDA, ERR_ARG, ERR_NOMEM, ERR_TIMEOUT, ERR_UNSUPPORTED, ERR_STATE = range(7)
NEIGH_VN1, NEIGH_M1, NEIGH_VNV, NEIGH_MV = range(4)
ENGINE_AUTO, ENGINE_WAVEFRONT, ENGINE_BITPLANE, ENGINE_DIAGONAL = range(4)
ENGINE_NAMES = {ENGINE_AUTO: "auto", ENGINE_WAVEFRONT: "wavefront", ENGINE_BITPLANE: "bitplane",
                ENGINE_DIAGONAL: "diagonal"}


class ClapcaError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"clapca status {status}: {message}")
        self.status = status


class RunStats(Structure):
    _fields_ = [("total_ms", c_float), ("kernel_ms", c_float), ("launches", c_int), ("engine", c_int),
                ("planes", c_int), ("workers", c_int), ("streamed", c_int)]


# every symbol include/clapca.h declares: (restype, argtypes)
SIGNATURES = {
    "clapca_device_count": (c_int, []),
    "clapca_init": (c_int, [c_int]),
    "clapca_shutdown": (None, []),
    "clapca_last_error": (c_char_p, []),
    "clapca_sm_count": (c_int, []),
    "clapca_device_mem_bytes": (c_size_t, []),
    "clapca_ca3d_run": (c_int, [c_void_p, POINTER(c_int64), c_uint32, c_uint32, c_uint32, c_int, c_int,
                                POINTER(c_int64)]),
    "clapca_ca3d_rule": (c_int, [c_int, POINTER(c_uint32), POINTER(c_uint32), POINTER(c_uint32)]),
    "clapca_ca2d_run": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_uint32, c_uint32, c_uint32, c_int, c_int,
                                c_int, c_int]),
    "clapca_ca2d_generate": (c_int, [c_void_p, c_int64, c_uint32, c_uint32, c_uint32, c_int, c_int, c_int, c_int,
                                     c_uint64, POINTER(c_uint64)]),
    "clapca_grid_seed2d": (c_int, [c_void_p, c_int64, c_uint32, c_uint64, POINTER(c_uint64)]),
    "clapca_noise_grad3d_bake_rgba8": (c_int, [c_void_p, c_size_t, c_int, c_float, c_float, c_float, c_uint32]),
    "clapca_noise_blue2d_rgba32f": (c_int, [c_void_p, c_int, c_uint64, POINTER(c_uint64)]),
    "clapca_noise_blue2d_device": (c_int, [c_void_p, c_int, c_uint64, POINTER(c_uint64), POINTER(c_float)]),
    "clapca_noise_fbm3": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_float, c_float, c_int, c_uint32]),
    "clapca_terrain_map0": (c_int, [c_void_p, c_long, c_uint]),
    "clapca_terrain_heightmap": (c_int, [c_void_p, c_long, c_uint, c_float, c_void_p, c_uint, c_float, c_int]),
    "clapca_grid_create": (c_int, [POINTER(c_void_p), c_int64, c_int64, c_int64]),
    "clapca_grid_destroy": (c_int, [c_void_p]),
    "clapca_grid_upload": (c_int, [c_void_p, c_void_p]),
    "clapca_grid_download": (c_int, [c_void_p, c_void_p]),
    "clapca_grid_device_ptr": (c_void_p, [c_void_p]),
    "clapca_grid_stream": (c_void_p, [c_void_p]),
    "clapca_grid_make3d": (c_int, [c_void_p, c_uint64, POINTER(c_uint64), POINTER(c_int64)]),
    "clapca_grid_run3d": (c_int, [c_void_p, c_uint32, c_uint32, c_uint32, c_int, c_int, POINTER(c_int64)]),
    "clapca_grid_run3d_streamed": (c_int, [c_void_p, c_void_p, c_void_p, c_uint, c_uint32, c_uint32, c_uint32, c_int,
                                           POINTER(c_int64)]),
    "clapca_grid_run2d": (c_int, [c_void_p, c_int64, c_uint32, c_uint32, c_uint32, c_int, c_int, c_int, c_int]),
    "clapca_grid_count": (c_int, [c_void_p, POINTER(c_int64)]),
    "clapca_grid_last_stats": (c_int, [c_void_p, POINTER(RunStats)]),
    "clapca_slab_create": (c_int, [POINTER(c_void_p), c_int64, c_int64, c_int64, c_int, c_int, c_int, c_int, c_uint]),
    "clapca_slab_destroy": (c_int, [c_void_p]),
    "clapca_slab_local_planes": (c_int, [c_void_p, POINTER(c_int)]),
    "clapca_slab_plane_map": (c_int, [c_void_p, POINTER(c_int64)]),
    "clapca_slab_device_ptr": (c_void_p, [c_void_p]),
    "clapca_slab_ipc_handle": (c_int, [c_void_p, c_void_p]),
    "clapca_slab_connect": (c_int, [c_void_p, c_void_p, c_void_p]),
    "clapca_slab_prepare_streamed": (c_int, [c_void_p, c_uint32, c_uint32, c_uint32, c_int]),
    "clapca_slab_run_streamed": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_int64)]),
    "clapca_slab_halo_ptr": (c_void_p, [c_void_p]),
    "clapca_slab_connect_local": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "clapca_noise_bake_array": (c_int, [POINTER(c_void_p), c_size_t, c_int, c_float, c_float, c_float, c_uint32, POINTER(c_float)]),
    "clapca_tex3d_array": (c_void_p, [c_void_p]),
    "clapca_tex3d_download": (c_int, [c_void_p, c_void_p]),
    "clapca_tex3d_destroy": (c_int, [c_void_p]),
    "clapca_hash_planes": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p]),
    "clapca_slab_upload": (c_int, [c_void_p, c_void_p]),
    "clapca_slab_download": (c_int, [c_void_p, c_void_p]),
    "clapca_slab_prepare": (c_int, [c_void_p, c_uint32, c_uint32, c_uint32, c_int]),
    "clapca_slab_run": (c_int, [c_void_p, POINTER(c_int64)]),
    "clapca_slab_last_stats": (c_int, [c_void_p, POINTER(RunStats)]),
    "clapca_noise_bake_device": (c_int, [c_void_p, c_size_t, c_int, c_float, c_float, c_float, c_uint32,
                                         POINTER(c_float)]),
    "clapca_terrain_heightmap_device": (c_int, [c_void_p, c_void_p, c_long, c_uint, c_float, c_void_p, c_uint,
                                                c_float, c_int, POINTER(c_float), POINTER(c_float)]),
    "clapca_terrain_mesh": (c_int, [c_void_p, c_uint, c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "clapca_terrain_mesh_device": (c_int, [c_void_p, c_uint, c_float, c_float, c_float, c_float, c_void_p, c_void_p,
                                           c_void_p, c_void_p, POINTER(c_float)]),
    "clapca_terrain_instantiators": (c_int, [c_void_p, c_uint, POINTER(c_uint32), c_int, c_void_p, c_uint, c_float, c_float,
                                             c_float, c_void_p, c_size_t, POINTER(c_size_t)]),
    "clapca_terrain_instantiators_device": (c_int, [c_void_p, c_uint, POINTER(c_uint32), c_int, c_void_p, c_uint, c_float,
                                                    c_float, c_float, c_void_p, c_size_t, POINTER(c_size_t),
                                                    POINTER(c_float)]),
    "clapca_device_alloc": (c_void_p, [c_size_t]),
    "clapca_device_free": (c_int, [c_void_p]),
    "clapca_memcpy_h2d": (c_int, [c_void_p, c_void_p, c_size_t]),
    "clapca_memcpy_d2h": (c_int, [c_void_p, c_void_p, c_size_t]),
}

_cuda = None


def load_cuda_library():
    """dlopen libclapca_cuda.so and attach the prototypes.  Raises if it was never built."""
    global _cuda
    if _cuda is not None:
        return _cuda
    if not os.path.exists(CUDA_LIB):
        raise ClapcaError(ERR_STATE, f"{CUDA_LIB} is missing: run __graft_entry__.build() / "
                                     "make -C clap_b200/csrc (there is no CPU fallback)")
    lib = ctypes.CDLL(CUDA_LIB, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _cuda = lib
    return lib


def check(lib, status):
    if status != OK:
        msg = lib.clapca_last_error()
        raise ClapcaError(status, msg.decode() if msg else "?")


_bound_device = None


def init(device=None):
    """Bind this process to one GPU (default: $LOCAL_RANK, else 0)."""
    global _bound_device
    lib = load_cuda_library()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", os.environ.get("CLAPCA_DEVICE", "0")))
    if _bound_device != device:
        check(lib, lib.clapca_init(int(device)))
        _bound_device = device
    return lib


def lib():
    return init(_bound_device) if _bound_device is not None else init()
