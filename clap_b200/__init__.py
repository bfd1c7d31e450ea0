"""clap_b200 -- B200-native procedural-generation hot path of virtuoso/clap.

Python host layer over the C ABI (``include/clapca.h``): rule descriptors, one-shot calls
on numpy arrays that mirror the reference's C entry points, device-resident grids for
pipelines and benchmarks, and the multi-GPU slab driver.  All compute happens in the
hand-written sm_100a kernels of ``libclapca_cuda``; nothing here falls back to the CPU.
"""
from ._lib import (ClapcaError, ENGINE_AUTO, ENGINE_BITPLANE, ENGINE_DIAGONAL, ENGINE_WAVEFRONT, NEIGH_M1, NEIGH_MV, NEIGH_VN1,
                   NEIGH_VNV, init)
from .rules import CA3D_RULES, CA_INSTORS, CA_TEST, CellAutomaton, ca3d_rule
from .ca import Grid, ca2d_generate, ca2d_seed, ca2d_step, ca3d_run
from .fields import NoiseTexture3D, blue_noise2d, noise_fbm3, noise_grad3d_bake_rgba8, terrain_heightmap, terrain_instantiators, terrain_map0, terrain_mesh

__all__ = [
    "ClapcaError", "ENGINE_AUTO", "ENGINE_BITPLANE", "ENGINE_DIAGONAL", "ENGINE_WAVEFRONT", "NEIGH_M1", "NEIGH_MV", "NEIGH_VN1",
    "NEIGH_VNV", "init", "CA3D_RULES", "CA_INSTORS", "CA_TEST", "CellAutomaton", "ca3d_rule", "Grid",
    "ca2d_generate", "ca2d_seed", "ca2d_step", "ca3d_run", "NoiseTexture3D", "blue_noise2d", "noise_fbm3", "noise_grad3d_bake_rgba8",
    "terrain_heightmap", "terrain_instantiators", "terrain_map0", "terrain_mesh",
]
