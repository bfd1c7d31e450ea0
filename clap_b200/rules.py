"""Rule descriptors: the Python face of ``struct cell_automaton`` (core/ca-common.h:10-32)
and the reference's built-in rule tables (core/ca3d.c:110-122, core/terrain.c:391-415)."""
from dataclasses import dataclass

from ._lib import NEIGH_M1, NEIGH_MV


def _bits(*ns):
    m = 0
    for n in ns:
        m |= 1 << n
    return m


def ca_range(start, end):
    """CA_RANGE(start, end): neighbour counts start .. end-1 (core/ca3d.h:34)."""
    return ((1 << (end - start)) - 1) << start


@dataclass(frozen=True)
class CellAutomaton:
    name: str
    born_mask: int
    surv_mask: int
    nr_states: int
    decay: bool = False          # 2D only; the 3D sweep always decays (core/ca3d.c:136-137)
    neigh: int = NEIGH_M1        # 2D neighbourhood selector; 3D is always the 26-cell Moore count


# cas[]: core/ca3d.c:110-122, in enum order (core/ca3d.h:36-47)
CA3D_RULES = (
    CellAutomaton("ca_445m", born_mask=_bits(4), surv_mask=_bits(4), nr_states=5),
    CellAutomaton("ca_678_678_3m", born_mask=_bits(6, 7, 8), surv_mask=_bits(6, 7, 8), nr_states=3),
    CellAutomaton("ca_pyroclastic", born_mask=_bits(6, 7, 8), surv_mask=_bits(4, 5, 6, 7), nr_states=10),
    CellAutomaton("ca_amoeba", born_mask=_bits(5, 6, 7, 12, 13, 15), surv_mask=ca_range(9, 26), nr_states=5),
    CellAutomaton("ca_builder", born_mask=_bits(4, 6, 8, 9), surv_mask=_bits(2, 6, 9), nr_states=10),
    CellAutomaton("ca_slow_decay", born_mask=ca_range(13, 26),
                  surv_mask=_bits(1, 4, 8, 11) | ca_range(13, 26), nr_states=5),
    CellAutomaton("ca_spiky_growth", born_mask=_bits(4, 13, 17, 26) | ca_range(20, 24),
                  surv_mask=ca_range(0, 3) | ca_range(7, 9) | ca_range(11, 13) | _bits(18, 21, 22, 24, 26),
                  nr_states=4),
    CellAutomaton("ca_coral", born_mask=ca_range(6, 7) | _bits(9, 12), surv_mask=ca_range(5, 8), nr_states=4),
    CellAutomaton("ca_crystal_1", born_mask=_bits(1, 3), surv_mask=ca_range(0, 6), nr_states=2),
)


def ca3d_rule(nca):
    """Rule selected by ca3d_run(xyz, nca, ...): index nca % 9 computed in size_t (core/ca3d.c:126), so a negative
    index wraps through 2^64 first (-1 selects rule 6, as in the reference)."""
    return CA3D_RULES[(int(nca) & ((1 << 64) - 1)) % len(CA3D_RULES)]


# core/terrain.c:391-398 and :400-415
CA_TEST = CellAutomaton("test", born_mask=3 << 2, surv_mask=3 << 7, nr_states=4, decay=True, neigh=NEIGH_M1)
CA_INSTORS = (
    CellAutomaton("cool tree", born_mask=0x1e, surv_mask=0xff, nr_states=20, decay=False, neigh=NEIGH_MV),
    CellAutomaton("ash pinus", born_mask=0xffffff, surv_mask=0xffffff, nr_states=21, decay=False, neigh=NEIGH_MV),
)
