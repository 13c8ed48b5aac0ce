"""voxactb_b200 -- B200-native (sm_100a) voxel-policy hot path of VoxAct-B / PerAct.

Public surface mirrors the reference's classes for this path:
  VoxelGrid                      (reference peract/voxel/voxel_grid.py)
  PerceiverVoxelLangEncoder      (reference peract/agents/peract_bc/perceiver_lang_io.py)
  QFunction                      (reference peract/agents/peract_bc/qattention_peract_bc_agent.py:31-135)
All arithmetic runs in libvoxactb.so (hand-written CUDA behind the C ABI of include/voxactb.h).
"""
from ._lib import MATH_BF16X3, MATH_FP32_SIMT, LIB_PATH, lib  # noqa: F401
from .perceiver_lang_io import PerceiverVoxelLangEncoder, PerceiverVoxelLang2RobotsEncoder  # noqa: F401
from .qfunction import QFunction, QFunction2Robots  # noqa: F401
from .voxel_grid import VoxelGrid  # noqa: F401

__all__ = ['VoxelGrid', 'PerceiverVoxelLangEncoder', 'PerceiverVoxelLang2RobotsEncoder', 'QFunction',
           'QFunction2Robots', 'lib', 'install_shims']


def install_shims():
    """Make the reference's import sites resolve to this package:
    ``from voxel.voxel_grid import VoxelGrid`` and
    ``from agents.peract_bc.perceiver_lang_io import PerceiverVoxelLangEncoder``.
    Call before importing the reference's agent modules."""
    import sys
    import types
    from . import perceiver_lang_io as _p, voxel_grid as _v
    pkg = sys.modules.setdefault('voxel', types.ModuleType('voxel'))
    pkg.voxel_grid = _v
    sys.modules['voxel.voxel_grid'] = _v
    sys.modules['agents.peract_bc.perceiver_lang_io'] = _p
