"""voxactb_b200 -- B200-native (sm_100a) voxel-policy hot path of VoxAct-B / PerAct.

Public surface mirrors the reference's classes for this path:
  VoxelGrid                      (reference peract/voxel/voxel_grid.py)
  PerceiverVoxelLangEncoder      (reference peract/agents/peract_bc/perceiver_lang_io.py)
  QFunction                      (reference peract/agents/peract_bc/qattention_peract_bc_agent.py:31-135)
All arithmetic runs in libvoxactb.so (hand-written CUDA behind the C ABI of include/voxactb.h).
"""
from ._lib import MATH_F16X3, MATH_F16F8C, MATH_FP32_SIMT, LIB_PATH, lib  # noqa: F401
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
    import importlib
    import sys
    import types
    from . import perceiver_lang_io as _p, voxel_grid as _v
    # keep the reference's real `voxel` package (voxel.augmentation is imported by the agent,
    # qattention_peract_bc_agent.py:18) and replace only its voxel_grid submodule
    pkg = sys.modules.get('voxel')
    if pkg is None:
        try:
            pkg = importlib.import_module('voxel')
        except ImportError:
            pkg = types.ModuleType('voxel')
            pkg.__path__ = []          # a package without other submodules
            sys.modules['voxel'] = pkg
    pkg.voxel_grid = _v
    sys.modules['voxel.voxel_grid'] = _v
    sys.modules['agents.peract_bc.perceiver_lang_io'] = _p
    # the SE(3) augmentation entry points the agent imports by name (qattention_peract_bc_agent.py:18): the real module keeps
    # everything else it defines
    from . import augmentation as _a
    try:
        aug = importlib.import_module('voxel.augmentation')
        aug.perturb_se3 = _a.perturb_se3
        aug.apply_se3_augmentation = _a.apply_se3_augmentation
        aug.apply_se3_augmentation_2Robots = _a.apply_se3_augmentation_2Robots
    except ImportError:                 # reference tree (or pytorch3d) absent: this package's module stands in
        sys.modules['voxel.augmentation'] = _a
        pkg.augmentation = _a
    agents = sys.modules.get('agents.peract_bc')
    if agents is not None:
        agents.perceiver_lang_io = _p
