"""Import the reference's own modules (authoring container only: /root/reference is not on the GPU box)."""
import os
import sys
import types

REF = '/root/reference/peract'


def available():
    return os.path.isdir(REF)


def load():
    """Returns (RefVoxelGrid, RefPerceiverVoxelLangEncoder) from the unmodified reference tree.
    agents/peract_bc/__init__.py imports rlbench; empty package stubs keep that from running
    (SURVEY.md section 8c)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ('agents', 'agents.peract_bc'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF, name.replace('.', '/'))]
            sys.modules[name] = m
    from voxel.voxel_grid import VoxelGrid
    from agents.peract_bc.perceiver_lang_io import PerceiverVoxelLangEncoder
    return VoxelGrid, PerceiverVoxelLangEncoder


def load2():
    """PerceiverVoxelLang2RobotsEncoder from the unmodified reference tree."""
    load()
    from agents.peract_bc.perceiver_lang_io import PerceiverVoxelLang2RobotsEncoder
    return PerceiverVoxelLang2RobotsEncoder
