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
    # by file path under private names: voxactb_b200.install_shims() (tests/test_agent_dropin.py) rebinds the public module
    # names `voxel.voxel_grid` / `agents.peract_bc.perceiver_lang_io` to this package for the rest of the process
    return _by_path('voxel/voxel_grid.py').VoxelGrid, _by_path('agents/peract_bc/perceiver_lang_io.py').PerceiverVoxelLangEncoder


_LOADED = {}


def _by_path(rel):
    import importlib.util
    if rel not in _LOADED:
        spec = importlib.util.spec_from_file_location('_vxb_ref_' + rel.replace('/', '_')[:-3], os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _LOADED[rel] = mod
    return _LOADED[rel]


def load2():
    """PerceiverVoxelLang2RobotsEncoder from the unmodified reference tree."""
    load()
    return _by_path('agents/peract_bc/perceiver_lang_io.py').PerceiverVoxelLang2RobotsEncoder
