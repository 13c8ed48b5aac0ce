"""The C-ABI library loads and exports every symbol include/voxactb.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'voxactb.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(vxb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from voxactb_b200 import _lib, build
    build.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), 'libvoxactb.so does not export %s' % n
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_channel():
    from voxactb_b200 import _lib
    L = _lib.lib()
    assert L.vxb_version() == 100
    # argument validation happens before any device work: exercisable without a GPU
    rc = L.vxb_voxelize_f32(None, None, None, 1, 1, 1, 3, 4, None, 0, None, None, 0, None)
    assert rc == -1
    assert b'null pointer' in L.vxb_last_error()
    assert L.vxb_voxelize_workspace_bytes(16, 65536, 100, 3) > 0


def test_descriptor_validation_mirrors_reference_failure():
    """k=s=4 at V=32 gives 9^3 patches vs an 8^3 positional encoding: the reference raises at
    perceiver_lang_io.py:422; the library refuses the shape."""
    from voxactb_b200 import _lib, PerceiverVoxelLangEncoder
    enc = PerceiverVoxelLangEncoder(depth=1, iterations=1, voxel_size=32, initial_dim=10, low_dim_size=4,
                                    num_latents=32, activation='lrelu', voxel_patch_size=4,
                                    voxel_patch_stride=4)
    d = enc._desc()
    assert _lib.lib().vxb_qnet_workspace_bytes(ctypes.byref(d), 1) == 0
    assert b'positional encoding' in _lib.lib().vxb_last_error()
    good = PerceiverVoxelLangEncoder(depth=1, iterations=1, voxel_size=32, initial_dim=10, low_dim_size=4,
                                     num_latents=32, activation='lrelu', voxel_patch_size=5,
                                     voxel_patch_stride=4)
    assert _lib.lib().vxb_qnet_workspace_bytes(ctypes.byref(good._desc()), 1) > 0
    assert _lib.lib().vxb_qnet_num_params(ctypes.byref(good._desc())) == len(__import__("voxactb_b200.perceiver_lang_io", fromlist=["x"])._FIXED_SLOTS) + 12


def test_no_cpu_fallback():
    import torch
    from voxactb_b200 import VoxelGrid, PerceiverVoxelLangEncoder
    vg = VoxelGrid([-1, -1, -1, 1, 1, 1], 8, 'cpu', 1, 3, 16)
    with pytest.raises(RuntimeError, match='CUDA only'):
        vg.coords_to_bounding_voxel_grid(torch.zeros(1, 16, 3), torch.zeros(1, 16, 3))
    enc = PerceiverVoxelLangEncoder(depth=1, iterations=1, voxel_size=20, initial_dim=10, low_dim_size=4,
                                    num_latents=32, activation='lrelu', voxel_patch_size=5,
                                    voxel_patch_stride=5).eval()
    with pytest.raises(RuntimeError, match='CUDA only'):
        enc(torch.zeros(1, 10, 20, 20, 20), torch.zeros(1, 4), None, torch.zeros(1, 77, 512), None, None, None)


def test_install_shims_redirects_the_reference_import_sites():
    """INTEGRATION.md route 1: after install_shims() the import statements of the reference's agent modules
    (qattention_peract_bc_agent.py:17, launch_utils.py:21) resolve to this package's classes."""
    import subprocess
    import sys
    code = (
        "import voxactb_b200\n"
        "voxactb_b200.install_shims()\n"
        "from voxel.voxel_grid import VoxelGrid\n"
        "from agents.peract_bc.perceiver_lang_io import PerceiverVoxelLangEncoder, PerceiverVoxelLang2RobotsEncoder\n"
        "assert VoxelGrid is voxactb_b200.VoxelGrid\n"
        "assert PerceiverVoxelLangEncoder is voxactb_b200.PerceiverVoxelLangEncoder\n"
        "assert PerceiverVoxelLang2RobotsEncoder is voxactb_b200.PerceiverVoxelLang2RobotsEncoder\n"
        "print('ok')\n")
    r = subprocess.run([sys.executable, '-c', code], cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == 'ok', r.stderr[-1500:]
