"""In-tree build of libvoxactb.so (hand-written sm_100a CUDA behind a C ABI).

``python -m voxactb_b200.build`` or ``voxactb_b200.build.build()``.  nvcc cross-compiles
without a GPU; the resulting .so sits next to this file so it travels with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libvoxactb.so')
SOURCES = ['voxelize.cu', 'qnet.cu', 'api_ops.cu', 'umma_ops.cu', 'train_ops.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, '..', 'include', 'voxactb.h'))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile_one(args):
    src, obj, verbose = args
    cmd = [_nvcc()] + NVCC_FLAGS[:-1] + ['-c', src, '-o', obj]
    if verbose:
        print(' '.join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    return res.returncode, res.stdout + res.stderr


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into voxactb_b200/libvoxactb.so (objects in parallel)."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = [(os.path.join(CSRC, s), os.path.join(objdir, s[:-3] + '.o'), verbose) for s in srcs]
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        results = list(ex.map(_compile_one, jobs))
    for rc, out in results:
        if rc != 0:
            raise RuntimeError('nvcc failed:\n' + out)
    cmd = [_nvcc(), '-shared', '-o', LIB] + [j[1] for j in jobs]
    if verbose:
        print(' '.join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc link failed:\n' + res.stdout + res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
