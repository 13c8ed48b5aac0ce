#!/usr/bin/env python
"""Copy the handful of UNMODIFIED reference files the agent-level drop-in test drives into baseline/_ref/ (git-ignored,
but shipped to the GPU box with the repo snapshot -- /root/reference does not exist there).  Authoring container only:

    python tools/vendor_reference.py

Nothing under baseline/_ref is product code or committed; tests/test_agent_dropin.py imports these files with this
package's classes shimmed in (voxactb_b200.install_shims), which is the "drops into train.py / eval.py unchanged" proof at
the level SURVEY.md section 9 describes (the hydra launcher and the simulator are not installable here)."""
import os
import shutil

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, 'baseline', '_ref')
FILES = [
    'peract/agents/peract_bc/qattention_peract_bc_agent.py',
    'peract/agents/peract_bc/qattention_stack_agent.py',
    'peract/agents/peract_bc/perceiver_lang_io.py',
    'peract/helpers/__init__.py',
    'peract/helpers/utils.py',
    'peract/helpers/network_utils.py',
    'peract/helpers/preprocess_agent.py',
    'peract/helpers/optim/lamb.py',
    'peract/voxel/__init__.py',
    'peract/voxel/voxel_grid.py',
    'peract/voxel/augmentation.py',
    'YARR/yarr/__init__.py',
    'YARR/yarr/agents/__init__.py',
    'YARR/yarr/agents/agent.py',
    'PyRep/pyrep/objects/vision_sensor.py',
]


def main():
    n = 0
    for f in FILES:
        src = os.path.join(REF, f)
        if not os.path.exists(src):
            print('missing in the reference tree (skipped):', f)
            continue
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    for d in ('peract/helpers/optim',):
        init = os.path.join(DST, d, '__init__.py')
        if not os.path.exists(init) and os.path.exists(os.path.join(REF, d, '__init__.py')):
            shutil.copyfile(os.path.join(REF, d, '__init__.py'), init)
    print('vendored %d reference files into %s' % (n, DST))


if __name__ == '__main__':
    main()
