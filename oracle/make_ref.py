"""Recipe for `oracle/_ref/`: a byte-for-byte COPY of the reference's pure-Python
path (environment/, agent/, configs/ -- about 1.3 k lines, no build step) taken
from /root/reference in the build container.  `oracle/_ref/` is git-ignored (the
reference's sources never enter this repository's history) but not
gpurun-ignored, so the unmodified reference travels to the GPU box, where
`bench.py --impl reference` and `bench.py`'s `cpu_baseline` leg time it on the
host cores (`"kind": "reference"`), and where tests may compare against it.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under self_play_racing_b200/
imports it.  gymnasium is absent from the image; `tools/gym_stub.py` supplies the
few gymnasium names the reference touches (see its header).

    python oracle/make_ref.py          # no-op when /root/reference is absent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference'
DST = os.path.join(HERE, '_ref')
PACKAGES = ('environment', 'agent', 'configs')


def make_ref(verbose=True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f'oracle/make_ref: {SRC} not present (GPU box?) -- keeping whatever is in {DST}')
        return os.path.isdir(os.path.join(DST, 'environment'))
    os.makedirs(DST, exist_ok=True)
    n = 0
    for pkg in PACKAGES:
        out = os.path.join(DST, pkg)
        os.makedirs(out, exist_ok=True)
        for name in sorted(os.listdir(os.path.join(SRC, pkg))):
            if name.endswith('.py'):
                shutil.copyfile(os.path.join(SRC, pkg, name), os.path.join(out, name))
                n += 1
    if verbose:
        print(f'oracle/make_ref: copied {n} files of the unmodified reference into {DST}')
    return True


def import_ref():
    """Put oracle/_ref (and the gymnasium stand-in) on sys.path; returns False when the copy is absent."""
    if not os.path.isdir(os.path.join(DST, 'environment')):
        return False
    root = os.path.dirname(HERE)
    tools = os.path.join(root, 'tools')
    if tools not in sys.path:
        sys.path.insert(0, tools)
    import gym_stub
    gym_stub.install()
    if DST not in sys.path:
        sys.path.insert(0, DST)
    sys.dont_write_bytecode = True
    return True


if __name__ == '__main__':
    make_ref()
