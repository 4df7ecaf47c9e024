"""Installs the UNMODIFIED reference RAFT (its Python package `RAFT/core`) under baseline/_ref/ so that the
reference-on-B200 baseline (tools/reference_gpu.py, bench.py's `reference_gpu` key) can run on the GPU box, where
/root/reference does not exist.

    python baseline/install_ref.py        # authoring container only; a no-op when the sources are absent

baseline/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it travels with
the snapshot like the built .so files.  Nothing under sd_animation_optical_flow_b200/ imports it.  The reference's
compiled op (`alt_cuda_corr`) is built separately into oracle/_ref/ by oracle/build_ref.py.

A `pip install --target baseline/_ref /root/reference` is not possible: the reference has no setup.py / pyproject
for the repository (only RAFT/alt_cuda_corr/setup.py for the extension), it is a tree of scripts.
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('SDOF_REFERENCE', '/root/reference')
DST = os.path.join(HERE, '_ref')


def installed_core():
    p = os.path.join(DST, 'RAFT', 'core')
    return p if os.path.exists(os.path.join(p, 'raft.py')) else None


def install(force: bool = False):
    """Returns the installed RAFT/core path, or None when neither the sources nor a previous install exist."""
    have = installed_core()
    src_core = os.path.join(SRC, 'RAFT', 'core')
    if have and not force:
        return have
    if not os.path.exists(os.path.join(src_core, 'raft.py')):
        return have
    dst_core = os.path.join(DST, 'RAFT', 'core')
    if os.path.exists(dst_core):
        shutil.rmtree(dst_core)
    shutil.copytree(src_core, dst_core, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    return dst_core


if __name__ == '__main__':
    p = install(force=True)
    print(p if p else 'reference sources not present: nothing installed')
