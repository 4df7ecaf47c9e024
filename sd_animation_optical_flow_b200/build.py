"""Build recipe for libsdof_b200.so (the C-ABI library, include/sdof_b200.h).

Plain nvcc, sm_100a only, no torch headers: `python -m sd_animation_optical_flow_b200.build`.
The library is built IN-TREE (sd_animation_optical_flow_b200/lib/) so it travels with the
repository snapshot to the GPU box; it is git-ignored.  cudart is linked statically, so the
library loads on a machine without a driver (symbol checks on CPU) and binds to libcuda lazily.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_DIR = os.path.join(PKG_DIR, 'lib')
OBJ_DIR = os.path.join(PKG_DIR, 'build')
LIB_PATH = os.path.join(LIB_DIR, 'libsdof_b200.so')

SOURCES = ['sdof_core.cu', 'warp.cu', 'mask.cu', 'fused.cu', 'corr_simt.cu', 'corr_lookup.cu', 'corr_tc.cu', 'corr_tc_res.cu', 'raft_glue.cu', 'blur.cu', 'keyframe.cu', 'resize.cu', 'conv_tc.cu', 'raft_glue16.cu']

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17',
    '--expt-relaxed-constexpr',
    '-Xcompiler', '-fPIC,-ffp-contract=off',
    '-Xptxas', '-v',
]
# experiments: SDOF_NVCC_EXTRA="-DSDOF_RES_TRACE" python -m sd_animation_optical_flow_b200.build   (tools/corr_trace.py)
NVCC_FLAGS += os.environ.get('SDOF_NVCC_EXTRA', '').split()


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: set NVCC or put /usr/local/cuda/bin on PATH')


def _stamp(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, 'rb') as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link libsdof_b200.so.  Returns its path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sources()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(PKG_DIR, '..', 'include', 'sdof_b200.h'))
    stamp = _stamp(srcs + headers)
    stamp_file = os.path.join(OBJ_DIR, 'stamp.txt')
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB_PATH
    nvcc = _nvcc()
    logs = {}

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src) + '.o')
        cmd = [nvcc, *NVCC_FLAGS, '-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[os.path.basename(src)] = r.stderr
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    link = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-cudart', 'static',
            '-Xcompiler', '-fPIC', *objs, '-o', LIB_PATH]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    with open(os.path.join(OBJ_DIR, 'ptxas.log'), 'w') as f:
        for k in sorted(logs):
            f.write(f'==== {k}\n{logs[k]}\n')
    with open(stamp_file, 'w') as f:
        f.write(stamp)
    if verbose:
        for k in sorted(logs):
            print(f'==== {k}\n{logs[k]}')
    return LIB_PATH


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(path)
