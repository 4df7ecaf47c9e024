"""Recipe for oracle/_ref/: the reference's OWN native kernel for this path, compiled unmodified.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The only compiled code on the reference's hot path is the
two-file torch extension `RAFT/alt_cuda_corr/{correlation.cpp, correlation_kernel.cu}` (pybind module
`alt_cuda_corr`, `forward`/`backward`; correlation.cpp:51-54).  This script compiles those two files FROM WHERE
THEY LIE under /root/reference (nothing is copied into the repository) with torch's cpp_extension for sm_100,
and writes only into oracle/_ref/ (git-ignored, NOT gpurun-ignored, so the built module travels to the GPU
box).  The reference's own setup.py is not run (it passes no arch flags, setup.py:8-10).

    python oracle/build_ref.py            # ~3 min; no-op when oracle/_ref/alt_cuda_corr_ref*.so exists

It runs only where /root/reference exists (the authoring container); on the GPU box the prebuilt module is
used by tests/test_gpu_ref_kernel.py and tools/gpu_diag.py as the reference kernel K1 to check
`sdof_alt_corr_forward` against and to time beside it.  It is a CUDA kernel: it cannot serve as a CPU baseline.
"""
from __future__ import annotations

import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/RAFT/alt_cuda_corr'
OUT = os.path.join(HERE, '_ref')
NAME = 'alt_cuda_corr_ref'


def built_module_path():
    hits = glob.glob(os.path.join(OUT, NAME + '*.so'))
    return hits[0] if hits else None


def build(force: bool = False):
    """Returns the path of the built extension module, or None when the reference sources are absent."""
    have = built_module_path()
    if have and not force:
        return have
    srcs = [os.path.join(REF_SRC, 'correlation.cpp'), os.path.join(REF_SRC, 'correlation_kernel.cu')]
    if not all(os.path.exists(s) for s in srcs):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0')
    os.environ.setdefault('MAX_JOBS', '4')
    from torch.utils import cpp_extension
    cpp_extension.load(name=NAME, sources=srcs, extra_cuda_cflags=['-O3'], build_directory=OUT, verbose=False,
                       is_python_module=False)
    return built_module_path()


def load():
    """Import the prebuilt reference op (GPU box).  Raises FileNotFoundError if it was never built."""
    path = built_module_path()
    if path is None:
        raise FileNotFoundError('oracle/_ref/alt_cuda_corr_ref*.so not built: run `python oracle/build_ref.py` in the authoring container')
    import importlib.util
    import torch  # noqa: F401  (the module links against libtorch)
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    p = build(force='--force' in sys.argv)
    print(p if p else 'reference sources not present: nothing built')
