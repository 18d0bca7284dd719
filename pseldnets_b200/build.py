"""Build libseldfeat.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m pseldnets_b200.build          # rebuilds pseldnets_b200/libseldfeat.so

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libseldfeat.so')
SOURCES = ['seld_foa.cu', 'seld_foa_iv2.cu', 'seld_mic.cu', 'seld_epilogue.cu', 'seld_augment.cu', 'seld_abi.cu']
# SELD_EXPERIMENTS=1 python -m pseldnets_b200.build  adds the kernels that lost their A/B runs (tensor-core mel
# projection iv5, two-warps-per-frame iv3, 12-warp iv2), selectable with SELD_IV_KERNEL=5 / 3 / SELD_IV2_WARPS=12;
# the product library ships without them.
EXPERIMENT_SOURCES = ['seld_foa_iv5.cu', 'seld_foa_iv3.cu']
# -rdc=true + cudadevrt: the FOA / MIC kernels launch their own redo form from the device (cudaStreamTailLaunch) for the rare
# frames whose channels are too unbalanced for the packed transform -- ordinary input then costs no second launch
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo', '-rdc=true',
              '-Xcompiler', '-fPIC', '-shared']
NVCC_LIBS = ['-lcudadevrt']


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), 'include')):
        for fn in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, fn)))
    return m


def build(force=False, verbose=False):
    """Compile if the library is missing or older than any source.  Returns the .so path."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    experiments = os.environ.get('SELD_EXPERIMENTS', '') not in ('', '0')
    sources = SOURCES + (EXPERIMENT_SOURCES if experiments else [])
    cmd = [nvcc] + NVCC_FLAGS + (['-DSELD_EXPERIMENTS'] if experiments else []) + (['-Xptxas', '-v'] if verbose else []) + \
          ['-o', LIB] + [os.path.join(CSRC, s) for s in sources] + NVCC_LIBS
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose='-v' in sys.argv))
