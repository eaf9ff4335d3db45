"""Build libegt_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m egt_b200.build            # rebuild if any source is newer than the library
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libegt_b200.so')
SOURCES = ['abi.cu', 'attn_staged.cu', 'attn_fast.cu', 'edge_kernels.cu', 'edge_fast.cu', 'node_kernels.cu', 'umma_probe.cu', 'mma_timing.cu', 'fused_prep.cu', 'fused_fwd.cu',
           'fused_bwd.cu', 'wide_prep.cu', 'wide_fwd.cu', 'wide_bwd.cu', 'node_blas.cu', 'node_tc.cu', 'peer_allreduce.cu', 'ffn_kernels.cu', 'ffn_tc.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def _deps():
    out = [os.path.join(HERE, '..', 'include', 'egt_b200.h')]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ and link egt_b200/lib/libegt_b200.so."""
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, 'obj')
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src):
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(os.path.join(objdir, src + '.ptxas.log'), 'w') as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{log}')
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-lcublas']   # static cudart; libcuda is resolved lazily at run time;
    # cuBLAS (node_blas.cu: plain node-channel GEMMs) resolves to the copy already loaded by the process, if any
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    return LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
